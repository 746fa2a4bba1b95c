// decompose_tensor! on the device (reference: src/layer1.jl:146-184, called through
// src/backends/interactive.jl:130-152): permute the tensor to [left | right], view it as an
// m x n matrix, take its SVD, keep the chi singular values above the relative threshold
// (at most max_rank) and return  B = U_chi * sqrt(S_chi),  C = sqrt(S_chi) * V_chi^H.
//
// The SVD is a one-sided (Hestenes) Jacobi iteration written for this use: the matrices are
// small or moderate (a 4x4 gate, a (2 chi) x (2 chi) bond between MPS sites), so the work is
// latency- not throughput-bound, and Jacobi needs nothing but column dot products and plane
// rotations -- one WARP per column pair, n/2 disjoint pairs per step of a round-robin
// tournament, n-1 steps per sweep.  It also delivers singular values to high relative
// accuracy, which matters here because chi is decided by thresholding them.
//
//   * n_pad/2 <= 32 pairs: ONE CTA of up to 32 warps runs all sweeps with __syncthreads()
//     between tournament steps (k_jacobi_cta) -- a single launch for a gate or a small bond.
//   * larger: one launch per tournament step (k_jacobi_step); the host reads the rotation
//     counter once per sweep.  decompose_tensor! returns chi to the host, so the call is
//     synchronous by contract anyway.
//
// Column pair (p, q) with Gram entries alpha = |a_p|^2, beta = |a_q|^2, gamma = a_p^H a_q:
// with w = gamma/|gamma| and b = a_q * conj(w) the Gram matrix of (a_p, b) is real,
// [[alpha, |gamma|], [|gamma|, beta]], and the real Jacobi rotation with
// zeta = (beta - alpha) / (2 |gamma|), t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)),
// c = 1/sqrt(1 + t^2), s = c t makes the new columns  a_p' = c a_p - s b,  a_q' = s a_p + c b
// orthogonal.  The same transformation is applied to the columns of V (initially I), so
// A_0 = A' V^H holds throughout; at convergence A' = U S.
#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>

#include "handle.h"

namespace pq {

namespace {

template <typename R> struct C2;
template <> struct C2<float> { using type = float2; };
template <> struct C2<double> { using type = double2; };

// round-robin tournament on n (even) players: step in [0, n-1), k in [0, n/2)
__device__ __forceinline__ void pair_of(int n, int step, int k, int& p, int& q) {
  int a, b;
  if (k == 0) {
    a = n - 1;
    b = step;
  } else {
    a = (step + k) % (n - 1);
    b = (step - k + (n - 1)) % (n - 1);
  }
  p = a < b ? a : b;
  q = a < b ? b : a;
}

template <typename R>
__device__ __forceinline__ R warp_sum(R v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// one warp orthogonalises columns p and q of A (m rows) and applies the same rotation to V
template <typename R>
__device__ __forceinline__ bool rotate_pair(typename C2<R>::type* __restrict__ A, long long m,
                                            typename C2<R>::type* __restrict__ Vm, int n, int p,
                                            int q, R tol2, int lane) {
  using V = typename C2<R>::type;
  V* ap = A + (long long)p * m;
  V* aq = A + (long long)q * m;
  R alpha = 0, beta = 0, gr = 0, gi = 0;
  for (long long i = lane; i < m; i += 32) {
    const V x = ap[i], y = aq[i];
    alpha += x.x * x.x + x.y * x.y;
    beta += y.x * y.x + y.y * y.y;
    gr += x.x * y.x + x.y * y.y;   // conj(x) * y
    gi += x.x * y.y - x.y * y.x;
  }
  alpha = warp_sum(alpha);
  beta = warp_sum(beta);
  gr = warp_sum(gr);
  gi = warp_sum(gi);
  const R g2 = gr * gr + gi * gi;
  if (!(g2 > tol2 * alpha * beta)) return false;   // already orthogonal (or a zero column)
  const R g = sqrt(g2);
  const R wr = gr / g, wi = gi / g;
  const R zeta = (beta - alpha) / (2 * g);
  const R t = copysign(R(1), zeta) / (fabs(zeta) + sqrt(R(1) + zeta * zeta));
  const R c = R(1) / sqrt(R(1) + t * t), s = c * t;
  for (long long i = lane; i < m; i += 32) {
    const V x = ap[i], y = aq[i];
    V b;
    b.x = y.x * wr + y.y * wi;   // y * conj(w)
    b.y = y.y * wr - y.x * wi;
    V xo, yo;
    xo.x = c * x.x - s * b.x;
    xo.y = c * x.y - s * b.y;
    yo.x = s * x.x + c * b.x;
    yo.y = s * x.y + c * b.y;
    ap[i] = xo;
    aq[i] = yo;
  }
  V* vp = Vm + (long long)p * n;
  V* vq = Vm + (long long)q * n;
  for (int i = lane; i < n; i += 32) {
    const V x = vp[i], y = vq[i];
    V b;
    b.x = y.x * wr + y.y * wi;
    b.y = y.y * wr - y.x * wi;
    V xo, yo;
    xo.x = c * x.x - s * b.x;
    xo.y = c * x.y - s * b.y;
    yo.x = s * x.x + c * b.x;
    yo.y = s * x.y + c * b.y;
    vp[i] = xo;
    vq[i] = yo;
  }
  return true;
}

// all sweeps in one CTA: warp w owns pair w of every tournament step
template <typename R>
__global__ void __launch_bounds__(1024)
k_jacobi_cta(typename C2<R>::type* __restrict__ A, long long m, typename C2<R>::type* __restrict__ Vm,
             int n, int npad, R tol2, int max_sweeps, int* __restrict__ info) {
  __shared__ int rotations;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    if (threadIdx.x == 0) rotations = 0;
    __syncthreads();
    for (int step = 0; step < npad - 1; ++step) {
      int p, q;
      pair_of(npad, step, warp, p, q);
      if (q < n && rotate_pair<R>(A, m, Vm, n, p, q, tol2, lane) && lane == 0)
        atomicAdd(&rotations, 1);
      __syncthreads();
    }
    const int done = rotations == 0;
    __syncthreads();
    if (done) break;
  }
  if (threadIdx.x == 0) {
    info[0] = sweep;                       // sweeps used
    info[1] = sweep < max_sweeps ? 1 : 0;  // converged
  }
}

// one tournament step over the whole grid, 8 warps (pairs) per CTA
template <typename R>
__global__ void __launch_bounds__(256)
k_jacobi_step(typename C2<R>::type* __restrict__ A, long long m, typename C2<R>::type* __restrict__ Vm,
              int n, int npad, int step, R tol2, int* __restrict__ rotations) {
  const int k = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= npad / 2) return;
  int p, q;
  pair_of(npad, step, k, p, q);
  if (q >= n) return;   // the padding column of an odd n
  if (rotate_pair<R>(A, m, Vm, n, p, q, tol2, lane) && lane == 0) atomicAdd(rotations, 1);
}

template <typename R>
__global__ void k_identity(typename C2<R>::type* __restrict__ Vm, int n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)n * n) return;
  typename C2<R>::type v;
  v.x = (i / n == i % n) ? R(1) : R(0);
  v.y = 0;
  Vm[i] = v;
}

// out (cols x rows) = in (rows x cols)^H
template <typename R>
__global__ void k_adjoint(const typename C2<R>::type* __restrict__ in,
                          typename C2<R>::type* __restrict__ out, long long rows, long long cols) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const long long c = i / rows, r = i - c * rows;
  typename C2<R>::type v = in[i];
  v.y = -v.y;
  out[c + cols * r] = v;
}

// sigma[j] = |a_j| (one warp per column), in double for the host-side thresholding
template <typename R>
__global__ void __launch_bounds__(256)
k_col_norms(const typename C2<R>::type* __restrict__ A, long long m, int n, double* __restrict__ sigma) {
  const int j = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= n) return;
  double s = 0;
  for (long long i = lane; i < m; i += 32) {
    const typename C2<R>::type x = A[(long long)j * m + i];
    s += (double)x.x * x.x + (double)x.y * x.y;
  }
  s = warp_sum(s);
  if (lane == 0) sigma[j] = sqrt(s);
}

// dst[i + rows * j] = src[i + rows * perm[j]] * scale[j]            (conj_t == 0)
// dst[j + chi * i]  = conj(src[i + rows * perm[j]]) * scale[j]      (conj_t == 1)
template <typename R>
__global__ void k_factor(const typename C2<R>::type* __restrict__ src, long long rows,
                         const int* __restrict__ perm, const double* __restrict__ scale, int chi,
                         int conj_t, typename C2<R>::type* __restrict__ dst) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= rows * chi) return;
  long long i, j;
  if (conj_t) {
    j = e % chi;
    i = e / chi;
  } else {
    i = e % rows;
    j = e / rows;
  }
  typename C2<R>::type v = src[i + rows * perm[j]];
  const R sc = (R)scale[j];
  v.x *= sc;
  v.y *= conj_t ? -sc : sc;
  dst[e] = v;
}

template <typename R>
int decompose_impl(pq_handle* h, const Launch& L, void* work, int64_t m, int64_t n, double threshold,
                   int max_rank, std::shared_ptr<Buffer>& Bout, std::shared_ptr<Buffer>& Cout) {
  using V = typename C2<R>::type;
  cudaStream_t st = L.stream;
  // Jacobi runs on the columns of a tall matrix: for m < n decompose A^H = U' S V'^H instead,
  // then A = V' S U'^H
  const bool transposed = m < n;
  const int64_t rows = transposed ? n : m;   // of the matrix Jacobi works on
  const int64_t cols = transposed ? m : n;
  PQ_REQUIRE(cols <= 16384, PQ_ERR_UNSUPPORTED, "decompose: more than 16384 columns");
  std::shared_ptr<Buffer> At;
  V* A = (V*)work;
  if (transposed) {
    At = std::make_shared<Buffer>(size_t(rows * cols) * sizeof(V), st);
    L.begin(KC_SVD, 2.0 * rows * cols * sizeof(V), 0);
    k_adjoint<R><<<(unsigned)((rows * cols + 255) / 256), 256, 0, st>>>((const V*)work, (V*)At->ptr, m, n);
    L.end();
    A = (V*)At->ptr;
  }
  const int nc = (int)cols;
  const int npad = nc + (nc & 1);
  Buffer Vm(size_t(nc) * nc * sizeof(V), st);
  Buffer sig(size_t(nc) * sizeof(double) + 16, st);
  Buffer info(4 * sizeof(int), st);
  L.begin(KC_SVD, 0, 0);
  k_identity<R><<<(unsigned)(((int64_t)nc * nc + 255) / 256), 256, 0, st>>>((V*)Vm.ptr, nc);
  L.end();
  const R eps = std::numeric_limits<R>::epsilon();
  // |gamma| is computed with a rounding error of about sqrt(rows) eps sqrt(alpha beta)
  const R tol = eps * std::max(R(4), (R)std::sqrt((double)rows));
  const R tol2 = tol * tol;
  const int max_sweeps = 40;
  int sweeps_used = 0;
  if (nc >= 2) {
    if (npad / 2 <= 32) {
      L.begin(KC_SVD, 0, 0);
      k_jacobi_cta<R><<<1, 32 * (npad / 2), 0, st>>>(A, rows, (V*)Vm.ptr, nc, npad, tol2, max_sweeps,
                                                    (int*)info.ptr);
      L.end();
      int hinfo[2] = {0, 0};
      PQ_CUDA(cudaMemcpyAsync(hinfo, info.ptr, sizeof(hinfo), cudaMemcpyDeviceToHost, st));
      PQ_CUDA(cudaStreamSynchronize(st));
      sweeps_used = hinfo[0];
    } else {
      const unsigned grid = (unsigned)((npad / 2 + 7) / 8);
      for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        PQ_CUDA(cudaMemsetAsync(info.ptr, 0, sizeof(int), st));
        for (int step = 0; step < npad - 1; ++step) {
          L.begin(KC_SVD, 0, 0);
          k_jacobi_step<R><<<grid, 256, 0, st>>>(A, rows, (V*)Vm.ptr, nc, npad, step, tol2,
                                                 (int*)info.ptr);
          L.end();
        }
        int rot = 0;
        PQ_CUDA(cudaMemcpyAsync(&rot, info.ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
        PQ_CUDA(cudaStreamSynchronize(st));
        sweeps_used = sweep + 1;
        if (rot == 0) break;
      }
    }
  }
  (void)sweeps_used;
  // singular values -> host: order, threshold, chi (layer1.jl:166-176)
  L.begin(KC_SVD, 0, 0);
  k_col_norms<R><<<(unsigned)((nc * 32LL + 255) / 256), 256, 0, st>>>(A, rows, nc, (double*)sig.ptr);
  L.end();
  std::vector<double> S(nc);
  PQ_CUDA(cudaMemcpyAsync(S.data(), sig.ptr, sizeof(double) * nc, cudaMemcpyDeviceToHost, st));
  PQ_CUDA(cudaStreamSynchronize(st));
  std::vector<int> order(nc);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return S[a] > S[b]; });
  double thr = std::max(threshold, std::sqrt((double)eps));   // at least sqrt(eps(real(T)))
  double s_norm2 = 0;
  for (double s : S) s_norm2 += s * s;
  const double s_norm = std::sqrt(s_norm2);
  int chi = 0;
  for (int j = 0; j < nc; ++j)
    if (S[order[j]] / s_norm > thr) ++chi;   // NaN (all-zero tensor) compares false, as in Julia
  if (max_rank > 0) chi = std::min(max_rank, chi);
  // factors
  Bout = std::make_shared<Buffer>(size_t(m) * chi * sizeof(V), st);
  Cout = std::make_shared<Buffer>(size_t(chi) * n * sizeof(V), st);
  if (chi > 0) {
    std::vector<double> up(chi), down(chi);
    for (int j = 0; j < chi; ++j) {
      up[j] = std::sqrt(S[order[j]]);
      down[j] = 1.0 / up[j];
    }
    Buffer dperm(sizeof(int) * chi, st), dup(sizeof(double) * chi, st), ddown(sizeof(double) * chi, st);
    PQ_CUDA(cudaMemcpyAsync(dperm.ptr, order.data(), sizeof(int) * chi, cudaMemcpyHostToDevice, st));
    PQ_CUDA(cudaMemcpyAsync(dup.ptr, up.data(), sizeof(double) * chi, cudaMemcpyHostToDevice, st));
    PQ_CUDA(cudaMemcpyAsync(ddown.ptr, down.data(), sizeof(double) * chi, cudaMemcpyHostToDevice, st));
    const unsigned gB = (unsigned)((m * chi + 255) / 256), gC = (unsigned)((n * chi + 255) / 256);
    L.begin(KC_SVD, 0, 0);
    if (!transposed) {
      // A' = U S (m x n), V (n x n):  B = a_j / sqrt(s_j),  C[j, k] = sqrt(s_j) conj(V[k, j])
      k_factor<R><<<gB, 256, 0, st>>>(A, m, (const int*)dperm.ptr, (const double*)ddown.ptr, chi, 0,
                                     (V*)Bout->ptr);
      k_factor<R><<<gC, 256, 0, st>>>((const V*)Vm.ptr, n, (const int*)dperm.ptr,
                                     (const double*)dup.ptr, chi, 1, (V*)Cout->ptr);
    } else {
      // At' = U' S (n x m), V' (m x m):  B = V'_j sqrt(s_j),  C[j, k] = conj(At'[k, j]) / sqrt(s_j)
      k_factor<R><<<gB, 256, 0, st>>>((const V*)Vm.ptr, m, (const int*)dperm.ptr,
                                     (const double*)dup.ptr, chi, 0, (V*)Bout->ptr);
      k_factor<R><<<gC, 256, 0, st>>>(A, n, (const int*)dperm.ptr, (const double*)ddown.ptr, chi, 1,
                                     (V*)Cout->ptr);
    }
    L.end();
    PQ_CUDA(cudaStreamSynchronize(st));   // the host vectors above are copied from asynchronously
  }
  PQ_CUDA(cudaGetLastError());
  return chi;
}

}  // namespace

int run_decompose(pq_handle* h, const Launch& L, void* work, int64_t m, int64_t n, double threshold,
                  int max_rank, std::shared_ptr<Buffer>& Bout, std::shared_ptr<Buffer>& Cout) {
  if (h->elem_size == 16)
    return decompose_impl<double>(h, L, work, m, n, threshold, max_rank, Bout, Cout);
  return decompose_impl<float>(h, L, work, m, n, threshold, max_rank, Bout, Cout);
}

}  // namespace pq
