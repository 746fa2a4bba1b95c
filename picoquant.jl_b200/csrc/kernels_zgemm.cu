// K2: ComplexF64 GEMM on the FP64 tensor pipe (DMMA), the second half of TTGT for the
// c128 backend.  tcgen05 has no f64 kind, so the legacy mma.sync path is the only
// tensor-core route for double precision on sm_100a (SASS: DMMA.8x8x4).
//
// Canonical TTGT layouts (both operands "open index fastest", produced by K1):
//     A[m + M*k]   B[n + N*k]   C[m + M*n]        (interleaved re/im, double2)
// so C(M,N) = A(M,K) * B(N,K)^T and C is written directly in the reference's result
// order (A-open ++ B-open, column-major) with no output permute.
//
// CTA tile 64x64x16, 8 warps (2 along m x 4 along n, warp tile 32x16), 3-stage
// cp.async pipeline.  Shared tiles are [k][m] with a pitch of 66 double2: the eight
// lanes of a quarter-warp read 4 k-rows x 2 consecutive m, i.e. byte offsets
// {0,32,64,96} + {0,16} mod 128 -> conflict-free LDS.128.  A complex product is four
// real DMMAs (ar*br, -ai*bi -> re; ar*bi, ai*br -> im).
// Roofline: 8*M*N*K real flops against the FP64 pipe; operands are re-read from L2.
#include "common.h"

namespace pq {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, STAGES = 3, PITCH = BM + 2;
constexpr int STAGE_ELEMS = BK * PITCH;                       // per operand
constexpr size_t SMEM_BYTES = size_t(STAGES) * 2 * STAGE_ELEMS * sizeof(double2);

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 2)
k_zgemm_dmma(const double2* __restrict__ A, const double2* __restrict__ B, double2* __restrict__ C,
             long long M, long long N, long long K) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* sA = reinterpret_cast<double2*>(smem_raw);
  double2* sB = sA + STAGES * STAGE_ELEMS;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 1, wn = warp >> 1;  // 2 x 4 warps
  const long long m0 = (long long)blockIdx.x * BM, n0 = (long long)blockIdx.y * BN;
  const long long KT = (K + BK - 1) / BK;

  auto load_stage = [&](int stage, long long kt) {
    double2* dA = sA + stage * STAGE_ELEMS;
    double2* dB = sB + stage * STAGE_ELEMS;
    const long long k0 = kt * BK;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int idx = tid + q * 256;  // 0..1023
      int kk = idx >> 6, mm = idx & 63;
      long long k = k0 + kk;
      bool va = (k < K) && (m0 + mm < M);
      bool vb = (k < K) && (n0 + mm < N);
      const double2* ga = va ? (A + (m0 + mm) + M * k) : A;
      const double2* gb = vb ? (B + (n0 + mm) + N * k) : B;
      cp_async16(dA + kk * PITCH + mm, ga, va);
      cp_async16(dB + kk * PITCH + mm, gb, vb);
    }
  };

  double cr[4][2][2], ci[4][2][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      cr[i][j][0] = cr[i][j][1] = 0.0;
      ci[i][j][0] = ci[i][j][1] = 0.0;
    }

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) load_stage(s, s);
    cp_async_commit();
  }

  const int frow = lane >> 2, fk = lane & 3;
  for (long long kt = 0; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      long long nk = kt + STAGES - 1;
      if (nk < KT) load_stage((int)(nk % STAGES), nk);
      cp_async_commit();
    }
    const double2* tA = sA + (int)(kt % STAGES) * STAGE_ELEMS + wm * 32 + frow;
    const double2* tB = sB + (int)(kt % STAGES) * STAGE_ELEMS + wn * 16 + frow;
#pragma unroll
    for (int ks = 0; ks < BK; ks += 4) {
      double2 a[4], b[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = tA[(ks + fk) * PITCH + i * 8];
#pragma unroll
      for (int j = 0; j < 2; ++j) b[j] = tB[(ks + fk) * PITCH + j * 8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double nai = -a[i].y;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          dmma(cr[i][j][0], cr[i][j][1], a[i].x, b[j].x);
          dmma(cr[i][j][0], cr[i][j][1], nai, b[j].y);
          dmma(ci[i][j][0], ci[i][j][1], a[i].x, b[j].y);
          dmma(ci[i][j][0], ci[i][j][1], a[i].y, b[j].x);
        }
      }
    }
  }
  cp_async_wait<0>();

  // epilogue: lane holds C[row = lane>>2][col = 2*(lane&3) + c] of every 8x8 sub-tile
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      long long n = n0 + wn * 16 + j * 8 + 2 * fk + c;
      if (n >= N) continue;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        long long m = m0 + wm * 32 + i * 8 + frow;
        if (m < M) C[m + M * n] = make_double2(cr[i][j][c], ci[i][j][c]);
      }
    }
}

// FP64 issue-rate probes for the roofline denominators (dependent chains per warp are
// kept short enough to saturate the pipe with 8 warps x 4 independent accumulators).
__global__ void __launch_bounds__(256) k_probe_dmma(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_probe_dfma(double* out, int iters) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = i;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

// per-device one-time kernel attributes (called from pq_create, outside any capture)
void init_kernels() {
  PQ_CUDA(cudaFuncSetAttribute(k_zgemm_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)SMEM_BYTES));
}

void run_zgemm_dmma(const Launch& L, const void* A, const void* B, void* C, int64_t M, int64_t N,
                    int64_t K) {
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN));
  PQ_REQUIRE(grid.y <= 65535, PQ_ERR_UNSUPPORTED, "N too large for the ZGEMM grid");
  double bytes = double(M * K + N * K + M * N) * 16.0, flops = 8.0 * M * N * K;
  L.begin(KC_GEMM_TENSOR, bytes, flops);
  k_zgemm_dmma<<<grid, 256, SMEM_BYTES, L.stream>>>((const double2*)A, (const double2*)B,
                                                    (double2*)C, M, N, K);
  L.end();
  PQ_CUDA(cudaGetLastError());
}

// returns achieved TFLOP/s of the probe ("dmma" or "dfma")
double run_fp64_probe(const Launch& L, bool tensor) {
  const int blocks = L.num_sms * 8, iters = 4096;
  double* out = nullptr;
  PQ_CUDA(cudaMalloc(&out, size_t(blocks) * 256 * sizeof(double)));
  cudaEvent_t e0, e1;
  PQ_CUDA(cudaEventCreate(&e0));
  PQ_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    PQ_CUDA(cudaEventRecord(e0, L.stream));
    if (tensor)
      k_probe_dmma<<<blocks, 256, 0, L.stream>>>(out, iters);
    else
      k_probe_dfma<<<blocks, 256, 0, L.stream>>>(out, iters);
    PQ_CUDA(cudaEventRecord(e1, L.stream));
    PQ_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    PQ_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  PQ_CUDA(cudaGetLastError());
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  double flops;
  if (tensor)  // per warp-instruction 8*8*4 FMAs
    flops = double(blocks) * 8.0 * iters * 8.0 * (8 * 8 * 4 * 2);
  else
    flops = double(blocks) * 256.0 * iters * 16.0 * 2.0;
  return flops / (best * 1e-3) / 1e12;
}

}  // namespace pq
