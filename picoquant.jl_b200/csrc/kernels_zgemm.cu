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
#include <cstdlib>

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
      // 16 independent accumulators first, their dependent second products afterwards:
      // back-to-back dependent DMMAs would stall on the pipe latency ("wait" stalls)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          dmma(cr[i][j][0], cr[i][j][1], a[i].x, b[j].x);
          dmma(ci[i][j][0], ci[i][j][1], a[i].x, b[j].y);
        }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double nai = -a[i].y;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          dmma(cr[i][j][0], cr[i][j][1], nai, b[j].y);
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

// ---------------------------------------------------------------------------
// Fused TTGT: the same tile pipeline, but the operand tiles are gathered straight from
// the un-permuted tensors -- the "transpose" of TTGT happens in the cp.async that fills
// shared memory, so no permuted copy of A or B is ever written to HBM.  An element of
// the A tile is A[row_off(m) + k_off(k)]: row offsets (64 per tile) and k offsets (K,
// 32-bit) are tabulated in shared memory from the fused index maps.  Because the 6
// lowest open bits of A are its 6 lowest-address open bits, a k-row of the tile is made
// of long contiguous runs and the gather stays sector-efficient.
// ---------------------------------------------------------------------------
struct FusedParams {
  IdxMap mA, kA, nB, kB;
  long long M, N, K;
};

// Persistent: each CTA walks tiles t = blockIdx.x, +gridDim.x, ... (m-tiles fastest) and the
// cp.async ring runs ahead ACROSS tile boundaries, so the fill of tile i+1 overlaps the
// last k-blocks and the epilogue of tile i (the skinny K=64 steps of an RQC sweep only have
// 4 k-blocks per tile; without this the pipeline would drain on every tile).
constexpr int ROW_RING = 4;

__global__ void __launch_bounds__(256, 2)
k_zgemm_fused(const double2* __restrict__ A, const double2* __restrict__ B, double2* __restrict__ C,
              const FusedParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* sA = reinterpret_cast<double2*>(smem_raw);
  double2* sB = sA + STAGES * STAGE_ELEMS;
  long long* rowA = reinterpret_cast<long long*>(sB + STAGES * STAGE_ELEMS);  // [ROW_RING][BM]
  long long* rowB = rowA + ROW_RING * BM;                                      // [ROW_RING][BN]
  int* koffA = reinterpret_cast<int*>(rowB + ROW_RING * BN);                   // [K]
  int* koffB = koffA + p.K;                                                    // [K]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 1, wn = warp >> 1;
  const long long M = p.M, N = p.N;
  const int K = (int)p.K;
  const long long tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  const long long ntiles = tiles_m * tiles_n;
  const int KT = (K + BK - 1) / BK;
  if ((long long)blockIdx.x >= ntiles) return;
  const int my_count = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

  auto tile_origin = [&](int i, long long& m0, long long& n0) {
    long long t = (long long)blockIdx.x + (long long)i * gridDim.x;
    long long tn = t / tiles_m, tm = t - tn * tiles_m;
    m0 = tm * BM;
    n0 = tn * BN;
  };
  auto compute_rows = [&](int i) {  // tables of tile i into ring slot i % ROW_RING
    if (tid >= BM + BN) return;
    long long m0, n0;
    tile_origin(i, m0, n0);
    const int slot = i & (ROW_RING - 1);
    if (tid < BM)
      rowA[slot * BM + tid] = (m0 + tid < M) ? map_offset(p.mA, m0 + tid) : -1;
    else
      rowB[slot * BN + (tid - BM)] =
          (n0 + (tid - BM) < N) ? map_offset(p.nB, n0 + (tid - BM)) : -1;
  };

  // ---- prefetch stream: (pf_i, pf_kt) walks this CTA's flattened (tile, k-block) list ----
  const int mm = tid & 63, kk0 = tid >> 6;          // this thread's column and first k-row
  int pf_i = 0, pf_kt = 0, pf_ring = 0;
  long long pf_ra = -1, pf_rb = -1;
  auto prefetch = [&]() {
    if (pf_i < my_count) {
      if (pf_kt == 0) {
        const int slot = pf_i & (ROW_RING - 1);
        pf_ra = rowA[slot * BM + mm];
        pf_rb = rowB[slot * BN + mm];
      }
      double2* dA = sA + pf_ring * STAGE_ELEMS + kk0 * PITCH + mm;
      double2* dB = sB + pf_ring * STAGE_ELEMS + kk0 * PITCH + mm;
      const int kbase = pf_kt * BK + kk0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = kbase + 4 * q;
        const bool kin = k < K;
        const bool va = kin && (pf_ra >= 0), vb = kin && (pf_rb >= 0);
        const int ka = kin ? koffA[k] : 0, kb = kin ? koffB[k] : 0;
        cp_async16(dA + 4 * q * PITCH, va ? (A + pf_ra + ka) : A, va);
        cp_async16(dB + 4 * q * PITCH, vb ? (B + pf_rb + kb) : B, vb);
      }
      if (++pf_kt == KT) {
        pf_kt = 0;
        ++pf_i;
      }
      pf_ring = (pf_ring + 1 == STAGES) ? 0 : pf_ring + 1;
    }
    cp_async_commit();
  };

  for (int k = tid; k < K; k += 256) {
    koffA[k] = (int)map_offset(p.kA, k);
    koffB[k] = (int)map_offset(p.kB, k);
  }
  for (int i = 0; i < ROW_RING - 1 && i < my_count; ++i) compute_rows(i);
  __syncthreads();

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) prefetch();

  double cr[4][2][2], ci[4][2][2];
  const int frow = lane >> 2, fk = lane & 3;
  int ring = 0;
  for (int i = 0; i < my_count; ++i) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        cr[a][b][0] = cr[a][b][1] = 0.0;
        ci[a][b][0] = ci[a][b][1] = 0.0;
      }
    for (int kt = 0; kt < KT; ++kt) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      // ring slot of tile i-1 is free: all of its stages were issued before tile i began
      if (kt == 0 && i + ROW_RING - 1 < my_count) compute_rows(i + ROW_RING - 1);
      prefetch();
      const double2* tA = sA + ring * STAGE_ELEMS + wm * 32 + frow;
      const double2* tB = sB + ring * STAGE_ELEMS + wn * 16 + frow;
#pragma unroll
      for (int ks = 0; ks < BK; ks += 4) {
        double2 a[4], b[2];
#pragma unroll
        for (int x = 0; x < 4; ++x) a[x] = tA[(ks + fk) * PITCH + x * 8];
#pragma unroll
        for (int y = 0; y < 2; ++y) b[y] = tB[(ks + fk) * PITCH + y * 8];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
          for (int y = 0; y < 2; ++y) {
            dmma(cr[x][y][0], cr[x][y][1], a[x].x, b[y].x);
            dmma(ci[x][y][0], ci[x][y][1], a[x].x, b[y].y);
          }
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const double nai = -a[x].y;
#pragma unroll
          for (int y = 0; y < 2; ++y) {
            dmma(cr[x][y][0], cr[x][y][1], nai, b[y].y);
            dmma(ci[x][y][0], ci[x][y][1], a[x].y, b[y].x);
          }
        }
      }
      ring = (ring + 1 == STAGES) ? 0 : ring + 1;
    }
    // epilogue of tile i; the loads of the next tile are already in flight
    long long m0, n0;
    tile_origin(i, m0, n0);
#pragma unroll
    for (int y = 0; y < 2; ++y)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        long long n = n0 + wn * 16 + y * 8 + 2 * fk + c;
        if (n >= N) continue;
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          long long m = m0 + wm * 32 + x * 8 + frow;
          if (m < M) C[m + M * n] = make_double2(cr[x][y][c], ci[x][y][c]);
        }
      }
  }
  cp_async_wait<0>();
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
constexpr int FUSED_MAX_K = 1024;
constexpr size_t FUSED_SMEM_MAX =
    SMEM_BYTES + size_t(ROW_RING) * (BM + BN) * 8 + size_t(FUSED_MAX_K) * 8;

void init_kernels() {
  PQ_CUDA(cudaFuncSetAttribute(k_zgemm_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)SMEM_BYTES));
  PQ_CUDA(cudaFuncSetAttribute(k_zgemm_fused, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)FUSED_SMEM_MAX));
}

// fused TTGT ZGEMM straight from the un-permuted operands (K <= 1024, 32-bit k offsets)
void run_zgemm_fused(const Launch& L, const ContractPlan& cp, const void* A, const void* B,
                     void* C) {
  FusedParams fp;
  fp.mA = cp.mA;
  fp.kA = cp.kA;
  fp.nB = cp.nB;
  fp.kB = cp.kB;
  fp.M = cp.M;
  fp.N = cp.N;
  fp.K = cp.K;
  long long ntiles = ((cp.M + BM - 1) / BM) * ((cp.N + BN - 1) / BN);
  // One tile per CTA by default: on B200 the hardware CTA scheduler (2 resident CTAs per SM,
  // naturally de-phased) beat a persistent 2-CTA/SM grid by ~7% on the skinny sweep steps
  // (measured, profiles/); PQ_ZGEMM_CTAS_PER_SM=k caps the grid at k*SMs persistent CTAs.
  static const long long per_sm =
      getenv("PQ_ZGEMM_CTAS_PER_SM") ? atoll(getenv("PQ_ZGEMM_CTAS_PER_SM")) : (1LL << 40);
  long long cap = per_sm <= 0 ? (1LL << 40) : per_sm * L.num_sms;
  if (cap > 0x7fffffffLL) cap = 0x7fffffffLL;
  unsigned grid = (unsigned)(ntiles < cap ? ntiles : cap);
  PQ_REQUIRE(cp.K <= FUSED_MAX_K, PQ_ERR_INVALID, "fused ZGEMM: K too large");
  size_t smem = SMEM_BYTES + size_t(ROW_RING) * (BM + BN) * 8 + size_t(cp.K) * 8;
  double bytes = double(cp.M * cp.K + cp.N * cp.K + cp.M * cp.N) * 16.0;
  double flops = 8.0 * double(cp.M) * double(cp.N) * double(cp.K);
  L.begin(KC_GEMM_TENSOR, bytes, flops);
  k_zgemm_fused<<<grid, 256, smem, L.stream>>>((const double2*)A, (const double2*)B, (double2*)C,
                                               fp);
  L.end();
  PQ_CUDA(cudaGetLastError());
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
