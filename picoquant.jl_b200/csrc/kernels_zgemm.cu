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
// The same GEMM with 3M complex products (see k_zgemm_fused_t / k_zgemm_skinny): long
// contractions on canonical layouts are bound by the FP64 pipe, so issuing three DMMAs per
// complex product instead of four is worth ~25 %.  The third accumulator set does not fit the
// 128-register budget of two 256-thread CTAs per SM with 32x16 warp tiles, hence ONE CTA of
// 16 warps (4 x 4, warp tile 16x16) per SM, 64x64x32 tiles, 3 stages (203 KB).
// ---------------------------------------------------------------------------
constexpr int M3_BK = 32, M3_STAGES = 3, M3_THREADS = 512;
constexpr int M3_STAGE_ELEMS = M3_BK * PITCH;
constexpr size_t M3_SMEM_BYTES = size_t(M3_STAGES) * 2 * M3_STAGE_ELEMS * sizeof(double2);

__global__ void __launch_bounds__(M3_THREADS, 1)
k_zgemm_dmma3m(const double2* __restrict__ A, const double2* __restrict__ B, double2* __restrict__ C,
               long long M, long long N, long long K) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* sA = reinterpret_cast<double2*>(smem_raw);
  double2* sB = sA + M3_STAGES * M3_STAGE_ELEMS;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;  // 4 x 4 warps
  // n-tiles in groups of 8 per sweep over m: a wave of CTAs re-uses 8 column tiles of B and a
  // band of A out of L2 instead of streaming all of A once per column tile
  const long long tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  constexpr long long GROUP_N = 8;
  const long long bid = blockIdx.x;
  const long long group = bid / (GROUP_N * tiles_m);
  const long long gn = (tiles_n - group * GROUP_N) < GROUP_N ? (tiles_n - group * GROUP_N) : GROUP_N;
  const long long rem = bid - group * GROUP_N * tiles_m;
  const long long m0 = (rem / gn) * BM, n0 = (group * GROUP_N + rem % gn) * BN;
  const long long KT = (K + M3_BK - 1) / M3_BK;

  auto load_stage = [&](int stage, long long kt) {
    double2* dA = sA + stage * M3_STAGE_ELEMS;
    double2* dB = sB + stage * M3_STAGE_ELEMS;
    const long long k0 = kt * M3_BK;
#pragma unroll
    for (int q = 0; q < BM * M3_BK / M3_THREADS; ++q) {
      int idx = tid + q * M3_THREADS;
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

  double p1[2][2][2], p2[2][2][2], p3[2][2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      p1[i][j][0] = p1[i][j][1] = 0.0;
      p2[i][j][0] = p2[i][j][1] = 0.0;
      p3[i][j][0] = p3[i][j][1] = 0.0;
    }

#pragma unroll
  for (int s = 0; s < M3_STAGES - 1; ++s) {
    if (s < KT) load_stage(s, s);
    cp_async_commit();
  }

  const int frow = lane >> 2, fk = lane & 3;
  int ring = 0, pring = M3_STAGES - 1;
  for (long long kt = 0; kt < KT; ++kt) {
    cp_async_wait<M3_STAGES - 2>();
    __syncthreads();
    if (kt + M3_STAGES - 1 < KT) load_stage(pring, kt + M3_STAGES - 1);
    cp_async_commit();
    const double2* tA = sA + ring * M3_STAGE_ELEMS + wm * 16 + frow;
    const double2* tB = sB + ring * M3_STAGE_ELEMS + wn * 16 + frow;
#pragma unroll
    for (int ks = 0; ks < M3_BK; ks += 4) {
      double2 a[2], b[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) a[i] = tA[(ks + fk) * PITCH + i * 8];
#pragma unroll
      for (int j = 0; j < 2; ++j) b[j] = tB[(ks + fk) * PITCH + j * 8];
      double as[2], bs[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) as[i] = a[i].x + a[i].y;
#pragma unroll
      for (int j = 0; j < 2; ++j) bs[j] = b[j].x + b[j].y;
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          dmma(p1[i][j][0], p1[i][j][1], a[i].x, b[j].x);
          dmma(p2[i][j][0], p2[i][j][1], a[i].y, b[j].y);
          dmma(p3[i][j][0], p3[i][j][1], as[i], bs[j]);
        }
    }
    ring = (ring + 1 == M3_STAGES) ? 0 : ring + 1;
    pring = (pring + 1 == M3_STAGES) ? 0 : pring + 1;
  }
  cp_async_wait<0>();

#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      long long n = n0 + wn * 16 + j * 8 + 2 * fk + c;
      if (n >= N) continue;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        long long m = m0 + wm * 16 + i * 8 + frow;
        if (m < M)
          C[m + M * n] = make_double2(p1[i][j][c] - p2[i][j][c],
                                      p3[i][j][c] - p1[i][j][c] - p2[i][j][c]);
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
// (struct FusedParams: common.h)

// One output tile per CTA (a persistent variant with a cross-tile cp.async ring was
// measured ~7 % slower on B200: the hardware CTA scheduler de-phases resident CTAs better).
// Two tile configurations:
//   <64,64, 2x4 warps, 3 stages, 2 CTAs/SM>  warp tile 32x16 -- large N and K
//   <64,32, 4x2 warps, 2 stages, 4 CTAs/SM>  warp tile 16x16 -- the skinny K<=128 sweep steps:
//     64 registers and 51 KB of shared memory per CTA put four independent CTAs (32 warps)
//     on every SM, which hides the per-tile fill latency that dominates when a tile has only
//     four k-blocks (27.4 -> 30.7 TFLOP/s on M=2^18, N=K=64; cuBLAS ZGEMM on pre-permuted
//     operands: 30.7-31.7).  A Gauss-3M variant (3 DMMAs per complex product) was tried and
//     gave no gain: these steps are bound by tile fill latency, not by the DMMA pipe.
//   <128,8, 8x1 warps, 2 stages, 3 CTAs/SM>  warp tile 16x8 -- N <= 16 with K >= 32 (a site
//     tensor with one open bond of 8 absorbed into the boundary): one m8n8k4 DMMA column
//     covers all of N, so no tensor-pipe work is wasted on padding; HBM-bound.
// AKF / BKF ("k first"): gather order of the A / B tile loads.  false: consecutive threads
// walk the rows of the tile (m or n), true: they walk k first.  The host picks, per operand,
// the order in which consecutive threads touch consecutive addresses (k first when the
// fastest axis of that tensor is a contracted one); otherwise every 16-byte cp.async of a
// warp would land in a different 128-byte line.
// M3 ("3M", Karatsuba): a complex product from THREE real DMMAs instead of four,
//     P1 = Ar Br,  P2 = Ai Bi,  P3 = (Ar + Ai)(Br + Bi);   Cr = P1 - P2,  Ci = P3 - P1 - P2.
// The tensor pipe is the binding resource of the K = 32..64 sweep steps (ncu: pipe 89 % active,
// math_pipe_throttle the dominant stall), so 25 % fewer DMMAs is worth a third accumulator set
// and two DADDs per fragment.  Norm-wise the rounding error stays at a few eps |A||B|.
template <int TBM, int TBN, int WGM, int WGN, int NST, int MINB, bool AKF, bool BKF, bool M3 = false>
__global__ void __launch_bounds__(256, MINB)
k_zgemm_fused_t(const double2* __restrict__ A, const double2* __restrict__ B,
                double2* __restrict__ C, const FusedParams p) {
  constexpr int PA = TBM + 2, PB = TBN + 2;              // smem pitches (double2)
  constexpr int XT = TBM / WGM / 8, YT = TBN / WGN / 8;  // 8x8 sub-tiles per warp
  constexpr int A_ELEMS = BK * PA, B_ELEMS = BK * PB;
  constexpr int A_STEP = 256 / TBM, A_CNT = BK / A_STEP;  // k-rows per pass / passes
  // a B tile narrower than 256 / BK columns is loaded by the first TBN * BK threads only
  constexpr int B_STEP = (256 / TBN < BK) ? 256 / TBN : BK, B_CNT = BK / B_STEP;
  static_assert(WGM * WGN == 8, "eight warps");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* sA = reinterpret_cast<double2*>(smem_raw);
  double2* sB = sA + NST * A_ELEMS;
  long long* rowA = reinterpret_cast<long long*>(sB + NST * B_ELEMS);  // [TBM]
  long long* rowB = rowA + TBM;                                        // [TBN]
  int* koffA = reinterpret_cast<int*>(rowB + TBN);                     // [K]
  int* koffB = koffA + p.K;                                            // [K]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % WGM, wn = warp / WGM;
  const long long M = p.M, N = p.N;
  const int K = (int)p.K;
  const long long tiles_n = (N + TBN - 1) / TBN;
  // n-tiles fastest: CTAs that share the same rows of A are scheduled back to back (L2 reuse)
  const long long tm = (long long)blockIdx.x / tiles_n, tn = (long long)blockIdx.x - tm * tiles_n;
  const long long m0 = tm * TBM, n0 = tn * TBN;
  const int KT = (K + BK - 1) / BK;

  if (p.stagger_ns > 0 && (int)blockIdx.x < p.first_wave && tid == 0) {
    const unsigned long long wait_ns = (unsigned long long)(blockIdx.x / p.num_sms) * p.stagger_ns;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
      __nanosleep(256);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    } while (t1 - t0 < wait_ns);
  }
  for (int k = tid; k < K; k += 256) {
    koffA[k] = (int)map_offset(p.kA, k);
    koffB[k] = (int)map_offset(p.kB, k);
  }
  if (tid < TBM)
    rowA[tid] = (m0 + tid < M) ? map_offset(p.mA, m0 + tid) : -1;
  else if (tid < TBM + TBN)
    rowB[tid - TBM] = (n0 + (tid - TBM) < N) ? map_offset(p.nB, n0 + (tid - TBM)) : -1;
  __syncthreads();

  // row-first: thread -> (row = tid % R, k-row = tid / R + q * (256 / R))
  // k-first:   thread -> (k-row = tid % BK, row = tid / BK + q * (256 / BK))
  constexpr int KF_STEP = 256 / BK, A_KF_CNT = TBM / KF_STEP;
  constexpr int B_KF_CNT = TBN >= KF_STEP ? TBN / KF_STEP : 1;
  const int ma = tid % TBM, ka0 = tid / TBM, mb = tid % TBN, kb0 = tid / TBN;
  const int kf_k = tid % BK, kf_r = tid / BK;
  const long long ra = AKF ? 0 : rowA[ma], rb = BKF ? 0 : rowB[mb];
  auto load_stage = [&](int stage, int kt) {
    const int k0 = kt * BK;
    if (!AKF) {
      double2* dA = sA + stage * A_ELEMS + ka0 * PA + ma;
#pragma unroll
      for (int q = 0; q < A_CNT; ++q) {
        const int k = k0 + ka0 + q * A_STEP;
        const bool v = (k < K) && (ra >= 0);
        cp_async16(dA + q * A_STEP * PA, v ? (A + ra + koffA[k]) : A, v);
      }
    } else {
      double2* dA = sA + stage * A_ELEMS + kf_k * PA + kf_r;
      const int k = k0 + kf_k;
      const int ko = (k < K) ? koffA[k] : 0;
#pragma unroll
      for (int q = 0; q < A_KF_CNT; ++q) {
        const long long r = rowA[kf_r + q * KF_STEP];
        const bool v = (k < K) && (r >= 0);
        cp_async16(dA + q * KF_STEP, v ? (A + r + ko) : A, v);
      }
    }
    if (!BKF) {
      if (TBN * BK >= 256 || kb0 < BK) {
        double2* dB = sB + stage * B_ELEMS + kb0 * PB + mb;
#pragma unroll
        for (int q = 0; q < B_CNT; ++q) {
          const int k = k0 + kb0 + q * B_STEP;
          const bool v = (k < K) && (rb >= 0);
          cp_async16(dB + q * B_STEP * PB, v ? (B + rb + koffB[k]) : B, v);
        }
      }
    } else {
      if (TBN >= KF_STEP || kf_r < TBN) {
        double2* dB = sB + stage * B_ELEMS + kf_k * PB + kf_r;
        const int k = k0 + kf_k;
        const int ko = (k < K) ? koffB[k] : 0;
#pragma unroll
        for (int q = 0; q < B_KF_CNT; ++q) {
          const long long r = rowB[kf_r + q * KF_STEP];
          const bool v = (k < K) && (r >= 0);
          cp_async16(dB + q * KF_STEP, v ? (B + r + ko) : B, v);
        }
      }
    }
  };

  // M3: cr = P1, ci = P2, p3 = P3
  double cr[XT][YT][2], ci[XT][YT][2], p3[M3 ? XT : 1][M3 ? YT : 1][2];
#pragma unroll
  for (int x = 0; x < XT; ++x)
#pragma unroll
    for (int y = 0; y < YT; ++y) {
      cr[x][y][0] = cr[x][y][1] = 0.0;
      ci[x][y][0] = ci[x][y][1] = 0.0;
      if (M3) p3[x][y][0] = p3[x][y][1] = 0.0;
    }

#pragma unroll
  for (int s = 0; s < NST - 1; ++s) {
    if (s < KT) load_stage(s, s);
    cp_async_commit();
  }

  const int frow = lane >> 2, fk = lane & 3;
  int ring = 0, pring = NST - 1;
  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<NST - 2>();
    __syncthreads();
    if (kt + NST - 1 < KT) load_stage(pring, kt + NST - 1);
    cp_async_commit();
    const double2* tA = sA + ring * A_ELEMS + wm * (XT * 8) + frow;
    const double2* tB = sB + ring * B_ELEMS + wn * (YT * 8) + frow;
#pragma unroll
    for (int ks = 0; ks < BK; ks += 4) {
      double2 a[XT], b[YT];
#pragma unroll
      for (int x = 0; x < XT; ++x) a[x] = tA[(ks + fk) * PA + x * 8];
#pragma unroll
      for (int y = 0; y < YT; ++y) b[y] = tB[(ks + fk) * PB + y * 8];
      if (M3) {
        double as[XT], bs[YT];
#pragma unroll
        for (int x = 0; x < XT; ++x) as[x] = a[x].x + a[x].y;
#pragma unroll
        for (int y = 0; y < YT; ++y) bs[y] = b[y].x + b[y].y;
#pragma unroll
        for (int x = 0; x < XT; ++x)
#pragma unroll
          for (int y = 0; y < YT; ++y) {
            dmma(cr[x][y][0], cr[x][y][1], a[x].x, b[y].x);
            dmma(ci[x][y][0], ci[x][y][1], a[x].y, b[y].y);
            dmma(p3[x][y][0], p3[x][y][1], as[x], bs[y]);
          }
      } else {
#pragma unroll
        for (int x = 0; x < XT; ++x)
#pragma unroll
          for (int y = 0; y < YT; ++y) {
            dmma(cr[x][y][0], cr[x][y][1], a[x].x, b[y].x);
            dmma(ci[x][y][0], ci[x][y][1], a[x].x, b[y].y);
          }
#pragma unroll
        for (int x = 0; x < XT; ++x) {
          const double nai = -a[x].y;
#pragma unroll
          for (int y = 0; y < YT; ++y) {
            dmma(cr[x][y][0], cr[x][y][1], nai, b[y].y);
            dmma(ci[x][y][0], ci[x][y][1], a[x].y, b[y].x);
          }
        }
      }
    }
    ring = (ring + 1 == NST) ? 0 : ring + 1;
    pring = (pring + 1 == NST) ? 0 : pring + 1;
  }
  cp_async_wait<0>();

#pragma unroll
  for (int y = 0; y < YT; ++y)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const long long n = n0 + wn * (YT * 8) + y * 8 + 2 * fk + c;
      if (n >= N) continue;
#pragma unroll
      for (int x = 0; x < XT; ++x) {
        const long long m = m0 + wm * (XT * 8) + x * 8 + frow;
        if (m >= M) continue;
        if (M3)
          C[m + M * n] = make_double2(cr[x][y][c] - ci[x][y][c],
                                      p3[x][y][c] - cr[x][y][c] - ci[x][y][c]);
        else
          C[m + M * n] = make_double2(cr[x][y][c], ci[x][y][c]);
      }
    }
}

// ---------------------------------------------------------------------------
// Thin-N fused TTGT ZGEMM: N <= 16, K <= 64, M huge (a site tensor with one open bond of 8
// absorbed into the boundary: per row of A, 1 KB in and 128 B out -- a stream).  No shared-
// memory staging of A at all: the m8n8k4 DMMA's A fragment is one element per lane (row
// lane / 4, k lane % 4), so a warp owns 8 rows per step and fetches its fragments of all K
// straight from the un-permuted tensor with K / 4 independent 16-byte loads per lane (8 KB in
// flight per warp, 16 warps per SM), then runs the 4 * K / 4 DMMAs against B in shared memory
// (K x N, gathered once per CTA).  C[m + M n]: 8 lanes x 16 B contiguous per column.
//
// What decides the speed of such a gather is how many 128-byte lines a QUARTER warp (8 lanes,
// one L1 tag pass) touches per load -- measured on M = 2^18, N = 8, K = 64 (ncu durations):
//     1 line  57 us (tiled kernel, rows fastest)     2 lines 56 us (this kernel, k fastest)
//     4 lines 70 us (this kernel, rows fastest, fragment order)     8 lines 92 us (tiled, k fastest)
// So the load order follows the operand: KF (a contracted axis is the fastest of A): lane ->
// (row lane / 4, k lane % 4), the fragment order itself, 4 x 16 B contiguous per row; otherwise
// lane -> (row lane % 8, k lane / 8), 8 x 16 B contiguous per k, and the 8 x 4 lane grid is
// transposed into fragment order with two 64-bit shuffles per load.  Loads and DMMAs are
// volatile asm, i.e. kept in program order: all loads of a step are in flight before the first
// DMMA waits on one (left to itself the compiler re-uses fragment registers and issues the
// loads in dependent batches between the DMMAs).  A half-step software pipeline (loads of the
// next half of K in flight under the DMMAs of this one) measured the same and was dropped.
// Replaces the 128 x 8 tile-per-CTA configuration of k_zgemm_fused_t on these steps.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double2 thin_load(const double2* ptr, bool pred) {
  double2 v;
  asm volatile(
      "{\n .reg .pred q;\n setp.ne.b32 q, %3, 0;\n mov.f64 %0, 0d0000000000000000;\n"
      " mov.f64 %1, 0d0000000000000000;\n @q ld.global.nc.v2.f64 {%0, %1}, [%2];\n}\n"
      : "=d"(v.x), "=d"(v.y)
      : "l"(ptr), "r"((int)pred));
  return v;
}

template <int KS, int NB, bool KF>   // K <= 4 KS, N <= 8 NB; KF: load in fragment order
__global__ void __launch_bounds__(256, 2)
k_zgemm_thin(const double2* __restrict__ A, const double2* __restrict__ B,
             double2* __restrict__ C, const FusedParams p) {
  constexpr int KP = 4 * KS, NP = 8 * NB;
  // (pitch = 2 (mod 8) double2: the four k rows a quarter warp reads fall into different banks)
  __shared__ __align__(16) double2 sB[KP][NP + 2];
  __shared__ int koffA[KP];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int frow = lane >> 2, fk = lane & 3;                     // fragment order
  const int lrow = KF ? frow : (lane & 7), lk = KF ? fk : (lane >> 3);   // load order
  const int src = frow + 8 * fk;   // !KF: the lane that loaded this lane's fragment
  const long long M = p.M;
  const int N = (int)p.N, K = (int)p.K;

  for (int k = tid; k < KP; k += 256) koffA[k] = k < K ? (int)map_offset(p.kA, k) : -1;
  for (int i = tid; i < KP * NP; i += 256) {
    const int k = i / NP, n = i % NP;
    sB[k][n] = (k < K && n < N) ? B[map_offset(p.nB, n) + map_offset(p.kB, k)] : make_double2(0.0, 0.0);
  }
  __syncthreads();
  int ko[KS];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) ko[ks] = koffA[4 * ks + lk];

  const long long blocks = (M + 7) / 8, stride = (long long)gridDim.x * 8;
  long long blk = (long long)blockIdx.x * 8 + warp;
  bool lvalid = blk < blocks && blk * 8 + lrow < M;
  long long ra = lvalid ? map_offset(p.mA, blk * 8 + lrow) : 0;
  while (blk < blocks) {
    const long long m = blk * 8 + frow;
    double2 a[KS];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) a[ks] = thin_load(A + ra + ko[ks], lvalid && ko[ks] >= 0);
    // the row offset of the warp's next step, behind the loads
    const long long nblk = blk + stride;
    const bool nvalid = nblk < blocks && nblk * 8 + lrow < M;
    const long long nra = nvalid ? map_offset(p.mA, nblk * 8 + lrow) : 0;
    double cr[NB][2], ci[NB][2];
#pragma unroll
    for (int y = 0; y < NB; ++y) cr[y][0] = cr[y][1] = ci[y][0] = ci[y][1] = 0.0;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      double ax = a[ks].x, ay = a[ks].y;
      if (!KF) {
        ax = __shfl_sync(0xffffffffu, ax, src);
        ay = __shfl_sync(0xffffffffu, ay, src);
      }
      const double nay = -ay;
#pragma unroll
      for (int y = 0; y < NB; ++y) {
        const double2 b = sB[4 * ks + fk][8 * y + frow];
        dmma(cr[y][0], cr[y][1], ax, b.x);
        dmma(ci[y][0], ci[y][1], ax, b.y);
        dmma(cr[y][0], cr[y][1], nay, b.y);
        dmma(ci[y][0], ci[y][1], ay, b.x);
      }
    }
    if (m < M) {
#pragma unroll
      for (int y = 0; y < NB; ++y)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int n = 8 * y + 2 * fk + c;
          if (n < N) C[m + M * n] = make_double2(cr[y][c], ci[y][c]);
        }
    }
    blk = nblk;
    lvalid = nvalid;
    ra = nra;
  }
}

// ---------------------------------------------------------------------------
// Persistent fused TTGT ZGEMM for the skinny sweep steps (K <= 64, 16 < N <= 64, M huge).
//
// Measured on the tile-per-CTA kernel above (M = 2^18, N = K = 64): with the DMMAs removed
// the data movement alone takes 183 us of the 294 us, and a 3M variant with 25 % fewer
// DMMAs runs in the same time -- each short-lived CTA pays a serial fill (tables, first
// k-block, L2 latency) for 1 MFLOP of work, every A tile is fetched twice (two n-tiles) and
// B 8192 times.  Here one CTA per SM lives for the whole GEMM:
//   * all of B (K x N <= 64 KB) is gathered into shared memory ONCE per CTA;
//   * a tile is 32 rows of A x ALL of K and N, so A is read exactly once and C written
//     exactly once;
//   * the 16 warps form TWO independent groups of 8 (2 x 4 warps, warp tile 16 x TBN/4).  Each
//     group walks its own tiles with its own two A stages and its own named barrier: the
//     gather of its next tile (cp.async) overlaps the DMMAs of the current one, and while one
//     group is between tiles (stores, barrier, issuing the next gather) the other group keeps
//     the FP64 pipe fed -- group 1 starts half a tile late so the two stay out of phase.
// Probes on M = 2^18, N = K = 64 (B200, throw-away builds): data movement alone (one DMMA per
// k-step) 132 us = 4.1 TB/s; without the C stores -5 %; without the 3M DADDs -1 %; four groups
// of four warps or a deeper A ring: no gain; fragments kept in registers across the k loop (no
// LDS) -7.5 %; leaving 4-8 SMs free for the tiny kernels of other graph branches: +-0.
// The two groups keep one 32 KB tile each in flight per SM (all the shared memory left beside
// B at K = N = 64), which at the ~4 TB/s this gather + store pattern sustains arrives in
// ~2.4 us -- about the DMMA time of a tile -- so the remaining gap between the 173 us of
// pure 3M DMMA issue time and the measured ~255 us is load latency not fully hidden.
// ---------------------------------------------------------------------------
constexpr int SK_MAXK = 64, SK_THREADS = 512, SK_MAX_STAGES = 6;
constexpr size_t SK_SMEM_BUDGET = 224 * 1024;

// NG groups of 16 / NG warps; a group's tile is TBM = 64 / NG rows
template <int TBN, int NG>
constexpr size_t skinny_smem(int K, int stages) {
  const int K4 = (K + 3) / 4 * 4;
  return size_t(K4) * ((TBN + 2) + size_t(NG) * stages * (64 / NG + 2)) * sizeof(double2) +
         size_t(K4) * 8;
}
template <int TBN, bool AKF, bool M3, int NG, int S>
__global__ void __launch_bounds__(SK_THREADS, 1)
k_zgemm_skinny(const double2* __restrict__ A, const double2* __restrict__ B,
               double2* __restrict__ C, const FusedParams p) {
  constexpr int GT = SK_THREADS / NG;          // threads per group
  constexpr int WM = 4 / NG;                   // warps along m per group (x 4 along n)
  constexpr int TBM = 16 * WM, PA = TBM + 2, PB = TBN + 2;
  constexpr int XT = 2, YT = TBN / 4 / 8;      // warp tile 16 x (TBN / 4)
  static_assert(NG == 2 || NG == 4, "two or four groups");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int K = (int)p.K, K4 = (K + 3) / 4 * 4;
  double2* sB = reinterpret_cast<double2*>(smem_raw);   // [K4][PB]
  double2* sAall = sB + K4 * PB;                        // [group][S][K4][PA]
  int* koffA = reinterpret_cast<int*>(sAall + NG * S * K4 * PA);
  int* koffB = koffA + K4;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int group = warp / (16 / NG), wg = warp % (16 / NG), gtid = tid % GT;
  const int wm = wg % WM, wn = wg / WM;
  double2* sA = sAall + group * S * K4 * PA;
  const long long M = p.M, N = p.N;
  const long long tiles = (M + TBM - 1) / TBM;

  for (int k = tid; k < K4; k += SK_THREADS) {
    koffA[k] = k < K ? (int)map_offset(p.kA, k) : 0;
    koffB[k] = k < K ? (int)map_offset(p.kB, k) : 0;
  }
  __syncthreads();

  // B: K x TBN gathered once (columns n >= N and rows k >= K are zero-filled)
  {
    const int n = tid % TBN, kb0 = tid / TBN;
    const long long rb = n < N ? map_offset(p.nB, n) : -1;
    for (int k = kb0; k < K4; k += SK_THREADS / TBN) {
      const bool v = (k < K) && (rb >= 0);
      cp_async16(sB + k * PB + n, v ? (B + rb + koffB[k]) : B, v);
    }
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  auto issue_A = [&](int stage, long long tile) {
    const long long m0 = tile * TBM;
    double2* dst = sA + stage * K4 * PA;
    if (!AKF) {   // consecutive threads walk the rows of the tile
      const int ma = gtid % TBM, ka0 = gtid / TBM;
      const long long ra = (m0 + ma < M) ? map_offset(p.mA, m0 + ma) : -1;
#pragma unroll 4
      for (int k = ka0; k < K4; k += GT / TBM) {
        const bool v = (k < K) && (ra >= 0);
        cp_async16(dst + k * PA + ma, v ? (A + ra + koffA[k]) : A, v);
      }
    } else {      // consecutive threads walk k (the fastest axis of A is a contracted one)
      constexpr int RS = GT / 16, RQ = TBM / RS;   // rows per pass / passes
      const int kf_k = gtid % 16, kf_r = gtid / 16;
      long long r[RQ];
#pragma unroll
      for (int q = 0; q < RQ; ++q)
        r[q] = (m0 + kf_r + q * RS < M) ? map_offset(p.mA, m0 + kf_r + q * RS) : -1;
      for (int kb = 0; kb < K4; kb += 16) {
        const int k = kb + kf_k;
        if (k >= K4) break;
        const int ko = koffA[k];
#pragma unroll
        for (int q = 0; q < RQ; ++q) {
          const bool v = (k < K) && (r[q] >= 0);
          cp_async16(dst + k * PA + kf_r + q * RS, v ? (A + r[q] + ko) : A, v);
        }
      }
    }
  };
  auto group_barrier = [&]() {
    asm volatile("bar.sync %0, %1;\n" ::"r"(1 + group), "r"(GT) : "memory");
  };

  const int frow = lane >> 2, fk = lane & 3;
  const long long stride = (long long)gridDim.x * NG;
  long long tile = (long long)blockIdx.x * NG + group;
  // prologue: S - 1 tiles in flight (one commit group per tile, empty groups included)
  for (int s = 0; s < S - 1; ++s) {
    if (tile + s * stride < tiles) issue_A(s, tile + s * stride);
    cp_async_commit();
  }
  if (group > 0 && tiles > 4 * stride) {
    // the groups start 1/NG of a tile apart so that their between-tile phases do not coincide
    // (a 64-row slab is ~8 us of FP64 pipe time at K = 64)
    unsigned long long t0, t1;
    const unsigned long long wait_ns = 8000ull * (unsigned)K4 / 64 * group / (NG * NG);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
      __nanosleep(200);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    } while (t1 - t0 < wait_ns);
  }
  int stage = 0, pstage = S - 1;
  for (; tile < tiles; tile += stride) {
    cp_async_wait<S - 2>();
    group_barrier();   // this tile has landed; every warp of the group left the previous stage
    const long long next = tile + (S - 1) * stride;
    if (next < tiles) issue_A(pstage, next);
    cp_async_commit();

    // M3 (see k_zgemm_fused_t): cr = P1 = Ar Br, ci = P2 = Ai Bi, p3 = (Ar + Ai)(Br + Bi)
    double cr[XT][YT][2], ci[XT][YT][2], p3[M3 ? XT : 1][M3 ? YT : 1][2];
#pragma unroll
    for (int x = 0; x < XT; ++x)
#pragma unroll
      for (int y = 0; y < YT; ++y) {
        cr[x][y][0] = cr[x][y][1] = 0.0;
        ci[x][y][0] = ci[x][y][1] = 0.0;
        if (M3) p3[x][y][0] = p3[x][y][1] = 0.0;
      }
    const double2* tA = sA + stage * K4 * PA + wm * (XT * 8) + frow;
    const double2* tB = sB + wn * (YT * 8) + frow;
#pragma unroll 4
    for (int ks = 0; ks < K4; ks += 4) {
      double2 a[XT], b[YT];
#pragma unroll
      for (int x = 0; x < XT; ++x) a[x] = tA[(ks + fk) * PA + x * 8];
#pragma unroll
      for (int y = 0; y < YT; ++y) b[y] = tB[(ks + fk) * PB + y * 8];
      if (M3) {
        double as[XT], bs[YT];
#pragma unroll
        for (int x = 0; x < XT; ++x) as[x] = a[x].x + a[x].y;
#pragma unroll
        for (int y = 0; y < YT; ++y) bs[y] = b[y].x + b[y].y;
#pragma unroll
        for (int x = 0; x < XT; ++x)
#pragma unroll
          for (int y = 0; y < YT; ++y) {
            dmma(cr[x][y][0], cr[x][y][1], a[x].x, b[y].x);
            dmma(ci[x][y][0], ci[x][y][1], a[x].y, b[y].y);
            dmma(p3[x][y][0], p3[x][y][1], as[x], bs[y]);
          }
      } else {
#pragma unroll
        for (int x = 0; x < XT; ++x)
#pragma unroll
          for (int y = 0; y < YT; ++y) {
            dmma(cr[x][y][0], cr[x][y][1], a[x].x, b[y].x);
            dmma(ci[x][y][0], ci[x][y][1], a[x].x, b[y].y);
          }
#pragma unroll
        for (int x = 0; x < XT; ++x) {
          const double nai = -a[x].y;
#pragma unroll
          for (int y = 0; y < YT; ++y) {
            dmma(cr[x][y][0], cr[x][y][1], nai, b[y].y);
            dmma(ci[x][y][0], ci[x][y][1], a[x].y, b[y].x);
          }
        }
      }
    }
    const long long m0 = tile * TBM;
#pragma unroll
    for (int y = 0; y < YT; ++y)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const long long n = wn * (YT * 8) + y * 8 + 2 * fk + c;
        if (n >= N) continue;
#pragma unroll
        for (int x = 0; x < XT; ++x) {
          const long long m = m0 + wm * (XT * 8) + x * 8 + frow;
          if (m >= M) continue;
          if (M3)
            C[m + M * n] = make_double2(cr[x][y][c] - ci[x][y][c],
                                        p3[x][y][c] - cr[x][y][c] - ci[x][y][c]);
          else
            C[m + M * n] = make_double2(cr[x][y][c], ci[x][y][c]);
        }
      }
    stage = (stage + 1 == S) ? 0 : stage + 1;
    pstage = (pstage + 1 == S) ? 0 : pstage + 1;
  }
  cp_async_wait<0>();
}

template <int TBM, int TBN, int NST>
constexpr size_t fused_smem(int K) {
  return size_t(NST) * BK * ((TBM + 2) + (TBN + 2)) * sizeof(double2) + size_t(TBM + TBN) * 8 +
         size_t(K) * 8;
}

// FP64 issue-rate probes for the roofline denominators (dependent chains per warp are
// kept short enough to saturate the pipe with 8 warps x 4 independent accumulators).
__global__ void __launch_bounds__(1024) k_probe_dmma(double* out, int iters) {
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

template <bool AKF, bool BKF>
static void init_fused() {
  PQ_CUDA(cudaFuncSetAttribute(k_zgemm_fused_t<64, 64, 2, 4, 3, 2, AKF, BKF>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)fused_smem<64, 64, 3>(FUSED_MAX_K)));
  PQ_CUDA(cudaFuncSetAttribute(k_zgemm_fused_t<64, 32, 4, 2, 2, 4, AKF, BKF>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)fused_smem<64, 32, 2>(FUSED_MAX_K)));
  PQ_CUDA(cudaFuncSetAttribute(k_zgemm_fused_t<128, 8, 8, 1, 2, 3, AKF, BKF>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)fused_smem<128, 8, 2>(FUSED_MAX_K)));
  PQ_CUDA(cudaFuncSetAttribute(k_zgemm_fused_t<64, 32, 4, 2, 2, 3, AKF, BKF, true>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)fused_smem<64, 32, 2>(FUSED_MAX_K)));
}

template <bool AKF, bool BKF>
static void launch_fused(int cfg, const Launch& L, const FusedParams& fp, const void* A,
                         const void* B, void* C) {
  if (cfg == 4) {   // 64x32 tiles with 3M products, 3 CTAs/SM (A/B: same time as cfg 2)
    long long tiles = ((fp.M + 63) / 64) * ((fp.N + 31) / 32);
    PQ_REQUIRE(tiles <= 0x7fffffffLL, PQ_ERR_UNSUPPORTED, "too many tiles");
    k_zgemm_fused_t<64, 32, 4, 2, 2, 3, AKF, BKF, true>
        <<<(unsigned)tiles, 256, fused_smem<64, 32, 2>((int)fp.K), L.stream>>>(
            (const double2*)A, (const double2*)B, (double2*)C, fp);
  } else if (cfg == 3) {
    long long tiles = ((fp.M + 127) / 128) * ((fp.N + 7) / 8);
    PQ_REQUIRE(tiles <= 0x7fffffffLL, PQ_ERR_UNSUPPORTED, "too many tiles");
    k_zgemm_fused_t<128, 8, 8, 1, 2, 3, AKF, BKF>
        <<<(unsigned)tiles, 256, fused_smem<128, 8, 2>((int)fp.K), L.stream>>>(
            (const double2*)A, (const double2*)B, (double2*)C, fp);
  } else if (cfg == 2) {
    long long tiles = ((fp.M + 63) / 64) * ((fp.N + 31) / 32);
    PQ_REQUIRE(tiles <= 0x7fffffffLL, PQ_ERR_UNSUPPORTED, "too many tiles");
    k_zgemm_fused_t<64, 32, 4, 2, 2, 4, AKF, BKF>
        <<<(unsigned)tiles, 256, fused_smem<64, 32, 2>((int)fp.K), L.stream>>>(
            (const double2*)A, (const double2*)B, (double2*)C, fp);
  } else {
    long long tiles = ((fp.M + 63) / 64) * ((fp.N + 63) / 64);
    PQ_REQUIRE(tiles <= 0x7fffffffLL, PQ_ERR_UNSUPPORTED, "too many tiles");
    k_zgemm_fused_t<64, 64, 2, 4, 3, 2, AKF, BKF>
        <<<(unsigned)tiles, 256, fused_smem<64, 64, 3>((int)fp.K), L.stream>>>(
            (const double2*)A, (const double2*)B, (double2*)C, fp);
  }
}

template <bool KF>
static void launch_thin(const Launch& L, const FusedParams& fp, const void* A, const void* B, void* C) {
  const unsigned grid = (unsigned)(2 * L.num_sms);
  const double2 *a = (const double2*)A, *b = (const double2*)B;
  double2* c = (double2*)C;
  if (fp.N > 16) {   // short contraction, up to 64 columns: K <= 16
    if (fp.K <= 8) {
      if (fp.N <= 32) k_zgemm_thin<2, 4, KF><<<grid, 256, 0, L.stream>>>(a, b, c, fp);
      else k_zgemm_thin<2, 8, KF><<<grid, 256, 0, L.stream>>>(a, b, c, fp);
    } else {
      if (fp.N <= 32) k_zgemm_thin<4, 4, KF><<<grid, 256, 0, L.stream>>>(a, b, c, fp);
      else k_zgemm_thin<4, 8, KF><<<grid, 256, 0, L.stream>>>(a, b, c, fp);
    }
  } else if (fp.K <= 32) {
    if (fp.N <= 8) k_zgemm_thin<8, 1, KF><<<grid, 256, 0, L.stream>>>(a, b, c, fp);
    else k_zgemm_thin<8, 2, KF><<<grid, 256, 0, L.stream>>>(a, b, c, fp);
  } else {
    if (fp.N <= 8) k_zgemm_thin<16, 1, KF><<<grid, 256, 0, L.stream>>>(a, b, c, fp);
    else k_zgemm_thin<16, 2, KF><<<grid, 256, 0, L.stream>>>(a, b, c, fp);
  }
}

constexpr int SK_NG = 2, SK_S = 2;   // measured: four groups / deeper rings are not faster (six stages at K = 8: 72 vs 67 us)

template <int TBN, bool AKF>
static void init_skinny() {
  PQ_CUDA(cudaFuncSetAttribute(k_zgemm_skinny<TBN, AKF, false, SK_NG, SK_S>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SK_SMEM_BUDGET));
  PQ_CUDA(cudaFuncSetAttribute(k_zgemm_skinny<TBN, AKF, true, SK_NG, SK_S>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SK_SMEM_BUDGET));
}

template <int TBN, bool AKF>
static void launch_skinny(bool m3, unsigned grid, const Launch& L, const FusedParams& fp,
                          const void* A, const void* B, void* C) {
  const size_t smem = skinny_smem<TBN, SK_NG>((int)fp.K, SK_S);
  if (m3)
    k_zgemm_skinny<TBN, AKF, true, SK_NG, SK_S><<<grid, SK_THREADS, smem, L.stream>>>(
        (const double2*)A, (const double2*)B, (double2*)C, fp);
  else
    k_zgemm_skinny<TBN, AKF, false, SK_NG, SK_S><<<grid, SK_THREADS, smem, L.stream>>>(
        (const double2*)A, (const double2*)B, (double2*)C, fp);
}

void init_kernels() {
  PQ_CUDA(cudaFuncSetAttribute(k_zgemm_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)SMEM_BYTES));
  PQ_CUDA(cudaFuncSetAttribute(k_zgemm_dmma3m, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)M3_SMEM_BYTES));
  init_skinny<64, false>();
  init_skinny<64, true>();
  init_skinny<32, false>();
  init_skinny<32, true>();
  init_fused<false, false>();
  init_fused<false, true>();
  init_fused<true, false>();
  init_fused<true, true>();
}

void run_zgemm_dmma(const Launch& L, const void* A, const void* B, void* C, int64_t M, int64_t N,
                    int64_t K) {
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN));
  PQ_REQUIRE(grid.y <= 65535, PQ_ERR_UNSUPPORTED, "N too large for the ZGEMM grid");
  double bytes = double(M * K + N * K + M * N) * 16.0, flops = 8.0 * M * N * K;
  // long contractions are bound by the FP64 pipe: 3M products (option zgemm_3m=1: four DMMAs)
  const long long tiles = (long long)grid.x * grid.y;
  const bool m3 = !(L.opt && L.opt->zgemm_3m == 1) && K >= 256 && tiles <= 0x7fffffffLL;
  L.begin(KC_GEMM_TENSOR, bytes, flops);
  if (m3)
    k_zgemm_dmma3m<<<(unsigned)tiles, M3_THREADS, M3_SMEM_BYTES, L.stream>>>(
        (const double2*)A, (const double2*)B, (double2*)C, M, N, K);
  else
    k_zgemm_dmma<<<grid, 256, SMEM_BYTES, L.stream>>>((const double2*)A, (const double2*)B,
                                                      (double2*)C, M, N, K);
  L.end();
  PQ_CUDA(cudaGetLastError());
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
  PQ_REQUIRE(cp.K <= FUSED_MAX_K, PQ_ERR_INVALID, "fused ZGEMM: K too large");
  double bytes = double(cp.M * cp.K + cp.N * cp.K + cp.M * cp.N) * 16.0;
  double flops = 8.0 * double(cp.M) * double(cp.N) * double(cp.K);
  // short contractions (few k-blocks per tile) are fill-latency bound: use the small-CTA
  // configuration with four resident CTAs per SM; option "zgemm_cfg" forces 1 (64x64) / 2 (64x32)
  int cfg = L.opt ? L.opt->zgemm_cfg : 0;
  if (cfg == 0) cfg = (cp.N <= 16) ? 3 : (cp.K <= 128) ? 2 : 1;
  {
    const int per_sm = (cfg == 1) ? 2 : (cfg == 3 || cfg == 4) ? 3 : 4;
    PQ_REQUIRE(cfg >= 1 && cfg <= 4 || cfg == 7, PQ_ERR_INVALID, "zgemm_cfg must be 0..4 or 7");
    fp.num_sms = L.num_sms;
    fp.first_wave = L.num_sms * per_sm;
    fp.stagger_ns = L.opt ? L.opt->zgemm_stagger : 0;
  }
  auto min_stride = [](const IdxMap& m) {
    int64_t best = INT64_MAX;
    for (int d = 0; d < m.nd; ++d) best = m.str[d] < best ? m.str[d] : best;
    return best;
  };
  bool akf = min_stride(cp.kA) < min_stride(cp.mA), bkf = min_stride(cp.kB) < min_stride(cp.nB);
  if (L.opt && L.opt->zgemm_kfirst == 1) akf = bkf = false;  // A/B check knob
  if (cfg == 3) akf = bkf = false;  // measured: no gain for the narrow tiles (HBM-bound)
  // INT8 tensor-core Ozaki product (default policy ozaki_t_preferred, or forced by zgemm_ozaki = 6)
  {
    int oz = L.opt ? L.opt->zgemm_ozaki : 0;
    if (oz == 0 && (L.opt == nullptr || (L.opt->ozaki_auto != 0 && L.opt->zgemm_cfg == 0 && L.opt->zgemm_skinny == 0)) &&
        ozaki_t_preferred(16, cp.M, cp.N, cp.K))
      oz = 6;   // default policy: k_ozaki_t
    if (oz != 0 && zgemm_ozaki_eligible(cp.M, cp.N, cp.K)) {
      L.begin(KC_GEMM_INT8, bytes, flops);
      run_zgemm_ozaki_t(L, fp, A, B, C);
      L.end();
      PQ_CUDA(cudaGetLastError());
      return;
    }
  }
  // thin N, a contracted axis fastest in A: fragments straight from the un-permuted tensor
  // (k_zgemm_thin, 58 us against 93 us of the tiled kernel on M = 2^18, N = 8, K = 64); with an
  // open axis fastest the tiled kernel's row-first gather already touches one line per quarter
  // warp and stays ahead (58 vs 61 us).  Options: an explicit zgemm_cfg or zgemm_thin = 1 keep
  // the tiled / skinny kernels, zgemm_thin = 2 / 3 force k_zgemm_thin with the k-first /
  // rows-first order on N <= 16.
  {
    const int thin = L.opt ? L.opt->zgemm_thin : 0;
    PQ_REQUIRE(thin >= 0 && thin <= 4, PQ_ERR_INVALID, "zgemm_thin must be 0..4");
    const bool k_fastest = min_stride(cp.kA) < min_stride(cp.mA);
    // short contraction with a wide small side (16 < N <= 64, output-bound): the same kernel with
    // up to eight 8-column blocks; ncu durations against the persistent skinny kernel at M = 2^18,
    // K = 8: N = 64 63.8 vs 68.3 us (75.8 when a contracted axis is fastest), N = 32 35.9 vs 42.3;
    // K = 16: 91.3 vs 91.4, left to the skinny kernel (option zgemm_thin = 4 forces K <= 16)
    const bool wide = cp.N > 16 && cp.N <= 64 && cp.M >= 4096 &&
                      ((thin == 0 && cp.K <= 8) || (thin == 4 && cp.K <= 16));
    if ((L.opt ? L.opt->zgemm_cfg : 0) == 0 && ((cp.N <= 16 && cp.K <= 64 &&
        (thin == 2 || thin == 3 || (thin == 0 && k_fastest))) || wide)) {
      const bool kf = thin == 2 || ((thin == 0 || thin == 4) && k_fastest);
      L.begin(KC_GEMM_TENSOR, bytes, flops);
      if (kf)
        launch_thin<true>(L, fp, A, B, C);
      else
        launch_thin<false>(L, fp, A, B, C);
      L.end();
      PQ_CUDA(cudaGetLastError());
      return;
    }
  }
  // persistent skinny kernel: K and N small enough for B to live in shared memory, and
  // enough 64-row tiles to keep every SM busy for several tiles
  const bool skinny_ok = cp.K <= SK_MAXK && cp.N <= 64 && cp.N > 16 && cp.M >= 256LL * L.num_sms;
  const int skinny = L.opt ? L.opt->zgemm_skinny : 0;   // 0 auto, 1 off
  if ((cfg == 7 || ((L.opt ? L.opt->zgemm_cfg : 0) == 0 && skinny == 0)) && skinny_ok) {
    const unsigned grid = (unsigned)L.num_sms;
    if (L.opt && L.opt->zgemm_kfirst == 1) akf = false;
    L.begin(KC_GEMM_TENSOR, bytes, flops);
    const bool m3 = !(L.opt && L.opt->zgemm_3m == 1);   // option zgemm_3m=1: four DMMAs per product
    if (cp.N > 32) {
      if (akf)
        launch_skinny<64, true>(m3, grid, L, fp, A, B, C);
      else
        launch_skinny<64, false>(m3, grid, L, fp, A, B, C);
    } else {
      if (akf)
        launch_skinny<32, true>(m3, grid, L, fp, A, B, C);
      else
        launch_skinny<32, false>(m3, grid, L, fp, A, B, C);
    }
    L.end();
    PQ_CUDA(cudaGetLastError());
    return;
  }
  if (cfg == 7) cfg = (cp.K <= 128) ? 2 : 1;
  L.begin(KC_GEMM_TENSOR, bytes, flops);
  if (akf && bkf)
    launch_fused<true, true>(cfg, L, fp, A, B, C);
  else if (akf)
    launch_fused<true, false>(cfg, L, fp, A, B, C);
  else if (bkf)
    launch_fused<false, true>(cfg, L, fp, A, B, C);
  else
    launch_fused<false, false>(cfg, L, fp, A, B, C);
  L.end();
  PQ_CUDA(cudaGetLastError());
}

// returns achieved TFLOP/s of the probe ("dmma" or "dfma")
// warps_per_sm = 0: the default saturating shape (8 CTAs of 8 warps per SM); otherwise ONE
// CTA of that many warps per SM -- how many resident warps the FP64 pipe needs
double run_fp64_probe(const Launch& L, bool tensor, int warps_per_sm) {
  if (warps_per_sm > 0) {
    const int blocks = L.num_sms, threads = 32 * warps_per_sm, iters = 8192;
    double* out = nullptr;
    PQ_CUDA(cudaMalloc(&out, size_t(blocks) * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    PQ_CUDA(cudaEventCreate(&e0));
    PQ_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      PQ_CUDA(cudaEventRecord(e0, L.stream));
      k_probe_dmma<<<blocks, threads, 0, L.stream>>>(out, iters);
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
    return double(blocks) * warps_per_sm * iters * 8.0 * (8 * 8 * 4 * 2) / (best * 1e-3) / 1e12;
  }
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
