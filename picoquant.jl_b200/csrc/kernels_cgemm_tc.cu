// K3: ComplexF32 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), with
// split-precision TF32 emulation ("3xTF32") so that results stay inside the 1e-5 rel-L2
// tolerance of the ComplexF32 backend.
//
// Canonical TTGT layouts in (produced by K1, interleaved complex float2):
//     A[m + M*k]   B[n + N*k]   ->   C[m + M*n]
// Per CTA: a 128 x BN output tile (UMMA M = 128, N = BN, cta_group::1), accumulators in
// TMEM: Cr in columns [0, BN), Ci in columns [BN, 2*BN), fp32.
//
// Operand staging.  tcgen05.mma reads its operands from shared memory through matrix
// descriptors.  The loader threads read the interleaved complex tile, split every real
// number x into hi = tf32(x) and lo = x - hi, and write FOUR real planes per operand
// (re_hi, re_lo, im_hi, im_lo) in the no-swizzle K-major core-matrix layout
//     offset(row, k) = (k/4)*LBO + (row/8)*128 + (row%8)*16 + (k%4)*4        [bytes]
// (core matrix = 8 rows x 16 bytes, SBO = 128 B between 8-row groups, LBO between
// 16-byte K chunks).  The split therefore costs no extra pass over HBM.
//
// Math per K = 8 step (tf32 UMMA K):  12 MMAs
//     Cr += ArH*BrH - AiH*BiH                    Xr += ArH*BrL + ArL*BrH - (AiH*BiL + AiL*BiH)
//     Ci += ArH*BiH + AiH*BrH                    Xi += ArH*BiL + ArL*BiH +  AiH*BrL + AiL*BrH
// i.e. hi*hi + hi*lo + lo*hi per real product (the lo*lo term, ~2^-22 relative, is dropped;
// `a_negate` gives the minus sign).  The cross terms go to their OWN accumulators (Xr, Xi):
// every MMA truncates when it adds into its accumulator, an error relative to the
// accumulator's magnitude.  Kept apart, the 8 cross-term MMAs only disturb sums that are
// 2^-11 of the result, and the main accumulators see 2 instead of 6 truncations per K-step.
//
// Pipeline: two shared-memory stages; all 8 warps load+split stage s while the tensor core
// works on stage s^1; one elected thread issues the MMAs and tcgen05.commit signals an
// mbarrier when the stage may be overwritten.  Epilogue: tcgen05.ld (32 lanes x 32 bit per
// warp quadrant) -> registers -> interleaved float2 stores, coalesced along m.
#include "common.h"

namespace pq {

namespace {

constexpr int TM = 128;      // UMMA M
constexpr int TK = 32;       // k elements per stage (4 UMMA K-steps of 8)
constexpr int NSTAGE = 2;
constexpr int NTHREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// no-swizzle K-major matrix descriptor (sm_100 "version 1")
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version for Blackwell
  return d;                // layout_type (bits 61..63) = 0: SWIZZLE_NONE
}

// instruction descriptor, kind::tf32, fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool neg_a) {
  return (1u << 4)                       // c_format = F32
         | (2u << 7)                     // a_format = TF32
         | (2u << 10)                    // b_format = TF32
         | ((neg_a ? 1u : 0u) << 13)     // a_negate
         | ((uint32_t)(N >> 3) << 17)    // n_dim
         | ((uint32_t)(M >> 4) << 24);   // m_dim
}

// whole-warp form: every lane executes the (uniform) surrounding code, one elected lane issues
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}

// hi part of the split: x rounded to TF32 (10 explicit mantissa bits), done with one integer
// add and one mask.  cvt.rna.tf32.f32 compiles to a ~10-instruction emulation on sm_100a
// (ncu source view: FSETP/SEL/LOP3 chains were a third of all issued instructions); the
// integer form differs only for NaN/Inf inputs, which the tensor core would propagate anyway.
__device__ __forceinline__ float tf32_hi(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

// byte offset of (row, k) inside one plane of an operand tile with `rows` rows
__device__ __forceinline__ uint32_t plane_off(int rows, int row, int k) {
  return (uint32_t)((k >> 2) * (rows * 16) + (row >> 3) * 128 + (row & 7) * 16 + (k & 3) * 4);
}

template <int BN>
struct Cfg {
  static constexpr int A_PLANE = TM * TK * 4;   // bytes
  static constexpr int B_PLANE = BN * TK * 4;
  static constexpr int STAGE = 4 * A_PLANE + 4 * B_PLANE;
  static constexpr int SMEM = NSTAGE * STAGE + 128;
  // two hi*hi accumulator sets (ping-pong over K chunks) + the running sums + the
  // cross-term accumulators, 2*BN fp32 columns each
  static constexpr int TMEM_COLS = (8 * BN <= 32) ? 32 : (8 * BN <= 64) ? 64 : (8 * BN <= 128) ? 128
                                   : (8 * BN <= 256) ? 256 : 512;
  static_assert(8 * BN <= 512, "TMEM has 512 columns");
};

// One loader work item = (row, 16-byte K chunk) = 4 consecutive k of one row of a
// ROWS x TK complex tile stored as src[row + ld*k].  Items are strided over the loader
// threads.  LoadMap holds what does not change from stage to stage (global base pointer,
// shared-memory offset); LoadRegs is one register set of fetched values.  `fetch` only
// issues the global loads, `split_store` converts and writes the four planes, so loads of
// later stages are in flight while the current one is converted and fed to the tensor core.
template <int ROWS>
struct LoadMap {
  static constexpr int ITEMS = ROWS * (TK / 4);
  static constexpr int NITEMS = (ITEMS + NTHREADS - 1) / NTHREADS;
  const float2* base[NITEMS];  // element (row, k = 4*chunk) of stage 0; nullptr: row out of range
  uint32_t off[NITEMS];        // byte offset inside a plane
  int kfirst[NITEMS];          // 4*chunk
  long long ld;

  __device__ __forceinline__ void init(const float2* __restrict__ src, long long ld_, long long row0,
                                       long long nrows, int tid) {
    ld = ld_;
#pragma unroll
    for (int i = 0; i < NITEMS; ++i) {
      const int it = tid + i * NTHREADS;
      const int row = it % ROWS, chunk = it / ROWS;
      const bool ok = (it < ITEMS) && (row0 + row < nrows);
      kfirst[i] = chunk * 4;
      base[i] = ok ? src + (row0 + row) + ld * (chunk * 4) : nullptr;
      off[i] = (it < ITEMS) ? plane_off(ROWS, row, chunk * 4) : 0xffffffffu;
    }
  }
};

template <int ROWS>
struct LoadRegs {
  static constexpr int NITEMS = LoadMap<ROWS>::NITEMS;
  float2 v[NITEMS][4];

  __device__ __forceinline__ void fetch(const LoadMap<ROWS>& m, long long kt, long long K) {
    const long long k0 = kt * TK;
    const bool full = k0 + TK <= K;  // uniform: only the last stage can be ragged
#pragma unroll
    for (int i = 0; i < NITEMS; ++i) {
      const float2* p = m.base[i];
      if (p != nullptr) {
        p += k0 * m.ld;
        if (full) {
#pragma unroll
          for (int j = 0; j < 4; ++j) v[i][j] = p[j * m.ld];
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            v[i][j] = (k0 + m.kfirst[i] + j < K) ? p[j * m.ld] : make_float2(0.f, 0.f);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[i][j] = make_float2(0.f, 0.f);
      }
    }
  }
  __device__ __forceinline__ void split_store(const LoadMap<ROWS>& m, unsigned char* planes,
                                              int plane_bytes) {
#pragma unroll
    for (int i = 0; i < NITEMS; ++i) {
      if (m.off[i] == 0xffffffffu) continue;
      float4 rh, rl, ih, il;
      rh.x = tf32_hi(v[i][0].x); rh.y = tf32_hi(v[i][1].x);
      rh.z = tf32_hi(v[i][2].x); rh.w = tf32_hi(v[i][3].x);
      ih.x = tf32_hi(v[i][0].y); ih.y = tf32_hi(v[i][1].y);
      ih.z = tf32_hi(v[i][2].y); ih.w = tf32_hi(v[i][3].y);
      // (the tensor core truncates the lo parts to tf32; rounding them to nearest here was
      // measured to change the result error by < 2 % while costing loader ALU time)
      rl.x = v[i][0].x - rh.x; rl.y = v[i][1].x - rh.y;
      rl.z = v[i][2].x - rh.z; rl.w = v[i][3].x - rh.w;
      il.x = v[i][0].y - ih.x; il.y = v[i][1].y - ih.y;
      il.z = v[i][2].y - ih.z; il.w = v[i][3].y - ih.w;
      unsigned char* dst = planes + m.off[i];
      *reinterpret_cast<float4*>(dst + 0 * plane_bytes) = rh;
      *reinterpret_cast<float4*>(dst + 1 * plane_bytes) = rl;
      *reinterpret_cast<float4*>(dst + 2 * plane_bytes) = ih;
      *reinterpret_cast<float4*>(dst + 3 * plane_bytes) = il;
    }
  }
};

// The tensor core adds products into its fp32 accumulator with truncation, so the error of
// one long accumulation grows linearly with K (5.7e-5 rel-L2 at K = 4096, measured).  K is
// therefore cut into chunks of KCHUNK elements: each chunk accumulates in its own TMEM set
// (ping-pong) starting from zero, and the finished chunk is added with round-to-nearest to
// fp32 running sums kept in a third TMEM region by the four epilogue warps.
//
// Warp roles (288 threads): warps 0-7 load + split operand tiles (warps 0-3 additionally fold
// finished chunks and write C); warp 8, one elected lane, only issues tcgen05.mma.  The roles
// are coupled by mbarriers, never by __syncthreads, so loaders run up to NSTAGE stages ahead.
constexpr int KCHUNK = 256;
constexpr int STAGES_PER_CHUNK = KCHUNK / TK;
constexpr int NLOAD = NTHREADS;          // loader threads
constexpr int NALL = NTHREADS + 32;      // + MMA warp

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

#define PQ_TMEM_LD8(r, addr)                                                                  \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"    \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]),     \
                 "=r"(r[6]), "=r"(r[7])                                                       \
               : "r"(addr))
#define PQ_TMEM_ST8(addr, r)                                                                  \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"r"(addr), \
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),  \
               "r"(r[7])                                                                      \
               : "memory")

template <int BN>
__global__ void __launch_bounds__(NALL, 1)
k_cgemm_tcgen05(const float2* __restrict__ A, const float2* __restrict__ B, float2* __restrict__ C,
                long long M, long long N, long long K) {
  using cfg = Cfg<BN>;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTAGE * cfg::STAGE);
  uint64_t* full = bars;                    // [NSTAGE] loaders -> MMA   (count NLOAD)
  uint64_t* empty = bars + NSTAGE;          // [NSTAGE] MMA -> loaders   (tcgen05.commit)
  uint64_t* chunk_done = bars + 2 * NSTAGE; // [2] chunk accumulated in TMEM set (commit)
  uint64_t* drained = chunk_done + 2;       // [2] set folded into the sums (count 128)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(drained + 2);

  // (warp index through a shuffle: provably warp-uniform, so the role branches are convergent for
  // ptxas and the MMA warp's descriptor arithmetic runs on the uniform datapath -- with `if (lane ==
  // 0)` around the issue loop every UTCHMMA was wrapped in an ELECT / BRA.U.ANY loop behind four
  // R2URs, ~17 instructions per MMA on a scheduler shared with two loader warps)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  // Rasterisation: consecutive CTAs walk GROUP_M m-tiles x all n-tiles, so one wave of CTAs
  // re-uses a band of A (GROUP_M*128 rows) and a few column tiles of B out of L2 instead of
  // streaming all of A once per wave (2.16 GB -> ~0.5 GB of DRAM traffic at 4096^3).
  constexpr long long GROUP_M = 16;
  const long long tiles_m = (M + TM - 1) / TM, tiles_n = (N + BN - 1) / BN;
  const long long bid = blockIdx.x;
  const long long group = bid / (GROUP_M * tiles_n);
  const long long gm = (tiles_m - group * GROUP_M) < GROUP_M ? (tiles_m - group * GROUP_M) : GROUP_M;
  const long long rem = bid - group * GROUP_M * tiles_n;
  const long long m0 = (group * GROUP_M + rem % gm) * TM, n0 = (rem / gm) * BN;
  const long long KT = (K + TK - 1) / TK;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], NLOAD / 32);   // one arrival per loader warp
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&chunk_done[s], 1);
      mbar_init(&drained[s], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"((uint32_t)cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_sum = tmem_base + 4 * BN;   // running sums: re in [0,BN), im in [BN,2BN)
  const uint32_t tmem_x = tmem_base + 6 * BN;     // cross terms (hi*lo + lo*hi), never folded

  constexpr uint32_t IDESC = make_idesc(TM, BN, false);
  constexpr uint32_t IDESC_NEG = make_idesc(TM, BN, true);
  constexpr uint32_t A_LBO = TM * 16, B_LBO = BN * 16, SBO = 128;

  if (warp == 8) {
    // ===================== MMA issuer (all 32 lanes, uniform; one elected lane issues) =====
    {
      for (long long kt = 0; kt < KT; ++kt) {
        const int s = (int)(kt % NSTAGE);
        const long long chunk = kt / STAGES_PER_CHUNK;
        const int kin = (int)(kt - chunk * STAGES_PER_CHUNK);
        const int set = (int)(chunk & 1);
        if (kin == 0 && chunk >= 2)   // the set must have been folded (chunk - 2) before reuse
          mbar_wait(&drained[set], (uint32_t)(((chunk >> 1) - 1) & 1));
        mbar_wait(&full[s], (uint32_t)((kt / NSTAGE) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        unsigned char* stage = smem + s * cfg::STAGE;
        const uint32_t a0 = smem_u32(stage), b0 = smem_u32(stage + 4 * cfg::A_PLANE);
        const uint32_t tmem_cr = tmem_base + set * 2 * BN, tmem_ci = tmem_cr + BN;
#pragma unroll
        for (int ks = 0; ks < TK / 8; ++ks) {
          const uint32_t ao = a0 + ks * 2 * A_LBO, bo = b0 + ks * 2 * B_LBO;
          const uint64_t arh = make_desc(ao + 0 * cfg::A_PLANE, A_LBO, SBO);
          const uint64_t arl = make_desc(ao + 1 * cfg::A_PLANE, A_LBO, SBO);
          const uint64_t aih = make_desc(ao + 2 * cfg::A_PLANE, A_LBO, SBO);
          const uint64_t ail = make_desc(ao + 3 * cfg::A_PLANE, A_LBO, SBO);
          const uint64_t brh = make_desc(bo + 0 * cfg::B_PLANE, B_LBO, SBO);
          const uint64_t brl = make_desc(bo + 1 * cfg::B_PLANE, B_LBO, SBO);
          const uint64_t bih = make_desc(bo + 2 * cfg::B_PLANE, B_LBO, SBO);
          const uint64_t bil = make_desc(bo + 3 * cfg::B_PLANE, B_LBO, SBO);
          const uint32_t acc = (kin > 0 || ks > 0) ? 1u : 0u;   // a chunk starts from zero
          const uint32_t accx = (kt > 0 || ks > 0) ? 1u : 0u;   // cross terms: whole K
          const uint32_t tmem_xr = tmem_x, tmem_xi = tmem_x + BN;
          umma_tf32(tmem_cr, arh, brh, IDESC, acc);
          umma_tf32(tmem_ci, arh, bih, IDESC, acc);
          umma_tf32(tmem_xr, arh, brl, IDESC, accx);
          umma_tf32(tmem_xi, arh, bil, IDESC, accx);
          umma_tf32(tmem_cr, aih, bih, IDESC_NEG, 1u);
          umma_tf32(tmem_ci, aih, brh, IDESC, 1u);
          umma_tf32(tmem_xr, arl, brh, IDESC, 1u);
          umma_tf32(tmem_xi, arl, bih, IDESC, 1u);
          umma_tf32(tmem_xr, aih, bil, IDESC_NEG, 1u);
          umma_tf32(tmem_xi, aih, brl, IDESC, 1u);
          umma_tf32(tmem_xr, ail, bih, IDESC_NEG, 1u);
          umma_tf32(tmem_xi, ail, brh, IDESC, 1u);
        }
        umma_commit(&empty[s]);   // the stage may be refilled once these MMAs have completed
        if (kin == STAGES_PER_CHUNK - 1 || kt == KT - 1) umma_commit(&chunk_done[set]);
      }
    }
  } else {
    // ===================== loaders (+ chunk folding on warps 0-3) =====================
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    // sum (+)= TMEM set of chunk c; `final` writes C instead of storing the sums back
    auto fold = [&](long long c, bool final) {
      const int set = (int)(c & 1);
      mbar_wait(&chunk_done[set], (uint32_t)((c >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const uint32_t tr = tmem_base + set * 2 * BN + lane_base, ti = tr + BN;
      const uint32_t sr = tmem_sum + lane_base, si = sr + BN;
      const long long m = m0 + (warp & 3) * 32 + lane;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 8) {
        uint32_t r[8], q[8], ar[8], aq[8];
        PQ_TMEM_LD8(r, tr + c0);
        PQ_TMEM_LD8(q, ti + c0);
        if (c > 0) {
          PQ_TMEM_LD8(ar, sr + c0);
          PQ_TMEM_LD8(aq, si + c0);
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        if (c > 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(ar[j]));
            q[j] = __float_as_uint(__uint_as_float(q[j]) + __uint_as_float(aq[j]));
          }
        }
        if (final) {
          // + the cross terms of the whole contraction (their last MMA precedes the commit
          // this fold waited for)
          PQ_TMEM_LD8(ar, tmem_x + lane_base + c0);
          PQ_TMEM_LD8(aq, tmem_x + lane_base + BN + c0);
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(ar[j]));
            q[j] = __float_as_uint(__uint_as_float(q[j]) + __uint_as_float(aq[j]));
          }
          if (m < M) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const long long n = n0 + c0 + j;
              if (n < N) C[m + M * n] = make_float2(__uint_as_float(r[j]), __uint_as_float(q[j]));
            }
          }
        } else {
          PQ_TMEM_ST8(sr + c0, r);
          PQ_TMEM_ST8(si + c0, q);
        }
      }
      if (!final) {
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        mbar_arrive(&drained[set]);
      }
    };

    // two register sets: the global loads of stages kt+1 and kt+2 are both in flight while
    // stage kt is converted (the loaders are latency-bound, so bytes in flight = throughput)
    LoadMap<TM> ma;
    LoadMap<BN> mb;
    ma.init(A, M, m0, M, tid);
    mb.init(B, N, n0, N, tid);
    LoadRegs<TM> la0, la1;
    LoadRegs<BN> lb0, lb1;
    la0.fetch(ma, 0, K);
    lb0.fetch(mb, 0, K);
    if (KT > 1) {
      la1.fetch(ma, 1, K);
      lb1.fetch(mb, 1, K);
    }
    auto stage_body = [&](LoadRegs<TM>& la, LoadRegs<BN>& lb, long long kt) {
      const int s = (int)(kt % NSTAGE);
      const long long chunk = kt / STAGES_PER_CHUNK;
      const int kin = (int)(kt - chunk * STAGES_PER_CHUNK);
      unsigned char* stage = smem + s * cfg::STAGE;
      if (kt >= NSTAGE) mbar_wait(&empty[s], (uint32_t)(((kt / NSTAGE) - 1) & 1));
      la.split_store(ma, stage, cfg::A_PLANE);
      lb.split_store(mb, stage + 4 * cfg::A_PLANE, cfg::B_PLANE);
      if (kt + 2 < KT) {  // refill this register set with the stage after next
        la.fetch(ma, kt + 2, K);
        lb.fetch(mb, kt + 2, K);
      }
      // generic-proxy writes -> visible to the tensor core; one arrival per warp
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
      // the previous chunk is complete (or about to be): fold it while the tensor core
      // works on this chunk in the other TMEM set
      if (kin == 0 && chunk >= 1 && warp < 4) fold(chunk - 1, false);
    };
    for (long long kt = 0; kt < KT; kt += 2) {
      stage_body(la0, lb0, kt);
      if (kt + 1 < KT) stage_body(la1, lb1, kt + 1);
    }
    if (warp < 4) fold((KT - 1) / STAGES_PER_CHUNK, true);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base),
                 "r"((uint32_t)cfg::TMEM_COLS)
                 : "memory");
  }
}

template <int BN>
void launch(const Launch& L, const void* A, const void* B, void* C, int64_t M, int64_t N,
            int64_t K) {
  using cfg = Cfg<BN>;
  long long tiles = ((M + TM - 1) / TM) * ((N + BN - 1) / BN);
  PQ_REQUIRE(tiles <= 0x7fffffffLL, PQ_ERR_UNSUPPORTED, "too many tiles for the tcgen05 CGEMM grid");
  unsigned grid = (unsigned)tiles;
  k_cgemm_tcgen05<BN><<<grid, NALL, cfg::SMEM, L.stream>>>((const float2*)A, (const float2*)B,
                                                              (float2*)C, M, N, K);
}

}  // namespace

void init_kernels_cgemm() {
  PQ_CUDA(cudaFuncSetAttribute(k_cgemm_tcgen05<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               Cfg<64>::SMEM));
  PQ_CUDA(cudaFuncSetAttribute(k_cgemm_tcgen05<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               Cfg<32>::SMEM));
}

// ComplexF32 GEMM on canonical TTGT layouts, tensor-core path
void run_cgemm_tcgen05(const Launch& L, const void* A, const void* B, void* C, int64_t M, int64_t N,
                       int64_t K) {
  double bytes = double(M * K + N * K + M * N) * 8.0, flops = 8.0 * M * N * K;
  L.begin(KC_GEMM_TENSOR, bytes, flops);
  if (N > 32)
    launch<64>(L, A, B, C, M, N, K);
  else
    launch<32>(L, A, B, C, M, N, K);
  L.end();
  PQ_CUDA(cudaGetLastError());
}

}  // namespace pq
