// K2t: second-generation INT8 tensor-core complex GEMM for the skinny sweep steps
// (K <= 64, N <= 64, gather fused) -- "transposed, K-concatenated, warp-specialised".
//
// The arithmetic (ozaki_math.h; executed on the host by test_lower.cpp: test_ozaki_t): an
// Ozaki-scheme product.  Every row of A (all k, re and im) gets one power-of-two scale, every
// column of B likewise; a real x of a row becomes q = rint(x 2^(46 - E)), written in balanced
// base 256 as six int8 digits (ComplexF32: four digits of a 30-bit q); products of digit planes
// with equal weight are summed EXACTLY by tcgen05.mma kind::i8 in one int32 accumulator group,
// pairs beyond the sixth (fourth) group are dropped (rel-L2 ~2.5e-13 / 2.7e-8 per contraction).
// A first kernel with this arithmetic ran its phases (gather, slice, MMA, TMEM drain) one after
// the other on 128-row tiles with A as the M-side operand and measured 282 us on the dominant
// step against 261 us for DMMA (profiles/ozaki_probe_r02_first.json; removed).  Here:
//
//  * B is the RESIDENT M-side operand: W = 128 rows, row 2n = [Br(.,n) | -Bi(.,n)], row 2n+1 =
//    [Bi(.,n) | Br(.,n)] (contraction length 2K: the four real products of a complex one are
//    ONE real GEMM).  A tile of 64 rows of A is the N-side operand X, row j = [Ar | Ai].
//    D = W X^T: TMEM lanes (2n, 2n+1) hold (Cr, Ci) of column n, TMEM columns the tile rows.
//    One MMA is 128 x 64 x 32; per tile and digit pair there are K/16 of them (21 pairs for
//    ComplexF64, 10 for ComplexF32) -- half the instruction count of the first kernel, whose
//    128 x 32 x 32 MMAs were bound by re-reading the 4 KB A operand from shared memory.
//  * A 64-row tile needs 48 KB of digit planes (32 KB ComplexF32), so X is DOUBLE-BUFFERED
//    next to the resident W (96 / 64 KB): slicing of tile t+1 overlaps the MMAs of tile t.
//  * Warp-specialised, coupled only by mbarriers:
//      8 producer warps   gather 64 rows x K of A (thread = row, 16-k chunk; the row exponent
//                         is a warp shuffle), slice, wait empty[stage], write planes,
//                         arrive full[stage];
//      1 MMA warp         wait full[stage]; per accumulator group g: wait freed[g], issue,
//                         tcgen05.commit -> done[g]; after the last group commit -> empty[stage];
//      8 epilogue warps   (two per TMEM lane quadrant, 32 tile rows each) wait done[g],
//                         tcgen05.ld.16x256b (re and im of two rows per thread), exact integer
//                         recombination (two int32 pair sums, one int64), conversion by magic
//                         numbers (no I2F), one 32-byte store per thread.
//    ComplexF32 has 4 groups of 64 columns: the accumulators are double-buffered as well.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "common.h"
#include "ozaki_math.h"

namespace pq {

namespace {

namespace ot = oz::ot;

constexpr int OT_PW = 8;                                // producer warps
constexpr int OT_EW = 8;                                // epilogue warps
constexpr int OT_THREADS = (OT_PW + OT_EW + 1) * 32;    // + the MMA warp
constexpr int OT_NST = 2;                               // X stages
constexpr int OT_SLOTS = 4;                             // row-scale ring (tiles in flight P -> E)
constexpr int OT_NBAR = 2 * OT_NST + 2 * 8 + 2 * 8 + 1;   // full, empty, done, freed, wready

template <class Real> struct OtVec;
template <> struct OtVec<double> { using type = double2; };
template <> struct OtVec<float> { using type = float2; };

template <int S>
struct OtSmem {
  static constexpr int kW = 0;
  static constexpr int kX = kW + S * ot::W_PLANE;
  static constexpr int kKoffA = kX + OT_NST * S * ot::X_PLANE;   // int[64]
  static constexpr int kKoffB = kKoffA + 64 * 4;                 // int[64]
  static constexpr int kRowS = kKoffB + 64 * 4;                  // double[OT_SLOTS][64]
  static constexpr int kColS = kRowS + OT_SLOTS * 64 * 8;        // double[64]
  static constexpr int kBars = kColS + 64 * 8;
  static constexpr int kTotal = kBars + OT_NBAR * 8 + 16;        // + tmem slot, abort flag
};
static_assert(OtSmem<6>::kTotal <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ uint32_t ot_smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void ot_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(ot_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void ot_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(ot_smem_u32(bar)) : "memory");
}
// Watchdog (as in the first kernel): a wait that does not complete within ~2 s records which
// one it was and raises a CTA-wide abort flag, so a protocol mistake ends as a finished kernel
// with a diagnosis (pq_microbench "ozaki_t_debug"), not as a hung GPU.
//   g_ot_debug = {flag, wait id, tile, group, block, warp, -, -}
//   wait ids: 1 full (MMA), 2 freed (MMA), 3 done (epilogue), 4 empty (producers), 5 wready (MMA)
__device__ int g_ot_debug[8];
__device__ long long g_ot_trace[32 * 32];   // block 0: [tile < 32][event]
__device__ long long g_ot_block_ns[2 * 160];   // TR: globaltimer at the start / end of every block
#define OT_TRACE(tile_no, ev)                                                                   \
  do {                                                                                          \
    if (TR && blockIdx.x == 0 && (tile_no) < 32) g_ot_trace[(tile_no) * 32 + (ev)] = clock64(); \
  } while (0)
constexpr long long OT_WAIT_LIMIT = 4000000000ll;

__device__ __forceinline__ bool ot_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(ot_smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __noinline__ void ot_wait_slow(uint64_t* bar, uint32_t parity, volatile int* abort_flag,
                                          int id, int it, int g) {
  const long long t0 = clock64();
  while (!ot_try_wait(bar, parity)) {
    if (*abort_flag) return;
    if (clock64() - t0 > OT_WAIT_LIMIT) {
      *abort_flag = 1;
      if (atomicCAS(&g_ot_debug[0], 0, 1) == 0) {
        g_ot_debug[1] = id;
        g_ot_debug[2] = it;
        g_ot_debug[3] = g;
        g_ot_debug[4] = (int)blockIdx.x;
        g_ot_debug[5] = (int)(threadIdx.x >> 5);
        __threadfence();
      }
      return;
    }
  }
}
__device__ __forceinline__ void ot_wait(uint64_t* bar, uint32_t parity, volatile int* abort_flag,
                                        int id, int it, int g) {
  if (!ot_try_wait(bar, parity)) ot_wait_slow(bar, parity, abort_flag, id, it, g);
}

// no-swizzle K-major shared-memory matrix descriptor (core matrix = 8 rows x 16 bytes; LBO =
// bytes between core matrices adjacent in K, SBO = between 8-row groups), version 1
__device__ __forceinline__ uint64_t ot_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// kind::i8 instruction descriptor: D = S32, A = B = signed int8, both K-major
__host__ __device__ constexpr uint32_t ot_idesc(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void ot_umma_i8(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                           uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "setp.ne.b32 p, %6, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand (the M side) from tensor memory: lane = row, 8 columns = the 32 int8 of one k-step
__device__ __forceinline__ void ot_umma_i8_ts_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc,
                                                    uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Whole-warp variants: every lane executes the (uniform) surrounding code, one elected lane
// issues.  With `if (lane == 0)` around the issue loop ptxas cannot prove uniformity: it wraps
// every UTCIMMA in an ELECT / BRA.U.ANY loop and moves each descriptor through R2UR, ~17
// instructions per MMA -- measured 115 clocks per MMA in the running kernel (the issuing warp
// shares its scheduler with four busy warps) against 48 for the tensor core itself.
__device__ __forceinline__ void ot_umma_i8_elect(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void ot_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(ot_smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void ot_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                   ot_smem_u32(bar))
               : "memory");
}
#define OT_TMEM_LD8(r, addr)                                                                  \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"    \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]),     \
                 "=r"(r[6]), "=r"(r[7])                                                       \
               : "r"(addr))

// 16 lanes x 8 columns as an accumulator fragment: thread t gets lane t/4 (r0, r1) and lane
// t/4 + 8 (r2, r3), columns 2 (t%4), 2 (t%4) + 1
#define OT_TMEM_LD_16x256(r, addr)                                               \
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];\n" \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])                 \
               : "r"(addr))
#define OT_TMEM_LD4(r, addr)                                                \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n" \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])               \
               : "r"(addr))

// Accumulator placement in the 512 TMEM columns (64 per group).  ComplexF32 (4 groups): every
// group double-buffered, 64 g + 256 buf.  ComplexF64 (6 groups, 384 columns): the two spare
// groups of columns double-buffer groups 4 and 5 -- the longest ones (20 + 24 of the 84 MMAs
// of a tile), issued FIRST, so that the tensor core has ~2100 clocks of work on the next tile
// while the epilogue still drains groups 0..3 of the current one.
// WT > 0 (ComplexF64 only): the spare 128 columns hold digit planes 0..WT-1 of W instead (8
// columns per plane and k-step, lane = W row), the A operand of the TS form of the MMA -- no
// double buffering then; groups in natural order.
template <int G, int WT>
__device__ __forceinline__ constexpr bool ot_dbuf(int g) { return G == 4 || (WT == 0 && g >= 4); }
template <int G, int WT>
__device__ __forceinline__ uint32_t ot_acc_col(int g, uint32_t buf) {
  return G == 4 ? (uint32_t)(64 * g) + 256u * buf
                : (g < 4 || WT > 0 ? (uint32_t)(64 * g) : 256u + 64u * (uint32_t)(g - 4) + 128u * buf);
}
template <int G, int WT>
__device__ __forceinline__ constexpr int ot_issue_order(int i) {
  return (G == 4 || WT > 0) ? i : (i == 0 ? 5 : i == 1 ? 4 : i - 2);
}
__device__ __forceinline__ constexpr uint32_t ot_w_col(int wp, int ks) { return 384u + (uint32_t)((wp * 4 + ks) * 8); }

// Output scaling, kept lean in ISSUE SLOTS and FP64 instructions (the epilogue's two scarce
// resources: ncu showed ~80 instructions per real number for an integer-only version, and the
// DADD / DMULs of the first versions waiting on the FP64 pipe for a third of the epilogue's
// time).  ComplexF64: both scales ride in the exponent field of the magic constant of the
// integer -> double conversion (ot::to_double51_scaled: ONE DADD, no multiplication).
// ComplexF32: int32 -> float conversions, one FFMA, two FMULs (row scale first).
template <class Real> struct OtScale;
template <> struct OtScale<double> {
  using row_type = int;            // pre-shifted exponent word (ot::row_word)
  using col_type = ot::DoubleMagic;
  static __device__ __forceinline__ int row_field(int ef) { return ot::field_clamp(ef); }
  static __device__ __forceinline__ row_type row(int ef) { return ot::row_word(ef); }
  template <int G>
  static __device__ __forceinline__ int col_field(int ef) { return ot::field_clamp(ef); }
  template <int G>
  static __device__ __forceinline__ col_type col_load(int ef) {   // ef: the clamped field from shared memory
    const bool dead = ef < oz::Traits<double>::MIN_EF || ef >= 2047;   // flushed (V = 0 anyway) / Inf, NaN
    return ot::double_magic(dead ? 0 : oz::Traits<double>::out_exp(ef) - 8 * (G - 1), ef >= 2047);
  }
  template <int G>
  static __device__ __forceinline__ double apply(const int* rr, const col_type& c, row_type sr) {
    static_assert(G == 6, "ComplexF64: six digits, six groups");
    return ot::to_double51_scaled(ot::combine51(rr), c, sr);
  }
};
template <> struct OtScale<float> {
  using row_type = float;
  using col_type = float;
  static __device__ __forceinline__ row_type row(int ef) { return oz::Traits<float>::out_scale_f(ef, 0); }
  static __device__ __forceinline__ int row_field(int ef) { return ef; }
  template <int G>
  static __device__ __forceinline__ int col_field(int ef) { return ef; }
  template <int G>
  static __device__ __forceinline__ col_type col_load(int ef) { return oz::Traits<float>::out_scale_f(ef, -8 * (G - 1)); }
  template <int G>
  static __device__ __forceinline__ float apply(const int* rr, col_type sc, row_type sr) {
    static_assert(G == 4, "ComplexF32: four digits, four groups");
    return (ot::combine_f32(rr) * sr) * sc;
  }
};

// 2 consecutive complex numbers (one lane's share of 4 rows of a column of C)
__device__ __forceinline__ void ot_store2(double2* dst, const double* re, const double* im, bool wide) {
  if (wide) {   // 32-byte aligned: one 256-bit store (STG.E.ENL2.256), a whole sector
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};\n" ::"l"(dst), "d"(re[0]), "d"(im[0]),
                 "d"(re[1]), "d"(im[1])
                 : "memory");
  } else {
    dst[0] = make_double2(re[0], im[0]);
    dst[1] = make_double2(re[1], im[1]);
  }
}
__device__ __forceinline__ void ot_store2(float2* dst, const float* re, const float* im, bool wide) {
  if (wide) {   // 16-byte aligned
    *reinterpret_cast<float4*>(dst) = make_float4(re[0], im[0], re[1], im[1]);
  } else {
    dst[0] = make_float2(re[0], im[0]);
    dst[1] = make_float2(re[1], im[1]);
  }
}

// 4 consecutive complex numbers (one lane's share of a column of C)
__device__ __forceinline__ void ot_store4(double2* dst, const double* re, const double* im, bool wide) {
  if (wide) {   // 32-byte aligned: two 256-bit stores (STG.E.ENL2.256), whole sectors
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};\n" ::"l"(dst), "d"(re[0]), "d"(im[0]),
                 "d"(re[1]), "d"(im[1])
                 : "memory");
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};\n" ::"l"(dst + 2), "d"(re[2]), "d"(im[2]),
                 "d"(re[3]), "d"(im[3])
                 : "memory");
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = make_double2(re[i], im[i]);
  }
}
__device__ __forceinline__ void ot_store4(float2* dst, const double* re, const double* im, bool wide) {
  if (wide) {   // 32-byte aligned
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(dst),
                 "f"((float)re[0]), "f"((float)im[0]), "f"((float)re[1]), "f"((float)im[1]),
                 "f"((float)re[2]), "f"((float)im[2]), "f"((float)re[3]), "f"((float)im[3])
                 : "memory");
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = make_float2((float)re[i], (float)im[i]);
  }
}

template <class Real, bool TR = false, int WT = 0>
__global__ void __maxnreg__(96)
k_ozaki_t(const typename OtVec<Real>::type* __restrict__ A, const typename OtVec<Real>::type* __restrict__ B,
          typename OtVec<Real>::type* __restrict__ C, const FusedParams p) {
  using Tr = oz::Traits<Real>;
  using V2 = typename OtVec<Real>::type;
  using Sm = OtSmem<Tr::S>;
  constexpr int S = Tr::S, G = Tr::S;
  static_assert(G == 4 || G == 6, "accumulator placement (ot_acc_col) is written for 4 or 6 groups");
  static_assert(WT == 0 || (G == 6 && WT <= 4), "W planes in tensor memory: ComplexF64 only, at most 4 planes (128 columns)");
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sW = smem + Sm::kW;
  unsigned char* sX = smem + Sm::kX;
  int* koffA = reinterpret_cast<int*>(smem + Sm::kKoffA);
  int* koffB = reinterpret_cast<int*>(smem + Sm::kKoffB);
  using Sc = OtScale<Real>;
  using RowT = typename Sc::row_type;
  RowT* rowS = reinterpret_cast<RowT*>(smem + Sm::kRowS);   // [OT_SLOTS][64] row scales
  int* colS = reinterpret_cast<int*>(smem + Sm::kColS);      // [64] column exponent fields
  // full[s]: the producers have written stage s (count 8).  empty[s]: the MMAs reading stage s
  // have completed (commit).  done[b][g]: the MMAs of group g into accumulator buffer b have
  // completed (commit).  freed[b][g]: every epilogue warp has read it out of TMEM (count 8).
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Sm::kBars);
  uint64_t* empty = full + OT_NST;
  uint64_t* done = empty + OT_NST;
  uint64_t* freed = done + 16;
  uint64_t* wready = freed + 16;   // W (and its copy in tensor memory) is complete (count 8: the epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wready + 1);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  // (the shuffle makes the warp index provably warp-uniform: the role branches below are then
  // convergent for ptxas, and the MMA warp's descriptor arithmetic runs on the uniform datapath)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  if (TR && tid == 0 && blockIdx.x < 160) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    g_ot_block_ns[2 * blockIdx.x] = (long long)ns;
  }
  const int K = (int)p.K, N = (int)p.N;
  const long long M = p.M;
  const int KC = (K + 15) / 16;   // 16-k chunks per half = MMA k-steps
  const long long tiles = (M + ot::ROWS - 1) / ot::ROWS;

  if (tid < 64) {
    koffA[tid] = tid < K ? (int)map_offset(p.kA, tid) : 0;
    koffB[tid] = tid < K ? (int)map_offset(p.kB, tid) : 0;
  }
  if (tid == 0) {
    *abort_flag = 0;
    for (int s = 0; s < OT_NST; ++s) {
      ot_mbar_init(&full[s], OT_PW);
      ot_mbar_init(&empty[s], 1);
    }
    for (int g = 0; g < 16; ++g) {
      ot_mbar_init(&done[g], 1);
      ot_mbar_init(&freed[g], OT_EW);
    }
    ot_mbar_init(wready, OT_EW);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     ot_smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // ---- W: B gathered, scaled per column, sliced, once per CTA -- by the EPILOGUE warps, which
  // have nothing to drain yet, while the producers already fetch and slice the first tile (the
  // set-up used to cost every launch 3-4 us in front of the pipeline; a slice has 25 launches).
  // The MMA warp waits for `wready` before its first MMA; the epilogue warps meet at a named
  // barrier of their own (colS is theirs to read).
  if (warp >= OT_PW && warp < OT_PW + OT_EW) {
    const int ew = warp - OT_PW;
    const int n = ew * 8 + (lane & 7), c = lane >> 3;
    const bool on = c < KC;
    const long long rb = n < N ? map_offset(p.nB, n) : -1;
    Real xr[16], xi[16];
    int key = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int k = c * 16 + i;
      V2 v;
      v.x = v.y = (Real)0;
      if (on && rb >= 0 && k < K) v = B[rb + koffB[k]];
      xr[i] = v.x;
      xi[i] = v.y;
      key = max(key, max(Tr::key(v.x), Tr::key(v.y)));
    }
    key = max(key, __shfl_xor_sync(0xffffffffu, key, 8));
    key = max(key, __shfl_xor_sync(0xffffffffu, key, 16));
    const int eb = Sc::template col_field<G>(Tr::exp_field(key));
    if (on) ot::w_item<Real>(sW, n, c, KC, xr, xi, Tr::slice_scale(eb));
    if (c == 0) colS[n] = eb;
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    if constexpr (WT > 0) {
      // digit planes 0..WT-1 of W into tensor memory: lane = W row, 8 columns = the 32 digits of
      // one k-step (two 16-byte chunks of the row).  Epilogue warp w < 4 writes the lanes of
      // quadrant w once every epilogue warp has stored its rows.
      asm volatile("bar.sync 1, %0;\n" ::"n"(OT_EW * 32) : "memory");
      if (ew < 4) {
        const int r = 32 * ew + lane;
#pragma unroll
        for (int wp = 0; wp < WT; ++wp)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint4 lo = *reinterpret_cast<const uint4*>(sW + wp * ot::W_PLANE + oz::plane_off(ot::WROWS, r, 2 * ks));
            const uint4 hi = *reinterpret_cast<const uint4*>(sW + wp * ot::W_PLANE + oz::plane_off(ot::WROWS, r, 2 * ks + 1));
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(
                             tmem_base + ((uint32_t)(32 * ew) << 16) + ot_w_col(wp, ks)),
                         "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
                         : "memory");
          }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    }
    asm volatile("bar.sync 1, %0;\n" ::"n"(OT_EW * 32) : "memory");
    if (lane == 0) ot_mbar_arrive(wready);
  }

  if (warp < OT_PW) {
    // ===================== producers =====================
    // thread = (tile row j, 16-k chunk c).  The gather of a tile is one batch of 16 loads per
    // thread; the row offset of the NEXT tile is computed behind the issued loads (its chain of
    // constant-bank reads cost ~1000 clocks per tile at the top of the loop).  ComplexF32
    // (32 data registers per batch) keeps TWO batches in flight: the loads of tile t + 1 are
    // issued before tile t is sliced.  ComplexF64 (64 registers) has room for one; an L2
    // prefetch of the next tile was tried and changed nothing (pq: PQ_OT_PF experiment).
    const int j = warp * 8 + (lane & 7), c = lane >> 3;
    const bool on = c < KC;
    constexpr bool TWO = sizeof(Real) == 4;
    auto row_offset = [&](long long tile) -> long long {
      const long long m = tile * ot::ROWS + j;
      return (tile < tiles && m < M) ? map_offset(p.mA, m) : -1;
    };
    auto issue = [&](Real (&xr)[16], Real (&xi)[16], long long ra) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int k = c * 16 + i;
        V2 v;
        v.x = v.y = (Real)0;
        if (on && ra >= 0 && k < K) v = A[ra + koffA[k]];
        xr[i] = v.x;
        xi[i] = v.y;
      }
    };
    auto finish = [&](const Real (&xr)[16], const Real (&xi)[16], uint32_t t) {
      const uint32_t stage = t & 1u, use = t >> 1;
      int key = 0;
#pragma unroll
      for (int i = 0; i < 16; ++i) key = max(key, max(Tr::key(xr[i]), Tr::key(xi[i])));
      key = max(key, __shfl_xor_sync(0xffffffffu, key, 8));
      key = max(key, __shfl_xor_sync(0xffffffffu, key, 16));
      const int ea = Sc::row_field(Tr::exp_field(key));
      if (tid == 0) OT_TRACE(t, 1);
      // the MMAs that read this stage two tiles ago must have completed
      if (use > 0) ot_wait(&empty[stage], (use - 1) & 1u, abort_flag, 4, (int)t, -1);
      if (tid == 0) OT_TRACE(t, 2);
      if (on) ot::x_item<Real>(sX + stage * (S * ot::X_PLANE), j, c, KC, xr, xi, Tr::slice_scale(ea));
      if (c == 0) rowS[(t & (OT_SLOTS - 1)) * ot::ROWS + j] = Sc::row(ea);
      // generic-proxy stores -> visible to the tensor core
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) ot_mbar_arrive(&full[stage]);
      if (tid == 0) OT_TRACE(t, 3);
    };
    const long long step = gridDim.x;
    long long tile = blockIdx.x;
    uint32_t t = 0;
    Real xa_r[16], xa_i[16];
    if constexpr (TWO) {
      Real xb_r[16], xb_i[16];
      long long ra = row_offset(tile);
      issue(xa_r, xa_i, ra);
      ra = row_offset(tile + step);
      while (tile < tiles) {
        if (tid == 0) OT_TRACE(t, 0);
        issue(xb_r, xb_i, ra);               // tile + step (zeros beyond the last tile)
        ra = row_offset(tile + 2 * step);
        finish(xa_r, xa_i, t);
        tile += step, ++t;
        if (tile >= tiles) break;
        if (tid == 0) OT_TRACE(t, 0);
        issue(xa_r, xa_i, ra);
        ra = row_offset(tile + 2 * step);
        finish(xb_r, xb_i, t);
        tile += step, ++t;
      }
    } else {
      long long ra = row_offset(tile);
      for (; tile < tiles; tile += step, ++t) {
        if (tid == 0) OT_TRACE(t, 0);
        issue(xa_r, xa_i, ra);
        ra = row_offset(tile + step);        // behind the loads of this tile
        finish(xa_r, xa_i, t);
      }
    }
  } else if (warp == OT_PW + OT_EW) {
    // ===================== MMA issuer =====================
    // (all 32 lanes run this uniform loop; one elected lane issues, see ot_umma_i8_elect)
    constexpr uint32_t IDESC = ot_idesc(ot::WROWS, ot::ROWS);
    const uint64_t w_base = ot_desc(ot_smem_u32(sW), ot::W_LBO, ot::SBO);
    const uint64_t x_base = ot_desc(ot_smem_u32(sX), ot::X_LBO, ot::SBO);
    uint32_t t = 0;
    ot_wait(wready, 0u, abort_flag, 5, 0, -1);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++t) {
      const uint32_t stage = t & 1u;
      ot_wait(&full[stage], (t >> 1) & 1u, abort_flag, 1, (int)t, -1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      if (lane == 0) OT_TRACE(t, 8);
      const uint64_t xs = x_base + (uint64_t)((stage * (S * ot::X_PLANE)) >> 4);
#pragma unroll
      for (int gi = 0; gi < G; ++gi) {
        const int g = ot_issue_order<G, WT>(gi);
        const uint32_t buf = ot_dbuf<G, WT>(g) ? (t & 1u) : 0u, u = ot_dbuf<G, WT>(g) ? (t >> 1) : t;
        // the epilogue must have drained this accumulator (its previous use)
        if (u > 0) {
          ot_wait(&freed[buf * 8 + g], (u - 1) & 1u, abort_flag, 2, (int)t, g);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        }
        const uint32_t acc_addr = tmem_base + ot_acc_col<G, WT>(g, buf);
        uint32_t acc = 0;
#pragma unroll
        for (int sp = 0; sp < S; ++sp) {
          const int wp = g - sp;   // digit pair (wp of W, sp of X), wp + sp = g
          if (wp < 0 || wp >= S) continue;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (ks < KC) {
              if (WT > 0 && wp < WT)
                ot_umma_i8_ts_elect(acc_addr, tmem_base + ot_w_col(wp, ks),
                                    xs + (uint64_t)((sp * ot::X_PLANE + ks * 2 * ot::X_LBO) >> 4), IDESC, acc);
              else
                ot_umma_i8_elect(acc_addr, w_base + (uint64_t)((wp * ot::W_PLANE + ks * 2 * ot::W_LBO) >> 4),
                                 xs + (uint64_t)((sp * ot::X_PLANE + ks * 2 * ot::X_LBO) >> 4), IDESC, acc);
              acc = 1u;
            }
          }
        }
        ot_commit_elect(&done[buf * 8 + g]);
        if (lane == 0) OT_TRACE(t, 9 + gi);
      }
      ot_commit_elect(&empty[stage]);   // the stage may be rewritten once these MMAs have completed
    }
  } else {
    // ===================== epilogue =====================
    // Warp (quadrant q, half ch) drains lanes 32 q .. 32 q + 31 x tile rows 32 ch .. 32 ch + 31
    // of every group, in 8 blocks of 16 lanes x 8 rows (tcgen05.ld.16x256b): thread t of a
    // block owns column n = 16 q + 8 h + t / 4 of C and rows 2 (t % 4), 2 (t % 4) + 1 of the
    // block, re from lane t / 4, im from lane t / 4 + 8 -- two whole complex numbers, stored
    // as one 32-byte (ComplexF32: 16-byte) word; four threads cover 128 contiguous bytes.
    const int q = warp & 3;                 // TMEM lane quadrant of this warp
    const int ch = (warp - OT_PW) >> 2;     // which 32 of the 64 tile rows
    const int jt = 2 * (lane & 3);          // this thread's row pair inside a block
    typename Sc::col_type sb[2];
    V2* cbase[2];   // column n of C (null: n beyond N)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = 16 * q + 8 * h + (lane >> 2);
      sb[h] = Sc::template col_load<G>(colS[n]);
      cbase[h] = n < N ? C + M * (long long)n : nullptr;
    }
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * q) << 16);
    const bool wide = (M & 1) == 0 && (reinterpret_cast<uintptr_t>(C) & 31) == 0;
    uint32_t t = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++t) {
      const RowT* rs = rowS + (t & (OT_SLOTS - 1)) * ot::ROWS;
      // block b: lanes half h = b & 1, rows 32 ch + 8 (b >> 1) ..; software-pipelined over two
      // register sets (the loads of block b + 1 are in flight while block b is recombined)
      uint32_t ra[G][4], rb[G][4];
      auto load = [&](uint32_t (&r)[G][4], int b, bool first) {
        const uint32_t col0 = (uint32_t)(32 * ch + 8 * (b >> 1));
        const uint32_t la = lane_addr + ((uint32_t)(16 * (b & 1)) << 16) + col0;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const uint32_t buf = ot_dbuf<G, WT>(g) ? (t & 1u) : 0u, u = ot_dbuf<G, WT>(g) ? (t >> 1) : t;
          if (first) {
            ot_wait(&done[buf * 8 + g], u & 1u, abort_flag, 3, (int)t, g);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            if (tid == OT_PW * 32 && (g == 0 || g == G - 1)) OT_TRACE(t, g == 0 ? 16 : 17);
          }
          OT_TMEM_LD_16x256(r[g], la + (ot_acc_col<G, WT>(g, buf)));
        }
      };
      auto wait_ld = [&]() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); };
      // full tiles of an even-M, 32-byte aligned C (the sweep steps): one unchecked wide store
      const bool fast = wide && (tile + 1) * ot::ROWS <= M;
      auto process = [&](const uint32_t (&r)[G][4], int h, int cg) {   // lanes half h, 8-row group cg
        const int j0 = 32 * ch + 8 * cg + jt;
        Real v[4];   // re(m), re(m + 1), im(m), im(m + 1)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int rr[G];
#pragma unroll
          for (int g = 0; g < G; ++g) rr[g] = (int)r[g][i];
          v[i] = Sc::template apply<G>(rr, sb[h], rs[j0 + (i & 1)]);
        }
        const Real re[2] = {v[0], v[1]}, im[2] = {v[2], v[3]};
        const long long m0 = tile * ot::ROWS + j0;
        if (fast) {
          if (cbase[h] != nullptr) ot_store2(cbase[h] + m0, re, im, true);
        } else if (cbase[h] != nullptr && m0 < M) {
          V2* dst = cbase[h] + m0;
          if (m0 + 1 < M) {
            ot_store2(dst, re, im, wide);
          } else {
            V2 o;
            o.x = re[0];
            o.y = im[0];
            dst[0] = o;
          }
        }
      };
      load(ra, 0, true);
      wait_ld();
#pragma unroll 1
      for (int b = 0; b < 8; b += 2) {
        load(rb, b + 1, false);
        if (b + 2 == 8) {
          // the last loads of the tile: release the accumulators BEFORE the last two blocks are
          // recombined, so that the MMAs of the next tile's single-buffered groups (1900 clocks)
          // run under that work instead of after it
          wait_ld();
          asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
          __syncwarp();
          if (lane == 0) {
#pragma unroll
            for (int g = 0; g < G; ++g)
              ot_mbar_arrive(&freed[(ot_dbuf<G, WT>(g) ? (t & 1u) : 0u) * 8 + g]);
          }
          process(ra, 0, b >> 1);
          process(rb, 1, b >> 1);
        } else {
          process(ra, 0, b >> 1);
          wait_ld();
          load(ra, b + 2, false);
          process(rb, 1, b >> 1);
          wait_ld();
        }
      }
      if (tid == OT_PW * 32) OT_TRACE(t, 18);
      if (TR && tid == OT_PW * 32 && blockIdx.x == 0 && t < 32) {   // wall clock beside the SM clock
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        g_ot_trace[t * 32 + 19] = (long long)ns;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (TR && tid == 0 && blockIdx.x < 160) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    g_ot_block_ns[2 * blockIdx.x + 1] = (long long)ns;
  }
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u)
                 : "memory");
  }
}


// ---------------------------------------------------------------------------
// Issue / execution rate probe of the kernel's own MMA stream (pq_microbench
// "ozaki_t_rate_<mode>"): every SM issues `iters` tiles worth of the ComplexF64 schedule
// (21 digit pairs x 4 k-steps of 128 x 64 x 32 into 6 accumulators) on resident planes and
// reports SM clocks per MMA.  mode bits: 1 = every MMA reads the SAME W / X slices (operand
// reuse), 2 = operand-major order (consecutive MMAs rotate over the accumulators instead of
// chaining into one), 4 = the 16 idle warps poll the final mbarrier with one lane per warp
// instead of all 32, 8 = no idle warps at all (they exit), 16 = N = 128 tiles (two X planes side by side),
// 32 / 64 / 128 = an arithmetic loop beside the MMA stream, 256 = the W operand from TENSOR memory
// (TS form; accumulators 0..3 only, W in columns 384..511), 512 = the eight producer warps stream
// 64 KB batches from global memory beside the MMA stream like the kernel's producers (16 loads of
// 16 bytes per thread, one batch in flight; their clocks per batch in out[148 + block]),
// 1024 = no MMAs at all (the load stream alone), 2048 = N = 32: two half-tile MMAs per X plane and
// k-step (the reported clocks are per PAIR, i.e. per 64 rows as in the other modes).  TS and N = 32
// are template parameters: a run-time branch around the MMA costs the uniform-datapath issue path.
// ---------------------------------------------------------------------------
template <bool TS, bool NARROW>
__global__ void __launch_bounds__(OT_THREADS, 1) k_ot_mma_rate(int iters, int mode, long long* __restrict__ out,
                                                               const double2* __restrict__ stream_buf,
                                                               long long stream_elems) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ int abort_flag;
  constexpr int S = 6;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // uniform: the MMA loop issues on the uniform datapath
  unsigned char* sW = smem;
  unsigned char* sX = smem + S * ot::W_PLANE;
  for (int i = tid; i < (S * ot::W_PLANE + 2 * S * ot::X_PLANE) / 4; i += OT_THREADS)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (tid == 0) {
    abort_flag = 0;
    ot_mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(ot_smem_u32(&slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = slot;
  const bool same = mode & 1, opmajor = mode & 2, one_lane = mode & 4, no_idle = mode & 8, wide = mode & 16;
  const bool no_mma = mode & 1024;
  constexpr bool narrow = NARROW;   // mode 2048: N = 32 MMAs (half tiles), two per X plane and k-step
  long long t0 = 0;
  if (warp == OT_PW + OT_EW) {
    if (no_mma) {
      t0 = clock64();
      if (lane == 0) ot_mbar_arrive(&bar);
    } else {   // whole warp, one elected lane issues (the kernel's own issue path)
      const uint32_t IDESC = wide ? ot_idesc(ot::WROWS, 128) : narrow ? ot_idesc(ot::WROWS, 32) : ot_idesc(ot::WROWS, ot::ROWS);
      const uint64_t w_base = ot_desc(ot_smem_u32(sW), ot::W_LBO, ot::SBO);
      const uint64_t x_base = ot_desc(ot_smem_u32(sX), wide ? 2 * ot::X_LBO : ot::X_LBO, ot::SBO);
      const int ncol = wide ? 128 : 64;
      auto mma = [&](int g, int wp, int xp, int ks, uint32_t acc) {
        if (same) wp = xp = ks = 0;
        if (wide) xp >>= 1;
        if (TS) {
          if constexpr (NARROW) {   // two half-tile MMAs (rows 0..31 / 32..63 of the plane) into two 192-column accumulator sets
#pragma unroll
            for (int h = 0; h < 2; ++h)
              ot_umma_i8_ts_elect(tmem + (uint32_t)(h * 192 + g * 32), tmem + 384u + (uint32_t)((wp & 3) * 32 + ks * 8),
                                  x_base + (uint64_t)((xp * ot::X_PLANE + ks * 2 * ot::X_LBO + h * 512) >> 4), IDESC, acc);
            return;
          }
          ot_umma_i8_ts_elect(tmem + (uint32_t)((g & 3) * ncol), tmem + 384u + (uint32_t)((wp & 3) * 32 + ks * 8),
                              x_base + (uint64_t)((xp * ot::X_PLANE + ks * 2 * ot::X_LBO) >> 4), IDESC, acc);
          return;
        }
        if constexpr (NARROW) {
#pragma unroll
          for (int h = 0; h < 2; ++h)
            ot_umma_i8_elect(tmem + (uint32_t)(h * 192 + g * 32),
                             w_base + (uint64_t)((wp * ot::W_PLANE + ks * 2 * ot::W_LBO) >> 4),
                             x_base + (uint64_t)((xp * ot::X_PLANE + ks * 2 * ot::X_LBO + h * 512) >> 4), IDESC, acc);
          return;
        }
        ot_umma_i8_elect(tmem + (uint32_t)((wide ? (g & 3) : g) * ncol),
                         w_base + (uint64_t)((wp * ot::W_PLANE + ks * 2 * ot::W_LBO) >> 4),
                         x_base + (uint64_t)((xp * (wide ? 2 : 1) * ot::X_PLANE + ks * 2 * (wide ? 2 : 1) * ot::X_LBO) >> 4),
                         IDESC, acc);
      };
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        if (!opmajor) {
#pragma unroll
          for (int g = 0; g < S; ++g)
            ot::for_each_mma_of_group<S>(g, 4, [&](int wp, int xp, int ks, uint32_t acc) { mma(g, wp, xp, ks, acc | (it > 0)); });
        } else {
#pragma unroll
          for (int wp = 0; wp < S; ++wp)
            for (int ks = 0; ks < 4; ++ks)
#pragma unroll
              for (int xp = 0; xp + wp < S; ++xp) mma(wp + xp, wp, xp, ks, (uint32_t)(it > 0 || ks > 0 || wp > 0));
        }
      }
      ot_commit_elect(&bar);
    }
  } else if ((mode & 512) && warp < OT_PW) {
    // the producers' load stream beside the MMA stream
    const long long a0 = clock64();
    double acc = 0;
    for (int it = 0; it < iters; ++it) {
      const long long base = (((long long)it * gridDim.x + blockIdx.x) * 4096) % (stream_elems - 4096);
      double2 v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = stream_buf[base + i * 256 + tid];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc += v[i].x + v[i].y;
    }
    if (acc == 12345.0) out[400] = 1;
    if (tid == 0) out[148 + blockIdx.x] = clock64() - a0;
  } else if (mode & (32 | 64 | 128)) {
    // contention probe: the other 16 warps run an arithmetic loop beside the MMA stream
    // (32: DFMA, 64: FFMA, 128: IMAD; 8 independent chains, 4096 iterations) and report their
    // own duration in out[148 + blockIdx.x]
    const long long a0 = clock64();
    if (mode & 32) {
      double x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = 1.0 + tid * 1e-9 + i;
      for (int it = 0; it < 4096; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fma(x[i], 1.0000001, 1e-9);
      }
      double sum = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) sum += x[i];
      if (sum == 12345.0) out[400] = 1;
    } else if (mode & 64) {
      float x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = 1.0f + tid * 1e-6f + i;
      for (int it = 0; it < 4096; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], 1.0000001f, 1e-9f);
      }
      float sum = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) sum += x[i];
      if (sum == 12345.0f) out[400] = 1;
    } else {
      int x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = tid + i;
      for (int it = 0; it < 4096; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = x[i] * 1664525 + 1013904223;
      }
      int sum = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) sum += x[i];
      if (sum == 12345) out[400] = 1;
    }
    if (tid == 0) out[148 + blockIdx.x] = clock64() - a0;
  }
  if (!(no_idle && warp != OT_PW + OT_EW)) {
    if (one_lane) {
      if (lane == 0) ot_wait(&bar, 0u, &abort_flag, 9, 0, -1);
      __syncwarp();
    } else {
      ot_wait(&bar, 0u, &abort_flag, 9, 0, -1);
    }
  }
  if (warp == OT_PW + OT_EW && lane == 0) out[blockIdx.x] = clock64() - t0;
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u) : "memory");
}


// TMEM read rate probe (pq_microbench "ozaki_t_ldtm_<W>_<warps>"): `warps` warps (4, 8 or 16:
// 1, 2 or 4 per lane quadrant) each issue 256 tcgen05.ld.32x32b.x<W> back to back (one
// wait::ld per load); returns bytes per SM clock.
template <int W>
__device__ __forceinline__ void ot_ldtm_probe(uint32_t addr, uint32_t& sink) {
  uint32_t r[W];
  if constexpr (W == 4) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
  } else if constexpr (W == 8) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(addr));
  } else if constexpr (W == 16) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(addr));
  } else {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(addr));
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < W; ++i) sink ^= r[i];
}
template <int W>
__global__ void __launch_bounds__(512, 1) k_ot_ldtm_rate(long long* __restrict__ out) {
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(ot_smem_u32(&slot)),
                 "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t sink = 0;
  const long long t0 = clock64();
#pragma unroll 4
  for (int it = 0; it < 256; ++it) ot_ldtm_probe<W>(tmem + (uint32_t)(((it + warp) * W) & (512 - W)), sink);
  const long long t1 = clock64();
  __syncthreads();
  if ((tid & 31) == 0) out[blockIdx.x * 16 + warp] = (t1 - t0) + (sink == 0x12345678u ? 1 : 0);
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(slot), "r"(512u) : "memory");
}

constexpr int OT_SELF_M = 128, OT_SELF_N = 64;                       // probe tiles: 128 x 64 (k) and 64 x 64 (k) int8 planes
constexpr int OT_SELF_A_PLANE = OT_SELF_M * 64, OT_SELF_B_PLANE = OT_SELF_N * 64;
// ---------------------------------------------------------------------------
// Bring-up aids (pq_microbench "umma_i8_selftest", "umma_i8_tops_n32", "umma_i8_tops_n64").
//
// Self-test: ONE 128 x 32 x 32 kind::i8 MMA on known int8 patterns laid out exactly like the
// kernel's planes (plane_off, LBO = rows * 16, SBO = 128), read back with tcgen05.ld and
// compared on the host with the integer dot products -- isolates the descriptor encodings
// and the TMEM lane / column mapping from everything else.  Returns the number of wrong
// entries (0 = pass).
// ---------------------------------------------------------------------------
__host__ __device__ inline int ot_pat_a(int r, int k) { return (r * 7 + k * 3 + r / 64) % 127 - 63; }
__host__ __device__ inline int ot_pat_b(int c, int k) { return (c * 5 + k * 11 + 1) % 127 - 63; }

__global__ void __launch_bounds__(128, 1) k_umma_i8_selftest(int* __restrict__ out) {
  __shared__ __align__(1024) unsigned char sa[OT_SELF_M * 32];
  __shared__ __align__(1024) unsigned char sb[32 * 32];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ int abort_flag;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < OT_SELF_M * 32; i += 128) {
    const int r = i >> 5, k = i & 31;
    sa[oz::plane_off(OT_SELF_M, r, k >> 4) + (k & 15)] = (unsigned char)(signed char)ot_pat_a(r, k);
  }
  for (int i = tid; i < 32 * 32; i += 128) {
    const int c = i >> 5, k = i & 31;
    sb[oz::plane_off(32, c, k >> 4) + (k & 15)] = (unsigned char)(signed char)ot_pat_b(c, k);
  }
  if (tid == 0) {
    abort_flag = 0;
    ot_mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     ot_smem_u32(&slot)),
                 "r"(32u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = slot;
  if (tid == 0) {
    const uint64_t ad = ot_desc(ot_smem_u32(sa), OT_SELF_M * 16, 128);
    const uint64_t bd = ot_desc(ot_smem_u32(sb), 32 * 16, 128);
    ot_umma_i8(tmem, (uint32_t)ad, (uint32_t)(ad >> 32), (uint32_t)bd, (uint32_t)(bd >> 32),
               ot_idesc(OT_SELF_M, 32), 0u);
    ot_commit(&bar);
  }
  ot_wait(&bar, 0u, &abort_flag, 4, 0, -1);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
#pragma unroll
  for (int c0 = 0; c0 < 32; c0 += 8) {
    uint32_t r[8];
    OT_TMEM_LD8(r, tmem + lane_base + c0);
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int j = 0; j < 8; ++j) out[(warp * 32 + lane) * 32 + c0 + j] = (int)r[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(32u)
                 : "memory");
}

// Issue-rate probe: every SM issues `iters` x 16 MMAs of 128 x NCOL x 32 on resident planes
// (4 accumulators, values irrelevant).  Returns int8 TOPS (2 ops per MAC).
template <int NCOL>
__global__ void __launch_bounds__(128, 1) k_umma_i8_rate(int iters, int* __restrict__ sink) {
  extern __shared__ __align__(1024) unsigned char smem[];   // A: 128 x 64, B: 64 x 64, zeroed
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ int abort_flag;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (OT_SELF_A_PLANE + OT_SELF_B_PLANE) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (tid == 0) {
    abort_flag = 0;
    ot_mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     ot_smem_u32(&slot)),
                 "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = slot;
  if (tid == 0) {
    const uint64_t ad = ot_desc(ot_smem_u32(smem), OT_SELF_M * 16, 128);
    const uint64_t bd = ot_desc(ot_smem_u32(smem + OT_SELF_A_PLANE), OT_SELF_N * 16, 128);
    constexpr uint32_t IDESC = ot_idesc(OT_SELF_M, NCOL);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t ks = (uint32_t)(j & 1) * ((2 * OT_SELF_M * 16) >> 4);
        const uint32_t kb = (uint32_t)(j & 1) * ((2 * OT_SELF_N * 16) >> 4);
        ot_umma_i8(tmem + (uint32_t)((j >> 2) * NCOL), (uint32_t)ad + ks, (uint32_t)(ad >> 32),
                   (uint32_t)bd + kb, (uint32_t)(bd >> 32), IDESC, (it | (j & 3)) ? 1u : 0u);
      }
    }
    ot_commit(&bar);
  }
  ot_wait(&bar, 0u, &abort_flag, 4, 0, -1);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  uint32_t r[8];
  OT_TMEM_LD8(r, tmem + ((uint32_t)(warp * 32) << 16));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
  if (r[0] == 0x7fffffffu) sink[blockIdx.x] = (int)r[1];   // keeps the loads alive
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(256u)
                 : "memory");
}


}  // namespace

static bool g_ozaki_t_ready = false;
static bool g_ozaki_t_launched = false;   // any k_ozaki_t launch in this process (see ozaki_t_check_watchdog)

void init_kernels_ozaki_t() {
  cudaError_t e[5];
  e[4] = cudaFuncSetAttribute(k_ozaki_t<double, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              OtSmem<6>::kTotal);
  e[0] = cudaFuncSetAttribute(k_ozaki_t<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, OtSmem<6>::kTotal);
  e[1] = cudaFuncSetAttribute(k_ozaki_t<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, OtSmem<4>::kTotal);
  e[2] = cudaFuncSetAttribute(k_ozaki_t<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              OtSmem<6>::kTotal);
  e[3] = cudaFuncSetAttribute(k_ozaki_t<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              OtSmem<4>::kTotal);
  g_ozaki_t_ready = true;
  for (cudaError_t x : e)
    if (x != cudaSuccess) {
      (void)cudaGetLastError();
      g_ozaki_t_ready = false;
    }
}

// ComplexF64 (6 digits, 6 accumulator groups) or ComplexF32 (4 / 4) by L.elem_size.  The caller
// brackets the launch with L.begin / L.end.
void run_zgemm_ozaki_t(const Launch& L, const FusedParams& fp, const void* A, const void* B, void* C) {
  PQ_REQUIRE(g_ozaki_t_ready, PQ_ERR_UNSUPPORTED, "ozaki GEMM: kernel attributes could not be set");
  PQ_REQUIRE(zgemm_ozaki_eligible(fp.M, fp.N, fp.K), PQ_ERR_INVALID, "ozaki GEMM: K, N <= 64 only");
  const long long tiles = (fp.M + ot::ROWS - 1) / ot::ROWS;
  const unsigned grid = (unsigned)(tiles < L.num_sms ? tiles : L.num_sms);
  const bool trace = std::getenv("PQ_OZAKI_TRACE") != nullptr;
  const FusedParams& fq = fp;
  g_ozaki_t_launched = true;
  if (L.elem_size == 16) {
    if (trace)
      k_ozaki_t<double, true><<<grid, OT_THREADS, OtSmem<6>::kTotal, L.stream>>>(
          (const double2*)A, (const double2*)B, (double2*)C, fq);
    else if (L.opt && L.opt->ozaki_tsw == 2)   // W planes 0..3 in tensor memory (TS form of the MMA)
      k_ozaki_t<double, false, 4><<<grid, OT_THREADS, OtSmem<6>::kTotal, L.stream>>>(
          (const double2*)A, (const double2*)B, (double2*)C, fq);
    else
      k_ozaki_t<double><<<grid, OT_THREADS, OtSmem<6>::kTotal, L.stream>>>(
          (const double2*)A, (const double2*)B, (double2*)C, fq);
  } else {
    if (trace)
      k_ozaki_t<float, true><<<grid, OT_THREADS, OtSmem<4>::kTotal, L.stream>>>(
          (const float2*)A, (const float2*)B, (float2*)C, fq);
    else
      k_ozaki_t<float><<<grid, OT_THREADS, OtSmem<4>::kTotal, L.stream>>>(
          (const float2*)A, (const float2*)B, (float2*)C, fq);
  }
}

// ComplexF32 contraction with the gather fused (plans lowered with fused_gemm, see lower.cpp)
void run_cgemm_ozaki_fused(const Launch& L, const ContractPlan& cp, const void* A, const void* B,
                           void* C) {
  PQ_REQUIRE(L.elem_size == 8, PQ_ERR_INVALID, "fused ComplexF32 GEMM plans run on the INT8 kernel only");
  FusedParams fp{};
  fp.mA = cp.mA;
  fp.kA = cp.kA;
  fp.nB = cp.nB;
  fp.kB = cp.kB;
  fp.M = cp.M;
  fp.N = cp.N;
  fp.K = cp.K;
  fp.num_sms = L.num_sms;
  L.begin(KC_GEMM_INT8, double(cp.M * cp.K + cp.N * cp.K + cp.M * cp.N) * 8.0,
          8.0 * double(cp.M) * double(cp.N) * double(cp.K));
  run_zgemm_ozaki_t(L, fp, A, B, C);
  L.end();
  PQ_CUDA(cudaGetLastError());
}

// Sync points of the library (pq_sync, pq_load_tensor) call this once the stream is idle: a
// watchdog event means a k_ozaki_t CTA gave up on an mbarrier wait and its results are garbage --
// that must surface as an error, not as numbers.  `launched` keeps the check off the path of
// handles that never ran the kernel.
void ozaki_t_check_watchdog() {
  if (!g_ozaki_t_launched) return;
  int rec[8] = {0};
  PQ_CUDA(cudaMemcpyFromSymbol(rec, g_ot_debug, sizeof(rec)));
  if (rec[0] == 0) return;
  int zero[8] = {0};
  PQ_CUDA(cudaMemcpyToSymbol(g_ot_debug, zero, sizeof(zero)));
  char msg[200];
  std::snprintf(msg, sizeof(msg), "k_ozaki_t watchdog: wait id %d (1 full, 2 freed, 3 done, 4 empty) tile %d group %d "
                "block %d warp %d never completed; results of that launch are invalid", rec[1], rec[2], rec[3], rec[4], rec[5]);
  throw Error(PQ_ERR_CUDA, msg);
}

// pq_microbench back ends of this kernel: "ozaki_t_debug" (watchdog record, 0 = none),
// "ozaki_t_trace" (block 0's phase stamps -> $PQ_OZAKI_TRACE)
double run_ozaki_t_microbench(const Launch& L, const std::string& what) {
  if (what == "ozaki_t_debug") {
    int rec[8] = {0};
    PQ_CUDA(cudaStreamSynchronize(L.stream));
    PQ_CUDA(cudaMemcpyFromSymbol(rec, g_ot_debug, sizeof(rec)));
    if (rec[0]) {
      std::fprintf(stderr, "ozaki_t watchdog: wait id %d (1 full, 2 freed, 3 done, 4 empty) tile %d "
                           "group %d block %d warp %d\n", rec[1], rec[2], rec[3], rec[4], rec[5]);
      int zero[8] = {0};
      PQ_CUDA(cudaMemcpyToSymbol(g_ot_debug, zero, sizeof(zero)));
      return rec[1];
    }
    return 0;
  }
  if (what.rfind("umma_i8_", 0) == 0) {
  if (what == "umma_i8_selftest") {
      int* d = nullptr;
      PQ_CUDA(cudaMalloc(&d, OT_SELF_M * 32 * sizeof(int)));
      PQ_CUDA(cudaMemsetAsync(d, 0xff, OT_SELF_M * 32 * sizeof(int), L.stream));
      k_umma_i8_selftest<<<1, 128, 0, L.stream>>>(d);
      std::vector<int> got(OT_SELF_M * 32);
      cudaError_t e = cudaMemcpyAsync(got.data(), d, got.size() * sizeof(int), cudaMemcpyDeviceToHost,
                                      L.stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(L.stream);
      cudaFree(d);
      PQ_CUDA(e);
      if (const char* path = std::getenv("PQ_OZAKI_DUMP")) {   // raw 128 x 32 int32, row-major
        if (FILE* f = std::fopen(path, "wb")) {
          std::fwrite(got.data(), sizeof(int), got.size(), f);
          std::fclose(f);
        }
      }
      int wrong = 0;
      for (int r = 0; r < OT_SELF_M; ++r)
        for (int c = 0; c < 32; ++c) {
          int want = 0;
          for (int k = 0; k < 32; ++k) want += ot_pat_a(r, k) * ot_pat_b(c, k);
          wrong += got[r * 32 + c] != want;
        }
      return wrong;
    }
    const bool n64 = what == "umma_i8_tops_n64";
    PQ_REQUIRE(n64 || what == "umma_i8_tops_n32", PQ_ERR_INVALID, "unknown microbench: " + what);
    const int iters = 4096, smem = OT_SELF_A_PLANE + OT_SELF_B_PLANE;
    int* sink = nullptr;
    PQ_CUDA(cudaMalloc(&sink, L.num_sms * sizeof(int)));
    cudaEvent_t e0, e1;
    PQ_CUDA(cudaEventCreate(&e0));
    PQ_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      PQ_CUDA(cudaEventRecord(e0, L.stream));
      if (n64)
        k_umma_i8_rate<64><<<L.num_sms, 128, smem, L.stream>>>(iters, sink);
      else
        k_umma_i8_rate<32><<<L.num_sms, 128, smem, L.stream>>>(iters, sink);
      PQ_CUDA(cudaEventRecord(e1, L.stream));
      PQ_CUDA(cudaEventSynchronize(e1));
      float ms = 0;
      PQ_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    PQ_CUDA(e);
    const double macs = double(L.num_sms) * iters * 16.0 * OT_SELF_M * (n64 ? 64 : 32) * 32;
    return 2.0 * macs / (best * 1e-3) / 1e12;
  
  }
  if (what.rfind("ozaki_t_ldtm_", 0) == 0) {   // TMEM read bytes per SM clock
    int W = 8, warps = 8;
    std::sscanf(what.c_str() + 13, "%d_%d", &W, &warps);
    PQ_REQUIRE((W == 4 || W == 8 || W == 16 || W == 32) && (warps == 4 || warps == 8 || warps == 16),
               PQ_ERR_INVALID, "ozaki_t_ldtm_<4|8|16|32>_<4|8|16>");
    long long* d = nullptr;
    PQ_CUDA(cudaMalloc(&d, L.num_sms * 16 * sizeof(long long)));
    for (int rep = 0; rep < 2; ++rep) {
      if (W == 4) k_ot_ldtm_rate<4><<<L.num_sms, warps * 32, 0, L.stream>>>(d);
      else if (W == 8) k_ot_ldtm_rate<8><<<L.num_sms, warps * 32, 0, L.stream>>>(d);
      else if (W == 16) k_ot_ldtm_rate<16><<<L.num_sms, warps * 32, 0, L.stream>>>(d);
      else k_ot_ldtm_rate<32><<<L.num_sms, warps * 32, 0, L.stream>>>(d);
    }
    std::vector<long long> h(L.num_sms * 16);
    cudaError_t e = cudaMemcpyAsync(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, L.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(L.stream);
    cudaFree(d);
    PQ_CUDA(e);
    double worst = 0;   // slowest warp of SM 0..n: the SM's drain time
    for (int b = 0; b < L.num_sms; ++b)
      for (int w = 0; w < warps; ++w) worst = std::max(worst, (double)h[b * 16 + w]);
    return double(warps) * 256.0 * W * 32 * 4 / worst;
  }
  if (what.rfind("ozaki_t_rate_", 0) == 0) {   // SM clocks per MMA, mean over the SMs
    const int mode = std::atoi(what.c_str() + 13), iters = 200;
    const int smem = 6 * ot::W_PLANE + 2 * 6 * ot::X_PLANE;
    auto kern = (mode & 2048) ? ((mode & 256) ? k_ot_mma_rate<true, true> : k_ot_mma_rate<false, true>)
                              : ((mode & 256) ? k_ot_mma_rate<true, false> : k_ot_mma_rate<false, false>);
    PQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    long long* d = nullptr;
    PQ_CUDA(cudaMalloc(&d, 512 * sizeof(long long)));
    PQ_CUDA(cudaMemsetAsync(d, 0, 512 * sizeof(long long), L.stream));
    std::vector<long long> h(512);
    const int grid = L.num_sms < 148 ? L.num_sms : 148;
    double2* buf = nullptr;
    const long long elems = (mode & 512) ? (1ll << 26) : 0;   // 1 GiB
    if (elems) {
      PQ_CUDA(cudaMalloc(&buf, elems * sizeof(double2)));
      PQ_CUDA(cudaMemsetAsync(buf, 0, elems * sizeof(double2), L.stream));
    }
    for (int rep = 0; rep < 2; ++rep) kern<<<grid, OT_THREADS, smem, L.stream>>>(iters, mode, d, buf, elems);
    cudaError_t e = cudaMemcpyAsync(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, L.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(L.stream);
    cudaFree(d);
    if (buf) cudaFree(buf);
    PQ_CUDA(e);
    double sum = 0, side = 0;
    for (int b = 0; b < grid; ++b) {
      sum += (double)h[b];
      side += (double)h[148 + b];
    }
    if (mode & 512)   // clocks per 64 KB batch of the load stream
      std::fprintf(stderr, "ozaki_t_rate mode %d: load stream %.0f clk per 64 KB batch\n", mode, side / grid / iters);
    else if (mode & (32 | 64 | 128))   // clocks per warp-level arithmetic instruction of the side loop (16 warps)
      std::fprintf(stderr, "ozaki_t_rate mode %d: side loop %.2f clk per warp instruction per SMSP\n", mode,
                   side / grid / (4096.0 * 8.0 * 4.0));
    return sum / grid / (double(iters) * 84.0);
  }
  PQ_REQUIRE(what == "ozaki_t_trace", PQ_ERR_INVALID, "unknown microbench: " + what);
  std::vector<long long> tr(32 * 32);
  PQ_CUDA(cudaStreamSynchronize(L.stream));
  PQ_CUDA(cudaMemcpyFromSymbol(tr.data(), g_ot_trace, tr.size() * sizeof(long long)));
  {
    std::vector<long long> bn(2 * 160);
    PQ_CUDA(cudaMemcpyFromSymbol(bn.data(), g_ot_block_ns, bn.size() * sizeof(long long)));
    long long first = 0, last = 0, dmin = 1ll << 60, dmax = 0;
    for (int b = 0; b < L.num_sms && b < 160; ++b) {
      if (bn[2 * b] == 0) continue;
      if (first == 0 || bn[2 * b] < first) first = bn[2 * b];
      if (bn[2 * b + 1] > last) last = bn[2 * b + 1];
      dmin = std::min(dmin, bn[2 * b + 1] - bn[2 * b]);
      dmax = std::max(dmax, bn[2 * b + 1] - bn[2 * b]);
    }
    std::fprintf(stderr, "ozaki_t blocks: first start -> last end %.1f us, block lifetime min %.1f max %.1f us\n",
                 (last - first) / 1e3, dmin / 1e3, dmax / 1e3);
  }
  if (const char* path = std::getenv("PQ_OZAKI_TRACE"))
    if (FILE* f = std::fopen(path, "wb")) {
      std::fwrite(tr.data(), sizeof(long long), tr.size(), f);
      std::fclose(f);
    }
  return 0;
}

}  // namespace pq
