// K2o (EXPERIMENTAL, opt-in: option "zgemm_ozaki" = 6 or 7; NOT yet run on hardware, see
// DESIGN.md section 2.1): ComplexF64 GEMM for the skinny sweep steps on the INT8 tensor cores
// of sm_100a (tcgen05.mma kind::i8, int32 accumulators in TMEM) -- an Ozaki-scheme
// ("error-free slicing") product.
//
// Why.  tcgen05 has no f64 kind; the FP64 tensor pipe (DMMA) tops out at ~37 TFLOP/s, and the
// dominant sweep step of the bench workload (M = 2^18, N = K = 64) is bound by it: 246 us,
// although its operands and result stream through HBM in ~85 us.  The INT8 pipe is two orders
// of magnitude wider, and the backend's ComplexF64 contract is a 1e-10 relative L2 tolerance,
// not 53 mantissa bits per product.
//
// The arithmetic (ozaki_math.h; executed on the host by test_lower.cpp: test_ozaki, with the
// tensor core replaced by an integer GEMM reading through the same LBO / SBO addressing).
// Every row of A (all k, re and im together) gets one power-of-two scale 2^EA with
// |x| < 2^EA; likewise every column n of B.  A real x of that row becomes the integer
//     q = rint(x * 2^(46 - EA)),   |q| <= 2^46,
// written in balanced base 256:  q = sum_{i<6} d_i 256^i,  d_i in [-128, 127]  (int8).
// The digits are simply the bytes of q + BIAS (BIAS = sum 128*256^i) with the top bit flipped:
// no carry chain, no bit-field shuffling.  With s counting digits from the most significant
// one, the product of digit planes s of A and t of B has weight 256^(10 - s - t); all pairs
// with the same g = s + t are summed EXACTLY by the tensor core in one int32 accumulator
// (|acc| < 2^25 for K <= 64), and pairs with g >= G are dropped (G = 6: 21 pairs, rel-L2
// ~2e-13 on well-scaled rows; G = 7: 26 pairs, ~2e-14).  The complex product uses four real
// ones; the minus sign of Cr = Ar Br - Ai Bi is carried by a third, negated set of B planes
// (B is tiny and resident).  Epilogue:  C = 2^(EA + EB - 12) * sum_g acc_g 256^-g, evaluated
// as two int64 Horner sums (g < 3 and g >= 3), two int64 -> f64 conversions and one FMA per
// real number.
//
// The kernel (same envelope as k_zgemm_skinny: K <= 64, N <= 64, gather straight from the
// un-permuted operands).  One persistent CTA per SM, 16 worker warps + 1 MMA-issue warp:
//   * B (K x N) is gathered, scaled and sliced into 3 x 6 resident int8 planes once per CTA;
//   * a tile is 128 rows of A x all of K: worker thread (row, 16-k chunk) loads its 16 complex
//     numbers, the row exponent is an atomicMax over the four chunk threads, and each thread
//     writes one 16-byte core-matrix row per plane (6 re + 6 im planes, no-swizzle K-major
//     UMMA layout, conflict-free STS.128);
//   * the MMA warp issues, per 32-column half of N, G x (pairs) x 4 x (K/32) MMAs of shape
//     128 x 32 x 32 into 2 G accumulators of 32 TMEM columns (G = 7: 448 of the 512 columns);
//   * the worker warps drain TMEM (tcgen05.ld), combine, scale and store C[m + M n]
//     (lane = row: 512-byte coalesced stores); the next tile of A is L2-prefetched meanwhile.
// MMAs and TMEM drain are pipelined per accumulator group (mbarriers done[g] / freed[g]);
// gather + slicing of a tile is serial with them (one set of A planes).  Expected bound: TMEM
// read bandwidth + slicing ALU work, ~2x under the DMMA time -- to be measured.
//
// Long contractions: k_zgemm_ozaki_kloop below (canonical layouts, K in chunks of 64).
//
// ComplexF32 twin (option "cgemm_ozaki" = 4): the same kernel with 4 digits per real
// (q = rint(x * 2^(30 - E)), nothing of the float is lost), G = 4 groups = 10 plane pairs, all
// 64 columns in one pass (2 * 4 * 64 = 512 TMEM columns), result rounded once to float:
// rel-L2 ~3e-8 per contraction, and -- unlike the K1 + tcgen05 3xTF32 path -- the gather is
// fused, so A is read once and nothing is materialised.
#include <cstdio>
#include <cstdlib>

#include "common.h"
#include "ozaki_math.h"

namespace pq {

namespace {

constexpr int OZ_TM = 128;                 // rows per tile = UMMA M
constexpr int OZ_WORKERS = 512;            // 16 worker warps: (row, 16-k chunk)
constexpr int OZ_THREADS = OZ_WORKERS + 32;
constexpr int OZ_KMAX = 64, OZ_NMAX = 64;
constexpr int OZ_A_PLANE = OZ_TM * OZ_KMAX;     // bytes (int8)
constexpr int OZ_B_PLANE = OZ_NMAX * OZ_KMAX;

template <class Real> struct OzVec;
template <> struct OzVec<double> { using type = double2; };
template <> struct OzVec<float> { using type = float2; };

// shared-memory layout for S digits per real: A planes (re, im), B planes (re, im, -im), tables
template <int S>
struct OzSmem {
  static constexpr int kA = 0;
  static constexpr int kB = kA + 2 * S * OZ_A_PLANE;
  static constexpr int kKoffA = kB + 3 * S * OZ_B_PLANE;    // int[64]
  static constexpr int kKoffB = kKoffA + OZ_KMAX * 4;       // int[64]
  static constexpr int kRowE = kKoffB + OZ_KMAX * 4;        // int[2][128]
  static constexpr int kColE = kRowE + 2 * OZ_TM * 4;       // int[64]
  static constexpr int kBars = kColE + OZ_NMAX * 4;         // 1 + 8 + 8 mbarriers
  static constexpr int kTotal = kBars + 17 * 8 + 16;        // + tmem slot, abort flag
};
static_assert(OzSmem<6>::kTotal <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ uint32_t oz_smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void oz_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(oz_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void oz_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(oz_smem_u32(bar)) : "memory");
}
// Watchdog for the bring-up phase: a wait that does not complete within ~2 s records WHICH
// wait it was in g_oz_debug and raises a CTA-wide abort flag; every other wait of the CTA then
// returns at once, so a wrong barrier phase ends as a finished kernel with a diagnosis
// (pq_microbench "ozaki_debug") instead of a hung GPU.
//   g_oz_debug = {flag, wait id, iteration, group, block, warp, -, -}
__device__ int g_oz_debug[8];
// Phase trace of block 0 (template parameter TR; off in the product instantiations): clock64
// stamps per tile, see tools/ozaki_trace.py for the event numbering.
__device__ long long g_oz_trace[16 * 64];
#define OZ_TRACE(tile_no, ev)                                                             \
  do {                                                                                    \
    if (TR && blockIdx.x == 0 && (tile_no) < 16) g_oz_trace[(tile_no) * 64 + (ev)] = clock64(); \
  } while (0)
constexpr long long OZ_WAIT_LIMIT = 4000000000ll;   // clock64 ticks

__device__ __forceinline__ bool oz_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(oz_smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// wait ids: 1 planes (MMA warp), 2 freed[g] (MMA warp), 3 done[g] (workers), 4 probes
__device__ __noinline__ void oz_wait_slow(uint64_t* bar, uint32_t parity, volatile int* abort_flag,
                                          int id, int it, int g) {
  const long long t0 = clock64();
  while (!oz_try_wait(bar, parity)) {
    if (*abort_flag) return;
    if (clock64() - t0 > OZ_WAIT_LIMIT) {
      *abort_flag = 1;
      if (atomicCAS(&g_oz_debug[0], 0, 1) == 0) {
        g_oz_debug[1] = id;
        g_oz_debug[2] = it;
        g_oz_debug[3] = g;
        g_oz_debug[4] = (int)blockIdx.x;
        g_oz_debug[5] = (int)(threadIdx.x >> 5);
        __threadfence();
      }
      return;
    }
  }
}
__device__ __forceinline__ void oz_mbar_wait(uint64_t* bar, uint32_t parity, volatile int* abort_flag,
                                             int id, int it, int g) {
  if (!oz_try_wait(bar, parity)) oz_wait_slow(bar, parity, abort_flag, id, it, g);
}

// no-swizzle K-major shared-memory matrix descriptor: core matrix = 8 rows x 16 bytes
// (16 int8 k-elements); LBO = bytes between core matrices adjacent in K, SBO = between
// 8-row groups (cute/arch/mma_sm100_desc.hpp: SmemDescriptor, version 1, SWIZZLE_NONE)
__device__ __forceinline__ uint64_t oz_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor for kind::i8 (cute InstrDescriptor): c_format S32 = 2 at bit 4,
// a_format / b_format INT8 (signed) = 1 at bits 7 / 10, both K-major, N >> 3 at bit 17,
// M >> 4 at bit 24, no saturation
__host__ __device__ constexpr uint32_t oz_idesc(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
// The descriptors are passed as (low, high) words: only the 14-bit start-address field in the
// low word changes from MMA to MMA, so the issuing thread does 32-bit adds of constants.
__device__ __forceinline__ void oz_umma_i8(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi,
                                           uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                           uint32_t accumulate) {
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
__device__ __forceinline__ void oz_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                   oz_smem_u32(bar))
               : "memory");
}

#define OZ_TMEM_LD8(r, addr)                                                                  \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"    \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]),     \
                 "=r"(r[6]), "=r"(r[7])                                                       \
               : "r"(addr))

__device__ __forceinline__ void oz_store(unsigned char* dst, const oz::Word4& v) {
  *reinterpret_cast<uint4*>(dst) = make_uint4(v.w[0], v.w[1], v.w[2], v.w[3]);
}

// Real = double (ComplexF64: 6 digits, NC = 32 columns per pass, G = 6 / 7 accumulator groups)
// or float (ComplexF32: 4 digits, NC = 64, G = 3 / 4).  2 * G * NC accumulator columns <= 512.
template <class Real, int G, int NC, bool TR = false>
__global__ void __maxnreg__(96)
k_zgemm_ozaki(const typename OzVec<Real>::type* __restrict__ A,
              const typename OzVec<Real>::type* __restrict__ B,
              typename OzVec<Real>::type* __restrict__ C, const FusedParams p) {
  using Tr = oz::Traits<Real>;
  using V2 = typename OzVec<Real>::type;
  using Sm = OzSmem<Tr::S>;
  constexpr int S = Tr::S;
  constexpr int CW = NC / 4;   // accumulator columns drained by one worker warp
  static_assert(G >= oz::HI_GROUPS && 2 * G * NC <= 512, "accumulator columns must fit TMEM");
  static_assert(NC == 32 || NC == 64, "one or two passes over N <= 64");
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem + Sm::kA;
  unsigned char* sB = smem + Sm::kB;
  int* koffA = reinterpret_cast<int*>(smem + Sm::kKoffA);
  int* koffB = reinterpret_cast<int*>(smem + Sm::kKoffB);
  int* rowE = reinterpret_cast<int*>(smem + Sm::kRowE);     // [2][128] biased exponent fields
  int* colE = reinterpret_cast<int*>(smem + Sm::kColE);
  // mbarriers.  planes: the worker warps have written this tile's A planes (count 16).
  // done[g]: the MMAs of accumulator group g have completed (tcgen05.commit).  freed[g]: every
  // worker warp has read group g out of TMEM (count 16).  done / freed complete one phase per
  // (tile, column pass), planes one per tile.
  uint64_t* planes = reinterpret_cast<uint64_t*>(smem + Sm::kBars);
  uint64_t* done = planes + 1;
  uint64_t* freed = done + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(freed + 8);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = (int)p.K, N = (int)p.N;
  const long long M = p.M;
  const int KS = (K + 31) / 32;              // UMMA k-steps (32 int8 each)
  const int NH = (N + NC - 1) / NC;          // column passes over N
  const long long tiles = (M + OZ_TM - 1) / OZ_TM;

  if (tid < OZ_KMAX) {
    koffA[tid] = tid < K ? (int)map_offset(p.kA, tid) : 0;
    koffB[tid] = tid < K ? (int)map_offset(p.kB, tid) : 0;
    colE[tid] = 0;
  }
  if (tid < 2 * OZ_TM) rowE[tid] = 0;
  if (tid == 0) {
    *abort_flag = 0;
    oz_mbar_init(planes, OZ_WORKERS / 32);
    for (int g = 0; g < 8; ++g) {
      oz_mbar_init(&done[g], 1);
      oz_mbar_init(&freed[g], OZ_WORKERS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     oz_smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  constexpr uint32_t A_LBO = OZ_TM * 16, B_LBO = OZ_NMAX * 16, SBO = 128;

  if (warp == OZ_WORKERS / 32) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t IDESC = oz_idesc(OZ_TM, NC);
      const uint64_t a_base = oz_desc(oz_smem_u32(sA), A_LBO, SBO);
      const uint64_t b_base = oz_desc(oz_smem_u32(sB), B_LBO, SBO);
      const uint32_t a_lo = (uint32_t)a_base, a_hi = (uint32_t)(a_base >> 32);
      const uint32_t b_lo = (uint32_t)b_base, b_hi = (uint32_t)(b_base >> 32);
      uint32_t it = 0, tile_no = 0;
      for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tile_no) {
        oz_mbar_wait(planes, tile_no & 1u, abort_flag, 1, (int)it, -1);
        OZ_TRACE(tile_no, 48);
        for (int h = 0; h < NH; ++h, ++it) {
          const uint32_t bh_lo = b_lo + (uint32_t)((h * (NC / 8) * SBO) >> 4);   // rows NC h .. of B
#pragma unroll
          for (int g = 0; g < G; ++g) {
            // the accumulators of group g must have been drained by the previous (tile, pass)
            if (it > 0) oz_mbar_wait(&freed[g], (it - 1) & 1u, abort_flag, 2, (int)it, g);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            // descriptor address fields are in 16-byte units; one K-step = two core matrices
            auto mma = [&](int accum, int a_plane, int b_plane, int ks, uint32_t acc) {
              oz_umma_i8(tmem_base + (uint32_t)(accum * NC),
                         a_lo + (uint32_t)((a_plane * OZ_A_PLANE + ks * 2 * (int)A_LBO) >> 4), a_hi,
                         bh_lo + (uint32_t)((b_plane * OZ_B_PLANE + ks * 2 * (int)B_LBO) >> 4), b_hi,
                         IDESC, acc);
            };
            if (KS == 2)
              oz::for_each_mma_of_group<S, 2>(g, mma);
            else
              oz::for_each_mma_of_group<S, 1>(g, mma);
            oz_commit(&done[g]);   // group g may be read while the next groups are computed
            OZ_TRACE(tile_no, 49 + h * 7 + g);
          }
        }
      }
    }
  } else {
    // ===================== workers: gather + slice, then drain TMEM =====================
    const int row = tid & (OZ_TM - 1), chunk = tid >> 7;       // chunk = 16 consecutive k
    const bool chunk_on = chunk * 16 < KS * 32;                // inside the padded K
    auto workers_barrier = [&]() {
      asm volatile("bar.sync 1, %0;\n" ::"r"(OZ_WORKERS) : "memory");
    };

    // ---- B: gathered, scaled per column n, sliced into re / im / -im planes, once ----
    {
      const int n = tid & (OZ_NMAX - 1), bchunk = (tid >> 6) & 3;
      const bool mine = tid < 4 * OZ_NMAX && n < NH * NC && bchunk * 16 < KS * 32;
      Real xr[16], xi[16];
      int ef = 0;
      if (mine) {
        const long long rb = n < N ? map_offset(p.nB, n) : -1;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int k = bchunk * 16 + j;
          V2 v;
          v.x = v.y = (Real)0;
          if (rb >= 0 && k < K) v = B[rb + koffB[k]];
          xr[j] = v.x;
          xi[j] = v.y;
          ef = max(ef, max(Tr::key(v.x), Tr::key(v.y)));
        }
        atomicMax(&colE[n], Tr::exp_field(ef));
      }
      workers_barrier();
      if (mine) {
        const auto scale = Tr::slice_scale(colE[n]);
        oz::Word4 pl[S];
        unsigned char* dst = sB + oz::plane_off(OZ_NMAX, n, bchunk);
        Tr::slice16(xr, scale, false, pl);
#pragma unroll
        for (int s = 0; s < S; ++s) oz_store(dst + s * OZ_B_PLANE, pl[s]);
        Tr::slice16(xi, scale, false, pl);
#pragma unroll
        for (int s = 0; s < S; ++s) oz_store(dst + (S + s) * OZ_B_PLANE, pl[s]);
        Tr::slice16(xi, scale, true, pl);
#pragma unroll
        for (int s = 0; s < S; ++s) oz_store(dst + (2 * S + s) * OZ_B_PLANE, pl[s]);
      }
      // (made visible to the tensor core by the proxy fence before the first `planes` arrival)
    }

    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;   // this warp's TMEM lanes
    const int cpart = warp >> 2;                                    // CW of the NC columns
    uint32_t it = 0;
    int buf = 0;
    uint32_t tno = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, buf ^= 1, ++tno) {
      const long long m = tile * OZ_TM + row;
      if (tid == 0) OZ_TRACE(tno, 0);
      const long long ra = m < M ? map_offset(p.mA, m) : -1;
      // ---- gather this thread's 16 complex numbers, row exponent ----
      Real xr[16], xi[16];
      int ef = 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int k = chunk * 16 + j;
        V2 v;
        v.x = v.y = (Real)0;
        if (chunk_on && ra >= 0 && k < K) v = A[ra + koffA[k]];
        xr[j] = v.x;
        xi[j] = v.y;
        ef = max(ef, max(Tr::key(v.x), Tr::key(v.y)));
      }
      if (chunk_on) atomicMax(&rowE[buf * OZ_TM + row], Tr::exp_field(ef));
      if (tid == 0) OZ_TRACE(tno, 1);
      workers_barrier();
      if (tid == 0) OZ_TRACE(tno, 2);
      const int ea = rowE[buf * OZ_TM + row];
      if (tid < OZ_TM) rowE[(buf ^ 1) * OZ_TM + tid] = 0;    // for the next tile (see header)
      // ---- slice into the re and im digit planes of this tile ----
      if (chunk_on) {
        const auto scale = Tr::slice_scale(ea);
        oz::Word4 pl[S];
        unsigned char* dst = sA + oz::plane_off(OZ_TM, row, chunk);
        Tr::slice16(xr, scale, false, pl);
#pragma unroll
        for (int s = 0; s < S; ++s) oz_store(dst + s * OZ_A_PLANE, pl[s]);
        Tr::slice16(xi, scale, false, pl);
#pragma unroll
        for (int s = 0; s < S; ++s) oz_store(dst + (S + s) * OZ_A_PLANE, pl[s]);
      }
      if (tid == 0) OZ_TRACE(tno, 3);
      // L2 prefetch of the next tile's rows (registers are needed by the epilogue)
      {
        const long long mn = m + (long long)gridDim.x * OZ_TM;
        if (chunk_on && mn < M) {
          const long long rn = map_offset(p.mA, mn);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int k = chunk * 16 + j;
            if (k < K) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(A + rn + koffA[k]));
          }
        }
      }
      // output scale of this row: 2^(EA - 6) with |x| < 2^EA
      const double sa = Tr::out_scale(ea);

      // the planes of this tile are written: generic-proxy stores -> visible to the tensor core
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) oz_mbar_arrive(planes);
      if (tid == 0) OZ_TRACE(tno, 4);

      for (int h = 0; h < NH; ++h, ++it) {
        const int nbase = h * NC + cpart * CW;
        if (nbase >= N) {
          // this warp's columns lie beyond N (narrow steps): keep the barrier protocol in
          // step, skip the TMEM loads and the arithmetic
#pragma unroll 1
          for (int g = 0; g < G; ++g) {
            oz_mbar_wait(&done[g], it & 1u, abort_flag, 3, (int)it, g);
            __syncwarp();
            if (lane == 0) oz_mbar_arrive(&freed[g]);
          }
          continue;
        }
#pragma unroll 1
        for (int cb = 0; cb < CW / 8; ++cb) {   // 8 columns at a time (register budget)
          const int n0 = nbase + cb * 8;
          long long hr[8], hq[8];   // Horner sums over the groups, re / im
          long long fr[8], fq[8];   // the finished first sums (groups 0 .. HI_GROUPS-1)
#pragma unroll
          for (int j = 0; j < 8; ++j) hr[j] = hq[j] = fr[j] = fq[j] = 0;
#pragma unroll
          for (int g = 0; g < G; ++g) {
            uint32_t r[8], q[8];
            const uint32_t col = tmem_base + lane_base + (uint32_t)((2 * g) * NC + cpart * CW + cb * 8);
            // (already complete on the later column blocks of this pass)
            oz_mbar_wait(&done[g], it & 1u, abort_flag, 3, (int)it, g);
            if (tid == 0 && cb == 0) OZ_TRACE(tno, 5 + h * 16 + g);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            OZ_TMEM_LD8(r, col);
            OZ_TMEM_LD8(q, col + NC);
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            if (tid == 0 && cb == 0) OZ_TRACE(tno, 13 + h * 16 + g);
            if (cb == CW / 8 - 1) {
              // this warp is done with group g: the MMA warp may overwrite it (next pass / tile)
              asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
              __syncwarp();
              if (lane == 0) oz_mbar_arrive(&freed[g]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              hr[j] = hr[j] * 256 + (long long)(int)r[j];
              hq[j] = hq[j] * 256 + (long long)(int)q[j];
            }
            if (g == oz::HI_GROUPS - 1) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                fr[j] = hr[j];
                fq[j] = hq[j];
                hr[j] = hq[j] = 0;
              }
            }
          }
          if (m < M) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int n = n0 + j;
              if (n < N) {
                const double sc = sa * Tr::out_scale(colE[n]);
                V2 out;
                out.x = (Real)(oz::combine(fr[j], hr[j], G) * sc);
                out.y = (Real)(oz::combine(fq[j], hq[j], G) * sc);
                C[m + M * n] = out;
              }
            }
          }
        }
        if (tid == 0) OZ_TRACE(tno, 40 + h);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base),
                 "r"(512u)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------
// K2'o: the same product for LONG contractions on canonical layouts A[m + M k], B[n + N k]
// (behind K1, where k_zgemm_dmma3m runs today): K is walked in chunks of 64 with the int32
// accumulators staying in TMEM, so one CTA tile is 128 rows x NC columns x all of K.
//   * the row / column scales must be the same for every chunk: k_oz_row_exponents computes
//     them once per operand (max exponent field over the whole row) into the plan's workspace;
//   * per chunk the workers load + slice 128 x 64 of A and NC x 64 of B (B is not resident
//     here), wait until the MMAs of the previous chunk have consumed the planes
//     (`consumed`, tcgen05.commit), write the planes and arrive on `planes`; the loads of the
//     next chunk are already in flight while the tensor core works on the current one;
//   * the MMA thread accumulates chunk after chunk; on the last chunk it commits group by
//     group (`done[g]`) and the epilogue of the skinny kernel drains TMEM.
// |acc| <= 2^14 * K * 12 bounds K to 8192 (int32); longer contractions stay on DMMA.
// Tiles are rasterised in bands of 16 row tiles so that the CTAs of a wave share a band of A
// and a few column blocks of B in L2.
// ---------------------------------------------------------------------------

template <class Real>
__global__ void __launch_bounds__(128)
k_oz_row_exponents(const typename OzVec<Real>::type* __restrict__ X, long long R, long long K,
                   int* __restrict__ E) {
  using Tr = oz::Traits<Real>;
  const long long r = (long long)blockIdx.x * 128 + threadIdx.x;
  if (r >= R) return;
  const long long per = (K + gridDim.y - 1) / gridDim.y;
  const long long k0 = (long long)blockIdx.y * per, k1 = (k0 + per < K) ? k0 + per : K;
  int key = 0;
  for (long long k = k0; k < k1; ++k) {
    const typename OzVec<Real>::type v = X[r + R * k];
    key = max(key, max(Tr::key(v.x), Tr::key(v.y)));
  }
  atomicMax(&E[r], Tr::exp_field(key));   // E is zeroed by the launcher
}

template <class Real, int G, int NC>
__global__ void __maxnreg__(96)
k_zgemm_ozaki_kloop(const typename OzVec<Real>::type* __restrict__ A,
                    const typename OzVec<Real>::type* __restrict__ B,
                    typename OzVec<Real>::type* __restrict__ C, long long M, long long N, long long K,
                    const int* __restrict__ expA, const int* __restrict__ expB) {
  using Tr = oz::Traits<Real>;
  using V2 = typename OzVec<Real>::type;
  using Sm = OzSmem<Tr::S>;
  constexpr int S = Tr::S;
  constexpr int CW = NC / 4;
  static_assert(G >= oz::HI_GROUPS && 2 * G * NC <= 512, "accumulator columns must fit TMEM");
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sA = smem + Sm::kA;
  unsigned char* sB = smem + Sm::kB;
  // planes: this chunk's planes are written (count 16, one phase per chunk); consumed: the MMAs
  // of a chunk have completed (commit, one phase per chunk); done[g] / freed[g]: as in the
  // skinny kernel, one phase per tile
  uint64_t* planes = reinterpret_cast<uint64_t*>(smem + Sm::kBars);
  uint64_t* done = planes + 1;
  uint64_t* freed = done + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(freed + 8);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  uint64_t* consumed = reinterpret_cast<uint64_t*>(smem + Sm::kRowE);   // (the row table is unused here)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long tiles_m = (M + OZ_TM - 1) / OZ_TM, tiles_n = (N + NC - 1) / NC;
  const long long tiles = tiles_m * tiles_n;
  const long long chunks = (K + OZ_KMAX - 1) / OZ_KMAX;
  constexpr long long BAND = 16;
  auto tile_coords = [&](long long t, long long& mt, long long& nt) {
    const long long band = t / (BAND * tiles_n), within = t - band * BAND * tiles_n;
    const long long bm = (tiles_m - band * BAND) < BAND ? (tiles_m - band * BAND) : BAND;
    nt = within / bm;
    mt = band * BAND + within % bm;
  };

  if (tid == 0) {
    *abort_flag = 0;
    oz_mbar_init(planes, OZ_WORKERS / 32);
    oz_mbar_init(consumed, 1);
    for (int g = 0; g < 8; ++g) {
      oz_mbar_init(&done[g], 1);
      oz_mbar_init(&freed[g], OZ_WORKERS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     oz_smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  constexpr uint32_t A_LBO = OZ_TM * 16, B_LBO = OZ_NMAX * 16, SBO = 128;

  if (warp == OZ_WORKERS / 32) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t IDESC = oz_idesc(OZ_TM, NC);
      const uint64_t a_base = oz_desc(oz_smem_u32(sA), A_LBO, SBO);
      const uint64_t b_base = oz_desc(oz_smem_u32(sB), B_LBO, SBO);
      const uint32_t a_lo = (uint32_t)a_base, a_hi = (uint32_t)(a_base >> 32);
      const uint32_t b_lo = (uint32_t)b_base, b_hi = (uint32_t)(b_base >> 32);
      uint32_t it = 0, cc = 0;   // tiles / chunks done by this CTA
      for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        for (long long c = 0; c < chunks; ++c, ++cc) {
          const bool last = c == chunks - 1;
          const int kleft = (int)((K - c * OZ_KMAX) < OZ_KMAX ? (K - c * OZ_KMAX) : OZ_KMAX);
          const int KS = (kleft + 31) / 32;
          oz_mbar_wait(planes, cc & 1u, abort_flag, 1, (int)cc, -1);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
          for (int g = 0; g < G; ++g) {
            if (c == 0 && it > 0) {   // the previous tile's epilogue must have drained group g
              oz_mbar_wait(&freed[g], (it - 1) & 1u, abort_flag, 2, (int)it, g);
              asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            }
            auto mma = [&](int accum, int a_plane, int b_plane, int ks, uint32_t acc) {
              oz_umma_i8(tmem_base + (uint32_t)(accum * NC),
                         a_lo + (uint32_t)((a_plane * OZ_A_PLANE + ks * 2 * (int)A_LBO) >> 4), a_hi,
                         b_lo + (uint32_t)((b_plane * OZ_B_PLANE + ks * 2 * (int)B_LBO) >> 4), b_hi,
                         IDESC, c > 0 ? 1u : acc);
            };
            if (KS == 2)
              oz::for_each_mma_of_group<S, 2>(g, mma);
            else
              oz::for_each_mma_of_group<S, 1>(g, mma);
            if (last) oz_commit(&done[g]);
          }
          oz_commit(consumed);   // the planes may be rewritten once these MMAs have completed
        }
      }
    }
  } else {
    // ===================== workers =====================
    const int row = tid & (OZ_TM - 1), chunk = tid >> 7;
    const int bn = tid % NC, bchunk = tid / NC;            // B work item (threads < 4 * NC)
    const bool b_mine = tid < 4 * NC;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int cpart = warp >> 2;
    uint32_t it = 0, cc = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      long long mt, nt;
      tile_coords(tile, mt, nt);
      const long long m = mt * OZ_TM + row, n0 = nt * NC;
      const long long nb = n0 + bn;
      const int ea = m < M ? expA[m] : 0;
      const int eb = (b_mine && nb < N) ? expB[nb] : 0;
      const auto scale_a = Tr::slice_scale(ea);
      const auto scale_b = Tr::slice_scale(eb);
      for (long long c = 0; c < chunks; ++c, ++cc) {
        const long long kbase = c * OZ_KMAX;
        Real xr[16], xi[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const long long k = kbase + chunk * 16 + j;
          V2 v;
          v.x = v.y = (Real)0;
          if (m < M && k < K) v = A[m + M * k];
          xr[j] = v.x;
          xi[j] = v.y;
        }
        // the MMAs of the previous chunk must be done with the planes
        if (cc > 0) oz_mbar_wait(consumed, (cc - 1) & 1u, abort_flag, 5, (int)cc, -1);
        {
          oz::Word4 pl[S];
          unsigned char* dst = sA + oz::plane_off(OZ_TM, row, chunk);
          Tr::slice16(xr, scale_a, false, pl);
#pragma unroll
          for (int s = 0; s < S; ++s) oz_store(dst + s * OZ_A_PLANE, pl[s]);
          Tr::slice16(xi, scale_a, false, pl);
#pragma unroll
          for (int s = 0; s < S; ++s) oz_store(dst + (S + s) * OZ_A_PLANE, pl[s]);
        }
        if (b_mine) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const long long k = kbase + bchunk * 16 + j;
            V2 v;
            v.x = v.y = (Real)0;
            if (nb < N && k < K) v = B[nb + N * k];
            xr[j] = v.x;
            xi[j] = v.y;
          }
          oz::Word4 pl[S];
          unsigned char* dst = sB + oz::plane_off(OZ_NMAX, bn, bchunk);
          Tr::slice16(xr, scale_b, false, pl);
#pragma unroll
          for (int s = 0; s < S; ++s) oz_store(dst + s * OZ_B_PLANE, pl[s]);
          Tr::slice16(xi, scale_b, false, pl);
#pragma unroll
          for (int s = 0; s < S; ++s) oz_store(dst + (S + s) * OZ_B_PLANE, pl[s]);
          Tr::slice16(xi, scale_b, true, pl);
#pragma unroll
          for (int s = 0; s < S; ++s) oz_store(dst + (2 * S + s) * OZ_B_PLANE, pl[s]);
          // B's loads sit behind the slicing of A (registers): pull the next chunk into L2 now
          if (nb < N && c + 1 < chunks) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const long long k = kbase + OZ_KMAX + bchunk * 16 + j;
              if (k < K) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(B + nb + N * k));
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) oz_mbar_arrive(planes);
      }

      // ---- epilogue of the tile (as in the skinny kernel, one column pass) ----
      const double sa = Tr::out_scale(ea);
      const long long nbase = n0 + cpart * CW;
      if (nbase >= N) {
#pragma unroll 1
        for (int g = 0; g < G; ++g) {
          oz_mbar_wait(&done[g], it & 1u, abort_flag, 3, (int)it, g);
          __syncwarp();
          if (lane == 0) oz_mbar_arrive(&freed[g]);
        }
        continue;
      }
#pragma unroll 1
      for (int cb = 0; cb < CW / 8; ++cb) {
        long long hr[8], hq[8], fr[8], fq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) hr[j] = hq[j] = fr[j] = fq[j] = 0;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          uint32_t r[8], q[8];
          const uint32_t col = tmem_base + lane_base + (uint32_t)((2 * g) * NC + cpart * CW + cb * 8);
          oz_mbar_wait(&done[g], it & 1u, abort_flag, 3, (int)it, g);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          OZ_TMEM_LD8(r, col);
          OZ_TMEM_LD8(q, col + NC);
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
          if (cb == CW / 8 - 1) {
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncwarp();
            if (lane == 0) oz_mbar_arrive(&freed[g]);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            hr[j] = hr[j] * 256 + (long long)(int)r[j];
            hq[j] = hq[j] * 256 + (long long)(int)q[j];
          }
          if (g == oz::HI_GROUPS - 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              fr[j] = hr[j];
              fq[j] = hq[j];
              hr[j] = hq[j] = 0;
            }
          }
        }
        if (m < M) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const long long n = nbase + cb * 8 + j;
            if (n < N) {
              const double sc = sa * Tr::out_scale(__ldg(&expB[n]));
              V2 out;
              out.x = (Real)(oz::combine(fr[j], hr[j], G) * sc);
              out.y = (Real)(oz::combine(fq[j], hq[j], G) * sc);
              C[m + M * n] = out;
            }
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base),
                 "r"(512u)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------
// Bring-up aids (pq_microbench "umma_i8_selftest", "umma_i8_tops_n32", "umma_i8_tops_n64").
//
// Self-test: ONE 128 x 32 x 32 kind::i8 MMA on known int8 patterns laid out exactly like the
// kernel's planes (plane_off, LBO = rows * 16, SBO = 128), read back with tcgen05.ld and
// compared on the host with the integer dot products -- isolates the descriptor encodings
// and the TMEM lane / column mapping from everything else.  Returns the number of wrong
// entries (0 = pass).
// ---------------------------------------------------------------------------
__host__ __device__ inline int oz_pat_a(int r, int k) { return (r * 7 + k * 3 + r / 64) % 127 - 63; }
__host__ __device__ inline int oz_pat_b(int c, int k) { return (c * 5 + k * 11 + 1) % 127 - 63; }

__global__ void __launch_bounds__(128, 1) k_umma_i8_selftest(int* __restrict__ out) {
  __shared__ __align__(1024) unsigned char sa[OZ_TM * 32];
  __shared__ __align__(1024) unsigned char sb[32 * 32];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ int abort_flag;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < OZ_TM * 32; i += 128) {
    const int r = i >> 5, k = i & 31;
    sa[oz::plane_off(OZ_TM, r, k >> 4) + (k & 15)] = (unsigned char)(signed char)oz_pat_a(r, k);
  }
  for (int i = tid; i < 32 * 32; i += 128) {
    const int c = i >> 5, k = i & 31;
    sb[oz::plane_off(32, c, k >> 4) + (k & 15)] = (unsigned char)(signed char)oz_pat_b(c, k);
  }
  if (tid == 0) {
    abort_flag = 0;
    oz_mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     oz_smem_u32(&slot)),
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
    const uint64_t ad = oz_desc(oz_smem_u32(sa), OZ_TM * 16, 128);
    const uint64_t bd = oz_desc(oz_smem_u32(sb), 32 * 16, 128);
    oz_umma_i8(tmem, (uint32_t)ad, (uint32_t)(ad >> 32), (uint32_t)bd, (uint32_t)(bd >> 32),
               oz_idesc(OZ_TM, 32), 0u);
    oz_commit(&bar);
  }
  oz_mbar_wait(&bar, 0u, &abort_flag, 4, 0, -1);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
#pragma unroll
  for (int c0 = 0; c0 < 32; c0 += 8) {
    uint32_t r[8];
    OZ_TMEM_LD8(r, tmem + lane_base + c0);
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
  for (int i = tid; i < (OZ_A_PLANE + OZ_B_PLANE) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (tid == 0) {
    abort_flag = 0;
    oz_mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     oz_smem_u32(&slot)),
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
    const uint64_t ad = oz_desc(oz_smem_u32(smem), OZ_TM * 16, 128);
    const uint64_t bd = oz_desc(oz_smem_u32(smem + OZ_A_PLANE), OZ_NMAX * 16, 128);
    constexpr uint32_t IDESC = oz_idesc(OZ_TM, NCOL);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t ks = (uint32_t)(j & 1) * ((2 * OZ_TM * 16) >> 4);
        const uint32_t kb = (uint32_t)(j & 1) * ((2 * OZ_NMAX * 16) >> 4);
        oz_umma_i8(tmem + (uint32_t)((j >> 2) * NCOL), (uint32_t)ad + ks, (uint32_t)(ad >> 32),
                   (uint32_t)bd + kb, (uint32_t)(bd >> 32), IDESC, (it | (j & 3)) ? 1u : 0u);
      }
    }
    oz_commit(&bar);
  }
  oz_mbar_wait(&bar, 0u, &abort_flag, 4, 0, -1);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  uint32_t r[8];
  OZ_TMEM_LD8(r, tmem + ((uint32_t)(warp * 32) << 16));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
  if (r[0] == 0x7fffffffu) sink[blockIdx.x] = (int)r[1];   // keeps the loads alive
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(256u)
                 : "memory");
}

}  // namespace

// The default path never depends on this experimental kernel: if its attributes cannot be set
// the error is swallowed here and the option is refused later (run_zgemm_ozaki).
static bool g_ozaki_ready = false;

void init_kernels_ozaki() {
  cudaError_t e[4];
  e[0] = cudaFuncSetAttribute(k_zgemm_ozaki<double, 6, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              OzSmem<6>::kTotal);
  e[1] = cudaFuncSetAttribute(k_zgemm_ozaki<double, 7, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              OzSmem<6>::kTotal);
  e[2] = cudaFuncSetAttribute(k_zgemm_ozaki<float, 3, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              OzSmem<4>::kTotal);
  e[3] = cudaFuncSetAttribute(k_zgemm_ozaki<float, 4, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              OzSmem<4>::kTotal);
  cudaError_t f[4];
  f[0] = cudaFuncSetAttribute(k_zgemm_ozaki_kloop<double, 6, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              OzSmem<6>::kTotal);
  f[1] = cudaFuncSetAttribute(k_zgemm_ozaki_kloop<double, 7, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              OzSmem<6>::kTotal);
  f[2] = cudaFuncSetAttribute(k_zgemm_ozaki_kloop<float, 3, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              OzSmem<4>::kTotal);
  f[3] = cudaFuncSetAttribute(k_zgemm_ozaki_kloop<float, 4, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              OzSmem<4>::kTotal);
  g_ozaki_ready = true;
  for (cudaError_t x : f)
    if (x != cudaSuccess) {
      (void)cudaGetLastError();
      g_ozaki_ready = false;
    }
  for (cudaError_t x : e)
    if (x != cudaSuccess) {
      (void)cudaGetLastError();
      g_ozaki_ready = false;
    }
}

// ComplexF64: groups = 6 or 7 (option "zgemm_ozaki"); ComplexF32: groups = 3 or 4
// (option "cgemm_ozaki").  The caller brackets the launch with L.begin / L.end.
void run_zgemm_ozaki(const Launch& L, const FusedParams& fp, int groups, const void* A,
                     const void* B, void* C) {
  PQ_REQUIRE(g_ozaki_ready, PQ_ERR_UNSUPPORTED, "ozaki GEMM: kernel attributes could not be set");
  PQ_REQUIRE(zgemm_ozaki_eligible(fp.M, fp.N, fp.K), PQ_ERR_INVALID, "ozaki GEMM: K, N <= 64 only");
  if ((L.opt == nullptr || L.opt->ozaki_gen == 0) && groups == (L.elem_size == 16 ? 6 : 4)) {
    run_zgemm_ozaki_t(L, fp, A, B, C);   // second-generation kernel (kernels_zgemm_ozaki2.cu)
    return;
  }
  const long long tiles = (fp.M + OZ_TM - 1) / OZ_TM;
  const unsigned grid = (unsigned)(tiles < L.num_sms ? tiles : L.num_sms);
  if (L.elem_size == 16) {
    PQ_REQUIRE(groups == 6 || groups == 7, PQ_ERR_INVALID, "zgemm_ozaki must be 0, 6 or 7");
    if (groups == 6 && std::getenv("PQ_OZAKI_TRACE")) {
      cudaFuncSetAttribute(k_zgemm_ozaki<double, 6, 32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           OzSmem<6>::kTotal);
      k_zgemm_ozaki<double, 6, 32, true><<<grid, OZ_THREADS, OzSmem<6>::kTotal, L.stream>>>(
          (const double2*)A, (const double2*)B, (double2*)C, fp);
    } else if (groups == 6)
      k_zgemm_ozaki<double, 6, 32><<<grid, OZ_THREADS, OzSmem<6>::kTotal, L.stream>>>(
          (const double2*)A, (const double2*)B, (double2*)C, fp);
    else
      k_zgemm_ozaki<double, 7, 32><<<grid, OZ_THREADS, OzSmem<6>::kTotal, L.stream>>>(
          (const double2*)A, (const double2*)B, (double2*)C, fp);
  } else {
    PQ_REQUIRE(groups == 3 || groups == 4, PQ_ERR_INVALID, "cgemm_ozaki must be 0, 3 or 4");
    if (groups == 3)
      k_zgemm_ozaki<float, 3, 64><<<grid, OZ_THREADS, OzSmem<4>::kTotal, L.stream>>>(
          (const float2*)A, (const float2*)B, (float2*)C, fp);
    else
      k_zgemm_ozaki<float, 4, 64><<<grid, OZ_THREADS, OzSmem<4>::kTotal, L.stream>>>(
          (const float2*)A, (const float2*)B, (float2*)C, fp);
  }
}

// Long contractions on canonical layouts A[m + M k], B[n + N k] (see k_zgemm_ozaki_kloop).
// `ws` holds (M + N) ints for the row / column exponent fields (ContractPlan::ws_bytes).
template <class Real, int G, int NC>
static void launch_kloop(const Launch& L, const void* A, const void* B, void* C, int64_t M, int64_t N,
                         int64_t K, int* expA, int* expB) {
  using V2 = typename OzVec<Real>::type;
  const long long tiles = ((M + OZ_TM - 1) / OZ_TM) * ((N + NC - 1) / NC);
  const unsigned grid = (unsigned)(tiles < L.num_sms ? tiles : L.num_sms);
  k_zgemm_ozaki_kloop<Real, G, NC><<<grid, OZ_THREADS, OzSmem<oz::Traits<Real>::S>::kTotal, L.stream>>>(
      (const V2*)A, (const V2*)B, (V2*)C, M, N, K, expA, expB);
}

void run_zgemm_ozaki_kloop(const Launch& L, int groups, const void* A, const void* B, void* C,
                           int64_t M, int64_t N, int64_t K, void* ws) {
  PQ_REQUIRE(g_ozaki_ready, PQ_ERR_UNSUPPORTED, "ozaki GEMM: kernel attributes could not be set");
  PQ_REQUIRE(ws != nullptr && zgemm_ozaki_kloop_eligible(M, N, K), PQ_ERR_INVALID,
             "ozaki GEMM (long K): needs the plan's workspace and 64 < K <= 8192");
  int* expA = static_cast<int*>(ws);
  int* expB = expA + M;
  PQ_CUDA(cudaMemsetAsync(ws, 0, size_t(M + N) * sizeof(int), L.stream));
  const unsigned ksplit = (unsigned)(K >= 4096 ? 16 : K >= 1024 ? 4 : 1);
  const dim3 ga((unsigned)((M + 127) / 128), ksplit), gb((unsigned)((N + 127) / 128), ksplit);
  if (L.elem_size == 16) {
    PQ_REQUIRE(groups == 6 || groups == 7, PQ_ERR_INVALID, "zgemm_ozaki must be 0, 6 or 7");
    k_oz_row_exponents<double><<<ga, 128, 0, L.stream>>>((const double2*)A, M, K, expA);
    k_oz_row_exponents<double><<<gb, 128, 0, L.stream>>>((const double2*)B, N, K, expB);
    if (groups == 6)
      launch_kloop<double, 6, 32>(L, A, B, C, M, N, K, expA, expB);
    else
      launch_kloop<double, 7, 32>(L, A, B, C, M, N, K, expA, expB);
  } else {
    PQ_REQUIRE(groups == 3 || groups == 4, PQ_ERR_INVALID, "cgemm_ozaki must be 0, 3 or 4");
    k_oz_row_exponents<float><<<ga, 128, 0, L.stream>>>((const float2*)A, M, K, expA);
    k_oz_row_exponents<float><<<gb, 128, 0, L.stream>>>((const float2*)B, N, K, expB);
    if (groups == 3)
      launch_kloop<float, 3, 64>(L, A, B, C, M, N, K, expA, expB);
    else
      launch_kloop<float, 4, 64>(L, A, B, C, M, N, K, expA, expB);
  }
}

void run_cgemm_ozaki_fused(const Launch& L, const ContractPlan& cp, const void* A, const void* B,
                           void* C) {
  PQ_REQUIRE(L.elem_size == 8, PQ_ERR_INVALID, "fused ComplexF32 GEMM plans run on the INT8 kernel only");
  const int groups = (L.opt && L.opt->cgemm_ozaki != 0) ? L.opt->cgemm_ozaki : 4;
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
  run_zgemm_ozaki(L, fp, groups, A, B, C);
  L.end();
  PQ_CUDA(cudaGetLastError());
}

// pq_microbench back ends: "umma_i8_selftest" (wrong entries, 0 = pass), "umma_i8_tops_n32",
// "umma_i8_tops_n64" (int8 TOPS at the kernel's MMA shape / at N = 64)
double run_ozaki_microbench(const Launch& L, const std::string& what) {
  if (what == "ozaki_debug") {   // 0 = no watchdog event since the last call; else the wait id
    int rec[8] = {0};
    PQ_CUDA(cudaStreamSynchronize(L.stream));
    PQ_CUDA(cudaMemcpyFromSymbol(rec, g_oz_debug, sizeof(rec)));
    if (rec[0]) {
      std::fprintf(stderr, "ozaki watchdog: wait id %d (1 planes, 2 freed, 3 done, 4 probe) iteration %d "
                           "group %d block %d warp %d\n", rec[1], rec[2], rec[3], rec[4], rec[5]);
      int zero[8] = {0};
      PQ_CUDA(cudaMemcpyToSymbol(g_oz_debug, zero, sizeof(zero)));
      return rec[1];
    }
    return 0;
  }
  if (what == "ozaki_trace") {   // dumps block 0's phase stamps to $PQ_OZAKI_TRACE
    std::vector<long long> tr(16 * 64);
    PQ_CUDA(cudaStreamSynchronize(L.stream));
    PQ_CUDA(cudaMemcpyFromSymbol(tr.data(), g_oz_trace, tr.size() * sizeof(long long)));
    if (const char* path = std::getenv("PQ_OZAKI_TRACE"))
      if (FILE* f = std::fopen(path, "wb")) {
        std::fwrite(tr.data(), sizeof(long long), tr.size(), f);
        std::fclose(f);
      }
    return 0;
  }
  if (what == "umma_i8_selftest") {
    int* d = nullptr;
    PQ_CUDA(cudaMalloc(&d, OZ_TM * 32 * sizeof(int)));
    PQ_CUDA(cudaMemsetAsync(d, 0xff, OZ_TM * 32 * sizeof(int), L.stream));
    k_umma_i8_selftest<<<1, 128, 0, L.stream>>>(d);
    std::vector<int> got(OZ_TM * 32);
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
    for (int r = 0; r < OZ_TM; ++r)
      for (int c = 0; c < 32; ++c) {
        int want = 0;
        for (int k = 0; k < 32; ++k) want += oz_pat_a(r, k) * oz_pat_b(c, k);
        wrong += got[r * 32 + c] != want;
      }
    return wrong;
  }
  const bool n64 = what == "umma_i8_tops_n64";
  PQ_REQUIRE(n64 || what == "umma_i8_tops_n32", PQ_ERR_INVALID, "unknown microbench: " + what);
  const int iters = 4096, smem = OZ_A_PLANE + OZ_B_PLANE;
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
  const double macs = double(L.num_sms) * iters * 16.0 * OZ_TM * (n64 ? 64 : 32) * 32;
  return 2.0 * macs / (best * 1e-3) / 1e12;
}

}  // namespace pq
