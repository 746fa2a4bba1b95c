// Contraction kernels that do not go through a GEMM launch:
//
//  * k_contract_small  (K4/K5) -- one operand is small (gate tensor, input cap, outer
//    product factor): it is staged, already permuted, in shared memory; every thread
//    owns one row of the big operand, gathers its K elements straight from the
//    un-permuted tensor (no TTGT temporary), and writes NS outputs.  Reads the big
//    operand once and writes C once => HBM-bound, algorithmic bytes
//    (M*K + K*N + M*N) * sizeof(element).
//  * k_contract_dot    (K6) -- tiny output, long contraction (the final inner product of
//    a single-amplitude plan): split-K partial sums per CTA + a fixed-order second pass,
//    so the result is deterministic.
//  * k_contract_direct -- one thread per output element, gathers from both operands.
//    Used for tiny problems and as the always-correct cross-check path.
//  * k_gemm_simt       -- tiled SIMT complex GEMM on the canonical TTGT layouts; the
//    c64 GEMM until the tcgen05 kernel lands and the cross-check for the DMMA kernel.
//
// Reference semantics: src/layer1.jl:85-92 (tensorcontract), C axes = A-open ++ B-open.
#include "common.h"

namespace pq {

template <typename R> struct C2;
template <> struct C2<float> { using type = float2; };
template <> struct C2<double> { using type = double2; };

template <typename V, typename R>
__device__ __forceinline__ void cfma(V& acc, const V& a, const V& b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}

// ---------------------------------------------------------------------------
// small-operand fused kernel
// ---------------------------------------------------------------------------
struct SmallParams {
  IdxMap rmap;    // row index r of the big operand -> element offset in it
  IdxMap kbig;    // k -> element offset in the big operand
  IdxMap ksmall;  // k -> element offset in the small operand
  IdxMap ssmall;  // s (open index of the small operand) -> element offset in it
  long long R;    // rows of the big operand
  int K;
  int S;          // open extent of the small operand (<= NS)
  long long out_rs, out_ss;  // C offset = r * out_rs + s * out_ss
};

template <typename R, int NS>
__global__ void __launch_bounds__(128)
k_contract_small(const typename C2<R>::type* __restrict__ big,
                 const typename C2<R>::type* __restrict__ small,
                 typename C2<R>::type* __restrict__ out, const SmallParams p) {
  using V = typename C2<R>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V* Q = reinterpret_cast<V*>(smem_raw);                       // [K][NS]
  long long* koff = reinterpret_cast<long long*>(Q + (size_t)p.K * NS);  // [K]
  for (int i = threadIdx.x; i < p.K * NS; i += blockDim.x) {
    int k = i / NS, s = i - k * NS;
    V v;
    v.x = 0;
    v.y = 0;
    if (s < p.S) v = small[map_offset(p.ksmall, k) + map_offset(p.ssmall, s)];
    Q[i] = v;
  }
  for (int k = threadIdx.x; k < p.K; k += blockDim.x) koff[k] = map_offset(p.kbig, k);
  __syncthreads();

  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < p.R; r += stride) {
    const V* row = big + map_offset(p.rmap, r);
    V acc[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      acc[s].x = 0;
      acc[s].y = 0;
    }
#pragma unroll 4
    for (int k = 0; k < p.K; ++k) {
      V a = row[koff[k]];
#pragma unroll
      for (int s = 0; s < NS; ++s) cfma<V, R>(acc[s], a, Q[k * NS + s]);
    }
    V* dst = out + r * p.out_rs;
#pragma unroll
    for (int s = 0; s < NS; ++s)
      if (s < p.S) dst[s * p.out_ss] = acc[s];
  }
}

// ComplexF32, small operand on the right, lowest axis of the big operand open with even
// extent: every thread owns TWO adjacent rows, so all global traffic is 16-byte
// (LDG.128 of (re,im,re,im), STG.128 of two results).
template <int NS>
__global__ void __launch_bounds__(128)
k_contract_small_c64x2(const float2* __restrict__ big, const float2* __restrict__ small,
                       float2* __restrict__ out, const SmallParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* Q = reinterpret_cast<float2*>(smem_raw);                           // [K][NS]
  long long* koff = reinterpret_cast<long long*>(Q + (size_t)p.K * NS);      // [K]
  for (int i = threadIdx.x; i < p.K * NS; i += blockDim.x) {
    int k = i / NS, s = i - k * NS;
    float2 v = make_float2(0.f, 0.f);
    if (s < p.S) v = small[map_offset(p.ksmall, k) + map_offset(p.ssmall, s)];
    Q[i] = v;
  }
  for (int k = threadIdx.x; k < p.K; k += blockDim.x) koff[k] = map_offset(p.kbig, k);
  __syncthreads();
  const long long pairs = p.R >> 1;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < pairs; t += stride) {
    const float2* row = big + map_offset(p.rmap, 2 * t);
    float2 acc0[NS], acc1[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      acc0[s] = make_float2(0.f, 0.f);
      acc1[s] = make_float2(0.f, 0.f);
    }
#pragma unroll 4
    for (int k = 0; k < p.K; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(row + koff[k]);
      const float2 a0 = make_float2(a.x, a.y), a1 = make_float2(a.z, a.w);
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const float2 q = Q[k * NS + s];
        cfma<float2, float>(acc0[s], a0, q);
        cfma<float2, float>(acc1[s], a1, q);
      }
    }
    float2* dst = out + 2 * t;  // out_rs == 1
#pragma unroll
    for (int s = 0; s < NS; ++s)
      if (s < p.S)
        *reinterpret_cast<float4*>(dst + s * p.out_ss) =
            make_float4(acc0[s].x, acc0[s].y, acc1[s].x, acc1[s].y);
  }
}

// The same with FOUR adjacent rows per thread: 256-bit global accesses (LDG.E.ENL2.256 /
// STG.E.ENL2.256, sm_100), half the memory instructions per byte and twice the bytes in flight
// per thread -- the gate applications of QFT-26 in ComplexF32 ran at 5.5 TB/s with 128-bit
// accesses against 6.2 TB/s for their ComplexF64 twins.  NS <= 4 (register budget).
template <int NS>
__global__ void __launch_bounds__(128)
k_contract_small_c64x4(const float2* __restrict__ big, const float2* __restrict__ small,
                       float2* __restrict__ out, const SmallParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* Q = reinterpret_cast<float2*>(smem_raw);                           // [K][NS]
  long long* koff = reinterpret_cast<long long*>(Q + (size_t)p.K * NS);      // [K]
  for (int i = threadIdx.x; i < p.K * NS; i += blockDim.x) {
    int k = i / NS, s = i - k * NS;
    float2 v = make_float2(0.f, 0.f);
    if (s < p.S) v = small[map_offset(p.ksmall, k) + map_offset(p.ssmall, s)];
    Q[i] = v;
  }
  for (int k = threadIdx.x; k < p.K; k += blockDim.x) koff[k] = map_offset(p.kbig, k);
  __syncthreads();
  const long long quads = p.R >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < quads; t += stride) {
    const float2* row = big + map_offset(p.rmap, 4 * t);
    float2 acc[4][NS];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int s = 0; s < NS; ++s) acc[r][s] = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int k = 0; k < p.K; ++k) {
      float a[8];
      asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                   : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]), "=f"(a[4]), "=f"(a[5]), "=f"(a[6]), "=f"(a[7])
                   : "l"(row + koff[k]));
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const float2 q = Q[k * NS + s];
#pragma unroll
        for (int r = 0; r < 4; ++r) cfma<float2, float>(acc[r][s], make_float2(a[2 * r], a[2 * r + 1]), q);
      }
    }
    float2* dst = out + 4 * t;  // out_rs == 1
#pragma unroll
    for (int s = 0; s < NS; ++s)
      if (s < p.S)
        asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"l"(dst + s * p.out_ss),
                     "f"(acc[0][s].x), "f"(acc[0][s].y), "f"(acc[1][s].x), "f"(acc[1][s].y), "f"(acc[2][s].x),
                     "f"(acc[2][s].y), "f"(acc[3][s].x), "f"(acc[3][s].y)
                     : "memory");
  }
}

// ComplexF32, small operand on the right with a short open bond (S <= 8) over 16 < K <= 64, the
// big operand with a CONTRACTED axis fastest (the sweep's N = 8 steps: the boundary absorbs a
// site tensor whose bond bits sit at the lowest addresses).  One row per thread makes every lane
// of a load touch its own 128-byte line (k_contract_small: 70 us = 2.1 TB/s on R = 2^18, S = 8,
// K = 64; an FP32-FMA variant with lanes over k re-read the small operand from shared memory
// for every row block and was slower still).  Here a warp owns 16 rows per step and multiplies
// on the legacy tensor path, mma.sync.m16n8k8.tf32 with 3xTF32 splitting (hi * hi + hi * lo +
// lo * hi, 12 MMAs per 8 k's for the four real products): the A fragment IS the load layout --
// lane (g, t) holds rows g, g + 8 and k = 8 s + t, 8 s + t + 4, so a quarter warp reads 2 rows x
// 32 contiguous bytes -- and all loads of a step are issued before the first MMA.
__device__ __forceinline__ float2 tcs_load(const float2* ptr, bool pred) {
  float2 v;
  asm volatile(
      "{\n .reg .pred q;\n setp.ne.b32 q, %3, 0;\n mov.f32 %0, 0f00000000;\n mov.f32 %1, 0f00000000;\n"
      " @q ld.global.nc.v2.f32 {%0, %1}, [%2];\n}\n"
      : "=f"(v.x), "=f"(v.y)
      : "l"(ptr), "r"((int)pred));
  return v;
}
__device__ __forceinline__ uint32_t tcs_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void tcs_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int KS, int NB>   // K <= 8 KS, S <= 8 NB
__global__ void __launch_bounds__(256, 2)
k_contract_small_c64tc(const float2* __restrict__ big, const float2* __restrict__ small,
                       float2* __restrict__ out, const SmallParams p) {
  constexpr int KP = 8 * KS, NP = 8 * NB;
  constexpr int QP = NB > 1 ? NP + 8 : NP;   // pitch = 8 (mod 32) words: conflict-free fragment reads
  // the small operand, split and signed once per CTA: [variant][k][n], variants Br_hi, Br_lo,
  // Bi_hi, Bi_lo, -Bi_hi, -Bi_lo (TF32 bit patterns)
  __shared__ uint32_t Q[6][KP][QP];
  __shared__ int koff[KP];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  for (int i = tid; i < KP * NP; i += 256) {
    const int k = i / NP, n = i % NP;
    float2 v = make_float2(0.f, 0.f);
    if (k < p.K && n < p.S) v = small[map_offset(p.ksmall, k) + map_offset(p.ssmall, n)];
    const uint32_t rh = tcs_tf32(v.x), ih = tcs_tf32(v.y);
    const uint32_t rl = tcs_tf32(v.x - __uint_as_float(rh)), il = tcs_tf32(v.y - __uint_as_float(ih));
    Q[0][k][n] = rh;
    Q[1][k][n] = rl;
    Q[2][k][n] = ih;
    Q[3][k][n] = il;
    Q[4][k][n] = ih ^ 0x80000000u;
    Q[5][k][n] = il ^ 0x80000000u;
  }
  for (int k = tid; k < KP; k += 256) koff[k] = k < p.K ? (int)map_offset(p.kbig, k) : -1;
  __syncthreads();
  int ko[2 * KS];   // k = 8 s + t, 8 s + t + 4
#pragma unroll
  for (int s2 = 0; s2 < KS; ++s2) {
    ko[2 * s2] = koff[8 * s2 + t];
    ko[2 * s2 + 1] = koff[8 * s2 + t + 4];
  }
  const long long blocks = (p.R + 15) / 16, stride = (long long)gridDim.x * 8;
  for (long long blk = (long long)blockIdx.x * 8 + warp; blk < blocks; blk += stride) {
    const long long r0 = blk * 16 + g, r1 = r0 + 8;
    const bool v0 = r0 < p.R, v1 = r1 < p.R;
    const float2* row0 = big + (v0 ? map_offset(p.rmap, r0) : 0);
    const float2* row1 = big + (v1 ? map_offset(p.rmap, r1) : 0);
    float2 a[KS][4];   // fragment order: (g, t), (g + 8, t), (g, t + 4), (g + 8, t + 4)
#pragma unroll
    for (int s2 = 0; s2 < KS; ++s2) {
      a[s2][0] = tcs_load(row0 + ko[2 * s2], v0 && ko[2 * s2] >= 0);
      a[s2][1] = tcs_load(row1 + ko[2 * s2], v1 && ko[2 * s2] >= 0);
      a[s2][2] = tcs_load(row0 + ko[2 * s2 + 1], v0 && ko[2 * s2 + 1] >= 0);
      a[s2][3] = tcs_load(row1 + ko[2 * s2 + 1], v1 && ko[2 * s2 + 1] >= 0);
    }
    float cr[NB][4], ci[NB][4];
#pragma unroll
    for (int y = 0; y < NB; ++y)
#pragma unroll
      for (int j = 0; j < 4; ++j) cr[y][j] = ci[y][j] = 0.f;
#pragma unroll
    for (int s2 = 0; s2 < KS; ++s2) {
      uint32_t rh[4], rl[4], ih[4], il[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        rh[j] = tcs_tf32(a[s2][j].x);
        ih[j] = tcs_tf32(a[s2][j].y);
        rl[j] = tcs_tf32(a[s2][j].x - __uint_as_float(rh[j]));
        il[j] = tcs_tf32(a[s2][j].y - __uint_as_float(ih[j]));
      }
#pragma unroll
      for (int y = 0; y < NB; ++y) {
        uint32_t b[6][2];
#pragma unroll
        for (int v = 0; v < 6; ++v) {
          b[v][0] = Q[v][8 * s2 + t][8 * y + g];
          b[v][1] = Q[v][8 * s2 + t + 4][8 * y + g];
        }
        // small terms first
        tcs_mma(cr[y], rl, b[0][0], b[0][1]);   // Ar_lo Br_hi
        tcs_mma(cr[y], rh, b[1][0], b[1][1]);   // Ar_hi Br_lo
        tcs_mma(cr[y], il, b[4][0], b[4][1]);   // Ai_lo (-Bi_hi)
        tcs_mma(cr[y], ih, b[5][0], b[5][1]);   // Ai_hi (-Bi_lo)
        tcs_mma(ci[y], rl, b[2][0], b[2][1]);   // Ar_lo Bi_hi
        tcs_mma(ci[y], rh, b[3][0], b[3][1]);   // Ar_hi Bi_lo
        tcs_mma(ci[y], il, b[0][0], b[0][1]);   // Ai_lo Br_hi
        tcs_mma(ci[y], ih, b[1][0], b[1][1]);   // Ai_hi Br_lo
        tcs_mma(cr[y], rh, b[0][0], b[0][1]);   // Ar_hi Br_hi
        tcs_mma(cr[y], ih, b[4][0], b[4][1]);   // Ai_hi (-Bi_hi)
        tcs_mma(ci[y], rh, b[2][0], b[2][1]);   // Ar_hi Bi_hi
        tcs_mma(ci[y], ih, b[0][0], b[0][1]);   // Ai_hi Br_hi
      }
    }
    // C fragment: (row g, cols 2 t, 2 t + 1), (row g + 8, cols 2 t, 2 t + 1) of every n block
#pragma unroll
    for (int y = 0; y < NB; ++y)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int n = 8 * y + 2 * t + j;
        if (n < p.S) {
          if (v0) out[r0 * p.out_rs + n * p.out_ss] = make_float2(cr[y][j], ci[y][j]);
          if (v1) out[r1 * p.out_rs + n * p.out_ss] = make_float2(cr[y][2 + j], ci[y][2 + j]);
        }
      }
  }
}

static int64_t map_min_stride(const IdxMap& m) {
  int64_t best = INT64_MAX;
  for (int d = 0; d < m.nd; ++d) best = m.str[d] < best ? m.str[d] : best;
  return best;
}

// ComplexF32, many rows, and either a short bond over a longer contraction with a contracted
// axis fastest in the big operand (S <= 8, 16 < K <= 64), or a short contraction with up to 64
// open on the small side (K <= 16, 16 < S <= 64: output-bound; lower.cpp sends those here
// instead of to the INT8 kernel)
static bool small_c64tc_ok(const SmallParams& p) {
  if (p.R < 4096 || p.kbig.nd < 1 || p.rmap.nd < 1) return false;
  if (p.S > 16 && p.S <= 64 && p.K <= 16) return true;
  return p.S > 4 && p.S <= 8 && p.K > 16 && p.K <= 64 && map_min_stride(p.kbig) < map_min_stride(p.rmap);
}

static void launch_small_c64tc(const Launch& L, const SmallParams& p, const void* big, const void* small,
                               void* out) {
  const unsigned grid = (unsigned)(2 * L.num_sms);
  const float2 *b = (const float2*)big, *s = (const float2*)small;
  float2* o = (float2*)out;
  if (p.S > 8) {
    if (p.K <= 8) {
      if (p.S <= 32) k_contract_small_c64tc<1, 4><<<grid, 256, 0, L.stream>>>(b, s, o, p);
      else k_contract_small_c64tc<1, 8><<<grid, 256, 0, L.stream>>>(b, s, o, p);
    } else {
      if (p.S <= 32) k_contract_small_c64tc<2, 4><<<grid, 256, 0, L.stream>>>(b, s, o, p);
      else k_contract_small_c64tc<2, 8><<<grid, 256, 0, L.stream>>>(b, s, o, p);
    }
  } else if (p.K <= 32) {
    k_contract_small_c64tc<4, 1><<<grid, 256, 0, L.stream>>>(b, s, o, p);
  } else {
    k_contract_small_c64tc<8, 1><<<grid, 256, 0, L.stream>>>(b, s, o, p);
  }
}

template <int NS>
static void launch_small_c64x4(const Launch& L, const SmallParams& p, const void* big,
                               const void* small, void* out) {
  size_t smem = (size_t)p.K * NS * sizeof(float2) + (size_t)p.K * sizeof(long long);
  long long blocks = (p.R / 4 + 127) / 128;
  long long cap = (long long)L.num_sms * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k_contract_small_c64x4<NS><<<(unsigned)blocks, 128, smem, L.stream>>>(
      (const float2*)big, (const float2*)small, (float2*)out, p);
}

template <int NS>
static void launch_small_c64x2(const Launch& L, const SmallParams& p, const void* big,
                               const void* small, void* out) {
  size_t smem = (size_t)p.K * NS * sizeof(float2) + (size_t)p.K * sizeof(long long);
  long long blocks = (p.R / 2 + 127) / 128;
  long long cap = (long long)L.num_sms * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k_contract_small_c64x2<NS><<<(unsigned)blocks, 128, smem, L.stream>>>(
      (const float2*)big, (const float2*)small, (float2*)out, p);
}

template <typename R, int NS>
static void launch_small_ns(const Launch& L, const SmallParams& p, const void* big,
                            const void* small, void* out) {
  using V = typename C2<R>::type;
  size_t smem = (size_t)p.K * NS * sizeof(V) + (size_t)p.K * sizeof(long long);
  long long blocks = (p.R + 127) / 128;
  long long cap = (long long)L.num_sms * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k_contract_small<R, NS><<<(unsigned)blocks, 128, smem, L.stream>>>((const V*)big, (const V*)small,
                                                                   (V*)out, p);
}

// 16-byte path condition: c64, result rows contiguous (small operand on the right), the
// fastest axis of the big operand is an open axis of even extent (then every other stride
// of that tensor, the k offsets included, is even and all accesses are 16-byte aligned)
static bool small_vec2_ok(const SmallParams& p) {
  if (p.out_rs != 1 || (p.R & 1) || (p.out_ss & 1) || p.rmap.nd < 1) return false;
  if (p.rmap.str[0] != 1 || (p.rmap.ext[0] & 1)) return false;
  for (int d = 1; d < p.rmap.nd; ++d)
    if (p.rmap.str[d] & 1) return false;
  for (int d = 0; d < p.kbig.nd; ++d)
    if (p.kbig.str[d] & 1) return false;
  return true;
}

// 32-byte path condition: the same with extent and strides multiples of 4
static bool small_vec4_ok(const SmallParams& p) {
  if (!small_vec2_ok(p) || (p.R & 3) || (p.out_ss & 3) || (p.rmap.ext[0] & 3) || p.S > 4) return false;
  for (int d = 1; d < p.rmap.nd; ++d)
    if (p.rmap.str[d] & 3) return false;
  for (int d = 0; d < p.kbig.nd; ++d)
    if (p.kbig.str[d] & 3) return false;
  return true;
}

template <typename R>
static void launch_small(const Launch& L, const SmallParams& p, const void* big, const void* small,
                         void* out) {
  if (sizeof(R) == 4 && small_c64tc_ok(p) && !(L.opt && L.opt->small_tc == 1)) {
    launch_small_c64tc(L, p, big, small, out);
    return;
  }
  PQ_REQUIRE(p.S <= 16, PQ_ERR_UNSUPPORTED, "small-operand kernel: more than 16 open elements on the small side");
  if (sizeof(R) == 4 && small_vec4_ok(p) && ((uintptr_t)big % 32 == 0) && ((uintptr_t)out % 32 == 0) &&
      p.R >= (1 << 16)) {
    if (p.S <= 1)
      launch_small_c64x4<1>(L, p, big, small, out);
    else if (p.S <= 2)
      launch_small_c64x4<2>(L, p, big, small, out);
    else
      launch_small_c64x4<4>(L, p, big, small, out);
    return;
  }
  if (sizeof(R) == 4 && small_vec2_ok(p) && ((uintptr_t)big % 16 == 0) &&
      ((uintptr_t)out % 16 == 0)) {
    if (p.S <= 1)
      launch_small_c64x2<1>(L, p, big, small, out);
    else if (p.S <= 2)
      launch_small_c64x2<2>(L, p, big, small, out);
    else if (p.S <= 4)
      launch_small_c64x2<4>(L, p, big, small, out);
    else if (p.S <= 8)
      launch_small_c64x2<8>(L, p, big, small, out);
    else
      launch_small_c64x2<16>(L, p, big, small, out);
    return;
  }
  if (p.S <= 1)
    launch_small_ns<R, 1>(L, p, big, small, out);
  else if (p.S <= 2)
    launch_small_ns<R, 2>(L, p, big, small, out);
  else if (p.S <= 4)
    launch_small_ns<R, 4>(L, p, big, small, out);
  else if (p.S <= 8)
    launch_small_ns<R, 8>(L, p, big, small, out);
  else
    launch_small_ns<R, 16>(L, p, big, small, out);
}

// ---------------------------------------------------------------------------
// direct kernel
// ---------------------------------------------------------------------------
struct DirectParams {
  IdxMap mA, kA, nB, kB;
  long long M, N, K;
};

template <typename R>
__global__ void __launch_bounds__(128)
k_contract_direct(const typename C2<R>::type* __restrict__ A,
                  const typename C2<R>::type* __restrict__ B,
                  typename C2<R>::type* __restrict__ C, const DirectParams p) {
  using V = typename C2<R>::type;
  const long long total = p.M * p.N;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += stride) {
    long long n = c / p.M, m = c - n * p.M;
    const V* a = A + map_offset(p.mA, m);
    const V* b = B + map_offset(p.nB, n);
    V acc;
    acc.x = 0;
    acc.y = 0;
    for (long long k = 0; k < p.K; ++k)
      cfma<V, R>(acc, a[map_offset(p.kA, k)], b[map_offset(p.kB, k)]);
    C[c] = acc;
  }
}

// ---------------------------------------------------------------------------
// chains of tiny contractions: one CTA executes a list of contractions one after another.
// A slice of an RQC amplitude starts with ~10^3 contractions of a few dozen elements (every
// qubit's world-line collapses gate by gate); each is one dependent step of a chain, so as
// separate launches they cost a kernel dispatch apiece.  Here a chain is a single CTA walking
// its descriptors (copied to shared memory one at a time), chains that become ready together
// share one launch, and the intermediate tensors stay in L2.  Operands are read with ld.cg:
// they were written by other threads of the same CTA a barrier earlier.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int mini_offset(const MiniMap& m, int i) {
  int off = 0;
  if (m.pow2) {
    for (int d = 0; d < m.nd; ++d) {
      off += (i & (m.ext[d] - 1)) * m.str[d];
      i >>= m.sh[d];
    }
  } else {
    for (int d = 0; d < m.nd; ++d) {
      const int q = i / m.ext[d];
      off += (i - q * m.ext[d]) * m.str[d];
      i = q;
    }
  }
  return off;
}

constexpr int CHAIN_THREADS = 256;

template <typename R>
__global__ void __launch_bounds__(CHAIN_THREADS)
k_contract_chain(const ChainItem* __restrict__ items, const ChainRange* __restrict__ ranges) {
  using V = typename C2<R>::type;
  constexpr int WORDS = (int)(sizeof(ChainItem) / sizeof(int));
  constexpr int PER = (WORDS + CHAIN_THREADS - 1) / CHAIN_THREADS;   // descriptor words per thread
  __shared__ ChainItem it;
  const ChainRange rg = ranges[blockIdx.x];
  // the descriptor of item s + 1 is fetched into registers while item s is computed, so
  // only the operand round trips through L2 remain on the chain's critical path
  int pre[PER];
  {
    const int* src = reinterpret_cast<const int*>(items + rg.begin);
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int i = threadIdx.x + q * CHAIN_THREADS;
      pre[q] = (i < WORDS && rg.count > 0) ? src[i] : 0;
    }
  }
  for (int s = 0; s < rg.count; ++s) {
    __syncthreads();   // results of the previous item are visible; `it` may be overwritten
    {
      int* dst = reinterpret_cast<int*>(&it);
#pragma unroll
      for (int q = 0; q < PER; ++q) {
        const int i = threadIdx.x + q * CHAIN_THREADS;
        if (i < WORDS) dst[i] = pre[q];
      }
    }
    __syncthreads();
    if (s + 1 < rg.count) {
      const int* src = reinterpret_cast<const int*>(items + rg.begin + s + 1);
#pragma unroll
      for (int q = 0; q < PER; ++q) {
        const int i = threadIdx.x + q * CHAIN_THREADS;
        if (i < WORDS) pre[q] = src[i];
      }
    }
    const V* A = static_cast<const V*>(it.A);
    const V* B = static_cast<const V*>(it.B);
    V* C = static_cast<V*>(it.C);
    const int total = it.M * it.N;
    for (int c = threadIdx.x; c < total; c += blockDim.x) {
      const int n = c / it.M, m = c - n * it.M;
      const V* a = A + mini_offset(it.mA, m);
      const V* b = B + mini_offset(it.nB, n);
      V acc;
      acc.x = 0;
      acc.y = 0;
      for (int k = 0; k < it.K; ++k)
        cfma<V, R>(acc, __ldcg(a + mini_offset(it.kA, k)), __ldcg(b + mini_offset(it.kB, k)));
      C[c] = acc;
    }
  }
}

static bool to_mini(const IdxMap& m, MiniMap& o) {
  if (m.nd > MINI_ND) return false;
  o.nd = m.nd;
  o.pow2 = m.pow2;
  for (int d = 0; d < MINI_ND; ++d) {
    o.sh[d] = 0;
    o.ext[d] = 1;
    o.str[d] = 0;
  }
  for (int d = 0; d < m.nd; ++d) {
    if (m.ext[d] >= (int64_t(1) << 30) || m.str[d] >= (int64_t(1) << 30) || m.str[d] < 0) return false;
    o.sh[d] = m.sh[d];
    o.ext[d] = (int)m.ext[d];
    o.str[d] = (int)m.str[d];
  }
  return true;
}

bool chain_item_from_plan(const ContractPlan& p, ChainItem& it) {
  if (p.M * p.N >= (int64_t(1) << 24) || p.K >= (int64_t(1) << 24)) return false;
  if (!to_mini(p.mA, it.mA) || !to_mini(p.kA, it.kA) || !to_mini(p.nB, it.nB) ||
      !to_mini(p.kB, it.kB))
    return false;
  it.M = (int)p.M;
  it.N = (int)p.N;
  it.K = (int)p.K;
  it.pad = 0;
  it.A = it.B = nullptr;
  it.C = nullptr;
  return true;
}

void run_chains(const Launch& L, const ChainItem* d_items, const ChainRange* d_ranges, int nchains,
                double bytes, double flops) {
  if (nchains <= 0) return;
  L.begin(KC_CONTRACT_CHAIN, bytes, flops);
  if (L.elem_size == 16)
    k_contract_chain<double><<<nchains, CHAIN_THREADS, 0, L.stream>>>(d_items, d_ranges);
  else
    k_contract_chain<float><<<nchains, CHAIN_THREADS, 0, L.stream>>>(d_items, d_ranges);
  L.end();
  PQ_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// dot kernel (M*N <= 16, long K)
// ---------------------------------------------------------------------------
struct DotParams {
  IdxMap mA, kA, nB, kB;
  int M, N;
  long long K;
  int split;   // > 0 (M = N = 1, power-of-two extents, 2^split threads): k = thread + j * 2^split
};

// MNT = compile-time bound on M*N (1, 4 or 16), U = independent k-iterations in flight per
// thread (the loads of one operand are usually scattered, so memory-level parallelism, not
// arithmetic, sets the speed).  The partial sums are combined in a fixed order
// (thread-strided k, shuffle tree, warp order, then k_dot_finish), so the result does not
// depend on scheduling.
template <typename R, int MNT, int U>
__global__ void __launch_bounds__(256)
k_contract_dot(const typename C2<R>::type* __restrict__ A,
               const typename C2<R>::type* __restrict__ B,
               typename C2<R>::type* __restrict__ partial, const DotParams p) {
  using V = typename C2<R>::type;
  __shared__ long long offA[16], offB[16];
  __shared__ V red[8][16];
  const int MN = p.M * p.N;
  if (threadIdx.x < MN) {
    int n = threadIdx.x / p.M, m = threadIdx.x - n * p.M;
    offA[threadIdx.x] = map_offset(p.mA, m);
    offB[threadIdx.x] = map_offset(p.nB, n);
  }
  __syncthreads();
  V acc[MNT];
#pragma unroll
  for (int j = 0; j < MNT; ++j) {
    acc[j].x = 0;
    acc[j].y = 0;
  }
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (MNT == 1 && p.split > 0) {
    // offset(k_low + j 2^s) = offset(k_low) + offset(j 2^s): one index-map walk per thread, the
    // j part from a table (<= 64 entries per operand)
    __shared__ long long tabA[64], tabB[64];
    const int J = (int)(p.K >> p.split);
    if (threadIdx.x < J) {
      tabA[threadIdx.x] = map_offset(p.kA, (long long)threadIdx.x << p.split);
      tabB[threadIdx.x] = map_offset(p.kB, (long long)threadIdx.x << p.split);
    }
    __syncthreads();
    const V* a0 = A + map_offset(p.kA, k) + offA[0];
    const V* b0 = B + map_offset(p.kB, k) + offB[0];
    for (int j = 0; j < J; j += 8) {
      V a[8], b[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        a[u] = a0[tabA[j + u]];
        b[u] = b0[tabB[j + u]];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) cfma<V, R>(acc[0], a[u], b[u]);
    }
    k = p.K;   // nothing left for the generic loops
  } else if (MNT == 1) {
    const long long oa = offA[0], ob = offB[0];
    for (; k + (U - 1) * stride < p.K; k += U * stride) {
      V a[U], b[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        a[u] = A[map_offset(p.kA, k + u * stride) + oa];
        b[u] = B[map_offset(p.kB, k + u * stride) + ob];
      }
#pragma unroll
      for (int u = 0; u < U; ++u) cfma<V, R>(acc[0], a[u], b[u]);
    }
  }
  for (; k < p.K; k += stride) {
    const V* a = A + map_offset(p.kA, k);
    const V* b = B + map_offset(p.kB, k);
#pragma unroll
    for (int j = 0; j < MNT; ++j)
      if (j < MN) cfma<V, R>(acc[j], a[offA[j]], b[offB[j]]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < MNT; ++j) {
    if (j < MN) {
      V v = acc[j];
      for (int d = 16; d > 0; d >>= 1) {
        v.x += __shfl_down_sync(0xffffffffu, v.x, d);
        v.y += __shfl_down_sync(0xffffffffu, v.y, d);
      }
      if (lane == 0) red[warp][j] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < MN) {
    V v = red[0][threadIdx.x];
    for (int w = 1; w < 8; ++w) {
      v.x += red[w][threadIdx.x].x;
      v.y += red[w][threadIdx.x].y;
    }
    partial[(long long)blockIdx.x * 16 + threadIdx.x] = v;
  }
}

// 64 groups of 16 threads: group g sums the partials of blocks g, g+64, ... for output j,
// then a fixed-order pass over the 64 group sums.
template <typename R>
__global__ void __launch_bounds__(1024)
k_dot_finish(const typename C2<R>::type* __restrict__ partial,
             typename C2<R>::type* __restrict__ C, int MN, int blocks) {
  using V = typename C2<R>::type;
  __shared__ V part[64][16];
  const int j = threadIdx.x & 15, g = threadIdx.x >> 4;
  V v;
  v.x = 0;
  v.y = 0;
  if (j < MN) {
#pragma unroll 4
    for (int b = g; b < blocks; b += 64) {
      const V q = partial[(long long)b * 16 + j];
      v.x += q.x;
      v.y += q.y;
    }
  }
  part[g][j] = v;
  __syncthreads();
  if (threadIdx.x < MN) {
    V t = part[0][threadIdx.x];
    for (int q = 1; q < 64; ++q) {
      t.x += part[q][threadIdx.x].x;
      t.y += part[q][threadIdx.x].y;
    }
    C[threadIdx.x] = t;
  }
}

// ---------------------------------------------------------------------------
// SIMT GEMM on canonical layouts: A[m + M k], B[n + N k], C[m + M n]
// ---------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(256)
k_gemm_simt(const typename C2<R>::type* __restrict__ A, const typename C2<R>::type* __restrict__ B,
            typename C2<R>::type* __restrict__ C, long long M, long long N, long long K) {
  using V = typename C2<R>::type;
  constexpr int BM = 64, BN = 64, BK = 8;
  __shared__ V As[BK][BM];
  __shared__ V Bs[BK][BN];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long m0 = (long long)blockIdx.x * BM, n0 = (long long)blockIdx.y * BN;
  V acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[i][j].x = 0;
      acc[i][j].y = 0;
    }
  V zero;
  zero.x = 0;
  zero.y = 0;
  for (long long k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      int idx = threadIdx.x + q * 256;  // 0..511
      int kk = idx >> 6, mm = idx & 63;
      long long k = k0 + kk;
      As[kk][mm] = (k < K && m0 + mm < M) ? A[(m0 + mm) + M * k] : zero;
      Bs[kk][mm] = (k < K && n0 + mm < N) ? B[(n0 + mm) + N * k] : zero;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      V a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][tx + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][ty + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) cfma<V, R>(acc[i][j], a[i], b[j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    long long n = n0 + ty + 16 * j;
    if (n >= N) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      long long m = m0 + tx + 16 * i;
      if (m < M) C[m + M * n] = acc[i][j];
    }
  }
}

void run_gemm_simt(const Launch& L, const void* A, const void* B, void* C, int64_t M, int64_t N,
                   int64_t K) {
  dim3 grid((unsigned)((M + 63) / 64), (unsigned)((N + 63) / 64));
  PQ_REQUIRE(grid.y <= 65535, PQ_ERR_UNSUPPORTED, "N too large for the SIMT GEMM grid");
  double bytes = double(M * K + N * K + M * N) * L.elem_size, flops = 8.0 * M * N * K;
  L.begin(KC_GEMM_SIMT, bytes, flops);
  if (L.elem_size == 16)
    k_gemm_simt<double><<<grid, 256, 0, L.stream>>>((const double2*)A, (const double2*)B,
                                                    (double2*)C, M, N, K);
  else
    k_gemm_simt<float><<<grid, 256, 0, L.stream>>>((const float2*)A, (const float2*)B, (float2*)C,
                                                   M, N, K);
  L.end();
  PQ_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// dispatcher
// ---------------------------------------------------------------------------
template <typename R>
static void run_contract_t(const Launch& L, const ContractPlan& p, const void* A, const void* B,
                           void* C, void* tempA, void* tempB, void* ws) {
  using V = typename C2<R>::type;
  const double bytes = double(p.M * p.K + p.K * p.N + p.M * p.N) * sizeof(V);
  const double flops = 8.0 * double(p.M) * double(p.N) * double(p.K);
  switch (p.kind) {
    case CK_SMALL_RIGHT: {
      SmallParams sp;
      sp.rmap = p.mA;
      sp.kbig = p.kA;
      sp.ksmall = p.kB;
      sp.ssmall = p.nB;
      sp.R = p.M;
      sp.K = (int)p.K;
      sp.S = (int)p.N;
      sp.out_rs = 1;
      sp.out_ss = p.M;
      L.begin(KC_CONTRACT_SMALL, bytes, flops);
      launch_small<R>(L, sp, A, B, C);
      L.end();
      break;
    }
    case CK_SMALL_LEFT: {
      SmallParams sp;
      sp.rmap = p.nB;
      sp.kbig = p.kB;
      sp.ksmall = p.kA;
      sp.ssmall = p.mA;
      sp.R = p.N;
      sp.K = (int)p.K;
      sp.S = (int)p.M;
      sp.out_rs = p.M;
      sp.out_ss = 1;
      L.begin(KC_CONTRACT_SMALL, bytes, flops);
      launch_small<R>(L, sp, B, A, C);
      L.end();
      break;
    }
    case CK_DOT: {
      DotParams dp;
      dp.mA = p.mA;
      dp.kA = p.kA;
      dp.nB = p.nB;
      dp.kB = p.kB;
      dp.M = (int)p.M;
      dp.N = (int)p.N;
      dp.K = p.K;
      dp.split = p.dot_split;
      L.begin(KC_CONTRACT_DOT, bytes, flops);
      if (p.M * p.N == 1)
        k_contract_dot<R, 1, 4><<<p.dot_blocks, 256, 0, L.stream>>>((const V*)A, (const V*)B, (V*)ws, dp);
      else if (p.M * p.N <= 4)
        k_contract_dot<R, 4, 1><<<p.dot_blocks, 256, 0, L.stream>>>((const V*)A, (const V*)B, (V*)ws, dp);
      else
        k_contract_dot<R, 16, 1><<<p.dot_blocks, 256, 0, L.stream>>>((const V*)A, (const V*)B, (V*)ws, dp);
      L.end();
      L.begin(KC_CONTRACT_DOT, 0, 0);
      k_dot_finish<R><<<1, 1024, 0, L.stream>>>((const V*)ws, (V*)C, (int)(p.M * p.N), p.dot_blocks);
      L.end();
      break;
    }
    case CK_DIRECT: {
      DirectParams dp;
      dp.mA = p.mA;
      dp.kA = p.kA;
      dp.nB = p.nB;
      dp.kB = p.kB;
      dp.M = p.M;
      dp.N = p.N;
      dp.K = p.K;
      long long total = p.M * p.N;
      long long blocks = (total + 127) / 128;
      long long cap = (long long)L.num_sms * 32;
      if (blocks > cap) blocks = cap;
      if (blocks < 1) blocks = 1;
      L.begin(KC_CONTRACT_DIRECT, bytes, flops);
      k_contract_direct<R><<<(unsigned)blocks, 128, 0, L.stream>>>((const V*)A, (const V*)B, (V*)C,
                                                                  dp);
      L.end();
      break;
    }
    case CK_GEMM: {
      if (p.fused_gemm) {
        if (sizeof(R) == 8)
          run_zgemm_fused(L, p, A, B, C);
        else
          run_cgemm_ozaki_fused(L, p, A, B, C);   // ComplexF32 skinny steps: k_ozaki_t, gather fused
        break;
      }
      const void* Ap = A;
      const void* Bp = B;
      if (!p.permA.identity) {
        run_permute(L, p.permA, A, tempA);
        Ap = tempA;
      }
      if (!p.permB.identity) {
        run_permute(L, p.permB, B, tempB);
        Bp = tempB;
      }
      const bool tensor = (L.opt == nullptr || L.opt->gemm != 1);
      if (tensor && sizeof(R) == 8)
        run_zgemm_dmma(L, Ap, Bp, C, p.M, p.N, p.K);      // FP64 DMMA
      else if (tensor)
        run_cgemm_tcgen05(L, Ap, Bp, C, p.M, p.N, p.K);   // tcgen05, 3xTF32
      else
        run_gemm_simt(L, Ap, Bp, C, p.M, p.N, p.K);
      break;
    }
  }
  PQ_CUDA(cudaGetLastError());
}

void run_contract(const Launch& L, const ContractPlan& p, const void* A, const void* B, void* C,
                  void* tempA, void* tempB, void* ws) {
  if (L.elem_size == 16)
    run_contract_t<double>(L, p, A, B, C, tempA, tempB, ws);
  else
    run_contract_t<float>(L, p, A, B, C, tempA, tempB, ws);
}

}  // namespace pq
