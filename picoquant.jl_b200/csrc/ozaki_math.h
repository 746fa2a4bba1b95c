// Integer arithmetic of the INT8 Ozaki-scheme complex GEMM (kernels_zgemm_ozaki2.cu), shared
// between the CUDA kernel and the host-side emulation in test_lower.cpp, so that the slicing,
// the packing into the UMMA core-matrix layout and the recombination are checked without a GPU
// (the same role tile_math.h plays for the permute kernel).
//
// A real x of a row whose largest magnitude has biased exponent field `ef` becomes
// q = rint(x * 2^(QBITS - E)) with |x| < 2^E, |q| <= 2^QBITS, written in balanced base 256:
// q = sum_i d_i 256^i with d_i in [-128, 127] (int8).  The digits are the BYTES of q + BIAS
// (BIAS = sum_i 128 * 256^i) with their top bit flipped, so there are no carries and no
// bit-field shuffling.  Plane s holds digit 256^(S-1-s).
//   double: S = 6 digits, QBITS = 46      float: S = 4 digits, QBITS = 30
// (one bit of headroom below the bias in both cases).
#pragma once
#include <stdint.h>

#include <cmath>
#include <cstring>
#ifdef __CUDACC__
#define OZ_HD __host__ __device__ __forceinline__
#else
#define OZ_HD inline
#endif

namespace pq {
namespace oz {

struct Word4 {
  uint32_t w[4];
};

OZ_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
#ifdef __CUDA_ARCH__
  return __byte_perm(a, b, sel);
#else
  const unsigned long long ab = ((unsigned long long)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) {
    const uint32_t n = (sel >> (4 * i)) & 7u;
    r |= (uint32_t)((ab >> (8 * n)) & 0xFFu) << (8 * i);
  }
  return r;
#endif
}
OZ_HD long long d2ll_rn(double x) {
#ifdef __CUDA_ARCH__
  return __double2ll_rn(x);
#else
  return std::llrint(x);   // round-to-nearest-even in the default rounding mode
#endif
}
OZ_HD int f2i_rn(float x) {
#ifdef __CUDA_ARCH__
  return __float2int_rn(x);
#else
  return (int)std::lrintf(x);
#endif
}
// 2^(f - 1023) as a double from a biased exponent field f in [1, 2046]
OZ_HD double pow2_field(int f) {
#ifdef __CUDA_ARCH__
  return __hiloint2double(f << 20, 0);
#else
  const unsigned long long bits = (unsigned long long)(uint32_t)f << 52;
  double d;
  std::memcpy(&d, &bits, 8);
  return d;
#endif
}
// 2^(f - 127) as a float from a biased exponent field f in [1, 254]
OZ_HD float pow2_field_f(int f) {
#ifdef __CUDA_ARCH__
  return __int_as_float(f << 23);
#else
  const uint32_t bits = (uint32_t)f << 23;
  float d;
  std::memcpy(&d, &bits, 4);
  return d;
#endif
}

OZ_HD unsigned long long double_bits(double x) {
#ifdef __CUDA_ARCH__
  return (unsigned long long)__double_as_longlong(x);
#else
  unsigned long long b;
  std::memcpy(&b, &x, 8);
  return b;
#endif
}
OZ_HD double bits_double(unsigned long long b) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)b);
#else
  double x;
  std::memcpy(&x, &b, 8);
  return x;
#endif
}
OZ_HD double fma_rn(double a, double b, double c) {
#ifdef __CUDA_ARCH__
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}
// 64-bit float <-> integer conversion instructions (F2I.S64.F64, I2F.F64.S64) run at a small
// fraction of the FP64 rate on sm_100 and block the FP64 pipe while they do: measured, they
// were the bound of k_ozaki_t (stall_math on every DMUL behind them).  The kernels therefore
// convert with the magic-number identity instead: for an integer |q| < 2^51,
//     bits(1.5 * 2^52 + q) = bits(1.5 * 2^52) + q     (two's complement in the mantissa),
// and the addition is exactly round-to-nearest-even to an integer -- one DFMA / DADD on the
// full-rate pipe plus 64-bit integer adds.
constexpr unsigned long long MAGIC52_BITS = 0x4338000000000000ull;   // 1.5 * 2^52
OZ_HD double magic52() { return 6755399441055744.0; }

// 4 x 4 byte transpose of w[0..3]: o_i = (byte i of w0, byte i of w1, byte i of w2, byte i of w3)
OZ_HD void transpose4(const uint32_t* w, uint32_t& o0, uint32_t& o1, uint32_t& o2, uint32_t& o3) {
  const uint32_t t0 = byte_perm(w[0], w[1], 0x5140), t1 = byte_perm(w[0], w[1], 0x7362);
  const uint32_t t2 = byte_perm(w[2], w[3], 0x5140), t3 = byte_perm(w[2], w[3], 0x7362);
  o0 = byte_perm(t0, t2, 0x5410);
  o1 = byte_perm(t0, t2, 0x7632);
  o2 = byte_perm(t1, t3, 0x5410);
  o3 = byte_perm(t1, t3, 0x7632);
}

template <class Real>
struct Traits;

template <>
struct Traits<double> {
  static constexpr int S = 6;                // int8 digits per real number
  static constexpr int QBITS = 8 * S - 2;    // |q| <= 2^46
  static constexpr unsigned long long BIAS = 0x808080808080ull;   // sum_{i<S} 128 * 256^i
  static constexpr int MIN_EF = 64;          // rows below 2^-958 flush to zero
  // key: monotonic in |x|; exp_field(key) = biased exponent field, |x| < 2^(field - 1022)
  static OZ_HD int key(double x) {
#ifdef __CUDA_ARCH__
    return __double2hiint(x) & 0x7fffffff;
#else
    unsigned long long bits;
    std::memcpy(&bits, &x, 8);
    return (int)((bits >> 32) & 0x7fffffffu);
#endif
  }
  static OZ_HD int exp_field(int key) { return key >> 20; }
  // slicing scale 2^(QBITS - E), E = ef - 1022
  static OZ_HD double slice_scale(int ef) { return ef >= MIN_EF ? pow2_field(QBITS + 2045 - ef) : 0.0; }
  // output scale 2^(E - 6):  C = 2^(EA + EB - 2 QBITS) 256^(2 (S - 1)) sum_g acc_g 256^-g
  //                            = 2^(EA - 6) 2^(EB - 6) sum_g acc_g 256^-g
  // (a row holding an Inf or a NaN has the all-ones field: its results are NaN, as with DMMA)
  static OZ_HD double out_scale(int ef) {
    return ef >= 2047 ? pow2_field(2047) * 0.0 : ef >= MIN_EF ? pow2_field(ef - 5) : 0.0;
  }

  // exponent of out_scale: out_scale(ef) = 2^(ef - 1028) for MIN_EF <= ef < 2047
  static OZ_HD int out_exp(int ef) { return ef - 1028; }

  // Slices 16 reals of one row (one 16-byte K chunk) into S planes: out[s] holds digit s
  // (s = 0 most significant) of the 16 numbers, byte j = number j.
  static OZ_HD void slice16(const double* x, double scale, bool negate, Word4* out) {
#pragma unroll
    for (int jg = 0; jg < 4; ++jg) {
      uint32_t lo[4], hi[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // q = rint(x * scale) without a conversion instruction (see magic52 above); |q| <= 2^46.
        // A negated plane is the product with -scale (exact: scale is a power of two).
        const double y = fma_rn(x[4 * jg + j], negate ? -scale : scale, magic52());
        const long long q = (long long)(double_bits(y) - MAGIC52_BITS);
        const unsigned long long u = (unsigned long long)(q + (long long)BIAS) ^ BIAS;
        lo[j] = (uint32_t)u;           // digits 256^0 .. 256^3
        hi[j] = (uint32_t)(u >> 32);   // digits 256^4, 256^5 (upper half zero)
      }
      uint32_t unused0, unused1;
      transpose4(lo, out[5].w[jg], out[4].w[jg], out[3].w[jg], out[2].w[jg]);   // 256^0 = plane S-1
      transpose4(hi, out[1].w[jg], out[0].w[jg], unused0, unused1);
    }
  }
};

template <>
struct Traits<float> {
  static constexpr int S = 4;                // (three digits would quantise the row to 22 bits,
  static constexpr int QBITS = 8 * S - 2;    //  coarser than the float itself) |q| <= 2^30
  static constexpr unsigned BIAS = 0x80808080u;
  static constexpr int MIN_EF = 32;          // rows below 2^-94 flush to zero
  static OZ_HD int key(float x) {
#ifdef __CUDA_ARCH__
    return __float_as_int(x) & 0x7fffffff;
#else
    uint32_t bits;
    std::memcpy(&bits, &x, 4);
    return (int)(bits & 0x7fffffffu);
#endif
  }
  static OZ_HD int exp_field(int key) { return key >> 23; }   // |x| < 2^(field - 126)
  // slicing scale 2^(QBITS - E), E = ef - 126: field = QBITS + 126 - ef + 127 in [29, 251]
  static OZ_HD float slice_scale(int ef) { return ef >= MIN_EF ? pow2_field_f(QBITS + 253 - ef) : 0.0f; }
  // output scale 2^(E - 6) (the same formula as for double: 2 QBITS - 16 (S - 1) = 12), kept in
  // double so that the product of a row and a column scale cannot under- or overflow
  static OZ_HD double out_scale(int ef) {
    return ef >= 255 ? pow2_field(2047) * 0.0 : ef >= MIN_EF ? pow2_field(ef + 891) : 0.0;
  }

  // the same scale as a float, 2^(ef - 132) in [2^-100, 2^122] (k_ozaki_t scales its float
  // results with two FMULs, row scale first); `extra` = -8 (G - 1) for the column scale
  static OZ_HD float out_scale_f(int ef, int extra) {
    return ef >= 255 ? pow2_field_f(255) * 0.0f : ef >= MIN_EF ? pow2_field_f(ef - 5 + extra) : 0.0f;
  }

  static OZ_HD void slice16(const float* x, float scale, bool negate, Word4* out) {
#pragma unroll
    for (int jg = 0; jg < 4; ++jg) {
      uint32_t u[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int q = f2i_rn(x[4 * jg + j] * scale);
        if (negate) q = -q;
        u[j] = ((uint32_t)q + BIAS) ^ BIAS;   // digits 256^0 .. 256^3 (no overflow: |q| <= 2^30)
      }
      transpose4(u, out[3].w[jg], out[2].w[jg], out[1].w[jg], out[0].w[jg]);
    }
  }
};

// byte offset of (row, 16-k chunk) inside a no-swizzle K-major plane with `rows` rows:
// core matrix = 8 rows x 16 bytes, 128 bytes between 8-row groups, rows * 16 bytes between
// K chunks (the UMMA descriptor's SBO and LBO)
OZ_HD uint32_t plane_off(int rows, int row, int chunk) {
  return (uint32_t)(chunk * rows * 16 + (row >> 3) * 128 + (row & 7) * 16);
}

// ---------------------------------------------------------------------------
// The kernel's operand layout (kernels_zgemm_ozaki2.cu: k_ozaki_t), "transposed, K-concatenated":
// the SMALL operand B is the resident M-side operand of the MMA, a tile of 64 rows of A is the
// N-side operand, and the complex product is folded into ONE real contraction of length 2K:
//     Cr[n, m] = sum_k  Br[k, n] Ar[m, k] + (-Bi[k, n]) Ai[m, k]
//     Ci[n, m] = sum_k  Bi[k, n] Ar[m, k] +   Br[k, n]  Ai[m, k]
// W (resident, 128 rows): row w_row(n, 0) = [Br(., n) | -Bi(., n)], row w_row(n, 1) = [Bi(., n) |
// Br(., n)];  X (per tile, 64 rows): row j = [Ar(m_j, .) | Ai(m_j, .)].
// D = W X^T puts the rows of the tile into TMEM columns and Cr / Ci of column n into the TMEM
// lanes w_row(n, 0 / 1) = 16 (n / 8) + 8 part + n % 8: EIGHT LANES APART inside a block of 16,
// which is the accumulator-fragment layout of tcgen05.ld.16x256b (thread t of a warp receives
// lane t / 4 and lane t / 4 + 8, columns 2 (t % 4) and 2 (t % 4) + 1) -- one load hands every
// thread re AND im of two consecutive rows of C: whole complex numbers, no exchange.  A plane row holds KCH = 8 chunks of 16 int8: chunks [0, KC) are the
// "re half", [KC, 2 KC) the "im half" (KC = ceil(K / 16) <= 4), so one MMA k-step (32 bytes)
// is chunks (2 ks, 2 ks + 1) and there are KC k-steps.
// ---------------------------------------------------------------------------
namespace ot {

constexpr int ROWS = 64;     // rows of A per tile = UMMA N
constexpr int WROWS = 128;   // rows of W = UMMA M
constexpr int KCH = 8;       // chunk slots per plane row
constexpr int W_PLANE = WROWS * KCH * 16;   // bytes
constexpr int X_PLANE = ROWS * KCH * 16;
constexpr int W_LBO = WROWS * 16, X_LBO = ROWS * 16, SBO = 128;

OZ_HD int w_row(int n, int part) { return 16 * (n >> 3) + 8 * part + (n & 7); }

OZ_HD void store16(unsigned char* dst, const Word4& v) {
#ifdef __CUDA_ARCH__
  *reinterpret_cast<uint4*>(dst) = make_uint4(v.w[0], v.w[1], v.w[2], v.w[3]);
#else
  std::memcpy(dst, v.w, 16);
#endif
}

// digit planes of one (tile row j, chunk c) work item of X: 16 complex numbers of A's row
template <class Real, class Scale>
OZ_HD void x_item(unsigned char* X, int j, int c, int KC, const Real* xr, const Real* xi, Scale scale) {
  using Tr = Traits<Real>;
  Word4 pl[Tr::S];
  Tr::slice16(xr, scale, false, pl);
#pragma unroll
  for (int s = 0; s < Tr::S; ++s) store16(X + s * X_PLANE + plane_off(ROWS, j, c), pl[s]);
  Tr::slice16(xi, scale, false, pl);
#pragma unroll
  for (int s = 0; s < Tr::S; ++s) store16(X + s * X_PLANE + plane_off(ROWS, j, KC + c), pl[s]);
}

// digit planes of one (column n of B, chunk c) work item of W: rows w_row(n, 0) (Cr), w_row(n, 1) (Ci)
template <class Real, class Scale>
OZ_HD void w_item(unsigned char* W, int n, int c, int KC, const Real* xr, const Real* xi, Scale scale) {
  using Tr = Traits<Real>;
  Word4 pl[Tr::S];
  Tr::slice16(xr, scale, false, pl);   // Br: re half of the Cr row, im half of the Ci row
#pragma unroll
  for (int s = 0; s < Tr::S; ++s) {
    store16(W + s * W_PLANE + plane_off(WROWS, w_row(n, 0), c), pl[s]);
    store16(W + s * W_PLANE + plane_off(WROWS, w_row(n, 1), KC + c), pl[s]);
  }
  Tr::slice16(xi, scale, false, pl);   // Bi: re half of the Ci row
#pragma unroll
  for (int s = 0; s < Tr::S; ++s) store16(W + s * W_PLANE + plane_off(WROWS, w_row(n, 1), c), pl[s]);
  Tr::slice16(xi, scale, true, pl);    // -Bi: im half of the Cr row
#pragma unroll
  for (int s = 0; s < Tr::S; ++s) store16(W + s * W_PLANE + plane_off(WROWS, w_row(n, 0), KC + c), pl[s]);
}

// MMA schedule of accumulator group g (all digit pairs (t of W, s of X) with s + t = g):
// f(w_plane, x_plane, kstep, accumulate) in issue order
template <int S, class F>
OZ_HD void for_each_mma_of_group(int g, int KC, F&& f) {
  unsigned acc = 0;
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int t = g - s;
    if (t < 0 || t >= S) continue;
    for (int ks = 0; ks < KC; ++ks) {
      f(t, s, ks, acc);
      acc = 1u;
    }
  }
}

// V = sum_g r_g 256^(G-1-g) as an exact integer.  With |digit| <= 128 and a contraction of at
// most 128 bytes, |r_g| <= (g + 1) 2^21: the pair sums r0 256 + r1 and r2 256 + r3 fit int32.
template <int G>
OZ_HD long long combine(const int* r);
template <>
OZ_HD long long combine<6>(const int* r) {
  const int p01 = r[0] * 256 + r[1], p23 = r[2] * 256 + r[3];
  return (long long)p01 * 4294967296ll + (long long)p23 * 65536ll + (long long)r[4] * 256ll + (long long)r[5];
}
template <>
OZ_HD long long combine<4>(const int* r) {
  const int p01 = r[0] * 256 + r[1], p23 = r[2] * 256 + r[3];
  return (long long)p01 * 65536ll + (long long)p23;
}
// (double)(V 2^c), correctly rounded, without I2F.F64.S64 and with the power-of-two column
// scale 2^c for free: V = hi 2^32 + lo (hi = V >> 32 signed, lo the unsigned low word),
//     bits((0x453 + c) << 52 | (hi ^ 2^31)) = 2^c (2^84 + (hi + 2^31) 2^32)
//     bits((0x433 + c) << 52 | lo)          = 2^c (2^52 + lo)
// With C = 2^84 + 2^63 + 2^52 (exactly representable), dhi - 2^c C = 2^c (hi 2^32 - 2^52) is
// exact (32 significant bits) and adding dlo rounds once: two DADDs.  (FP64 instructions are
// scarce beside a running MMA stream -- one per ~9 clocks per warp and scheduler, measured by
// pq_microbench "ozaki_t_rate_32" -- but integer sequences cost issue slots, which the
// epilogue has even fewer of: the first integer-only version of this epilogue ran ~80
// instructions per real number and was slower than three FP64 operations.)
// Both scales ride in the exponent fields: e = c (column) + r (row).  The high words of the two
// magic doubles are K + (e << 20) and the offset 2^e C has the SAME high word as dhi's magic
// (C = bits 0x45300000 80100000), so a thread keeps two 32-bit constants per column and adds
// the row's pre-shifted exponent word: per real number 2 DADDs and ~3 integer instructions, no
// multiplication at all.  Row words: r << 20 for a normal row, ROW_NAN for a row holding Inf /
// NaN (the conversion then yields NaN); flushed rows need nothing (all their digits are zero).
// The fields 0x433 + e .. 0x453 + e must stay normal, so field_clamp limits the magnitudes: a
// row (column) whose largest magnitude is below 2^-494 (~1e-149) is sliced as if it were 2^-494
// and loses the digits below 2^-540 -- nobody can see them in a tensor-network amplitude --,
// and one at or above 2^+486 (~1e146) counts as Inf.
struct DoubleMagic {   // per-thread constants for one column exponent c
  int k_hi, k_lo;      // (0x453 + c) << 20, (0x433 + c) << 20
  bool nan_column;
};
constexpr int ROW_NAN = 0x7fffffff;
// exponent fields (|x| < 2^(ef - 1022)) a row or column may have without clamping: with
// r = efa - 1028 in [-500, 480] and c = efb - 1028 - 8 (G - 1) in [-540, 440] the combined
// exponent stays in [-1040, 920] and every field normal
constexpr int FIELD_LO = 528, FIELD_HI = 1508;
OZ_HD int field_clamp(int ef) {   // fields below MIN_EF (flushed rows) pass through
  return ef > FIELD_HI ? 2047 : (ef >= 64 && ef < FIELD_LO) ? FIELD_LO : ef;
}
OZ_HD int row_word(int ef) { return ef >= 2047 ? ROW_NAN : ef >= 64 ? (ef - 1028) * (1 << 20) : 0; }
OZ_HD DoubleMagic double_magic(int c, bool nan_column) {
  DoubleMagic m;
  m.k_hi = (0x453 + c) << 20;
  m.k_lo = (0x433 + c) << 20;
  m.nan_column = nan_column;
  return m;
}
// (double)(V 2^(c + r)), correctly rounded; rw = the row word
OZ_HD double to_double_scaled(long long v, const DoubleMagic& m, int rw) {
  const unsigned lo = (unsigned)(unsigned long long)v;
  const int hi = (int)(v >> 32);
  const bool nan = rw == ROW_NAN || m.nan_column;
  const unsigned w1 = nan ? 0x7ff80000u : (unsigned)(m.k_hi + rw), w2 = (unsigned)(m.k_lo + rw);
  const double dhi = bits_double(((unsigned long long)w1 << 32) | (unsigned)(hi ^ 0x80000000));
  const double dlo = bits_double(((unsigned long long)w2 << 32) | lo);
  const double off = bits_double(((unsigned long long)w1 << 32) | 0x80100000u);
  return (dhi - off) + dlo;
}
// ONE DADD per real number (ComplexF64, 6 groups): V has up to 62 bits, the magic-number
// conversion takes 51.  The low 11 bits of V weigh about as much as the groups the scheme
// drops anyway (sum_{g >= 6} r_g 256^(5-g): ~2^9 for random digits), so V is first rounded to
// V' = round(V / 2^11), |V'| < 2^51 -- from the group sums, int32 up to two widening
// multiply-adds -- and V' 2^(11 + e) is one conversion:
//     bits(1.5 2^52 2^(11 + e)) + V'  =  bits of  2^(11 + e) (1.5 2^52 + V').
// FP64 instructions are the scarcest resource of this kernel (they share hardware with the
// running MMA stream: measured, the epilogue's drain time halves when its DADDs are removed).
constexpr int V_DROP = 11;
OZ_HD long long combine51(const int* r) {   // V' = round(sum_g r_g 256^(5-g) / 2^11), G = 6
  const int p01 = r[0] * 256 + r[1], p23 = r[2] * 256 + r[3];
  const int s45 = r[4] * 32 + ((r[5] + 4) >> 3);    // (r4 256 + r5) / 8, |r4| < 2^24
  const int t45 = (s45 + 128) >> 8;                 // ... / 2^11, rounded
  return (long long)p01 * 2097152ll + ((long long)p23 * 32ll + (long long)t45);
}
OZ_HD double to_double51_scaled(long long v51, const DoubleMagic& m, int rw) {
  const bool nan = rw == ROW_NAN || m.nan_column;
  // magic = 1.5 * 2^52 * 2^(11 + e): high word (0x433 + 11 + e) << 20 | 0x80000, low word 0
  const unsigned wm = (unsigned)(m.k_lo + rw) + ((unsigned)V_DROP << 20) + 0x80000u;
  const double magic = bits_double((unsigned long long)wm << 32);
  const double y = bits_double(((unsigned long long)wm << 32) + (unsigned long long)v51);
  const double d = y - magic;
  return nan ? bits_double(0x7ff8000000000000ull) : d;
}
// ComplexF32 result: V = p01 2^16 + p23 rounded to float through two int32 -> float
// conversions and one FMA (within one ulp of V; no 64-bit conversion)
OZ_HD float combine_f32(const int* r) {
  const int p01 = r[0] * 256 + r[1], p23 = r[2] * 256 + r[3];
#ifdef __CUDA_ARCH__
  return __fmaf_rn(__int2float_rn(p01), 65536.0f, __int2float_rn(p23));
#else
  return std::fmaf((float)p01, 65536.0f, (float)p23);
#endif
}

// 256^-(G-1): C = out_scale(row) * out_scale(column) * 256^-(G-1) * V
template <int G>
OZ_HD double group_weight() {
  return 1.0 / (double)(1ull << (8 * (G - 1)));
}

}  // namespace ot

}  // namespace oz
}  // namespace pq
