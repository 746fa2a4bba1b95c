// Integer arithmetic of the INT8 Ozaki-scheme ZGEMM (kernels_zgemm_ozaki.cu), shared between
// the CUDA kernel and the host-side emulation in test_lower.cpp, so that the slicing, the
// packing into the UMMA core-matrix layout and the recombination are checked without a GPU
// (the same role tile_math.h plays for the permute kernel).
//
// A real x of a row whose largest magnitude has biased exponent field `ef` (|x| < 2^(ef-1022))
// becomes q = rint(x * 2^(QBITS - (ef - 1022))), |q| <= 2^QBITS, written in balanced base 128:
// q = sum_i d_i 128^i with d_i in [-64, 63].  The digits are the 7-bit fields of q + BIAS minus
// 64 (BIAS = sum_i 64 * 128^i), so no carries propagate.  Plane s holds digit 128^(S-1-s).
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

constexpr int S = 7;                       // int8 digits per real number
constexpr int QBITS = 7 * S - 2;           // |q| <= 2^47
constexpr unsigned long long BIAS = 64ull * ((1ull << (7 * S)) - 1ull) / 127ull;
constexpr int MIN_EF = 64;                 // rows below 2^-958 flush to zero

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
OZ_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, int sh) {
#ifdef __CUDA_ARCH__
  return __funnelshift_r(lo, hi, sh);
#else
  return (uint32_t)(((((unsigned long long)hi) << 32) | lo) >> sh);
#endif
}
OZ_HD long long d2ll_rn(double x) {
#ifdef __CUDA_ARCH__
  return __double2ll_rn(x);
#else
  return std::llrint(x);   // round-to-nearest-even in the default rounding mode
#endif
}
// 2^(f - 1023) from a biased exponent field f in [1, 2046]
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
// high word of |x|: monotonic in |x|, >> 20 gives the biased exponent field
OZ_HD int abs_hi(double x) {
#ifdef __CUDA_ARCH__
  return __double2hiint(x) & 0x7fffffff;
#else
  unsigned long long bits;
  std::memcpy(&bits, &x, 8);
  return (int)((bits >> 32) & 0x7fffffffu);
#endif
}
// slicing scale 2^(QBITS - (ef - 1022)) and output scale 2^((ef - 1022) - 5) of a row / column
OZ_HD double slice_scale(int ef) { return ef >= MIN_EF ? pow2_field(QBITS + 2045 - ef) : 0.0; }
OZ_HD double out_scale(int ef) { return ef >= MIN_EF ? pow2_field(ef - 4) : 0.0; }

// 7-bit fields -> bytes: bits [7i, 7i+7) of t go to byte i (i < 4)
OZ_HD uint32_t spread(uint32_t t) {
  return (t & 0x7Fu) | ((t & 0x3F80u) << 1) | ((t & 0x1FC000u) << 2) | ((t & 0xFE00000u) << 3);
}
// field e in [0, 127] -> int8 digit e - 64, on four packed bytes
OZ_HD uint32_t unbias(uint32_t p) {
  p ^= 0x40404040u;
  return p | ((p & 0x40404040u) << 1);
}

// Slices 16 reals of one row (one 16-byte K chunk) into S planes: out[s] holds digit s
// (s = 0 most significant) of the 16 numbers, byte j = number j.
OZ_HD void slice16(const double* x, double scale, bool negate, Word4* out) {
#pragma unroll
  for (int jg = 0; jg < 4; ++jg) {
    uint32_t p0[4], p1[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      long long q = d2ll_rn(x[4 * jg + j] * scale);
      if (negate) q = -q;
      const unsigned long long u = (unsigned long long)(q + (long long)BIAS);
      const uint32_t lo = (uint32_t)u, hi = (uint32_t)(u >> 32);
      p0[j] = unbias(spread(lo & 0x0FFFFFFFu));                   // digits 128^0 .. 128^3
      p1[j] = unbias(spread(funnel_r(lo, hi, 28) & 0x1FFFFFu));   // digits 128^4 .. 128^6
    }
    // 4 x 4 byte transposes: the word of plane i = byte i of the four numbers
    const uint32_t t0 = byte_perm(p0[0], p0[1], 0x5140), t1 = byte_perm(p0[0], p0[1], 0x7362);
    const uint32_t t2 = byte_perm(p0[2], p0[3], 0x5140), t3 = byte_perm(p0[2], p0[3], 0x7362);
    out[6].w[jg] = byte_perm(t0, t2, 0x5410);   // 128^0 = least significant = plane 6
    out[5].w[jg] = byte_perm(t0, t2, 0x7632);
    out[4].w[jg] = byte_perm(t1, t3, 0x5410);
    out[3].w[jg] = byte_perm(t1, t3, 0x7632);
    const uint32_t v0 = byte_perm(p1[0], p1[1], 0x5140), v1 = byte_perm(p1[0], p1[1], 0x7362);
    const uint32_t v2 = byte_perm(p1[2], p1[3], 0x5140), v3 = byte_perm(p1[2], p1[3], 0x7362);
    out[2].w[jg] = byte_perm(v0, v2, 0x5410);
    out[1].w[jg] = byte_perm(v0, v2, 0x7632);
    out[0].w[jg] = byte_perm(v1, v3, 0x5410);
  }
}

// byte offset of (row, 16-k chunk) inside a no-swizzle K-major plane with `rows` rows:
// core matrix = 8 rows x 16 bytes, 128 bytes between 8-row groups, rows * 16 bytes between
// K chunks (the UMMA descriptor's SBO and LBO)
OZ_HD uint32_t plane_off(int rows, int row, int chunk) {
  return (uint32_t)(chunk * rows * 16 + (row >> 3) * 128 + (row & 7) * 16);
}

// The MMA schedule of accumulator group g of one (tile, 32-column half): calls
//     f(accumulator, a_plane, b_plane, ks, accumulate)
// for every 128 x 32 x 32 MMA in issue order.  Accumulator 2g is Cr of group g, 2g + 1 is Ci.
// A planes: [0, S) re digits, [S, 2S) im digits.  B planes: [0, S) re, [S, 2S) im,
// [2S, 3S) digits of -im (Cr = Ar Br + Ai (-Bi), Ci = Ar Bi + Ai Br).
template <int KS, class F>
OZ_HD void for_each_mma_of_group(int g, F&& f) {
  unsigned acc = 0;
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int t = g - s;
    if (t < 0 || t >= S) continue;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      f(2 * g, s, t, ks, acc);
      f(2 * g + 1, s, S + t, ks, acc);
      f(2 * g, S + s, 2 * S + t, ks, 1u);
      f(2 * g + 1, S + s, t, ks, 1u);
      acc = 1u;
    }
  }
}

// value of sum_g acc_g 2^(-7 g) from the two Horner sums (groups 0..3 and 4..G-1)
OZ_HD double combine(long long hi, long long lo, int G) {
  const double low = 1.0 / (double)(1ull << (7 * (G - 1)));
  if (G <= 4) return (double)hi * low;   // (then `hi` holds groups 0..G-1)
#ifdef __CUDA_ARCH__
  return fma((double)lo, low, (double)hi * (1.0 / 2097152.0));
#else
  return std::fma((double)lo, low, (double)hi * (1.0 / 2097152.0));
#endif
}

}  // namespace oz
}  // namespace pq
