// Integer arithmetic of the INT8 Ozaki-scheme ZGEMM (kernels_zgemm_ozaki.cu), shared between
// the CUDA kernel and the host-side emulation in test_lower.cpp, so that the slicing, the
// packing into the UMMA core-matrix layout and the recombination are checked without a GPU
// (the same role tile_math.h plays for the permute kernel).
//
// A real x of a row whose largest magnitude has biased exponent field `ef` (|x| < 2^(ef-1022))
// becomes q = rint(x * 2^(QBITS - (ef - 1022))), |q| <= 2^QBITS, written in balanced base 256:
// q = sum_i d_i 256^i with d_i in [-128, 127] (int8).  The digits are the BYTES of q + BIAS
// (BIAS = sum_i 128 * 256^i) with their top bit flipped, so there are no carries and no
// bit-field shuffling.  Plane s holds digit 256^(S-1-s).
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

constexpr int S = 6;                       // int8 digits per real number
constexpr int QBITS = 8 * S - 2;           // |q| <= 2^46 (one bit of headroom below the bias)
constexpr unsigned long long BIAS = 0x808080808080ull;   // sum_{i<S} 128 * 256^i
constexpr int MIN_EF = 64;                 // rows below 2^-958 flush to zero
constexpr int HI_GROUPS = 3;               // accumulator groups summed in the first Horner sum
static_assert(S == 6, "BIAS, slice16 and the transposes are written for six digits");

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
// slicing scale 2^(QBITS - (ef - 1022)) and output scale 2^((ef - 1022) - 6) of a row / column:
// C = 2^(EA + EB - 2 QBITS) * 256^(2 (S - 1)) * sum_g acc_g 256^-g = 2^(EA - 6) 2^(EB - 6) * sum_g ...
OZ_HD double slice_scale(int ef) { return ef >= MIN_EF ? pow2_field(QBITS + 2045 - ef) : 0.0; }
OZ_HD double out_scale(int ef) { return ef >= MIN_EF ? pow2_field(ef - 5) : 0.0; }

// Slices 16 reals of one row (one 16-byte K chunk) into S planes: out[s] holds digit s
// (s = 0 most significant) of the 16 numbers, byte j = number j.
OZ_HD void slice16(const double* x, double scale, bool negate, Word4* out) {
#pragma unroll
  for (int jg = 0; jg < 4; ++jg) {
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      long long q = d2ll_rn(x[4 * jg + j] * scale);
      if (negate) q = -q;
      const unsigned long long u = (unsigned long long)(q + (long long)BIAS) ^ BIAS;
      lo[j] = (uint32_t)u;           // digits 256^0 .. 256^3
      hi[j] = (uint32_t)(u >> 32);   // digits 256^4, 256^5 (upper half zero)
    }
    // 4 x 4 byte transposes: the word of plane i = byte i of the four numbers
    const uint32_t t0 = byte_perm(lo[0], lo[1], 0x5140), t1 = byte_perm(lo[0], lo[1], 0x7362);
    const uint32_t t2 = byte_perm(lo[2], lo[3], 0x5140), t3 = byte_perm(lo[2], lo[3], 0x7362);
    out[5].w[jg] = byte_perm(t0, t2, 0x5410);   // 256^0 = least significant = plane S - 1
    out[4].w[jg] = byte_perm(t0, t2, 0x7632);
    out[3].w[jg] = byte_perm(t1, t3, 0x5410);
    out[2].w[jg] = byte_perm(t1, t3, 0x7632);
    const uint32_t v0 = byte_perm(hi[0], hi[1], 0x5140), v2 = byte_perm(hi[2], hi[3], 0x5140);
    out[1].w[jg] = byte_perm(v0, v2, 0x5410);
    out[0].w[jg] = byte_perm(v0, v2, 0x7632);
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

// value of sum_g acc_g 256^-g from the two Horner sums hi = sum_{g < HI_GROUPS} acc_g
// 256^(HI_GROUPS-1-g) and lo = sum_{g >= HI_GROUPS} acc_g 256^(G-1-g)   (G > HI_GROUPS)
OZ_HD double combine(long long hi, long long lo, int G) {
  const double whi = 1.0 / (double)(1ull << (8 * (HI_GROUPS - 1)));
  const double wlo = 1.0 / (double)(1ull << (8 * (G - 1)));
#ifdef __CUDA_ARCH__
  return fma((double)lo, wlo, (double)hi * whi);
#else
  return std::fma((double)lo, wlo, (double)hi * whi);
#endif
}

}  // namespace oz
}  // namespace pq
