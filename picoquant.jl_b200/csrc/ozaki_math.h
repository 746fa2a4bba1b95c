// Integer arithmetic of the INT8 Ozaki-scheme complex GEMM (kernels_zgemm_ozaki.cu), shared
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

constexpr int HI_GROUPS = 3;   // accumulator groups summed in the first Horner sum

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

  // Slices 16 reals of one row (one 16-byte K chunk) into S planes: out[s] holds digit s
  // (s = 0 most significant) of the 16 numbers, byte j = number j.
  static OZ_HD void slice16(const double* x, double scale, bool negate, Word4* out) {
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

// The MMA schedule of accumulator group g of one (tile, column block): calls
//     f(accumulator, a_plane, b_plane, ks, accumulate)
// for every MMA in issue order.  Accumulator 2g is Cr of group g, 2g + 1 is Ci.
// A planes: [0, S) re digits, [S, 2S) im digits.  B planes: [0, S) re, [S, 2S) im,
// [2S, 3S) digits of -im (Cr = Ar Br + Ai (-Bi), Ci = Ar Bi + Ai Br).
template <int S, int KS, class F>
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
// 256^(HI_GROUPS-1-g) and lo = sum_{g >= HI_GROUPS} acc_g 256^(G-1-g)  (lo = 0 if G = HI_GROUPS)
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
