// CPU harness for the integer half of the CUDA path: emulates the tiled
// bit-permutation kernel and the gather maps of the contraction kernels on the host,
// using the very same index functions the kernels call (tile_math.h, map_offset), and
// compares them with brute-force multi-index arithmetic.  Run by tests/test_lowering.py.
#include <algorithm>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <numeric>
#include <random>

#include "common.h"
#include "ozaki_math.h"
#include "tile_math.h"

using namespace pq;
typedef std::complex<double> cd;

static int failures = 0;
#define CHECK(cond, ...)            \
  do {                              \
    if (!(cond)) {                  \
      ++failures;                   \
      std::printf("FAIL: " __VA_ARGS__); \
      std::printf("\n");            \
    }                               \
  } while (0)

// brute force: out[o] = in[i] where the multi-index of o over out dims maps to perm
static std::vector<int64_t> ref_permute_sources(const std::vector<int64_t>& dims,
                                                const std::vector<int>& perm) {
  int r = (int)dims.size();
  int64_t total = prod(dims);
  std::vector<int64_t> istr(r);
  int64_t acc = 1;
  for (int i = 0; i < r; ++i) {
    istr[i] = acc;
    acc *= dims[i];
  }
  std::vector<int64_t> src(total);
  for (int64_t o = 0; o < total; ++o) {
    int64_t rem = o, off = 0;
    for (int k = 0; k < r; ++k) {
      int64_t e = dims[perm[k]];
      off += (rem % e) * istr[perm[k]];
      rem /= e;
    }
    src[o] = off;
  }
  return src;
}

static void emulate_tiled(const TileParams& tp, std::vector<int64_t>& out_src, int elem_size) {
  // out_src[o] = input offset that lands at output offset o
  const int T = 1 << tp.t;
  std::vector<int64_t> tile(T);
  const int c = elem_size == 16 ? 3 : 4;
  for (long long r = 0; r < tp.ntiles; ++r) {
    long long ib, ob;
    tile_bases(tp, r, ib, ob);
    std::vector<char> used(T, 0);
    for (int e = 0; e < T; ++e) {
      int s = tile_swizzle(tp, e);
      CHECK(s >= 0 && s < T && !used[s], "swizzle not a bijection (e=%d s=%d)", e, s);
      used[s] = 1;
      tile[s] = ib + ((long long)(e & 31) | tile_in_hi(tp, e >> TILE_LO));
    }
    for (int o = 0; o < T; ++o) {
      int e = tile_e_lo(tp, o & 31) | tile_e_hi(tp, o >> TILE_LO);
      long long dst = ob + ((long long)(o & 31) | tile_out_hi(tp, o >> TILE_LO));
      out_src[dst] = tile[tile_swizzle(tp, e)];
    }
    // bank-conflict freedom of both phases for the first tile
    if (r == 0) {
      const int group = 1 << c;
      for (int base = 0; base < T; base += group) {
        unsigned seen_w = 0, seen_r = 0;
        for (int q = 0; q < group; ++q) {
          int sw = tile_swizzle(tp, base + q) & (group - 1);
          int o = base + q;
          int e = tile_e_lo(tp, o & 31) | tile_e_hi(tp, o >> TILE_LO);
          int sr = tile_swizzle(tp, e) & (group - 1);
          seen_w |= 1u << sw;
          seen_r |= 1u << sr;
        }
        CHECK(seen_w == (1u << group) - 1, "smem write conflict at e-base %d", base);
        CHECK(seen_r == (1u << group) - 1, "smem read conflict at o-base %d", base);
      }
    }
  }
}

static void test_permutations(std::mt19937& rng) {
  Options opt;
  int tiled_cases = 0, generic_cases = 0;
  for (int iter = 0; iter < 300; ++iter) {
    int elem = (iter & 1) ? 16 : 8;
    std::vector<int64_t> dims;
    int64_t total = 1;
    bool pow2 = (iter % 3) != 0;
    int target_bits = 4 + (int)(rng() % 13);  // up to 2^16 elements
    while (true) {
      int64_t e = pow2 ? (int64_t(1) << (rng() % 3)) : (int64_t)(1 + rng() % 5);
      if (total * e > (int64_t(1) << target_bits)) break;
      dims.push_back(e);
      total *= e;
      if (dims.size() >= 20) break;
    }
    if (dims.empty()) dims.push_back(1);
    std::vector<int> perm(dims.size());
    std::iota(perm.begin(), perm.end(), 0);
    if (iter % 7 == 0) {
      // a "gate application" style permutation: move one axis to the end
      int a = (int)(rng() % perm.size());
      std::rotate(perm.begin() + a, perm.begin() + a + 1, perm.end());
    } else if (iter % 11 != 0) {
      std::shuffle(perm.begin(), perm.end(), rng);
    }
    PermutePlan P = lower_permute(dims, perm, elem, opt);
    std::vector<int64_t> ref = ref_permute_sources(dims, perm);
    if (P.identity) {
      for (int64_t o = 0; o < total; ++o) CHECK(ref[o] == o, "identity plan for a non-identity permutation");
      continue;
    }
    // a paired plan moves 16-byte pairs: unit u of the plan covers elements 2u, 2u+1
    const int64_t units = P.paired ? total / 2 : total;
    const int64_t w = P.paired ? 2 : 1;
    CHECK(!P.paired || elem == 8, "pairing is a c64-only trick");
    std::vector<int64_t> uref(units);
    for (int64_t u = 0; u < units; ++u) {
      uref[u] = ref[u * w] / w;
      if (P.paired) CHECK(ref[2 * u] % 2 == 0 && ref[2 * u + 1] == ref[2 * u] + 1, "pair split");
    }
    ref = uref;
    total = units;
    for (int64_t o = 0; o < total; ++o) {
      int64_t got = map_offset(P.gmap, o);
      if (got != ref[o]) {
        CHECK(false, "generic map mismatch at %lld (iter %d)", (long long)o, iter);
        break;
      }
    }
    ++generic_cases;
    if (P.tiled) {
      ++tiled_cases;
      CHECK(P.tp.a >= 5 && P.tp.b >= 5, "tile runs too short");
      std::vector<int64_t> got(total, -1);
      emulate_tiled(P.tp, got, P.paired ? 16 : elem);
      for (int64_t o = 0; o < total; ++o)
        if (got[o] != ref[o]) {
          CHECK(false, "tiled permute mismatch at %lld (iter %d, n=%d t=%d)", (long long)o, iter,
                P.tp.n, P.tp.t);
          break;
        }
    }
  }
  std::printf("permutations: %d generic, %d tiled cases\n", generic_cases, tiled_cases);
  CHECK(tiled_cases > 20, "too few tiled cases exercised");
}

// shaped = 0: random small label sets (small / direct kernels).  shaped = 1..3: label sets that
// reach the other lowerings -- 1: long contraction with a tiny output (CK_DOT, with its
// re-ordered k enumeration), 2: GEMM-shaped (fused and materialised TTGT), 3: narrow-N GEMM.
static void test_contractions(std::mt19937& rng, int shaped, int iters) {
  int kinds[5] = {0, 0, 0, 0, 0};
  for (int iter = 0; iter < iters; ++iter) {
    Options opt;
    if (iter % 5 == 1) opt.fused = 1;
    if (iter % 5 == 2 && shaped == 0) opt.gemm = 3;
    // random label assignment: each label goes to A only, B only, or both
    int nlab = 1 + (int)(rng() % (iter % 3 == 0 ? 12 : 7));
    std::vector<int64_t> ad, bd;
    std::vector<int32_t> ai, bi;
    int nopen = 0, ncon = 0;
    if (shaped != 0) {
      // numbers of extent-2 labels that are A-open, B-open and contracted
      int na2, nb2, nc2;
      if (shaped == 1) { na2 = (int)(rng() % 3); nb2 = (int)(rng() % (3 - na2)); nc2 = 9 + (int)(rng() % 4); }
      else if (shaped == 2) { na2 = 6 + (int)(rng() % 2); nb2 = 6; nc2 = 5 + (int)(rng() % 2); }
      else { na2 = 12; nb2 = 3 + (int)(rng() % 2); nc2 = 5 + (int)(rng() % 2); }
      for (int l = 0; l < na2; ++l) { ad.push_back(2); ai.push_back(-(++nopen)); }
      for (int l = 0; l < nb2; ++l) { bd.push_back(2); bi.push_back(-(++nopen)); }
      for (int l = 0; l < nc2; ++l) { ++ncon; ad.push_back(2); ai.push_back(ncon); bd.push_back(2); bi.push_back(ncon); }
      nlab = 0;
    }
    for (int l = 0; l < nlab; ++l) {
      int64_t e = (iter % 4 == 0) ? (int64_t)(1 + rng() % 4) : (int64_t(1) << (rng() % 3));
      int where = (int)(rng() % 3);
      if (where == 0) { ad.push_back(e); ai.push_back(-(++nopen)); }
      else if (where == 1) { bd.push_back(e); bi.push_back(-(++nopen)); }
      else { ++ncon; ad.push_back(e); ai.push_back(ncon); bd.push_back(e); bi.push_back(ncon); }
    }
    // shuffle axis order of each operand independently
    auto shuffle_axes = [&](std::vector<int64_t>& d, std::vector<int32_t>& ix) {
      std::vector<int> p(d.size());
      std::iota(p.begin(), p.end(), 0);
      std::shuffle(p.begin(), p.end(), rng);
      std::vector<int64_t> d2(d.size());
      std::vector<int32_t> i2(d.size());
      for (size_t k = 0; k < p.size(); ++k) { d2[k] = d[p[k]]; i2[k] = ix[p[k]]; }
      d = d2; ix = i2;
    };
    shuffle_axes(ad, ai);
    shuffle_axes(bd, bi);
    ContractPlan P = lower_contract(ad, ai, bd, bi, 16, opt);
    kinds[P.kind]++;
    int64_t na = prod(ad), nb = prod(bd);
    std::vector<cd> A(na), B(nb);
    for (auto& x : A) x = cd((double)(rng() % 17) - 8, (double)(rng() % 13) - 6);
    for (auto& x : B) x = cd((double)(rng() % 11) - 5, (double)(rng() % 7) - 3);
    // brute force over all labels
    std::vector<int64_t> astr(ad.size()), bstr(bd.size());
    { int64_t s = 1; for (size_t i = 0; i < ad.size(); ++i) { astr[i] = s; s *= ad[i]; } }
    { int64_t s = 1; for (size_t i = 0; i < bd.size(); ++i) { bstr[i] = s; s *= bd[i]; } }
    // C axes: A-open (A order) then B-open (B order)
    std::vector<int> a_open, b_open, a_con, b_con;
    for (size_t i = 0; i < ai.size(); ++i) {
      auto it = std::find(bi.begin(), bi.end(), ai[i]);
      if (it == bi.end()) a_open.push_back((int)i);
      else { a_con.push_back((int)i); b_con.push_back((int)(it - bi.begin())); }
    }
    for (size_t i = 0; i < bi.size(); ++i)
      if (std::find(ai.begin(), ai.end(), bi[i]) == ai.end()) b_open.push_back((int)i);
    int64_t M = 1, N = 1, K = 1;
    for (int i : a_open) M *= ad[i];
    for (int i : b_open) N *= bd[i];
    for (int i : a_con) K *= ad[i];
    CHECK(M == P.M && N == P.N && K == P.K, "M/N/K mismatch");
    CHECK(prod(P.cdims) == M * N, "cdims product mismatch");
    std::vector<cd> Cref(M * N), Cgot(M * N);
    for (int64_t m = 0; m < M; ++m)
      for (int64_t n = 0; n < N; ++n) {
        int64_t oa = 0, ob = 0, r = m;
        for (int i : a_open) { oa += (r % ad[i]) * astr[i]; r /= ad[i]; }
        r = n;
        for (int i : b_open) { ob += (r % bd[i]) * bstr[i]; r /= bd[i]; }
        cd acc = 0;
        for (int64_t k = 0; k < K; ++k) {
          int64_t ka = 0, kb = 0, rk = k;
          for (size_t q = 0; q < a_con.size(); ++q) {
            int64_t e = ad[a_con[q]];
            ka += (rk % e) * astr[a_con[q]];
            kb += (rk % e) * bstr[b_con[q]];
            rk /= e;
          }
          acc += A[oa + ka] * B[ob + kb];
        }
        Cref[m + M * n] = acc;
      }
    // the maps the kernels use
    for (int64_t m = 0; m < M; ++m)
      for (int64_t n = 0; n < N; ++n) {
        cd acc = 0;
        for (int64_t k = 0; k < K; ++k)
          acc += A[map_offset(P.mA, m) + map_offset(P.kA, k)] *
                 B[map_offset(P.nB, n) + map_offset(P.kB, k)];
        Cgot[m + M * n] = acc;
      }
    for (int64_t i = 0; i < M * N; ++i)
      if (Cgot[i] != Cref[i]) {
        CHECK(false, "contraction map mismatch (iter %d)", iter);
        break;
      }
    // canonical TTGT layouts: A' = [M|K], B' = [N|K]
    if (P.kind == CK_GEMM && !P.fused_gemm) {
      std::vector<cd> Ap(M * K), Bp(N * K);
      if (P.permA.identity) Ap = A;
      else for (int64_t i = 0; i < M * K; ++i) Ap[i] = A[map_offset(P.permA.gmap, i)];
      if (P.permB.identity) Bp = B;
      else for (int64_t i = 0; i < N * K; ++i) Bp[i] = B[map_offset(P.permB.gmap, i)];
      for (int64_t m = 0; m < M; ++m)
        for (int64_t n = 0; n < N; ++n) {
          cd acc = 0;
          for (int64_t k = 0; k < K; ++k) acc += Ap[m + M * k] * Bp[n + N * k];
          if (acc != Cref[m + M * n]) {
            CHECK(false, "TTGT layout mismatch (iter %d)", iter);
            m = M; break;
          }
        }
    }
  }
  std::printf("contractions (shape class %d): small_right %d small_left %d direct %d dot %d gemm %d\n",
              shaped, kinds[0], kinds[1], kinds[2], kinds[3], kinds[4]);
  if (shaped == 1) CHECK(kinds[3] > iters / 2, "dot lowering not exercised");
  if (shaped == 2 || shaped == 3) CHECK(kinds[4] > iters / 2, "GEMM lowering not exercised");
}

static void test_big_shapes() {
  // QFT-26 style gate application and an RQC sweep step: only check plan selection
  Options opt;
  std::vector<int64_t> ad(26, 2), bd(4, 2);
  std::vector<int32_t> ai(26), bi = {1, 2, -25, -26};
  int o = 0;
  for (int i = 0; i < 26; ++i) ai[i] = (i == 7) ? 1 : (i == 19) ? 2 : -(++o);
  ContractPlan P = lower_contract(ad, ai, bd, bi, 16, opt);
  CHECK(P.kind == CK_SMALL_RIGHT && P.M == (1 << 24) && P.N == 4 && P.K == 4, "QFT gate plan");
  CHECK(P.mA.nd == 3 && P.kA.nd == 2, "QFT gate fusion: got %d m-groups %d k-groups", P.mA.nd, P.kA.nd);
  std::vector<int64_t> ad2(24, 2), bd2(12, 2);
  std::vector<int32_t> ai2(24), bi2(12);
  o = 0;
  for (int i = 0; i < 24; ++i) ai2[i] = (i % 4 == 1) ? (i / 4 + 1) : -(++o);
  for (int i = 0; i < 12; ++i) bi2[i] = (i < 6) ? (6 - i) : -(++o);
  P = lower_contract(ad2, ai2, bd2, bi2, 16, opt);
  CHECK(P.kind == CK_GEMM && P.M == (1 << 18) && P.N == 64 && P.K == 64, "sweep step plan");
  CHECK(P.fused_gemm && P.tempA_bytes == 0 && P.tempB_bytes == 0, "sweep step should be fused TTGT");
  opt.fused = 1;  // materialised TTGT
  P = lower_contract(ad2, ai2, bd2, bi2, 16, opt);
  CHECK(P.kind == CK_GEMM && !P.fused_gemm, "unfused sweep step plan");
  CHECK(!P.permA.identity && P.permA.tiled, "sweep step A permute should be tiled");
  std::printf("big shapes: ok (sweep A-permute tile t=%d a=%d b=%d)\n", P.permA.tp.t, P.permA.tp.a,
              P.permA.tp.b);
}


// ---------------------------------------------------------------------------
// INT8 Ozaki-scheme complex GEMM (kernels_zgemm_ozaki2.cu: k_ozaki_t): the kernel's own slicing,
// plane layout, MMA schedule, recombination and scaling (ozaki_math.h) executed on the host, with
// the tensor core replaced by an integer GEMM that reads its operands through the UMMA
// no-swizzle K-major addressing (LBO / SBO), against a long double reference.
// W = resident B planes (128 rows,
// rows 2n / 2n+1 = Cr / Ci of column n, contraction [re half | im half]), X = a 64-row tile of
// A; the same host-side execution of the kernel's own arithmetic (ozaki_math.h, namespace ot).
// constant != 0: every entry of A and B is (constant, constant) -- the largest accumulator sums.
template <class Real>
static double ozaki_t_tile_error(std::mt19937& rng, int Mrows, int N, int K, double spread_sigma,
                                 double sparsity, double big = 1.0, double constant = 0.0) {
  using Tr = oz::Traits<Real>;
  namespace ot = oz::ot;
  typedef std::complex<Real> cr;
  constexpr int S = Tr::S, G = Tr::S;
  const int KC = (K + 15) / 16;
  std::normal_distribution<double> g(0.0, 1.0);
  std::uniform_real_distribution<double> u(0.0, 1.0);
  auto draw = [&]() {
    if (constant != 0.0) return (Real)constant;
    if (u(rng) < sparsity) return (Real)0;
    return (Real)(g(rng) * std::exp(spread_sigma * g(rng)));
  };
  const Real tiny = sizeof(Real) == 8 ? (Real)1e-300 : (Real)1e-30;
  std::vector<cr> A((size_t)Mrows * K), B((size_t)N * K);   // A[m + Mrows k], B[n + N k]
  for (auto& x : A) x = cr(draw(), draw());
  for (auto& x : B) x = cr(draw(), draw());
  if (big != 1.0) {
    for (int r = 0; r < Mrows; ++r)
      for (int k = 0; k < K; ++k) A[r + (size_t)Mrows * k] *= (Real)((r & 1) ? big : 1.0 / big);
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) B[n + (size_t)N * k] *= (Real)((n & 1) ? 1.0 / big : big);
  }
  if (Mrows > 3 && constant == 0.0)
    for (int k = 0; k < K; ++k) A[3 + (size_t)Mrows * k] = (Real)0;
  if (Mrows > 5 && constant == 0.0)
    for (int k = 0; k < K; ++k) A[5 + (size_t)Mrows * k] *= tiny;

  std::vector<unsigned char> sW((size_t)S * ot::W_PLANE, 0xAB), sX((size_t)S * ot::X_PLANE, 0xCD);
  std::vector<int> rowE(ot::ROWS, 0), colE(64, 0);
  // the kernel's work items: (row or column, 16-k chunk), exponent = max over the chunks
  auto gather = [&](const cr* src, size_t ld, bool valid, int c, Real* xr, Real* xi) {
    int key = 0;
    for (int i = 0; i < 16; ++i) {
      const int k = c * 16 + i;
      const cr v = (valid && k < K) ? src[(size_t)k * ld] : cr(0, 0);
      xr[i] = v.real();
      xi[i] = v.imag();
      key = std::max(key, std::max(Tr::key(v.real()), Tr::key(v.imag())));
    }
    return key;
  };
  Real xr[16], xi[16];
  for (int n = 0; n < 64; ++n) {
    int key = 0;
    for (int c = 0; c < 4; ++c) key = std::max(key, gather(&B[n < N ? n : 0], N, n < N && c < KC, c, xr, xi));
    colE[n] = Tr::exp_field(key);
    if (sizeof(Real) == 8) colE[n] = ot::field_clamp(colE[n]);   // OtScale<double>::col_field
    for (int c = 0; c < KC; ++c) {
      gather(&B[n < N ? n : 0], N, n < N, c, xr, xi);
      ot::w_item<Real>(sW.data(), n, c, KC, xr, xi, Tr::slice_scale(colE[n]));
    }
  }
  for (int j = 0; j < ot::ROWS; ++j) {
    int key = 0;
    for (int c = 0; c < 4; ++c)
      key = std::max(key, gather(&A[j < Mrows ? j : 0], Mrows, j < Mrows && c < KC, c, xr, xi));
    rowE[j] = Tr::exp_field(key);
    if (sizeof(Real) == 8) rowE[j] = ot::field_clamp(rowE[j]);   // OtScale<double>::row_field
    for (int c = 0; c < KC; ++c) {
      gather(&A[j < Mrows ? j : 0], Mrows, j < Mrows, c, xr, xi);
      ot::x_item<Real>(sX.data(), j, c, KC, xr, xi, Tr::slice_scale(rowE[j]));
    }
  }
  auto elem = [&](const std::vector<unsigned char>& buf, size_t base, int lbo, int row, int k) {
    return (int)(int8_t)buf[base + (size_t)(k / 16) * lbo + (size_t)(row / 8) * ot::SBO + (row % 8) * 16 + k % 16];
  };
  // the digits of X reconstruct q exactly, re half and im half
  for (int j = 0; j < std::min(Mrows, 8); ++j)
    for (int k = 0; k < K; ++k) {
      long long qr = 0, qi = 0;
      for (int s = 0; s < S; ++s) {
        qr = qr * 256 + elem(sX, (size_t)s * ot::X_PLANE, ot::X_LBO, j, k);
        qi = qi * 256 + elem(sX, (size_t)s * ot::X_PLANE, ot::X_LBO, j, KC * 16 + k);
      }
      const double sc = (double)Tr::slice_scale(rowE[j]);
      CHECK(qr == std::llrint((double)A[j + (size_t)Mrows * k].real() * sc) &&
                qi == std::llrint((double)A[j + (size_t)Mrows * k].imag() * sc),
            "ozaki_t digits of A[%d,%d]", j, k);
    }
  // MMA schedule: D[group][W row][tile row], int32 with an overflow watch
  std::vector<int32_t> acc((size_t)G * ot::WROWS * ot::ROWS, 0);
  long long worst = 0;
  int mmas = 0;
  for (int grp = 0; grp < G; ++grp)
    ot::for_each_mma_of_group<S>(grp, KC, [&](int wp, int xp, int ks, unsigned accumulate) {
      ++mmas;
      const size_t wb = (size_t)wp * ot::W_PLANE + (size_t)ks * 2 * ot::W_LBO;
      const size_t xb = (size_t)xp * ot::X_PLANE + (size_t)ks * 2 * ot::X_LBO;
      for (int l = 0; l < ot::WROWS; ++l)
        for (int j = 0; j < ot::ROWS; ++j) {
          long long sum = 0;
          for (int k = 0; k < 32; ++k) sum += elem(sW, wb, ot::W_LBO, l, k) * elem(sX, xb, ot::X_LBO, j, k);
          int32_t& a = acc[((size_t)grp * ot::WROWS + l) * ot::ROWS + j];
          const long long v = (accumulate ? (long long)a : 0) + sum;
          worst = std::max(worst, std::llabs(v));
          a = (int32_t)v;
        }
    });
  CHECK(mmas == (S * (S + 1) / 2) * KC, "ozaki_t MMA count %d", mmas);
  CHECK(worst < (1ll << 31), "ozaki_t int32 accumulator overflow: %lld", worst);
  std::vector<std::complex<double>> C((size_t)Mrows * N);
  for (int n = 0; n < N; ++n)
    for (int j = 0; j < Mrows; ++j) {
      double out[2];
      for (int part = 0; part < 2; ++part) {
        int r[G];
        long long exact = 0;
        for (int grp = 0; grp < G; ++grp) {
          r[grp] = acc[((size_t)grp * ot::WROWS + ot::w_row(n, part)) * ot::ROWS + j];
          exact = exact * 256 + r[grp];
        }
        // the int32 pair sums inside combine() must not wrap
        CHECK(std::llabs((long long)r[0] * 256 + r[1]) < (1ll << 31) &&
                  std::llabs((long long)r[2] * 256 + r[3]) < (1ll << 31),
              "ozaki_t pair sum overflow");
        const long long v = ot::combine<G>(r);
        CHECK(v == exact, "ozaki_t combine: %lld vs %lld", v, exact);
        const double want = (double)v * (Tr::out_scale(colE[n]) * ot::group_weight<G>() * Tr::out_scale(rowE[j]));
        if (sizeof(Real) == 8) {   // OtScale<double>::apply: both scales inside the conversion, two DADDs
          using Td = oz::Traits<double>;
          const bool dead = colE[n] < Td::MIN_EF || colE[n] >= 2047;
          const ot::DoubleMagic m = ot::double_magic(dead ? 0 : Td::out_exp(colE[n]) - 8 * (G - 1), colE[n] >= 2047);
          CHECK(ot::to_double_scaled(v, ot::double_magic(0, false), 0) == (double)v, "ozaki_t to_double(%lld)", v);
          // the two-DADD conversion of the full 62-bit V: identical to the plain product
          const double full = ot::to_double_scaled(v, m, ot::row_word(rowE[j]));
          CHECK(full == want || (v == 0 && full == 0.0), "ozaki_t double scaling: %g vs %g", full, want);
          // the kernel's one-DADD conversion of V' = round(V / 2^11): within 2^11 units of V
          const long long v51 = ot::combine51(r);
          CHECK(std::llabs(v51) < (1ll << 51) && std::llabs(v51 * 2048 - v) <= 1300, "ozaki_t combine51: %lld vs %lld", v51 * 2048, v);
          out[part] = ot::to_double51_scaled(v51, m, ot::row_word(rowE[j]));
          const double unit = std::fabs(Tr::out_scale(colE[n]) * ot::group_weight<G>() * Tr::out_scale(rowE[j]));
          CHECK(std::fabs(out[part] - want) <= 1300.0 * unit + std::fabs(want) * 3e-16, "ozaki_t 51-bit conversion: %g vs %g", out[part], want);
        } else {                   // OtScale<float>::apply: FP32 only, row scale first
          using Tf = oz::Traits<float>;
          const float f = (ot::combine_f32(r) * Tf::out_scale_f(rowE[j], 0)) * Tf::out_scale_f(colE[n], -8 * (G - 1));
          out[part] = (double)f;
          // within 1.5 ulp of the correctly rounded result (two int32 -> float roundings and one FMA)
          CHECK(std::fabs(out[part] - want) <= std::fabs(want) * 1.8e-7 + 1e-44, "ozaki_t float scaling: %g vs %g", out[part], want);
        }
      }
      C[j + (size_t)Mrows * n] = std::complex<double>(out[0], out[1]);
    }
  long double num = 0, den = 0;
  for (int r = 0; r < Mrows; ++r)
    for (int n = 0; n < N; ++n) {
      long double rr = 0, ri = 0;
      for (int k = 0; k < K; ++k) {
        const cr a = A[r + (size_t)Mrows * k], b = B[n + (size_t)N * k];
        rr += (long double)a.real() * b.real() - (long double)a.imag() * b.imag();
        ri += (long double)a.real() * b.imag() + (long double)a.imag() * b.real();
      }
      const std::complex<double> got = C[r + (size_t)Mrows * n];
      if (r == 5 && constant == 0.0) continue;
      long double w = 1.0L;
      if (big != 1.0) w = ((r & 1) ? 1.0L / big : (long double)big) * ((n & 1) ? (long double)big : 1.0L / big);
      num += w * w * ((got.real() - rr) * (got.real() - rr) + (got.imag() - ri) * (got.imag() - ri));
      den += w * w * (rr * rr + ri * ri);
      if (r == 3 && constant == 0.0) CHECK(got == std::complex<double>(0, 0), "ozaki_t: zero row must give exact zeros");
    }
  return (double)std::sqrt(num / den);
}

static void test_ozaki_t(std::mt19937& rng) {
  struct Case { int M, N, K; double sigma, sparsity, tol64, tol32; };
  const Case cases[] = {
      {64, 64, 64, 0.0, 0.0, 1e-12, 1e-7}, {64, 64, 64, 3.0, 0.0, 5e-11, 1e-6}, {50, 33, 40, 0.0, 0.3, 1e-12, 1e-7},
      {64, 64, 32, 0.0, 0.0, 1e-12, 1e-7}, {64, 64, 8, 0.0, 0.0, 1e-12, 1e-7},  {17, 5, 1, 0.0, 0.0, 1e-12, 1e-7},
      {64, 32, 17, 1.0, 0.5, 5e-12, 2e-7}, {64, 64, 48, 0.0, 0.0, 1e-12, 1e-7},
  };
  for (const Case& c : cases) {
    const double e64 = ozaki_t_tile_error<double>(rng, c.M, c.N, c.K, c.sigma, c.sparsity);
    const double e32 = ozaki_t_tile_error<float>(rng, c.M, c.N, c.K, c.sigma, c.sparsity);
    CHECK(e64 < c.tol64, "ozaki_t c128 M=%d N=%d K=%d: rel-L2 %.3e", c.M, c.N, c.K, e64);
    CHECK(e32 < c.tol32, "ozaki_t c64 M=%d N=%d K=%d: rel-L2 %.3e", c.M, c.N, c.K, e32);
    std::printf("ozaki_t M=%d N=%d K=%d sigma=%.0f: rel-L2 c128 %.2e, c64 %.2e\n", c.M, c.N, c.K, c.sigma, e64, e32);
  }
  {
    const double e64 = ozaki_t_tile_error<double>(rng, 64, 48, 64, 0.0, 0.0, 1e140);
    const double e32 = ozaki_t_tile_error<float>(rng, 64, 48, 64, 0.0, 0.0, 1e15);
    CHECK(e64 < 1e-12 && e32 < 1e-7, "ozaki_t scale range: %.3e %.3e", e64, e32);
  }
  for (double cst : {1.0, 1.9999999, -1.0, 1.0000001}) {   // largest sums: every product has the same sign
    const double e64 = ozaki_t_tile_error<double>(rng, 64, 64, 64, 0.0, 0.0, 1.0, cst);
    const double e32 = ozaki_t_tile_error<float>(rng, 64, 64, 64, 0.0, 0.0, 1.0, cst);
    CHECK(e64 < 1e-12 && e32 < 1e-7, "ozaki_t constant %.7f: %.3e %.3e", cst, e64, e32);
  }
  std::printf("ozaki_t (second-generation kernel arithmetic): ok\n");
}

template <class Real>
static void check_digits_by_hand() {
  using Tr = oz::Traits<Real>;
  constexpr int S = Tr::S;
  unsigned long long bias = 0;
  for (int i = 0; i < S; ++i) bias |= 128ull << (8 * i);
  CHECK((unsigned long long)Tr::BIAS == bias, "bias constant");
  Real x[16] = {-1, 1, 128, -129};
  oz::Word4 pl[S];
  Tr::slice16(x, (Real)1, false, pl);
  auto digit = [&](int s, int j) { return (int)(int8_t)((pl[s].w[j / 4] >> (8 * (j % 4))) & 0xff); };
  CHECK(digit(S - 1, 0) == -1 && digit(S - 2, 0) == 0 && digit(0, 0) == 0, "digits of -1");
  CHECK(digit(S - 1, 1) == 1 && digit(S - 2, 1) == 0, "digits of 1");
  CHECK(digit(S - 1, 2) == -128 && digit(S - 2, 2) == 1, "digits of 128 = 1*256 - 128");
  CHECK(digit(S - 1, 3) == 127 && digit(S - 2, 3) == -1, "digits of -129 = -256 + 127");
  CHECK(digit(S - 1, 4) == 0 && digit(0, 15) == 0, "digits of 0");
  Tr::slice16(x, (Real)1, true, pl);
  CHECK(digit(S - 1, 0) == 1 && digit(S - 1, 2) == -128 && digit(S - 2, 2) == 0 && digit(S - 2, 3) == 1,
        "negated digits");
}

static void test_ozaki_lowering() {
  // ComplexF32 GEMM-shaped steps are lowered with the gather fused only under option cgemm_ozaki
  std::vector<int64_t> ad(20, 2), bd(12, 2);
  std::vector<int32_t> ai, bi;
  int o = 0, k = 0;
  for (int i = 0; i < 20; ++i) {
    if (i == 3 || i == 4 || i == 5 || i == 15 || i == 17 || i == 19) ai.push_back(++k);
    else ai.push_back(-(++o));
  }
  for (int i = 6; i >= 1; --i) bi.push_back(i);
  for (int j = 0; j < 6; ++j) bi.push_back(-(o + 1 + j));
  Options opt;
  ContractPlan P = lower_contract(ad, ai, bd, bi, 8, opt);
  CHECK(P.kind == CK_GEMM && P.fused_gemm && P.tempA_bytes == 0, "c64 default (ozaki_auto): gather fused, k_ozaki_t");
  opt.ozaki_auto = 0;
  P = lower_contract(ad, ai, bd, bi, 8, opt);
  CHECK(P.kind == CK_GEMM && !P.fused_gemm && P.tempA_bytes > 0, "c64 with ozaki_auto = 0: materialised TTGT");
  {   // below the policy's envelope (M < 4096) the materialised path stays
    std::vector<int64_t> as(12, 2);
    std::vector<int32_t> ais = {-1, -2, -3, 1, 2, 3, -4, -5, -6, 4, 5, 6};
    Options od;
    ContractPlan Q = lower_contract(as, ais, bd, bi, 8, od);
    CHECK(Q.kind == CK_GEMM && !Q.fused_gemm && Q.M == 64, "c64 default, small M: materialised TTGT");
    CHECK(ozaki_t_preferred(16, 1 << 18, 64, 64) && !ozaki_t_preferred(16, 1 << 18, 64, 8) &&
              !ozaki_t_preferred(16, 1 << 18, 8, 64) && ozaki_t_preferred(8, 1 << 18, 64, 8) &&
              !ozaki_t_preferred(8, 1 << 18, 65, 8) && !ozaki_t_preferred(16, 1000, 64, 64),
          "ozaki_t_preferred envelope");
  }
  opt.cgemm_ozaki = 4;
  P = lower_contract(ad, ai, bd, bi, 8, opt);
  CHECK(P.kind == CK_GEMM && P.fused_gemm && P.tempA_bytes == 0 && P.M == (1 << 14) && P.N == 64 && P.K == 64,
        "c64 with cgemm_ozaki: fused gather");
  P = lower_contract(ad, ai, bd, bi, 16, opt);
  CHECK(P.kind == CK_GEMM && P.fused_gemm, "c128 unaffected by cgemm_ozaki");
  {   // c64, short contraction (K = 8) with 64 open on the small side: output-bound, the small-operand
      // tensor-core kernel (S > 16 is its territory only) -- unless it is switched off or the INT8
      // kernel is forced; c128 keeps the GEMM path (DMMA thin / skinny kernels); the dot split
    std::vector<int64_t> aw(19, 2), bw(9, 2);
    std::vector<int32_t> aiw, biw;
    int ow = 0, kw = 0;
    for (int i = 0; i < 19; ++i) {
      if (i >= 16) aiw.push_back(++kw);
      else aiw.push_back(-(++ow));
    }
    for (int i = 1; i <= 3; ++i) biw.push_back(i);
    for (int j = 0; j < 6; ++j) biw.push_back(-(ow + 1 + j));
    Options ow0;
    ContractPlan W = lower_contract(aw, aiw, bw, biw, 8, ow0);
    CHECK(W.kind == CK_SMALL_RIGHT && W.M == (1 << 16) && W.N == 64 && W.K == 8, "c64 K = 8, N = 64: small-operand tensor-core kernel");
    ow0.small_tc = 1;
    W = lower_contract(aw, aiw, bw, biw, 8, ow0);
    CHECK(W.kind == CK_GEMM && W.fused_gemm, "c64 K = 8, N = 64 with small_tc = 1: INT8 kernel");
    Options ow1;
    ow1.cgemm_ozaki = 4;
    W = lower_contract(aw, aiw, bw, biw, 8, ow1);
    CHECK(W.kind == CK_GEMM && W.fused_gemm, "c64 K = 8, N = 64 with cgemm_ozaki forced: INT8 kernel");
    Options ow2;
    W = lower_contract(aw, aiw, bw, biw, 16, ow2);
    CHECK(W.kind == CK_GEMM, "c128 K = 8, N = 64: GEMM path");
    // inner product over 2^21 elements with power-of-two extents: 2^18 threads, 8 values of j each
    std::vector<int64_t> dv(21, 2);
    std::vector<int32_t> di;
    for (int i = 1; i <= 21; ++i) di.push_back(i);
    ContractPlan D = lower_contract(dv, di, dv, di, 16, ow2);
    CHECK(D.kind == CK_DOT && D.dot_split == 18 && D.dot_blocks == 1024, "dot over 2^21: offset split");
    std::vector<int64_t> d3 = {3, 7, 1000};
    std::vector<int32_t> d3i = {1, 2, 3};
    D = lower_contract(d3, d3i, d3, d3i, 16, ow2);
    CHECK(D.kind == CK_DOT && D.dot_split == 0, "dot over non-power-of-two extents: generic loop");
  }
  std::printf("ozaki lowering: ok\n");
}

static void test_ozaki(std::mt19937& rng) {
  test_ozaki_lowering();
  CHECK(oz::pow2_field(1023) == 1.0 && oz::pow2_field(1033) == 1024.0, "pow2_field");
  CHECK(oz::pow2_field_f(127) == 1.0f && oz::pow2_field_f(137) == 1024.0f, "pow2_field_f");
  CHECK(oz::Traits<float>::out_scale(126 + 6) == 1.0 && oz::Traits<double>::out_scale(1022 + 6) == 1.0,
        "out_scale");
  check_digits_by_hand<double>();
  check_digits_by_hand<float>();
  test_ozaki_t(rng);
}

int main() {
  std::mt19937 rng(12345);
  try {
    test_permutations(rng);
    test_contractions(rng, 0, 600);
    test_contractions(rng, 1, 40);
    test_contractions(rng, 2, 30);
    test_contractions(rng, 3, 6);
    test_big_shapes();
    test_ozaki(rng);
  } catch (const Error& e) {
    std::printf("FAIL: exception %d %s\n", e.code, e.what());
    return 2;
  }
  if (failures) {
    std::printf("FAILED %d checks\n", failures);
    return 1;
  }
  std::printf("ALL OK\n");
  return 0;
}
