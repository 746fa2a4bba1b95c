// Host-side lowering: turns (dims, ncon labels) / (dims, permutation) into fused
// index maps, tile parameters and a kernel choice.  Pure integer code, no CUDA
// calls -- also compiled into the CPU test harness (test_lower.cpp).
//
// Contraction semantics follow the reference's contract_tensors
// (src/layer1.jl:85-92 -> TensorOperations.tensorcontract): a label present in
// both index lists is contracted, all others are open, and the result axes are
// A's open axes (A order) followed by B's open axes (B order).  In GEMM terms
// C(M,N) = A'(M,K) * B'(K,N), column-major, so C never needs an output permute.
#include <algorithm>
#include <set>

#include "common.h"

namespace pq {

namespace {

struct D3 {
  int64_t ext, sa, sb;
};

void fill_map(IdxMap& m, const std::vector<D3>& dims, bool use_sb) {
  PQ_REQUIRE((int)dims.size() <= MAXF, PQ_ERR_UNSUPPORTED, "too many fused dims");
  m.nd = (int)dims.size();
  m.pow2 = 1;
  for (int i = 0; i < m.nd; ++i) {
    m.ext[i] = dims[i].ext;
    m.str[i] = use_sb ? dims[i].sb : dims[i].sa;
    if (!is_pow2(dims[i].ext)) m.pow2 = 0;
    m.sh[i] = ilog2(dims[i].ext);
  }
  for (int i = m.nd; i < MAXF; ++i) {
    m.ext[i] = 1;
    m.str[i] = 0;
    m.sh[i] = 0;
  }
}

// drop extent-1 dims and merge neighbours that are contiguous in every stride set
std::vector<D3> fuse(const std::vector<D3>& in, bool both) {
  std::vector<D3> out;
  for (const D3& d : in) {
    if (d.ext == 1) continue;
    if (!out.empty()) {
      D3& p = out.back();
      bool ca = d.sa == p.sa * p.ext;
      bool cb = !both || d.sb == p.sb * p.ext;
      if (ca && cb) {
        p.ext *= d.ext;
        continue;
      }
    }
    out.push_back(d);
  }
  return out;
}

std::vector<int64_t> strides_of(const std::vector<int64_t>& dims) {
  std::vector<int64_t> s(dims.size());
  int64_t acc = 1;
  for (size_t i = 0; i < dims.size(); ++i) {
    s[i] = acc;
    acc *= dims[i];
  }
  return s;
}

}  // namespace

// ---------------------------------------------------------------------------
// permutation
// ---------------------------------------------------------------------------
PermutePlan lower_permute(const std::vector<int64_t>& in_dims, const std::vector<int>& perm,
                          int elem_size, const Options& opt) {
  const int rank = (int)in_dims.size();
  PQ_REQUIRE((int)perm.size() == rank, PQ_ERR_INVALID, "permutation length != rank");
  std::vector<char> seen(rank, 0);
  for (int p : perm) {
    PQ_REQUIRE(p >= 0 && p < rank && !seen[p], PQ_ERR_INVALID, "not a permutation");
    seen[p] = 1;
  }
  PermutePlan P;
  P.total = prod(in_dims);
  std::vector<int64_t> istr = strides_of(in_dims);
  std::vector<D3> list;
  for (int k = 0; k < rank; ++k) list.push_back({in_dims[perm[k]], istr[perm[k]], 0});
  list = fuse(list, false);
  P.identity = list.empty() || (list.size() == 1 && list[0].sa == 1) || P.total <= 1;
  int es = elem_size;
  if (!P.identity && elem_size == 8 && list[0].sa == 1 && list[0].ext % 2 == 0) {
    // ComplexF32 with an untouched even lowest axis: permute (re,im,re,im) pairs as
    // 16-byte elements -- same kernel as c128 on half as many elements.
    P.paired = true;
    list[0].ext /= 2;
    for (size_t i = 1; i < list.size(); ++i) list[i].sa /= 2;
    list = fuse(list, false);
    es = 16;
  }
  fill_map(P.gmap, list, false);
  if (P.identity) return P;
  const int64_t total = P.paired ? P.total / 2 : P.total;

  bool all_pow2 = P.gmap.pow2 != 0;
  if (!all_pow2 || total < 4096 || opt.permute == 1) return P;

  // ---- bit permutation: output bit j is input bit src[j] ------------------------
  const int n = ilog2(total);
  if (n > 47) return P;
  std::vector<int> src(n), dst(n);
  {
    int j = 0;
    for (const D3& d : list) {
      int e = ilog2(d.ext), s = ilog2(d.sa);
      for (int q = 0; q < e; ++q) src[j++] = s + q;
    }
    for (int q = 0; q < n; ++q) dst[src[q]] = q;
  }
  const int t_target = std::min(n, es == 16 ? 10 : 11);
  std::vector<char> inT(n, 0);
  int count = 0;
  auto add = [&](int bit) {
    if (!inT[bit]) {
      inT[bit] = 1;
      ++count;
    }
  };
  const int lo = std::min(5, n);
  for (int q = 0; q < lo; ++q) add(q);
  for (int q = 0; q < lo; ++q) add(src[q]);
  bool turn_in = true;
  while (count < t_target) {
    if (turn_in) {
      for (int q = 0; q < n; ++q)
        if (!inT[q]) {
          add(q);
          break;
        }
    } else {
      for (int q = 0; q < n; ++q)
        if (!inT[src[q]]) {
          add(src[q]);
          break;
        }
    }
    turn_in = !turn_in;
  }
  TileParams& tp = P.tp;
  tp.n = n;
  tp.t = count;
  PQ_REQUIRE(tp.t <= MAXTILEBITS, PQ_ERR_UNSUPPORTED, "tile too large");
  std::vector<int> tin;  // tile bits, ascending input position
  for (int q = 0; q < n; ++q)
    if (inT[q]) tin.push_back(q);
  std::vector<int> tout = tin;  // same bits, ascending output position
  std::sort(tout.begin(), tout.end(), [&](int x, int y) { return dst[x] < dst[y]; });
  for (int u = 0; u < tp.t; ++u) tp.tin_pos[u] = tin[u];
  for (int v = 0; v < tp.t; ++v) {
    tp.tout_pos[v] = dst[tout[v]];
    tp.emap[v] = int(std::find(tin.begin(), tin.end(), tout[v]) - tin.begin());
  }
  tp.a = 0;
  while (tp.a < n && inT[tp.a]) ++tp.a;
  tp.b = 0;
  while (tp.b < n && inT[src[tp.b]]) ++tp.b;
  tp.nrest = 0;
  for (int q = 0; q < n; ++q)
    if (!inT[q]) {
      tp.rest_in[tp.nrest] = q;
      tp.rest_out[tp.nrest] = dst[q];
      ++tp.nrest;
    }
  tp.ntiles = 1LL << tp.nrest;
  // xor swizzle: the c lowest output bits must spread over the c lowest slot bits
  const int c = es == 16 ? 3 : 4;
  std::vector<int> hi;
  std::vector<char> lo_used(c, 0);
  for (int v = 0; v < c && v < tp.t; ++v) {
    if (tp.emap[v] < c)
      lo_used[tp.emap[v]] = 1;
    else
      hi.push_back(tp.emap[v]);
  }
  tp.nswz = 0;
  int f = 0;
  for (int u : hi) {
    while (f < c && lo_used[f]) ++f;
    if (f >= c) break;
    tp.swz_src[tp.nswz] = u;
    tp.swz_dst[tp.nswz] = f;
    ++tp.nswz;
    ++f;
  }
  P.tiled = (tp.a >= 5 && tp.b >= 5);
  return P;
}

// ---------------------------------------------------------------------------
// contraction
// ---------------------------------------------------------------------------
ContractPlan lower_contract(const std::vector<int64_t>& a_dims, const std::vector<int32_t>& a_idx,
                            const std::vector<int64_t>& b_dims, const std::vector<int32_t>& b_idx,
                            int elem_size, const Options& opt) {
  PQ_REQUIRE(a_dims.size() == a_idx.size(), PQ_ERR_INVALID, "A: index list length != rank");
  PQ_REQUIRE(b_dims.size() == b_idx.size(), PQ_ERR_INVALID, "B: index list length != rank");
  {
    std::set<int32_t> sa(a_idx.begin(), a_idx.end()), sb(b_idx.begin(), b_idx.end());
    PQ_REQUIRE(sa.size() == a_idx.size() && sb.size() == b_idx.size(), PQ_ERR_UNSUPPORTED,
               "repeated label inside one tensor (partial trace) is not supported");
  }
  std::map<int32_t, int> bpos, apos;
  for (int i = 0; i < (int)b_idx.size(); ++i) bpos[b_idx[i]] = i;
  for (int i = 0; i < (int)a_idx.size(); ++i) apos[a_idx[i]] = i;

  std::vector<int> a_open, a_con, b_con, b_open;
  for (int i = 0; i < (int)a_idx.size(); ++i) {
    auto it = bpos.find(a_idx[i]);
    if (it == bpos.end()) {
      a_open.push_back(i);
    } else {
      a_con.push_back(i);
      b_con.push_back(it->second);
      PQ_REQUIRE(a_dims[i] == b_dims[it->second], PQ_ERR_SHAPE,
                 "DimensionMismatch on a contracted axis");
    }
  }
  for (int i = 0; i < (int)b_idx.size(); ++i)
    if (!apos.count(b_idx[i])) b_open.push_back(i);

  ContractPlan P;
  std::vector<int64_t> sa = strides_of(a_dims), sb = strides_of(b_dims);
  std::vector<D3> md, nd, kd;
  for (int i : a_open) {
    md.push_back({a_dims[i], sa[i], 0});
    P.cdims.push_back(a_dims[i]);
    P.M *= a_dims[i];
  }
  for (int i : b_open) {
    nd.push_back({b_dims[i], sb[i], 0});
    P.cdims.push_back(b_dims[i]);
    P.N *= b_dims[i];
  }
  for (size_t q = 0; q < a_con.size(); ++q) {
    kd.push_back({a_dims[a_con[q]], sa[a_con[q]], sb[b_con[q]]});
    P.K *= a_dims[a_con[q]];
  }
  md = fuse(md, false);
  nd = fuse(nd, false);
  kd = fuse(kd, true);
  fill_map(P.mA, md, false);
  fill_map(P.nB, nd, false);
  fill_map(P.kA, kd, false);
  fill_map(P.kB, kd, true);

  const int64_t M = P.M, N = P.N, K = P.K;
  const bool fused_ok = opt.fused == 0 && opt.gemm != 3;
  const int64_t SMALL_Q = 1024;  // elements of the small operand kept in shared memory
  // c128, one open bond of 8-16 on the small side, K >= 32: the per-row FP64 FMA work of the
  // small-operand kernel (2 N K per row) exceeds what HBM delivers; the narrow-tile DMMA
  // GEMM (128x8 tiles) is HBM-bound instead
  const bool narrow_gemm = elem_size == 16 && opt.fused == 0 && (opt.gemm == 0 || opt.gemm == 2) &&
                           N >= 8 && N <= 16 && K >= 32 && K <= 1024 && M >= 4096 &&
                           (opt.zgemm_cfg == 0 || opt.zgemm_cfg == 3);
  if (opt.gemm == 3) {
    P.kind = CK_DIRECT;
  } else if (narrow_gemm) {
    P.kind = CK_GEMM;
  } else if (fused_ok && N <= 16 && K <= 256 && N * K <= SMALL_Q && M >= N) {
    P.kind = CK_SMALL_RIGHT;
  } else if (fused_ok && M <= 16 && K <= 256 && M * K <= SMALL_Q) {
    P.kind = CK_SMALL_LEFT;
  } else if (fused_ok && elem_size == 8 && opt.small_tc == 0 && opt.cgemm_ozaki == 0 && K <= 16 && N <= 64 &&
             M >= 4096) {
    // c64, short contraction, up to 64 open on the small side: output-bound; the small-operand
    // tensor-core kernel (k_contract_small_c64tc) writes C at the speed the INT8 kernel's
    // epilogue cannot (58 us for 151 MB at K = 8)
    P.kind = CK_SMALL_RIGHT;
  } else if (fused_ok && elem_size == 8 && opt.small_tc == 0 && opt.cgemm_ozaki == 0 && K <= 16 && M <= 64 &&
             N >= 4096) {
    P.kind = CK_SMALL_LEFT;
  } else if (fused_ok && M * N <= 16 && K >= 512) {
    P.kind = CK_DOT;
  } else if (M * N * K <= 65536 || (M * N <= 64 && K <= 4096)) {
    P.kind = CK_DIRECT;
  } else {
    P.kind = CK_GEMM;
  }

  if (P.kind == CK_DOT) {
    // A sum over k may visit k in any order: enumerate the contracted bits so that the five
    // lowest (= the lanes of a warp) alternate between the lowest-address bits of A and of B.
    // Both operands are then read in runs of >= 4 consecutive elements (whole 32-byte
    // sectors) instead of one of them being a 16-byte gather.
    bool pow2 = true;
    for (const D3& d : kd) pow2 = pow2 && is_pow2(d.ext);
    if (pow2 && kd.size() > 1) {
      std::vector<D3> bits;
      for (const D3& d : kd)
        for (int q = 0; q < ilog2(d.ext); ++q) bits.push_back({2, d.sa << q, d.sb << q});
      std::vector<D3> order;
      std::vector<char> used(bits.size(), 0);
      for (int pick = 0; pick < 6 && order.size() < bits.size(); ++pick) {
        int best = -1;
        for (int i = 0; i < (int)bits.size(); ++i) {
          if (used[i]) continue;
          const int64_t key = (pick & 1) ? bits[i].sb : bits[i].sa;
          if (best < 0 || key < ((pick & 1) ? bits[best].sb : bits[best].sa)) best = i;
        }
        used[best] = 1;
        order.push_back(bits[best]);
      }
      for (size_t i = 0; i < bits.size(); ++i)
        if (!used[i]) order.push_back(bits[i]);
      order = fuse(order, true);
      if ((int)order.size() <= MAXF) {
        fill_map(P.kA, order, false);
        fill_map(P.kB, order, true);
      }
    }
    int64_t blocks = (K + 1023) / 1024;   // >= 4 k per thread
    if (blocks > 1184) blocks = 1184;     // 8 resident CTAs on each of the 148 SMs
    if (blocks < 1) blocks = 1;
    // M = N = 1 over power-of-two extents: the offset of k = k_low + j 2^s is the sum of the
    // offsets of its two parts (disjoint address bits), so a thread computes the offset of ITS
    // k_low once and walks j through a small table instead of re-deriving every offset from the
    // fused index map (that arithmetic, not HBM, set the speed: 46 us for 67 MB).  2^s threads,
    // 8..64 values of j each.
    P.dot_split = 0;
    if (pow2 && is_pow2(K) && M * N == 1 && K >= (int64_t(1) << 14)) {
      int64_t per = 8, b2 = K / (256 * per);
      while (b2 > 2048 && per < 64) {
        per *= 2;
        b2 /= 2;
      }
      if (b2 <= 2048) {
        blocks = b2;
        P.dot_split = ilog2(b2 * 256);
      }
    }
    P.dot_blocks = (int)blocks;
    P.ws_bytes = size_t(blocks) * 16 * elem_size;
  }
  if (P.kind == CK_GEMM && elem_size == 16 && opt.fused == 0 && (opt.gemm == 0 || opt.gemm == 2) &&
      K <= 1024) {
    // fused TTGT: k offsets are tabulated as 32-bit element offsets
    int64_t ka = 0, kb = 0;
    for (int d = 0; d < P.kA.nd; ++d) {
      ka += (P.kA.ext[d] - 1) * P.kA.str[d];
      kb += (P.kB.ext[d] - 1) * P.kB.str[d];
    }
    P.fused_gemm = ka < (int64_t(1) << 31) && kb < (int64_t(1) << 31);
  }
  if (P.kind == CK_GEMM && elem_size == 8 && opt.fused == 0 && (opt.gemm == 0 || opt.gemm == 2) &&
      ((opt.cgemm_ozaki != 0 && zgemm_ozaki_eligible(M, N, K)) ||
       (opt.cgemm_ozaki == 0 && opt.ozaki_auto != 0 && ozaki_t_preferred(8, M, N, K)))) {
    // ComplexF32 skinny steps on the INT8 tensor-core kernel, gather fused (no K1 pass)
    int64_t ka = 0, kb = 0;
    for (int d = 0; d < P.kA.nd; ++d) {
      ka += (P.kA.ext[d] - 1) * P.kA.str[d];
      kb += (P.kB.ext[d] - 1) * P.kB.str[d];
    }
    P.fused_gemm = ka < (int64_t(1) << 31) && kb < (int64_t(1) << 31);
  }
  if (P.kind == CK_GEMM && !P.fused_gemm) {
    std::vector<int> pa = a_open, pb = b_open;
    pa.insert(pa.end(), a_con.begin(), a_con.end());
    pb.insert(pb.end(), b_con.begin(), b_con.end());
    P.permA = lower_permute(a_dims, pa, elem_size, opt);
    P.permB = lower_permute(b_dims, pb, elem_size, opt);
    if (!P.permA.identity) P.tempA_bytes = size_t(M) * K * elem_size;
    if (!P.permB.identity) P.tempB_bytes = size_t(N) * K * elem_size;
  }
  return P;
}

}  // namespace pq
