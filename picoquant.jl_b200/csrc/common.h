// Shared host/device definitions for libpq_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/pq_b200.h"

namespace pq {

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define PQ_CUDA(call)                                                              \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess)                                                        \
      throw ::pq::Error(PQ_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

#define PQ_REQUIRE(cond, code, msg)            \
  do {                                         \
    if (!(cond)) throw ::pq::Error((code), (msg)); \
  } while (0)

// ---------------------------------------------------------------------------
// index maps: linear index over a list of fused dims -> element offset
// ---------------------------------------------------------------------------
constexpr int MAXF = 40;  // max fused dims (extent >= 2 each => 2^40 elements)

struct IdxMap {
  int nd;
  int pow2;            // every extent is a power of two -> shifts instead of div/mod
  int sh[MAXF];        // log2(ext) when pow2
  int64_t ext[MAXF];
  int64_t str[MAXF];   // element stride of each fused dim
};

__host__ __device__ inline int64_t map_offset(const IdxMap& m, int64_t i) {
  int64_t off = 0;
  if (m.pow2) {
#pragma unroll 1
    for (int d = 0; d < m.nd; ++d) {
      off += (i & (m.ext[d] - 1)) * m.str[d];
      i >>= m.sh[d];
    }
  } else {
#pragma unroll 1
    for (int d = 0; d < m.nd; ++d) {
      int64_t q = i / m.ext[d];
      off += (i - q * m.ext[d]) * m.str[d];
      i = q;
    }
  }
  return off;
}

// ---------------------------------------------------------------------------
// tiled bit-permutation parameters (kernels_permute.cu)
// ---------------------------------------------------------------------------
constexpr int MAXTILEBITS = 12;
struct TileParams {
  int n;        // log2(total elements)
  int t;        // tile bits
  int a, b;     // low input / output bits that are contiguous inside the tile (>= 5)
  int tin_pos[MAXTILEBITS];   // input bit position of tile bit u (input order)
  int tout_pos[MAXTILEBITS];  // output bit position of tile bit v (output order)
  int emap[MAXTILEBITS];      // tile bit (input order) feeding output-order tile bit v
  int nrest;
  int rest_in[48];   // input bit position of tile-counter bit w
  int rest_out[48];  // output bit position of the same bit
  int nswz;
  int swz_src[4];    // smem swizzle: bit swz_src[k] of e is xor-ed into bit swz_dst[k]
  int swz_dst[4];
  long long ntiles;
};

// ---------------------------------------------------------------------------
// kernel classes (profiling / roofline bookkeeping)
// ---------------------------------------------------------------------------
enum KClass {
  KC_PERMUTE_TILED = 0,
  KC_PERMUTE_GENERIC = 1,
  KC_CONTRACT_SMALL = 2,
  KC_CONTRACT_DIRECT = 3,
  KC_CONTRACT_DOT = 4,
  KC_GEMM_SIMT = 5,
  KC_GEMM_TENSOR = 6,
  KC_VIEW = 7,
  KC_ACCUMULATE = 8,
  KC_COPY = 9,
  KC_ALLREDUCE = 10,
  KC_SVD = 11,
  KC_OTHER = 12,
  KC_CONTRACT_CHAIN = 13,  // k_contract_chain: a group of chains of tiny contractions
  KC_GEMM_INT8 = 14        // k_ozaki_t: GEMM-shaped steps on the INT8 tensor pipe (tcgen05 kind::i8)
};
static_assert(KC_GEMM_INT8 + 1 == PQ_NUM_KERNEL_CLASSES, "class count");

struct ProfRecord {
  cudaEvent_t e0, e1;
  int cls;
  double bytes, flops;
};

struct Options {
  int gemm = 0;     // 0 auto, 1 SIMT, 2 tensor (DMMA), 3 direct everywhere
  int permute = 0;  // 0 auto, 1 generic, 2 tiled
  int fused = 0;    // 0 auto, 1 disable the fused small-operand / dot kernels
  int zgemm_kfirst = 0;  // 1: always use the row-first gather order in the fused ZGEMM (A/B check)
  int small_tc = 0;    // 1 disables k_contract_small_c64tc (ComplexF32 small-operand steps with a contracted axis fastest, 3xTF32 mma.sync)
  int ozaki_tsw = 0;   // k_ozaki_t<double>: 2 = digit planes 0..3 of W in tensor memory (TS form of the MMA)
  int zgemm_thin = 0;  // N <= 16, K <= 64 steps: 0 auto (k_zgemm_thin when a contracted axis is the fastest of A, or K <= 8 with 16 < N <= 64), 1 off (tiled 128x8 / skinny), 2 / 3 force k_zgemm_thin on N <= 16, k-first / rows-first loads, 4 also K <= 16 with 16 < N <= 64
  int zgemm_cfg = 0;  // fused ZGEMM: 0 auto, 1 = 64x64 (2 CTAs/SM), 2 = 64x32 (4 CTAs/SM), 3 = 128x8, 4 = 64x32 3M, 7 = persistent skinny where eligible
  int zgemm_3m = 0;       // persistent skinny ZGEMM: 0 = 3M (three DMMAs per complex product), 1 = 4M
  int zgemm_skinny = 0;   // persistent skinny fused ZGEMM: 0 auto, 1 off
  int zgemm_ozaki = 0;    // 6: force the INT8 tensor-core kernel k_ozaki_t for every eligible ComplexF64 step (K, N <= 64)
  int cgemm_ozaki = 0;    // 4: the same for ComplexF32
  int ozaki_auto = 1;     // 1 (default): GEMM-shaped steps inside ozaki_t_preferred() run on k_ozaki_t; 0: only when forced
  int zgemm_stagger = 0;  // ns of start delay per resident-CTA slot in the first wave (0 = off)
  int chain = 0;    // compiled programs: 0 = batch chains of tiny contractions into one launch, 1 = off
  int prio = 0;     // 0: small-grid graph nodes get the highest launch priority, 1: off
  int graph = 0;    // 0 CUDA graph with parallel branches, 1 eager replay, 2 single-chain graph
};

// launch context handed to every kernel launcher
struct Launch {
  cudaStream_t stream = nullptr;
  int elem_size = 16;   // 8 (c64) or 16 (c128)
  int num_sms = 148;
  int64_t* launch_counter = nullptr;
  // profiling: eager launches record plain events; while a profiling graph is captured
  // (`external`), the events become event-record NODES of the graph (cudaEventRecordExternal)
  bool profile = false;
  bool external = false;
  std::vector<ProfRecord>* prof = nullptr;
  mutable ProfRecord cur{};
  void begin(int cls, double bytes, double flops) const;
  void end() const;
  const Options* opt = nullptr;
};

// ---------------------------------------------------------------------------
// lowering results
// ---------------------------------------------------------------------------
struct PermutePlan {
  bool identity = true;
  bool tiled = false;
  bool paired = false;   // c64 only: lowest axis untouched, pairs move as 16-byte elements
  int64_t total = 1;
  IdxMap gmap{};     // output-order fused dims with input strides
  TileParams tp{};
};

PermutePlan lower_permute(const std::vector<int64_t>& in_dims, const std::vector<int>& perm,
                          int elem_size, const Options& opt);
void run_permute(const Launch& L, const PermutePlan& p, const void* in, void* out);

enum ContractKind { CK_SMALL_RIGHT, CK_SMALL_LEFT, CK_DIRECT, CK_DOT, CK_GEMM };

struct ContractPlan {
  ContractKind kind = CK_DIRECT;
  int64_t M = 1, N = 1, K = 1;
  std::vector<int64_t> cdims;          // logical dims of C (A-open then B-open)
  IdxMap mA{}, kA{}, nB{}, kB{};       // fused index maps into A and B
  // TTGT
  PermutePlan permA, permB;            // A -> [M|K], B -> [N|K] canonical layouts
  size_t tempA_bytes = 0, tempB_bytes = 0, ws_bytes = 0;
  bool fused_gemm = false;             // gather straight from A and B inside the GEMM
  int dot_blocks = 0;
  int dot_split = 0;                   // > 0: k = k_low + j * 2^dot_split, k_low = the thread (k_contract_dot, power-of-two extents)
};

ContractPlan lower_contract(const std::vector<int64_t>& a_dims, const std::vector<int32_t>& a_idx,
                            const std::vector<int64_t>& b_dims, const std::vector<int32_t>& b_idx,
                            int elem_size, const Options& opt);
// tempA/tempB/ws must hold the sizes the plan asks for (may be null when 0)
void run_contract(const Launch& L, const ContractPlan& p, const void* A, const void* B, void* C,
                  void* tempA, void* tempB, void* ws);

// ---------------------------------------------------------------------------
// chains of tiny contractions (kernels_contract.cu, used by compiled programs)
// ---------------------------------------------------------------------------
constexpr int MINI_ND = 12;
struct MiniMap {   // IdxMap for tensors of < 2^31 elements and <= MINI_ND fused dims
  int nd, pow2;
  int sh[MINI_ND], ext[MINI_ND], str[MINI_ND];
};
struct ChainItem {   // C[m + M n] = sum_k A[mA(m) + kA(k)] * B[nB(n) + kB(k)]
  MiniMap mA, kA, nB, kB;
  int M, N, K, pad;
  const void* A;
  const void* B;
  void* C;
};
struct ChainRange {  // items [begin, begin + count) are executed in order by one CTA
  int begin, count;
};
// false when the plan does not fit a ChainItem (too many fused dims, sizes beyond int)
bool chain_item_from_plan(const ContractPlan& p, ChainItem& it);
void run_chains(const Launch& L, const ChainItem* d_items, const ChainRange* d_ranges, int nchains,
                double bytes = 0, double flops = 0);

// operands of the gather-fused GEMM kernels (kernels_zgemm.cu, kernels_zgemm_ozaki2.cu):
// C[m + M n] = sum_k A[mA(m) + kA(k)] * B[nB(n) + kB(k)]
struct FusedParams {
  IdxMap mA, kA, nB, kB;
  long long M, N, K;
  // De-phasing of the first wave: CTA b < first_wave starts (b / num_sms) * stagger_ns late, so
  // the CTAs that share an SM are in different phases of their tile (fill / DMMA / store) and
  // stay so for the whole grid, because every later CTA starts when an earlier one retires.
  int stagger_ns, first_wave, num_sms;
};

// INT8 tensor-core complex GEMM for the skinny sweep steps (kernels_zgemm_ozaki2.cu: k_ozaki_t)
void init_kernels_ozaki_t();
double run_ozaki_t_microbench(const Launch& L, const std::string& what);
void run_zgemm_ozaki_t(const Launch& L, const FusedParams& fp, const void* A, const void* B, void* C);
void ozaki_t_check_watchdog();   // throws PQ_ERR_CUDA if a k_ozaki_t launch hit its mbarrier watchdog (call on an idle stream)
// envelope of the kernel: all of K and N resident per tile
inline bool zgemm_ozaki_eligible(int64_t M, int64_t N, int64_t K) {
  return K >= 1 && K <= 64 && N >= 1 && N <= 64 && M >= 1;
}
// Default policy (option ozaki_auto): which GEMM-shaped steps run on k_ozaki_t, from the
// per-shape timings on B200 (profiles/ozaki_t_probe_r02.json).  ComplexF64: the steps the FP64
// tensor pipe bounds (K >= 32 with N >= 32: 155 vs 267 us at M = 2^18, N = K = 64); the
// HBM-bound K <= 16 steps stay on the persistent DMMA kernel (81 vs 108 us).  ComplexF32: every
// skinny step (the alternative is a K1 permute + tcgen05 3xTF32 on canonical layouts: 77 vs
// 245 us, and 4.5e-8 instead of 3.4e-7 relative error).
inline bool ozaki_t_preferred(int elem_size, int64_t M, int64_t N, int64_t K) {
  if (!zgemm_ozaki_eligible(M, N, K) || M < 4096) return false;
  if (elem_size == 16) return K >= 32 && N >= 32;
  return N > 16 || K > 16;
}
// ComplexF32 contraction with the gather fused, on the INT8 kernel (plan lowered with fused_gemm)
void run_cgemm_ozaki_fused(const Launch& L, const ContractPlan& cp, const void* A, const void* B, void* C);

void init_kernels();
void init_kernels_cgemm();

// misc kernels
void run_view(const Launch& L, const void* in, void* out, int64_t inner, int64_t ext_in,
              int64_t nsel, int64_t outer, int start0, const int32_t* start_dev);
void run_accumulate(const Launch& L, void* dst, const void* src, int64_t n);
// pq_save_tensors: `table` (device) holds n ScatterItem; item i copies `words` 8-byte words from
// the staging block at `src_off` to `dst`
struct ScatterItem {
  void* dst;
  unsigned long long src_off;
  unsigned long long words;
};
void run_scatter(const Launch& L, const unsigned char* stage_dev, const ScatterItem* table, int n,
                 double bytes);
double run_microbench(const Launch& L, const std::string& what);

// gemm back ends on canonical layouts A'[m + M k], B''[n + N k], C[m + M n]
void run_gemm_simt(const Launch& L, const void* A, const void* B, void* C, int64_t M, int64_t N,
                   int64_t K);
void run_zgemm_dmma(const Launch& L, const void* A, const void* B, void* C, int64_t M, int64_t N,
                    int64_t K);
void run_zgemm_fused(const Launch& L, const ContractPlan& cp, const void* A, const void* B, void* C);
void run_cgemm_tcgen05(const Launch& L, const void* A, const void* B, void* C, int64_t M, int64_t N,
                       int64_t K);

inline int64_t prod(const std::vector<int64_t>& v) {
  int64_t p = 1;
  for (auto x : v) p *= x;
  return p;
}
inline bool is_pow2(int64_t x) { return x > 0 && (x & (x - 1)) == 0; }
inline int ilog2(int64_t x) {
  int l = 0;
  while ((int64_t(1) << l) < x) ++l;
  return l;
}

}  // namespace pq
