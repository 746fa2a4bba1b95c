// execute_dsl_file on the device (reference interpreter: src/layer1.jl:211-315; grammar
// emitted by src/backends/dsl.jl:65-214).  A `.tl` command stream is compiled once:
// shapes are propagated symbolically, every contraction is lowered to its kernel(s),
// every intermediate (results, TTGT temporaries, split-K workspaces) gets a fixed
// offset in one arena by replaying the alloc/free timeline through a best-fit
// allocator, and the whole launch sequence is captured into a CUDA graph.  Replaying
// the graph costs one host call per slice; `view` start indices live in device
// memory so the same graph serves every slice of a sliced contraction
// (reference flow: src/layer2/slicing.jl:100-110 + examples/dist_slicing_example.jl).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <sstream>

#include "handle.h"

using namespace pq;

namespace {

constexpr size_t ALIGN = 256;
inline size_t round_up(size_t x) { return (x + ALIGN - 1) / ALIGN * ALIGN; }

// Host -> device upload that has LANDED when the call returns.  cudaMemcpy from pageable
// memory may return once the data sits in the driver's staging buffer, and the handle's /
// lanes' streams are non-blocking, i.e. not ordered after the legacy default stream: a graph
// launched right afterwards could read the destination before the DMA has finished.
void upload_now(void* dst, const void* src, size_t bytes) {
  cudaStream_t up = nullptr;   // per call: these uploads happen a few times per program
  PQ_CUDA(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, up);
  if (e == cudaSuccess) e = cudaStreamSynchronize(up);
  cudaStreamDestroy(up);
  PQ_CUDA(e);
}

// offset allocator used at compile time only
struct ArenaSim {
  std::map<size_t, size_t> free_blocks;  // offset -> size
  size_t top = 0;
  size_t alloc(size_t bytes) {
    bytes = round_up(bytes == 0 ? 1 : bytes);
    auto best = free_blocks.end();
    for (auto it = free_blocks.begin(); it != free_blocks.end(); ++it)
      if (it->second >= bytes && (best == free_blocks.end() || it->second < best->second)) best = it;
    if (best != free_blocks.end()) {
      size_t off = best->first, sz = best->second;
      free_blocks.erase(best);
      if (sz > bytes) free_blocks[off + bytes] = sz - bytes;
      return off;
    }
    // grow: extend a trailing free block if there is one
    if (!free_blocks.empty()) {
      auto last = std::prev(free_blocks.end());
      if (last->first + last->second == top) {
        size_t off = last->first;
        top = off + bytes;
        free_blocks.erase(last);
        return off;
      }
    }
    size_t off = top;
    top += bytes;
    return off;
  }
  void release(size_t off, size_t bytes) {
    bytes = round_up(bytes == 0 ? 1 : bytes);
    auto it = free_blocks.emplace(off, bytes).first;
    auto nx = std::next(it);
    if (nx != free_blocks.end() && it->first + it->second == nx->first) {
      it->second += nx->second;
      free_blocks.erase(nx);
    }
    if (it != free_blocks.begin()) {
      auto pv = std::prev(it);
      if (pv->first + pv->second == it->first) {
        pv->second += it->second;
        free_blocks.erase(it);
      }
    }
  }
};

struct Sym {
  bool leaf = false;
  std::shared_ptr<Buffer> leafbuf;  // keeps a bound handle tensor alive
  size_t offset = 0, bytes = 0;     // arena placement when !leaf
  bool small = false;               // lives in the no-reuse arena for small tensors
  bool dep = false;                 // depends on a `view` (slice) parameter
  bool used_by_dep = false;         // slice-invariant tensor read by a slice-dependent step
  std::vector<int64_t> dims;
};

enum StepKind { ST_CONTRACT, ST_PERMUTE, ST_VIEW, ST_SAVE };

struct Ref {  // pointer = leaf ? leafptr : (small ? arena_small : arena) + offset
  void* leafptr = nullptr;
  size_t offset = 0;
  bool leaf = false;
  bool small = false;
  bool dep = false;   // which of the two big arenas (slice-invariant / slice-dependent)
  bool null = true;
};

// Tensors up to this size are bump-allocated and never recycled within a program, so
// that the only hazards between the ~10^3 tiny contractions of a slice are true data
// dependencies and independent world-lines can run on parallel graph branches.
constexpr size_t SMALL_TENSOR_BYTES = 256 * 1024;

struct Step {
  StepKind kind;
  bool dep = false;  // slice-dependent (must run for every slice) vs slice-invariant
  ContractPlan cp;
  PermutePlan pp;
  Ref a, b, c, ta, tb, ws;
  // view
  int64_t inner = 1, ext = 1, nsel = 1, outer = 1;
  int start0 = 1, view_slot = -1;
  // save
  std::string key;
  std::vector<int64_t> dims;
  int save_slot = -1;  // index into LaneMem::outs
};

std::vector<int32_t> parse_ints(const std::string& s) {
  std::vector<int32_t> v;
  std::stringstream ss(s);
  std::string item;
  while (std::getline(ss, item, ',')) {
    if (item.empty()) continue;
    v.push_back((int32_t)std::stol(item));
  }
  return v;
}

}  // namespace

// Per-lane memory of a program.  Lane 0 is the program's own memory; further lanes are
// private copies (arenas, view parameters, `save` buffers, graph instances) that let
// pq_program_run_slices keep several slices in flight on separate streams.
struct LaneMem {
  void* arena2[2] = {nullptr, nullptr};  // [0] slice-invariant, [1] slice-dependent
  void* arena_small = nullptr;
  int32_t* d_starts = nullptr;
  // graph instances: invariant part (lane 0 only), dependent part, whole stream
  cudaGraphExec_t exec2[3] = {nullptr, nullptr, nullptr};
  void* chain_items[3] = {nullptr, nullptr, nullptr};   // resolved ChainItem arrays, per graph
  int64_t graph_launches[3] = {0, 0, 0};                // kernel launches inside each graph
  std::vector<std::shared_ptr<Buffer>> outs;  // one per `save` step
  cudaStream_t stream = nullptr;              // run_slices: the lane's own stream
  cudaEvent_t done_ev = nullptr, acc_ev = nullptr;
};

struct pq_program {
  std::vector<Step> steps;
  // [0]: slice-invariant steps, [1]: slice-dependent steps (disjoint arenas, so that the
  // invariant part can be executed once per amplitude and the dependent part per slice)
  size_t arena_bytes2[2] = {0, 0};
  size_t arena_bytes = 0, arena_small_bytes = 0;
  static constexpr int MAX_LANES = 8;
  std::vector<LaneMem> lanes;  // lanes[0] always exists
  int nsaves = 0;
  bool hoist = false;      // run() executes only the slice-dependent part
  bool prepared = false;   // the invariant part has been executed since hoisting was enabled
  int64_t macs2[2] = {0, 0}, launches2[2] = {0, 0}, ncontract2[2] = {0, 0};
  int nviews = 0;
  std::vector<int32_t> default_starts;
  // pinned ring for asynchronous parameter uploads
  static constexpr int RING = 32;
  int32_t* h_ring = nullptr;
  cudaEvent_t ring_ev[RING];
  bool ring_used[RING];
  int ring_pos = 0;
  // run_slices: view parameters of a whole batch of slices (pinned staging + device table)
  int32_t* h_table = nullptr;
  int32_t* d_table = nullptr;
  size_t table_cap = 0;
  cudaEvent_t table_ev = nullptr, batch_ev = nullptr;
  bool table_used = false;
  int64_t launches = 0, macs = 0, ncontract = 0, max_elems = 0;
  int64_t launches_batched = 0;   // launches per run when chains are batched (graph mode)
  int device = 0;
  // leaf tensors the program is bound to: (store key, pinned buffer)
  std::vector<std::pair<std::string, std::shared_ptr<Buffer>>> leaves;
  // what each leaf looked like when the program was compiled / last prepared: an in-place
  // re-save bumps Buffer::gen (hoisted invariant tensors are then stale), and must keep dims
  std::vector<uint64_t> leaf_gen;
  std::vector<std::vector<int64_t>> leaf_dims;
  // dependency DAG for multi-stream capture: deps[j] = earlier steps j must wait for
  std::vector<std::vector<int>> deps;
  // chains of tiny contractions that one CTA executes back to back; chains with the same
  // external dependencies form a group = one launch (see k_contract_chain)
  struct ChainGroup {
    std::vector<int> steps;          // member steps, chain by chain
    std::vector<ChainRange> ranges;  // one per chain, `begin` relative to the group's items
    std::vector<int> deps;           // steps outside the group that it waits for
    int first = 0;                   // issue position: the smallest member index
    int item_base = 0, range_base = 0;
    bool dep = false;
  };
  std::vector<ChainGroup> groups;
  std::vector<int> group_of;         // per step: its group or -1
  ChainRange* d_ranges = nullptr;
  static constexpr int NSTREAMS = 8;
  cudaStream_t side[NSTREAMS] = {nullptr};
  std::vector<cudaEvent_t> step_ev;
  cudaEvent_t fork_ev = nullptr;
};

namespace {
struct Range {
  const char* lo;
  const char* hi;
};
inline bool overlap(const Range& a, const Range& b) { return a.lo < b.hi && b.lo < a.hi; }
}  // namespace

// Where a step's operands live for a given lane.  `split` = the invariant part runs on its
// own (hoisting): slice-invariant tensors then exist once, in lane 0, for every lane.
struct Where {
  int lane = 0;
  bool split = false;
};

static void* resolve(const pq_program* p, const Ref& r, Where w = Where()) {
  if (r.null) return nullptr;
  if (r.leaf) return r.leafptr;
  const LaneMem& m = p->lanes[(w.split && !r.dep) ? 0 : w.lane];
  return (void*)((char*)(r.small ? m.arena_small : m.arena2[r.dep ? 1 : 0]) + r.offset);
}

static void issue_step(pq_handle* h, pq_program* p, Step& s, Launch& L, Where w);

// phase 0: slice-invariant steps, phase 1: slice-dependent steps (stream order within a phase)
static void issue_steps(pq_handle* h, pq_program* p, Launch& L, int phase, Where w) {
  for (Step& s : p->steps)
    if (phase < 0 || (s.dep ? 1 : 0) == phase) issue_step(h, p, s, L, w);
}

// read / write address ranges of a step (after the arena has been allocated)
static void step_ranges(pq_handle* h, const pq_program* p, const Step& s, std::vector<Range>& rd,
                        std::vector<Range>& wr) {
  const size_t es = h->elem_size;
  auto rng = [&](const Ref& r, size_t bytes) {
    const char* b = (const char*)resolve(p, r);
    return Range{b, b + (bytes ? bytes : 1)};
  };
  switch (s.kind) {
    case ST_CONTRACT:
      rd.push_back(rng(s.a, size_t(s.cp.M * s.cp.K) * es));
      rd.push_back(rng(s.b, size_t(s.cp.N * s.cp.K) * es));
      wr.push_back(rng(s.c, size_t(s.cp.M * s.cp.N) * es));
      if (!s.ta.null) wr.push_back(rng(s.ta, s.cp.tempA_bytes));
      if (!s.tb.null) wr.push_back(rng(s.tb, s.cp.tempB_bytes));
      if (!s.ws.null) wr.push_back(rng(s.ws, s.cp.ws_bytes));
      break;
    case ST_PERMUTE:
      rd.push_back(rng(s.a, size_t(s.pp.total) * es));
      wr.push_back(rng(s.c, size_t(s.pp.total) * es));
      break;
    case ST_VIEW:
      rd.push_back(rng(s.a, size_t(s.inner * s.ext * s.outer) * es));
      wr.push_back(rng(s.c, size_t(s.inner * s.nsel * s.outer) * es));
      break;
    case ST_SAVE: {
      size_t bytes = size_t(prod(s.dims)) * es;
      rd.push_back(rng(s.a, bytes));
      const char* o = (const char*)p->lanes[0].outs[s.save_slot]->ptr;
      wr.push_back(Range{o, o + (bytes ? bytes : 1)});
      break;
    }
  }
}

// RAW / WAR / WAW hazards between steps, including those created by arena reuse
static void build_dag(pq_handle* h, pq_program* p) {
  const int n = (int)p->steps.size();
  std::vector<std::vector<Range>> rd(n), wr(n);
  for (int j = 0; j < n; ++j) step_ranges(h, p, p->steps[j], rd[j], wr[j]);
  p->deps.assign(n, {});
  for (int j = 0; j < n; ++j) {
    for (int i = j - 1; i >= 0; --i) {
      bool hit = false;
      for (const Range& w : wr[j]) {
        for (const Range& r : rd[i]) hit = hit || overlap(w, r);
        for (const Range& r : wr[i]) hit = hit || overlap(w, r);
      }
      for (const Range& r : rd[j])
        for (const Range& w : wr[i]) hit = hit || overlap(r, w);
      if (hit) p->deps[j].push_back(i);
    }
  }
  if (getenv("PQ_B200_DEBUG")) {
    std::vector<int> depth(n, 1);
    int maxd = 0;
    size_t nd = 0;
    for (int j = 0; j < n; ++j) {
      for (int d : p->deps[j]) depth[j] = std::max(depth[j], depth[d] + 1);
      maxd = std::max(maxd, depth[j]);
      nd += p->deps[j].size();
    }
    fprintf(stderr, "[pq_b200] program DAG: %d steps, %zu hazard edges, critical path %d steps\n", n,
            nd, maxd);
  }
}

// Finds the chains of tiny contractions (each member depends on the previous member and on
// steps that precede the whole chain) and groups chains with identical outside dependencies.
// Members write to the never-recycled small arena, so running a chain earlier than its
// position in the stream cannot create a hazard.
static void build_chains(pq_handle* h, pq_program* p) {
  const int n = (int)p->steps.size();
  p->group_of.assign(n, -1);
  p->groups.clear();
  if (h->opt.chain != 0) return;
  auto tiny = [&](int j) {
    const Step& s = p->steps[j];
    if (s.kind != ST_CONTRACT) return false;
    const ContractPlan& c = s.cp;
    if (c.kind != CK_SMALL_RIGHT && c.kind != CK_SMALL_LEFT && c.kind != CK_DIRECT) return false;
    if (c.M * c.N > 8192 || c.K > 64 || c.M * c.K > 8192 || c.N * c.K > 8192) return false;
    if (!s.c.small) return false;
    ChainItem it;
    return chain_item_from_plan(c, it);
  };
  struct Chain {
    std::vector<int> m;
    bool dep;
  };
  std::vector<Chain> chains;
  std::vector<int> chain_of(n, -1);
  for (int j = 0; j < n; ++j) {
    if (!tiny(j)) continue;
    int best = -1;
    for (int d : p->deps[j]) {
      const int c = chain_of[d];
      if (c < 0 || chains[c].m.back() != d || chains[c].dep != p->steps[j].dep) continue;
      bool ok = true;
      for (int e : p->deps[j])
        if (e != d && chain_of[e] != c && e >= chains[c].m.front()) ok = false;
      if (ok) {
        best = c;
        break;
      }
    }
    if (best < 0) {
      chains.push_back(Chain{{j}, p->steps[j].dep});
      best = (int)chains.size() - 1;
    } else {
      chains[best].m.push_back(j);
    }
    chain_of[j] = best;
  }
  // group by (phase, outside dependencies)
  std::map<std::pair<bool, std::vector<int>>, int> index;
  for (size_t c = 0; c < chains.size(); ++c) {
    std::set<int> ext;
    for (int m : chains[c].m)
      for (int d : p->deps[m])
        if (chain_of[d] != (int)c) ext.insert(d);
    std::vector<int> key(ext.begin(), ext.end());
    auto it = index.find({chains[c].dep, key});
    int g;
    if (it == index.end()) {
      g = (int)p->groups.size();
      index[{chains[c].dep, key}] = g;
      pq_program::ChainGroup G;
      G.deps = key;
      G.dep = chains[c].dep;
      G.first = chains[c].m.front();
      p->groups.push_back(G);
    } else {
      g = it->second;
    }
    pq_program::ChainGroup& G = p->groups[g];
    G.ranges.push_back(ChainRange{(int)G.steps.size(), (int)chains[c].m.size()});
    for (int m : chains[c].m) {
      G.steps.push_back(m);
      p->group_of[m] = g;
    }
    if (chains[c].m.front() < G.first) G.first = chains[c].m.front();
  }
  // a group of one single-step chain is just that step
  for (size_t g = 0; g < p->groups.size(); ++g)
    if (p->groups[g].steps.size() == 1) {
      p->group_of[p->groups[g].steps[0]] = -1;
      p->groups[g].steps.clear();
      p->groups[g].ranges.clear();
    }
  int ib = 0, rb = 0;
  std::vector<ChainRange> all;
  for (auto& G : p->groups) {
    G.item_base = ib;
    G.range_base = rb;
    ib += (int)G.steps.size();
    rb += (int)G.ranges.size();
    all.insert(all.end(), G.ranges.begin(), G.ranges.end());
  }
  if (!all.empty()) {
    PQ_CUDA(cudaMalloc(&p->d_ranges, sizeof(ChainRange) * all.size()));
    upload_now(p->d_ranges, all.data(), sizeof(ChainRange) * all.size());
  }
  if (getenv("PQ_B200_DEBUG")) {
    size_t members = 0, nch = 0, ng = 0;
    for (auto& G : p->groups)
      if (!G.steps.empty()) {
        members += G.steps.size();
        nch += G.ranges.size();
        ++ng;
      }
    fprintf(stderr, "[pq_b200] chains: %zu tiny steps in %zu chains, %zu group launches\n", members,
            nch, ng);
  }
}

// device descriptors of every group for one (lane, graph): operand addresses resolved
static void upload_chain_items(pq_handle* h, pq_program* p, int slot, Where w) {
  LaneMem& m = p->lanes[w.lane];
  if (m.chain_items[slot]) return;
  std::vector<ChainItem> items;
  for (auto& G : p->groups)
    for (int j : G.steps) {
      ChainItem it;
      const Step& s = p->steps[j];
      chain_item_from_plan(s.cp, it);
      it.A = resolve(p, s.a, w);
      it.B = resolve(p, s.b, w);
      it.C = resolve(p, s.c, w);
      items.push_back(it);
    }
  if (items.empty()) return;
  PQ_CUDA(cudaMalloc(&m.chain_items[slot], sizeof(ChainItem) * items.size()));
  upload_now(m.chain_items[slot], items.data(), sizeof(ChainItem) * items.size());
}

// Issues every step during stream capture, spreading independent steps over several
// capture streams joined by events, so the instantiated graph is a DAG rather than a
// chain: the ~10^3 tiny world-line contractions of a slice run concurrently.
static void issue_steps_dag(pq_handle* h, pq_program* p, Launch& L0, int phase, Where w, int slot) {
  const int n = (int)p->steps.size();
  const int S = pq_program::NSTREAMS;
  cudaStream_t streams[pq_program::NSTREAMS + 1];
  streams[0] = L0.stream;
  for (int s = 0; s < S; ++s) streams[s + 1] = p->side[s];
  const int NS = S + 1;
  PQ_CUDA(cudaEventRecord(p->fork_ev, streams[0]));
  for (int s = 1; s < NS; ++s) PQ_CUDA(cudaStreamWaitEvent(streams[s], p->fork_ev, 0));
  // issue order: nodes (a step, or a whole group of chains) get increasing sequence numbers;
  // events recorded on a stream cover everything issued on it up to that sequence number
  std::vector<long long> seq(n, -1);
  std::vector<int> where(n, 0);
  std::vector<long long> tail(NS, -1);   // sequence number of the last node on each stream
  std::vector<std::vector<long long>> synced(NS, std::vector<long long>(NS, -1));
  long long counter = 0;
  const ChainItem* items = static_cast<const ChainItem*>(p->lanes[w.lane].chain_items[slot]);
  for (int j = 0; j < n; ++j) {
    if (phase >= 0 && (p->steps[j].dep ? 1 : 0) != phase) continue;
    const int g = p->group_of[j];
    if (g >= 0 && j != p->groups[g].first) continue;   // issued with its group
    // hazards towards the other phase are ordered by the graph launches themselves
    std::vector<int> deps;
    for (int d : (g >= 0 ? p->groups[g].deps : p->deps[j]))
      if (phase < 0 || (p->steps[d].dep ? 1 : 0) == phase) deps.push_back(d);
    int best = -1;
    for (int d : deps) {
      const int sd = where[d];
      if (seq[d] == tail[sd] && (best < 0 || tail[sd] > tail[best])) best = sd;
    }
    if (best < 0) {
      best = 0;
      for (int s = 1; s < NS; ++s)
        if (tail[s] < tail[best]) best = s;
    }
    for (int d : deps) {
      const int sd = where[d];
      if (sd == best || synced[best][sd] >= seq[d]) continue;
      const int gd = p->group_of[d];   // a group records one event, on its first member
      PQ_CUDA(cudaStreamWaitEvent(streams[best], p->step_ev[gd >= 0 ? p->groups[gd].first : d], 0));
      synced[best][sd] = seq[d];
    }
    Launch L = L0;
    L.stream = streams[best];
    const long long me = counter++;
    if (g >= 0) {
      const pq_program::ChainGroup& G = p->groups[g];
      double gbytes = 0, gflops = 0;
      for (int m : G.steps) {
        const ContractPlan& c = p->steps[m].cp;
        gbytes += double(c.M * c.K + c.N * c.K + c.M * c.N) * h->elem_size;
        gflops += 8.0 * double(c.M) * double(c.N) * double(c.K);
      }
      run_chains(L, items + G.item_base, p->d_ranges + G.range_base, (int)G.ranges.size(), gbytes,
                 gflops);
      // one event for the group: every member maps to it
      PQ_CUDA(cudaEventRecord(p->step_ev[G.first], streams[best]));
      for (int m : G.steps) {
        seq[m] = me;
        where[m] = best;
      }
    } else {
      issue_step(h, p, p->steps[j], L, w);
      PQ_CUDA(cudaEventRecord(p->step_ev[j], streams[best]));
      seq[j] = me;
      where[j] = best;
    }
    tail[best] = me;
  }
  for (int s = 1; s < NS; ++s) {  // join every side stream back into the origin
    PQ_CUDA(cudaEventRecord(p->fork_ev, streams[s]));
    PQ_CUDA(cudaStreamWaitEvent(streams[0], p->fork_ev, 0));
  }
}

static void issue_step(pq_handle* h, pq_program* p, Step& s, Launch& L, Where w) {
  switch (s.kind) {
    case ST_CONTRACT:
      run_contract(L, s.cp, resolve(p, s.a, w), resolve(p, s.b, w), resolve(p, s.c, w),
                   resolve(p, s.ta, w), resolve(p, s.tb, w), resolve(p, s.ws, w));
      break;
    case ST_PERMUTE:
      run_permute(L, s.pp, resolve(p, s.a, w), resolve(p, s.c, w));
      break;
    case ST_VIEW: {
      int32_t* starts = p->lanes[w.lane].d_starts;
      run_view(L, resolve(p, s.a, w), resolve(p, s.c, w), s.inner, s.ext, s.nsel, s.outer, s.start0,
               starts ? starts + s.view_slot : nullptr);
      break;
    }
    case ST_SAVE: {
      size_t bytes = size_t(prod(s.dims)) * h->elem_size;
      L.begin(KC_COPY, 2.0 * bytes, 0);
      PQ_CUDA(cudaMemcpyAsync(p->lanes[w.lane].outs[s.save_slot]->ptr, resolve(p, s.a, w), bytes,
                              cudaMemcpyDeviceToDevice, L.stream));
      L.end();
      break;
    }
  }
}

// Kernel nodes with small grids get the highest launch priority, nodes with large grids
// (GEMM-shaped steps, state-sized permutes) the lowest.  The block scheduler otherwise
// drains a large grid completely before it dispatches the first CTA of a later kernel, so
// the latency-bound chains of tiny contractions -- of this slice or of another slice in
// flight on a different lane -- would sit behind every GEMM instead of slipping into the
// SM slots its retiring CTAs free.
static void prioritise_small_nodes(cudaGraph_t graph, int num_sms) {
  int least = 0, greatest = 0;
  if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess || least == greatest) return;
  size_t n = 0;
  if (cudaGraphGetNodes(graph, nullptr, &n) != cudaSuccess || n == 0) return;
  std::vector<cudaGraphNode_t> nodes(n);
  if (cudaGraphGetNodes(graph, nodes.data(), &n) != cudaSuccess) return;
  for (cudaGraphNode_t nd : nodes) {
    cudaGraphNodeType t;
    if (cudaGraphNodeGetType(nd, &t) != cudaSuccess || t != cudaGraphNodeTypeKernel) continue;
    cudaKernelNodeParams kp;
    if (cudaGraphKernelNodeGetParams(nd, &kp) != cudaSuccess) continue;
    const long long ctas = (long long)kp.gridDim.x * kp.gridDim.y * kp.gridDim.z;
    cudaLaunchAttributeValue v;
    memset(&v, 0, sizeof(v));
    v.priority = ctas > 2LL * num_sms ? least : greatest;
    cudaGraphKernelNodeSetAttribute(nd, cudaLaunchAttributePriority, &v);
  }
  cudaGetLastError();
}

// Executes one phase of the program on the handle's stream: eagerly (profiling / option
// graph=1) or by launching its CUDA graph, captured on first use.
// A profiling instance of one graph: the same capture with event-record nodes around every
// kernel node (pq_program_profile_slices).
struct ProfCapture {
  cudaGraphExec_t exec = nullptr;
  std::vector<ProfRecord> recs;
};

static void run_phase(pq_handle* h, pq_program* p, Launch& L, int phase, int lane = 0,
                      cudaStream_t on = nullptr, ProfCapture* pc = nullptr) {
  // phase -1 = the whole stream as ONE graph (no hoisting): invariant and dependent chains
  // then share the parallel branches
  const int slot = phase < 0 ? 2 : phase;
  const int64_t nl = phase < 0 ? p->launches : p->launches2[phase];
  if (nl == 0) return;
  Where w;
  w.lane = lane;
  w.split = phase >= 0;
  LaneMem& m = p->lanes[lane];
  const bool eager = h->profile || h->opt.graph == 1;
  if (eager) {
    issue_steps(h, p, L, phase, w);
  } else {
    cudaGraphExec_t& exec = pc ? pc->exec : m.exec2[slot];
    if (!exec) {
      // graphs are captured on the handle's stream (nothing executes) and may be launched
      // on any stream
      Launch LC = L;
      LC.stream = h->stream;
      int64_t captured = 0;
      LC.launch_counter = &captured;
      LC.profile = false;
      if (pc) {   // event-record nodes around every kernel node
        LC.profile = true;
        LC.external = true;
        LC.prof = &pc->recs;
      }
      if (h->opt.graph != 2 && !m.chain_items[slot]) upload_chain_items(h, p, slot, w);
      cudaGraph_t graph = nullptr;
      PQ_CUDA(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
      try {
        if (h->opt.graph == 2)
          issue_steps(h, p, LC, phase, w);       // single-stream chain (A/B against the DAG)
        else
          issue_steps_dag(h, p, LC, phase, w, slot);   // independent steps on parallel branches
      } catch (...) {
        cudaStreamEndCapture(h->stream, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
      }
      PQ_CUDA(cudaStreamEndCapture(h->stream, &graph));
      if (h->opt.prio == 0) prioritise_small_nodes(graph, h->num_sms);
      cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
      cudaGraphDestroy(graph);
      PQ_CUDA(e);
      m.graph_launches[slot] = captured;
    }
    PQ_CUDA(cudaGraphLaunch(exec, on ? on : h->stream));
    h->launches += m.graph_launches[slot];
  }
  h->n_contract += phase < 0 ? p->ncontract : p->ncontract2[phase];
  h->macs += phase < 0 ? p->macs : p->macs2[phase];
}

// Allocates the memory of lanes [1, n) (lane 0 is allocated by pq_program_compile).
static void ensure_lanes(pq_handle* h, pq_program* p, int n) {
  while ((int)p->lanes.size() < n) {
    LaneMem m;
    for (int a = 0; a < 2; ++a)
      PQ_CUDA(cudaMalloc(&m.arena2[a], p->arena_bytes2[a] ? p->arena_bytes2[a] : ALIGN));
    PQ_CUDA(cudaMalloc(&m.arena_small, p->arena_small_bytes ? p->arena_small_bytes : ALIGN));
    if (p->nviews > 0) {
      PQ_CUDA(cudaMalloc(&m.d_starts, sizeof(int32_t) * p->nviews));
      upload_now(m.d_starts, p->default_starts.data(), sizeof(int32_t) * p->nviews);
    }
    for (const Step& s : p->steps)
      if (s.kind == ST_SAVE)
        m.outs.push_back(std::make_shared<Buffer>(size_t(prod(s.dims)) * h->elem_size, h->stream));
    p->lanes.push_back(std::move(m));
  }
  for (int l = 0; l < n; ++l) {
    LaneMem& m = p->lanes[l];
    if (m.stream) continue;
    PQ_CUDA(cudaStreamCreateWithFlags(&m.stream, cudaStreamNonBlocking));
    PQ_CUDA(cudaEventCreateWithFlags(&m.done_ev, cudaEventDisableTiming));
    PQ_CUDA(cudaEventCreateWithFlags(&m.acc_ev, cudaEventDisableTiming));
  }
  if (!p->batch_ev) {
    PQ_CUDA(cudaEventCreateWithFlags(&p->batch_ev, cudaEventDisableTiming));
    PQ_CUDA(cudaEventCreateWithFlags(&p->table_ev, cudaEventDisableTiming));
  }
}

static void free_lanes(pq_program* p) {
  for (LaneMem& m : p->lanes) {
    for (int a = 0; a < 3; ++a) {
      if (m.exec2[a]) cudaGraphExecDestroy(m.exec2[a]);
      if (m.chain_items[a]) cudaFree(m.chain_items[a]);
    }
    for (int a = 0; a < 2; ++a)
      if (m.arena2[a]) cudaFree(m.arena2[a]);
    if (m.arena_small) cudaFree(m.arena_small);
    if (m.d_starts) cudaFree(m.d_starts);
    if (m.stream) cudaStreamDestroy(m.stream);
    if (m.done_ev) cudaEventDestroy(m.done_ev);
    if (m.acc_ev) cudaEventDestroy(m.acc_ev);
  }
  p->lanes.clear();
}

extern "C" int pq_program_compile(pq_handle* h, const char* tl_text, pq_program** out) {
  if (!h || !tl_text || !out) return PQ_ERR_INVALID;
  *out = nullptr;
  pq_program* p = new pq_program();
  p->lanes.emplace_back();
  try {
    PQ_CUDA(cudaSetDevice(h->device));
    p->device = h->device;
    std::map<std::string, Sym> syms;
    ArenaSim sims[2];  // [0] slice-invariant, [1] slice-dependent
    const int es = h->elem_size;
    size_t small_top = 0;
    auto place = [&](Sym& s) {  // assigns an arena slot to a freshly created tensor
      if (s.bytes <= SMALL_TENSOR_BYTES) {
        s.small = true;
        s.offset = small_top;
        small_top += round_up(s.bytes ? s.bytes : 1);
      } else {
        s.small = false;
        s.offset = sims[s.dep ? 1 : 0].alloc(s.bytes);
      }
    };
    auto ref_of = [&](const Sym& s) {
      Ref r;
      r.null = false;
      r.leaf = s.leaf;
      r.leafptr = s.leaf ? s.leafbuf->ptr : nullptr;
      r.offset = s.offset;
      r.small = s.small;
      r.dep = s.dep;
      return r;
    };
    auto temp_ref = [&](size_t off, bool dep) {
      Ref r;
      r.null = false;
      r.offset = off;
      r.dep = dep;
      return r;
    };
    auto lookup = [&](const std::string& name) -> Sym& {
      auto it = syms.find(name);
      if (it == syms.end())
        throw Error(PQ_ERR_NOT_FOUND, "KeyError: tensor '" + name + "' is not defined in the program");
      return it->second;
    };
    auto drop = [&](const std::string& name) {
      auto it = syms.find(name);
      if (it == syms.end()) return;
      // an invariant tensor that a slice-dependent step reads must survive every replay
      const Sym& y = it->second;
      if (!y.leaf && !y.small && !(!y.dep && y.used_by_dep))
        sims[y.dep ? 1 : 0].release(y.offset, y.bytes);
      syms.erase(it);
    };

    std::stringstream text(tl_text);
    std::string line;
    int lineno = 0;
    while (std::getline(text, line)) {
      ++lineno;
      std::stringstream ls(line);
      std::vector<std::string> tok;
      std::string w;
      while (ls >> w) tok.push_back(w);
      if (tok.empty()) continue;
      const std::string& cmd = tok[0];
      auto need = [&](size_t n) {
        if (tok.size() < n)
          throw Error(PQ_ERR_PARSE, "line " + std::to_string(lineno) + ": too few arguments for '" +
                                        cmd + "'");
      };
      if (cmd == "tensor") {
        need(3);
        Tensor& t = h->get(tok[2]);
        Sym s;
        s.leaf = true;
        s.leafbuf = t.buf;
        s.dims = t.dims;
        s.bytes = t.buf->bytes;
        t.buf->pins += 1;
        p->leaves.emplace_back(tok[2], t.buf);
        p->leaf_gen.push_back(t.buf->gen);
        p->leaf_dims.push_back(t.dims);
        drop(tok[1]);
        syms[tok[1]] = s;
      } else if (cmd == "del") {
        need(2);
        drop(tok[1]);
      } else if (cmd == "ncon") {
        need(4);
        size_t pos = 1;
        std::string C = tok[pos++], A = tok[pos++];
        Sym sa = lookup(A);
        std::vector<int32_t> ai, bi;
        if (!sa.dims.empty()) {
          need(pos + 1);
          ai = parse_ints(tok[pos++]);
        }
        need(pos + 1);
        std::string B = tok[pos++];
        Sym sb = lookup(B);
        if (!sb.dims.empty()) {
          need(pos + 1);
          bi = parse_ints(tok[pos++]);
        }
        Step st;
        st.kind = ST_CONTRACT;
        st.dep = sa.dep || sb.dep;
        if (st.dep) {
          if (!sa.dep) lookup(A).used_by_dep = true;
          if (!sb.dep) lookup(B).used_by_dep = true;
        }
        st.cp = lower_contract(sa.dims, ai, sb.dims, bi, es, h->opt);
        st.a = ref_of(sa);
        st.b = ref_of(sb);
        Sym sc;
        sc.dep = st.dep;
        sc.dims = st.cp.cdims;
        sc.bytes = size_t(st.cp.M * st.cp.N) * es;
        place(sc);
        st.c = ref_of(sc);
        ArenaSim& sim = sims[st.dep ? 1 : 0];
        size_t oa = 0, ob = 0, ow = 0;
        if (st.cp.tempA_bytes) st.ta = temp_ref(oa = sim.alloc(st.cp.tempA_bytes), st.dep);
        if (st.cp.tempB_bytes) st.tb = temp_ref(ob = sim.alloc(st.cp.tempB_bytes), st.dep);
        if (st.cp.ws_bytes) st.ws = temp_ref(ow = sim.alloc(st.cp.ws_bytes), st.dep);
        if (st.cp.tempA_bytes) sim.release(oa, st.cp.tempA_bytes);
        if (st.cp.tempB_bytes) sim.release(ob, st.cp.tempB_bytes);
        if (st.cp.ws_bytes) sim.release(ow, st.cp.ws_bytes);
        p->macs += st.cp.M * st.cp.N * st.cp.K;
        p->macs2[st.dep ? 1 : 0] += st.cp.M * st.cp.N * st.cp.K;
        p->ncontract2[st.dep ? 1 : 0] += 1;
        p->ncontract += 1;
        if (st.cp.M * st.cp.N > p->max_elems) p->max_elems = st.cp.M * st.cp.N;
        p->steps.push_back(std::move(st));
        drop(C);
        syms[C] = sc;  // the DSL's own `del A` / `del B` lines release the operands
      } else if (cmd == "permute") {
        need(3);
        Sym& s = lookup(tok[1]);
        std::vector<int32_t> axes = parse_ints(tok[2]);
        PQ_REQUIRE(axes.size() == s.dims.size(), PQ_ERR_INVALID, "permute: axes length != rank");
        std::vector<int> perm(axes.size());
        for (size_t k = 0; k < axes.size(); ++k) perm[k] = axes[k] - 1;
        PermutePlan pp = lower_permute(s.dims, perm, es, h->opt);
        std::vector<int64_t> nd(axes.size());
        for (size_t k = 0; k < axes.size(); ++k) nd[k] = s.dims[perm[k]];
        if (!pp.identity) {
          Step st;
          st.kind = ST_PERMUTE;
          st.dep = s.dep;
          st.pp = pp;
          st.a = ref_of(s);
          Sym ns;
          ns.dep = s.dep;
          ns.dims = nd;
          ns.bytes = size_t(pp.total) * es;
          place(ns);
          st.c = ref_of(ns);
          p->steps.push_back(std::move(st));
          std::string name = tok[1];
          drop(name);
          syms[name] = ns;
        } else {
          s.dims = nd;
        }
      } else if (cmd == "reshape") {
        need(3);
        Sym& s = lookup(tok[1]);
        std::vector<int64_t> nd;
        std::stringstream gs(tok[2]);
        std::string g;
        while (std::getline(gs, g, ';')) {
          int64_t d = 1;
          for (int32_t ax : parse_ints(g)) {
            PQ_REQUIRE(ax >= 1 && ax <= (int)s.dims.size(), PQ_ERR_INVALID, "reshape: axis out of range");
            d *= s.dims[ax - 1];
          }
          nd.push_back(d);
        }
        PQ_REQUIRE(prod(nd) == prod(s.dims), PQ_ERR_SHAPE, "reshape changes the number of elements");
        s.dims = nd;
      } else if (cmd == "view") {
        need(5);
        Sym src = lookup(tok[2]);
        int axis = std::stoi(tok[3]);
        std::vector<int32_t> idx = parse_ints(tok[4]);
        PQ_REQUIRE(axis >= 1 && axis <= (int)src.dims.size(), PQ_ERR_INVALID, "view: axis out of range");
        PQ_REQUIRE(!idx.empty(), PQ_ERR_INVALID, "view: empty index list");
        for (size_t j = 1; j < idx.size(); ++j)
          PQ_REQUIRE(idx[j] == idx[j - 1] + 1, PQ_ERR_UNSUPPORTED,
                     "view: programs support contiguous index ranges only");
        const int64_t ext = src.dims[axis - 1];
        PQ_REQUIRE(idx.front() >= 1 && idx.back() <= ext, PQ_ERR_INVALID, "view: index out of range");
        Step st;
        st.kind = ST_VIEW;
        st.dep = true;  // its start index is a per-slice parameter
        if (!src.dep) lookup(tok[2]).used_by_dep = true;
        for (int d = 0; d < axis - 1; ++d) st.inner *= src.dims[d];
        for (size_t d = axis; d < src.dims.size(); ++d) st.outer *= src.dims[d];
        st.ext = ext;
        st.nsel = (int64_t)idx.size();
        st.start0 = idx.front();
        st.view_slot = p->nviews++;
        p->default_starts.push_back(idx.front());
        st.a = ref_of(src);
        Sym v;
        v.dep = true;
        v.dims = src.dims;
        v.dims[axis - 1] = st.nsel;
        v.bytes = size_t(prod(v.dims)) * es;
        place(v);
        st.c = ref_of(v);
        p->steps.push_back(std::move(st));
        drop(tok[1]);
        syms[tok[1]] = v;
      } else if (cmd == "save") {
        need(4);
        Sym& s = lookup(tok[1]);
        Step st;
        st.kind = ST_SAVE;
        st.dep = s.dep;
        st.a = ref_of(s);
        st.key = tok[3];
        st.dims = s.dims;
        st.save_slot = p->nsaves++;
        p->lanes[0].outs.push_back(std::make_shared<Buffer>(size_t(prod(s.dims)) * es, h->stream));
        p->steps.push_back(std::move(st));
      } else if (cmd == "decompose") {
        throw Error(PQ_ERR_UNSUPPORTED, "decompose (SVD) is outside the contraction hot path");
      }
      // unknown commands are ignored, like the reference interpreter does
    }

    p->arena_small_bytes = small_top;
    LaneMem& m0 = p->lanes[0];
    for (int a = 0; a < 2; ++a) {
      p->arena_bytes2[a] = sims[a].top;
      PQ_CUDA(cudaMalloc(&m0.arena2[a], sims[a].top ? sims[a].top : ALIGN));
    }
    p->arena_bytes = sims[0].top + sims[1].top;
    PQ_CUDA(cudaMalloc(&m0.arena_small, small_top ? small_top : ALIGN));
    if (p->nviews > 0) {
      PQ_CUDA(cudaMalloc(&m0.d_starts, sizeof(int32_t) * p->nviews));
      upload_now(m0.d_starts, p->default_starts.data(), sizeof(int32_t) * p->nviews);
      PQ_CUDA(cudaMallocHost(&p->h_ring, sizeof(int32_t) * p->nviews * pq_program::RING));
      for (int i = 0; i < pq_program::RING; ++i) {
        PQ_CUDA(cudaEventCreateWithFlags(&p->ring_ev[i], cudaEventDisableTiming));
        p->ring_used[i] = false;
      }
    }
    build_dag(h, p);
    build_chains(h, p);
    for (int i = 0; i < pq_program::NSTREAMS; ++i)
      PQ_CUDA(cudaStreamCreateWithFlags(&p->side[i], cudaStreamNonBlocking));
    PQ_CUDA(cudaEventCreateWithFlags(&p->fork_ev, cudaEventDisableTiming));
    p->step_ev.resize(p->steps.size());
    for (auto& e : p->step_ev) PQ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    // count launches with a dry bookkeeping pass (no device work)
    {
      for (const Step& s : p->steps) {
        int64_t n = 1;
        if (s.kind == ST_CONTRACT) {
          if (s.cp.kind == CK_GEMM && !s.cp.fused_gemm)
            n = 1 + (s.cp.permA.identity ? 0 : 1) + (s.cp.permB.identity ? 0 : 1);
          else if (s.cp.kind == CK_DOT)
            n = 2;
        }
        p->launches2[s.dep ? 1 : 0] += n;
      }
      p->launches = p->launches2[0] + p->launches2[1];
      // what one replay launches in graph mode: a group of chains is a single launch
      p->launches_batched = p->launches;
      if (h->opt.graph == 0 || h->opt.graph > 2)
        for (const auto& G : p->groups)
          if (!G.steps.empty()) p->launches_batched -= (int64_t)G.steps.size() - 1;
    }
  } catch (const Error& e) {
    h->last_error = e.what();
    free_lanes(p);
    if (p->d_ranges) cudaFree(p->d_ranges);
    for (auto& l : p->leaves) l.second->pins -= 1;
    delete p;
    return e.code;
  } catch (const std::exception& e) {
    h->last_error = e.what();
    free_lanes(p);
    if (p->d_ranges) cudaFree(p->d_ranges);
    for (auto& l : p->leaves) l.second->pins -= 1;
    delete p;
    return PQ_ERR_PARSE;
  }
  *out = p;
  return PQ_OK;
}

extern "C" int pq_program_num_views(const pq_program* p) { return p ? p->nviews : PQ_ERR_INVALID; }

extern "C" int pq_program_stats(const pq_program* p, int64_t* arena_bytes, int64_t* launches,
                                int64_t* macs) {
  if (!p) return PQ_ERR_INVALID;
  if (arena_bytes) *arena_bytes = (int64_t)(p->arena_bytes + p->arena_small_bytes);
  if (launches) *launches = p->launches_batched;
  if (macs) *macs = p->macs;
  return PQ_OK;
}

extern "C" int pq_program_set_hoist(pq_program* p, int on) {
  if (!p) return PQ_ERR_INVALID;
  p->hoist = on != 0;
  p->prepared = false;
  return PQ_OK;
}

extern "C" int pq_program_hoist_stats(const pq_program* p, int64_t* macs_invariant,
                                      int64_t* macs_dependent, int64_t* launches_invariant,
                                      int64_t* launches_dependent) {
  if (!p) return PQ_ERR_INVALID;
  if (macs_invariant) *macs_invariant = p->macs2[0];
  if (macs_dependent) *macs_dependent = p->macs2[1];
  if (launches_invariant) *launches_invariant = p->launches2[0];
  if (launches_dependent) *launches_dependent = p->launches2[1];
  return PQ_OK;
}

static void check_leaves(pq_handle* h, pq_program* p) {
  PQ_REQUIRE(p->device == h->device, PQ_ERR_INVALID, "program belongs to another device");
  for (size_t i = 0; i < p->leaves.size(); ++i) {
    auto& l = p->leaves[i];
    auto it = h->tensors.find(l.first);
    PQ_REQUIRE(it != h->tensors.end() && it->second.buf.get() == l.second.get(), PQ_ERR_INVALID,
               "tensor '" + l.first + "' was deleted or rebound since the program was compiled; "
               "recompile the program");
    PQ_REQUIRE(it->second.dims == p->leaf_dims[i], PQ_ERR_SHAPE,
               "tensor '" + l.first + "' changed shape since the program was compiled; "
               "recompile the program");
    if (l.second->gen != p->leaf_gen[i]) {  // re-saved in place: hoisted tensors are stale
      p->leaf_gen[i] = l.second->gen;
      p->prepared = false;
    }
  }
}

// Per-slice view parameters come from the caller: the same bounds the compile-time defaults
// and pq_view get (the reference raises BoundsError, src/layer1.jl:191-194), checked on the
// host before anything is uploaded -- k_view itself does not check.
static void check_starts(const pq_program* p, const int32_t* view_starts, int nslices) {
  if (!view_starts || p->nviews == 0) return;
  for (const Step& s : p->steps) {
    if (s.kind != ST_VIEW) continue;
    for (int i = 0; i < nslices; ++i) {
      const int64_t st = view_starts[size_t(i) * p->nviews + s.view_slot];
      PQ_REQUIRE(st >= 1 && st + s.nsel - 1 <= s.ext, PQ_ERR_INVALID,
                 "view start " + std::to_string(st) + " (slice " + std::to_string(i) + ", view " +
                     std::to_string(s.view_slot) + ") outside [1, " +
                     std::to_string(s.ext - s.nsel + 1) + "]");
    }
  }
}

extern "C" int pq_program_prepare(pq_handle* h, pq_program* p) {
  if (!h || !p) return PQ_ERR_INVALID;
  try {
    PQ_CUDA(cudaSetDevice(h->device));
    check_leaves(h, p);
    Launch L = h->launch_ctx();
    run_phase(h, p, L, 0);
    p->prepared = p->hoist;
  } catch (const Error& e) {
    h->last_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    h->last_error = e.what();
    return PQ_ERR_INVALID;
  }
  return PQ_OK;
}

// Publishes the `save` results of one run of `lane` in the handle's store and, optionally,
// adds them to `accumulate_into` -- on the handle's stream.
static void publish_saves(pq_handle* h, pq_program* p, Launch& L, int lane,
                          const char* accumulate_into, bool split = false) {
  PQ_REQUIRE(!accumulate_into || p->nsaves <= 1, PQ_ERR_UNSUPPORTED,
             "accumulate_into needs a program with a single `save` (it has " +
                 std::to_string(p->nsaves) + ")");
  for (Step& s : p->steps) {
    if (s.kind != ST_SAVE) continue;
    // with hoisting a slice-invariant `save` ran once, in phase 0, into lane 0's buffer
    const std::shared_ptr<Buffer>& out = p->lanes[(split && !s.dep) ? 0 : lane].outs[s.save_slot];
    Tensor t;
    t.dims = s.dims;
    t.buf = out;
    h->tensors[s.key] = t;
    if (!accumulate_into) continue;
    auto it = h->tensors.find(accumulate_into);
    int64_t n = prod(s.dims);
    if (it == h->tensors.end()) {
      Tensor d;
      d.dims = s.dims;
      d.buf = std::make_shared<Buffer>(size_t(n) * h->elem_size, h->stream);
      L.begin(KC_COPY, 2.0 * n * h->elem_size, 0);
      PQ_CUDA(cudaMemcpyAsync(d.buf->ptr, out->ptr, size_t(n) * h->elem_size,
                              cudaMemcpyDeviceToDevice, h->stream));
      L.end();
      h->tensors[accumulate_into] = d;
    } else {
      PQ_REQUIRE(it->second.numel() == n, PQ_ERR_SHAPE, "accumulate: size mismatch");
      run_accumulate(L, it->second.buf->ptr, out->ptr, n);
    }
  }
}

// one slice on lane 0, everything on the handle's stream
static void run_one(pq_handle* h, pq_program* p, const int32_t* view_starts,
                    const char* accumulate_into) {
  if (view_starts && p->nviews > 0) {
    int slot = p->ring_pos;
    p->ring_pos = (p->ring_pos + 1) % pq_program::RING;
    if (p->ring_used[slot]) PQ_CUDA(cudaEventSynchronize(p->ring_ev[slot]));
    int32_t* src = p->h_ring + size_t(slot) * p->nviews;
    memcpy(src, view_starts, sizeof(int32_t) * p->nviews);
    PQ_CUDA(cudaMemcpyAsync(p->lanes[0].d_starts, src, sizeof(int32_t) * p->nviews,
                            cudaMemcpyHostToDevice, h->stream));
    PQ_CUDA(cudaEventRecord(p->ring_ev[slot], h->stream));
    p->ring_used[slot] = true;
  }
  Launch L = h->launch_ctx();
  // hoisting: the slice-invariant part runs once (pq_program_prepare), only the
  // slice-dependent part is replayed per slice
  if (!p->hoist) {
    run_phase(h, p, L, -1);
  } else {
    if (!p->prepared) {
      run_phase(h, p, L, 0);
      p->prepared = true;
    }
    run_phase(h, p, L, 1);
  }
  h->note_tensor(p->max_elems);
  publish_saves(h, p, L, 0, accumulate_into);
}

extern "C" int pq_program_run(pq_handle* h, pq_program* p, const int32_t* view_starts, int nviews,
                              const char* accumulate_into) {
  if (!h || !p) return PQ_ERR_INVALID;
  try {
    PQ_CUDA(cudaSetDevice(h->device));
    check_leaves(h, p);
    if (view_starts)
      PQ_REQUIRE(nviews == p->nviews, PQ_ERR_INVALID, "pq_program_run: wrong number of view starts");
    check_starts(p, view_starts, 1);
    run_one(h, p, view_starts, accumulate_into);
  } catch (const Error& e) {
    h->last_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    h->last_error = e.what();
    return PQ_ERR_INVALID;
  }
  return PQ_OK;
}

// The slice loop of a sliced contraction in one call: for s in [0, nslices) run the program
// with view_starts[s * nviews ...] and add its saved result to `accumulate_into`, in slice
// order (so the sum is bit-identical to nslices pq_program_run calls).  Slices are
// independent, so up to `nlanes` of them are kept in flight: each lane has private arenas,
// view parameters and graph instance and runs on its own stream, which hides the
// latency-bound stretch of one slice (its ~10^3 tiny contractions) behind the GEMM steps of
// the others.  Only the accumulations are ordered, on the handle's stream.
extern "C" int pq_program_run_slices(pq_handle* h, pq_program* p, const int32_t* view_starts,
                                     int nslices, int nviews, const char* accumulate_into,
                                     int nlanes) {
  if (!h || !p || nslices < 0) return PQ_ERR_INVALID;
  try {
    PQ_CUDA(cudaSetDevice(h->device));
    check_leaves(h, p);
    if (view_starts)
      PQ_REQUIRE(nviews == p->nviews, PQ_ERR_INVALID,
                 "pq_program_run_slices: wrong number of view starts");
    if (nslices == 0) return PQ_OK;
    check_starts(p, view_starts, nslices);
    const bool eager = h->profile || h->opt.graph == 1;
    if (nlanes > pq_program::MAX_LANES) nlanes = pq_program::MAX_LANES;
    if (nlanes > nslices) nlanes = nslices;
    if (eager || nlanes <= 1) {
      for (int s = 0; s < nslices; ++s)
        run_one(h, p, view_starts ? view_starts + size_t(s) * p->nviews : nullptr, accumulate_into);
      return PQ_OK;
    }
    ensure_lanes(h, p, nlanes);
    const bool params = view_starts && p->nviews > 0;
    if (params) {  // one upload for the whole batch
      const size_t n = size_t(nslices) * p->nviews;
      if (p->table_used) PQ_CUDA(cudaEventSynchronize(p->table_ev));
      if (n > p->table_cap) {
        if (p->h_table) PQ_CUDA(cudaFreeHost(p->h_table));
        if (p->d_table) PQ_CUDA(cudaFree(p->d_table));
        p->h_table = nullptr;
        p->d_table = nullptr;
        PQ_CUDA(cudaMallocHost(&p->h_table, sizeof(int32_t) * n));
        PQ_CUDA(cudaMalloc(&p->d_table, sizeof(int32_t) * n));
        p->table_cap = n;
      }
      memcpy(p->h_table, view_starts, sizeof(int32_t) * n);
      PQ_CUDA(cudaMemcpyAsync(p->d_table, p->h_table, sizeof(int32_t) * n, cudaMemcpyHostToDevice,
                              h->stream));
      PQ_CUDA(cudaEventRecord(p->table_ev, h->stream));
      p->table_used = true;
    }
    Launch L = h->launch_ctx();
    const bool split = p->hoist;
    if (split && !p->prepared) {
      run_phase(h, p, L, 0);
      p->prepared = true;
    }
    // fork: the lanes start after everything queued on the handle's stream so far
    PQ_CUDA(cudaEventRecord(p->batch_ev, h->stream));
    for (int l = 0; l < nlanes; ++l)
      PQ_CUDA(cudaStreamWaitEvent(p->lanes[l].stream, p->batch_ev, 0));
    for (int s = 0; s < nslices; ++s) {
      const int lane = s % nlanes;
      LaneMem& m = p->lanes[lane];
      // the lane's `save` buffer must have been consumed by the previous accumulation
      if (s >= nlanes) PQ_CUDA(cudaStreamWaitEvent(m.stream, m.acc_ev, 0));
      if (params)
        PQ_CUDA(cudaMemcpyAsync(m.d_starts, p->d_table + size_t(s) * p->nviews,
                                sizeof(int32_t) * p->nviews, cudaMemcpyDeviceToDevice, m.stream));
      run_phase(h, p, L, split ? 1 : -1, lane, m.stream);
      PQ_CUDA(cudaEventRecord(m.done_ev, m.stream));
      // join: publish / accumulate in slice order on the handle's stream
      PQ_CUDA(cudaStreamWaitEvent(h->stream, m.done_ev, 0));
      publish_saves(h, p, L, lane, accumulate_into, split);
      PQ_CUDA(cudaEventRecord(m.acc_ev, h->stream));
    }
    h->note_tensor(p->max_elems);
  } catch (const Error& e) {
    h->last_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    h->last_error = e.what();
    return PQ_ERR_INVALID;
  }
  return PQ_OK;
}

// Per-class timing measured inside the graph replays (see include/pq_b200.h).  Same launch
// structure as pq_program_run_slices -- lanes on their own streams, one graph per slice with
// its parallel branches -- but every (lane, round) uses a profiling instance of the graph whose
// kernel nodes are bracketed by event-record nodes.  Nothing is published or accumulated.
extern "C" int pq_program_profile_slices(pq_handle* h, pq_program* p, const int32_t* view_starts,
                                         int nslices, int nviews, int nlanes, double* busy_ms,
                                         double* sum_ms, int64_t* launches, double* bytes,
                                         double* flops, double* wall_ms) {
  if (!h || !p || nslices <= 0) return PQ_ERR_INVALID;
  std::vector<std::vector<ProfCapture>> caps;   // [lane][round]
  cudaEvent_t base = nullptr;
  int rc = PQ_OK;
  try {
    PQ_CUDA(cudaSetDevice(h->device));
    check_leaves(h, p);
    PQ_REQUIRE(!(h->profile || h->opt.graph == 1), PQ_ERR_INVALID,
               "pq_program_profile_slices measures graph replays: not available in eager mode");
    if (view_starts)
      PQ_REQUIRE(nviews == p->nviews, PQ_ERR_INVALID,
                 "pq_program_profile_slices: wrong number of view starts");
    check_starts(p, view_starts, nslices);
    if (nlanes < 1) nlanes = 1;
    if (nlanes > pq_program::MAX_LANES) nlanes = pq_program::MAX_LANES;
    if (nlanes > nslices) nlanes = nslices;
    PQ_REQUIRE(nslices <= 16 * nlanes, PQ_ERR_INVALID, "pq_program_profile_slices: too many slices");
    ensure_lanes(h, p, nlanes);
    const int rounds = (nslices + nlanes - 1) / nlanes;
    caps.resize(nlanes);
    for (auto& c : caps) c.resize(rounds);
    const bool params = view_starts && p->nviews > 0;
    int32_t* d_table = nullptr;
    if (params) {
      const size_t n = size_t(nslices) * p->nviews;
      PQ_CUDA(cudaMalloc(&d_table, sizeof(int32_t) * n));
      cudaError_t e = cudaMemcpy(d_table, view_starts, sizeof(int32_t) * n, cudaMemcpyHostToDevice);
      if (e != cudaSuccess) {
        cudaFree(d_table);
        PQ_CUDA(e);
      }
    }
    Launch L = h->launch_ctx();
    int64_t dummy = 0;
    L.launch_counter = &dummy;   // a measurement pass does not count as product launches
    const int64_t nc0 = h->n_contract, macs0 = h->macs, l0 = h->launches;
    const bool split = p->hoist;
    if (split && !p->prepared) {
      run_phase(h, p, L, 0);
      p->prepared = true;
    }
    // pass 0 captures + instantiates every profiling instance (host-bound: the lanes do not
    // overlap yet); pass 1 launches them back to back like the timed run and is the one the
    // events keep
    PQ_CUDA(cudaEventCreate(&base));
    for (int pass = 0; pass < 2; ++pass) {
      PQ_CUDA(cudaStreamSynchronize(h->stream));
      PQ_CUDA(cudaEventRecord(base, h->stream));
      PQ_CUDA(cudaEventRecord(p->batch_ev, h->stream));
      for (int l = 0; l < nlanes; ++l)
        PQ_CUDA(cudaStreamWaitEvent(p->lanes[l].stream, p->batch_ev, 0));
      for (int s = 0; s < nslices; ++s) {
        const int lane = s % nlanes, round = s / nlanes;
        LaneMem& m = p->lanes[lane];
        if (params)
          PQ_CUDA(cudaMemcpyAsync(m.d_starts, d_table + size_t(s) * p->nviews,
                                  sizeof(int32_t) * p->nviews, cudaMemcpyDeviceToDevice, m.stream));
        run_phase(h, p, L, split ? 1 : -1, lane, m.stream, &caps[lane][round]);
      }
      for (int l = 0; l < nlanes; ++l) PQ_CUDA(cudaStreamSynchronize(p->lanes[l].stream));
    }
    if (d_table) cudaFree(d_table);
    h->n_contract = nc0;
    h->macs = macs0;
    h->launches = l0;
    // intervals relative to `base`
    struct Iv {
      float a, b;
    };
    std::vector<std::vector<Iv>> iv(PQ_NUM_KERNEL_CLASSES);
    float first = 1e30f, last = 0;
    for (int c = 0; c < PQ_NUM_KERNEL_CLASSES; ++c) {
      if (busy_ms) busy_ms[c] = 0;
      if (sum_ms) sum_ms[c] = 0;
      if (launches) launches[c] = 0;
      if (bytes) bytes[c] = 0;
      if (flops) flops[c] = 0;
    }
    for (auto& lane : caps)
      for (auto& cap : lane)
        for (auto& r : cap.recs) {
          float a = 0, b = 0;
          PQ_CUDA(cudaEventElapsedTime(&a, base, r.e0));
          PQ_CUDA(cudaEventElapsedTime(&b, base, r.e1));
          iv[r.cls].push_back({a, b});
          if (a < first) first = a;
          if (b > last) last = b;
          if (sum_ms) sum_ms[r.cls] += double(b - a);
          if (launches) launches[r.cls] += 1;
          if (bytes) bytes[r.cls] += r.bytes;
          if (flops) flops[r.cls] += r.flops;
        }
    for (int c = 0; c < PQ_NUM_KERNEL_CLASSES; ++c) {
      auto& v = iv[c];
      std::sort(v.begin(), v.end(), [](const Iv& x, const Iv& y) { return x.a < y.a; });
      double busy = 0;
      float ca = 0, cb = -1;
      for (const Iv& x : v) {
        if (cb < 0 || x.a > cb) {
          if (cb >= 0) busy += double(cb - ca);
          ca = x.a;
          cb = x.b;
        } else if (x.b > cb) {
          cb = x.b;
        }
      }
      if (cb >= 0) busy += double(cb - ca);
      if (busy_ms) busy_ms[c] = busy;
    }
    if (wall_ms) *wall_ms = last > first ? double(last - first) : 0.0;
  } catch (const Error& e) {
    h->last_error = e.what();
    rc = e.code;
  } catch (const std::exception& e) {
    h->last_error = e.what();
    rc = PQ_ERR_INVALID;
  }
  for (auto& lane : caps)
    for (auto& cap : lane) {
      if (cap.exec) cudaGraphExecDestroy(cap.exec);
      for (auto& r : cap.recs) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
      }
    }
  if (base) cudaEventDestroy(base);
  return rc;
}

extern "C" int pq_program_destroy(pq_handle* h, pq_program* p) {
  if (!p) return PQ_OK;
  if (h) {
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
  }
  for (LaneMem& m : p->lanes)
    if (m.stream) cudaStreamSynchronize(m.stream);
  for (auto& e : p->step_ev) cudaEventDestroy(e);
  if (p->fork_ev) cudaEventDestroy(p->fork_ev);
  for (int i = 0; i < pq_program::NSTREAMS; ++i)
    if (p->side[i]) cudaStreamDestroy(p->side[i]);
  for (auto& l : p->leaves) l.second->pins -= 1;
  free_lanes(p);
  if (p->h_ring) {
    cudaFreeHost(p->h_ring);
    for (int i = 0; i < pq_program::RING; ++i) cudaEventDestroy(p->ring_ev[i]);
  }
  if (p->h_table) cudaFreeHost(p->h_table);
  if (p->d_table) cudaFree(p->d_table);
  if (p->d_ranges) cudaFree(p->d_ranges);
  if (p->batch_ev) cudaEventDestroy(p->batch_ev);
  if (p->table_ev) cudaEventDestroy(p->table_ev);
  delete p;
  return PQ_OK;
}
