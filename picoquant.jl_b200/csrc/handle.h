// pq_handle: one GPU, one stream, one label -> device tensor store.
#pragma once
#include "common.h"

namespace pq {

// device allocation, stream-ordered (cudaMallocAsync from the default pool, whose
// release threshold is raised so that freed blocks are recycled without going back to
// the driver: intermediates are allocated and freed on every contraction step).
struct Buffer {
  void* ptr = nullptr;
  size_t bytes = 0;
  cudaStream_t stream = nullptr;
  bool external = false;  // memory owned by somebody else (program arenas)
  int pins = 0;           // compiled programs bound to this buffer (leaf tensors)
  uint64_t gen = 0;       // bumped by every in-place re-save (programs re-run their invariant part)
  Buffer(size_t n, cudaStream_t s);
  Buffer(void* p, size_t n) : ptr(p), bytes(n), external(true) {}
  ~Buffer();
  Buffer(const Buffer&) = delete;
  Buffer& operator=(const Buffer&) = delete;
};

struct Tensor {
  std::shared_ptr<Buffer> buf;  // shared by aliases (save_output)
  std::vector<int64_t> dims;
  int64_t numel() const { return prod(dims); }
};

struct Comm;  // nccl_dyn.cu

}  // namespace pq

struct pq_handle {
  int device = 0;
  int dtype = PQ_C128;
  int elem_size = 16;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  std::map<std::string, pq::Tensor> tensors;
  std::string last_error;
  pq::Options opt;
  // counters
  int64_t n_contract = 0, macs = 0, max_elems = 0, launches = 0;
  // profiling
  bool profile = false;
  std::vector<pq::ProfRecord> prof;
  std::vector<cudaEvent_t> event_pool;
  double prof_ms[PQ_NUM_KERNEL_CLASSES] = {0};
  int64_t prof_launches[PQ_NUM_KERNEL_CLASSES] = {0};
  double prof_bytes[PQ_NUM_KERNEL_CLASSES] = {0};
  double prof_flops[PQ_NUM_KERNEL_CLASSES] = {0};
  pq::Comm* comm = nullptr;
  cudaEvent_t timer0 = nullptr, timer1 = nullptr;
  // pq_save_tensors: one pinned staging block (table + packed data) and its device twin
  unsigned char* stage_host = nullptr;
  unsigned char* stage_dev = nullptr;
  size_t stage_cap = 0;
  cudaEvent_t stage_ev = nullptr;
  // pq_load_tensor of large tensors: two pinned bounce blocks + their copy-done events
  unsigned char* d2h_stage[2] = {nullptr, nullptr};
  cudaEvent_t d2h_ev[2] = {nullptr, nullptr};
  bool stage_busy = false;

  pq::Launch launch_ctx();
  pq::Tensor& get(const std::string& label);
  void note_tensor(int64_t elems) {
    if (elems > max_elems) max_elems = elems;
  }
  void drain_profile();
};

namespace pq {
void comm_destroy(Comm* c);
// kernels_svd.cu: SVD-based split of the m x n matrix in `work` (overwritten); returns chi
int run_decompose(pq_handle* h, const Launch& L, void* work, int64_t m, int64_t n, double threshold,
                  int max_rank, std::shared_ptr<Buffer>& Bout, std::shared_ptr<Buffer>& Cout);
}
