// C ABI of libpq_b200: the device-side twin of PicoQuant's InteractiveBackend
// (reference: src/backends/interactive.jl).  See include/pq_b200.h for the contract
// of every entry point and the reference line it replaces.
#include <cstring>
#include <thread>
#include <complex>

#include "handle.h"

using namespace pq;

// ---------------------------------------------------------------------------
// Buffer / Launch / handle helpers
// ---------------------------------------------------------------------------
namespace pq {

Buffer::Buffer(size_t n, cudaStream_t s) : bytes(n), stream(s) {
  size_t alloc = n == 0 ? 16 : n;
  PQ_CUDA(cudaMallocAsync(&ptr, alloc, s));
}
Buffer::~Buffer() {
  if (ptr && !external) cudaFreeAsync(ptr, stream);
}

void Launch::begin(int cls, double bytes, double flops) const {
  if (launch_counter) ++*launch_counter;
  if (profile && prof) {
    cur.cls = cls;
    cur.bytes = bytes;
    cur.flops = flops;
    cudaEventCreate(&cur.e0);
    cudaEventCreate(&cur.e1);
    if (external)
      cudaEventRecordWithFlags(cur.e0, stream, cudaEventRecordExternal);
    else
      cudaEventRecord(cur.e0, stream);
  }
}
void Launch::end() const {
  if (profile && prof) {
    if (external)
      cudaEventRecordWithFlags(cur.e1, stream, cudaEventRecordExternal);
    else
      cudaEventRecord(cur.e1, stream);
    prof->push_back(cur);
  }
}

}  // namespace pq

Launch pq_handle::launch_ctx() {
  Launch L;
  L.stream = stream;
  L.elem_size = elem_size;
  L.num_sms = num_sms;
  L.launch_counter = &launches;
  L.profile = profile;
  L.prof = &prof;
  L.opt = &opt;
  return L;
}

Tensor& pq_handle::get(const std::string& label) {
  auto it = tensors.find(label);
  if (it == tensors.end()) throw Error(PQ_ERR_NOT_FOUND, "KeyError: tensor '" + label + "' not found");
  return it->second;
}

void pq_handle::drain_profile() {
  if (prof.empty()) return;
  cudaStreamSynchronize(stream);
  for (auto& r : prof) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      prof_ms[r.cls] += ms;
      prof_launches[r.cls] += 1;
      prof_bytes[r.cls] += r.bytes;
      prof_flops[r.cls] += r.flops;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  prof.clear();
}

#define PQ_TRY(h) try {
#define PQ_CATCH(h)                                   \
  }                                                   \
  catch (const Error& e) {                            \
    if (h) (h)->last_error = e.what();                \
    return e.code;                                    \
  }                                                   \
  catch (const std::exception& e) {                   \
    if (h) (h)->last_error = e.what();                \
    return PQ_ERR_INVALID;                            \
  }                                                   \
  return PQ_OK;

static void set_device(pq_handle* h) { PQ_CUDA(cudaSetDevice(h->device)); }

// ---------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------
extern "C" const char* pq_version(void) { return "pq_b200 0.1.0 (sm_100a)"; }

extern "C" int pq_create(int device, int dtype, pq_handle** out) {
  if (!out) return PQ_ERR_INVALID;
  *out = nullptr;
  if (dtype != PQ_C64 && dtype != PQ_C128) return PQ_ERR_INVALID;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count)
    return PQ_ERR_CUDA;  // no CPU fallback: a usable GPU is mandatory
  pq_handle* h = new pq_handle();
  try {
    h->device = device;
    h->dtype = dtype;
    h->elem_size = dtype == PQ_C128 ? 16 : 8;
    PQ_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PQ_CUDA(cudaGetDeviceProperties(&prop, device));
    h->num_sms = prop.multiProcessorCount;
    PQ_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    cudaMemPool_t pool;
    PQ_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t threshold = UINT64_MAX;
    PQ_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
    init_kernels();
    init_kernels_cgemm();
    init_kernels_ozaki_t();
  } catch (const std::exception&) {
    delete h;
    return PQ_ERR_CUDA;
  }
  *out = h;
  return PQ_OK;
}

extern "C" int pq_destroy(pq_handle* h) {
  if (!h) return PQ_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  h->drain_profile();
  h->tensors.clear();
  if (h->comm) comm_destroy(h->comm);
  if (h->timer0) {
    cudaEventDestroy(h->timer0);
    cudaEventDestroy(h->timer1);
  }
  if (h->stage_host) cudaFreeHost(h->stage_host);
  if (h->stage_dev) cudaFree(h->stage_dev);
  if (h->stage_ev) cudaEventDestroy(h->stage_ev);
  for (int i = 0; i < 2; ++i) {
    if (h->d2h_stage[i]) cudaFreeHost(h->d2h_stage[i]);
    if (h->d2h_ev[i]) cudaEventDestroy(h->d2h_ev[i]);
  }
  cudaStreamSynchronize(h->stream);
  cudaStreamDestroy(h->stream);
  delete h;
  return PQ_OK;
}

extern "C" const char* pq_last_error(const pq_handle* h) { return h ? h->last_error.c_str() : ""; }

// ---------------------------------------------------------------------------
// save / info / load
// ---------------------------------------------------------------------------
template <typename Dst>
static void convert_host(const void* src, int src_dtype, Dst* dst_ri, int64_t n) {
  // dst_ri: interleaved (re, im) pairs of type Dst
  switch (src_dtype) {
    case PQ_HOST_F32: {
      const float* s = (const float*)src;
      for (int64_t i = 0; i < n; ++i) { dst_ri[2 * i] = (Dst)s[i]; dst_ri[2 * i + 1] = 0; }
      break;
    }
    case PQ_HOST_F64: {
      const double* s = (const double*)src;
      for (int64_t i = 0; i < n; ++i) { dst_ri[2 * i] = (Dst)s[i]; dst_ri[2 * i + 1] = 0; }
      break;
    }
    case PQ_HOST_C64: {
      const float* s = (const float*)src;
      for (int64_t i = 0; i < 2 * n; ++i) dst_ri[i] = (Dst)s[i];
      break;
    }
    case PQ_HOST_C128: {
      const double* s = (const double*)src;
      for (int64_t i = 0; i < 2 * n; ++i) dst_ri[i] = (Dst)s[i];
      break;
    }
    default:
      throw Error(PQ_ERR_INVALID, "bad host dtype");
  }
}

extern "C" int pq_save_tensor(pq_handle* h, const char* label, int rank, const int64_t* dims,
                              const void* host, int host_dtype) {
  if (!h) return PQ_ERR_INVALID;
  PQ_TRY(h)
  PQ_REQUIRE(label && rank >= 0 && rank <= PQ_MAX_RANK && (rank == 0 || dims) && host,
             PQ_ERR_INVALID, "pq_save_tensor: bad arguments");
  set_device(h);
  Tensor t;
  t.dims.assign(dims, dims + rank);
  for (auto d : t.dims) PQ_REQUIRE(d >= 1, PQ_ERR_INVALID, "extents must be >= 1");
  int64_t n = t.numel();
  // Re-saving a tensor of the same size whose buffer is referenced only by this label
  // (and by compiled programs bound to it) updates the data in place, so programs see
  // the new values -- like execute_dsl_file re-reading the HDF5 file.  Otherwise the
  // label is rebound to a fresh buffer and aliases keep the old array.
  auto old = h->tensors.find(label);
  if (old != h->tensors.end() && old->second.buf && !old->second.buf->external &&
      old->second.buf->bytes == size_t(n) * h->elem_size &&
      old->second.buf.use_count() - old->second.buf->pins == 1) {
    t.buf = old->second.buf;
    t.buf->gen += 1;
  } else
    t.buf = std::make_shared<Buffer>(size_t(n) * h->elem_size, h->stream);
  bool direct = (h->dtype == PQ_C128 && host_dtype == PQ_HOST_C128) ||
                (h->dtype == PQ_C64 && host_dtype == PQ_HOST_C64);
  if (direct) {
    PQ_CUDA(cudaMemcpyAsync(t.buf->ptr, host, size_t(n) * h->elem_size, cudaMemcpyHostToDevice,
                            h->stream));
    // pageable source: the runtime stages it before returning; pinned source: the
    // caller must keep it alive until pq_sync (documented in INTEGRATION.md)
  } else {
    std::vector<unsigned char> tmp(size_t(n) * h->elem_size);
    if (h->dtype == PQ_C128)
      convert_host<double>(host, host_dtype, (double*)tmp.data(), n);
    else
      convert_host<float>(host, host_dtype, (float*)tmp.data(), n);
    // pageable source: the runtime copies it to its staging buffer before returning, so
    // `tmp` may be released right away and the host does not have to wait for the GPU
    PQ_CUDA(cudaMemcpyAsync(t.buf->ptr, tmp.data(), tmp.size(), cudaMemcpyHostToDevice, h->stream));
  }
  h->note_tensor(n);
  h->tensors[label] = std::move(t);
  PQ_CATCH(h)
}

// Batched save_tensor_data: the O(#gates) tiny uploads of a network (interactive.jl:32-36 called
// once per node from src/layer3.jl:195,226,277,308; re-done per rank in
// examples/dist_slicing_example.jl:22-27) as ONE pinned staging block, ONE host->device copy and
// ONE scatter launch.  Same per-tensor semantics as pq_save_tensor (convert to the backend
// dtype, in-place update of a uniquely held buffer of the same size, rebind otherwise).
extern "C" int pq_save_tensors(pq_handle* h, int n, const char* const* labels, const int* ranks,
                               const int64_t* dims_flat, const void* const* hosts,
                               const int* host_dtypes) {
  if (!h) return PQ_ERR_INVALID;
  PQ_TRY(h)
  PQ_REQUIRE(n >= 0 && (n == 0 || (labels && ranks && hosts && host_dtypes)), PQ_ERR_INVALID,
             "pq_save_tensors: bad arguments");
  if (n == 0) return PQ_OK;
  set_device(h);
  // validate everything before touching the store
  std::vector<int64_t> numel(n);
  std::vector<size_t> dim_at(n), off(n);
  size_t pos = 0, bytes = 0;
  const size_t table_bytes = (size_t(n) * sizeof(ScatterItem) + 255) & ~size_t(255);
  for (int i = 0; i < n; ++i) {
    PQ_REQUIRE(labels[i] && hosts[i] && ranks[i] >= 0 && ranks[i] <= PQ_MAX_RANK &&
                   (ranks[i] == 0 || dims_flat),
               PQ_ERR_INVALID, "pq_save_tensors: bad arguments for tensor " + std::to_string(i));
    PQ_REQUIRE(host_dtypes[i] >= PQ_HOST_F32 && host_dtypes[i] <= PQ_HOST_C128, PQ_ERR_INVALID,
               "bad host dtype");
    dim_at[i] = pos;
    int64_t m = 1;
    for (int d = 0; d < ranks[i]; ++d) {
      PQ_REQUIRE(dims_flat[pos + d] >= 1, PQ_ERR_INVALID, "extents must be >= 1");
      m *= dims_flat[pos + d];
    }
    pos += ranks[i];
    numel[i] = m;
    off[i] = table_bytes + bytes;
    bytes += (size_t(m) * h->elem_size + 15) & ~size_t(15);
  }
  const size_t total = table_bytes + bytes;
  if (h->stage_busy) PQ_CUDA(cudaEventSynchronize(h->stage_ev));   // previous batch still copying
  if (total > h->stage_cap) {
    if (h->stage_host) PQ_CUDA(cudaFreeHost(h->stage_host));
    if (h->stage_dev) PQ_CUDA(cudaFree(h->stage_dev));
    h->stage_host = h->stage_dev = nullptr;
    h->stage_cap = 0;
    const size_t cap = total * 2;
    PQ_CUDA(cudaMallocHost(&h->stage_host, cap));
    PQ_CUDA(cudaMalloc(&h->stage_dev, cap));
    h->stage_cap = cap;
  }
  if (!h->stage_ev) PQ_CUDA(cudaEventCreateWithFlags(&h->stage_ev, cudaEventDisableTiming));
  ScatterItem* table = reinterpret_cast<ScatterItem*>(h->stage_host);
  for (int i = 0; i < n; ++i) {
    Tensor t;
    t.dims.assign(dims_flat + dim_at[i], dims_flat + dim_at[i] + ranks[i]);
    const size_t nb = size_t(numel[i]) * h->elem_size;
    auto old = h->tensors.find(labels[i]);
    if (old != h->tensors.end() && old->second.buf && !old->second.buf->external &&
        old->second.buf->bytes == nb && old->second.buf.use_count() - old->second.buf->pins == 1) {
      t.buf = old->second.buf;
      t.buf->gen += 1;
    } else {
      t.buf = std::make_shared<Buffer>(nb, h->stream);
    }
    unsigned char* dst = h->stage_host + off[i];
    const bool direct = (h->dtype == PQ_C128 && host_dtypes[i] == PQ_HOST_C128) ||
                        (h->dtype == PQ_C64 && host_dtypes[i] == PQ_HOST_C64);
    if (direct)
      memcpy(dst, hosts[i], nb);
    else if (h->dtype == PQ_C128)
      convert_host<double>(hosts[i], host_dtypes[i], (double*)dst, numel[i]);
    else
      convert_host<float>(hosts[i], host_dtypes[i], (float*)dst, numel[i]);
    table[i].dst = t.buf->ptr;
    table[i].src_off = off[i];
    table[i].words = nb / 8;
    h->note_tensor(numel[i]);
    h->tensors[labels[i]] = std::move(t);
  }
  PQ_CUDA(cudaMemcpyAsync(h->stage_dev, h->stage_host, total, cudaMemcpyHostToDevice, h->stream));
  PQ_CUDA(cudaEventRecord(h->stage_ev, h->stream));
  h->stage_busy = true;
  Launch L = h->launch_ctx();
  run_scatter(L, h->stage_dev, reinterpret_cast<const ScatterItem*>(h->stage_dev), n, double(bytes));
  PQ_CATCH(h)
}

extern "C" int pq_tensor_info(pq_handle* h, const char* label, int* rank, int64_t* dims) {
  if (!h || !label) return PQ_ERR_INVALID;
  auto it = h->tensors.find(label);
  if (it == h->tensors.end()) return PQ_ERR_NOT_FOUND;
  if (rank) *rank = (int)it->second.dims.size();
  if (dims)
    for (size_t i = 0; i < it->second.dims.size(); ++i) dims[i] = it->second.dims[i];
  return PQ_OK;
}

// Device -> pageable host memory for large tensors (a 2^26-element state vector is 1 GiB).  A
// plain cudaMemcpy into pageable memory goes through the driver's own bounce buffer, one chunk
// at a time, and pays the first-touch page faults of a freshly allocated destination inside the
// copy (measured: ~3.4 GB/s).  Here two pinned blocks alternate: while the DMA engine fills one,
// a few host threads copy the other into the destination.
static void d2h_pipelined(pq_handle* h, unsigned char* dst, const unsigned char* src, size_t bytes) {
  constexpr size_t CHUNK = size_t(32) << 20;
  constexpr int THREADS = 4;
  for (int i = 0; i < 2; ++i) {
    if (!h->d2h_stage[i]) PQ_CUDA(cudaMallocHost(&h->d2h_stage[i], CHUNK));
    if (!h->d2h_ev[i]) PQ_CUDA(cudaEventCreateWithFlags(&h->d2h_ev[i], cudaEventDisableTiming));
  }
  const size_t chunks = (bytes + CHUNK - 1) / CHUNK;
  auto issue = [&](size_t c) {
    const size_t off = c * CHUNK, len = std::min(CHUNK, bytes - off);
    PQ_CUDA(cudaMemcpyAsync(h->d2h_stage[c & 1], src + off, len, cudaMemcpyDeviceToHost, h->stream));
    PQ_CUDA(cudaEventRecord(h->d2h_ev[c & 1], h->stream));
  };
  issue(0);
  for (size_t c = 0; c < chunks; ++c) {
    PQ_CUDA(cudaEventSynchronize(h->d2h_ev[c & 1]));
    if (c + 1 < chunks) issue(c + 1);   // the other block: its previous contents were copied out last turn
    const size_t off = c * CHUNK, len = std::min(CHUNK, bytes - off);
    const unsigned char* s = h->d2h_stage[c & 1];
    const size_t part = (len / THREADS + 4095) & ~size_t(4095);
    std::thread workers[THREADS - 1];
    int started = 0;
    for (int t = 1; t < THREADS; ++t) {
      const size_t b = size_t(t) * part;
      if (b >= len) break;
      workers[started++] = std::thread([=] { std::memcpy(dst + off + b, s + b, std::min(part, len - b)); });
    }
    std::memcpy(dst + off, s, std::min(part, len));
    for (int t = 0; t < started; ++t) workers[t].join();
  }
}

extern "C" int pq_load_tensor(pq_handle* h, const char* label, void* host_out, int host_dtype) {
  if (!h) return PQ_ERR_INVALID;
  PQ_TRY(h)
  PQ_REQUIRE(label && host_out, PQ_ERR_INVALID, "pq_load_tensor: bad arguments");
  PQ_REQUIRE(host_dtype == PQ_HOST_C64 || host_dtype == PQ_HOST_C128, PQ_ERR_INVALID,
             "pq_load_tensor: host dtype must be complex");
  set_device(h);
  Tensor& t = h->get(label);
  int64_t n = t.numel();
  bool direct = (h->dtype == PQ_C128 && host_dtype == PQ_HOST_C128) ||
                (h->dtype == PQ_C64 && host_dtype == PQ_HOST_C64);
  if (direct && size_t(n) * h->elem_size >= (size_t(64) << 20)) {
    d2h_pipelined(h, (unsigned char*)host_out, (const unsigned char*)t.buf->ptr, size_t(n) * h->elem_size);
    PQ_CUDA(cudaStreamSynchronize(h->stream));
  } else if (direct) {
    PQ_CUDA(cudaMemcpyAsync(host_out, t.buf->ptr, size_t(n) * h->elem_size, cudaMemcpyDeviceToHost,
                            h->stream));
    PQ_CUDA(cudaStreamSynchronize(h->stream));
  } else {
    std::vector<unsigned char> tmp(size_t(n) * h->elem_size);
    PQ_CUDA(cudaMemcpyAsync(tmp.data(), t.buf->ptr, tmp.size(), cudaMemcpyDeviceToHost, h->stream));
    PQ_CUDA(cudaStreamSynchronize(h->stream));
    if (h->dtype == PQ_C128) {
      const double* s = (const double*)tmp.data();
      float* d = (float*)host_out;
      for (int64_t i = 0; i < 2 * n; ++i) d[i] = (float)s[i];
    } else {
      const float* s = (const float*)tmp.data();
      double* d = (double*)host_out;
      for (int64_t i = 0; i < 2 * n; ++i) d[i] = (double)s[i];
    }
  }
  ozaki_t_check_watchdog();   // the stream is idle here
  PQ_CATCH(h)
}

// ---------------------------------------------------------------------------
// contract / permute / reshape / view / delete / save_output
// ---------------------------------------------------------------------------
extern "C" int pq_contract(pq_handle* h, const char* A, const int32_t* a_idx, int na, const char* B,
                           const int32_t* b_idx, int nb, const char* C) {
  if (!h) return PQ_ERR_INVALID;
  PQ_TRY(h)
  PQ_REQUIRE(A && B && C && na >= 0 && nb >= 0 && (na == 0 || a_idx) && (nb == 0 || b_idx),
             PQ_ERR_INVALID, "pq_contract: bad arguments");
  set_device(h);
  Tensor& ta = h->get(A);
  Tensor& tb = h->get(B);
  std::vector<int32_t> ai(a_idx, a_idx + na), bi(b_idx, b_idx + nb);
  ContractPlan p = lower_contract(ta.dims, ai, tb.dims, bi, h->elem_size, h->opt);
  Tensor tc;
  tc.dims = p.cdims;
  tc.buf = std::make_shared<Buffer>(size_t(p.M * p.N) * h->elem_size, h->stream);
  std::unique_ptr<Buffer> tA, tB, ws;
  if (p.tempA_bytes) tA.reset(new Buffer(p.tempA_bytes, h->stream));
  if (p.tempB_bytes) tB.reset(new Buffer(p.tempB_bytes, h->stream));
  if (p.ws_bytes) ws.reset(new Buffer(p.ws_bytes, h->stream));
  Launch L = h->launch_ctx();
  run_contract(L, p, ta.buf->ptr, tb.buf->ptr, tc.buf->ptr, tA ? tA->ptr : nullptr,
               tB ? tB->ptr : nullptr, ws ? ws->ptr : nullptr);
  h->n_contract += 1;
  h->macs += p.M * p.N * p.K;
  h->note_tensor(p.M * p.N);
  // save C, then delete A and B (interactive.jl:72-74); stream order keeps this safe
  std::string la(A), lb(B);
  h->tensors[C] = std::move(tc);
  if (la != C) h->tensors.erase(la);
  if (lb != C) h->tensors.erase(lb);
  PQ_CATCH(h)
}

extern "C" int pq_permute(pq_handle* h, const char* label, const int32_t* axes, int n) {
  if (!h) return PQ_ERR_INVALID;
  PQ_TRY(h)
  PQ_REQUIRE(label && n >= 0 && (n == 0 || axes), PQ_ERR_INVALID, "pq_permute: bad arguments");
  set_device(h);
  Tensor& t = h->get(label);
  PQ_REQUIRE((int)t.dims.size() == n, PQ_ERR_INVALID, "pq_permute: axes length != rank");
  std::vector<int> perm(n);
  for (int k = 0; k < n; ++k) perm[k] = axes[k] - 1;
  PermutePlan p = lower_permute(t.dims, perm, h->elem_size, h->opt);
  std::vector<int64_t> nd(n);
  for (int k = 0; k < n; ++k) nd[k] = t.dims[perm[k]];
  if (!p.identity) {
    auto nb = std::make_shared<Buffer>(size_t(p.total) * h->elem_size, h->stream);
    Launch L = h->launch_ctx();
    run_permute(L, p, t.buf->ptr, nb->ptr);
    t.buf = nb;  // the old buffer survives while an alias (save_output) still holds it
  }
  t.dims = nd;
  PQ_CATCH(h)
}

extern "C" int pq_reshape(pq_handle* h, const char* label, const int32_t* groups_flat,
                          const int32_t* group_sizes, int ngroups) {
  if (!h) return PQ_ERR_INVALID;
  PQ_TRY(h)
  PQ_REQUIRE(label && ngroups >= 0, PQ_ERR_INVALID, "pq_reshape: bad arguments");
  Tensor& t = h->get(label);
  std::vector<int64_t> nd;
  int pos = 0;
  for (int g = 0; g < ngroups; ++g) {
    int64_t d = 1;
    for (int q = 0; q < group_sizes[g]; ++q) {
      int ax = groups_flat[pos++];
      PQ_REQUIRE(ax >= 1 && ax <= (int)t.dims.size(), PQ_ERR_INVALID, "pq_reshape: axis out of range");
      d *= t.dims[ax - 1];
    }
    nd.push_back(d);
  }
  PQ_REQUIRE(prod(nd) == t.numel(), PQ_ERR_SHAPE, "DimensionMismatch: reshape changes the number of elements");
  t.dims = nd;
  PQ_CATCH(h)
}

extern "C" int pq_view(pq_handle* h, const char* view, const char* src, int axis, const int32_t* idx,
                       int nidx) {
  if (!h) return PQ_ERR_INVALID;
  PQ_TRY(h)
  PQ_REQUIRE(view && src && idx && nidx >= 1, PQ_ERR_INVALID, "pq_view: bad arguments");
  set_device(h);
  Tensor& t = h->get(src);
  PQ_REQUIRE(axis >= 1 && axis <= (int)t.dims.size(), PQ_ERR_INVALID, "pq_view: axis out of range");
  const int64_t ext = t.dims[axis - 1];
  for (int j = 0; j < nidx; ++j)
    PQ_REQUIRE(idx[j] >= 1 && idx[j] <= ext, PQ_ERR_INVALID, "BoundsError: view index out of range");
  int64_t inner = 1, outer = 1;
  for (int d = 0; d < axis - 1; ++d) inner *= t.dims[d];
  for (size_t d = axis; d < t.dims.size(); ++d) outer *= t.dims[d];
  Tensor v;
  v.dims = t.dims;
  v.dims[axis - 1] = nidx;
  v.buf = std::make_shared<Buffer>(size_t(v.numel()) * h->elem_size, h->stream);
  Launch L = h->launch_ctx();
  // split the index list into contiguous runs (a UnitRange is a single run)
  int j = 0;
  while (j < nidx) {
    int j1 = j + 1;
    while (j1 < nidx && idx[j1] == idx[j1 - 1] + 1) ++j1;
    if (j == 0 && j1 == nidx) {
      run_view(L, t.buf->ptr, v.buf->ptr, inner, ext, nidx, outer, idx[0], nullptr);
    } else {
      // run [j, j1): one launch per outer slab keeps the kernel simple
      for (int64_t o = 0; o < outer; ++o) {
        const char* in = (const char*)t.buf->ptr + size_t(inner * ext * o) * h->elem_size;
        char* out = (char*)v.buf->ptr + size_t(inner * (j + nidx * o)) * h->elem_size;
        run_view(L, in, out, inner, ext, j1 - j, 1, idx[j], nullptr);
      }
    }
    j = j1;
  }
  h->note_tensor(v.numel());
  h->tensors[view] = std::move(v);
  PQ_CATCH(h)
}

extern "C" int pq_decompose(pq_handle* h, const char* tensor, const int32_t* left_positions,
                            int nleft, const int32_t* right_positions, int nright, double threshold,
                            int max_rank, const char* left_label, const char* right_label,
                            int* chi_out) {
  if (!h) return PQ_ERR_INVALID;
  PQ_TRY(h)
  PQ_REQUIRE(tensor && left_label && right_label && nleft >= 0 && nright >= 0 &&
                 (nleft == 0 || left_positions) && (nright == 0 || right_positions),
             PQ_ERR_INVALID, "pq_decompose: bad arguments");
  set_device(h);
  Tensor& t = h->get(tensor);
  const int rank = (int)t.dims.size();
  PQ_REQUIRE(nleft + nright == rank, PQ_ERR_INVALID,
             "pq_decompose: left and right positions must cover every axis exactly once");
  std::vector<int> perm;
  std::vector<int64_t> ldims, rdims;
  for (int k = 0; k < nleft; ++k) {
    PQ_REQUIRE(left_positions[k] >= 1 && left_positions[k] <= rank, PQ_ERR_INVALID,
               "pq_decompose: position out of range");
    perm.push_back(left_positions[k] - 1);
    ldims.push_back(t.dims[left_positions[k] - 1]);
  }
  for (int k = 0; k < nright; ++k) {
    PQ_REQUIRE(right_positions[k] >= 1 && right_positions[k] <= rank, PQ_ERR_INVALID,
               "pq_decompose: position out of range");
    perm.push_back(right_positions[k] - 1);
    rdims.push_back(t.dims[right_positions[k] - 1]);
  }
  PermutePlan pp = lower_permute(t.dims, perm, h->elem_size, h->opt);  // validates the permutation
  const int64_t m = prod(ldims), n = prod(rdims);
  Launch L = h->launch_ctx();
  // the SVD overwrites its input: always work on a private [left | right] copy
  auto work = std::make_shared<Buffer>(size_t(m * n) * h->elem_size, h->stream);
  if (pp.identity) {
    L.begin(KC_COPY, 2.0 * m * n * h->elem_size, 0);
    PQ_CUDA(cudaMemcpyAsync(work->ptr, t.buf->ptr, size_t(m * n) * h->elem_size,
                            cudaMemcpyDeviceToDevice, h->stream));
    L.end();
  } else {
    run_permute(L, pp, t.buf->ptr, work->ptr);
  }
  Tensor B, C;
  const int chi = run_decompose(h, L, work->ptr, m, n, threshold, max_rank, B.buf, C.buf);
  B.dims = ldims;
  B.dims.push_back(chi);
  C.dims.push_back(chi);
  C.dims.insert(C.dims.end(), rdims.begin(), rdims.end());
  const std::string ll = left_label, rl = right_label, src = tensor;
  h->tensors[ll] = std::move(B);
  h->tensors[rl] = std::move(C);
  if (src != ll && src != rl) h->tensors.erase(src);
  if (chi_out) *chi_out = chi;
  PQ_CATCH(h)
}

extern "C" int pq_delete(pq_handle* h, const char* label) {
  if (!h || !label) return PQ_ERR_INVALID;
  h->tensors.erase(label);  // a missing label is not an error (interactive.jl:159-161)
  return PQ_OK;
}

extern "C" int pq_save_output(pq_handle* h, const char* node, const char* name) {
  if (!h) return PQ_ERR_INVALID;
  PQ_TRY(h)
  PQ_REQUIRE(node && name, PQ_ERR_INVALID, "pq_save_output: bad arguments");
  Tensor t = h->get(node);  // alias: shares the buffer (interactive.jl:84-88)
  h->tensors[name] = t;
  PQ_CATCH(h)
}

extern "C" int pq_sync(pq_handle* h) {
  if (!h) return PQ_ERR_INVALID;
  PQ_TRY(h)
  set_device(h);
  PQ_CUDA(cudaStreamSynchronize(h->stream));
  PQ_CUDA(cudaGetLastError());
  ozaki_t_check_watchdog();
  PQ_CATCH(h)
}

extern "C" int pq_accumulate(pq_handle* h, const char* dst, const char* src) {
  if (!h) return PQ_ERR_INVALID;
  PQ_TRY(h)
  PQ_REQUIRE(dst && src, PQ_ERR_INVALID, "pq_accumulate: bad arguments");
  set_device(h);
  Tensor& s = h->get(src);
  auto it = h->tensors.find(dst);
  Launch L = h->launch_ctx();
  if (it == h->tensors.end()) {
    Tensor d;
    d.dims = s.dims;
    d.buf = std::make_shared<Buffer>(size_t(s.numel()) * h->elem_size, h->stream);
    L.begin(KC_COPY, 2.0 * s.numel() * h->elem_size, 0);
    PQ_CUDA(cudaMemcpyAsync(d.buf->ptr, s.buf->ptr, size_t(s.numel()) * h->elem_size,
                            cudaMemcpyDeviceToDevice, h->stream));
    L.end();
    h->tensors[dst] = std::move(d);
  } else {
    Tensor& d = it->second;
    PQ_REQUIRE(d.numel() == s.numel(), PQ_ERR_SHAPE, "pq_accumulate: size mismatch");
    if (d.buf.use_count() > 1) {  // copy-on-write: never mutate an aliased buffer
      auto nb = std::make_shared<Buffer>(size_t(d.numel()) * h->elem_size, h->stream);
      PQ_CUDA(cudaMemcpyAsync(nb->ptr, d.buf->ptr, size_t(d.numel()) * h->elem_size,
                              cudaMemcpyDeviceToDevice, h->stream));
      d.buf = nb;
    }
    run_accumulate(L, d.buf->ptr, s.buf->ptr, s.numel());
  }
  PQ_CATCH(h)
}

// ---------------------------------------------------------------------------
// instrumentation
// ---------------------------------------------------------------------------
extern "C" int pq_get_counters(pq_handle* h, int64_t* n_contract, int64_t* macs, int64_t* max_elems,
                               int64_t* kernel_launches) {
  if (!h) return PQ_ERR_INVALID;
  if (n_contract) *n_contract = h->n_contract;
  if (macs) *macs = h->macs;
  if (max_elems) *max_elems = h->max_elems;
  if (kernel_launches) *kernel_launches = h->launches;
  return PQ_OK;
}

extern "C" int pq_reset_counters(pq_handle* h) {
  if (!h) return PQ_ERR_INVALID;
  h->n_contract = h->macs = h->max_elems = h->launches = 0;
  return PQ_OK;
}

extern "C" int pq_profile_enable(pq_handle* h, int on) {
  if (!h) return PQ_ERR_INVALID;
  h->drain_profile();
  h->profile = on != 0;
  if (on) {
    for (int c = 0; c < PQ_NUM_KERNEL_CLASSES; ++c) {
      h->prof_ms[c] = 0;
      h->prof_launches[c] = 0;
      h->prof_bytes[c] = 0;
      h->prof_flops[c] = 0;
    }
  }
  return PQ_OK;
}

extern "C" int pq_profile_read(pq_handle* h, double* ms, int64_t* launches, double* bytes,
                               double* flops) {
  if (!h) return PQ_ERR_INVALID;
  cudaSetDevice(h->device);
  h->drain_profile();
  for (int c = 0; c < PQ_NUM_KERNEL_CLASSES; ++c) {
    if (ms) ms[c] = h->prof_ms[c];
    if (launches) launches[c] = h->prof_launches[c];
    if (bytes) bytes[c] = h->prof_bytes[c];
    if (flops) flops[c] = h->prof_flops[c];
  }
  return PQ_OK;
}

extern "C" const char* pq_kernel_class_name(int cls) {
  static const char* names[PQ_NUM_KERNEL_CLASSES] = {
      "permute_tiled", "permute_generic", "contract_small", "contract_direct", "contract_dot",
      "gemm_simt",     "gemm_tensor",     "view",           "accumulate",      "copy",
      "allreduce",     "svd",             "other",          "contract_chain",  "gemm_int8"};
  return (cls >= 0 && cls < PQ_NUM_KERNEL_CLASSES) ? names[cls] : "?";
}

extern "C" int pq_timer_begin(pq_handle* h) {
  if (!h) return PQ_ERR_INVALID;
  PQ_TRY(h)
  set_device(h);
  if (!h->timer0) {
    PQ_CUDA(cudaEventCreate(&h->timer0));
    PQ_CUDA(cudaEventCreate(&h->timer1));
  }
  PQ_CUDA(cudaEventRecord(h->timer0, h->stream));
  PQ_CATCH(h)
}

extern "C" int pq_timer_end(pq_handle* h, double* ms) {
  if (!h) return PQ_ERR_INVALID;
  PQ_TRY(h)
  PQ_REQUIRE(ms && h->timer0, PQ_ERR_INVALID, "pq_timer_end without pq_timer_begin");
  set_device(h);
  PQ_CUDA(cudaEventRecord(h->timer1, h->stream));
  PQ_CUDA(cudaEventSynchronize(h->timer1));
  float f = 0;
  PQ_CUDA(cudaEventElapsedTime(&f, h->timer0, h->timer1));
  *ms = f;
  PQ_CATCH(h)
}

extern "C" int pq_set_option(pq_handle* h, const char* key, int value) {
  if (!h || !key) return PQ_ERR_INVALID;
  std::string k(key);
  if (k == "gemm") h->opt.gemm = value;
  else if (k == "permute") h->opt.permute = value;
  else if (k == "fused") h->opt.fused = value;
  else if (k == "graph") h->opt.graph = value;
  else if (k == "prio") h->opt.prio = value;
  else if (k == "chain") h->opt.chain = value;
  else if (k == "zgemm_stagger") h->opt.zgemm_stagger = value;
  else if (k == "zgemm_skinny") h->opt.zgemm_skinny = value;
  else if (k == "zgemm_3m") h->opt.zgemm_3m = value;
  else if (k == "cgemm_ozaki") {
    if (value != 0 && value != 4) {
      h->last_error = "cgemm_ozaki must be 0 or 4";
      return PQ_ERR_INVALID;
    }
    h->opt.cgemm_ozaki = value;
  }
  else if (k == "zgemm_ozaki") {
    if (value != 0 && value != 6) {
      h->last_error = "zgemm_ozaki must be 0 or 6";
      return PQ_ERR_INVALID;
    }
    h->opt.zgemm_ozaki = value;
  }
  else if (k == "ozaki_auto") h->opt.ozaki_auto = value;
  else if (k == "zgemm_cfg") h->opt.zgemm_cfg = value;
  else if (k == "zgemm_thin") h->opt.zgemm_thin = value;
  else if (k == "ozaki_tsw") h->opt.ozaki_tsw = value;
  else if (k == "small_tc") h->opt.small_tc = value;
  else if (k == "zgemm_kfirst") h->opt.zgemm_kfirst = value;
  else {
    h->last_error = "unknown option: " + k;
    return PQ_ERR_INVALID;
  }
  return PQ_OK;
}

extern "C" int pq_microbench(pq_handle* h, const char* what, double* result) {
  if (!h) return PQ_ERR_INVALID;
  PQ_TRY(h)
  PQ_REQUIRE(what && result, PQ_ERR_INVALID, "pq_microbench: bad arguments");
  set_device(h);
  PQ_CUDA(cudaStreamSynchronize(h->stream));
  Launch L = h->launch_ctx();
  L.profile = false;
  *result = run_microbench(L, what);
  PQ_CATCH(h)
}
