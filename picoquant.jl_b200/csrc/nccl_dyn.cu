// The one collective of the path: summing slice partials across the GPUs of a box
// (reference: MPI.Reduce! at examples/dist_slicing_example.jl:30).  NCCL is bound at
// run time with dlopen so that the library loads on boxes without it and shares the
// libnccl.so.2 already mapped by torch when driven from Python.
#include <dlfcn.h>
#include <nccl.h>

#include "handle.h"

namespace pq {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi& api() {
  static NcclApi a;
  if (a.lib) return a;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (a.lib) break;
  }
  if (!a.lib) throw Error(PQ_ERR_NCCL, std::string("cannot load libnccl: ") + dlerror());
  auto sym = [&](const char* s) {
    void* p = dlsym(a.lib, s);
    if (!p) throw Error(PQ_ERR_NCCL, std::string("missing NCCL symbol ") + s);
    return p;
  };
  a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
  a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
  a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
  a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
  return a;
}

#define PQ_NCCL(call)                                                                \
  do {                                                                               \
    ncclResult_t r__ = (call);                                                       \
    if (r__ != ncclSuccess)                                                          \
      throw Error(PQ_ERR_NCCL, std::string(#call) + ": " + api().GetErrorString(r__)); \
  } while (0)

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
};

void comm_destroy(Comm* c) {
  if (!c) return;
  if (c->comm) api().CommDestroy(c->comm);
  delete c;
}

}  // namespace pq

using namespace pq;

static_assert(sizeof(ncclUniqueId) == 128, "pq_comm_unique_id hands out 128 bytes");

extern "C" int pq_comm_unique_id(void* id128) {
  if (!id128) return PQ_ERR_INVALID;
  try {
    ncclUniqueId id;
    PQ_NCCL(api().GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
  } catch (const Error& e) {
    return e.code;
  }
  return PQ_OK;
}

extern "C" int pq_comm_init(pq_handle* h, const void* id128, int rank, int nranks) {
  if (!h || !id128) return PQ_ERR_INVALID;
  try {
    PQ_CUDA(cudaSetDevice(h->device));
    if (h->comm) {
      comm_destroy(h->comm);
      h->comm = nullptr;
    }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    Comm* c = new Comm();
    c->rank = rank;
    c->nranks = nranks;
    ncclResult_t r = api().CommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) {
      delete c;
      throw Error(PQ_ERR_NCCL, std::string("ncclCommInitRank: ") + api().GetErrorString(r));
    }
    h->comm = c;
  } catch (const Error& e) {
    h->last_error = e.what();
    return e.code;
  }
  return PQ_OK;
}

extern "C" int pq_allreduce_sum(pq_handle* h, const char* label) {
  if (!h || !label) return PQ_ERR_INVALID;
  try {
    PQ_CUDA(cudaSetDevice(h->device));
    PQ_REQUIRE(h->comm, PQ_ERR_NCCL, "pq_allreduce_sum: call pq_comm_init first");
    Tensor& t = h->get(label);
    if (t.buf.use_count() > 1) {  // never mutate an aliased buffer
      auto nb = std::make_shared<Buffer>(size_t(t.numel()) * h->elem_size, h->stream);
      PQ_CUDA(cudaMemcpyAsync(nb->ptr, t.buf->ptr, size_t(t.numel()) * h->elem_size,
                              cudaMemcpyDeviceToDevice, h->stream));
      t.buf = nb;
    }
    Launch L = h->launch_ctx();
    L.begin(KC_ALLREDUCE, 2.0 * t.numel() * h->elem_size, 0);
    PQ_NCCL(api().AllReduce(t.buf->ptr, t.buf->ptr, size_t(2 * t.numel()),
                            h->dtype == PQ_C128 ? ncclDouble : ncclFloat, ncclSum, h->comm->comm,
                            h->stream));
    L.end();
  } catch (const Error& e) {
    h->last_error = e.what();
    return e.code;
  }
  return PQ_OK;
}
