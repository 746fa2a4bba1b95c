// Small data-movement kernels: K7 slice copy (view_tensor!, reference
// src/layer1.jl:191-194), slice-partial accumulation, and the bandwidth probe.
#include <cstdio>
#include "common.h"

namespace pq {

double run_fp64_probe(const Launch& L, bool tensor, int warps_per_sm = 0);  // kernels_zgemm.cu

// out[i + inner*(j + nsel*o)] = in[i + inner*((start-1+j) + ext_in*o)]
// `start` (1-based) comes from device memory when start_dev != nullptr, so that one
// captured launch can serve every slice of a sliced contraction.
template <typename E>
__global__ void __launch_bounds__(256)
k_view(const E* __restrict__ in, E* __restrict__ out, long long inner, long long ext_in,
       long long nsel, long long outer, int start0, const int* __restrict__ start_dev) {
  const long long start = (start_dev ? (long long)(*start_dev) : (long long)start0) - 1;
  const long long total = inner * nsel * outer;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    long long i = t % inner, rest = t / inner;
    long long j = rest % nsel, o = rest / nsel;
    out[t] = in[i + inner * ((start + j) + ext_in * o)];
  }
}

void run_view(const Launch& L, const void* in, void* out, int64_t inner, int64_t ext_in,
              int64_t nsel, int64_t outer, int start0, const int32_t* start_dev) {
  long long total = inner * nsel * outer;
  if (total <= 0) return;
  long long blocks = (total + 255) / 256;
  long long cap = (long long)L.num_sms * 32;
  if (blocks > cap) blocks = cap;
  L.begin(KC_VIEW, 2.0 * double(total) * L.elem_size, 0);
  if (L.elem_size == 16)
    k_view<double2><<<(unsigned)blocks, 256, 0, L.stream>>>((const double2*)in, (double2*)out, inner,
                                                          ext_in, nsel, outer, start0, start_dev);
  else
    k_view<float2><<<(unsigned)blocks, 256, 0, L.stream>>>((const float2*)in, (float2*)out, inner,
                                                         ext_in, nsel, outer, start0, start_dev);
  L.end();
  PQ_CUDA(cudaGetLastError());
}

template <typename R>
__global__ void __launch_bounds__(256) k_accumulate(R* __restrict__ dst, const R* __restrict__ src,
                                                    long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] += src[i];
}

void run_accumulate(const Launch& L, void* dst, const void* src, int64_t n) {
  if (n <= 0) return;
  long long reals = 2 * n;
  long long blocks = (reals + 255) / 256;
  long long cap = (long long)L.num_sms * 32;
  if (blocks > cap) blocks = cap;
  L.begin(KC_ACCUMULATE, 3.0 * double(n) * L.elem_size, 2.0 * n);
  if (L.elem_size == 16)
    k_accumulate<double><<<(unsigned)blocks, 256, 0, L.stream>>>((double*)dst, (const double*)src,
                                                               reals);
  else
    k_accumulate<float><<<(unsigned)blocks, 256, 0, L.stream>>>((float*)dst, (const float*)src,
                                                              reals);
  L.end();
  PQ_CUDA(cudaGetLastError());
}

// one CTA per tensor of a pq_save_tensors batch: staging block -> the tensor's own buffer
__global__ void __launch_bounds__(128) k_scatter(const unsigned char* __restrict__ stage,
                                                 const ScatterItem* __restrict__ table) {
  const ScatterItem it = table[blockIdx.x];
  const uint2* src = reinterpret_cast<const uint2*>(stage + it.src_off);
  uint2* dst = reinterpret_cast<uint2*>(it.dst);
  for (unsigned long long i = threadIdx.x; i < it.words; i += blockDim.x) dst[i] = src[i];
}

void run_scatter(const Launch& L, const unsigned char* stage_dev, const ScatterItem* table, int n,
                 double bytes) {
  if (n <= 0) return;
  L.begin(KC_COPY, 2.0 * bytes, 0);
  k_scatter<<<(unsigned)n, 128, 0, L.stream>>>(stage_dev, table);
  L.end();
  PQ_CUDA(cudaGetLastError());
}

__global__ void __launch_bounds__(256) k_copy16(const uint4* __restrict__ in, uint4* __restrict__ out,
                                                long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = in[i];
}

// Memory-system probe for the sweep-step access pattern (pq_microbench "stream_<mode>_<ctas>"): a
// pure copy of 2^24 16-byte elements in 64 KB tiles, with the READ side laid out like the A operand
// of the dominant sweep step (mode bit 0: 8 runs of 8 KB per tile, 4 / 16 / 64 MB apart; else one
// contiguous 64 KB block) and the WRITE side like its result C[m + M n] (mode bit 1: 64 runs of
// 1 KB per tile, 4 MB apart; else contiguous).  `ctas` resident CTAs per SM x 256 threads x 16
// loads of 16 bytes in flight.  Returns GB/s (read + write).
__global__ void __launch_bounds__(256)
k_stream_probe(const uint4* __restrict__ in, uint4* __restrict__ out, int mode) {
  const long long tiles = 4096;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    // tile bits 0..8 -> element bits 9..17, tile bits 9, 10, 11 -> element bits 19, 21, 23
    const long long rbase = ((tile & 511) << 9) | (((tile >> 9) & 1) << 19) | (((tile >> 10) & 1) << 21) |
                            (((tile >> 11) & 1) << 23);
    uint4 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int u = threadIdx.x + 256 * i;   // 16-byte unit of the tile
      long long src;
      if (mode & 4) {
        // the producers' own lane mapping: thread = (tile row j, 16-k chunk c), load i reads k = 16 c + i:
        // a warp instruction touches 4 separate 128-byte pieces (8 rows x 16 bytes each)
        const int j = (threadIdx.x >> 5) * 8 + (threadIdx.x & 7), c = (threadIdx.x & 31) >> 3, k = 16 * c + i;
        src = rbase + (j & 7) + ((long long)(k & 7) << 3) + ((long long)(j >> 3) << 6) + ((long long)((k >> 3) & 1) << 18) +
              ((long long)((k >> 4) & 1) << 20) + ((long long)((k >> 5) & 1) << 22);
      } else if (mode & 1) {
        const int run = u >> 9, within = u & 511;   // run bits 0, 1, 2 -> element bits 18, 20, 22
        src = rbase + within + ((long long)(run & 1) << 18) + ((long long)((run >> 1) & 1) << 20) +
              ((long long)((run >> 2) & 1) << 22);
      } else {
        src = tile * 4096 + u;
      }
      v[i] = in[src];
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int u = threadIdx.x + 256 * i;
      long long dst;
      if (mode & 2) {
        const int n = u >> 6, row = u & 63;
        dst = tile * 64 + row + ((long long)n << 18);
      } else {
        dst = tile * 4096 + u;
      }
      out[dst] = v[i];
    }
  }
}

double run_microbench(const Launch& L, const std::string& what) {
  if (what.rfind("stream_", 0) == 0) {
    int mode = 0, ctas = 2;
    std::sscanf(what.c_str() + 7, "%d_%d", &mode, &ctas);
    const long long n = 1LL << 24;
    uint4 *a = nullptr, *b = nullptr;
    PQ_CUDA(cudaMalloc(&a, n * 16));
    PQ_CUDA(cudaMalloc(&b, n * 16));
    PQ_CUDA(cudaMemsetAsync(a, 1, n * 16, L.stream));
    cudaEvent_t e0, e1;
    PQ_CUDA(cudaEventCreate(&e0));
    PQ_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
      PQ_CUDA(cudaEventRecord(e0, L.stream));
      k_stream_probe<<<L.num_sms * ctas, 256, 0, L.stream>>>(a, b, mode);
      PQ_CUDA(cudaEventRecord(e1, L.stream));
      PQ_CUDA(cudaEventSynchronize(e1));
      float ms = 0;
      PQ_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(a);
    cudaFree(b);
    PQ_CUDA(err);
    return 2.0 * double(n) * 16.0 / (best * 1e-3) / 1e9;
  }
  if (what.rfind("ozaki_t_", 0) == 0 || what.rfind("umma_i8_", 0) == 0) return run_ozaki_t_microbench(L, what);
  if (what == "dmma_tflops") return run_fp64_probe(L, true);
  if (what.rfind("dmma_tflops_w", 0) == 0) return run_fp64_probe(L, true, std::stoi(what.substr(13)));
  if (what == "dfma_tflops") return run_fp64_probe(L, false);
  if (what == "copy_gbs") {
    const long long n = 1LL << 26;  // 1 GiB in + 1 GiB out, larger than L2
    uint4 *a = nullptr, *b = nullptr;
    PQ_CUDA(cudaMalloc(&a, n * 16));
    PQ_CUDA(cudaMalloc(&b, n * 16));
    PQ_CUDA(cudaMemsetAsync(a, 1, n * 16, L.stream));
    cudaEvent_t e0, e1;
    PQ_CUDA(cudaEventCreate(&e0));
    PQ_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
      PQ_CUDA(cudaEventRecord(e0, L.stream));
      k_copy16<<<L.num_sms * 16, 256, 0, L.stream>>>(a, b, n);
      PQ_CUDA(cudaEventRecord(e1, L.stream));
      PQ_CUDA(cudaEventSynchronize(e1));
      float ms = 0;
      PQ_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best) best = ms;
    }
    PQ_CUDA(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(a);
    cudaFree(b);
    return 2.0 * double(n) * 16.0 / (best * 1e-3) / 1e9;
  }
  throw Error(PQ_ERR_INVALID, "unknown microbench: " + what);
}

}  // namespace pq
