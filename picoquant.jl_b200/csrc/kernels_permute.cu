// K1: out-of-place index permutation (permutedims), the first half of TTGT and the
// backend's permute_tensor (reference: src/layer1.jl:111-114).  HBM-bound:
// algorithmic bytes = 2 * numel * sizeof(element).
//
//  * k_permute_tiled   -- power-of-two extents: the permutation is a permutation of
//    address bits.  Each CTA moves tiles of 2^t elements through shared memory; the
//    tile holds the >=5 lowest input bits and the >=5 lowest output bits, so both the
//    global reads (input order) and the global writes (output order) are runs of >=32
//    consecutive elements (512 B for c128).  An xor swizzle keeps both shared-memory
//    phases bank-conflict free.  Tiles are double-buffered with cp.async.
//  * k_permute_generic -- any extents: one thread per output element, coalesced
//    writes, gathered reads.  Fallback and cross-check.
#include "common.h"
#include "tile_math.h"

namespace pq {

template <typename E>
__global__ void __launch_bounds__(256)
k_permute_generic(const E* __restrict__ in, E* __restrict__ out, IdxMap map, long long total) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
    out[i] = in[map_offset(map, i)];
}

template <int BYTES>
__device__ __forceinline__ void cp_async_elem(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  if (BYTES == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}

// Two tile buffers: the cp.async gather of tile r+1 (global -> swizzled shared memory,
// no register staging) is in flight while tile r is being written out, so every CTA keeps
// reads and writes outstanding at the same time.
template <typename E>
__global__ void __launch_bounds__(256)
k_permute_tiled(const E* __restrict__ in, E* __restrict__ out, const TileParams tp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tile_elems = 1 << tp.t;
  const int nhi = 1 << (tp.t - TILE_LO);
  E* tile0 = reinterpret_cast<E*>(smem_raw);
  long long* in_hi = reinterpret_cast<long long*>(smem_raw + 2 * sizeof(E) * tile_elems);
  long long* out_hi = in_hi + nhi;
  int* e_hi = reinterpret_cast<int*>(out_hi + nhi);
  int* e_lo = e_hi + nhi;

  for (int x = threadIdx.x; x < nhi; x += blockDim.x) {
    in_hi[x] = tile_in_hi(tp, x);
    out_hi[x] = tile_out_hi(tp, x);
    e_hi[x] = tile_e_hi(tp, x);
  }
  if (threadIdx.x < 32) e_lo[threadIdx.x] = tile_e_lo(tp, threadIdx.x);
  __syncthreads();

  auto gather = [&](long long r, E* tile) {
    long long in_base, out_base;
    tile_bases(tp, r, in_base, out_base);
    const E* src = in + in_base;
#pragma unroll 4
    for (int e = threadIdx.x; e < tile_elems; e += 256)
      cp_async_elem<sizeof(E)>(tile + tile_swizzle(tp, e),
                               src + ((long long)(e & 31) | in_hi[e >> TILE_LO]));
  };

  long long r = blockIdx.x;
  if (r < tp.ntiles) gather(r, tile0);
  asm volatile("cp.async.commit_group;\n" ::);
  int buf = 0;
  for (; r < tp.ntiles; r += gridDim.x) {
    const long long rn = r + gridDim.x;
    if (rn < tp.ntiles) gather(rn, tile0 + (buf ^ 1) * tile_elems);
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 1;\n" ::);
    __syncthreads();
    long long in_base, out_base;
    tile_bases(tp, r, in_base, out_base);
    E* dst = out + out_base;
    const E* tile = tile0 + buf * tile_elems;
#pragma unroll 4
    for (int o = threadIdx.x; o < tile_elems; o += 256) {
      int e = e_lo[o & 31] | e_hi[o >> TILE_LO];
      dst[(long long)(o & 31) | out_hi[o >> TILE_LO]] = tile[tile_swizzle(tp, e)];
    }
    __syncthreads();  // the buffer just drained is the prefetch target of the next iteration
    buf ^= 1;
  }
  asm volatile("cp.async.wait_group 0;\n" ::);
}

template <typename E>
static void launch_permute(const Launch& L, const PermutePlan& p, const void* in, void* out) {
  const long long total = p.paired ? p.total / 2 : p.total;
  const double bytes = 2.0 * double(total) * sizeof(E);
  if (p.tiled && (L.opt == nullptr || L.opt->permute != 1)) {
    const TileParams& tp = p.tp;
    const int nhi = 1 << (tp.t - TILE_LO);
    size_t smem = 2 * sizeof(E) * (size_t(1) << tp.t) + size_t(nhi) * (8 + 8 + 4) + 32 * 4;
    long long grid = tp.ntiles;
    long long cap = (long long)L.num_sms * 6;  // ~6 resident CTAs per SM, several tiles each
    if (grid > cap) grid = cap;
    L.begin(KC_PERMUTE_TILED, bytes, 0);
    k_permute_tiled<E><<<(unsigned)grid, 256, smem, L.stream>>>((const E*)in, (E*)out, tp);
    L.end();
  } else {
    long long blocks = (total + 255) / 256;
    long long cap = (long long)L.num_sms * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    L.begin(KC_PERMUTE_GENERIC, bytes, 0);
    k_permute_generic<E><<<(unsigned)blocks, 256, 0, L.stream>>>((const E*)in, (E*)out, p.gmap,
                                                               total);
    L.end();
  }
  PQ_CUDA(cudaGetLastError());
}

void run_permute(const Launch& L, const PermutePlan& p, const void* in, void* out) {
  if (p.identity) {
    // pure copy (only reached when the caller needs a distinct buffer)
    L.begin(KC_COPY, 2.0 * double(p.total) * L.elem_size, 0);
    PQ_CUDA(cudaMemcpyAsync(out, in, size_t(p.total) * L.elem_size, cudaMemcpyDeviceToDevice,
                            L.stream));
    L.end();
    return;
  }
  if (L.elem_size == 16 || p.paired)
    launch_permute<double2>(L, p, in, out);  // c64 pairs travel as 16-byte elements
  else
    launch_permute<float2>(L, p, in, out);
}

}  // namespace pq
