// Index arithmetic of the tiled bit-permutation kernel, shared between the CUDA
// kernel (kernels_permute.cu) and the host-side emulation used by the CPU tests
// (test_lower.cpp), so the trickiest integer code is checked without a GPU.
//
// A tensor whose extents are all powers of two is a vector of 2^n elements and an
// index permutation is a permutation of the n address bits.  A tile is the set of
// 2^t elements obtained by fixing the n-t "rest" bits; the tile bits contain the
// lowest `a` input bits (contiguous reads) and the lowest `b` output bits
// (contiguous writes).  Inside the tile, `e` enumerates elements in input order and
// `o` in output order; both are split into a 5-bit low part (identity, because
// a, b >= 5) and a high part looked up in small shared-memory tables.
#pragma once
#include "common.h"

namespace pq {

constexpr int TILE_LO = 5;

// offset contributed by the high part x = e >> 5 of an input-order tile index
__host__ __device__ inline long long tile_in_hi(const TileParams& tp, int x) {
  long long off = 0;
  for (int u = TILE_LO; u < tp.t; ++u)
    off |= (long long)((x >> (u - TILE_LO)) & 1) << tp.tin_pos[u];
  return off;
}
// offset contributed by the high part y = o >> 5 of an output-order tile index
__host__ __device__ inline long long tile_out_hi(const TileParams& tp, int y) {
  long long off = 0;
  for (int v = TILE_LO; v < tp.t; ++v)
    off |= (long long)((y >> (v - TILE_LO)) & 1) << tp.tout_pos[v];
  return off;
}
// input-order tile index e of the element whose output-order index has low part z
__host__ __device__ inline int tile_e_lo(const TileParams& tp, int z) {
  int e = 0;
  for (int v = 0; v < TILE_LO && v < tp.t; ++v) e |= ((z >> v) & 1) << tp.emap[v];
  return e;
}
__host__ __device__ inline int tile_e_hi(const TileParams& tp, int y) {
  int e = 0;
  for (int v = TILE_LO; v < tp.t; ++v) e |= ((y >> (v - TILE_LO)) & 1) << tp.emap[v];
  return e;
}
// shared-memory slot of tile element e (xor swizzle, a bijection on [0, 2^t))
__host__ __device__ inline int tile_swizzle(const TileParams& tp, int e) {
  int s = e;
  for (int k = 0; k < tp.nswz; ++k) s ^= ((e >> tp.swz_src[k]) & 1) << tp.swz_dst[k];
  return s;
}
// base offsets of tile r in the input and output tensors
__host__ __device__ inline void tile_bases(const TileParams& tp, long long r, long long& in_base,
                                           long long& out_base) {
  long long ib = 0, ob = 0;
  for (int w = 0; w < tp.nrest; ++w) {
    long long bit = (r >> w) & 1;
    ib |= bit << tp.rest_in[w];
    ob |= bit << tp.rest_out[w];
  }
  in_base = ib;
  out_base = ob;
}

}  // namespace pq
