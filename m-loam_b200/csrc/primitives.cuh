// primitives.cuh — hand-written device-wide building blocks used by the extraction / voxel-grid kernels:
// exclusive scan over int arrays and a stable LSD radix sort of (u64 key, u32 payload) pairs.
// Stability matters: it is what reproduces "ascending voxel index, then input order" for the centroid sums.
#pragma once
#include "common.cuh"

namespace mloam {

constexpr int PRIM_THREADS = 256;
constexpr int PRIM_ITEMS = 8;
constexpr int PRIM_TILE = PRIM_THREADS * PRIM_ITEMS;

// Exclusive scan of one int per thread across the block; *total (optional, shared or local) gets the block sum.
__device__ __forceinline__ int prim_block_scan(int v, int *total) {
  __shared__ int ws[PRIM_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(MLOAM_FULL_MASK, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) ws[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = lane < PRIM_THREADS / 32 ? ws[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(MLOAM_FULL_MASK, w, o);
      if (lane >= o) w += t;
    }
    if (lane < PRIM_THREADS / 32) ws[lane] = w;
  }
  __syncthreads();
  const int base = wid > 0 ? ws[wid - 1] : 0;
  if (total) *total = ws[PRIM_THREADS / 32 - 1];
  __syncthreads();
  return base + inc - v;
}

__global__ void k_prim_tile_sums(const int *__restrict__ in, int n, int *__restrict__ tile_sums);
__global__ void k_prim_scan_tiles(int *tile_sums, int n_tiles, int *total_out);
__global__ void k_prim_scan_apply(const int *__restrict__ in, int n, const int *__restrict__ tile_sums, int *__restrict__ out);
__global__ void k_rs_hist(const unsigned long long *__restrict__ keys, int n, int shift, int *__restrict__ hist, int nblk);
__global__ void k_rs_scatter(const unsigned long long *__restrict__ keys, const unsigned *__restrict__ vals, int n, int shift,
                             const int *__restrict__ offs, int nblk, unsigned long long *__restrict__ keys_out,
                             unsigned *__restrict__ vals_out);

}  // namespace mloam
