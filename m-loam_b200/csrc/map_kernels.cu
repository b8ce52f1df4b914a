// map_kernels.cu — GPU voxel-hash map build, exact kNN and the scan-to-map matcher (sm_100a).
//
//   map build  : replaces pcl::KdTreeFLANN::setInputCloud (lidar_mapper_keyframe.cpp:433-434)
//   k_knn      : replaces nearestKSearch (feature_extract.hpp:406,570,666,813)
//   k_match    : FeatureExtract::matchCornerFromMap / matchSurfFromMap, one warp per feature
//                (feature_extract.hpp:378-643; per-point forms :645-883)
//
// HBM layout: points are float4 (x,y,z,w); the map keeps a cell-major sorted copy whose w carries the
// original index, so a cell is one contiguous run of 16 B records (a 128 B line holds 8 points).  The hash
// table is open addressing over 16 B {key,start,count} records: one probe = one 16 B load.
#include "ctx.h"
#include "knn.cuh"

namespace mloam {

// ------------------------------------------------------------------------------------------- build
__global__ void k_table_clear(HashEntry *table, unsigned long long *block_mask, unsigned cap, int *cursor) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *cursor = 0;
  if (i < cap) {
    uint4 v;
    v.x = 0xffffffffu, v.y = 0xffffffffu, v.z = 0u, v.w = 0u;
    reinterpret_cast<uint4 *>(table)[i] = v;
    block_mask[i] = 0ull;
  }
}

// Pass 1: cell key per point, insert-or-find its slot, count.  rank_of = arrival order inside the cell
// (only the order of points inside a cell depends on it; results never do — kNN ties break on the index).
__global__ void k_map_insert(const float4 *__restrict__ pts, int m, float inv_cell, HashEntry *table, unsigned mask,
                             int *__restrict__ slot_of, int *__restrict__ rank_of) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const float4 p = pts[i];
  const int fx = (int)floorf(p.x * inv_cell), fy = (int)floorf(p.y * inv_cell), fz = (int)floorf(p.z * inv_cell);
  const unsigned long long key = pack_cell(fx, fy, fz);
  unsigned h = hash_cell(key) & mask;
  while (true) {
    unsigned long long *kp = &table[h].key;
    unsigned long long prev = *kp;
    if (prev == MLOAM_EMPTY_KEY) prev = atomicCAS(kp, MLOAM_EMPTY_KEY, key);
    if (prev == MLOAM_EMPTY_KEY || prev == key) break;
    h = (h + 1) & mask;
  }
  slot_of[i] = (int)h;
  rank_of[i] = atomicAdd(&table[h].count, 1);
}

// Pass 2, one acting thread per occupied CELL (the point that arrived first): the cell's run of `sorted` is claimed
// with one atomicAdd on a cursor (runs need to be contiguous per cell, not ordered across cells — no prefix scan over
// the 2m-slot table), and the cell is entered once into the occupancy record of its 4x4x4 block: same table, tagged
// key, point count in `start` (its `count` stays 0), occupied-cell bit in block_mask.
__device__ __forceinline__ int block_exclusive_scan_256(int v, int *total) {
  __shared__ int warp_sums[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(MLOAM_FULL_MASK, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = lane < 8 ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      int t = __shfl_up_sync(MLOAM_FULL_MASK, w, o);
      if (lane >= o) w += t;
    }
    if (lane < 8) warp_sums[lane] = w;
  }
  __syncthreads();
  const int base = wid > 0 ? warp_sums[wid - 1] : 0;
  *total = warp_sums[7];
  return base + inc - v;
}

__global__ void __launch_bounds__(256)
    k_map_assign(const float4 *__restrict__ pts, int m, float inv_cell, HashEntry *table, unsigned long long *block_mask, unsigned mask,
                 const int *__restrict__ slot_of, const int *__restrict__ rank_of, int *cursor) {
  __shared__ int block_base;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool act = i < m && rank_of[i] == 0;
  int h = 0, cnt = 0;
  if (act) h = slot_of[i], cnt = table[h].count;
  // one cursor atomic per CTA: block-wide exclusive scan of the cells' counts (a per-cell atomic on one address serialises)
  int total;
  const int excl = block_exclusive_scan_256(cnt, &total);
  if (threadIdx.x == 0) block_base = total > 0 ? atomicAdd(cursor, total) : 0;
  __syncthreads();
  if (!act) return;
  table[h].start = block_base + excl;
  const float4 p = pts[i];
  const int fx = (int)floorf(p.x * inv_cell), fy = (int)floorf(p.y * inv_cell), fz = (int)floorf(p.z * inv_cell);
  const unsigned long long ckey = coarse_key(fx >> MLOAM_COARSE_SHIFT, fy >> MLOAM_COARSE_SHIFT, fz >> MLOAM_COARSE_SHIFT);
  unsigned hc = hash_cell(ckey) & mask;
  while (true) {
    unsigned long long *kp = &table[hc].key;
    unsigned long long prev = *kp;
    if (prev == MLOAM_EMPTY_KEY) prev = atomicCAS(kp, MLOAM_EMPTY_KEY, ckey);
    if (prev == MLOAM_EMPTY_KEY || prev == ckey) break;
    hc = (hc + 1) & mask;
  }
  atomicAdd(&table[hc].start, cnt);
  atomicOr(&block_mask[hc], 1ull << (((fz & 3) << 4) | ((fy & 3) << 2) | (fx & 3)));
}

__global__ void k_map_scatter(const float4 *__restrict__ pts, int m, const HashEntry *__restrict__ table,
                              const int *__restrict__ slot_of, const int *__restrict__ rank_of, float4 *__restrict__ sorted,
                              float4 *__restrict__ orig) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  float4 p = pts[i];
  orig[i] = p;
  p.w = __int_as_float(i);
  sorted[table[slot_of[i]].start + rank_of[i]] = p;
}

static unsigned next_pow2(unsigned v) {
  unsigned p = 1;
  while (p < v) p <<= 1;
  return p;
}

int map_build_device(Ctx *c, int slot, const float4 *d_pts, int m, float cell) {
  if (slot < 0 || slot >= MLOAM_NUM_MAPS || m < 0 || !(cell > 0.f)) {
    c->err = "map_build: bad slot / size / cell";
    return MLOAM_E_INVALID;
  }
  MapStorage &M = c->maps[slot];
  ProfScope ps(c, "map_build");
  // fine cells + coarse blocks <= 2m records: keep at least two slots empty so every probe sequence terminates
  const unsigned cap = next_pow2((unsigned)(m > 511 ? 2 * (unsigned)m + 2 : 1024u));
  MLOAM_CUDA_OK(c, M.sorted.reserve(sizeof(float4) * (size_t)(m + 1)));
  MLOAM_CUDA_OK(c, M.orig.reserve(sizeof(float4) * (size_t)(m + 1)));
  MLOAM_CUDA_OK(c, M.table.reserve(sizeof(HashEntry) * (size_t)cap));
  MLOAM_CUDA_OK(c, M.block_mask.reserve(sizeof(unsigned long long) * (size_t)cap));
  MLOAM_CUDA_OK(c, M.slot_of.reserve(sizeof(int) * (size_t)(m + 1)));
  MLOAM_CUDA_OK(c, M.rank_of.reserve(sizeof(int) * (size_t)(m + 1)));
  MLOAM_CUDA_OK(c, M.scan_tmp.reserve(64));  // the run cursor
  M.capacity = cap, M.m = m, M.cell = cell, M.built = true;
  cudaStream_t st = c->stream;
  k_table_clear<<<(cap + 255) / 256, 256, 0, st>>>(M.table.as<HashEntry>(), M.block_mask.as<unsigned long long>(), cap, M.scan_tmp.as<int>());
  c->launches++;
  if (m > 0) {
    const int nb = (m + 255) / 256;
    k_map_insert<<<nb, 256, 0, st>>>(d_pts, m, 1.0f / cell, M.table.as<HashEntry>(), cap - 1, M.slot_of.as<int>(), M.rank_of.as<int>());
    k_map_assign<<<nb, 256, 0, st>>>(d_pts, m, 1.0f / cell, M.table.as<HashEntry>(), M.block_mask.as<unsigned long long>(), cap - 1,
                                     M.slot_of.as<int>(), M.rank_of.as<int>(), M.scan_tmp.as<int>());
    k_map_scatter<<<nb, 256, 0, st>>>(d_pts, m, M.table.as<HashEntry>(), M.slot_of.as<int>(), M.rank_of.as<int>(),
                                      M.sorted.as<float4>(), M.orig.as<float4>());
    c->launches += 3;
  }
  MLOAM_CUDA_OK(c, cudaGetLastError());
  return MLOAM_OK;
}

// ------------------------------------------------------------------------------------------- kNN
constexpr int QWARPS = 8;  // warps (= queries) per block

template <int K>
__global__ void __launch_bounds__(QWARPS * 32)
    k_knn(MapView map, const float4 *__restrict__ q, int nq, const double *__restrict__ pose7, float max_sqdist,
          int *__restrict__ idx, float *__restrict__ sqd) {
  __shared__ RunBuf rbuf[QWARPS];
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * QWARPS + (threadIdx.x >> 5); i < nq; i += gridDim.x * QWARPS) {
    const float4 p = __ldg(q + i);
    float3 s = make_float3(p.x, p.y, p.z);
    if (pose7) s = associate(pose_from_param(pose7), p.x, p.y, p.z);
    Best best;
    warp_knn<K, false>(map, rbuf[threadIdx.x >> 5], s.x, s.y, s.z, max_sqdist, lane, best);
    if (lane < K) {  // lane r holds the r-th neighbour
      const float d2 = key_d2(best.key);
      const bool ok = best.key != MLOAM_KEY_NONE && d2 < max_sqdist;
      idx[(size_t)i * K + lane] = ok ? (int)(unsigned)(best.key & 0xffffffffu) : -1;
      sqd[(size_t)i * K + lane] = ok ? d2 : INFINITY;
    }
  }
}

int knn_device(Ctx *c, int slot, const float4 *d_q, int nq, const double *d_pose7, int k, float max_sqdist, int *d_idx,
               float *d_sqd) {
  if (slot < 0 || slot >= MLOAM_NUM_MAPS || !c->maps[slot].built) {
    c->err = "knn: map slot not built";
    return MLOAM_E_STATE;
  }
  if (nq <= 0) return MLOAM_OK;
  ProfScope ps(c, "knn");
  MapView mv = c->maps[slot].view();
  int nb = (nq + QWARPS - 1) / QWARPS;
  if (nb > 8 * c->sm_count) nb = 8 * c->sm_count;  // warps stride over the queries
  cudaStream_t st = c->stream;
  switch (k) {
    case 1: k_knn<1><<<nb, QWARPS * 32, 0, st>>>(mv, d_q, nq, d_pose7, max_sqdist, d_idx, d_sqd); break;
    case 5: k_knn<5><<<nb, QWARPS * 32, 0, st>>>(mv, d_q, nq, d_pose7, max_sqdist, d_idx, d_sqd); break;
    case 10: k_knn<10><<<nb, QWARPS * 32, 0, st>>>(mv, d_q, nq, d_pose7, max_sqdist, d_idx, d_sqd); break;
    default: c->err = "knn: k must be 1, 5 or 10"; return MLOAM_E_INVALID;
  }
  c->launches++;
  MLOAM_CUDA_OK(c, cudaGetLastError());
  return MLOAM_OK;
}

}  // namespace mloam
