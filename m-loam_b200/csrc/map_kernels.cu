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
#include "fit.cuh"
#include "knn.cuh"

namespace mloam {

// ------------------------------------------------------------------------------------------- build
__global__ void k_table_clear(HashEntry *table, unsigned long long *block_mask, unsigned cap) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap) {
    uint4 v;
    v.x = 0xffffffffu, v.y = 0xffffffffu, v.z = 0u, v.w = 0u;
    reinterpret_cast<uint4 *>(table)[i] = v;
    block_mask[i] = 0ull;
  }
}

// Pass 1: cell key per point, insert-or-find its slot, count.  rank_of = arrival order inside the cell
// (only the order of points inside a cell depends on it; results never do — kNN ties break on the index).
__global__ void k_map_insert(const float4 *__restrict__ pts, int m, float inv_cell, HashEntry *table, unsigned long long *block_mask,
                             unsigned mask, int *__restrict__ slot_of, int *__restrict__ rank_of) {
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
  // Coarse occupancy record of the 4x4x4-cell block this cell belongs to: same table, tagged key, the point
  // count lives in `start` (its `count` stays 0 so the start-offset scan ignores it).
  const unsigned long long ckey = coarse_key(fx >> MLOAM_COARSE_SHIFT, fy >> MLOAM_COARSE_SHIFT, fz >> MLOAM_COARSE_SHIFT);
  unsigned hc = hash_cell(ckey) & mask;
  while (true) {
    unsigned long long *kp = &table[hc].key;
    unsigned long long prev = *kp;
    if (prev == MLOAM_EMPTY_KEY) prev = atomicCAS(kp, MLOAM_EMPTY_KEY, ckey);
    if (prev == MLOAM_EMPTY_KEY || prev == ckey) break;
    hc = (hc + 1) & mask;
  }
  atomicAdd(&table[hc].start, 1);
  const unsigned long long bit = 1ull << (((fz & 3) << 4) | ((fy & 3) << 2) | (fx & 3));
  if (!(block_mask[hc] & bit)) atomicOr(&block_mask[hc], bit);
}

// Exclusive scan of table[].count into table[].start: block totals -> scan of totals -> apply.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int *total) {
  __shared__ int warp_sums[SCAN_THREADS / 32];
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
    int w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(MLOAM_FULL_MASK, w, o);
      if (lane >= o) w += t;
    }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = w;
  }
  __syncthreads();
  const int base = wid > 0 ? warp_sums[wid - 1] : 0;
  if (total) *total = warp_sums[SCAN_THREADS / 32 - 1];
  __syncthreads();
  return base + inc - v;
}

__global__ void k_scan_tile_sums(const HashEntry *table, unsigned cap, int *tile_sums) {
  const unsigned base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++)
    if (base + k < cap) s += table[base + k].count;
  int total;
  block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
__global__ void k_scan_tiles(int *tile_sums, int n_tiles) {
  // single block; n_tiles can exceed the block size, so walk in chunks carrying the running total
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_tiles; base += SCAN_THREADS) {
    int i = base + threadIdx.x;
    int v = i < n_tiles ? tile_sums[i] : 0;
    int total;
    int ex = block_exclusive_scan(v, &total);
    if (i < n_tiles) tile_sums[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
}
__global__ void k_scan_apply(HashEntry *table, unsigned cap, const int *tile_sums) {
  const unsigned base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int c[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    c[k] = (base + k < cap) ? table[base + k].count : 0;
    s += c[k];
  }
  int ex = block_exclusive_scan(s, nullptr) + tile_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    // tagged (coarse) and empty records keep their `start` (bit 63 is set in both)
    if (base + k < cap && !(table[base + k].key & MLOAM_COARSE_TAG)) table[base + k].start = ex;
    ex += c[k];
  }
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
  const int n_tiles = (int)((cap + SCAN_TILE - 1) / SCAN_TILE);
  MLOAM_CUDA_OK(c, M.scan_tmp.reserve(sizeof(int) * (size_t)n_tiles));
  M.capacity = cap, M.m = m, M.cell = cell, M.built = true;
  cudaStream_t st = c->stream;
  k_table_clear<<<(cap + 255) / 256, 256, 0, st>>>(M.table.as<HashEntry>(), M.block_mask.as<unsigned long long>(), cap);
  c->launches++;
  if (m > 0) {
    const int nb = (m + 255) / 256;
    k_map_insert<<<nb, 256, 0, st>>>(d_pts, m, 1.0f / cell, M.table.as<HashEntry>(), M.block_mask.as<unsigned long long>(), cap - 1,
                                     M.slot_of.as<int>(), M.rank_of.as<int>());
    k_scan_tile_sums<<<n_tiles, SCAN_THREADS, 0, st>>>(M.table.as<HashEntry>(), cap, M.scan_tmp.as<int>());
    k_scan_tiles<<<1, SCAN_THREADS, 0, st>>>(M.scan_tmp.as<int>(), n_tiles);
    k_scan_apply<<<n_tiles, SCAN_THREADS, 0, st>>>(M.table.as<HashEntry>(), cap, M.scan_tmp.as<int>());
    k_map_scatter<<<nb, 256, 0, st>>>(d_pts, m, M.table.as<HashEntry>(), M.slot_of.as<int>(), M.rank_of.as<int>(),
                                      M.sorted.as<float4>(), M.orig.as<float4>());
    c->launches += 5;
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
    TopK<K> best;
    warp_knn<K, false>(map, rbuf[threadIdx.x >> 5], s.x, s.y, s.z, max_sqdist, lane, best);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < K; k++) {
        const float d2 = __uint_as_float((unsigned)(best.key[k] >> 32));
        const bool ok = best.key[k] != MLOAM_KEY_NONE && d2 < max_sqdist;
        idx[(size_t)i * K + k] = ok ? (int)(unsigned)(best.key[k] & 0xffffffffu) : -1;
        sqd[(size_t)i * K + k] = ok ? d2 : INFINITY;
      }
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

// ------------------------------------------------------------------------------------------- match
// FOV gate, feature_extract.hpp:696-715 (and :434-458, :599-618, :842-861)
__device__ __forceinline__ bool in_laser_fov(const PoseD &T, const float3 &sel) {
  const float3 zt = associate(T, 0.0f, 0.0f, 10.0f);
  const double ex = T.t.x - (double)sel.x, ey = T.t.y - (double)sel.y, ez = T.t.z - (double)sel.z;
  const float s1 = (float)(ex * ex + ey * ey + ez * ez);
  const float ax = zt.x - sel.x, ay = zt.y - sel.y, az = zt.z - sel.z;
  const float s2 = ax * ax + ay * ay + az * az;
  const float check1 = 100.0f + s1 - s2 - 10.0f * sqrtf(3.0f) * sqrtf(s1);
  const float check2 = 100.0f + s1 - s2 + 10.0f * sqrtf(3.0f) * sqrtf(s1);
  return check1 < 0 && check2 > 0;
}

template <int K, bool IS_PLANE>
__global__ void __launch_bounds__(QWARPS * 32)
    k_match(MapView map, const float4 *__restrict__ pts, int n, const int *__restrict__ d_n, const double *__restrict__ pose7,
            float min_match_sq_dis, float min_plane_dis, int check_fov, unsigned char *__restrict__ valid,
            float *__restrict__ coeff, int *__restrict__ nn, int *__restrict__ work) {
  __shared__ RunBuf rbuf[QWARPS];
  const int lane = threadIdx.x & 31;
  if (d_n) n = min(n, *d_n);  // feature count produced on the device (no host round trip)
  const PoseD T = pose_from_param(pose7);
  // The grid is sized for the SM count, not for the (loose) upper bound.  With a work-queue head (`work`, zeroed
  // by k_lm between launches) every warp pulls the next feature when it finishes one, so a long query (sparse
  // neighbourhood) does not stall a whole wave; without one, warps stride statically.
  int i = blockIdx.x * QWARPS + (threadIdx.x >> 5);
  while (true) {
  if (work) {
    int t = 0;
    if (lane == 0) t = atomicAdd(work, 1);
    i = __shfl_sync(MLOAM_FULL_MASK, t, 0);
  }
  if (i >= n) break;
  const float4 p = __ldg(pts + i);
  const float3 sel = associate(T, p.x, p.y, p.z);  // pointAssociateToMap, utility.h:103-117
  TopK<K> best;
  warp_knn<K, true>(map, rbuf[threadIdx.x >> 5], sel.x, sel.y, sel.z, min_match_sq_dis, lane, best);
  bool ok = best.key[K - 1] != MLOAM_KEY_NONE &&
            __uint_as_float((unsigned)(best.key[K - 1] >> 32)) < min_match_sq_dis;  // :407,571,667,814
  float out[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (ok) {
    float X[K][3];
#pragma unroll
    for (int j = 0; j < K; j++) {
      const float4 v = __ldg(map.sorted + best.pos[j]);
      X[j][0] = v.x, X[j][1] = v.y, X[j][2] = v.z;
    }
    if (IS_PLANE) {
      // :573-594 / :817-837
      float A[K][3];
#pragma unroll
      for (int j = 0; j < K; j++) A[j][0] = X[j][0], A[j][1] = X[j][1], A[j][2] = X[j][2];
      float nv[3];
      ok = lsq_plane_dev<K>(A, nv);
      if (ok) {
        const float nrm = sqrtf(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]);
        const float d = 1 / nrm;
        nv[0] = nv[0] / nrm, nv[1] = nv[1] / nrm, nv[2] = nv[2] / nrm;
#pragma unroll
        for (int j = 0; j < K; j++)
          if (fabsf(nv[0] * X[j][0] + nv[1] * X[j][1] + nv[2] * X[j][2] + d) > min_plane_dis) ok = false;
        out[0] = nv[0], out[1] = nv[1], out[2] = nv[2], out[3] = d;
      }
    } else {
      // :410-432 / :670-693
      float cx = 0.f, cy = 0.f, cz = 0.f;
#pragma unroll
      for (int j = 0; j < K; j++) cx = cx + X[j][0], cy = cy + X[j][1], cz = cz + X[j][2];
      const float kf = (float)K;
      cx = cx / kf, cy = cy / kf, cz = cz / kf;
      float c00 = 0.f, c01 = 0.f, c02 = 0.f, c11 = 0.f, c12 = 0.f, c22 = 0.f;
#pragma unroll
      for (int j = 0; j < K; j++) {
        const float a = X[j][0] - cx, b = X[j][1] - cy, c = X[j][2] - cz;
        c00 = c00 + a * a, c01 = c01 + a * b, c02 = c02 + a * c;
        c11 = c11 + b * b, c12 = c12 + b * c, c22 = c22 + c * c;
      }
      float w[3], V[3][3];
      eig3f_dev(c00, c01, c02, c11, c12, c22, w, V);
      ok = w[2] > 3 * w[1];
      const float k01 = 0.1f;
      out[0] = k01 * V[0][2] + cx, out[1] = k01 * V[1][2] + cy, out[2] = k01 * V[2][2] + cz;
      out[3] = -k01 * V[0][2] + cx, out[4] = -k01 * V[1][2] + cy, out[5] = -k01 * V[2][2] + cz;
    }
    if (ok && check_fov) ok = in_laser_fov(T, sel);
  }
  if (lane == 0) {
    valid[i] = ok ? 1 : 0;
#pragma unroll
    for (int j = 0; j < 6; j++) coeff[(size_t)i * 6 + j] = ok ? out[j] : 0.f;
    if (nn) {
#pragma unroll
      for (int j = 0; j < K; j++)
        nn[(size_t)i * K + j] = (ok && best.key[j] != MLOAM_KEY_NONE) ? (int)(unsigned)(best.key[j] & 0xffffffffu) : -1;
    }
  }
  if (!work) i += gridDim.x * QWARPS;
  }  // feature loop
}

int match_from_map_device(Ctx *c, int slot, int type, const float4 *d_pts, int n, const int *d_n, const double *d_pose7,
                          const MatchCfg &cfg, unsigned char *d_valid, float *d_coeff, int *d_nn, int *d_work) {
  if (slot < 0 || slot >= MLOAM_NUM_MAPS || !c->maps[slot].built) {
    c->err = "match_from_map: map slot not built";
    return MLOAM_E_STATE;
  }
  if (type != 'c' && type != 's') {
    c->err = "match_from_map: type must be 'c' or 's'";
    return MLOAM_E_INVALID;
  }
  if (n <= 0) return MLOAM_OK;
  ProfScope ps(c, "match");
  MapView mv = c->maps[slot].view();
  int nb = (n + QWARPS - 1) / QWARPS;
  if (nb > 8 * c->sm_count) nb = 8 * c->sm_count;  // 8 CTAs x 8 warps per SM; warps stride over the features
  cudaStream_t st = c->stream;
#define MLOAM_LAUNCH_MATCH(KK, PL)                                                                                     \
  k_match<KK, PL><<<nb, QWARPS * 32, 0, st>>>(mv, d_pts, n, d_n, d_pose7, cfg.min_match_sq_dis, cfg.min_plane_dis, cfg.check_fov, \
                                              d_valid, d_coeff, d_nn, d_work)
  if (cfg.n_neigh == 5) {
    if (type == 's') MLOAM_LAUNCH_MATCH(5, true);
    else MLOAM_LAUNCH_MATCH(5, false);
  } else if (cfg.n_neigh == 10) {
    if (type == 's') MLOAM_LAUNCH_MATCH(10, true);
    else MLOAM_LAUNCH_MATCH(10, false);
  } else {
    c->err = "match_from_map: n_neigh must be 5 or 10";
    return MLOAM_E_INVALID;
  }
#undef MLOAM_LAUNCH_MATCH
  c->launches++;
  MLOAM_CUDA_OK(c, cudaGetLastError());
  return MLOAM_OK;
}

}  // namespace mloam
