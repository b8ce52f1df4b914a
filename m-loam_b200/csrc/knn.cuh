// knn.cuh — exact K-nearest-neighbour search over the GPU voxel-hash map, one WARP per query.
//
// Replaces pcl::KdTreeFLANN::nearestKSearch (feature_extract.hpp:155,293,406,570,666,813).
// Semantics: the K nearest map points in ascending (squared distance, original index) order, where the
// squared distance is FLANN's L2_Simple in float (dx*dx + dy*dy + dz*dz, rounded per operation), limited
// to points with d2 < max_sqdist — which is all the callers look at: every reference caller rejects the
// query unless sqdist[K-1] < MIN_MATCH_SQ_DIS (or sqdist[0] < DISTANCE_SQ_THRESHOLD for K=1).
//
// Search: cells are visited in Chebyshev shells around the query's cell.  In a shell each lane owns one
// cell: it probes the hash (one 16 B load) and scans that cell's points (contiguous float4, L1-resident
// after the first touch) into a private sorted top-K.  After a shell the 32 private lists are merged with
// warp reductions; the search stops when the K-th distance is below the distance to the faces of the cube
// visited so far (nothing outside can be closer), or when that face distance exceeds the search radius.
#pragma once
#include "common.cuh"

namespace mloam {

#define MLOAM_KEY_NONE 0xffffffffffffffffull

template <int K>
struct TopK {
  unsigned long long key[K];  // (float bits of d2) << 32 | original index
  int pos[K];                 // position in MapView::sorted
};

template <int K>
__device__ __forceinline__ void topk_reset(TopK<K> &t) {
#pragma unroll
  for (int i = 0; i < K; i++) t.key[i] = MLOAM_KEY_NONE, t.pos[i] = -1;
}

template <int K>
__device__ __forceinline__ void topk_insert(TopK<K> &t, unsigned long long key, int pos) {
  if (key < t.key[K - 1]) {
    t.key[K - 1] = key;
    t.pos[K - 1] = pos;
#pragma unroll
    for (int i = K - 1; i > 0; --i) {
      if (t.key[i] < t.key[i - 1]) {
        unsigned long long tk = t.key[i];
        t.key[i] = t.key[i - 1];
        t.key[i - 1] = tk;
        int tp = t.pos[i];
        t.pos[i] = t.pos[i - 1];
        t.pos[i - 1] = tp;
      }
    }
  }
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
  unsigned hi = (unsigned)(v >> 32);
  unsigned mh = __reduce_min_sync(MLOAM_FULL_MASK, hi);
  unsigned lo = (hi == mh) ? (unsigned)v : 0xffffffffu;
  unsigned ml = __reduce_min_sync(MLOAM_FULL_MASK, lo);
  return ((unsigned long long)mh << 32) | ml;
}

// Merge the 32 private lists into the warp-wide top-K.  On return lane 0 holds the merged list and every
// other lane an empty one (the union of private lists stays the best K seen so far); `out` is replicated.
template <int K>
__device__ __forceinline__ void warp_merge(TopK<K> &mine, TopK<K> &out, int lane) {
#pragma unroll
  for (int r = 0; r < K; r++) {
    unsigned long long cur = mine.key[0];
    unsigned long long m = warp_min_u64(cur);
    unsigned owners = __ballot_sync(MLOAM_FULL_MASK, cur == m);
    int src = __ffs(owners) - 1;
    int p = __shfl_sync(MLOAM_FULL_MASK, mine.pos[0], src);
    out.key[r] = m;
    out.pos[r] = (m == MLOAM_KEY_NONE) ? -1 : p;
    if (lane == src && m != MLOAM_KEY_NONE) {  // pop
#pragma unroll
      for (int i = 0; i < K - 1; i++) mine.key[i] = mine.key[i + 1], mine.pos[i] = mine.pos[i + 1];
      mine.key[K - 1] = MLOAM_KEY_NONE;
      mine.pos[K - 1] = -1;
    }
  }
  if (lane == 0) mine = out;
  else topk_reset(mine);
}

// Probe the open-addressing table.  *slot (optional) receives the record's slot (or -1).
__device__ __forceinline__ HashEntry hash_lookup(const MapView &map, unsigned long long key, int *slot = nullptr) {
  unsigned h = hash_cell(key) & map.mask;
  while (true) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(map.table + h));
    unsigned long long k = ((unsigned long long)raw.y << 32) | raw.x;
    if (k == key) {
      HashEntry e;
      e.key = k, e.start = (int)raw.z, e.count = (int)raw.w;
      if (slot) *slot = (int)h;
      return e;
    }
    if (k == MLOAM_EMPTY_KEY) {
      HashEntry e;
      e.key = k, e.start = 0, e.count = 0;
      if (slot) *slot = -1;
      return e;
    }
    h = (h + 1) & map.mask;
  }
}

// Scan one cell's points (contiguous float4 run) into a private top-K, 4 independent loads in flight.
template <int K>
__device__ __forceinline__ void scan_cell(const MapView &map, const HashEntry &e, float qx, float qy, float qz, TopK<K> &mine) {
  const float4 *p = map.sorted + e.start;
  for (int j = 0; j < e.count; j += 4) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) v[u] = __ldg(p + (j + u < e.count ? j + u : j));
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (j + u < e.count) {
        const float ex = v[u].x - qx, ey = v[u].y - qy, ez = v[u].z - qz;
        const float d2 = ex * ex + ey * ey + ez * ez;
        const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)__float_as_int(v[u].w);
        topk_insert(mine, key, e.start + j + u);
      }
    }
  }
}

__device__ __forceinline__ int cell_bit(int fx, int fy, int fz) { return ((fz & 3) << 4) | ((fy & 3) << 2) | (fx & 3); }

// Warp-cooperative search.  All lanes pass the same query; `out` is replicated in every lane.
//
// Step 0: 27 lanes fetch the occupancy records of the 27 blocks (4x4x4 cells) around the query's block: point
//   count + a 64-bit mask of occupied cells.  When a block edge is at least the search radius those blocks
//   contain every point of the search ball, so fewer than K points in them means no result can exist
//   (REJECT_PARTIAL callers stop after this one round — the fate of a feature with no map support).
// Ring 1: the 27 cells around the query's cell, one lane per cell, probing only cells whose mask bit is set;
//   warp-merge; stop if the K-th distance is inside the visited cube (the common case on a dense map).
// Otherwise: each lane walks the set bits of its own block's mask and visits the cells whose box is closer than
//   the current bound min(radius^2, K-th distance) — never an empty cell, never a cell outside the ball.
// If blocks are smaller than the radius (caller chose a tiny cell) the search falls back to plain shells.
template <int K, bool REJECT_PARTIAL>
__device__ __forceinline__ void warp_knn(const MapView &map, float qx, float qy, float qz, float max_sqdist, int lane,
                                         TopK<K> &out) {
  TopK<K> mine;
  topk_reset(mine);
  topk_reset(out);
  const int cx = (int)floorf(qx * map.inv_cell), cy = (int)floorf(qy * map.inv_cell), cz = (int)floorf(qz * map.inv_cell);
  const float eps = 1e-6f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + 8.0f * map.cell) + 1e-6f;
  const float radius = sqrtf(max_sqdist);
  const int ccx = cx >> MLOAM_COARSE_SHIFT, ccy = cy >> MLOAM_COARSE_SHIFT, ccz = cz >> MLOAM_COARSE_SHIFT;
  const bool coarse_ok = map.cell * (float)(1 << MLOAM_COARSE_SHIFT) >= radius * 1.0002f + 64.0f * eps;
  // ---- step 0: block occupancy
  unsigned long long bmask = 0ull;
  {
    int cnt = 0;
    if (lane < 27) {
      int slot;
      const HashEntry e = hash_lookup(map, coarse_key(ccx + lane % 3 - 1, ccy + (lane % 9) / 3 - 1, ccz + lane / 9 - 1), &slot);
      cnt = e.start;  // block records keep their point count in `start`
      if (slot >= 0) bmask = __ldg(map.block_mask + slot);
    }
    const int total = __reduce_add_sync(MLOAM_FULL_MASK, cnt);
    if (REJECT_PARTIAL && coarse_ok && total < K) return;
  }
  // ---- ring 1
  {
    const int dx = lane % 3 - 1, dy = (lane % 9) / 3 - 1, dz = lane / 9 - 1;  // lanes 27..31 idle
    const int fx = cx + dx, fy = cy + dy, fz = cz + dz;
    const int bl = ((fz >> MLOAM_COARSE_SHIFT) - ccz + 1) * 9 + ((fy >> MLOAM_COARSE_SHIFT) - ccy + 1) * 3 +
                   ((fx >> MLOAM_COARSE_SHIFT) - ccx + 1);
    const unsigned long long m = __shfl_sync(MLOAM_FULL_MASK, bmask, lane < 27 ? bl : 0);
    if (lane < 27 && ((m >> cell_bit(fx, fy, fz)) & 1ull)) {
      const HashEntry e = hash_lookup(map, pack_cell(fx, fy, fz));
      scan_cell<K>(map, e, qx, qy, qz, mine);
    }
    warp_merge(mine, out, lane);
  }
  auto face_gap = [&](int r) {  // distance from the query to the nearest face of the visited cube [c-r, c+r+1) * cell
    float g = qx - (float)(cx - r) * map.cell;
    g = fminf(g, (float)(cx + r + 1) * map.cell - qx);
    g = fminf(g, qy - (float)(cy - r) * map.cell);
    g = fminf(g, (float)(cy + r + 1) * map.cell - qy);
    g = fminf(g, qz - (float)(cz - r) * map.cell);
    g = fminf(g, (float)(cz + r + 1) * map.cell - qz);
    return g - eps;
  };
  auto kth = [&]() { return __uint_as_float((unsigned)(out.key[K - 1] >> 32)); };
  {
    const float g = face_gap(1);
    if (g > 0.0f) {
      const float g2 = g * g;
      if (g2 >= max_sqdist) return;
      if (out.key[K - 1] != MLOAM_KEY_NONE && kth() < g2) return;
    }
  }
  if (coarse_ok) {
    // ---- mask-guided completion: lane = block; visit occupied cells closer than the bound
    float bound = max_sqdist;
    if (out.key[K - 1] != MLOAM_KEY_NONE) bound = fminf(bound, kth());
    const int B = 1 << MLOAM_COARSE_SHIFT;
    // (a) each lane (= block) selects, with ALU work only, its occupied cells whose box is within the bound
    unsigned long long cand = 0ull;
    if (lane < 27 && bmask) {
      const int b0x = (ccx + lane % 3 - 1) * B, b0y = (ccy + (lane % 9) / 3 - 1) * B, b0z = (ccz + lane / 9 - 1) * B;
      const float bw = map.cell * (float)B;
      const float bgx = fmaxf(fmaxf((float)b0x * map.cell - qx, qx - ((float)b0x * map.cell + bw)) - eps, 0.0f);
      const float bgy = fmaxf(fmaxf((float)b0y * map.cell - qy, qy - ((float)b0y * map.cell + bw)) - eps, 0.0f);
      const float bgz = fmaxf(fmaxf((float)b0z * map.cell - qz, qz - ((float)b0z * map.cell + bw)) - eps, 0.0f);
      if (bgx * bgx + bgy * bgy + bgz * bgz <= bound) {  // the block itself reaches into the ball
        unsigned long long m = bmask;
        while (m) {
          const int b = __ffsll((long long)m) - 1;
          m &= m - 1;
          const int fx = b0x + (b & 3), fy = b0y + ((b >> 2) & 3), fz = b0z + (b >> 4);
          if (abs(fx - cx) <= 1 && abs(fy - cy) <= 1 && abs(fz - cz) <= 1) continue;  // ring 1 already did it
          // squared distance from q to the cell box (conservative by eps)
          const float lox = (float)fx * map.cell, loy = (float)fy * map.cell, loz = (float)fz * map.cell;
          const float gx = fmaxf(fmaxf(lox - qx, qx - (lox + map.cell)) - eps, 0.0f);
          const float gy = fmaxf(fmaxf(loy - qy, qy - (loy + map.cell)) - eps, 0.0f);
          const float gz = fmaxf(fmaxf(loz - qz, qz - (loz + map.cell)) - eps, 0.0f);
          if (gx * gx + gy * gy + gz * gz <= bound) cand |= 1ull << b;  // ties at equal distance are kept
        }
      }
    }
    // (b) the warp walks the blocks; the selected cells of a block are spread over the lanes (one cell per lane)
    const unsigned nonempty = __ballot_sync(MLOAM_FULL_MASK, cand != 0ull);
    for (unsigned todo = nonempty; todo; todo &= todo - 1) {
      const int bl = __ffs(todo) - 1;
      unsigned long long m = __shfl_sync(MLOAM_FULL_MASK, cand, bl);
      const int b0x = (ccx + bl % 3 - 1) * B, b0y = (ccy + (bl % 9) / 3 - 1) * B, b0z = (ccz + bl / 9 - 1) * B;
      while (m) {
        const unsigned lo = (unsigned)m, hi = (unsigned)(m >> 32);
        const int nlo = __popc(lo), cnt = nlo + __popc(hi);
        if (lane < cnt) {
          const int b = lane < nlo ? (int)__fns(lo, 0, lane + 1) : 32 + (int)__fns(hi, 0, lane - nlo + 1);
          const HashEntry e = hash_lookup(map, pack_cell(b0x + (b & 3), b0y + ((b >> 2) & 3), b0z + (b >> 4)));
          scan_cell<K>(map, e, qx, qy, qz, mine);
        }
        if (cnt <= 32) break;
        for (int t = 0; t < 32; t++) m &= m - 1;  // drop the 32 cells just handled
      }
    }
    warp_merge(mine, out, lane);
    return;
  }
  // ---- fallback: plain Chebyshev shells (blocks do not cover the search ball)
  int rmax = (int)ceilf(radius * map.inv_cell) + 1;
  if (rmax > 16) rmax = 16;
  for (int r = 2; r <= rmax; r++) {
    const int s = 2 * r + 1;
    const int ncell = s * s * s;
    for (int base = 0; base < ncell; base += 32) {
      const int c = base + lane;
      if (c < ncell) {
        const int dz = c / (s * s) - r;
        const int rem = c % (s * s);
        const int dy = rem / s - r;
        const int dx = rem % s - r;
        if (max(max(abs(dx), abs(dy)), abs(dz)) == r) {
          const HashEntry e = hash_lookup(map, pack_cell(cx + dx, cy + dy, cz + dz));
          scan_cell<K>(map, e, qx, qy, qz, mine);
        }
      }
    }
    warp_merge(mine, out, lane);
    const float g = face_gap(r);
    if (g > 0.0f) {
      const float g2 = g * g;
      if (g2 >= max_sqdist) break;
      if (out.key[K - 1] != MLOAM_KEY_NONE && kth() < g2) break;
    }
  }
}

}  // namespace mloam
