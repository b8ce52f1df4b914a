// knn.cuh — exact K-nearest-neighbour search over the GPU voxel-hash map, one WARP per query.
//
// Replaces pcl::KdTreeFLANN::nearestKSearch (feature_extract.hpp:155,293,406,570,666,813).
// Semantics: the K nearest map points in ascending (squared distance, original index) order, where the
// squared distance is FLANN's L2_Simple in float (dx*dx + dy*dy + dz*dz, rounded per operation), limited
// to points with d2 < max_sqdist — which is all the callers look at: every reference caller rejects the
// query unless sqdist[K-1] < MIN_MATCH_SQ_DIS (or sqdist[0] < DISTANCE_SQ_THRESHOLD for K=1).
//
// Search plan per query (all 32 lanes cooperate; the running best list is DISTRIBUTED: lane r holds the r-th best):
//   step 0  27 lanes fetch the occupancy records of the 27 blocks (4x4x4 cells) around the query's block: point
//           count + 64-bit mask of occupied cells.  When a block edge is at least the search radius those blocks
//           contain every point of the search ball, so fewer than K points in them means no result can exist
//           (REJECT_PARTIAL callers stop here: the fate of a feature with no map support).
//   ring 1  the 27 cells around the query's cell, one lane per cell, probing only cells whose mask bit is set.
//           The cells' point runs (start, count) go to a per-warp shared-memory run table; the points of ALL runs
//           are then scanned as one flat list, 4 per lane per step (independent 16 B loads, perfectly balanced),
//           and reduced with a warp-wide K-selection (no per-lane sorted lists, no divergent insertion sort).
//           Stop if the K-th distance is inside the visited cube — the common case on a dense map.
//   finish  otherwise the occupied cells of all blocks whose box reaches into the current bound min(radius^2, K-th)
//           form one flat candidate list; 32 candidates per step are tested against the bound, the survivors probed
//           together and their runs packed into the run table (flushed through the same flat scan) — never an empty
//           cell, never a cell outside the ball, one dependent table access per 32 candidate cells.
// If blocks are smaller than the radius (caller chose a tiny cell) the finish falls back to plain shells.
//
// Code size matters as much as instruction count here: every launch starts with a cold instruction cache and the
// first query of each warp walks the whole search path, so the scan + selection (used from five places) is ONE
// out-of-line function with rolled selection rounds, and the best list costs three registers per lane.
#pragma once
#include "common.cuh"

namespace mloam {

#define MLOAM_KEY_NONE 0xffffffffffffffffull

// Running best list of a warp: lane r (r < N) holds the r-th smallest key seen so far, lanes >= N hold NONE.
// key = (float bits of d2) << 32 | original index; pos = position in MapView::sorted.
struct Best {
  unsigned long long key;
  int pos;
};
__device__ __forceinline__ Best best_none() { return Best{MLOAM_KEY_NONE, -1}; }
__device__ __forceinline__ unsigned long long best_key(const Best &b, int r) { return __shfl_sync(MLOAM_FULL_MASK, b.key, r); }
__device__ __forceinline__ float key_d2(unsigned long long k) { return __uint_as_float((unsigned)(k >> 32)); }

// Per-warp run table (shared memory): run r = points [start[r], start[r] + count) with pref[r] = points before it.
// The flat scan always reads all KNN_RUNS prefix entries: unused entries must hold the total.
constexpr int KNN_RUNS = 64;
struct RunBuf {
  int start[KNN_RUNS];
  int pref[KNN_RUNS];
};

// Optional per-query instrumentation of the blind search (stage profiling only).
struct KnnDbg {
  long long t_coarse, t_ring1, t_finish;  // SM cycles per phase
  int ring1_pts, finish_pts, finish_blocks, finish_cells;
};

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
  unsigned hi = (unsigned)(v >> 32);
  unsigned mh = __reduce_min_sync(MLOAM_FULL_MASK, hi);
  unsigned lo = (hi == mh) ? (unsigned)v : 0xffffffffu;
  unsigned ml = __reduce_min_sync(MLOAM_FULL_MASK, lo);
  return ((unsigned long long)mh << 32) | ml;
}

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int *total) {
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(MLOAM_FULL_MASK, inc, o);
    if (lane >= o) inc += t;
  }
  *total = __shfl_sync(MLOAM_FULL_MASK, inc, 31);
  return inc - v;
}

// Scan the flat concatenation of the runs in `rb` (`total` points) and merge them into the best list: the new list is
// the N smallest of (current best) U (scanned points).  Out of line on purpose (see the header comment).
template <int N>
__device__ __noinline__ Best scan_runs(const float4 *__restrict__ sorted, const RunBuf *rb, int total, float qx, float qy, float qz,
                                       Best best) {
  const int lane = threadIdx.x & 31;
  for (int base = 0; base < total; base += 128) {
    unsigned long long ck[4];
    int cp[4];
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int t = base + u * 32 + lane;
      cp[u] = -1;
      if (t < total) {
        int r = 0;
#pragma unroll
        for (int step = KNN_RUNS / 2; step > 0; step >>= 1)
          if (rb->pref[r + step] <= t) r += step;  // largest r with pref[r] <= t (empty runs share a prefix value)
        cp[u] = rb->start[r] + (t - rb->pref[r]);
        v[u] = __ldg(sorted + cp[u]);
      }
    }
    unsigned long long cmin = MLOAM_KEY_NONE;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      ck[u] = MLOAM_KEY_NONE;
      if (cp[u] >= 0) {
        const float ex = v[u].x - qx, ey = v[u].y - qy, ez = v[u].z - qz;
        const float d2 = ex * ex + ey * ey + ez * ez;
        ck[u] = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)__float_as_int(v[u].w);
      }
      cmin = ck[u] < cmin ? ck[u] : cmin;
    }
    // cheap exit: nothing beats the current N-th
    const unsigned long long kn = best_key(best, N - 1);
    if (!__any_sync(MLOAM_FULL_MASK, cmin < kn)) continue;
    // selection: N rounds, each extracts the smallest remaining key of (old list entry of this lane) U (its 4 candidates)
    unsigned long long ek = best.key, nk = MLOAM_KEY_NONE;
    int ep = best.pos, np = -1;
#pragma unroll 1
    for (int r = 0; r < N; r++) {
      unsigned long long lm = ek;
      int lp = ep, which = 4;
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (ck[u] < lm) lm = ck[u], lp = cp[u], which = u;
      const unsigned long long m = warp_min_u64(lm);
      if (m == MLOAM_KEY_NONE) break;  // fewer than N points so far
      const unsigned owners = __ballot_sync(MLOAM_FULL_MASK, lm == m);
      const int src = __ffs(owners) - 1;  // keys are unique (they embed the point index): one owner
      const int p = __shfl_sync(MLOAM_FULL_MASK, lp, src);
      if (lane == r) nk = m, np = p;
      if (lane == src) {
        if (which == 4) ek = MLOAM_KEY_NONE;
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (which == u) ck[u] = MLOAM_KEY_NONE;
      }
    }
    best.key = nk, best.pos = np;
  }
  return best;
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

__device__ __forceinline__ int cell_bit(int fx, int fy, int fz) { return ((fz & 3) << 4) | ((fy & 3) << 2) | (fx & 3); }

// Fill the whole run table from one (start, count) per lane (lanes beyond the used ones pass count 0).
__device__ __forceinline__ int fill_runs32(RunBuf &rb, int start, int count, int lane) {
  int total;
  const int excl = warp_excl_scan(count, lane, &total);
  __syncwarp();
  rb.start[lane] = start;
  rb.pref[lane] = excl;
  rb.pref[lane + 32] = total;
  __syncwarp();
  return total;
}

// Seeded search (temporal coherence between the re-association iterations of one scan2MapOptimization): the caller
// knows K map points — the previous iteration's neighbours — whose largest squared distance to the moved query is
// r2 < max_sqdist.  Every point of the true K-nearest set then lies in the ball of radius sqrt(r2), so scanning the
// cells that intersect that ball (usually 1-8 instead of the 27 + 27 probes of the blind search) gives the exact
// result, ties included.  pad > 0 widens the ball so that the (K+1)-th distance is seen too.  Returns false (nothing
// written) when the ball needs more than 32 cells.  *explored: every map point closer than this has been scanned.
template <int N>
__device__ __forceinline__ bool warp_knn_seeded(const MapView &map, RunBuf &rb, float qx, float qy, float qz, float r2, float pad,
                                                int lane, Best &out, float *explored) {
  const float eps = 1e-6f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + 8.0f * map.cell) + 1e-6f;
  const float rr = sqrtf(r2) * 1.0002f + eps + pad;
  const int lx = (int)floorf((qx - rr) * map.inv_cell), hx = (int)floorf((qx + rr) * map.inv_cell);
  const int ly = (int)floorf((qy - rr) * map.inv_cell), hy = (int)floorf((qy + rr) * map.inv_cell);
  const int lz = (int)floorf((qz - rr) * map.inv_cell), hz = (int)floorf((qz + rr) * map.inv_cell);
  const int nx = hx - lx + 1, ny = hy - ly + 1, nz = hz - lz + 1;
  if (nx > 32 || ny > 32 || nz > 32 || nx * ny * nz > 32) return false;
  const int ncell = nx * ny * nz;
  int start = 0, count = 0;
  if (lane < ncell) {
    const int fx = lx + lane % nx, fy = ly + (lane / nx) % ny, fz = lz + lane / (nx * ny);
    const float lox = (float)fx * map.cell, loy = (float)fy * map.cell, loz = (float)fz * map.cell;
    const float gx = fmaxf(fmaxf(lox - qx, qx - (lox + map.cell)) - eps, 0.0f);
    const float gy = fmaxf(fmaxf(loy - qy, qy - (loy + map.cell)) - eps, 0.0f);
    const float gz = fmaxf(fmaxf(loz - qz, qz - (loz + map.cell)) - eps, 0.0f);
    if (gx * gx + gy * gy + gz * gz <= rr * rr) {
      const HashEntry e = hash_lookup(map, pack_cell(fx, fy, fz));
      start = e.start, count = e.count;
    }
  }
  const int total = fill_runs32(rb, start, count, lane);
  out = scan_runs<N>(map.sorted, &rb, total, qx, qy, qz, best_none());
  __syncwarp();
  *explored = fmaxf(rr - 2.0f * eps, 0.0f);
  return true;
}

// Plain Chebyshev shells r = 2.. (blocks do not cover the search ball): rare, out of line.
template <int K, int N>
__device__ __noinline__ Best knn_shells(const MapView &map, RunBuf &rb, float qx, float qy, float qz, float max_sqdist, Best out) {
  const int lane = threadIdx.x & 31;
  const int cx = (int)floorf(qx * map.inv_cell), cy = (int)floorf(qy * map.inv_cell), cz = (int)floorf(qz * map.inv_cell);
  const float eps = 1e-6f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + 8.0f * map.cell) + 1e-6f;
  int rmax = (int)ceilf(sqrtf(max_sqdist) * map.inv_cell) + 1;
  if (rmax > 16) rmax = 16;
  for (int r = 2; r <= rmax; r++) {
    const int s = 2 * r + 1;
    const int ncell = s * s * s;
    for (int base = 0; base < ncell; base += 32) {
      const int c = base + lane;
      int start = 0, count = 0;
      if (c < ncell) {
        const int dz = c / (s * s) - r;
        const int rem = c % (s * s);
        const int dy = rem / s - r;
        const int dx = rem % s - r;
        if (max(max(abs(dx), abs(dy)), abs(dz)) == r) {
          const HashEntry e = hash_lookup(map, pack_cell(cx + dx, cy + dy, cz + dz));
          start = e.start, count = e.count;
        }
      }
      if (!__any_sync(MLOAM_FULL_MASK, count > 0)) continue;
      const int total = fill_runs32(rb, start, count, lane);
      out = scan_runs<N>(map.sorted, &rb, total, qx, qy, qz, out);
    }
    float g = qx - (float)(cx - r) * map.cell;
    g = fminf(g, (float)(cx + r + 1) * map.cell - qx);
    g = fminf(g, qy - (float)(cy - r) * map.cell);
    g = fminf(g, (float)(cy + r + 1) * map.cell - qy);
    g = fminf(g, qz - (float)(cz - r) * map.cell);
    g = fminf(g, (float)(cz + r + 1) * map.cell - qz);
    g -= eps;
    if (g > 0.0f) {
      const float g2 = g * g;
      if (g2 >= max_sqdist) break;
      const unsigned long long kk = best_key(out, K - 1);
      if (kk != MLOAM_KEY_NONE && key_d2(kk) < g2) break;
    }
  }
  return out;
}

// REJECT_PARTIAL: the caller only wants results when K neighbours exist inside the radius (every matcher gate).
// rb: this warp's run table in shared memory.
// N >= K: selection width.  The search is driven by the K-th distance; with N = K + 1 the extra slot holds the nearest
// point outside the K-set AMONG THE SCANNED ONES.
// *explored (nullable): every map point closer than this has been scanned into `out` — or, on the REJECT_PARTIAL
// block exit, counted: fewer than K points exist inside that distance.  Hence min(K-th scanned, *explored) bounds the
// true K-th distance from below, and min((K+1)-th scanned, *explored) the (K+1)-th.  pad > 0 lets the mask-guided
// finish look that much beyond min(radius, K-th), so that the caller also learns how isolated the K-set (or how far
// from K neighbours a rejected query) is; the result inside the radius is unaffected.
template <int K, bool REJECT_PARTIAL, int N = K>
__device__ __forceinline__ void warp_knn(const MapView &map, RunBuf &rb, float qx, float qy, float qz, float max_sqdist, int lane,
                                         Best &out, float *explored = nullptr, float pad = 0.0f, KnnDbg *dbg = nullptr) {
  long long t_mark = dbg ? clock64() : 0ll;
  out = best_none();
  if (explored) *explored = 0.0f;
  const int cx = (int)floorf(qx * map.inv_cell), cy = (int)floorf(qy * map.inv_cell), cz = (int)floorf(qz * map.inv_cell);
  const float eps = 1e-6f * (fabsf(qx) + fabsf(qy) + fabsf(qz) + 8.0f * map.cell) + 1e-6f;
  const float radius = sqrtf(max_sqdist);
  const int B = 1 << MLOAM_COARSE_SHIFT;
  const int ccx = cx >> MLOAM_COARSE_SHIFT, ccy = cy >> MLOAM_COARSE_SHIFT, ccz = cz >> MLOAM_COARSE_SHIFT;
  const bool coarse_ok = map.cell * (float)B >= radius * 1.0002f + 64.0f * eps;
  // ---- step 0 + ring 1 probes, issued together: lane l < 27 fetches the occupancy record of block l of the 3x3x3
  // blocks around the query's block (point count + 64-bit mask of occupied cells; the mask is read speculatively from
  // the record's home slot) AND the record of cell l of the 3x3x3 cells around the query's cell — three independent
  // loads in flight per lane, one memory round trip instead of three dependent ones.
  unsigned long long bmask = 0ull;
  int start = 0, count = 0;  // this lane's ring-1 run
  {
    int cnt = 0;
    if (lane < 27) {
      const int dx = lane % 3 - 1, dy = (lane % 9) / 3 - 1, dz = lane / 9 - 1;
      const unsigned long long kb = coarse_key(ccx + dx, ccy + dy, ccz + dz), kc = pack_cell(cx + dx, cy + dy, cz + dz);
      unsigned hb = hash_cell(kb) & map.mask, hc = hash_cell(kc) & map.mask;
      uint4 eb = __ldg(reinterpret_cast<const uint4 *>(map.table + hb));
      uint4 ec = __ldg(reinterpret_cast<const uint4 *>(map.table + hc));
      const unsigned long long m_home = __ldg(map.block_mask + hb);
      bool moved = false;
      while (true) {  // block record
        const unsigned long long k = ((unsigned long long)eb.y << 32) | eb.x;
        if (k == kb) {
          cnt = (int)eb.z;  // block records keep their point count in `start`
          bmask = moved ? __ldg(map.block_mask + hb) : m_home;
          break;
        }
        if (k == MLOAM_EMPTY_KEY) break;
        hb = (hb + 1) & map.mask, moved = true;
        eb = __ldg(reinterpret_cast<const uint4 *>(map.table + hb));
      }
      while (true) {  // cell record
        const unsigned long long k = ((unsigned long long)ec.y << 32) | ec.x;
        if (k == kc) {
          start = (int)ec.z, count = (int)ec.w;
          break;
        }
        if (k == MLOAM_EMPTY_KEY) break;
        hc = (hc + 1) & map.mask;
        ec = __ldg(reinterpret_cast<const uint4 *>(map.table + hc));
      }
    }
    const int total = __reduce_add_sync(MLOAM_FULL_MASK, cnt);
    if (dbg) {
      const long long t = clock64();
      dbg->t_coarse = t - t_mark, t_mark = t;
    }
    if (REJECT_PARTIAL && coarse_ok && total < K) {
      if (explored) {  // distance from the query to the hull of the 3x3x3 blocks
        const float bw = map.cell * (float)B;
        float g = qx - (float)((ccx - 1) * B) * map.cell;
        g = fminf(g, (float)((ccx + 2) * B) * map.cell - qx);
        g = fminf(g, qy - (float)((ccy - 1) * B) * map.cell);
        g = fminf(g, (float)((ccy + 2) * B) * map.cell - qy);
        g = fminf(g, qz - (float)((ccz - 1) * B) * map.cell);
        g = fminf(g, (float)((ccz + 2) * B) * map.cell - qz);
        *explored = fmaxf(fminf(g, 2.0f * bw) - eps, 0.0f);
      }
      return;
    }
  }
  // ---- ring 1: scan the 27 cells' runs as one flat list
  {
    const int total = fill_runs32(rb, start, count, lane);
    out = scan_runs<N>(map.sorted, &rb, total, qx, qy, qz, out);
    if (dbg) {
      const long long t = clock64();
      dbg->t_ring1 = t - t_mark, t_mark = t, dbg->ring1_pts = total;
    }
  }
  unsigned long long kk = best_key(out, K - 1);  // K-th so far
  {
    // distance from the query to the nearest face of the visited cube [c-1, c+2) * cell
    float g = qx - (float)(cx - 1) * map.cell;
    g = fminf(g, (float)(cx + 2) * map.cell - qx);
    g = fminf(g, qy - (float)(cy - 1) * map.cell);
    g = fminf(g, (float)(cy + 2) * map.cell - qy);
    g = fminf(g, qz - (float)(cz - 1) * map.cell);
    g = fminf(g, (float)(cz + 2) * map.cell - qz);
    g -= eps;
    if (g > 0.0f) {
      const float g2 = g * g;
      if (g2 >= max_sqdist || (kk != MLOAM_KEY_NONE && key_d2(kk) < g2)) {
        if (explored) *explored = g;
        return;
      }
    }
  }
  if (!coarse_ok) {
    out = knn_shells<K, N>(map, rb, qx, qy, qz, max_sqdist, out);
    return;
  }
  // ---- finish with the block masks
  // lim: min(radius^2, K-th so far) — what the result needs.  Cells are pruned against `bound`, which with pad > 0
  // reaches pad beyond sqrt(lim) (but never beyond what the 27 blocks cover).
  const float cover = fmaxf((map.cell * (float)B - 64.0f * eps) / 1.0002f, radius);
  auto prune_of = [&](float lim2) {
    if (!(pad > 0.0f)) return lim2;
    const float r = fminf(sqrtf(lim2) + pad, cover);
    return fmaxf(lim2, r * r);
  };
  float lim = max_sqdist;
  if (kk != MLOAM_KEY_NONE) lim = fminf(lim, key_d2(kk));
  float bound = prune_of(lim);
  // blocks that are non-empty and whose box reaches into the bound (ties at equal distance are kept)
  bool reach = false;
  if (lane < 27 && bmask) {
    const float bw = map.cell * (float)B;
    const float x0 = (float)((ccx + lane % 3 - 1) * B) * map.cell, y0 = (float)((ccy + (lane % 9) / 3 - 1) * B) * map.cell,
                z0 = (float)((ccz + lane / 9 - 1) * B) * map.cell;
    const float gx = fmaxf(fmaxf(x0 - qx, qx - (x0 + bw)) - eps, 0.0f);
    const float gy = fmaxf(fmaxf(y0 - qy, qy - (y0 + bw)) - eps, 0.0f);
    const float gz = fmaxf(fmaxf(z0 - qz, qz - (z0 + bw)) - eps, 0.0f);
    reach = gx * gx + gy * gy + gz * gz <= bound;
  }
  int nr = 0, npts = 0;  // runs / points currently in the table (warp-uniform)
  __syncwarp();          // ring 1 is done reading the table
  // The occupied cells of all reaching blocks form ONE flat candidate list (block b contributes popc(mask_b)
  // entries): 32 candidates per step are tested against the bound and probed together.
  const int my_cnt = reach ? __popcll(bmask) : 0;
  int n_cand;
  const int my_base = warp_excl_scan(my_cnt, lane, &n_cand);
  auto flush = [&]() {
    __syncwarp();
    for (int r = nr + lane; r < KNN_RUNS; r += 32) rb.pref[r] = npts;
    __syncwarp();
    out = scan_runs<N>(map.sorted, &rb, npts, qx, qy, qz, out);
    if (dbg) dbg->finish_pts += npts, dbg->finish_cells += nr;
    kk = best_key(out, K - 1);
    if (kk != MLOAM_KEY_NONE) lim = fminf(lim, key_d2(kk)), bound = prune_of(lim);
    nr = 0, npts = 0;
    __syncwarp();
  };
#pragma unroll 1
  for (int t0 = 0; t0 < n_cand; t0 += 32) {
    const int t = t0 + lane;
    int bl = 0;  // largest block index with base <= t
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
      const int cand = bl + step;
      const int v = __shfl_sync(MLOAM_FULL_MASK, my_base, cand < 27 ? cand : 26);
      if (cand < 27 && v <= t) bl = cand;
    }
    const unsigned long long m = __shfl_sync(MLOAM_FULL_MASK, bmask, bl);
    const int u = t - __shfl_sync(MLOAM_FULL_MASK, my_base, bl);
    bool take = t < n_cand;
    int fx = 0, fy = 0, fz = 0;
    if (take) {
      const unsigned lo = (unsigned)m, hi = (unsigned)(m >> 32);
      const int nlo = __popc(lo);
      const int b = u < nlo ? (int)__fns(lo, 0u, u + 1) : 32 + (int)__fns(hi, 0u, u - nlo + 1);
      fx = (ccx + bl % 3 - 1) * B + (b & 3), fy = (ccy + (bl % 9) / 3 - 1) * B + ((b >> 2) & 3), fz = (ccz + bl / 9 - 1) * B + (b >> 4);
      if (abs(fx - cx) <= 1 && abs(fy - cy) <= 1 && abs(fz - cz) <= 1) take = false;  // ring 1 did it
      if (take) {
        const float lox = (float)fx * map.cell, loy = (float)fy * map.cell, loz = (float)fz * map.cell;
        const float gx = fmaxf(fmaxf(lox - qx, qx - (lox + map.cell)) - eps, 0.0f);
        const float gy = fmaxf(fmaxf(loy - qy, qy - (loy + map.cell)) - eps, 0.0f);
        const float gz = fmaxf(fmaxf(loz - qz, qz - (loz + map.cell)) - eps, 0.0f);
        take = gx * gx + gy * gy + gz * gz <= bound;
      }
    }
    if (dbg) dbg->finish_blocks++;
    const unsigned tk = __ballot_sync(MLOAM_FULL_MASK, take);
    if (!tk) continue;
    const int ncell = __popc(tk);
    if (nr + ncell > KNN_RUNS) flush();
    int start = 0, count = 0;
    if (take) {
      const HashEntry e = hash_lookup(map, pack_cell(fx, fy, fz));
      start = e.start, count = e.count;
    }
    int tot;
    const int excl = warp_excl_scan(count, lane, &tot);
    if (take) {
      const int slot = nr + __popc(tk & ((1u << lane) - 1u));
      rb.start[slot] = start;
      rb.pref[slot] = npts + excl;
    }
    nr += ncell, npts += tot;
  }
  if (nr > 0) flush();
  if (dbg) dbg->t_finish = clock64() - t_mark;
  if (explored) *explored = fmaxf(sqrtf(prune_of(lim)) - eps, 0.0f);  // bounds never dropped below the final one
}

}  // namespace mloam
