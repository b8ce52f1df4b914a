// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
// Point clouds, exact kNN (stand-in for pcl::KdTreeFLANN / FLANN KDTreeSingleIndex, leaf 15),
// pcl::VoxelGrid (PCL 1.8.0, un-vendored; algorithm mirrored in-tree at
// mloam_pcl/include/mloam_pcl/voxel_grid_covariance_mloam_impl.hpp:84-116,206-250) and
// FeatureExtract::extractCloud (estimator/src/featureExtract/feature_extract.cpp:118-297).
#pragma once
#include "orc_math.hpp"
#include <numeric>

namespace orc {

// pcl::PointXYZI payload (common::PointI, mloam_common/.../types/type.h:20). 16 B here; the
// reference's 32 B struct only adds padding.
struct PointI {
  float x, y, z, intensity;
};
typedef std::vector<PointI> Cloud;

// ============================================================================ exact kNN
// Result order: ascending (squared distance, index) — FLANN/nanoflann return ascending distance;
// the index tie-break is ours (ties are implementation-ordered in the reference).
// Distance is FLANN's L2_Simple in float: ((dx*dx) + dy*dy) + dz*dz.
struct KnnHit {
  float d2;
  int idx;
};
inline bool hit_less(const KnnHit &a, const KnnHit &b) { return a.d2 < b.d2 || (a.d2 == b.d2 && a.idx < b.idx); }
inline float dist2f(const PointI &a, float qx, float qy, float qz) {
  float dx = a.x - qx, dy = a.y - qy, dz = a.z - qz;
  return dx * dx + dy * dy + dz * dz;
}

// Optional backend for the TIMED CPU arm (bench.py --impl reference / cpu_baseline): the reference's own vendored
// nanoflann (oracle/_ref/libref_knn.so, compiled in place from /root/reference) builds and searches the tree.  Same
// exact answers (tests pin both to each other); ties between equal distances are re-ordered by index afterwards.
struct RefTreeApi {
  void *(*create)(const float *pts, int m) = nullptr;
  int (*knn)(void *tree, float qx, float qy, float qz, int k, int *idx, float *sqd) = nullptr;
  void (*destroy)(void *tree) = nullptr;
};
inline RefTreeApi &ref_tree_api() {
  static RefTreeApi api;
  return api;
}

class KdTree {
 public:
  KdTree() = default;
  KdTree(const KdTree &) = delete;
  KdTree &operator=(const KdTree &) = delete;
  ~KdTree() {
    if (ref_) ref_tree_api().destroy(ref_);
  }
  // pcl::KdTreeFLANN::setInputCloud call sites: lidar_tracker.cpp:33-34, lidar_mapper_keyframe.cpp:433-434,
  // estimator.cpp:1129-1130,1231-1233
  void setInputCloud(const Cloud *cloud) {
    cloud_ = cloud;
    if (ref_) ref_tree_api().destroy(ref_), ref_ = nullptr;
    if (ref_tree_api().create && !cloud->empty()) {
      ref_ = ref_tree_api().create(&(*cloud)[0].x, (int)cloud->size());
      return;
    }
    const int n = (int)cloud->size();
    order_.resize(n);
    std::iota(order_.begin(), order_.end(), 0);
    nodes_.clear();
    nodes_.reserve(n / 4 + 16);
    if (n > 0) build(0, n);
  }
  // nearestKSearch (call sites feature_extract.hpp:155,293,406,570,666,813).  Writes min(K, size) hits and
  // returns that count.
  int nearestKSearch(float qx, float qy, float qz, int K, int *idx, float *sqd) const {
    if (ref_) {
      const int got = ref_tree_api().knn(ref_, qx, qy, qz, K, idx, sqd);
      for (int i = 1; i < got; i++)  // equal distances: ascending index (insertion sort over tie runs)
        for (int j = i; j > 0 && sqd[j - 1] == sqd[j] && idx[j - 1] > idx[j]; j--) std::swap(idx[j - 1], idx[j]);
      return got;
    }
    Best best;
    best.K = K < kMaxK ? K : kMaxK;
    if (!nodes_.empty()) search(0, qx, qy, qz, best);
    for (int i = 0; i < best.n; i++) idx[i] = best.h[i].idx, sqd[i] = best.h[i].d2;
    return best.n;
  }
  size_t size() const { return cloud_ ? cloud_->size() : 0; }

 private:
  struct Node {
    int lo, hi;        // range in order_
    int left, right;   // children or -1
    float bmin[3], bmax[3];
  };
  const Cloud *cloud_ = nullptr;
  void *ref_ = nullptr;
  std::vector<int> order_;
  std::vector<Node> nodes_;
  static constexpr int kLeaf = 15;
  static constexpr int kMaxK = 64;
  struct Best {  // fixed-capacity sorted result set (no heap traffic per query)
    KnnHit h[kMaxK + 1];
    int n = 0, K = 0;
    void offer(const KnnHit &x) {
      if (n == K && !hit_less(x, h[n - 1])) return;
      int i = n < K ? n : K - 1;
      while (i > 0 && hit_less(x, h[i - 1])) h[i] = h[i - 1], i--;
      h[i] = x;
      if (n < K) n++;
    }
  };

  static float coord(const PointI &p, int d) { return d == 0 ? p.x : (d == 1 ? p.y : p.z); }

  int build(int lo, int hi) {
    Node nd;
    nd.lo = lo, nd.hi = hi, nd.left = nd.right = -1;
    for (int d = 0; d < 3; d++) nd.bmin[d] = 3.4e38f, nd.bmax[d] = -3.4e38f;
    for (int i = lo; i < hi; i++) {
      const PointI &p = (*cloud_)[order_[i]];
      for (int d = 0; d < 3; d++) {
        float c = coord(p, d);
        nd.bmin[d] = std::min(nd.bmin[d], c);
        nd.bmax[d] = std::max(nd.bmax[d], c);
      }
    }
    int id = (int)nodes_.size();
    nodes_.push_back(nd);
    if (hi - lo > kLeaf) {
      int dim = 0;
      float ext = nd.bmax[0] - nd.bmin[0];
      for (int d = 1; d < 3; d++)
        if (nd.bmax[d] - nd.bmin[d] > ext) ext = nd.bmax[d] - nd.bmin[d], dim = d;
      int mid = (lo + hi) / 2;
      std::nth_element(order_.begin() + lo, order_.begin() + mid, order_.begin() + hi, [&](int a, int b) {
        float ca = coord((*cloud_)[a], dim), cb = coord((*cloud_)[b], dim);
        return ca < cb || (ca == cb && a < b);
      });
      int l = build(lo, mid);
      int r = build(mid, hi);
      nodes_[id].left = l;
      nodes_[id].right = r;
    }
    return id;
  }
  // lower bound (double, slightly deflated so float rounding of point distances cannot defeat it)
  double boxdist(const Node &nd, float qx, float qy, float qz) const {
    const float q[3] = {qx, qy, qz};
    double s = 0;
    for (int d = 0; d < 3; d++) {
      double e = 0;
      if (q[d] < nd.bmin[d]) e = (double)nd.bmin[d] - q[d];
      else if (q[d] > nd.bmax[d]) e = (double)q[d] - nd.bmax[d];
      s += e * e;
    }
    return s * (1.0 - 1e-5);
  }
  void search(int id, float qx, float qy, float qz, Best &best) const {
    const Node &nd = nodes_[id];
    if (best.n == best.K && boxdist(nd, qx, qy, qz) > (double)best.h[best.n - 1].d2) return;
    if (nd.left < 0) {
      for (int i = nd.lo; i < nd.hi; i++) best.offer(KnnHit{dist2f((*cloud_)[order_[i]], qx, qy, qz), order_[i]});
      return;
    }
    double dl = boxdist(nodes_[nd.left], qx, qy, qz), dr = boxdist(nodes_[nd.right], qx, qy, qz);
    if (dl <= dr) {
      search(nd.left, qx, qy, qz, best);
      search(nd.right, qx, qy, qz, best);
    } else {
      search(nd.right, qx, qy, qz, best);
      search(nd.left, qx, qy, qz, best);
    }
  }
};

// brute-force kNN for validating the tree (tests only)
inline int knn_brute(const Cloud &c, float qx, float qy, float qz, int K, int *idx, float *sqd) {
  std::vector<KnnHit> all(c.size());
  for (size_t i = 0; i < c.size(); i++) all[i] = KnnHit{dist2f(c[i], qx, qy, qz), (int)i};
  int k = std::min<int>(K, (int)c.size());
  std::partial_sort(all.begin(), all.begin() + k, all.end(), hit_less);
  for (int i = 0; i < k; i++) idx[i] = all[i].idx, sqd[i] = all[i].d2;
  return k;
}

// ============================================================================ pcl::VoxelGrid<PointXYZI>
// PCL 1.8.0 VoxelGrid::applyFilter with downsample_all_data_=true, min_points_per_voxel_=0, no
// filter field.  (In-tree mirror: voxel_grid_covariance_mloam_impl.hpp:84-116 bbox/divisions,
// :206-222 voxel index, :227 sort, :239-250 run detection.)  Centroid = arithmetic mean of
// x,y,z,intensity accumulated in float in sorted order, divided by the float count.
// std::sort on (idx) is unstable in the reference; the restatement orders equal idx by point index.
// Returns false (and copies the input) when the index space would overflow int32 (:92-101).
inline bool voxel_grid(const Cloud &in, float leaf, Cloud &out, bool intensity_last = false) {
  out.clear();
  if (in.empty()) return true;
  const float inv = 1.0f / leaf;
  float mn[3] = {3.4028235e38f, 3.4028235e38f, 3.4028235e38f}, mx[3] = {-3.4028235e38f, -3.4028235e38f, -3.4028235e38f};
  for (const PointI &p : in) {
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    mn[0] = std::min(mn[0], p.x), mn[1] = std::min(mn[1], p.y), mn[2] = std::min(mn[2], p.z);
    mx[0] = std::max(mx[0], p.x), mx[1] = std::max(mx[1], p.y), mx[2] = std::max(mx[2], p.z);
  }
  int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1, dy = (int64_t)((mx[1] - mn[1]) * inv) + 1,
          dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > (int64_t)std::numeric_limits<int32_t>::max()) {
    out = in;
    return false;
  }
  int minb[3], maxb[3], divb[3];
  for (int d = 0; d < 3; d++) {
    minb[d] = (int)std::floor(mn[d] * inv);
    maxb[d] = (int)std::floor(mx[d] * inv);
    divb[d] = maxb[d] - minb[d] + 1;
  }
  const int mul[3] = {1, divb[0], divb[0] * divb[1]};
  struct Item {
    unsigned idx;
    int pt;
  };
  std::vector<Item> items;
  items.reserve(in.size());
  for (int i = 0; i < (int)in.size(); i++) {
    const PointI &p = in[i];
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    int i0 = (int)(std::floor(p.x * inv) - (float)minb[0]);
    int i1 = (int)(std::floor(p.y * inv) - (float)minb[1]);
    int i2 = (int)(std::floor(p.z * inv) - (float)minb[2]);
    items.push_back(Item{(unsigned)(i0 * mul[0] + i1 * mul[1] + i2 * mul[2]), i});
  }
  std::stable_sort(items.begin(), items.end(), [](const Item &a, const Item &b) { return a.idx < b.idx; });
  size_t k = 0;
  while (k < items.size()) {
    size_t e = k + 1;
    while (e < items.size() && items[e].idx == items[k].idx) e++;
    float sx = 0, sy = 0, sz = 0, si = 0;
    for (size_t j = k; j < e; j++) {
      const PointI &p = in[items[j].pt];
      sx = sx + p.x, sy = sy + p.y, sz = sz + p.z, si = si + p.intensity;
    }
    float cnt = (float)(e - k);
    PointI o{sx / cnt, sy / cnt, sz / cnt, si / cnt};
    // VoxelGridCovarianceMLOAM on plain PointI keeps the LAST point's intensity
    // (voxel_grid_covariance_mloam_impl.hpp:417-418,428)
    if (intensity_last) o.intensity = in[items[e - 1].pt].intensity;
    out.push_back(o);
    k = e;
  }
  return true;
}

// ============================================================================ extractCloud
// ScanInfo, estimator/src/estimator/parameters.h:192-207.  start = ring_begin+5, end = ring_end-6
// (image_segmenter.hpp:385-387).
struct ScanInfo {
  std::vector<int> scan_start_ind, scan_end_ind;
};
// cloudFeature (parameters.h:161): keys laser_cloud / corner_points_sharp / corner_points_less_sharp /
// surf_points_flat / surf_points_less_flat (feature_extract.cpp:281-285)
struct CloudFeature {
  Cloud laser_cloud, corner_points_sharp, corner_points_less_sharp, surf_points_flat, surf_points_less_flat;
  std::vector<float> curvature;  // exposed for kernel-level parity tests
  std::vector<int> label;
};

// feature_extract.cpp:118-297
inline void extract_cloud(const Cloud &laser_cloud, const ScanInfo &scan_info, int n_scans, CloudFeature &out) {
  const int cloud_size = (int)laser_cloud.size();
  std::vector<float> curv(cloud_size, 0.0f);
  std::vector<int> sort_ind(cloud_size, 0), picked(cloud_size, 0), label(cloud_size, 0);
  const Cloud &P = laser_cloud;
  // :133-142 — 11-tap stencil over the flat ring-major array, evaluated left to right in float
  for (int i = 5; i < cloud_size - 5; i++) {
    float dx = P[i - 5].x + P[i - 4].x + P[i - 3].x + P[i - 2].x + P[i - 1].x - 10 * P[i].x + P[i + 1].x + P[i + 2].x +
               P[i + 3].x + P[i + 4].x + P[i + 5].x;
    float dy = P[i - 5].y + P[i - 4].y + P[i - 3].y + P[i - 2].y + P[i - 1].y - 10 * P[i].y + P[i + 1].y + P[i + 2].y +
               P[i + 3].y + P[i + 4].y + P[i + 5].y;
    float dz = P[i - 5].z + P[i - 4].z + P[i - 3].z + P[i - 2].z + P[i - 1].z - 10 * P[i].z + P[i + 1].z + P[i + 2].z +
               P[i + 3].z + P[i + 4].z + P[i + 5].z;
    curv[i] = dx * dx + dy * dy + dz * dz;
    sort_ind[i] = i;
  }
  out = CloudFeature();
  auto gap2 = [&](int a, int b) {
    float ex = P[a].x - P[b].x, ey = P[a].y - P[b].y, ez = P[a].z - P[b].z;
    return ex * ex + ey * ey + ez * ez;
  };
  auto suppress = [&](int ind) {  // :192-213 / :233-254
    for (int l = 1; l <= 5; l++) {
      if (gap2(ind + l, ind + l - 1) > 0.05) break;  // float vs double literal, as in the reference
      picked[ind + l] = 1;
    }
    for (int l = -1; l >= -5; l--) {
      if (gap2(ind + l, ind + l + 1) > 0.05) break;
      picked[ind + l] = 1;
    }
  };
  for (int i = 0; i < n_scans; i++) {
    const int s = scan_info.scan_start_ind[i], e = scan_info.scan_end_ind[i];
    if (e - s < 6) continue;  // :155
    Cloud less_flat_scan;
    for (int j = 0; j < 6; j++) {
      const int sp = s + (e - s) * j / 6;            // :160
      const int ep = s + (e - s) * (j + 1) / 6 - 1;  // :161
      // :162 std::sort by curvature; restated as stable (curvature, index) — ties are implementation-ordered
      std::stable_sort(sort_ind.begin() + sp, sort_ind.begin() + ep + 1,
                       [&](int a, int b) { return curv[a] < curv[b]; });
      // :165-215 edge pick, largest curvature first
      int largest = 0;
      for (int k = ep; k >= sp; k--) {
        int ind = sort_ind[k];
        if (picked[ind] == 0 && curv[ind] > 0.1) {
          largest++;
          if (largest <= 2) {
            label[ind] = 2;
            out.corner_points_sharp.push_back(P[ind]);
            out.corner_points_less_sharp.push_back(P[ind]);
          } else if (largest <= 20) {
            label[ind] = 1;
            out.corner_points_less_sharp.push_back(P[ind]);
          } else {
            break;
          }
          picked[ind] = 1;
          suppress(ind);
        }
      }
      // :218-256 flat pick, smallest curvature first; note the break precedes the marking (:227-231)
      int smallest = 0;
      for (int k = sp; k <= ep; k++) {
        int ind = sort_ind[k];
        if (picked[ind] == 0 && curv[ind] < 0.1) {
          label[ind] = -1;
          out.surf_points_flat.push_back(P[ind]);
          smallest++;
          if (smallest >= 4) break;
          picked[ind] = 1;
          suppress(ind);
        }
      }
      // :258-264 — k indexes the cloud
      for (int k = sp; k <= ep; k++)
        if (label[k] <= 0) less_flat_scan.push_back(P[k]);
    }
    // :266-271
    Cloud ds;
    voxel_grid(less_flat_scan, 0.2f, ds);
    out.surf_points_less_flat.insert(out.surf_points_less_flat.end(), ds.begin(), ds.end());
  }
  out.laser_cloud = laser_cloud;
  out.curvature = curv;
  out.label = label;
}

}  // namespace orc
