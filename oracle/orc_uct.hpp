// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
// Submap assembly with uncertainty:
//   compoundPoseWithCov            estimator/src/lidarMapper/associate_uct.hpp:9-88 (Barfoot's 4th-order compounding, method 2)
//   cloudUCTAssociateToMap         estimator/src/lidarMapper/lidar_mapper_keyframe.cpp:1116-1158
//   VoxelGridCovarianceMLOAM<PointIWithCov>::applyFilter with the covariance-weighted merge
//                                  mloam_pcl/include/mloam_pcl/voxel_grid_covariance_mloam_impl.hpp:69-457 (merge :293-333)
// Pose covariance ordering as the reference's: [translation (3) | rotation (3)] (pointToFS, associate_uct.hpp:149-156).
// PointXYZIWithCov (point_with_cov.hpp:45-53) is carried as PointI + float cov_vec[6] (xx xy xz yy yz zz) + float cov_trace.
#pragma once
#include "orc_pipeline.hpp"

namespace orc {

struct M6 {
  double m[36];
  double &operator()(int r, int c) { return m[r * 6 + c]; }
  double operator()(int r, int c) const { return m[r * 6 + c]; }
};
inline M6 m6_zero() {
  M6 z;
  for (double &v : z.m) v = 0.0;
  return z;
}
inline M6 m6_mul(const M6 &A, const M6 &B) {
  M6 C = m6_zero();
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) {
      double s = 0;
      for (int k = 0; k < 6; k++) s += A(i, k) * B(k, j);
      C(i, j) = s;
    }
  return C;
}
inline M6 m6_T(const M6 &A) {
  M6 T;
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) T(i, j) = A(j, i);
  return T;
}
inline M3 m6_block(const M6 &A, int r0, int c0) {
  M3 B;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) B(i, j) = A(r0 + i, c0 + j);
  return B;
}
inline void m6_set(M6 &A, int r0, int c0, const M3 &B) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) A(r0 + i, c0 + j) = B(i, j);
}
inline M3 m3_add(const M3 &A, const M3 &B) {
  M3 C;
  for (int i = 0; i < 9; i++) C.m[i] = A.m[i] + B.m[i];
  return C;
}
// associate_uct.hpp:18-28
inline M3 covop1(const M3 &B) {
  const double tr = B(0, 0) + B(1, 1) + B(2, 2);
  M3 A = B;
  for (int i = 0; i < 3; i++) A(i, i) = -tr + B(i, i);
  return A;
}
inline M3 covop2(const M3 &B, const M3 &C) { return m3_add(matmul(covop1(B), covop1(C)), covop1(matmul(C, B))); }

// associate_uct.hpp:30-88 (method == 2).  pose_cp = pose_1 * pose_2.
inline void compound_pose_with_cov(const Pose &p1, const M6 &cov1, const Pose &p2, const M6 &cov2, Pose &pcp, M6 &ccp) {
  pcp = Pose{qmul(p1.q, p2.q), qrot(p1.q, p2.t) + p1.t};  // Quaterniond product is not re-normalised here (:37-38)
  const M3 R = qmat(p1.q);
  M6 Ad = m6_zero();  // adjointMatrix :9-16
  m6_set(Ad, 0, 0, R), m6_set(Ad, 0, 3, matmul(skew(p1.t), R)), m6_set(Ad, 3, 3, R);
  const M6 c2p = m6_mul(m6_mul(Ad, cov2), m6_T(Ad));
  const M3 c1rr = m6_block(cov1, 0, 0), c1rp = m6_block(cov1, 0, 3), c1pp = m6_block(cov1, 3, 3);
  const M3 c2rr = m6_block(c2p, 0, 0), c2rp = m6_block(c2p, 0, 3), c2pp = m6_block(c2p, 3, 3);
  M6 A1 = m6_zero(), A2 = m6_zero(), B = m6_zero();
  m6_set(A1, 0, 0, covop1(c1pp)), m6_set(A1, 0, 3, covop1(m3_add(c1rp, transpose(c1rp)))), m6_set(A1, 3, 3, covop1(c1pp));
  m6_set(A2, 0, 0, covop1(c2pp)), m6_set(A2, 0, 3, covop1(m3_add(c2rp, transpose(c2rp)))), m6_set(A2, 3, 3, covop1(c2pp));
  const M3 Brr = m3_add(m3_add(m3_add(covop2(c1pp, c2rr), covop2(transpose(c1rp), c2rp)), covop2(c1rp, transpose(c2rp))), covop2(c1rr, c2pp));
  const M3 Brp = m3_add(covop2(c1pp, transpose(c2rp)), covop2(transpose(c1rp), c2pp));
  const M3 Bpp = covop2(c1pp, c2pp);
  m6_set(B, 0, 0, Brr), m6_set(B, 0, 3, Brp), m6_set(B, 3, 0, transpose(Brp)), m6_set(B, 3, 3, Bpp);
  const M6 t1 = m6_mul(A1, c2p), t2 = m6_mul(c2p, m6_T(A1)), t3 = m6_mul(A2, cov1), t4 = m6_mul(cov1, m6_T(A2));
  for (int i = 0; i < 36; i++) ccp.m[i] = cov1.m[i] + c2p.m[i] + (((t1.m[i] + t2.m[i]) + t3.m[i]) + t4.m[i]) / 12 + B.m[i] / 4;
}

struct CovCloud {  // PointICovCloud
  Cloud pts;
  std::vector<float> cov6;   // 6 per point
  std::vector<float> trace;  // cov_trace
};

// evalPointUncertainty in double (associate_uct.hpp:192-214), 3x3 symmetric out
inline void eval_point_uncertainty_d(const PointI &pi, const Pose &pose, const double cov_pose[36], const double cov_meas[9], double C[3][3]) {
  const V3 tp = qrot(pose.q, V3{(double)pi.x, (double)pi.y, (double)pi.z}) + pose.t;
  const M3 R = qmat(pose.q);
  const M3 S = skew(tp);
  double G[3][9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) G[i][j] = (i == j) ? 1.0 : 0.0, G[i][3 + j] = -S(i, j), G[i][6 + j] = R(i, j);
  double Sig[9][9] = {{0}};
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) Sig[i][j] = cov_pose[i * 6 + j];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Sig[6 + i][6 + j] = cov_meas[i * 3 + j];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int a = 0; a < 9; a++) {
        double t = 0;
        for (int b = 0; b < 9; b++) t += Sig[a][b] * G[j][b];
        s += G[i][a] * t;
      }
      C[i][j] = s;
    }
}

// cloudUCTAssociateToMap, lidar_mapper_keyframe.cpp:1116-1158.  pose_compound[n] / cov_compound[n] = compoundPoseWithCov(pose_global, ext[n]).
inline void cloud_uct_associate(const Cloud &cloud_local, const Pose &pose_global, const std::vector<Pose> &pose_ext, const std::vector<Pose> &pose_compound,
                                const std::vector<M6> &cov_compound, const double cov_meas[9], bool with_ua, double trace_threshold, CovCloud &out) {
  out.pts.clear(), out.cov6.clear(), out.trace.clear();
  for (const PointI &po : cloud_local) {
    const int ind = (int)po.intensity;
    double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    if (with_ua) {
      const PointI sel = associate(po, pose_inv(pose_ext[ind]));                                   // :1147
      eval_point_uncertainty_d(sel, pose_compound[ind], cov_compound[ind].m, cov_meas, C);         // :1148
      if (C[0][0] + C[1][1] + C[2][2] > trace_threshold) continue;                                 // :1150
    }
    const PointI pc = associate(po, pose_global);                                                  // :1152
    out.pts.push_back(pc);                                                                         // updateCov, point_with_cov.hpp:187-196
    const double c6[6] = {C[0][0], C[0][1], C[0][2], C[1][1], C[1][2], C[2][2]};
    for (double v : c6) out.cov6.push_back((float)v);
    out.trace.push_back((float)(C[0][0] + C[1][1] + C[2][2]));
  }
}

// VoxelGridCovarianceMLOAM<PointIWithCov>::applyFilter (no field filter, downsample_all_data, min_points_per_voxel 0).
// Indexing as pcl::VoxelGrid (see voxel_grid() in orc_cloud.hpp); the per-voxel merge is :293-333.  `abs` at :306 is taken as the float
// overload.  Returns false (input copied) on the int32 index overflow path.
inline bool voxel_grid_cov(const CovCloud &in, float leaf, float trace_threshold, CovCloud &out) {
  out.pts.clear(), out.cov6.clear(), out.trace.clear();
  const int n = (int)in.pts.size();
  if (n == 0) return true;
  const float inv = 1.0f / leaf;
  float mn[3] = {3.4028235e38f, 3.4028235e38f, 3.4028235e38f}, mx[3] = {-3.4028235e38f, -3.4028235e38f, -3.4028235e38f};
  for (const PointI &p : in.pts) {
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    mn[0] = std::min(mn[0], p.x), mn[1] = std::min(mn[1], p.y), mn[2] = std::min(mn[2], p.z);
    mx[0] = std::max(mx[0], p.x), mx[1] = std::max(mx[1], p.y), mx[2] = std::max(mx[2], p.z);
  }
  const int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1, dy = (int64_t)((mx[1] - mn[1]) * inv) + 1, dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > (int64_t)std::numeric_limits<int32_t>::max()) {
    out = in;
    return false;
  }
  int minb[3], divb[3];
  for (int d = 0; d < 3; d++) minb[d] = (int)std::floor(mn[d] * inv), divb[d] = (int)std::floor(mx[d] * inv) - minb[d] + 1;
  const int mul[3] = {1, divb[0], divb[0] * divb[1]};
  struct Item {
    unsigned idx;
    int pt;
  };
  std::vector<Item> items;
  for (int i = 0; i < n; i++) {
    const PointI &p = in.pts[i];
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    const int i0 = (int)(std::floor(p.x * inv) - (float)minb[0]), i1 = (int)(std::floor(p.y * inv) - (float)minb[1]),
              i2 = (int)(std::floor(p.z * inv) - (float)minb[2]);
    items.push_back(Item{(unsigned)(i0 * mul[0] + i1 * mul[1] + i2 * mul[2]), i});
  }
  std::stable_sort(items.begin(), items.end(), [](const Item &a, const Item &b) { return a.idx < b.idx; });
  size_t k = 0;
  while (k < items.size()) {
    size_t e = k + 1;
    while (e < items.size() && items[e].idx == items[k].idx) e++;
    float mu[3] = {0, 0, 0}, ity = 0, cov[7] = {0, 0, 0, 0, 0, 0, 0}, weight_total = 0, w_max = 0;
    for (size_t j = k; j < e; j++) {
      const int q = items[j].pt;
      const PointI &p = in.pts[q];
      const float *c6 = &in.cov6[(size_t)q * 6];
      const float t7[7] = {c6[0], c6[1], c6[2], c6[3], c6[4], c6[5], in.trace[q]};
      const float tr = c6[0] + c6[3] + c6[5];                     // temporary[4] + temporary[7] + temporary[9]
      if (std::fabs(tr) >= trace_threshold) continue;             // :306
      const float w = trace_threshold - tr;                      // :311
      mu[0] = mu[0] + w * p.x, mu[1] = mu[1] + w * p.y, mu[2] = mu[2] + w * p.z;
      ity = w > w_max ? p.intensity : ity;
      w_max = w > w_max ? w : w_max;
      const float ww = w * w;
      for (int a = 0; a < 7; a++) cov[a] = cov[a] + ww * t7[a];
      weight_total = weight_total + w;
    }
    if (weight_total == 0) weight_total = 1.0f;
    for (float &v : mu) v = v / weight_total;
    const float w2 = weight_total * weight_total;
    for (float &v : cov) v = v / w2;
    out.pts.push_back(PointI{mu[0], mu[1], mu[2], ity});
    for (int a = 0; a < 6; a++) out.cov6.push_back(cov[a]);
    out.trace.push_back(cov[0] + cov[3] + cov[5]);               // centroid[10] = centroid[4] + centroid[7] + centroid[9]
    k = e;
  }
  return true;
}

}  // namespace orc
