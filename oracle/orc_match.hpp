// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
// Correspondence search: FeatureExtract::match{Corner,Surf}From{Scan,Map} and the per-point
// variants, estimator/src/featureExtract/feature_extract.hpp:131-883.
#pragma once
#include "orc_cloud.hpp"

namespace orc {

// PointPlaneFeature, estimator/src/estimator/parameters.h:163-175.  coeffs: 4 used for 's' (n,d),
// 6 for 'c' ([X1;X2]).  point = query point in the SENSOR frame (double), coeffs in the map frame.
struct Feature {
  size_t idx = 0;
  size_t laser_idx = 0;
  V3 point{0, 0, 0};
  double coeffs[6] = {0, 0, 0, 0, 0, 0};
  char type = 'n';
  int nn[16];  // neighbour indices (oracle-only, for kernel-level parity tests)
};

// Globals of parameters.h:45-133 that steer the path, as one POD.
struct MatchParams {
  float distance_sq_threshold = 25.0f;  // DISTANCE_SQ_THRESHOLD  (config yaml :103)
  float nearby_scan = 2.5f;             // NEARBY_SCAN            (:104)
  float min_match_sq_dis = 1.0f;        // MIN_MATCH_SQ_DIS       (:110)
  float min_plane_dis = 0.2f;           // MIN_PLANE_DIS          (:111)
};

// pointAssociateToMap, utility.h:103-117: double math, float store.  TransformToStart with
// b_distortion=false (utility.h:55-77, s=1: slerp(1,q)=q) is the same map.
inline PointI associate(const PointI &pi, const Pose &pose) {
  V3 v = qrot(pose.q, V3{(double)pi.x, (double)pi.y, (double)pi.z}) + pose.t;
  return PointI{(float)v.x, (float)v.y, (float)v.z, pi.intensity};
}

// FOV gate, feature_extract.hpp:696-715 (identical at :434-458, :599-618, :842-861)
inline bool in_laser_fov(const PointI &sel, const Pose &pose) {
  PointI z{0.0f, 0.0f, 10.0f, 0.0f};
  PointI zt = associate(z, pose);
  // pose_local.t_(k) - point_sel.x is double - float -> sqrSum<double>, narrowed on assignment to float
  const double ex = pose.t.x - (double)sel.x, ey = pose.t.y - (double)sel.y, ez = pose.t.z - (double)sel.z;
  float s1 = (float)(ex * ex + ey * ey + ez * ez);
  float s2 = sqrSumf(zt.x - sel.x, zt.y - sel.y, zt.z - sel.z);
  float check1 = 100.0f + s1 - s2 - 10.0f * std::sqrt(3.0f) * std::sqrt(s1);
  float check2 = 100.0f + s1 - s2 + 10.0f * std::sqrt(3.0f) * std::sqrt(s1);
  return check1 < 0 && check2 > 0;
}

// matchCornerPointFromMap, feature_extract.hpp:645-788 (batch form :378-538 is the same per point)
inline bool match_corner_point_from_map(const KdTree &tree, const Cloud &map, const PointI &ori, const Pose &pose,
                                        Feature &f, size_t idx, int n_neigh, bool check_fov, const MatchParams &mp) {
  int nn[16];
  float sq[16];
  PointI sel = associate(ori, pose);
  int got = tree.nearestKSearch(sel.x, sel.y, sel.z, n_neigh, nn, sq);
  if (got < n_neigh) return false;  // FLANN leaves the tail untouched; the reference never has maps that small
  if (!(sq[n_neigh - 1] < mp.min_match_sq_dis)) return false;
  float cx = 0, cy = 0, cz = 0;
  for (int j = 0; j < n_neigh; j++) cx = cx + map[nn[j]].x, cy = cy + map[nn[j]].y, cz = cz + map[nn[j]].z;
  // center /= (1.0 * num_neighbors): Eigen casts the double scalar to float, true division
  const float kf = (float)(1.0 * n_neigh);
  cx = cx / kf, cy = cy / kf, cz = cz / kf;
  float C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < n_neigh; j++) {
    float a = map[nn[j]].x - cx, b = map[nn[j]].y - cy, c = map[nn[j]].z - cz;
    C[0] = C[0] + a * a, C[1] = C[1] + a * b, C[2] = C[2] + a * c;
    C[4] = C[4] + b * b, C[5] = C[5] + b * c, C[8] = C[8] + c * c;
  }
  C[3] = C[1], C[6] = C[2], C[7] = C[5];
  float w[3], V[9];
  eig3f(C, w, V);
  if (!(w[2] > 3 * w[1])) return false;  // :693
  if (check_fov && !in_laser_fov(sel, pose)) return false;
  const float ux = V[0 * 3 + 2], uy = V[1 * 3 + 2], uz = V[2 * 3 + 2];
  // X1 = 0.1*u + c, X2 = -0.1*u + c in float (:768-769)
  const float k = 0.1f;
  f.coeffs[0] = (double)(k * ux + cx), f.coeffs[1] = (double)(k * uy + cy), f.coeffs[2] = (double)(k * uz + cz);
  f.coeffs[3] = (double)(-k * ux + cx), f.coeffs[4] = (double)(-k * uy + cy), f.coeffs[5] = (double)(-k * uz + cz);
  f.idx = idx;
  f.point = V3{(double)ori.x, (double)ori.y, (double)ori.z};
  f.laser_idx = (size_t)ori.intensity;
  f.type = 'c';
  for (int j = 0; j < n_neigh; j++) f.nn[j] = nn[j];
  return true;
}

// matchSurfPointFromMap, feature_extract.hpp:790-883 (batch form :541-643)
inline bool match_surf_point_from_map(const KdTree &tree, const Cloud &map, const PointI &ori, const Pose &pose,
                                      Feature &f, size_t idx, int n_neigh, bool check_fov, const MatchParams &mp) {
  int nn[16];
  float sq[16];
  PointI sel = associate(ori, pose);
  int got = tree.nearestKSearch(sel.x, sel.y, sel.z, n_neigh, nn, sq);
  if (got < n_neigh) return false;
  if (!(sq[n_neigh - 1] < mp.min_match_sq_dis)) return false;
  float A[16 * 3];
  for (int j = 0; j < n_neigh; j++) A[j * 3 + 0] = map[nn[j]].x, A[j * 3 + 1] = map[nn[j]].y, A[j * 3 + 2] = map[nn[j]].z;
  float n[3];
  if (!lsq_plane_f(A, n_neigh, n)) return false;
  float nrm = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
  float d = 1 / nrm;  // negative_OA_dot_norm (:824)
  n[0] = n[0] / nrm, n[1] = n[1] / nrm, n[2] = n[2] / nrm;
  for (int j = 0; j < n_neigh; j++)  // :828-837
    if (std::fabs(n[0] * map[nn[j]].x + n[1] * map[nn[j]].y + n[2] * map[nn[j]].z + d) > mp.min_plane_dis) return false;
  if (check_fov && !in_laser_fov(sel, pose)) return false;
  f.coeffs[0] = n[0], f.coeffs[1] = n[1], f.coeffs[2] = n[2], f.coeffs[3] = d;
  f.coeffs[4] = f.coeffs[5] = 0;
  f.idx = idx;
  f.point = V3{(double)ori.x, (double)ori.y, (double)ori.z};
  f.laser_idx = (size_t)ori.intensity;
  f.type = 's';
  for (int j = 0; j < n_neigh; j++) f.nn[j] = nn[j];
  return true;
}

// matchCornerFromMap / matchSurfFromMap: compacted in query order (:378-643)
inline void match_from_map(char type, const KdTree &tree, const Cloud &map, const Cloud &data, const Pose &pose,
                           std::vector<Feature> &features, int n_neigh, bool check_fov, const MatchParams &mp) {
  features.clear();
  const int n = (int)data.size();
  // The reference loop is serial (the mapper has no OpenMP).  With more than one OpenMP thread
  // configured (bench "all host cores" arm only) queries are matched in parallel and compacted in query order,
  // which yields the identical feature list.
  std::vector<unsigned char> ok(n, 0);
  std::vector<Feature> all(n);
#pragma omp parallel for schedule(dynamic, 256)
  for (int i = 0; i < n; i++) {
    ok[i] = (type == 'c' ? match_corner_point_from_map(tree, map, data[i], pose, all[i], i, n_neigh, check_fov, mp)
                         : match_surf_point_from_map(tree, map, data[i], pose, all[i], i, n_neigh, check_fov, mp))
                ? 1
                : 0;
  }
  for (int i = 0; i < n; i++)
    if (ok[i]) features.push_back(all[i]);
}

// matchCornerFromScan, feature_extract.hpp:131-271.  cloud_scan must be ring-sorted; int(intensity)=ring.
inline void match_corner_from_scan(const KdTree &tree, const Cloud &scan, const Cloud &data, const Pose &pose,
                                   std::vector<Feature> &features, const MatchParams &mp) {
  features.clear();
  for (size_t i = 0; i < data.size(); i++) {
    PointI sel = associate(data[i], pose);
    int j0;
    float d0;
    if (tree.nearestKSearch(sel.x, sel.y, sel.z, 1, &j0, &d0) < 1) continue;
    int closest = -1, ind2 = -1;
    if (d0 < mp.distance_sq_threshold) {
      closest = j0;
      const int ring = (int)scan[closest].intensity;
      float best2 = mp.distance_sq_threshold;
      for (int j = closest + 1; j < (int)scan.size(); j++) {  // :165-182
        if ((int)scan[j].intensity <= ring) continue;
        if ((int)scan[j].intensity > (ring + mp.nearby_scan)) break;  // int vs float compare -> float
        float d = sqrSumf(scan[j].x - sel.x, scan[j].y - sel.y, scan[j].z - sel.z);
        if (d < best2) best2 = d, ind2 = j;
      }
      for (int j = closest - 1; j >= 0; j--) {  // :185-202
        if ((int)scan[j].intensity >= ring) continue;
        if ((int)scan[j].intensity < (ring - mp.nearby_scan)) break;
        float d = sqrSumf(scan[j].x - sel.x, scan[j].y - sel.y, scan[j].z - sel.z);
        if (d < best2) best2 = d, ind2 = j;
      }
    }
    if (ind2 >= 0) {  // :205-268
      Feature f;
      f.coeffs[0] = scan[closest].x, f.coeffs[1] = scan[closest].y, f.coeffs[2] = scan[closest].z;
      f.coeffs[3] = scan[ind2].x, f.coeffs[4] = scan[ind2].y, f.coeffs[5] = scan[ind2].z;
      f.idx = i;
      f.point = V3{(double)data[i].x, (double)data[i].y, (double)data[i].z};
      f.type = 'c';  // the reference leaves type_ at its default 'n' here; kept as 'c' for bookkeeping only
      f.nn[0] = closest, f.nn[1] = ind2;
      features.push_back(f);
    }
  }
}

// matchSurfFromScan, feature_extract.hpp:273-376
inline void match_surf_from_scan(const KdTree &tree, const Cloud &scan, const Cloud &data, const Pose &pose,
                                 std::vector<Feature> &features, const MatchParams &mp) {
  features.clear();
  for (size_t i = 0; i < data.size(); i++) {
    PointI sel = associate(data[i], pose);
    int j0;
    float d0;
    if (tree.nearestKSearch(sel.x, sel.y, sel.z, 1, &j0, &d0) < 1) continue;
    if (!(d0 < mp.distance_sq_threshold)) continue;
    const int closest = j0;
    const int ring = (int)scan[closest].intensity;
    int ind2 = -1, ind3 = -1;
    float best2 = mp.distance_sq_threshold, best3 = mp.distance_sq_threshold;
    for (int j = closest + 1; j < (int)scan.size(); j++) {  // :304-324
      if ((int)scan[j].intensity > (ring + mp.nearby_scan)) break;
      float d = sqrSumf(scan[j].x - sel.x, scan[j].y - sel.y, scan[j].z - sel.z);
      if ((int)scan[j].intensity <= ring && d < best2) best2 = d, ind2 = j;
      else if ((int)scan[j].intensity > ring && d < best3) best3 = d, ind3 = j;
    }
    for (int j = closest - 1; j >= 0; j--) {  // :327-347
      if ((int)scan[j].intensity < (ring - mp.nearby_scan)) break;
      float d = sqrSumf(scan[j].x - sel.x, scan[j].y - sel.y, scan[j].z - sel.z);
      if ((int)scan[j].intensity >= ring && d < best2) best2 = d, ind2 = j;
      else if ((int)scan[j].intensity < ring && d < best3) best3 = d, ind3 = j;
    }
    if (ind2 >= 0 && ind3 >= 0) {  // :349-373, Vector3f geometry
      const PointI &J = scan[closest], &L = scan[ind2], &M = scan[ind3];
      float ax = J.x - L.x, ay = J.y - L.y, az = J.z - L.z;
      float bx = J.x - M.x, by = J.y - M.y, bz = J.z - M.z;
      float wx = ay * bz - az * by, wy = az * bx - ax * bz, wz = ax * by - ay * bx;
      float nrm = std::sqrt(wx * wx + wy * wy + wz * wz);
      wx = wx / nrm, wy = wy / nrm, wz = wz / nrm;
      float d = -(wx * J.x + wy * J.y + wz * J.z);
      Feature f;
      f.coeffs[0] = wx, f.coeffs[1] = wy, f.coeffs[2] = wz, f.coeffs[3] = d;
      f.idx = i;
      f.point = V3{(double)data[i].x, (double)data[i].y, (double)data[i].z};
      f.type = 's';
      f.nn[0] = closest, f.nn[1] = ind2, f.nn[2] = ind3;
      features.push_back(f);
    }
  }
}

}  // namespace orc
