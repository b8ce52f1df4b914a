// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
// Good-feature selection: ActiveFeatureSelection::goodFeatureMatching, estimator/src/lidarMapper/lidar_mapper.h:229-573
// (the odometry twin is Estimator::goodFeatureMatching, estimator.cpp:1347-1517), with evaluateFeatJacobianMatching
// (:130-174) and common::logDet(M, true) (mloam_common/.../math.hpp:172-202).
//
// What the reference leaves to chance is made explicit here so that the path can be checked at all:
//   * its std::mt19937 seeded from std::random_device becomes gf_rng (PCG32, seed given by the caller);
//   * its wall-clock cap (MAX_FEATURE_SELECT_TIME, 20 ms) is dropped — selection runs until the quota is met, the
//     candidate pool is empty or MAX_RANDOM_QUEUE_TIME (20) consecutive draws hit candidates already tried this round;
//   * std::priority_queue ties (equal log-det) resolve to the candidate pushed first;
//   * fps: the reference never terminates once every point has been visited without filling the quota (it relies on
//     the time cap); here it stops when all points have been visited (the condition commented out at :411-414).
// Matching and the Jacobian depend on the feature and the pose only, so evaluating them for every feature up front
// (what the GPU does) or lazily per drawn candidate (what the reference does) selects the same features.
#pragma once
#include <cstdint>

#include "orc_factors.hpp"
#include "orc_match.hpp"

namespace orc {

enum GfMethod { GF_WO = 0, GF_RND = 1, GF_FPS = 2, GF_GD = 3 };
constexpr int kGfMaxRandomQueue = 20;  // MAX_RANDOM_QUEUE_TIME, lidar_mapper.h:83

// PCG32 (XSH-RR), the RNG shared by the oracle and the GPU kernel
inline uint32_t gf_next(uint64_t &s) {
  const uint64_t old = s;
  s = old * 6364136223846793005ull + 1442695040888963407ull;
  const uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
  const uint32_t rot = (uint32_t)(old >> 59u);
  return (xorshifted >> rot) | (xorshifted << ((32u - rot) & 31u));
}
inline uint64_t gf_seed(uint64_t seed) {
  uint64_t s = seed * 0x9e3779b97f4a7c15ull + 0xda3e39cb94b95bdbull;
  gf_next(s);
  return s;
}
// uniform integer in [lo, hi] (geneRandUniform(lo, hi), random_generator.hpp:62-66), multiply-shift range reduction
inline size_t gf_uniform(uint64_t &s, size_t lo, size_t hi) {
  const uint64_t span = (uint64_t)(hi - lo) + 1ull;
  return lo + (size_t)(((uint64_t)gf_next(s) * span) >> 32);
}

inline void gf_add_outer(double H[36], const double *j) {
  for (int a = 0; a < 6; a++)
    for (int b = 0; b < 6; b++) H[a * 6 + b] += j[a] * j[b];
}
inline double gf_logdet_with(const double H[36], const double *j) {
  double M[36];
  for (int a = 0; a < 6; a++)
    for (int b = 0; b < 6; b++) M[a * 6 + b] = H[a * 6 + b] + j[a] * j[b];
  return logdet_chol(6, M);
}

// matched[i] / jaco[i*6..] for every feature; xyz = sensor-frame points (fps distances, :392-396).
// sel: selected feature indices in selection order; H: sub_mat_H, initialised to 1e-6 I by the caller's convention
// (:504, :519) and updated with J^T J of the selected features.
inline void good_feature_select(int method, double gf_ratio, uint64_t seed, int n, const unsigned char *matched, const double *jaco,
                                const float *xyz4, std::vector<int> &sel, double H[36]) {
  sel.clear();
  for (int i = 0; i < 36; i++) H[i] = (i % 7 == 0) ? 1e-6 : 0.0;
  const size_t num_all = (size_t)n;
  const size_t num_use = (size_t)(num_all * gf_ratio);  // :248
  uint64_t rng = gf_seed(seed);
  std::vector<size_t> pool(num_all);
  for (size_t i = 0; i < num_all; i++) pool[i] = i;
  std::vector<int> visited(num_all, -1);
  if (method == GF_WO) {  // :257-299
    for (size_t q = 0; q < num_all; q++)
      if (matched[q]) gf_add_outer(H, jaco + q * 6), sel.push_back((int)q);
    return;
  }
  if (method == GF_RND) {  // :300-346
    while (true) {
      if (sel.size() >= num_use || pool.empty()) break;
      const size_t j = gf_uniform(rng, 0, pool.size() - 1);
      const size_t q = pool[j];
      if (matched[q]) gf_add_outer(H, jaco + q * 6), sel.push_back((int)q);
      pool.erase(pool.begin() + j);
    }
    return;
  }
  if (method == GF_FPS) {  // :347-449
    if (num_all == 0) return;
    size_t k = gf_uniform(rng, 0, pool.size() - 1);
    std::vector<char> vis(num_all, 0);
    vis[k] = 1;
    size_t cnt_visited = 1;
    size_t old = k;
    if (matched[k]) sel.push_back((int)k);  // the start point is selected but never added to sub_mat_H (:375-379)
    std::vector<float> dist(num_all, 1e5f);
    while (true) {
      if (sel.size() >= num_use || cnt_visited >= num_all) break;
      float best_d = -1;
      size_t best_j = 1;
      for (size_t j = 0; j < num_all; j++) {
        if (vis[j]) continue;
        const float dx = xyz4[old * 4] - xyz4[j * 4], dy = xyz4[old * 4 + 1] - xyz4[j * 4 + 1], dz = xyz4[old * 4 + 2] - xyz4[j * 4 + 2];
        const float d = std::sqrt(dx * dx + dy * dy + dz * dz);
        const float d2 = std::min(d, dist[j]);
        dist[j] = d2;
        best_j = d2 > best_d ? j : best_j;
        best_d = d2 > best_d ? d2 : best_d;
      }
      const size_t q = best_j;
      old = q;
      vis[q] = 1;
      cnt_visited++;
      if (matched[q]) gf_add_outer(H, jaco + q * 6), sel.push_back((int)q);
    }
    return;
  }
  // gd_fix / gd_float (:450-556): stochastic greedy on log det(sub_mat_H + J^T J)
  int num_rnd_que = 0;
  while (true) {
    if (sel.size() >= num_use || pool.empty()) break;
    const size_t size_rnd_subset = (size_t)(1.0 * num_all / num_use);
    size_t heap_n = 0, best_idx = 0;
    double best_score = 0;
    while (true) {
      if (pool.empty()) break;
      num_rnd_que = 0;
      size_t j = 0;
      while (num_rnd_que < kGfMaxRandomQueue) {
        j = gf_uniform(rng, 0, pool.size() - 1);
        if (visited[j] < (int)sel.size()) {
          visited[j] = (int)sel.size();
          break;
        }
        num_rnd_que++;
      }
      if (num_rnd_que >= kGfMaxRandomQueue) break;
      const size_t q = pool[j];
      if (!matched[q]) {  // "not found constraints or outlier constraints" (:518-523)
        pool.erase(pool.begin() + j);
        visited.erase(visited.begin() + j);
        continue;
      }
      const double cur = gf_logdet_with(H, jaco + q * 6);
      if (heap_n == 0 || cur > best_score) best_score = cur, best_idx = q;  // ties: first pushed
      heap_n++;
      if (heap_n >= size_rnd_subset) {
        size_t pos = 0;
        while (pos < pool.size() && pool[pos] != best_idx) pos++;
        gf_add_outer(H, jaco + best_idx * 6);
        pool.erase(pool.begin() + pos);
        visited.erase(visited.begin() + pos);
        sel.push_back((int)best_idx);
        break;
      }
    }
    if (num_rnd_que >= kGfMaxRandomQueue) break;
  }
}

// goodFeatureMatching for one feature set: match every feature, evaluate its 1x6 Jacobian row
// (evaluateFeatJacobianMatching: the map factor's Jacobian with sqrt_info from the point's covariance trace, no loss
// scaling), then select.  cov_trace: per-point traces (extractCov) or null for the default trace.
inline void good_feature_matching(char type, const KdTree &tree, const Cloud &map, const Cloud &scan, const Pose &pose,
                                  const std::vector<double> *cov_trace, double default_trace, int method, double gf_ratio,
                                  uint64_t seed, int n_neigh, const MatchParams &mp, std::vector<Feature> &all,
                                  std::vector<unsigned char> &matched, std::vector<double> &jaco, std::vector<int> &sel, double H[36]) {
  const size_t n = scan.size();
  all.assign(n, Feature());
  matched.assign(n, 0);
  jaco.assign(n * 6, 0.0);
  double x[7];
  pose_to_param(pose, x);
  for (size_t i = 0; i < n; i++) {
    bool ok = type == 's' ? match_surf_point_from_map(tree, map, scan[i], pose, all[i], i, n_neigh, false, mp)
                          : match_corner_point_from_map(tree, map, scan[i], pose, all[i], i, n_neigh, false, mp);
    if (!ok) continue;
    matched[i] = 1;
    const double si = map_sqrt_info(cov_trace ? (*cov_trace)[i] : default_trace);
    double r, J[7];
    if (type == 's') plane_factor(all[i].point, all[i].coeffs, si, x, &r, J);
    else edge_factor(all[i].point, all[i].coeffs, si, x, &r, J);
    for (int k = 0; k < 6; k++) jaco[i * 6 + k] = J[k];
  }
  std::vector<float> xyz(n * 4);
  for (size_t i = 0; i < n; i++) xyz[i * 4] = scan[i].x, xyz[i * 4 + 1] = scan[i].y, xyz[i * 4 + 2] = scan[i].z, xyz[i * 4 + 3] = 0.f;
  good_feature_select(method, gf_ratio, seed, (int)n, matched.data(), jaco.data(), xyz.data(), sel, H);
}

// Estimator::goodFeatureMatching (estimator.cpp:1347-1517) with evaluateFeatJacobian (:1273-1345): pose_local = pivot^-1 * pose_i * ext,
// n_neigh 5, CHECK_FOV false; surf rows = the pose_i block of LidarPureOdomPlaneNormFactor(point, coeffs, 1.0), corner rows = [1 0 0 0 0 0];
// gf_ratio == 1.0 takes every matched feature in order (:1380-1414), otherwise the stochastic greedy selection.
inline void good_feature_matching_odom(char type, const KdTree &tree, const Cloud &map, const Cloud &scan, const Pose &pivot, const Pose &pose_i,
                                       const Pose &ext, double gf_ratio, uint64_t seed, const MatchParams &mp, std::vector<Feature> &all,
                                       std::vector<unsigned char> &matched, std::vector<double> &jaco, std::vector<int> &sel, double H[36]) {
  const size_t n = scan.size();
  all.assign(n, Feature());
  matched.assign(n, 0);
  jaco.assign(n * 6, 0.0);
  const Pose pose_local = pose_mul(pose_inv(pivot), pose_mul(pose_i, ext));  // :1358
  double xp[7], xi[7], xe[7];
  pose_to_param(pivot, xp), pose_to_param(pose_i, xi), pose_to_param(ext, xe);
  for (size_t i = 0; i < n; i++) {
    const bool ok = type == 's' ? match_surf_point_from_map(tree, map, scan[i], pose_local, all[i], i, 5, false, mp)
                                : match_corner_point_from_map(tree, map, scan[i], pose_local, all[i], i, 5, false, mp);
    if (!ok) continue;
    matched[i] = 1;
    if (type == 's') {
      double r, Jp[7], Ji[7], Je[7];
      odom_plane_factor(all[i].point, all[i].coeffs, 1.0, xp, xi, xe, &r, Jp, Ji, Je);
      for (int k = 0; k < 6; k++) jaco[i * 6 + k] = Ji[k];
    } else {
      jaco[i * 6] = 1.0;  // Matrix<double, 1, 6>::Identity() (:1342)
    }
  }
  std::vector<float> xyz(n * 4, 0.f);
  good_feature_select(gf_ratio == 1.0 ? 0 : 3, gf_ratio, seed, (int)n, matched.data(), jaco.data(), xyz.data(), sel, H);
}

}  // namespace orc
