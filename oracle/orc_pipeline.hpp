// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
// Orchestrators of the hot path: scan2MapOptimization (estimator/src/lidarMapper/
// lidar_mapper_keyframe.cpp:423-639, gf_method "wo_gf") and LidarTracker::trackCloud
// (estimator/src/lidarTracker/lidar_tracker.cpp:23-129).
#pragma once
#include "orc_match.hpp"
#include "orc_gf.hpp"
#include "orc_solver.hpp"
#include <chrono>

namespace orc {

struct Scan2MapOptions {
  int max_outer = 2;          // max_iter, lidar_mapper_keyframe.cpp:439
  int max_inner = 30;         // options.max_num_iterations, :590
  double huber_a = 0.1;       // :443
  double eig_thre = 100.0;    // MAP_EIG_THRE, config yaml :140
  int n_neigh = 5;            // lidar_mapper.h:253
  bool check_fov = false;     // all *PointFromMap callers pass false (lidar_mapper.h:193-283)
  bool point_plane = true;    // POINT_PLANE_FACTOR
  bool point_edge = true;     // POINT_EDGE_FACTOR
  double cov_trace = 0.0075;  // trace(COV_MEASUREMENT) = 3 * 0.0025 (with_ua=false, :541-545)
  // with_ua=true (:541-545,556-560): per scan point covariance traces (extractCov of PointIWithCov); null = with_ua false
  const std::vector<double> *surf_cov_trace = nullptr, *corner_cov_trace = nullptr;
  MatchParams mp;
  // FLAGS_gf_method / FLAGS_gf_ratio_ini (:474-492): 0 wo_gf, 1 rnd, 2 fps, 3 gd_fix; seed of iteration i, set s (0 corner,
  // 1 surf) = gf_seed + 2 i + s (orc_gf.hpp makes the reference's random_device explicit)
  int gf_method = 0;
  double gf_ratio = 1.0;
  uint64_t gf_seed = 0;
  // sharded frame (one LiDAR group per GPU): the selection runs per consecutive group of the scans with the same seeds;
  // empty = one selection over the whole scan (the reference)
  std::vector<int> gf_groups_surf, gf_groups_corner;
};
struct Scan2MapResult {
  Pose pose;
  int ran = 0;                 // 0 when the map-size gate (:429) rejects
  int n_surf = 0, n_corner = 0;  // matches of the last outer iteration
  int lm_iterations = 0;       // summed over outer iterations
  double final_cost = 0;
  double H_last[36];           // loss-corrected J^T J evaluated before the last Solve (:575-581)
  double eig_last[6];
  int degenerate = 0;
  double t_kdtree = 0, t_match = 0, t_solver = 0;  // seconds; tags mapping_kdtree / mapping_match_feat / mapping_solver
};

inline double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

inline Scan2MapResult scan2map(const Cloud &surf_map, const Cloud &corner_map, const Cloud &surf_scan,
                               const Cloud &corner_scan, const Pose &pose_init, const Scan2MapOptions &o) {
  Scan2MapResult res;
  res.pose = pose_init;
  std::memset(res.H_last, 0, sizeof(res.H_last));
  std::memset(res.eig_last, 0, sizeof(res.eig_last));
  if (!((surf_map.size() > 50) && (corner_map.size() > 10))) return res;  // :429
  res.ran = 1;
  double t0 = now_s();
  KdTree kd_surf, kd_corner;  // :433-434
  kd_surf.setInputCloud(&surf_map);
  kd_corner.setInputCloud(&corner_map);
  res.t_kdtree = now_s() - t0;
  double para_pose[7];
  Pose pose_wmap_curr = pose_init;
  const double sinfo = map_sqrt_info(o.cov_trace);
  for (int iter_cnt = 0; iter_cnt < o.max_outer; iter_cnt++) {
    pose_to_param(pose_wmap_curr, para_pose);  // vector2Double :236-245
    Problem problem;
    problem.huber_a = o.huber_a;
    int pid = problem.add_param(para_pose);
    t0 = now_s();
    std::vector<Feature> corner_f, surf_f;
    if (o.gf_method == 0) {
      if (o.point_edge)  // :503-517, wo_gf branch lidar_mapper.h:257-299: every feature, in order
        match_from_map('c', kd_corner, corner_map, corner_scan, pose_wmap_curr, corner_f, o.n_neigh, o.check_fov, o.mp);
      if (o.point_plane)  // :518-532
        match_from_map('s', kd_surf, surf_map, surf_scan, pose_wmap_curr, surf_f, o.n_neigh, o.check_fov, o.mp);
    } else {  // :503-532 with a good-feature method: residual blocks of the SELECTED features, in selection order (:537,:552)
      std::vector<Feature> all;
      std::vector<unsigned char> mt;
      std::vector<double> jc;
      std::vector<int> sel;
      double subH[36];
      auto select = [&](char type, const KdTree &kd, const Cloud &map, const Cloud &scan, const std::vector<double> *cov,
                        const std::vector<int> &groups, uint64_t seed, std::vector<Feature> &out) {
        if (groups.empty()) {
          good_feature_matching(type, kd, map, scan, pose_wmap_curr, cov, o.cov_trace, o.gf_method, o.gf_ratio, seed, o.n_neigh, o.mp, all, mt,
                                jc, sel, subH);
          for (int q : sel) out.push_back(all[q]);
          return;
        }
        size_t off = 0;
        for (int gsz : groups) {  // per group, as each GPU of the sharded frame selects among its own features
          Cloud part(scan.begin() + off, scan.begin() + off + gsz);
          std::vector<double> cpart;
          if (cov) cpart.assign(cov->begin() + off, cov->begin() + off + gsz);
          good_feature_matching(type, kd, map, part, pose_wmap_curr, cov ? &cpart : nullptr, o.cov_trace, o.gf_method, o.gf_ratio, seed,
                                o.n_neigh, o.mp, all, mt, jc, sel, subH);
          for (int q : sel) {
            Feature f = all[q];
            f.idx += (int)off;
            out.push_back(f);
          }
          off += gsz;
        }
      };
      if (o.point_edge) select('c', kd_corner, corner_map, corner_scan, o.corner_cov_trace, o.gf_groups_corner, o.gf_seed + 2 * (uint64_t)iter_cnt, corner_f);
      if (o.point_plane) select('s', kd_surf, surf_map, surf_scan, o.surf_cov_trace, o.gf_groups_surf, o.gf_seed + 2 * (uint64_t)iter_cnt + 1, surf_f);
    }
    res.t_match += now_s() - t0;
    res.n_surf = (int)surf_f.size(), res.n_corner = (int)corner_f.size();
    for (const Feature &f : surf_f) {  // :537-549
      const double si = o.surf_cov_trace ? map_sqrt_info((*o.surf_cov_trace)[f.idx]) : sinfo;
      ResidualBlock b{F_PLANE, f.point, {f.coeffs[0], f.coeffs[1], f.coeffs[2], f.coeffs[3], 0, 0}, si, {pid, 0, 0}};
      problem.blocks.push_back(b);
    }
    for (const Feature &f : corner_f) {  // :552-571
      const double si = o.corner_cov_trace ? map_sqrt_info((*o.corner_cov_trace)[f.idx]) : sinfo;
      ResidualBlock b{F_EDGE, f.point, {f.coeffs[0], f.coeffs[1], f.coeffs[2], f.coeffs[3], f.coeffs[4], f.coeffs[5]},
                      si, {pid, 0, 0}};
      problem.blocks.push_back(b);
    }
    t0 = now_s();
    // :575-582 problem.Evaluate -> J^T J -> evalDegenracy
    {
      NormalEq ne;
      std::vector<const double *> xs{para_pose};
      problem.evaluate(xs, true, ne);
      if (ne.rows > 0) {
        std::memcpy(res.H_last, ne.H.data(), sizeof(res.H_last));
        eval_degeneracy(res.H_last, o.eig_thre, problem.local[pid], res.eig_last);
        res.degenerate = problem.local[pid].is_degenerate ? 1 : 0;
      }
    }
    SolveSummary s = solve(problem, o.max_inner);  // :586-596
    res.t_solver += now_s() - t0;
    res.lm_iterations += s.iterations;
    res.final_cost = s.final_cost;
    // double2Vector :247-252 (no normalisation)
    pose_wmap_curr = pose_from_param_raw(para_pose);
  }
  res.pose = pose_wmap_curr;
  return res;
}

// evalPointUncertainty, estimator/src/lidarMapper/associate_uct.hpp:164-215 (pointToFS :149-156):
// cov_point = top-left 3x3 of G diag(cov_pose(6x6), COV_MEASUREMENT(3x3)) G^T,  G = [ I3 | -[T p]x | R ].
// Output packed like PointIWithCov::cov_vec (float [xx xy xz yy yz zz], point_with_cov.hpp:45-53).
inline void eval_point_uncertainty(const PointI &pi, const Pose &pose, const double cov_pose[36], const double cov_meas[9], float cov6[6]) {
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
  double C[3][3];
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
  cov6[0] = (float)C[0][0], cov6[1] = (float)C[0][1], cov6[2] = (float)C[0][2], cov6[3] = (float)C[1][1], cov6[4] = (float)C[1][2],
  cov6[5] = (float)C[2][2];
}

struct TrackOptions {
  int max_outer = 2;     // lidar_tracker.cpp:44
  int max_inner = 4;     // :114
  double huber_a = 0.1;  // :47
  MatchParams mp;
};
struct TrackResult {
  Pose pose;
  int n_corner = 0, n_surf = 0;
  int lm_iterations = 0;
};

// LidarTracker::trackCloud, lidar_tracker.cpp:23-129
inline TrackResult track_cloud(const Cloud &prev_corner_less_sharp, const Cloud &prev_surf_less_flat,
                               const Cloud &cur_corner_sharp, const Cloud &cur_surf_flat, const Pose &pose_ini,
                               const TrackOptions &o) {
  TrackResult res;
  KdTree kd_corner, kd_surf;
  kd_corner.setInputCloud(&prev_corner_less_sharp);
  kd_surf.setInputCloud(&prev_surf_less_flat);
  double para_pose[7];
  pose_to_param(pose_ini, para_pose);
  for (int iter_cnt = 0; iter_cnt < o.max_outer; iter_cnt++) {
    Problem problem;
    problem.huber_a = o.huber_a;
    int pid = problem.add_param(para_pose);
    Pose pose_local = make_pose(Q4{para_pose[3], para_pose[4], para_pose[5], para_pose[6]},
                                V3{para_pose[0], para_pose[1], para_pose[2]});  // :54-55
    std::vector<Feature> cf, sf;
    match_corner_from_scan(kd_corner, prev_corner_less_sharp, cur_corner_sharp, pose_local, cf, o.mp);
    match_surf_from_scan(kd_surf, prev_surf_less_flat, cur_surf_flat, pose_local, sf, o.mp);
    res.n_corner = (int)cf.size(), res.n_surf = (int)sf.size();
    if (cf.size() + sf.size() < 10) continue;  // :64-68
    for (const Feature &f : sf)
      problem.blocks.push_back(
          ResidualBlock{F_PLANE, f.point, {f.coeffs[0], f.coeffs[1], f.coeffs[2], f.coeffs[3], 0, 0}, 1.0, {pid, 0, 0}});
    for (const Feature &f : cf)
      problem.blocks.push_back(ResidualBlock{
          F_EDGE_VEC, f.point, {f.coeffs[0], f.coeffs[1], f.coeffs[2], f.coeffs[3], f.coeffs[4], f.coeffs[5]}, 1.0, {pid, 0, 0}});
    SolveSummary s = solve(problem, o.max_inner);
    res.lm_iterations += s.iterations;
  }
  res.pose = make_pose(Q4{para_pose[3], para_pose[4], para_pose[5], para_pose[6]},
                       V3{para_pose[0], para_pose[1], para_pose[2]});  // :126-128
  return res;
}

}  // namespace orc
