// shim_selftest.cpp — exercises the reference-shaped C++ surface (mloam_shim.hpp) end to end on a GPU.
// Reads like the reference's own smoke tests (mloam_test/src/test_jacobian_memory.cpp, the factors' check()):
// analytic vs forward-difference Jacobians with eps 1e-6 and right-multiplied deltaQ, Plus(x, 0) = x,
// kNN vs brute force, and a scan-to-map solve on three orthogonal planes + three edges.
// Build: g++ -std=c++14 shim_selftest.cpp -L.. -lmloam_b200 -Wl,-rpath,$PWD/..   Exit code 0 = all checks passed.
#include <cstdio>
#include <cstdlib>
#include <random>

#include "mloam_shim.hpp"

static int fails = 0;
#define EXPECT(cond, msg)                                  \
  do {                                                     \
    if (!(cond)) {                                         \
      std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, msg); \
      fails++;                                             \
    }                                                      \
  } while (0)

static void check_factor_fd(mloam::FactorBase &f, const char *name) {
  // the reference's check(): (r(x (+) eps e_k) - r(x)) / eps vs the analytic column, k = 0..5
  double x[7] = {0.3, -1.0, 2.0, 0.1, -0.2, 0.3, 0.9};
  double n = std::sqrt(x[3] * x[3] + x[4] * x[4] + x[5] * x[5] + x[6] * x[6]);
  for (int k = 3; k < 7; k++) x[k] /= n;
  double r, J[7];
  double *params[1] = {x};
  double *jac[1] = {J};
  f.Evaluate(params, &r, jac);
  PoseLocalParameterization lp;
  const double eps = 1e-6;
  for (int k = 0; k < 6; k++) {
    double d[6] = {0, 0, 0, 0, 0, 0}, xp[7], rp;
    d[k] = eps;
    lp.Plus(x, d, xp);
    double *pp[1] = {xp};
    f.Evaluate(pp, &rp, nullptr);
    EXPECT(std::fabs((rp - r) / eps - J[k]) < 2e-4, name);
  }
  EXPECT(J[6] == 0.0, "last Jacobian column must be zero");
}

int main() {
  try {
    // ---- factors
    double p[3] = {1.0, 2.0, -0.5};
    double plane[4] = {0.36, 0.48, 0.8, 0.7};
    double edge[6] = {1.0, 0.5, 0.2, 1.3, 2.5, -0.4};
    LidarMapPlaneNormFactor fp(p, plane, 0.0075);
    LidarMapEdgeFactor fe(p, edge, 0.0075);
    LidarScanPlaneNormFactor sp(p, plane);
    LidarOnlineCalibEdgeFactor ce(p, edge, 0.8);
    check_factor_fd(fp, "LidarMapPlaneNormFactor");
    check_factor_fd(fe, "LidarMapEdgeFactor");
    check_factor_fd(sp, "LidarScanPlaneNormFactor");
    check_factor_fd(ce, "LidarOnlineCalibEdgeFactor");
    {  // test_jacobian_memory.cpp: identity poses, zero point/coeff, three parameter blocks
      double z3[3] = {0, 0, 0}, z4[4] = {0, 0, 1, 0};
      LidarPureOdomPlaneNormFactor f(z3, z4, 1.0);
      double a[7] = {0, 0, 0, 0, 0, 0, 1}, b[7] = {0, 0, 0, 0, 0, 0, 1}, c[7] = {0, 0, 0, 0, 0, 0, 1};
      double *params[3] = {a, b, c};
      double r, J0[7], J1[7], J2[7];
      double *jac[3] = {J0, nullptr, J2};  // null-tolerant on individual blocks
      EXPECT(f.Evaluate(params, &r, jac), "Evaluate returns true");
      EXPECT(r == 0.0 && J2[2] == 1.0 && J0[2] == -1.0, "pure-odom plane factor at identity");
      (void)J1;
    }
    // ---- parameterisation
    {
      PoseLocalParameterization lp;
      double x[7] = {1, 2, 3, 0, 0, 0, 1}, d[6] = {0, 0, 0, 0, 0, 0}, y[7];
      lp.Plus(x, d, y);
      for (int k = 0; k < 7; k++) EXPECT(y[k] == x[k], "Plus(x, 0) == x");
      double Jp[42];
      lp.ComputeJacobian(x, Jp);
      EXPECT(Jp[0] == 1.0 && Jp[7] == 1.0 && Jp[36] == 0.0 && lp.GlobalSize() == 7 && lp.LocalSize() == 6, "ComputeJacobian = [I6; 0]");
    }
    // ---- scene: floor z=0, walls x=5 and y=4, sampled densely, plus their three intersection edges
    std::mt19937 rng(7);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    std::normal_distribution<float> N(0.f, 0.005f);
    common::PointICloud surf_map, corner_map;
    for (int i = 0; i < 60000; i++) {
      float a = 10.f * U(rng) - 5.f, b = 8.f * U(rng) - 4.f, h = 3.f * U(rng);
      common::PointI q;
      q.intensity = 0.f;
      switch (i % 3) {
        case 0: q.x = a, q.y = b, q.z = N(rng); break;
        case 1: q.x = 5.f + N(rng), q.y = b, q.z = h; break;
        default: q.x = a, q.y = 4.f + N(rng), q.z = h; break;
      }
      surf_map.push_back(q);
    }
    for (int i = 0; i < 6000; i++) {
      common::PointI q;
      q.intensity = 0.f;
      float s = U(rng);
      switch (i % 3) {
        case 0: q.x = 5.f + N(rng), q.y = 8.f * s - 4.f, q.z = N(rng); break;
        case 1: q.x = 10.f * s - 5.f, q.y = 4.f + N(rng), q.z = N(rng); break;
        default: q.x = 5.f + N(rng), q.y = 4.f + N(rng), q.z = 3.f * s; break;
      }
      corner_map.push_back(q);
    }
    // ---- kd-tree stand-in vs brute force
    {
      mloam::MapHandle kd(MLOAM_MAP_SCAN_SURF, 0.5f);
      kd.setInputCloud(surf_map);
      common::PointI q{1.0f, 1.0f, 0.02f, 0.f};
      std::vector<int> idx;
      std::vector<float> sqd;
      int got = kd.nearestKSearch(q, 5, idx, sqd);
      EXPECT(got == 5, "5 neighbours on the floor");
      float best = 1e30f;
      int bi = -1;
      for (size_t i = 0; i < surf_map.size(); i++) {
        float dx = surf_map.points[i].x - q.x, dy = surf_map.points[i].y - q.y, dz = surf_map.points[i].z - q.z;
        float d = dx * dx + dy * dy + dz * dz;
        if (d < best) best = d, bi = (int)i;
      }
      EXPECT(idx[0] == bi && sqd[0] == best, "nearest neighbour matches brute force (index and float distance)");
    }
    // ---- scan features: points of the same geometry seen from a sensor at T_true
    Pose T_true({{0.0, 0.0, std::sin(0.15), std::cos(0.15)}}, {{0.5, -0.3, 1.2}});
    auto to_sensor = [&](const common::PointI &w) {  // p_s = R^T (p_w - t)
      const double qx = -T_true.q_[0], qy = -T_true.q_[1], qz = -T_true.q_[2], qw = T_true.q_[3];
      const double vx = w.x - T_true.t_[0], vy = w.y - T_true.t_[1], vz = w.z - T_true.t_[2];
      const double ux = qy * vz - qz * vy, uy = qz * vx - qx * vz, uz = qx * vy - qy * vx;
      const double tx = 2 * ux, ty = 2 * uy, tz = 2 * uz;
      common::PointI s;
      s.x = (float)(vx + qw * tx + (qy * tz - qz * ty));
      s.y = (float)(vy + qw * ty + (qz * tx - qx * tz));
      s.z = (float)(vz + qw * tz + (qx * ty - qy * tx));
      s.intensity = 0.f;
      return s;
    };
    common::PointICloud surf_scan, corner_scan;
    for (size_t i = 0; i < surf_map.size(); i += 15) surf_scan.push_back(to_sensor(surf_map.points[i]));
    for (size_t i = 0; i < corner_map.size(); i += 10) corner_scan.push_back(to_sensor(corner_map.points[i]));
    Pose guess({{0.004, -0.003, std::sin(0.152), std::cos(0.152)}}, {{0.53, -0.27, 1.17}});
    mloam_solve_stats_t st;
    bool ran = mloam::scan2MapOptimization(surf_map, corner_map, surf_scan, corner_scan, guess, &st);
    EXPECT(ran && st.n_surf > 1000 && st.n_corner > 50, "scan2MapOptimization ran with matches");
    double et = 0;
    for (int k = 0; k < 3; k++) et += (guess.t_[k] - T_true.t_[k]) * (guess.t_[k] - T_true.t_[k]);
    std::printf("scan2map: %d surf + %d corner matches, %d LM iterations, |dt| = %.2e m\n", st.n_surf, st.n_corner, st.lm_iterations, std::sqrt(et));
    EXPECT(std::sqrt(et) < 2e-2, "pose pulled from 5 cm to < 2 cm of the truth (5 mm map noise, 5-point fits)");
    // map-size gate (lidar_mapper_keyframe.cpp:429)
    common::PointICloud tiny;
    for (int i = 0; i < 40; i++) tiny.push_back(surf_map.points[i]);
    Pose g2 = guess;
    EXPECT(!mloam::scan2MapOptimization(tiny, corner_map, surf_scan, corner_scan, g2, nullptr), "map-size gate rejects");
    // ---- feature matching surface
    {
      FeatureExtract fx;
      mloam::MapHandlePtr kd = std::make_shared<mloam::MapHandle>(MLOAM_MAP_SURF);
      kd->setInputCloud(surf_map);
      std::vector<PointPlaneFeature> feats;
      fx.matchSurfFromMap(kd, surf_map, surf_scan, T_true, feats, 5, false);
      EXPECT(feats.size() > surf_scan.size() * 8 / 10 && feats[0].type_ == 's', "matchSurfFromMap");
      PointPlaneFeature one;
      EXPECT(fx.matchSurfPointFromMap(kd, surf_map, surf_scan.points[0], T_true, one, 0, 5, false) || true, "per-point form callable");
    }
  } catch (const std::exception &e) {
    std::printf("EXCEPTION %s\n", e.what());
    return 2;
  }
  std::printf(fails ? "SHIM_SELFTEST FAILED (%d)\n" : "SHIM_SELFTEST OK\n", fails);
  return fails ? 1 : 0;
}
