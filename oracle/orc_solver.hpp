// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
// Restatement of what the reference delegates to Ceres Solver 1.12.0 (un-vendored; pinned at
// docker/Dockerfile:3): problem assembly with a robust loss, and ceres::Solve with the options the
// reference sets (DENSE_SCHUR, max_num_iterations, everything else default: TRUST_REGION +
// LEVENBERG_MARQUARDT, initial_trust_region_radius 1e4, max_trust_region_radius 1e16,
// min_relative_decrease 1e-3, min/max_lm_diagonal 1e-6/1e32, function_tolerance 1e-6,
// gradient_tolerance 1e-10, parameter_tolerance 1e-8, jacobi_scaling, monotonic steps).
// Call sites: lidar_tracker.cpp:112-120 (4 it), lidar_mapper_keyframe.cpp:587-596 (30 it),
// estimator.cpp:605-615,861.  PARITY UNPINNED at this boundary (no Ceres here):
// the loop below follows the published algorithm of Ceres' trust_region_minimizer.cc /
// levenberg_marquardt_strategy.cc (1.12 pre-refactor ordering: the function-tolerance test runs
// before the step is applied).  With one 6-dof block (or a handful) the Schur/dense distinction is
// immaterial: the step solves (J^T J + D^T D) y = J^T f by dense Cholesky.
#pragma once
#include "orc_factors.hpp"

namespace orc {

enum FactorKind {
  F_PLANE = 0,        // LidarMapPlaneNormFactor / LidarScanPlaneNormFactor / LidarOnlineCalibPlaneNormFactor
  F_EDGE = 1,         // LidarMapEdgeFactor / LidarOnlineCalibEdgeFactor
  F_EDGE_VEC = 2,     // LidarScanEdgeFactorVector
  F_ODOM_PLANE = 3,   // LidarPureOdomPlaneNormFactor
  F_ODOM_EDGE = 4     // LidarPureOdomEdgeFactor
};

struct ResidualBlock {
  int kind;
  V3 point;
  double coeffs[6];
  double sqrt_info;
  int pb[3];  // parameter-block ids (1 used for single-pose factors)
};

struct NormalEq {
  int n = 0;               // local size
  double cost = 0;         // 1/2 sum rho(s)
  std::vector<double> H;   // n x n (full, symmetric), loss-corrected J^T J
  std::vector<double> g;   // n, loss-corrected J^T r
  int rows = 0;
};

struct Problem {
  std::vector<double *> params;                 // each 7 doubles, caller-owned (AddParameterBlock)
  std::vector<bool> constant;                   // SetParameterBlockConstant
  std::vector<PoseLocalParameterization> local; // one per block
  std::vector<ResidualBlock> blocks;
  double huber_a = 0.1;
  bool use_loss = true;

  int add_param(double *x, bool is_const = false) {
    params.push_back(x);
    constant.push_back(is_const);
    local.emplace_back();
    return (int)params.size() - 1;
  }
  std::vector<int> offsets(int *n_local = nullptr) const {
    std::vector<int> off(params.size(), -1);
    int n = 0;
    for (size_t i = 0; i < params.size(); i++)
      if (!constant[i]) off[i] = n, n += 6;
    if (n_local) *n_local = n;
    return off;
  }

  // Evaluate at the given global state (xs[i] -> 7 doubles).  Loss applied as Ceres' Corrector does
  // for rho'' <= 0: r <- sqrt(rho') r, J <- sqrt(rho') J per residual BLOCK.
  void evaluate(const std::vector<const double *> &xs, bool want_jac, NormalEq &ne) const {
    int n;
    std::vector<int> off = offsets(&n);
    ne.n = n;
    ne.cost = 0;
    ne.rows = 0;
    ne.H.assign((size_t)n * n, 0.0);
    ne.g.assign(n, 0.0);
    double r[3], J[3][21];
    for (const ResidualBlock &b : blocks) {
      int nr = 1, np = 1;
      switch (b.kind) {
        case F_PLANE:
          plane_factor(b.point, b.coeffs, b.sqrt_info, xs[b.pb[0]], r, want_jac ? J[0] : nullptr);
          break;
        case F_EDGE:
          edge_factor(b.point, b.coeffs, b.sqrt_info, xs[b.pb[0]], r, want_jac ? J[0] : nullptr);
          break;
        case F_EDGE_VEC: {
          double Jv[21];
          edge_vector_factor(b.point, b.coeffs, xs[b.pb[0]], r, want_jac ? Jv : nullptr);
          nr = 3;
          if (want_jac)
            for (int i = 0; i < 3; i++)
              for (int j = 0; j < 7; j++) J[i][j] = Jv[i * 7 + j];
          break;
        }
        case F_ODOM_PLANE:
        case F_ODOM_EDGE: {
          np = 3;
          double *Jp = want_jac ? &J[0][0] : nullptr, *Ji = want_jac ? &J[0][7] : nullptr, *Je = want_jac ? &J[0][14] : nullptr;
          if (b.kind == F_ODOM_PLANE)
            odom_plane_factor(b.point, b.coeffs, b.sqrt_info, xs[b.pb[0]], xs[b.pb[1]], xs[b.pb[2]], r, Jp, Ji, Je);
          else
            odom_edge_factor(b.point, b.coeffs, b.sqrt_info, xs[b.pb[0]], xs[b.pb[1]], xs[b.pb[2]], r, Jp, Ji, Je);
          break;
        }
      }
      double s = 0;
      for (int i = 0; i < nr; i++) s += r[i] * r[i];
      double rho = s, rho1 = 1.0;
      if (use_loss) huber(huber_a, s, &rho, &rho1);
      ne.cost += 0.5 * rho;
      ne.rows += nr;
      if (!want_jac) continue;
      const double sc = std::sqrt(rho1);
      for (int i = 0; i < nr; i++) {
        const double ri = sc * r[i];
        for (int k = 0; k < np; k++) {
          int ok = off[b.pb[k]];
          if (ok < 0) continue;
          const double *Jk = &J[i][k * 7];
          for (int a = 0; a < 6; a++) {
            const double ja = sc * Jk[a];
            ne.g[ok + a] += ja * ri;
            for (int l = 0; l < np; l++) {
              int ol = off[b.pb[l]];
              if (ol < 0) continue;
              const double *Jl = &J[i][l * 7];
              for (int c = 0; c < 6; c++) ne.H[(size_t)(ok + a) * n + ol + c] += ja * (sc * Jl[c]);
            }
          }
        }
      }
    }
  }
};

struct SolveSummary {
  int iterations = 0;       // LM iterations attempted (successful + unsuccessful)
  int successful = 0;
  double initial_cost = 0, final_cost = 0;
  int termination = 0;      // 0 max-iter, 1 function tol, 2 parameter tol, 3 gradient tol, 4 failure
};

// ceres::Solve restatement (see header comment).
inline SolveSummary solve(Problem &prob, int max_num_iterations) {
  SolveSummary sum;
  int n;
  std::vector<int> off = prob.offsets(&n);
  const size_t nb = prob.params.size();
  std::vector<std::vector<double>> x(nb, std::vector<double>(7)), xc(nb, std::vector<double>(7));
  for (size_t i = 0; i < nb; i++) std::memcpy(x[i].data(), prob.params[i], 7 * sizeof(double));
  auto ptrs = [&](std::vector<std::vector<double>> &v) {
    std::vector<const double *> p(nb);
    for (size_t i = 0; i < nb; i++) p[i] = v[i].data();
    return p;
  };
  auto xnorm = [&](std::vector<std::vector<double>> &v) {
    double s = 0;
    for (size_t i = 0; i < nb; i++)
      if (!prob.constant[i])
        for (int k = 0; k < 7; k++) s += v[i][k] * v[i][k];
    return std::sqrt(s);
  };
  auto plus = [&](std::vector<std::vector<double>> &from, const std::vector<double> &delta,
                  std::vector<std::vector<double>> &to) {
    for (size_t i = 0; i < nb; i++) {
      if (off[i] < 0) to[i] = from[i];
      else prob.local[i].Plus(from[i].data(), &delta[off[i]], to[i].data());
    }
  };
  auto writeback = [&]() {
    for (size_t i = 0; i < nb; i++)
      if (!prob.constant[i]) std::memcpy(prob.params[i], x[i].data(), 7 * sizeof(double));
  };
  const double min_diag = 1e-6, max_diag = 1e32, max_radius = 1e16;
  const double function_tolerance = 1e-6, parameter_tolerance = 1e-8, gradient_tolerance = 1e-10;
  const double min_relative_decrease = 1e-3;
  double radius = 1e4, decrease_factor = 2.0;
  bool reuse_diagonal = false;

  NormalEq ne;
  prob.evaluate(ptrs(x), true, ne);
  double cost = ne.cost;
  sum.initial_cost = sum.final_cost = cost;
  if (n == 0) return sum;
  double x_norm = xnorm(x);
  // Jacobi scaling from the initial Jacobian: scale_j = 1 / (1 + sqrt(sum_i J_ij^2))
  std::vector<double> scale(n), diag(n), lmd(n), step(n), delta(n), Hs((size_t)n * n), gs(n);
  for (int j = 0; j < n; j++) scale[j] = 1.0 / (1.0 + std::sqrt(ne.H[(size_t)j * n + j]));
  auto gradient_max_norm = [&]() {
    std::vector<double> neg(n);
    for (int j = 0; j < n; j++) neg[j] = -ne.g[j];
    plus(x, neg, xc);
    double m = 0;
    for (size_t i = 0; i < nb; i++)
      if (off[i] >= 0)
        for (int k = 0; k < 7; k++) m = std::max(m, std::fabs(x[i][k] - xc[i][k]));
    return m;
  };
  if (gradient_max_norm() <= gradient_tolerance) {
    sum.termination = 3;
    return sum;
  }
  int num_consecutive_invalid = 0;
  int iteration = 0;
  while (true) {
    if (iteration >= max_num_iterations) {
      sum.termination = 0;
      break;
    }
    // ---- LevenbergMarquardtStrategy::ComputeStep on the column-scaled system
    for (int a = 0; a < n; a++) {
      gs[a] = scale[a] * ne.g[a];
      for (int b = 0; b < n; b++) Hs[(size_t)a * n + b] = scale[a] * ne.H[(size_t)a * n + b] * scale[b];
    }
    if (!reuse_diagonal)
      for (int j = 0; j < n; j++) diag[j] = std::min(std::max(Hs[(size_t)j * n + j], min_diag), max_diag);
    for (int j = 0; j < n; j++) lmd[j] = std::sqrt(diag[j] / radius);
    std::vector<double> A(Hs);
    for (int j = 0; j < n; j++) A[(size_t)j * n + j] += lmd[j] * lmd[j];
    bool ok = cholesky(n, A.data());
    if (ok) {
      cholesky_solve(n, A.data(), gs.data(), step.data());
      for (int j = 0; j < n; j++) {
        step[j] = -step[j];
        if (!std::isfinite(step[j])) ok = false;
      }
    }
    reuse_diagonal = true;
    iteration++;
    sum.iterations = iteration;
    double model_cost_change = 0;
    if (ok) {
      // model_cost_change = -(J s).(r + J s / 2) = -(s.g + s.H s / 2) on the scaled system
      double sg = 0, sHs = 0;
      for (int a = 0; a < n; a++) {
        sg += step[a] * gs[a];
        double t = 0;
        for (int b = 0; b < n; b++) t += Hs[(size_t)a * n + b] * step[b];
        sHs += step[a] * t;
      }
      model_cost_change = -(sg + 0.5 * sHs);
      if (model_cost_change < 0) ok = false;
    }
    if (!ok) {
      if (++num_consecutive_invalid >= 5) {
        sum.termination = 4;
        break;
      }
      radius *= 0.5;  // StepIsInvalid
      reuse_diagonal = true;
      continue;
    }
    num_consecutive_invalid = 0;
    for (int j = 0; j < n; j++) delta[j] = step[j] * scale[j];
    plus(x, delta, xc);
    NormalEq nc;
    prob.evaluate(ptrs(xc), false, nc);
    const double new_cost = nc.cost;
    double step_norm = 0;
    for (size_t i = 0; i < nb; i++)
      if (off[i] >= 0)
        for (int k = 0; k < 7; k++) step_norm += (x[i][k] - xc[i][k]) * (x[i][k] - xc[i][k]);
    step_norm = std::sqrt(step_norm);
    if (step_norm <= parameter_tolerance * (x_norm + parameter_tolerance)) {
      sum.termination = 2;
      break;
    }
    const double cost_change = cost - new_cost;
    if (std::fabs(cost_change) <= function_tolerance * cost) {
      sum.termination = 1;
      break;
    }
    const double relative_decrease = cost_change / model_cost_change;
    if (relative_decrease > min_relative_decrease) {
      sum.successful++;
      // StepAccepted
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * relative_decrease - 1.0, 3));
      radius = std::min(max_radius, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
      x = xc;
      x_norm = xnorm(x);
      prob.evaluate(ptrs(x), true, ne);
      cost = ne.cost;
      if (gradient_max_norm() <= gradient_tolerance) {
        sum.termination = 3;
        break;
      }
    } else {
      // StepRejected
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = true;
    }
  }
  sum.final_cost = cost;
  writeback();
  return sum;
}

// evalDegenracy: lidar_mapper_keyframe.cpp:1172-1204 (threshold MAP_EIG_THRE), lidar_tracker.cpp:131-163
// (threshold 10).  H is the loss-corrected J^T J of problem.Evaluate (block 0,0,6,6).  Eigenvalues ascending;
// eigenvectors below the threshold are zeroed (stop at the first one above);
// V_update = (V_f^T)^-1 V_p^T  (V_f orthogonal => V_f V_p^T).
inline void eval_degeneracy(const double H[36], double eig_thre, PoseLocalParameterization &lp, double eig_out[6]) {
  double w[6], Vf[36], Vp[36];
  eig_sym(6, H, w, Vf);
  std::memcpy(Vp, Vf, sizeof(Vf));
  for (int j = 0; j < 6; j++) {
    if (w[j] < eig_thre) {
      for (int k = 0; k < 6; k++) Vp[k * 6 + j] = 0.0;
      lp.is_degenerate = true;
    } else {
      break;
    }
  }
  if (eig_out) std::memcpy(eig_out, w, sizeof(w));
  if (lp.is_degenerate) {
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 6; j++) {
        double s = 0;
        for (int k = 0; k < 6; k++) s += Vf[i * 6 + k] * Vp[j * 6 + k];
        lp.V_update[i * 6 + j] = s;
      }
  }
}

}  // namespace orc
