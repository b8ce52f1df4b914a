// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
// Cost functions of estimator/src/factor/: residual + analytic Jacobians, all double.
// Parameter block layout [tx ty tz qx qy qz qw]; Jacobians row-major rows x 7, column 6 == 0.
#pragma once
#include "orc_math.hpp"

namespace orc {

// sqrt_info of the map factors: sqrt(1/trace(cov)), then clamped (lidar_map_factor.hpp:34,41 / :137,140)
inline double map_sqrt_info(double cov_trace) {
  double s = std::sqrt(1 / cov_trace);
  return s >= 3.0 ? 1.0 : s / 3.0;
}

// Eigen 3.3 MatrixBase::normalized(): a zero vector is returned unchanged
inline V3 normalized(const V3 &v) {
  double n = norm(v);
  return n > 0.0 ? V3{v.x / n, v.y / n, v.z / n} : v;
}

inline void set_row(double *J, double s, const V3 &a, const V3 &b) {
  J[0] = s * a.x, J[1] = s * a.y, J[2] = s * a.z, J[3] = s * b.x, J[4] = s * b.y, J[5] = s * b.z, J[6] = 0.0;
}

// Point-to-plane, single pose.  LidarMapPlaneNormFactor::Evaluate lidar_map_factor.hpp:44-68;
// LidarScanPlaneNormFactor (s_=1) lidar_scan_factor.hpp:33-60; LidarOnlineCalibPlaneNormFactor
// lidar_online_calib_factor.hpp:34-60.   r = s (w.(Rp+t) + d);  J = s [w^T | -w^T R [p]x]
inline void plane_factor(const V3 &p, const double *coeff, double sinfo, const double *x, double *r, double *J) {
  Q4 q{x[3], x[4], x[5], x[6]};
  V3 t{x[0], x[1], x[2]};
  V3 w{coeff[0], coeff[1], coeff[2]};
  double a = dot(w, qrot(q, p) + t) + coeff[3];
  r[0] = sinfo * a;
  if (J) {
    M3 R = qmat(q);
    V3 wr = vecmat(w, R);                // w^T R
    V3 jr = -vecmat(wr, skew(p));        // -(w^T R) [p]x
    set_row(J, sinfo, w, jr);
  }
}

// Point-to-line, scalar.  LidarMapEdgeFactor::Evaluate lidar_map_factor.hpp:143-171;
// LidarScanEdgeFactor lidar_scan_factor.hpp:139-168; LidarOnlineCalibEdgeFactor
// lidar_online_calib_factor.hpp:135-163.
inline void edge_factor(const V3 &p, const double *coeff, double sinfo, const double *x, double *r, double *J) {
  Q4 q{x[3], x[4], x[5], x[6]};
  V3 t{x[0], x[1], x[2]};
  V3 lpa{coeff[0], coeff[1], coeff[2]}, lpb{coeff[3], coeff[4], coeff[5]};
  V3 lp = qrot(q, p) + t;
  V3 nu = cross(lp - lpa, lp - lpb);
  V3 de = lpa - lpb;
  double nun = norm(nu), den = norm(de);
  r[0] = sinfo * nun / den;
  if (J) {
    M3 R = qmat(q);
    V3 eta = (1.0 / den) * normalized(nu);  // 1/|de| * nu.normalized()^T
    M3 S = skew(de);
    V3 eS = vecmat(eta, S);
    V3 jt = -eS;
    V3 jr = vecmat(vecmat(eS, R), skew(p));
    set_row(J, sinfo, jt, jr);
  }
}

// Point-to-line, 3-vector (tracker).  LidarScanEdgeFactorVector::Evaluate lidar_scan_factor.hpp:245-279.
// r = nu/|de| (3);  J = 1/|de| [-[de]x | [de]x R [p]x]   (3 x 7 row-major)
inline void edge_vector_factor(const V3 &p, const double *coeff, const double *x, double *r, double *J) {
  Q4 q{x[3], x[4], x[5], x[6]};
  V3 t{x[0], x[1], x[2]};
  V3 lpa{coeff[0], coeff[1], coeff[2]}, lpb{coeff[3], coeff[4], coeff[5]};
  V3 lp = qrot(q, p) + t;
  V3 nu = cross(lp - lpa, lp - lpb);
  V3 de = lpa - lpb;
  double den = norm(de);
  r[0] = nu.x / den, r[1] = nu.y / den, r[2] = nu.z / den;
  if (J) {
    M3 R = qmat(q);
    double eta = 1.0 / den;
    M3 S = skew(de);
    M3 SRP = matmul(matmul(S, R), skew(p));
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) {
        J[i * 7 + j] = -eta * S(i, j);
        J[i * 7 + 3 + j] = eta * SRP(i, j);
      }
      J[i * 7 + 6] = 0.0;
    }
  }
}

// Three-pose chain (pivot, pose_i, ext).  LidarPureOdomPlaneNormFactor::Evaluate
// lidar_pure_odom_factor.hpp:38-101.  J[k] (k = pivot, i, ext) may be null.
inline void odom_plane_factor(const V3 &p, const double *coeff, double sinfo, const double *xp, const double *xi,
                              const double *xe, double *r, double *Jp, double *Ji, double *Je) {
  Q4 Qp{xp[3], xp[4], xp[5], xp[6]}, Qi{xi[3], xi[4], xi[5], xi[6]}, Qe{xe[3], xe[4], xe[5], xe[6]};
  V3 tp{xp[0], xp[1], xp[2]}, ti{xi[0], xi[1], xi[2]}, te{xe[0], xe[1], xe[2]};
  Q4 Qpi = qmul(qconj(Qp), Qi);
  V3 tpi = qrot(qconj(Qp), ti - tp);
  Q4 Qepi = qmul(Qpi, Qe);
  V3 tepi = qrot(Qpi, te) + tpi;
  V3 w{coeff[0], coeff[1], coeff[2]};
  r[0] = sinfo * (dot(w, qrot(Qepi, p) + tepi) + coeff[3]);
  if (!Jp && !Ji && !Je) return;
  M3 Rp = qmat(Qp), Ri = qmat(Qi), Re = qmat(Qe);
  M3 RpT = transpose(Rp);
  V3 wRpT = vecmat(w, RpT);  // w^T Rp^T
  if (Jp) {                  // :61-72
    V3 v = matvec(Ri, matvec(Re, p)) + matvec(Ri, te) + ti - tp;
    V3 jr = vecmat(w, matmul(RpT, skew(v)));
    set_row(Jp, sinfo, -wRpT, jr);
  }
  if (Ji) {  // :74-85
    V3 jr = -vecmat(vecmat(wRpT, Ri), skew(matvec(Re, p) + te));
    set_row(Ji, sinfo, wRpT, jr);
  }
  if (Je) {  // :87-97
    V3 wRpTRi = vecmat(wRpT, Ri);
    V3 jr = -vecmat(wRpTRi, skew(matvec(Re, p)));
    set_row(Je, sinfo, wRpTRi, jr);
  }
}

// LidarPureOdomEdgeFactor::Evaluate lidar_pure_odom_factor.hpp:209-281
inline void odom_edge_factor(const V3 &p, const double *coeff, double sinfo, const double *xp, const double *xi,
                             const double *xe, double *r, double *Jp, double *Ji, double *Je) {
  Q4 Qp{xp[3], xp[4], xp[5], xp[6]}, Qi{xi[3], xi[4], xi[5], xi[6]}, Qe{xe[3], xe[4], xe[5], xe[6]};
  V3 tp{xp[0], xp[1], xp[2]}, ti{xi[0], xi[1], xi[2]}, te{xe[0], xe[1], xe[2]};
  Q4 Qpi = qmul(qconj(Qp), Qi);
  V3 tpi = qrot(qconj(Qp), ti - tp);
  Q4 Qepi = qmul(Qpi, Qe);
  V3 tepi = qrot(Qpi, te) + tpi;
  V3 lpa{coeff[0], coeff[1], coeff[2]}, lpb{coeff[3], coeff[4], coeff[5]};
  V3 lp = qrot(Qepi, p) + tepi;
  V3 nu = cross(lp - lpa, lp - lpb);
  V3 de = lpa - lpb;
  double nun = norm(nu), den = norm(de);
  r[0] = sinfo * nun / den;
  if (!Jp && !Ji && !Je) return;
  M3 Rp = qmat(Qp), Ri = qmat(Qi), Re = qmat(Qe);
  M3 RpT = transpose(Rp);
  V3 eta = (1.0 / den) * normalized(nu);
  V3 ba = lp - lpa, bb = lp - lpb;
  V3 eS = vecmat(eta, skew(ba - bb));  // eta [ba-bb]x
  V3 eSRpT = vecmat(eS, RpT);
  if (Jp) {  // :243-253
    V3 v = matvec(RpT, matvec(Ri, matvec(Re, p)) + matvec(Ri, te) + ti - tp);
    set_row(Jp, sinfo, -eSRpT, vecmat(eS, skew(v)));
  }
  if (Ji) {  // :255-266
    V3 jr = -vecmat(vecmat(eSRpT, Ri), skew(matvec(Re, p) + te));
    set_row(Ji, sinfo, eSRpT, jr);
  }
  if (Je) {  // :268-279
    V3 eSRpTRi = vecmat(eSRpT, Ri);
    M3 A = matmul(Re, skew(p));
    M3 B = skew(te);
    M3 AB;
    for (int k = 0; k < 9; k++) AB.m[k] = A.m[k] + B.m[k];
    set_row(Je, sinfo, eSRpTRi, -vecmat(eSRpTRi, AB));
  }
}

// ceres::HuberLoss(a): rho(s) and rho'(s).  Ceres 1.12 (un-vendored, docker/Dockerfile:3);
// the corrector is restated in-tree at marginalization_factor.cpp:50-81: for rho'' <= 0 the block is
// scaled by sqrt(rho').
inline void huber(double a, double s, double *rho, double *rho1) {
  const double b = a * a;
  if (s > b) {
    const double r = std::sqrt(s);
    *rho = 2 * a * r - b;
    *rho1 = std::max(std::numeric_limits<double>::min(), a / r);
  } else {
    *rho = s;
    *rho1 = 1.0;
  }
}

// PoseLocalParameterization, pose_local_parameterization.{h,cpp}
struct PoseLocalParameterization {
  bool is_degenerate = false;
  double V_update[36];
  PoseLocalParameterization() { setParameter(); }
  void setParameter() {  // .cpp:16-20
    is_degenerate = false;
    for (int i = 0; i < 36; i++) V_update[i] = (i % 7 == 0) ? 1.0 : 0.0;
  }
  // .cpp:26-46   p+ = p + (V d)[0:3];  q+ = normalize(q * [1, (V d)[3:6]/2])
  bool Plus(const double *x, const double *delta, double *xpd) const {
    double dx[6];
    for (int i = 0; i < 6; i++) {
      double s = 0;
      for (int j = 0; j < 6; j++) s += V_update[i * 6 + j] * delta[j];
      dx[i] = s;
    }
    xpd[0] = x[0] + dx[0], xpd[1] = x[1] + dx[1], xpd[2] = x[2] + dx[2];
    Q4 dq{dx[3] / 2.0, dx[4] / 2.0, dx[5] / 2.0, 1.0};  // Utility::deltaQ utility.h:173-185
    Q4 qn = qnormalized(qmul(Q4{x[3], x[4], x[5], x[6]}, dq));
    xpd[3] = qn.x, xpd[4] = qn.y, xpd[5] = qn.z, xpd[6] = qn.w;
    return true;
  }
  // .cpp:50-56   7x6 row-major [I6; 0]
  bool ComputeJacobian(const double *, double *J) const {
    for (int i = 0; i < 42; i++) J[i] = 0;
    for (int i = 0; i < 6; i++) J[i * 6 + i] = 1;
    return true;
  }
  int GlobalSize() const { return 7; }
  int LocalSize() const { return 6; }
};

}  // namespace orc
