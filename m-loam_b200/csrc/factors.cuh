// factors.cuh — residuals and analytic Jacobians of the M-LOAM LiDAR factors, double precision,
// one thread per factor.  Rows are [d/dt (3) | d/dtheta (3)] per pose block; the 7th column of the
// reference's 1x7 rows is identically zero and is only materialised by the batch-evaluate kernel.
//   plane     : lidar_map_factor.hpp:44-68, lidar_scan_factor.hpp:33-60, lidar_online_calib_factor.hpp:34-60
//   edge      : lidar_map_factor.hpp:143-171, lidar_scan_factor.hpp:139-168, lidar_online_calib_factor.hpp:135-163
//   edge vec  : lidar_scan_factor.hpp:245-279
//   odom plane: lidar_pure_odom_factor.hpp:38-101        odom edge: lidar_pure_odom_factor.hpp:209-281
#pragma once
#include "common.cuh"

namespace mloam {

struct PoseR {  // pose with its rotation matrix cached
  Q4 q;
  D3 t;
  M33 R;
};
__device__ __forceinline__ PoseR make_poser(const double *x) {
  PoseR P;
  P.q = Q4{x[3], x[4], x[5], x[6]};
  P.t = D3{x[0], x[1], x[2]};
  P.R = qmat(P.q);
  return P;
}

// r = s (w.(Rp+t) + d);  J = s [w^T | -w^T R [p]x]
__device__ __forceinline__ double plane_factor(const PoseR &P, const D3 &p, const D3 &w, double d, double s, double J[6],
                                               bool want_j) {
  const double a = dot(w, qrot(P.q, p) + P.t) + d;
  if (want_j) {
    const D3 jr = neg(vec_skew(vecmat(w, P.R), p));
    J[0] = s * w.x, J[1] = s * w.y, J[2] = s * w.z, J[3] = s * jr.x, J[4] = s * jr.y, J[5] = s * jr.z;
  }
  return s * a;
}

// r = s |nu| / |de|,  nu = (lp-a)x(lp-b), de = a-b;  eta = nu^/|de|;  J = s [-eta [de]x | eta [de]x R [p]x]
// nu.normalized() follows Eigen 3.3: a zero vector stays zero.
__device__ __forceinline__ double edge_factor(const PoseR &P, const D3 &p, const D3 &lpa, const D3 &lpb, double s,
                                              double J[6], bool want_j) {
  const D3 lp = qrot(P.q, p) + P.t;
  const D3 nu = cross(lp - lpa, lp - lpb);
  const D3 de = lpa - lpb;
  const double nun = norm(nu), den = norm(de);
  if (want_j) {
    const D3 nh = nun > 0.0 ? D3{nu.x / nun, nu.y / nun, nu.z / nun} : nu;
    const D3 eta = (1.0 / den) * nh;
    const D3 eS = vec_skew(eta, de);
    const D3 jr = vec_skew(vecmat(eS, P.R), p);
    J[0] = s * -eS.x, J[1] = s * -eS.y, J[2] = s * -eS.z, J[3] = s * jr.x, J[4] = s * jr.y, J[5] = s * jr.z;
  }
  return s * nun / den;
}

// r = nu/|de| (3);  J = 1/|de| [-[de]x | [de]x R [p]x]   rows i: J[i*6 + k]
__device__ __forceinline__ void edge_vector_factor(const PoseR &P, const D3 &p, const D3 &lpa, const D3 &lpb, double r[3],
                                                   double J[18], bool want_j) {
  const D3 lp = qrot(P.q, p) + P.t;
  const D3 nu = cross(lp - lpa, lp - lpb);
  const D3 de = lpa - lpb;
  const double den = norm(de);
  r[0] = nu.x / den, r[1] = nu.y / den, r[2] = nu.z / den;
  if (want_j) {
    const double eta = 1.0 / den;
    // rows of [de]x
    const D3 S0{0.0, -de.z, de.y}, S1{de.z, 0.0, -de.x}, S2{-de.y, de.x, 0.0};
    const D3 rows[3] = {S0, S1, S2};
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const D3 srp = vec_skew(vecmat(rows[i], P.R), p);
      J[i * 6 + 0] = -eta * rows[i].x, J[i * 6 + 1] = -eta * rows[i].y, J[i * 6 + 2] = -eta * rows[i].z;
      J[i * 6 + 3] = eta * srp.x, J[i * 6 + 4] = eta * srp.y, J[i * 6 + 5] = eta * srp.z;
    }
  }
}

// Three-pose chain: lp = Rp^T (Ri (Re p + te) + ti - tp).  Jp/Ji/Je: 6 each (nullable).
struct Chain {
  PoseR P, I, E;
  Q4 Qepi;
  D3 tepi;
};
__device__ __forceinline__ Chain make_chain(const double *xp, const double *xi, const double *xe) {
  Chain c;
  c.P = make_poser(xp), c.I = make_poser(xi), c.E = make_poser(xe);
  const Q4 Qpi = qmul(qconj(c.P.q), c.I.q);
  const D3 tpi = qrot(qconj(c.P.q), c.I.t - c.P.t);
  c.Qepi = qmul(Qpi, c.E.q);
  c.tepi = qrot(Qpi, c.E.t) + tpi;
  return c;
}
__device__ __forceinline__ void put6(double *J, double s, const D3 &a, const D3 &b) {
  J[0] = s * a.x, J[1] = s * a.y, J[2] = s * a.z, J[3] = s * b.x, J[4] = s * b.y, J[5] = s * b.z;
}
__device__ __forceinline__ double odom_plane_factor(const Chain &c, const D3 &p, const D3 &w, double d, double s, double *Jp,
                                                    double *Ji, double *Je) {
  const double r = dot(w, qrot(c.Qepi, p) + c.tepi) + d;
  if (Jp || Ji || Je) {
    const D3 wRpT = vecmatT(w, c.P.R);  // w^T Rp^T
    if (Jp) {
      const D3 v = matvec(c.I.R, matvec(c.E.R, p)) + matvec(c.I.R, c.E.t) + c.I.t - c.P.t;
      put6(Jp, s, neg(wRpT), vec_skew(wRpT, v));
    }
    if (Ji) {
      const D3 jr = neg(vec_skew(vecmat(wRpT, c.I.R), matvec(c.E.R, p) + c.E.t));
      put6(Ji, s, wRpT, jr);
    }
    if (Je) {
      const D3 wi = vecmat(wRpT, c.I.R);
      put6(Je, s, wi, neg(vec_skew(wi, matvec(c.E.R, p))));
    }
  }
  return s * r;
}
__device__ __forceinline__ double odom_edge_factor(const Chain &c, const D3 &p, const D3 &lpa, const D3 &lpb, double s,
                                                   double *Jp, double *Ji, double *Je) {
  const D3 lp = qrot(c.Qepi, p) + c.tepi;
  const D3 nu = cross(lp - lpa, lp - lpb);
  const D3 de = lpa - lpb;
  const double nun = norm(nu), den = norm(de);
  if (Jp || Ji || Je) {
    const D3 nh = nun > 0.0 ? D3{nu.x / nun, nu.y / nun, nu.z / nun} : nu;
    const D3 eta = (1.0 / den) * nh;
    const D3 ba = lp - lpa, bb = lp - lpb;
    const D3 eS = vec_skew(eta, ba - bb);
    const D3 eSRpT = vecmatT(eS, c.P.R);
    if (Jp) {
      const D3 v = vecmat(matvec(c.I.R, matvec(c.E.R, p)) + matvec(c.I.R, c.E.t) + c.I.t - c.P.t, c.P.R);  // Rp^T (.)
      put6(Jp, s, neg(eSRpT), vec_skew(eS, v));
    }
    if (Ji) {
      const D3 jr = neg(vec_skew(vecmat(eSRpT, c.I.R), matvec(c.E.R, p) + c.E.t));
      put6(Ji, s, eSRpT, jr);
    }
    if (Je) {
      const D3 ei = vecmat(eSRpT, c.I.R);
      // -ei (Re [p]x + [te]x)
      const D3 a = vec_skew(vecmat(ei, c.E.R), p);
      const D3 b = vec_skew(ei, c.E.t);
      put6(Je, s, ei, neg(a + b));
    }
  }
  return s * nun / den;
}

// ceres::HuberLoss(a): rho(s), rho'(s)  (corrector for rho'' <= 0: scale block by sqrt(rho'); restated
// in-tree at marginalization_factor.cpp:50-81)
__device__ __forceinline__ void huber(double a, double s, double *rho, double *rho1) {
  const double b = a * a;
  if (s > b) {
    const double r = sqrt(s);
    *rho = 2 * a * r - b;
    *rho1 = fmax(2.2250738585072014e-308, a / r);
  } else {
    *rho = s;
    *rho1 = 1.0;
  }
}

}  // namespace mloam
