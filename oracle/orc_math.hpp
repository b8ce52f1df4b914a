// ORACLE — TEST INFRASTRUCTURE ONLY.
// CPU restatement of the arithmetic M-LOAM's per-scan hot path delegates to Eigen3
// (un-vendored third party; ROS-melodic apt package, unpinned — SURVEY.md §8c).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, load or call anything under oracle/.  The product (m-loam_b200/)
// never includes this file.
//
// PARITY UNPINNED: the reference ships no golden vectors for this path and Eigen is not
// installed here, so the small dense routines below restate the *published* algorithms
// (cyclic Jacobi for the symmetric eigenproblems, column-pivoted Householder QR for the
// 5x3 least squares).  They define the behaviour the CUDA kernels are checked against.
//
// Everything is compiled with -ffp-contract=off so each float/double operation rounds
// exactly once, which is what lets the CUDA path (compiled with -fmad=false) reproduce
// gate decisions bit for bit.
#pragma once
#include <cmath>
#include <limits>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <vector>

namespace orc {

// ----------------------------------------------------------------------------- double 3-vectors
struct V3 {
  double x, y, z;
};
inline V3 operator+(const V3 &a, const V3 &b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(const V3 &a, const V3 &b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, const V3 &a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator-(const V3 &a) { return {-a.x, -a.y, -a.z}; }
inline double dot(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(const V3 &a, const V3 &b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double norm(const V3 &a) { return std::sqrt(dot(a, a)); }

// row-major 3x3
struct M3 {
  double m[9];
  double &operator()(int r, int c) { return m[r * 3 + c]; }
  double operator()(int r, int c) const { return m[r * 3 + c]; }
};
inline M3 matmul(const M3 &A, const M3 &B) {
  M3 C;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C(i, j) = A(i, 0) * B(0, j) + A(i, 1) * B(1, j) + A(i, 2) * B(2, j);
  return C;
}
inline M3 transpose(const M3 &A) {
  M3 T;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) T(i, j) = A(j, i);
  return T;
}
inline V3 matvec(const M3 &A, const V3 &v) {
  return {A(0, 0) * v.x + A(0, 1) * v.y + A(0, 2) * v.z, A(1, 0) * v.x + A(1, 1) * v.y + A(1, 2) * v.z,
          A(2, 0) * v.x + A(2, 1) * v.y + A(2, 2) * v.z};
}
// row-vector times matrix: (v^T A)
inline V3 vecmat(const V3 &v, const M3 &A) {
  return {v.x * A(0, 0) + v.y * A(1, 0) + v.z * A(2, 0), v.x * A(0, 1) + v.y * A(1, 1) + v.z * A(2, 1),
          v.x * A(0, 2) + v.y * A(1, 2) + v.z * A(2, 2)};
}
// Utility::skewSymmetric, estimator/src/utility/utility.h:187-195
inline M3 skew(const V3 &q) { return M3{{0, -q.z, q.y, q.z, 0, -q.x, -q.y, q.x, 0}}; }

// ----------------------------------------------------------------------------- quaternion (x,y,z,w)
struct Q4 {
  double x, y, z, w;
};
// Hamilton product, Eigen::Quaterniond operator* convention (used at pose.cpp:110-113)
inline Q4 qmul(const Q4 &a, const Q4 &b) {
  return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
          a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
inline Q4 qconj(const Q4 &q) { return {-q.x, -q.y, -q.z, q.w}; }
inline Q4 qnormalized(const Q4 &q) {
  double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  return {q.x / n, q.y / n, q.z / n, q.w / n};
}
// q * v as Eigen computes it: v + 2w(u x v) + 2 u x (u x v)
inline V3 qrot(const Q4 &q, const V3 &v) {
  V3 u{q.x, q.y, q.z};
  V3 uv = cross(u, v);
  uv = uv + uv;
  return v + q.w * uv + cross(u, uv);
}
inline M3 qmat(const Q4 &q) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  return M3{{1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx,
             1 - (txx + tyy)}};
}

// Pose = (q, t); estimator/src/estimator/pose.{h,cpp}.  Pose(q,t) normalises q (pose.cpp:34-41).
struct Pose {
  Q4 q{0, 0, 0, 1};
  V3 t{0, 0, 0};
};
inline Pose make_pose(const Q4 &q, const V3 &t) { return Pose{qnormalized(q), t}; }
// pose.cpp:110-113
inline Pose pose_mul(const Pose &a, const Pose &b) { return make_pose(qmul(a.q, b.q), qrot(a.q, b.t) + a.t); }
// pose.cpp:99-102
inline Pose pose_inv(const Pose &a) {
  Q4 qi = qconj(a.q);
  return make_pose(qi, -qrot(qi, a.t));
}
// parameter block [tx ty tz qx qy qz qw]  (pose_local_parameterization.h:20)
inline void pose_to_param(const Pose &p, double *x) {
  x[0] = p.t.x, x[1] = p.t.y, x[2] = p.t.z, x[3] = p.q.x, x[4] = p.q.y, x[5] = p.q.z, x[6] = p.q.w;
}
inline Pose pose_from_param_raw(const double *x) { return Pose{Q4{x[3], x[4], x[5], x[6]}, V3{x[0], x[1], x[2]}}; }

// ----------------------------------------------------------------------------- float helpers
// common::sqrSum, mloam_common/libs/include/common/algos/math.hpp:11-14
inline float sqrSumf(float a, float b, float c) { return a * a + b * b + c * c; }

// Symmetric 3x3 eigen-decomposition in float (cyclic Jacobi, fixed sweep schedule).
// Replaces Eigen::SelfAdjointEigenSolver<Matrix3f> (feature_extract.hpp:427,688).
// A is row-major symmetric. Output: eigenvalues ascending in w[3], eigenvectors as COLUMNS of V.
inline void eig3f(const float Ain[9], float w[3], float V[9]) {
  float a[3][3] = {{Ain[0], Ain[1], Ain[2]}, {Ain[1], Ain[4], Ain[5]}, {Ain[2], Ain[5], Ain[8]}};
  float v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 12; sweep++) {
    float off = std::fabs(a[0][1]) + std::fabs(a[0][2]) + std::fabs(a[1][2]);
    float diag = std::fabs(a[0][0]) + std::fabs(a[1][1]) + std::fabs(a[2][2]);
    if (off <= 1e-12f * diag || off == 0.0f) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        float apq = a[p][q];
        if (apq == 0.0f) continue;
        float theta = (a[q][q] - a[p][p]) / (2.0f * apq);
        float t = 1.0f / (std::fabs(theta) + std::sqrt(theta * theta + 1.0f));
        if (theta < 0.0f) t = -t;
        float c = 1.0f / std::sqrt(t * t + 1.0f);
        float s = t * c;
        // A <- J^T A J
        a[p][p] = a[p][p] - t * apq;
        a[q][q] = a[q][q] + t * apq;
        a[p][q] = 0.0f;
        a[q][p] = 0.0f;
        int r = 3 - p - q;
        float arp = a[r][p], arq = a[r][q];
        a[r][p] = c * arp - s * arq;
        a[p][r] = a[r][p];
        a[r][q] = s * arp + c * arq;
        a[q][r] = a[r][q];
        for (int k = 0; k < 3; k++) {
          float vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int idx[3] = {0, 1, 2};
  float d[3] = {a[0][0], a[1][1], a[2][2]};
  // ascending, stable
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2 - i; j++)
      if (d[idx[j + 1]] < d[idx[j]]) std::swap(idx[j], idx[j + 1]);
  for (int j = 0; j < 3; j++) {
    w[j] = d[idx[j]];
    for (int k = 0; k < 3; k++) V[k * 3 + j] = v[k][idx[j]];
  }
}

// Least squares  min || A n - b ||  for A (K x 3, row-major, float), b = -1, via column-pivoted
// Householder QR.  Replaces mat_A.colPivHouseholderQr().solve(mat_B) (feature_extract.hpp:579,823).
// Column norms are recomputed directly at each step (K<=16, 3 columns) instead of down-dated.
// Returns false when R is numerically rank deficient.
inline bool lsq_plane_f(const float *Arow, int K, float n[3]) {
  float A[16][3];
  float b[16];
  for (int i = 0; i < K; i++) {
    A[i][0] = Arow[i * 3 + 0], A[i][1] = Arow[i * 3 + 1], A[i][2] = Arow[i * 3 + 2];
    b[i] = -1.0f;
  }
  int perm[3] = {0, 1, 2};
  float maxpivot = 0.0f;
  for (int k = 0; k < 3; k++) {
    // pivot: largest remaining column norm
    int best = k;
    float bestn = -1.0f;
    for (int j = k; j < 3; j++) {
      float s = 0.0f;
      for (int i = k; i < K; i++) s = s + A[i][j] * A[i][j];
      if (s > bestn) bestn = s, best = j;
    }
    if (best != k) {
      for (int i = 0; i < K; i++) std::swap(A[i][k], A[i][best]);
      std::swap(perm[k], perm[best]);
    }
    // Householder on A[k:K, k]
    float c0 = A[k][k];
    float tail = 0.0f;
    for (int i = k + 1; i < K; i++) tail = tail + A[i][k] * A[i][k];
    float tau, beta;
    float ess[16];
    if (tail <= 1.17549435e-38f) {
      tau = 0.0f;
      beta = c0;
      for (int i = k + 1; i < K; i++) ess[i] = 0.0f;
    } else {
      beta = std::sqrt(c0 * c0 + tail);
      if (c0 >= 0.0f) beta = -beta;
      float den = c0 - beta;
      for (int i = k + 1; i < K; i++) ess[i] = A[i][k] / den;
      tau = (beta - c0) / beta;
    }
    A[k][k] = beta;
    if (std::fabs(beta) > maxpivot) maxpivot = std::fabs(beta);
    // apply H = I - tau v v^T (v = [1; ess]) to remaining columns and to b
    for (int j = k + 1; j < 3; j++) {
      float s = A[k][j];
      for (int i = k + 1; i < K; i++) s = s + ess[i] * A[i][j];
      s = tau * s;
      A[k][j] = A[k][j] - s;
      for (int i = k + 1; i < K; i++) A[i][j] = A[i][j] - s * ess[i];
    }
    {
      float s = b[k];
      for (int i = k + 1; i < K; i++) s = s + ess[i] * b[i];
      s = tau * s;
      b[k] = b[k] - s;
      for (int i = k + 1; i < K; i++) b[i] = b[i] - s * ess[i];
    }
  }
  // rank test (Eigen: |pivot| > eps * diagSize * maxpivot)
  const float thr = 1.1920929e-7f * 3.0f * maxpivot;
  for (int k = 0; k < 3; k++)
    if (!(std::fabs(A[k][k]) > thr)) return false;
  // back substitution R y = c
  float y[3];
  y[2] = b[2] / A[2][2];
  y[1] = (b[1] - A[1][2] * y[2]) / A[1][1];
  y[0] = (b[0] - A[0][1] * y[1] - A[0][2] * y[2]) / A[0][0];
  n[perm[0]] = y[0];
  n[perm[1]] = y[1];
  n[perm[2]] = y[2];
  return true;
}

// ----------------------------------------------------------------------------- small dense double
// Symmetric NxN eigen-decomposition (cyclic Jacobi), ascending eigenvalues, eigenvectors in columns
// of V (row-major N x N).  Replaces Eigen::SelfAdjointEigenSolver<Matrix<double,6,6>>
// (lidar_mapper_keyframe.cpp:1174, estimator.cpp:1613, lidar_tracker.cpp:140).
inline void eig_sym(int N, const double *Ain, double *w, double *V) {
  std::vector<double> a(Ain, Ain + N * N), v(N * N, 0.0);
  for (int i = 0; i < N; i++) v[i * N + i] = 1.0;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0, diag = 0;
    for (int i = 0; i < N; i++) {
      diag += std::fabs(a[i * N + i]);
      for (int j = i + 1; j < N; j++) off += std::fabs(a[i * N + j]);
    }
    if (off <= 1e-22 * diag || off == 0.0) break;
    for (int p = 0; p < N - 1; p++)
      for (int q = p + 1; q < N; q++) {
        double apq = a[p * N + q];
        if (apq == 0.0) continue;
        double theta = (a[q * N + q] - a[p * N + p]) / (2.0 * apq);
        double t = 1.0 / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        if (theta < 0.0) t = -t;
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        a[p * N + p] -= t * apq;
        a[q * N + q] += t * apq;
        a[p * N + q] = a[q * N + p] = 0.0;
        for (int r = 0; r < N; r++) {
          if (r == p || r == q) continue;
          double arp = a[r * N + p], arq = a[r * N + q];
          a[r * N + p] = a[p * N + r] = c * arp - s * arq;
          a[r * N + q] = a[q * N + r] = s * arp + c * arq;
        }
        for (int k = 0; k < N; k++) {
          double vkp = v[k * N + p], vkq = v[k * N + q];
          v[k * N + p] = c * vkp - s * vkq;
          v[k * N + q] = s * vkp + c * vkq;
        }
      }
  }
  std::vector<int> idx(N);
  for (int i = 0; i < N; i++) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](int i, int j) { return a[i * N + i] < a[j * N + j]; });
  for (int j = 0; j < N; j++) {
    w[j] = a[idx[j] * N + idx[j]];
    for (int k = 0; k < N; k++) V[k * N + j] = v[k * N + idx[j]];
  }
}

// Cholesky A = L L^T (row-major, lower in place).  Returns false if not positive definite.
inline bool cholesky(int N, double *A) {
  for (int j = 0; j < N; j++) {
    double d = A[j * N + j];
    for (int k = 0; k < j; k++) d -= A[j * N + k] * A[j * N + k];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    A[j * N + j] = d;
    for (int i = j + 1; i < N; i++) {
      double s = A[i * N + j];
      for (int k = 0; k < j; k++) s -= A[i * N + k] * A[j * N + k];
      A[i * N + j] = s / d;
    }
  }
  return true;
}
inline void cholesky_solve(int N, const double *L, const double *b, double *x) {
  std::vector<double> y(N);
  for (int i = 0; i < N; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= L[i * N + k] * y[k];
    y[i] = s / L[i * N + i];
  }
  for (int i = N - 1; i >= 0; i--) {
    double s = y[i];
    for (int k = i + 1; k < N; k++) s -= L[k * N + i] * x[k];
    x[i] = s / L[i * N + i];
  }
}
// common::logDet(H, true): log-determinant through LLT; mloam_common/.../math.hpp:172-202
inline double logdet_chol(int N, const double *H) {
  std::vector<double> L(H, H + N * N);
  if (!cholesky(N, L.data())) return -std::numeric_limits<double>::infinity();
  double s = 0;
  for (int i = 0; i < N; i++) s += std::log(L[i * N + i]);
  return 2.0 * s;
}

}  // namespace orc
