// fit.cuh — per-feature local geometry fits in float, register resident.
//  * line fit: mean + 3x3 scatter + symmetric eigen-decomposition  (feature_extract.hpp:670-693)
//  * plane fit: K x 3 least squares A n = -1 by column-pivoted Householder QR (feature_extract.hpp:817-825)
// Operation order is fixed (and -fmad=false) so results are reproducible to the bit across launches and
// against the CPU oracle's statement of the same algorithms.
#pragma once
#include "common.cuh"

namespace mloam {

// Cyclic Jacobi, ascending eigenvalues w[3], eigenvectors in the COLUMNS of v (v[row][col]).
__device__ __forceinline__ void eig3f_dev(float a00, float a01, float a02, float a11, float a12, float a22, float w[3],
                                          float v[3][3]) {
  float a[3][3] = {{a00, a01, a02}, {a01, a11, a12}, {a02, a12, a22}};
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) v[i][j] = (i == j) ? 1.0f : 0.0f;
  for (int sweep = 0; sweep < 12; sweep++) {
    const float off = fabsf(a[0][1]) + fabsf(a[0][2]) + fabsf(a[1][2]);
    const float diag = fabsf(a[0][0]) + fabsf(a[1][1]) + fabsf(a[2][2]);
    if (off <= 1e-12f * diag || off == 0.0f) break;
#pragma unroll
    for (int p = 0; p < 2; p++) {
#pragma unroll
      for (int q = p + 1; q < 3; q++) {
        const float apq = a[p][q];
        if (apq != 0.0f) {
          const float theta = (a[q][q] - a[p][p]) / (2.0f * apq);
          float t = 1.0f / (fabsf(theta) + sqrtf(theta * theta + 1.0f));
          if (theta < 0.0f) t = -t;
          const float c = 1.0f / sqrtf(t * t + 1.0f);
          const float s = t * c;
          a[p][p] = a[p][p] - t * apq;
          a[q][q] = a[q][q] + t * apq;
          a[p][q] = 0.0f;
          a[q][p] = 0.0f;
          const int r = 3 - p - q;
          const float arp = a[r][p], arq = a[r][q];
          a[r][p] = c * arp - s * arq;
          a[p][r] = a[r][p];
          a[r][q] = s * arp + c * arq;
          a[q][r] = a[r][q];
#pragma unroll
          for (int k = 0; k < 3; k++) {
            const float vkp = v[k][p], vkq = v[k][q];
            v[k][p] = c * vkp - s * vkq;
            v[k][q] = s * vkp + c * vkq;
          }
        }
      }
    }
  }
  w[0] = a[0][0], w[1] = a[1][1], w[2] = a[2][2];
  // ascending, stable (compare-swap sequence (0,1) (1,2) (0,1))
#define MLOAM_CSWAP(i, j)                                  \
  if (w[j] < w[i]) {                                       \
    float tw = w[i];                                       \
    w[i] = w[j];                                           \
    w[j] = tw;                                             \
    _Pragma("unroll") for (int k = 0; k < 3; k++) {        \
      float tv = v[k][i];                                  \
      v[k][i] = v[k][j];                                   \
      v[k][j] = tv;                                        \
    }                                                      \
  }
  MLOAM_CSWAP(0, 1)
  MLOAM_CSWAP(1, 2)
  MLOAM_CSWAP(0, 1)
#undef MLOAM_CSWAP
}

// Least squares min ||A n + 1|| for A (K x 3).  Returns false if numerically rank deficient.
template <int K>
__device__ __forceinline__ bool lsq_plane_dev(float A[K][3], float n[3]) {
  float b[K];
#pragma unroll
  for (int i = 0; i < K; i++) b[i] = -1.0f;
  int p0 = 0, p1 = 1, p2 = 2;  // column permutation
  float maxpivot = 0.0f;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    // pivot: largest remaining column norm (first maximum wins)
    int best = k;
    float bestn = -1.0f;
#pragma unroll
    for (int j = k; j < 3; j++) {
      float s = 0.0f;
#pragma unroll
      for (int i = k; i < K; i++) s = s + A[i][j] * A[i][j];
      if (s > bestn) bestn = s, best = j;
    }
    if (k == 0) {
      if (best == 1) {
#pragma unroll
        for (int i = 0; i < K; i++) { float t = A[i][0]; A[i][0] = A[i][1]; A[i][1] = t; }
        int t = p0; p0 = p1; p1 = t;
      } else if (best == 2) {
#pragma unroll
        for (int i = 0; i < K; i++) { float t = A[i][0]; A[i][0] = A[i][2]; A[i][2] = t; }
        int t = p0; p0 = p2; p2 = t;
      }
    } else if (k == 1) {
      if (best == 2) {
#pragma unroll
        for (int i = 0; i < K; i++) { float t = A[i][1]; A[i][1] = A[i][2]; A[i][2] = t; }
        int t = p1; p1 = p2; p2 = t;
      }
    }
    const float c0 = A[k][k];
    float tail = 0.0f;
#pragma unroll
    for (int i = k + 1; i < K; i++) tail = tail + A[i][k] * A[i][k];
    float tau, beta;
    float ess[K];
    if (tail <= 1.17549435e-38f) {
      tau = 0.0f;
      beta = c0;
#pragma unroll
      for (int i = 0; i < K; i++) ess[i] = 0.0f;
    } else {
      beta = sqrtf(c0 * c0 + tail);
      if (c0 >= 0.0f) beta = -beta;
      const float den = c0 - beta;
#pragma unroll
      for (int i = 0; i < K; i++) ess[i] = (i > k) ? A[i][k] / den : 0.0f;
      tau = (beta - c0) / beta;
    }
    A[k][k] = beta;
    if (fabsf(beta) > maxpivot) maxpivot = fabsf(beta);
#pragma unroll
    for (int j = k + 1; j < 3; j++) {
      float s = A[k][j];
#pragma unroll
      for (int i = k + 1; i < K; i++) s = s + ess[i] * A[i][j];
      s = tau * s;
      A[k][j] = A[k][j] - s;
#pragma unroll
      for (int i = k + 1; i < K; i++) A[i][j] = A[i][j] - s * ess[i];
    }
    {
      float s = b[k];
#pragma unroll
      for (int i = k + 1; i < K; i++) s = s + ess[i] * b[i];
      s = tau * s;
      b[k] = b[k] - s;
#pragma unroll
      for (int i = k + 1; i < K; i++) b[i] = b[i] - s * ess[i];
    }
  }
  const float thr = 1.1920929e-7f * 3.0f * maxpivot;
  if (!(fabsf(A[0][0]) > thr) || !(fabsf(A[1][1]) > thr) || !(fabsf(A[2][2]) > thr)) return false;
  const float y2 = b[2] / A[2][2];
  const float y1 = (b[1] - A[1][2] * y2) / A[1][1];
  const float y0 = (b[0] - A[0][1] * y1 - A[0][2] * y2) / A[0][0];
  // n[perm[k]] = y[k]
  n[0] = (p0 == 0) ? y0 : ((p1 == 0) ? y1 : y2);
  n[1] = (p0 == 1) ? y0 : ((p1 == 1) ? y1 : y2);
  n[2] = (p0 == 2) ? y0 : ((p1 == 2) ? y1 : y2);
  return true;
}

}  // namespace mloam
