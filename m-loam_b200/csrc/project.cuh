// Range-image pixel of one point: the arithmetic of ImageSegmenter::projectCloud (image_segmenter.hpp:99-124), expression by expression,
// with explicitly rounded operations on the device (FD_* map to the __f*_rn / __d*_rn intrinsics there and to plain IEEE operations in a
// host build with -ffp-contract=off — tests/test_abi_cpu.py compiles this header for the host and compares it with the oracle).
#pragma once
#include "fd_atan.cuh"
#include <cmath>

#if defined(__CUDA_ARCH__)
#define FD_SQRT(a) __fsqrt_rn(a)
#define FD_DDIV(a, b) __ddiv_rn(a, b)
#define FD_DADD(a, b) __dadd_rn(a, b)
#define FD_DSUB(a, b) __dsub_rn(a, b)
#define FD_DMUL(a, b) __dmul_rn(a, b)
#else
#define FD_SQRT(a) sqrtf(a)
#define FD_DDIV(a, b) ((a) / (b))
#define FD_DADD(a, b) ((a) + (b))
#define FD_DSUB(a, b) ((a) - (b))
#define FD_DMUL(a, b) ((a) * (b))
#endif
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#if !defined(__CUDACC__)
struct float4 {
  float x, y, z, w;
};
#endif

struct ProjectParam {
  int vertical_scans, horizon_scans;
  float ang_res_x, ang_res_y, ang_bottom;
  int vlp64;
  double roi_range;
};
FD_HD int project_pixel(const ProjectParam &sp, float4 p, int *row_out) {
  const float xx = FD_MUL(p.x, p.x), yy = FD_MUL(p.y, p.y), zz = FD_MUL(p.z, p.z);
  const float range = FD_SQRT(FD_ADD(FD_ADD(xx, yy), zz));
  if ((double)range < sp.roi_range) return -1;
  const float t = fd::atanf(FD_DIV(p.z, FD_SQRT(FD_ADD(xx, yy))));
  const float vertical_angle = (float)FD_DDIV((double)FD_MUL(t, 180.0f), M_PI);
  int row;
  if (sp.vlp64) {
    const double va = (double)vertical_angle;
    if (!(va == va)) return -1;  // NaN: the reference's int cast gives INT_MIN -> row < 0 -> skipped
    if (va >= -8.83) row = (int)FD_DADD(FD_DMUL(FD_DSUB(2.0, va), 3.0), 0.5);
    else row = sp.vertical_scans / 2 + (int)FD_DADD(FD_DMUL(FD_DSUB(-8.83, va), 2.0), 0.5);
    if (va > 2.0 || va < -24.33 || row > 50 || row < 0) return -1;
  } else {
    const float q = FD_DIV(FD_ADD(vertical_angle, sp.ang_bottom), sp.ang_res_y);
    if (!(q > -1.0f && q < 2.0e9f)) return -1;  // NaN / out of int range: the reference's cast yields INT_MIN there -> row < 0 -> skipped
    row = (int)q;
    if (row < 0 || row >= sp.vertical_scans) return -1;
  }
  const float h = fd::atan2f(p.x, p.y);
  const float horizon_angle = (float)FD_DDIV((double)FD_MUL(h, 180.0f), M_PI);
  const double cq = FD_DDIV(FD_DSUB((double)horizon_angle, 90.0), (double)sp.ang_res_x);
  if (!(cq > -1.0e9 && cq < 1.0e9)) return -1;
  int col = (int)FD_DADD(-round(cq), (double)(sp.horizon_scans / 2));
  if (col >= sp.horizon_scans) col -= sp.horizon_scans;
  if (col < 0 || col >= sp.horizon_scans) return -1;
  *row_out = row;
  return row * sp.horizon_scans + col;
}
inline ProjectParam project_param(int vertical_scans, int horizon_scans, double roi_range) {  // image_segmenter.cpp:18-63
  ProjectParam sp;
  sp.vertical_scans = vertical_scans, sp.horizon_scans = horizon_scans, sp.roi_range = roi_range;
  sp.ang_res_x = (float)(360.0 / horizon_scans);
  sp.vlp64 = vertical_scans == 64;
  sp.ang_res_y = vertical_scans == 16 ? 2.0f : (vertical_scans == 32 ? (float)(41.33 / float(vertical_scans - 1)) : 0.f);
  sp.ang_bottom = vertical_scans == 16 ? (float)(15.0 + 0.1) : (vertical_scans == 32 ? (float)(30.0 + 0.67) : 0.f);
  return sp;
}
