// Host build of the product's projection arithmetic (csrc/project.cuh, csrc/fd_atan.cuh) for the CPU tests: the same expressions the
// kernel evaluates, with IEEE host operations in place of the device intrinsics.  Build: g++ -O2 -ffp-contract=off -shared -fPIC.
#include "project.cuh"
#include <cmath>
#include <cstring>

extern "C" {
void host_project_pixels(const float *cloud, int n, int vertical_scans, int horizon_scans, double roi_range, int *pix) {
  const ProjectParam sp = project_param(vertical_scans, horizon_scans, roi_range);
  for (int i = 0; i < n; i++) {
    int row = 0;
    pix[i] = project_pixel(sp, float4{cloud[4 * i], cloud[4 * i + 1], cloud[4 * i + 2], cloud[4 * i + 3]}, &row);
  }
}
// number of arguments where fd::atanf / fd::atan2f differ from the C library's (bit comparison)
long host_fd_atan_mismatches(const float *x, const float *y, long n) {
  long bad = 0;
  for (long i = 0; i < n; i++) {
    const float a = std::atan2(x[i], y[i]), b = fd::atan2f(x[i], y[i]);
    const float c = std::atan(x[i]), d = fd::atanf(x[i]);
    bad += std::memcmp(&a, &b, 4) != 0;
    bad += std::memcmp(&c, &d, 4) != 0;
  }
  return bad;
}
}
