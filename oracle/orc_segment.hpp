// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).
// ImageSegmenter::segmentCloud with scan_info.segment_flag_ == false (segment_cloud: 0 — the KITTI / Oxford / UTBM configurations):
// projectCloud (estimator/src/imageSegmenter/image_segmenter.hpp:88-136) + the ring-order output and ScanInfo (:381-389).  With the flag
// off the labels computed in between (:160-360) do not touch the output.  ImageSegmenter::setParameter: image_segmenter.cpp:18-63.
// Arithmetic types as written in the reference: `range`, `vertical_angle`, `horizon_angle` are floats, sqrt / atan / atan2 take float
// arguments (the float overloads), `* 180 / M_PI` promotes to double before the float store; the row / column formulas mix float and
// double literals exactly as below.
#pragma once
#include "orc_cloud.hpp"
#include <cfloat>

namespace orc {

struct SegmenterParam {
  int vertical_scans = 64, horizon_scans = 2048;
  float ang_res_x = 0, ang_res_y = 0, ang_bottom = 0;
};
inline SegmenterParam segmenter_param(int vertical_scans, int horizon_scans) {  // image_segmenter.cpp:18-63
  SegmenterParam p;
  p.vertical_scans = vertical_scans, p.horizon_scans = horizon_scans;
  p.ang_res_x = 360.0 / horizon_scans;
  if (vertical_scans == 16) p.ang_res_y = 2.0, p.ang_bottom = 15.0 + 0.1;
  else if (vertical_scans == 32) p.ang_res_y = 41.33 / float(vertical_scans - 1), p.ang_bottom = 30.0 + 0.67;
  else p.ang_res_y = FLT_MAX;  // 64: the VLP-64 / HDL-64E row formula
  return p;
}

// row / column of one point, or false when projectCloud skips it (:99-124)
inline bool project_point(const SegmenterParam &sp, const PointI &point, double roi_range, int *row_id, int *column_id, float *range_out) {
  const float range = std::sqrt(point.x * point.x + point.y * point.y + point.z * point.z);
  if (range < roi_range) return false;
  const float vertical_angle = std::atan(point.z / std::sqrt(point.x * point.x + point.y * point.y)) * 180 / M_PI;
  int row;
  if (sp.vertical_scans == 64 && sp.ang_res_y == FLT_MAX) {
    if (vertical_angle >= -8.83) row = static_cast<int>((2 - vertical_angle) * 3.0 + 0.5);
    else row = static_cast<int>(sp.vertical_scans / 2) + static_cast<int>((-8.83 - vertical_angle) * 2.0 + 0.5);
    if (vertical_angle > 2 || vertical_angle < -24.33 || row > 50 || row < 0) return false;
  } else {
    row = static_cast<int>((vertical_angle + sp.ang_bottom) / sp.ang_res_y);
    if (row < 0 || row >= sp.vertical_scans) return false;
  }
  const float horizon_angle = std::atan2(point.x, point.y) * 180 / M_PI;
  int col = -std::round((horizon_angle - 90.0) / sp.ang_res_x) + sp.horizon_scans / 2;
  if (col >= sp.horizon_scans) col -= sp.horizon_scans;
  if (col < 0 || col >= sp.horizon_scans) return false;
  *row_id = row, *column_id = col, *range_out = range;
  return true;
}

// laser_cloud_out + ScanInfo of segmentCloud with segment_flag_ == false.  start / end sized vertical_scans.
inline void project_cloud(const Cloud &in, int vertical_scans, int horizon_scans, double roi_range, Cloud &out, std::vector<int> &scan_start,
                          std::vector<int> &scan_end) {
  const SegmenterParam sp = segmenter_param(vertical_scans, horizon_scans);
  std::vector<unsigned char> taken((size_t)vertical_scans * horizon_scans, 0);  // range_mat != FLT_MAX
  std::vector<Cloud> cloud_scan(vertical_scans);
  for (const PointI &p0 : in) {
    int row, col;
    float range;
    if (!project_point(sp, p0, roi_range, &row, &col, &range)) continue;
    if (taken[(size_t)row * horizon_scans + col]) continue;  // :118-119: the first point of a pixel wins
    taken[(size_t)row * horizon_scans + col] = 1;
    PointI p = p0;
    p.intensity += row;  // :121
    cloud_scan[row].push_back(p);
  }
  out.clear();
  scan_start.assign(vertical_scans, 0), scan_end.assign(vertical_scans, 0);
  for (int i = 0; i < vertical_scans; i++) {  // :381-387
    scan_start[i] = (int)out.size() + 5;
    out.insert(out.end(), cloud_scan[i].begin(), cloud_scan[i].end());
    scan_end[i] = (int)out.size() - 6;
  }
}

}  // namespace orc
