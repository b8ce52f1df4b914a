// mloam_io.hpp — wire formats either side of the hot path, host-only, no ROS / PCL needed:
//
//   sensor_msgs/PointCloud2 <-> mloam_point_t[]   the buffers the ROS nodes exchange (rosNodeRVHercules.cpp, lidar_mapper_keyframe.cpp
//                                                :1258-1277) packed into / out of the 16-byte float4 layout of the C ABI
//   mloam_msgs/Extrinsics, mloam_msgs/Keyframes   (mloam_msgs/msg/*.msg): header + status + odometry / pose-with-covariance arrays as PODs
//   TUM trajectory dump                          save_statistics.hpp:104-119: "stamp x y z qx qy qz qw", stamp with 15 significant
//                                                digits, the rest with 8 (std::ostream::precision semantics, not fixed notation)
//
// PointCloud2 layout rules implemented (sensor_msgs/PointCloud2 definition): fields are located by NAME ("x", "y", "z",
// "intensity"), each with its byte offset and datatype (FLOAT32 = 7; FLOAT64 = 8 is converted); points are `point_step` bytes apart,
// rows `row_step` apart; is_bigendian data is byte-swapped; a missing "intensity" field gives 0 — the reference exits in that case
// (feature_extract.hpp:138-142), which callers can reproduce by checking PointCloud2View::has_intensity.
#pragma once
#include <cstdint>
#include <cstring>
#include <ostream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/mloam_b200.h"

namespace mloam {
namespace io {

enum PointFieldType { INT8 = 1, UINT8 = 2, INT16 = 3, UINT16 = 4, INT32 = 5, UINT32 = 6, FLOAT32 = 7, FLOAT64 = 8 };

struct PointField {  // sensor_msgs/PointField
  std::string name;
  uint32_t offset = 0;
  uint8_t datatype = FLOAT32;
  uint32_t count = 1;
};

struct PointCloud2View {  // the members of sensor_msgs/PointCloud2 the packing needs (data is not owned)
  uint32_t height = 1, width = 0;
  std::vector<PointField> fields;
  bool is_bigendian = false;
  uint32_t point_step = 0, row_step = 0;
  const uint8_t *data = nullptr;
  size_t data_size = 0;
  bool has_intensity = false;  // set by unpack
};

namespace detail {
inline float load_scalar(const uint8_t *p, uint8_t datatype, bool swap) {
  uint8_t b[8];
  const int n = datatype == FLOAT64 ? 8 : 4;
  for (int i = 0; i < n; i++) b[i] = swap ? p[n - 1 - i] : p[i];
  if (datatype == FLOAT64) {
    double d;
    std::memcpy(&d, b, 8);
    return (float)d;
  }
  float f;
  std::memcpy(&f, b, 4);
  return f;
}
inline bool host_is_bigendian() {
  const uint16_t one = 1;
  return *reinterpret_cast<const uint8_t *>(&one) == 0;
}
}  // namespace detail

// PointCloud2 -> float4 points (x, y, z, intensity).  Returns false when x / y / z are missing, a field is not FLOAT32 / FLOAT64, or
// the buffer is shorter than height * row_step.
inline bool unpackPointCloud2(PointCloud2View &msg, std::vector<mloam_point_t> &out) {
  const PointField *fx = nullptr, *fy = nullptr, *fz = nullptr, *fi = nullptr;
  for (const PointField &f : msg.fields) {
    if (f.name == "x") fx = &f;
    else if (f.name == "y") fy = &f;
    else if (f.name == "z") fz = &f;
    else if (f.name == "intensity") fi = &f;
  }
  msg.has_intensity = fi != nullptr;
  out.clear();
  if (!fx || !fy || !fz) return false;
  for (const PointField *f : {fx, fy, fz, fi})
    if (f && f->datatype != FLOAT32 && f->datatype != FLOAT64) return false;
  const uint32_t row_step = msg.row_step ? msg.row_step : msg.width * msg.point_step;
  if (!msg.data || msg.data_size < (size_t)msg.height * row_step) return false;
  const bool swap = msg.is_bigendian != detail::host_is_bigendian();
  out.resize((size_t)msg.height * msg.width);
  size_t k = 0;
  for (uint32_t r = 0; r < msg.height; r++)
    for (uint32_t c = 0; c < msg.width; c++, k++) {
      const uint8_t *p = msg.data + (size_t)r * row_step + (size_t)c * msg.point_step;
      out[k].x = detail::load_scalar(p + fx->offset, fx->datatype, swap);
      out[k].y = detail::load_scalar(p + fy->offset, fy->datatype, swap);
      out[k].z = detail::load_scalar(p + fz->offset, fz->datatype, swap);
      out[k].intensity = fi ? detail::load_scalar(p + fi->offset, fi->datatype, swap) : 0.0f;
    }
  return true;
}

// float4 points -> the PointCloud2 layout pcl::toROSMsg gives a pcl::PointXYZI cloud: point_step 32, x/y/z at 0/4/8, intensity at 16,
// unorganised (height 1), little endian.  `storage` receives the bytes the returned view points at.
inline PointCloud2View packPointCloud2(const mloam_point_t *pts, size_t n, std::vector<uint8_t> &storage) {
  PointCloud2View msg;
  msg.height = 1, msg.width = (uint32_t)n, msg.point_step = 32, msg.row_step = (uint32_t)(32 * n), msg.is_bigendian = false;
  const char *names[4] = {"x", "y", "z", "intensity"};
  const uint32_t offs[4] = {0, 4, 8, 16};
  for (int i = 0; i < 4; i++) {
    PointField f;
    f.name = names[i], f.offset = offs[i], f.datatype = FLOAT32, f.count = 1;
    msg.fields.push_back(f);
  }
  storage.assign(32 * n, 0);
  for (size_t i = 0; i < n; i++) {
    std::memcpy(&storage[32 * i], &pts[i].x, 12);
    std::memcpy(&storage[32 * i + 16], &pts[i].intensity, 4);
  }
  msg.data = storage.data(), msg.data_size = storage.size(), msg.has_intensity = true;
  return msg;
}

// ---- mloam_msgs (header reduced to stamp + frame_id; nav_msgs/Odometry and PoseWithCovarianceStamped to what the nodes fill)
struct Header {
  double stamp = 0.0;
  std::string frame_id;
};
struct PoseWithCovariance {
  double position[3] = {0, 0, 0};
  double orientation[4] = {0, 0, 0, 1};  // x y z w
  double covariance[36] = {0};           // row-major 6x6 (geometry_msgs/PoseWithCovariance)
};
struct Odometry {  // nav_msgs/Odometry: header, child_frame_id, pose (twist unused by M-LOAM)
  Header header;
  std::string child_frame_id;
  PoseWithCovariance pose;
};
struct Extrinsics {  // mloam_msgs/Extrinsics.msg
  Header header;
  uint8_t status = 0;
  std::vector<Odometry> odoms;
};
struct PoseWithCovarianceStamped {
  Header header;
  PoseWithCovariance pose;
};
struct Keyframes {  // mloam_msgs/Keyframes.msg
  Header header;
  uint8_t status = 0;
  std::vector<PoseWithCovarianceStamped> poses;
};
// parameter block [tx ty tz qx qy qz qw] (pose_local_parameterization.h:20) <-> message pose
inline void poseFromParam(const double *x7, PoseWithCovariance &p) {
  for (int k = 0; k < 3; k++) p.position[k] = x7[k];
  for (int k = 0; k < 4; k++) p.orientation[k] = x7[3 + k];
}
inline void poseToParam(const PoseWithCovariance &p, double *x7) {
  for (int k = 0; k < 3; k++) x7[k] = p.position[k];
  for (int k = 0; k < 4; k++) x7[3 + k] = p.orientation[k];
}

// ---- TUM trajectory line / file (save_statistics.hpp:104-119)
inline void writeTumLine(std::ostream &os, double stamp, const double *x7) {
  os.precision(15);
  os << stamp << " ";
  os.precision(8);
  os << x7[0] << " " << x7[1] << " " << x7[2] << " " << x7[3] << " " << x7[4] << " " << x7[5] << " " << x7[6] << std::endl;
}
inline std::string tumTrajectory(const std::vector<double> &stamps, const std::vector<double> &poses7) {
  std::ostringstream os;
  for (size_t i = 0; i < stamps.size(); i++) writeTumLine(os, stamps[i], &poses7[7 * i]);
  return os.str();
}

}  // namespace io
}  // namespace mloam
