// CPU self-test of mloam_io.hpp (no GPU, no library): PointCloud2 round trip incl. FLOAT64 / big-endian / padded rows,
// message PODs, TUM formatting against hand-written expectations.
#include <cassert>
#include <cmath>
#include <cstdio>
#include <iostream>

#include "mloam_io.hpp"

using namespace mloam::io;

static void put_be(uint8_t *dst, const void *src, int n) {
  const uint8_t *s = static_cast<const uint8_t *>(src);
  for (int i = 0; i < n; i++) dst[i] = s[n - 1 - i];
}

int main() {
  // round trip through the pcl::PointXYZI layout
  std::vector<mloam_point_t> pts;
  for (int i = 0; i < 1000; i++) pts.push_back(mloam_point_t{0.1f * i, -0.2f * i, 1.5f + i, (float)(i % 64) + 0.001f * i});
  std::vector<uint8_t> store;
  PointCloud2View msg = packPointCloud2(pts.data(), pts.size(), store);
  assert(msg.point_step == 32 && msg.width == 1000 && msg.fields.size() == 4 && msg.fields[3].offset == 16);
  std::vector<mloam_point_t> back;
  assert(unpackPointCloud2(msg, back) && msg.has_intensity && back.size() == pts.size());
  for (size_t i = 0; i < pts.size(); i++) assert(std::memcmp(&pts[i], &back[i], sizeof(mloam_point_t)) == 0);

  // organised cloud, padded rows, FLOAT64 z, big endian, fields in another order, no intensity
  PointCloud2View m2;
  m2.height = 2, m2.width = 3, m2.point_step = 24, m2.row_step = 80, m2.is_bigendian = true;
  PointField fz, fy, fx;
  fz.name = "z", fz.offset = 0, fz.datatype = FLOAT64;
  fy.name = "y", fy.offset = 8, fy.datatype = FLOAT32;
  fx.name = "x", fx.offset = 12, fx.datatype = FLOAT32;
  m2.fields = {fz, fy, fx};
  std::vector<uint8_t> raw(2 * 80, 0xab);
  for (int r = 0; r < 2; r++)
    for (int c = 0; c < 3; c++) {
      const double z = 100.0 * r + c + 0.25;
      const float y = -1.0f * c, x = 7.0f + r;
      uint8_t *p = &raw[r * 80 + c * 24];
      put_be(p, &z, 8), put_be(p + 8, &y, 4), put_be(p + 12, &x, 4);
    }
  m2.data = raw.data(), m2.data_size = raw.size();
  std::vector<mloam_point_t> o2;
  assert(unpackPointCloud2(m2, o2) && !m2.has_intensity && o2.size() == 6);
  assert(o2[4].x == 8.0f && o2[4].y == -1.0f && o2[4].z == 101.25f && o2[4].intensity == 0.0f);
  m2.data_size = 100;  // truncated buffer
  assert(!unpackPointCloud2(m2, o2));
  m2.data_size = raw.size(), m2.fields.pop_back();  // x missing
  assert(!unpackPointCloud2(m2, o2));

  // messages
  Extrinsics ex;
  ex.status = 1;
  ex.odoms.resize(2);
  const double x7[7] = {0.5355, 0.0393, -1.131, -0.0169, 0.0575, 0.0195, 0.998};
  poseFromParam(x7, ex.odoms[1].pose);
  double y7[7];
  poseToParam(ex.odoms[1].pose, y7);
  for (int k = 0; k < 7; k++) assert(x7[k] == y7[k]);
  Keyframes kf;
  kf.poses.resize(3);
  assert(kf.poses[2].pose.orientation[3] == 1.0 && kf.status == 0);

  // TUM line: stamp with 15 significant digits, the rest with 8 (ostream default notation)
  const double p7[7] = {1.23456789012, -0.5, 100.0, 0.0, 0.0, 0.70710678118654752, 0.70710678118654752};
  const std::string line = tumTrajectory({1317384506.40123456}, std::vector<double>(p7, p7 + 7));
  const std::string want = "1317384506.40123 1.2345679 -0.5 100 0 0 0.70710678 0.70710678\n";
  if (line != want) {
    std::cerr << "TUM line mismatch:\n" << line << want;
    return 1;
  }
  std::printf("io_selftest OK\n");
  return 0;
}
