// TEST STUB — minimal pcl::PointCloud / point types (see tests/stubs/Eigen/Dense).
#pragma once
#include <memory>
#include <vector>
namespace pcl {
struct PointXYZI {
  float x = 0, y = 0, z = 0, pad0 = 1, intensity = 0, pad1[3] = {0, 0, 0};  // 32 bytes like pcl::PointXYZI
};
struct PointXYZIWithCov {  // mloam_pcl/point_with_cov.hpp:45-53
  float x = 0, y = 0, z = 0, pad0 = 1, intensity = 0, cov_vec[6] = {0, 0, 0, 0, 0, 0}, cov_trace = 0;
};
template <typename P>
class PointCloud {
 public:
  typedef std::shared_ptr<PointCloud<P>> Ptr;
  typedef std::shared_ptr<const PointCloud<P>> ConstPtr;
  std::vector<P> points;
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); }
  void resize(size_t n) { points.resize(n); }
  void push_back(const P &p) { points.push_back(p); }
  P &operator[](size_t i) { return points[i]; }
  const P &operator[](size_t i) const { return points[i]; }
  typename std::vector<P>::const_iterator begin() const { return points.begin(); }
  typename std::vector<P>::const_iterator end() const { return points.end(); }
  PointCloud &operator+=(const PointCloud &o) {
    points.insert(points.end(), o.points.begin(), o.points.end());
    return *this;
  }
};
}  // namespace pcl
