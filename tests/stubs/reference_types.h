// TEST STUB — the declarations of the reference the adapter expects to find (estimator/src/estimator/parameters.h:161-207,
// estimator/src/estimator/pose.h:38-67, mloam_common .../types/type.h:13-33), reduced to members; written against the stub Eigen / PCL.
#pragma once
#include <map>
#include <string>
#include <vector>

#include <Eigen/Dense>
#include <pcl/point_cloud.h>

namespace common {
typedef pcl::PointXYZI PointI;
typedef pcl::PointCloud<PointI> PointICloud;
typedef pcl::PointXYZIWithCov PointIWithCov;
typedef pcl::PointCloud<PointIWithCov> PointICovCloud;
}  // namespace common
using common::PointI;
using common::PointICloud;
using common::PointIWithCov;
using common::PointICovCloud;

typedef std::map<std::string, common::PointICloud> cloudFeature;

class PointPlaneFeature {
 public:
  PointPlaneFeature() : idx_(0), laser_idx_(0), type_('n') {}
  size_t idx_;
  size_t laser_idx_;
  Eigen::Vector3d point_;
  Eigen::VectorXd coeffs_;
  Eigen::MatrixXd jaco_;
  char type_;
};

class ScanInfo {
 public:
  ScanInfo(const int &n_scan, const bool &segment_flag) {
    segment_flag_ = segment_flag;
    scan_start_ind_.resize(n_scan);
    scan_end_ind_.resize(n_scan);
  }
  std::vector<int> scan_start_ind_, scan_end_ind_;
  bool segment_flag_;
  std::vector<bool> ground_flag_;
};

class Pose {
 public:
  Pose() {}
  Pose(const Eigen::Quaterniond &q, const Eigen::Vector3d &t, const double &td = 0) : td_(td), q_(q), t_(t) { q_.normalize(); }
  double td_ = 0;
  Eigen::Quaterniond q_;
  Eigen::Vector3d t_;
  Eigen::Matrix<double, 6, 6> cov_;
};
