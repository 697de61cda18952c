// Minimal stand-in for <sophus/se3.hpp>: params() in Sophus' order qx qy qz qw tx ty tz, the (quaternion, translation) ctor.
#pragma once
#include <Eigen/Core>
namespace Sophus {
struct SE3d {
    Eigen::Matrix<double, 7, 1> p;
    SE3d() { p[3] = 1.0; }
    SE3d(const Eigen::Quaterniond& q, const Eigen::Vector3d& t)
    {
        p[0] = q.x(), p[1] = q.y(), p[2] = q.z(), p[3] = q.w(), p[4] = t[0], p[5] = t[1], p[6] = t[2];
    }
    Eigen::Matrix<double, 7, 1> params() const { return p; }
};
}  // namespace Sophus
