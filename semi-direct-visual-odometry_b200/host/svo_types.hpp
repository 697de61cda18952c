// svo_types.hpp -- minimal value types standing in for Eigen::Vector2d/3d, Sophus::SE3d and cv::Mat / cv::Size.
// The image used to build this project has no Eigen / Sophus / OpenCV C++ headers, so the host
// classes are written against these stand-ins; each offers the members the reference's hot path uses (x(), y(), z(),
// operator*, inverse(), params(), rows/cols/ptr()).  Where the real libraries exist, adapters are one-liners because
// the layouts agree (Vec = contiguous doubles, SE3::params() = qx qy qz qw tx ty tz as Sophus, Mat8 = continuous
// CV_8UC1).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

namespace svo {

struct Vec2 {
    double v[2]{0, 0};
    Vec2() = default;
    Vec2(double a, double b) : v{a, b} {}
    double& x() { return v[0]; }
    double& y() { return v[1]; }
    double x() const { return v[0]; }
    double y() const { return v[1]; }
    double& operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
};

struct Vec3 {
    double v[3]{0, 0, 0};
    Vec3() = default;
    Vec3(double a, double b, double c) : v{a, b, c} {}
    double& x() { return v[0]; }
    double& y() { return v[1]; }
    double& z() { return v[2]; }
    double x() const { return v[0]; }
    double y() const { return v[1]; }
    double z() const { return v[2]; }
    double& operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
    double norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
    Vec3 normalized() const
    {
        const double n = norm();
        return Vec3(v[0] / n, v[1] / n, v[2] / n);
    }
    Vec3 operator+(const Vec3& o) const { return Vec3(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
    Vec3 operator-(const Vec3& o) const { return Vec3(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    Vec3 operator*(double s) const { return Vec3(v[0] * s, v[1] * s, v[2] * s); }
    Vec3 operator-() const { return Vec3(-v[0], -v[1], -v[2]); }
};

inline Vec3 cross(const Vec3& a, const Vec3& b)
{
    return Vec3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}

// Rigid transform, unit quaternion (x y z w) + translation; params() order as Sophus::SE3d::params().
struct SE3 {
    double q[4]{0, 0, 0, 1};
    double t[3]{0, 0, 0};
    SE3() = default;
    static SE3 fromParams(const double p[7])
    {
        SE3 T;
        std::memcpy(T.q, p, sizeof(T.q));
        std::memcpy(T.t, p + 4, sizeof(T.t));
        return T;
    }
    void params(double p[7]) const
    {
        std::memcpy(p, q, sizeof(q));
        std::memcpy(p + 4, t, sizeof(t));
    }
    Vec3 translation() const { return Vec3(t[0], t[1], t[2]); }
    Vec3 rotate(const Vec3& p) const
    {
        const Vec3 qv(q[0], q[1], q[2]);
        const Vec3 uv = cross(qv, p) * 2.0;
        return p + uv * q[3] + cross(qv, uv);
    }
    Vec3 operator*(const Vec3& p) const { return rotate(p) + translation(); }
    SE3 inverse() const
    {
        SE3 o;
        o.q[0] = -q[0];
        o.q[1] = -q[1];
        o.q[2] = -q[2];
        o.q[3] = q[3];
        const Vec3 r = o.rotate(translation());
        o.t[0]       = -r[0];
        o.t[1]       = -r[1];
        o.t[2]       = -r[2];
        return o;
    }
    SE3 operator*(const SE3& b) const
    {
        SE3 o;
        const double ax = q[0], ay = q[1], az = q[2], aw = q[3];
        const double bx = b.q[0], by = b.q[1], bz = b.q[2], bw = b.q[3];
        double r[4] = {aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                       aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz};
        const double n = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
        for (int i = 0; i < 4; i++) o.q[i] = r[i] / n;
        const Vec3 rt = rotate(b.translation());
        for (int i = 0; i < 3; i++) o.t[i] = t[i] + rt[i];
        return o;
    }
};

struct Size {
    int width = 0, height = 0;
    Size() = default;
    Size(int w, int h) : width(w), height(h) {}
};

// Continuous 8-bit single-channel image with shared ownership (the subset of cv::Mat the hot path touches).
struct Mat8 {
    int rows = 0, cols = 0;
    std::shared_ptr<std::vector<uint8_t>> buf;
    Mat8() = default;
    Mat8(int r, int c) : rows(r), cols(c), buf(std::make_shared<std::vector<uint8_t>>((size_t)r * c)) {}
    Mat8(int r, int c, const uint8_t* src) : Mat8(r, c) { std::memcpy(buf->data(), src, (size_t)r * c); }
    bool empty() const { return !buf || rows == 0 || cols == 0; }
    uint8_t* ptr(int r = 0) { return buf->data() + (size_t)r * cols; }
    const uint8_t* ptr(int r = 0) const { return buf->data() + (size_t)r * cols; }
    uint8_t at(int r, int c) const { return (*buf)[(size_t)r * cols + c]; }
    Size size() const { return Size(cols, rows); }
};

}  // namespace svo

// ---------------------------------------------------------------------------------------------------------------------
// Adapters to the reference's real types, compiled in wherever their headers exist (the build image of this project has
// none of them; tests/test_integration_compile.py compiles this block and INTEGRATION.md's bindings against minimal
// stand-ins that carry the real signatures).  Frame::m_absPose is a Sophus::SE3d (include/frame.hpp:198), the pyramids
// are std::vector<cv::Mat> (include/image_pyramid.hpp:147-148), FeatureAlignment::align takes an Eigen::Vector2d&
// (include/feature_alignment.hpp:25).
// ---------------------------------------------------------------------------------------------------------------------
#if defined(__has_include)
#if __has_include(<Eigen/Core>)
#include <Eigen/Core>
#define SVO_HAVE_EIGEN 1
namespace svo {
inline Vec2 fromEigen(const Eigen::Vector2d& v) { return Vec2(v.x(), v.y()); }
inline Vec3 fromEigen(const Eigen::Vector3d& v) { return Vec3(v.x(), v.y(), v.z()); }
inline Eigen::Vector2d toEigen(const Vec2& v) { return Eigen::Vector2d(v.x(), v.y()); }
inline Eigen::Vector3d toEigen(const Vec3& v) { return Eigen::Vector3d(v.x(), v.y(), v.z()); }
}  // namespace svo
#endif
#if __has_include(<sophus/se3.hpp>)
#include <sophus/se3.hpp>
#define SVO_HAVE_SOPHUS 1
namespace svo {
// Sophus::SE3d::params() is qx qy qz qw tx ty tz: the order of every pose in include/svo_b200.h
inline SE3 fromSophus(const Sophus::SE3d& T)
{
    const auto p = T.params();
    return SE3::fromParams(p.data());
}
inline Sophus::SE3d toSophus(const SE3& T)
{
    return Sophus::SE3d(Eigen::Quaterniond(T.q[3], T.q[0], T.q[1], T.q[2]), Eigen::Vector3d(T.t[0], T.t[1], T.t[2]));
}
}  // namespace svo
#endif
#if __has_include(<opencv2/core.hpp>)
#include <opencv2/core.hpp>
#define SVO_HAVE_OPENCV 1
namespace svo {
inline Mat8 fromCv(const cv::Mat& m)  // CV_8UC1, any row step
{
    Mat8 o(m.rows, m.cols);
    for (int r = 0; r < m.rows; r++) std::memcpy(o.ptr(r), m.ptr<uint8_t>(r), (size_t)m.cols);
    return o;
}
inline cv::Mat toCv(const Mat8& m) { return cv::Mat(m.rows, m.cols, CV_8UC1, (void*)m.ptr()).clone(); }
}  // namespace svo
#endif
#endif

