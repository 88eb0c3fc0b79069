// cv_compat.h — the handful of OpenCV types the reference's TS / MultiCalib interface is
// written in (cv::Point2d, cv::Point3d, cv::Size, a double cv::Mat, cv::Rodrigues).
//
// OpenCV's C++ headers are not installed in this image, so the drop-in adapters
// (ts_camera.h, multi_calib_b200.h) compile against this shim; when
// <opencv2/opencv.hpp> IS available define TSCM_USE_OPENCV and the real types are
// used instead (the adapters only touch the API subset implemented here).
#pragma once

#ifdef TSCM_USE_OPENCV
#include <opencv2/opencv.hpp>
#else

#include <cassert>
#include <cmath>
#include <algorithm>
#include <cstddef>
#include <memory>
#include <vector>

namespace cv {

constexpr int CV_64F = 6;
constexpr int CV_32FC1 = 5;

template <typename T>
struct Point_ {
  T x{}, y{};
  Point_() = default;
  Point_(T x_, T y_) : x(x_), y(y_) {}
};
template <typename T>
struct Point3_ {
  T x{}, y{}, z{};
  Point3_() = default;
  Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point_<double> Point2d;
typedef Point3_<double> Point3d;

struct Size {
  int width = 0, height = 0;
  Size() = default;
  Size(int w, int h) : width(w), height(h) {}
};

// Dense row-major single-channel matrix with reference semantics on copy (like cv::Mat).
// CV_64F everywhere in the calibration path; CV_32FC1 for the remap tables (TS.cpp:286-287).
class Mat {
 public:
  int rows = 0, cols = 0;
  Mat() = default;
  Mat(int r, int c, int type = CV_64F)
      : rows(r), cols(c), type_(type),
        d_(std::make_shared<std::vector<unsigned char>>((size_t)r * c * elem_size(type), (unsigned char)0)) {}
  Mat(Size s, int type) : Mat(s.height, s.width, type) {}
  int type() const { return type_; }
  bool empty() const { return !d_ || d_->empty(); }
  template <typename T> T& at(int i, int j = 0) { return reinterpret_cast<T*>(d_->data())[(size_t)i * cols + j]; }
  template <typename T> const T& at(int i, int j = 0) const { return reinterpret_cast<const T*>(d_->data())[(size_t)i * cols + j]; }
  template <typename T> T* ptr(int i = 0) { return reinterpret_cast<T*>(d_->data()) + (size_t)i * cols; }
  Mat clone() const { Mat m(rows, cols, type_); if (d_) *m.d_ = *d_; return m; }
  static Mat eye(int r, int c, int = CV_64F) { Mat m(r, c); for (int i = 0; i < r && i < c; ++i) m.at<double>(i, i) = 1.0; return m; }
  static Mat zeros(int r, int c, int = CV_64F) { return Mat(r, c); }
  Mat t() const { Mat m(cols, rows); for (int i = 0; i < rows; ++i) for (int j = 0; j < cols; ++j) m.at<double>(j, i) = at<double>(i, j); return m; }
 private:
  static size_t elem_size(int type) { return type == CV_64F ? 8 : type == CV_32FC1 ? 4 : 1; }
  int type_ = CV_64F;
  std::shared_ptr<std::vector<unsigned char>> d_;
};

inline Mat operator*(const Mat& a, const Mat& b) {
  assert(a.cols == b.rows);
  Mat m(a.rows, b.cols);
  for (int i = 0; i < a.rows; ++i)
    for (int j = 0; j < b.cols; ++j) {
      double s = 0.0;
      for (int k = 0; k < a.cols; ++k) s += a.at<double>(i, k) * b.at<double>(k, j);
      m.at<double>(i, j) = s;
    }
  return m;
}
inline Mat operator+(const Mat& a, const Mat& b) {
  Mat m(a.rows, a.cols);
  for (int i = 0; i < a.rows; ++i) for (int j = 0; j < a.cols; ++j) m.at<double>(i, j) = a.at<double>(i, j) + b.at<double>(i, j);
  return m;
}
inline Mat operator-(const Mat& a, const Mat& b) {
  Mat m(a.rows, a.cols);
  for (int i = 0; i < a.rows; ++i) for (int j = 0; j < a.cols; ++j) m.at<double>(i, j) = a.at<double>(i, j) - b.at<double>(i, j);
  return m;
}

// (cv::Mat_<double>(r, c) << a, b, ...) initialiser.
template <typename T>
class Mat_ : public Mat {
 public:
  Mat_() = default;
  Mat_(int r, int c) : Mat(r, c) {}
  Mat_(const Mat& m) : Mat(m) {}
  T& operator()(int i, int j = 0) { return this->template at<T>(i, j); }
  const T& operator()(int i, int j = 0) const { return this->template at<T>(i, j); }
  struct Init {
    Mat_ m; int k;
    Init& operator,(T v) { m.template at<T>(k / m.cols, k % m.cols) = v; ++k; return *this; }
    operator Mat() const { return m; }
    operator Mat_() const { return m; }
  };
  Init operator<<(T v) { Init in{*this, 0}; in.m.template at<T>(0, 0) = v; in.k = 1; return in; }
};

// cv::Rodrigues: 3x1 (or 1x3) rotation vector <-> 3x3 rotation matrix.
inline void Rodrigues(const Mat& src, Mat& dst) {
  if (src.rows * src.cols == 3) {
    const double rx = src.at<double>(0), ry = src.rows == 3 ? src.at<double>(1, 0) : src.at<double>(0, 1),
                 rz = src.rows == 3 ? src.at<double>(2, 0) : src.at<double>(0, 2);
    const double theta = std::sqrt(rx * rx + ry * ry + rz * rz);
    dst = Mat::eye(3, 3);
    if (theta < 2.220446049250313e-16) return;
    const double c = std::cos(theta), s = std::sin(theta), c1 = 1.0 - c, it = 1.0 / theta;
    const double x = rx * it, y = ry * it, z = rz * it;
    double R[9] = {c + c1 * x * x,     c1 * x * y - s * z, c1 * x * z + s * y,
                   c1 * x * y + s * z, c + c1 * y * y,     c1 * y * z - s * x,
                   c1 * x * z - s * y, c1 * y * z + s * x, c + c1 * z * z};
    for (int i = 0; i < 9; ++i) dst.at<double>(i / 3, i % 3) = R[i];
  } else {
    // matrix -> vector.  Like OpenCV (which takes the SVD and uses U V^T) the input is first
    // replaced by the nearest rotation: the reference feeds matrices that are NOT orthonormal
    // (r3 = r1 x r2 from float-truncated columns, chained pose-graph products).  The orthogonal
    // polar factor — the same matrix as U V^T — comes from Newton's iteration X <- (X + X^-T) / 2.
    // Then through the quaternion (stable for every angle).
    assert(src.rows == 3 && src.cols == 3 && "cv::Rodrigues: 3x1, 1x3 or 3x3 input");
    double M[9];
    for (int i = 0; i < 9; ++i) M[i] = src.at<double>(i / 3, i % 3);
    for (int it = 0; it < 12; ++it) {
      const double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
      const double det = M[0] * c00 + M[1] * c01 + M[2] * c02;
      if (!(std::fabs(det) > 1e-300)) break;
      // inverse transpose = cofactor matrix / det
      const double T[9] = {c00, c01, c02,
                           M[2] * M[7] - M[1] * M[8], M[0] * M[8] - M[2] * M[6], M[1] * M[6] - M[0] * M[7],
                           M[1] * M[5] - M[2] * M[4], M[2] * M[3] - M[0] * M[5], M[0] * M[4] - M[1] * M[3]};
      double change = 0.0;
      for (int i = 0; i < 9; ++i) {
        const double v = 0.5 * (M[i] + T[i] / det);
        change = std::max(change, std::fabs(v - M[i]));
        M[i] = v;
      }
      if (change < 1e-16) break;
    }
    double q[4];
    const double tr = M[0] + M[4] + M[8];
    if (tr > 0) {
      const double s = std::sqrt(tr + 1.0) * 2;
      q[0] = 0.25 * s; q[1] = (M[7] - M[5]) / s; q[2] = (M[2] - M[6]) / s; q[3] = (M[3] - M[1]) / s;
    } else {
      int i = 0;
      if (M[4] > M[0]) i = 1;
      if (M[8] > M[4 * i]) i = 2;
      const int j = (i + 1) % 3, k = (i + 2) % 3;
      const double s = std::sqrt(1.0 + M[4 * i] - M[4 * j] - M[4 * k]) * 2;
      q[0] = (M[3 * k + j] - M[3 * j + k]) / s;
      q[1 + i] = 0.25 * s;
      q[1 + j] = (M[3 * j + i] + M[3 * i + j]) / s;
      q[1 + k] = (M[3 * k + i] + M[3 * i + k]) / s;
    }
    if (q[0] < 0) for (double& v : q) v = -v;
    const double vn = std::sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    dst = Mat(3, 1);
    if (vn > 0) {
      const double ang = 2.0 * std::atan2(vn, q[0]);
      for (int i = 0; i < 3; ++i) dst.at<double>(i, 0) = q[1 + i] / vn * ang;
    }
  }
}

}  // namespace cv
#endif  // TSCM_USE_OPENCV
