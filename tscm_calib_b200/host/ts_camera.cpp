// ts_camera.cpp — see ts_camera.h.  Host side: parameter packing and the cold-start
// initialisation (small dense linear algebra, as in the reference); the refinement is
// tscm_solve(), the remap tables are tscm_remap_tables().
#include "ts_camera.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>

namespace {
// One TS projection in plain doubles, skew included (TS.cpp:332-344).
inline cv::Point2d ts_project_point(double X, double Y, double Z, double fx, double fy, double cx,
                                    double cy, double xi, double lamda, double alpha, double b, double c) {
  const double r2 = X * X + Y * Y;
  const double d1 = std::sqrt(r2 + Z * Z);
  const double z1 = Z + xi * d1;
  const double d2 = std::sqrt(r2 + z1 * z1);
  const double z2 = z1 + lamda * d2;
  const double d3 = std::sqrt(r2 + z2 * z2);
  const double den = z2 + alpha / (1 - alpha) * d3;
  return cv::Point2d(fx * X / den + b * Y / den + cx, c * X / den + fy * Y / den + cy);
}
}  // namespace

TripleSphereCamera::TripleSphereCamera()
    : has_init_guess_(false), cx_(0), cy_(0), fx_(0), fy_(0), xi_(0), lamda_(0), alpha_(0), b_(0), c_(0),
      intrinsic_(TSCM_INTRINSIC_SIZE, 0.0) {
  tscm_options_init(&options_);
  options_.max_num_iterations = 100;   // TS.cpp:274
  options_.verbose = 1;                // summary.BriefReport() is printed (TS.cpp:280)
  std::memset(&summary_, 0, sizeof(summary_));
}

TripleSphereCamera::TripleSphereCamera(double fx, double fy, double cx, double cy, double xi,
                                       double lamda, double alpha)
    : TripleSphereCamera() {
  has_init_guess_ = true;
  fx_ = fx; fy_ = fy; cx_ = cx; cy_ = cy; xi_ = xi; lamda_ = lamda; alpha_ = alpha;
}

bool TripleSphereCamera::calibrate(const std::vector<std::vector<cv::Point2d>> pixels,
                                   std::vector<bool> has_chessboard,
                                   const std::vector<cv::Point3d>& worlds, const cv::Size img_size,
                                   const cv::Size chessboard_num) {
  const size_t img_num = pixels.size();
  if (has_chessboard.size() != img_num) return false;
  if (!initial_guess(pixels, has_chessboard, worlds, img_size, chessboard_num)) return false;   // TS.cpp:41-52
  rt_.assign(img_num, std::vector<double>(TSCM_POSE_SIZE, 0.0));
  const double packed[TSCM_INTRINSIC_SIZE] = {fx_, fy_, cx_, cy_, xi_, lamda_, alpha_, b_, c_};  // TS.cpp:53-61
  intrinsic_.assign(packed, packed + TSCM_INTRINSIC_SIZE);
  // [r1 r2 t] -> angle-axis + t, through single-precision r1, r2 like TS.cpp:66-72 (cv::Vec3f).
  for (size_t i = 0; i < img_num; ++i) {
    if (!has_chessboard_[i]) continue;
    const cv::Mat& M = Rt_[i];
    const float r1[3] = {(float)M.at<double>(0, 0), (float)M.at<double>(1, 0), (float)M.at<double>(2, 0)};
    const float r2[3] = {(float)M.at<double>(0, 1), (float)M.at<double>(1, 1), (float)M.at<double>(2, 1)};
    const float r3[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
    cv::Mat R(3, 3);
    for (int k = 0; k < 3; ++k) { R.at<double>(k, 0) = r1[k]; R.at<double>(k, 1) = r2[k]; R.at<double>(k, 2) = r3[k]; }
    cv::Mat r;
    cv::Rodrigues(R, r);
    for (int k = 0; k < 3; ++k) { rt_[i][k] = r.at<double>(k, 0); rt_[i][3 + k] = M.at<double>(k, 2); }
  }

  const bool status = refinement(pixels, worlds);
  if (status) has_init_guess_ = true;

  fx_ = intrinsic_[0]; fy_ = intrinsic_[1]; cx_ = intrinsic_[2]; cy_ = intrinsic_[3];
  xi_ = intrinsic_[4]; lamda_ = intrinsic_[5]; alpha_ = intrinsic_[6]; b_ = intrinsic_[7]; c_ = intrinsic_[8];
  for (size_t i = 0; i < img_num; ++i) {   // TS.cpp:89-105
    if (!has_chessboard_[i]) continue;
    cv::Mat r(3, 1), R;
    for (int k = 0; k < 3; ++k) r.at<double>(k, 0) = rt_[i][k];
    cv::Rodrigues(r, R);
    for (int k = 0; k < 3; ++k) {
      Rt_[i].at<double>(k, 0) = R.at<double>(k, 0);
      Rt_[i].at<double>(k, 1) = R.at<double>(k, 1);
      Rt_[i].at<double>(k, 2) = rt_[i][3 + k];
    }
  }
  return status;
}

// TS.cpp:36-52: the state calibrate() is in when it reaches the parameter packing.
bool TripleSphereCamera::initial_guess(const std::vector<std::vector<cv::Point2d>>& pixels,
                                       std::vector<bool> has_chessboard,
                                       const std::vector<cv::Point3d>& worlds, const cv::Size img_size,
                                       const cv::Size chessboard_num) {
  pixels_ = pixels;
  has_chessboard_ = has_chessboard;
  const size_t img_num = pixels.size();
  Rt_.resize(img_num);
  if (!has_init_guess_) {
    cx_ = img_size.width / 2 - 0.5;     // integer halves, as TS.cpp:43-44
    cy_ = img_size.height / 2 - 0.5;
    xi_ = 0.0;
    lamda_ = 0.0;
    alpha_ = 0.5;
    estimate_focal(pixels, worlds, img_size, chessboard_num);
    std::cout << "[initialize] focal:" << fx_ << ", cx:" << cx_ << ", cy:" << cy_ << ", xi:" << xi_
              << ", lamda:" << lamda_ << ", alpha:" << alpha_ << std::endl;
    if (fx_ == 0) return false;
  }
  estimate_extrinsic(pixels, worlds, chessboard_num);
  return true;
}

namespace {

// Focal length from ONE board row (TS.cpp:127-158).  With xi = lamda = 0, alpha = 0.5 the
// model is the unified sphere model with unit mirror parameter, under which a 3-D line images
// as the circle  x^2 + y^2 - 2 f (nx/nz) x - 2 f (ny/nz) y - f^2 = 0  (n = unit normal of the
// line's plane through the centre).  The row's corners, relative to the principal point, give
// the design matrix [x  y  1/2  -(x^2 + y^2)/2]; its null vector c fixes f = |c3 / (|..| nz)|.
// Rows whose plane is too oblique (nx^2 + ny^2 > 0.95) or whose fit is not a real circle are
// skipped; returns false for those.
bool focal_from_board_row(const cv::Point2d* row, int count, double cx, double cy, double* focal) {
  cv::Mat design(count, 4);
  for (int j = 0; j < count; ++j) {
    const double x = row[j].x - cx, y = row[j].y - cy;
    design.at<double>(j, 0) = x;
    design.at<double>(j, 1) = y;
    design.at<double>(j, 2) = 0.5;
    design.at<double>(j, 3) = -0.5 * (x * x + y * y);
  }
  cv::Mat nullvec;
  cv::SVD::solveZ(design, nullvec);
  const double c[4] = {nullvec.at<double>(0), nullvec.at<double>(1), nullvec.at<double>(2), nullvec.at<double>(3)};
  const double scale2 = c[0] * c[0] + c[1] * c[1] + c[2] * c[3];
  if (scale2 < 0) return false;
  const double inv_norm = std::sqrt(1 / scale2);
  const double nx = c[0] * inv_norm, ny = c[1] * inv_norm;
  const double oblique = nx * nx + ny * ny;
  if (oblique > 0.95) return false;
  *focal = std::fabs(c[2] * inv_norm / std::sqrt(1 - oblique));
  return true;
}

// Rotation that turns the viewing ray p (unit vector) onto +z: first about y by the azimuth,
// then about x by the elevation (TS.cpp:179-187).
cv::Mat rotation_facing(const cv::Point3d& p) {
  const double az = std::atan2(p.x, p.z), el = std::asin(p.y);
  cv::Mat about_y = (cv::Mat_<double>(3, 3) << std::cos(az), 0, -std::sin(az),
                                               0, 1, 0,
                                               std::sin(az), 0, std::cos(az));
  cv::Mat about_x = (cv::Mat_<double>(3, 3) << 1, 0, 0,
                                               0, std::cos(el), -std::sin(el),
                                               0, std::sin(el), std::cos(el));
  return about_x * about_y;
}

}  // namespace

// TS.cpp:110-168: mean of the per-row focal estimates over every frame with corners.
void TripleSphereCamera::estimate_focal(const std::vector<std::vector<cv::Point2d>>& pixels,
                                        const std::vector<cv::Point3d>& /*worlds*/, cv::Size /*img_size*/,
                                        const cv::Size chessboard_num) {
  double sum = 0;
  int rows_used = 0;
  for (const std::vector<cv::Point2d>& frame : pixels) {
    if (frame.empty()) continue;                        // no detection in this image
    for (int r = 0; r < chessboard_num.height; ++r) {
      double f;
      if (!focal_from_board_row(&frame[(size_t)r * chessboard_num.width], chessboard_num.width, cx_, cy_, &f)) continue;
      sum += f;
      ++rows_used;
    }
  }
  if (rows_used == 0) std::cout << "焦距估计失败" << std::endl;      // the reference's message (TS.cpp:165)
  fx_ = fy_ = rows_used > 0 ? sum / rows_used : 0.0;
}

// TS.cpp:170-203.  Per frame: the corners are lifted to the unit sphere with the current
// intrinsics, the sphere is turned so that a corner near the board centre looks down +z, the
// rays are projected to the normalised plane z = 1 and the board pose follows from
// solvePnPRansac with an identity camera matrix; turned back, it is kept in the [r1 r2 t] form.
void TripleSphereCamera::estimate_extrinsic(const std::vector<std::vector<cv::Point2d>>& pixels,
                                            const std::vector<cv::Point3d>& worlds,
                                            const cv::Size chessboard_num) {
  const cv::Mat identity = cv::Mat::eye(3, 3, cv::CV_64F);
  for (size_t k = 0; k < pixels.size(); ++k) {
    if (!has_chessboard_[k]) continue;
    const std::vector<cv::Point2d>& corners = pixels[k];
    const size_t centre = corners.size() / 2 - chessboard_num.width / 2 - 1;        // TS.cpp:177
    const cv::Mat facing = rotation_facing(get_unit_sphere_coordinate(corners[centre], identity));
    std::vector<cv::Point2d> plane(corners.size());
    for (size_t i = 0; i < corners.size(); ++i) {
      const cv::Point3d ray = get_unit_sphere_coordinate(corners[i], facing);
      plane[i] = cv::Point2d(ray.x / ray.z, ray.y / ray.z);
    }
    cv::Mat rvec, tvec, pose;
    const bool found = cv::solvePnPRansac(worlds, plane, identity, cv::Mat::zeros(4, 0, cv::CV_64F), rvec, tvec);
    if (!found || rvec.empty() || tvec.empty()) {
      // no pose (corners outside the model domain under the current guess give NaN rays): the
      // frame does not take part in the refinement — where OpenCV would raise a cv::Exception
      has_chessboard_[k] = false;
      continue;
    }
    cv::Rodrigues(rvec, pose);
    const cv::Mat back = facing.t();
    pose = back * pose;
    const cv::Mat t = back * tvec;
    for (int r = 0; r < 3; ++r) pose.at<double>(r, 2) = t.at<double>(r);
    Rt_[k] = pose;
  }
}

// Replaces TS.cpp:247-282: one residual block per corner of every frame with a board
// (intrinsic_ 9, rt_[i] 6), no loss, DENSE_SCHUR, 100 iterations -> tscm_solve().
bool TripleSphereCamera::refinement(const std::vector<std::vector<cv::Point2d>>& pixels,
                                    const std::vector<cv::Point3d>& worlds) {
  const int K = (int)worlds.size();
  std::vector<double> board_xy(2 * (size_t)K);
  for (int j = 0; j < K; ++j) { board_xy[2 * j] = worlds[j].x; board_xy[2 * j + 1] = worlds[j].y; }
  // frames with a board become the problem's frames (has_chessboard_, TS.cpp:253)
  std::vector<int> frame_of;
  for (size_t i = 0; i < pixels.size(); ++i)
    if (has_chessboard_[i] && (int)pixels[i].size() == K) frame_of.push_back((int)i);
  const int F = (int)frame_of.size();
  if (F == 0 || K == 0) return false;
  std::vector<int32_t> view_camera(F, 0), view_frame(F);
  std::vector<double> obs(2 * (size_t)F * K), board_rt(6 * (size_t)F);
  for (int f = 0; f < F; ++f) {
    view_frame[f] = f;
    const std::vector<cv::Point2d>& px = pixels[frame_of[f]];
    std::memcpy(&obs[2 * (size_t)f * K], px.data(), sizeof(double) * 2 * K);   // cv::Point2d = {x, y}
    std::memcpy(&board_rt[6 * (size_t)f], rt_[frame_of[f]].data(), sizeof(double) * 6);
  }
  double cam_rt[6] = {0, 0, 0, 0, 0, 0};   // mono = rig with an identity, constant camera
  tscm_problem p;
  p.num_cameras = 1; p.num_frames = F; p.corners_per_board = K; p.num_views = F;
  p.board_xy = board_xy.data(); p.view_camera = view_camera.data(); p.view_frame = view_frame.data();
  p.obs_xy = obs.data(); p.fixed_camera = 0;
  std::memset(&summary_, 0, sizeof(summary_));
  const int rc = tscm_solve(&p, &options_, intrinsic_.data(), cam_rt, board_rt.data(), &summary_, device);
  if (rc != TSCM_OK) {
    std::cout << "[tscm] refinement failed: " << tscm_last_error() << std::endl;
    return false;
  }
  for (int f = 0; f < F; ++f)
    std::memcpy(rt_[frame_of[f]].data(), &board_rt[6 * (size_t)f], sizeof(double) * 6);
  return summary_.termination_type == TSCM_CONVERGENCE;
}

void TripleSphereCamera::Reproject(const std::vector<cv::Point3d>& worlds, cv::Mat Rt,
                                   std::vector<cv::Point2d>& pixels) {
  pixels.resize(worlds.size());
  for (size_t i = 0; i < worlds.size(); ++i) {
    const double x = worlds[i].x, y = worlds[i].y;
    const double X = Rt.at<double>(0, 0) * x + Rt.at<double>(0, 1) * y + Rt.at<double>(0, 2);
    const double Y = Rt.at<double>(1, 0) * x + Rt.at<double>(1, 1) * y + Rt.at<double>(1, 2);
    const double Z = Rt.at<double>(2, 0) * x + Rt.at<double>(2, 1) * y + Rt.at<double>(2, 2);
    pixels[i] = ts_project_point(X, Y, Z, fx_, fy_, cx_, cy_, xi_, lamda_, alpha_, b_, c_);
  }
}

cv::Point2d TripleSphereCamera::project(cv::Mat P) {
  return ts_project_point(P.at<double>(0, 0), P.at<double>(1, 0), P.at<double>(2, 0), fx_, fy_, cx_, cy_,
                          xi_, lamda_, alpha_, b_, c_);
}

cv::Point3d TripleSphereCamera::get_unit_sphere_coordinate(cv::Point2d pixel, cv::Mat transform) {
  const double px = pixel.x - cx_, py = pixel.y - cy_;
  const double det = fx_ * fy_ - b_ * c_;
  const double mx = (fy_ * px - b_ * py) / det, my = (-c_ * px + fx_ * py) / det;
  const double k = alpha_ / (1 - alpha_);
  const double r2 = mx * mx + my * my;
  const double gamma = (k + std::sqrt(1 + (1 - k * k) * r2)) / (r2 + 1);
  const double gk = gamma - k;
  const double eta = lamda_ * gk + std::sqrt((gk * gk - 1) * lamda_ * lamda_ + 1);
  const double mz = eta * gk;
  const double ml = mz - lamda_;
  const double mu = xi_ * ml + std::sqrt(xi_ * xi_ * (ml * ml - 1) + 1);
  const double v[3] = {mu * eta * gamma * mx, mu * eta * gamma * my, mu * ml - xi_};
  double o[3];
  for (int r = 0; r < 3; ++r)
    o[r] = transform.at<double>(r, 0) * v[0] + transform.at<double>(r, 1) * v[1] + transform.at<double>(r, 2) * v[2];
  return cv::Point3d(o[0], o[1], o[2]);
}

double TripleSphereCamera::ReprojectError(const std::vector<cv::Point2d>& pixels,
                                          const std::vector<cv::Point3d>& worlds, cv::Mat R, cv::Mat t) {
  double error = 0;
  for (size_t i = 0; i < worlds.size(); ++i) {
    double P[3];
    for (int r = 0; r < 3; ++r)
      P[r] = R.at<double>(r, 0) * worlds[i].x + R.at<double>(r, 1) * worlds[i].y +
             R.at<double>(r, 2) * worlds[i].z + t.at<double>(r, 0);
    const cv::Point2d q = ts_project_point(P[0], P[1], P[2], fx_, fy_, cx_, cy_, xi_, lamda_, alpha_, b_, c_);
    error += std::sqrt((pixels[i].x - q.x) * (pixels[i].x - q.x) + (pixels[i].y - q.y) * (pixels[i].y - q.y));
  }
  return error;
}

namespace {
// One block covering the whole table, through the CUDA kernel (no CPU path).
bool fill_tables(const double intr[9], const cv::Mat& M, double rfx, double rfy, double rcx, double rcy,
                 cv::Size size, int device, cv::Mat& mapx, cv::Mat& mapy) {
  mapx = cv::Mat(size, cv::CV_32FC1);
  mapy = cv::Mat(size, cv::CV_32FC1);
  tscm_remap_job job;
  std::memset(&job, 0, sizeof(job));
  std::memcpy(job.intrinsics, intr, sizeof(job.intrinsics));
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) job.matrix[3 * r + c] = M.at<double>(r, c);
  job.ray_fx = rfx; job.ray_fy = rfy; job.ray_cx = rcx; job.ray_cy = rcy;
  job.width = size.width; job.height = size.height;
  const int rc = tscm_remap_tables(&job, 1, size.width, size.height, mapx.ptr<float>(), mapy.ptr<float>(), device, nullptr);
  if (rc != TSCM_OK) std::cout << "[tscm] remap tables failed: " << tscm_last_error() << std::endl;
  return rc == TSCM_OK;
}
}  // namespace

void TripleSphereCamera::undistort(double fx, double fy, double cx, double cy, cv::Size img_size, cv::Mat& mapx,
                                   cv::Mat& mapy) {
  const double intr[9] = {fx_, fy_, cx_, cy_, xi_, lamda_, alpha_, b_, c_};
  fill_tables(intr, cv::Mat::eye(3, 3, cv::CV_64F), fx, fy, cx, cy, img_size, device, mapx, mapy);
}

bool TripleSphereCamera::undistort_chessboard_maps(int index, cv::Size chessboard, double chessboard_size,
                                                   cv::Mat& mapx, cv::Mat& mapy) {
  if (index < 0 || index >= (int)has_chessboard_.size() || has_chessboard_[index] == false) return false;
  const cv::Size img_size((int)((chessboard.width + 1) * chessboard_size), (int)((chessboard.height + 1) * chessboard_size));
  const double intr[9] = {fx_, fy_, cx_, cy_, xi_, lamda_, alpha_, b_, c_};
  // P = Rt * (j - size, i - size, 1)   (TS.cpp:321-322)
  return fill_tables(intr, Rt_[index], 1.0, 1.0, chessboard_size, chessboard_size, img_size, device, mapx, mapy);
}

#ifdef TSCM_USE_OPENCV
cv::Mat TripleSphereCamera::undistort_chessboard(cv::Mat src, int index, cv::Size chessboard, double chessboard_size) {
  cv::Mat dst, mapx, mapy;
  if (!undistort_chessboard_maps(index, chessboard, chessboard_size, mapx, mapy)) return dst;
  cv::remap(src, dst, mapx, mapy, cv::INTER_LINEAR);
  return dst;
}
#endif
