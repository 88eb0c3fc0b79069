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

// TS.cpp:36-52: the state calibrate() is in when it reaches the parameter packing.  Both
// estimate_focal (TS.cpp:110-168) and estimate_extrinsic (TS.cpp:170-203) run on the GPU, batched
// over the frames, through tscm_mono_init(); there is no CPU path.
bool TripleSphereCamera::initial_guess(const std::vector<std::vector<cv::Point2d>>& pixels,
                                       std::vector<bool> has_chessboard,
                                       const std::vector<cv::Point3d>& worlds, const cv::Size img_size,
                                       const cv::Size chessboard_num) {
  pixels_ = pixels;
  has_chessboard_ = has_chessboard;
  const size_t img_num = pixels.size();
  Rt_.assign(img_num, cv::Mat());
  const int K = chessboard_num.width * chessboard_num.height;
  if (img_num == 0 || K <= 0 || (int)worlds.size() != K) {
    std::cout << "[tscm] initial_guess: " << img_num << " frames, " << worlds.size() << " board points for a "
              << chessboard_num.width << " x " << chessboard_num.height << " board" << std::endl;
    return false;
  }
  std::vector<double> w(3 * (size_t)K), px(2 * img_num * K, 0.0), rt(9 * img_num, 0.0);
  std::vector<uint8_t> has(img_num, 0), ok(img_num, 0);
  for (int j = 0; j < K; ++j) { w[3 * j] = worlds[j].x; w[3 * j + 1] = worlds[j].y; w[3 * j + 2] = worlds[j].z; }
  for (size_t k = 0; k < img_num; ++k) {
    // a frame takes part when it has corners (estimate_focal: pixels[k].size() != 0, TS.cpp:129)
    // and is flagged (estimate_extrinsic: has_chessboard_[k], TS.cpp:174); main.cpp:33-37 sets
    // the two together
    if (!has_chessboard_[k] || (int)pixels[k].size() != K) continue;
    has[k] = 1;
    std::memcpy(&px[2 * k * K], pixels[k].data(), sizeof(double) * 2 * K);      // cv::Point2d = {x, y}
  }
  tscm_mono_init_problem prob;
  prob.num_frames = (int32_t)img_num;
  prob.board_width = chessboard_num.width; prob.board_height = chessboard_num.height;
  prob.image_width = img_size.width; prob.image_height = img_size.height;
  prob.worlds = w.data(); prob.has_board = has.data(); prob.pixels = px.data();
  prob.has_init_guess = has_init_guess_ ? 1 : 0;
  tscm_mono_init_result res;
  std::memset(&res, 0, sizeof(res));
  const double guess[TSCM_INTRINSIC_SIZE] = {fx_, fy_, cx_, cy_, xi_, lamda_, alpha_, b_, c_};
  std::memcpy(res.intrinsics, guess, sizeof(guess));
  res.mono_rt = rt.data(); res.frame_ok = ok.data();
  const int rc = tscm_mono_init(&prob, device, &res);
  if (rc != TSCM_OK) {
    std::cout << "[tscm] initial guess failed: " << tscm_last_error() << std::endl;
    return false;
  }
  if (!has_init_guess_) {
    fx_ = res.intrinsics[0]; fy_ = res.intrinsics[1]; cx_ = res.intrinsics[2]; cy_ = res.intrinsics[3];
    xi_ = res.intrinsics[4]; lamda_ = res.intrinsics[5]; alpha_ = res.intrinsics[6];
    if (res.focal_rows_used == 0) std::cout << "焦距估计失败" << std::endl;      // the reference's message (TS.cpp:165)
    std::cout << "[initialize] focal:" << fx_ << ", cx:" << cx_ << ", cy:" << cy_ << ", xi:" << xi_
              << ", lamda:" << lamda_ << ", alpha:" << alpha_ << std::endl;
    if (fx_ == 0) return false;
  }
  for (size_t k = 0; k < img_num; ++k) {
    if (!has_chessboard_[k]) continue;
    if (!ok[k]) {
      // no pose (corners outside the model domain under the current guess give NaN rays): the
      // frame does not take part in the refinement — where OpenCV would raise a cv::Exception
      has_chessboard_[k] = false;
      continue;
    }
    cv::Mat pose(3, 3);
    for (int e = 0; e < 9; ++e) pose.at<double>(e / 3, e % 3) = rt[9 * k + e];
    Rt_[k] = pose;
  }
  return true;
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
