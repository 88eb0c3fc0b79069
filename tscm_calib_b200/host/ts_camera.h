// ts_camera.h — drop-in `TripleSphereCamera` whose non-linear refinement runs on the B200.
//
// Mirrors the public interface of the reference class (/root/reference/TS.h:10-69): same
// method names, argument meaning and return conventions.  What changed is the body of
// refinement() (TS.cpp:247-282): instead of building a ceres::Problem and calling
// ceres::Solve it hands the same arrays to tscm_solve() (include/tscm.h).
//
// The cold-start initialisation — estimate_focal (TS.cpp:110-168, circle fit
// through cv::SVD::solveZ) and estimate_extrinsic (TS.cpp:170-203, cv::solvePnPRansac on
// unit-sphere-normalised corners) — runs on the GPU as well, batched over the frames, behind
// tscm_mono_init() (csrc/tscm_monoinit.cu).
// The remap tables of undistort / undistort_chessboard (TS.cpp:284-330, §8f #4) are filled
// by the CUDA kernel behind tscm_remap_tables().
#pragma once

#include <vector>

#include "../../include/tscm.h"
#include "cv_compat.h"

class TripleSphereCamera {
 public:
  TripleSphereCamera();
  TripleSphereCamera(double fx, double fy, double cx, double cy, double xi, double lamda, double alpha);

  // TS.cpp:30-108.  Returns true iff the solve terminated with CONVERGENCE (TS.cpp:281).
  bool calibrate(const std::vector<std::vector<cv::Point2d>> pixels, std::vector<bool> has_chessboard,
                 const std::vector<cv::Point3d>& worlds, const cv::Size img_size,
                 const cv::Size chessboard_num);
  // TS.cpp:227-245 — Rt is the 3x3 [r1 r2 t] matrix applied to (x, y, 1).
  void Reproject(const std::vector<cv::Point3d>& worlds, cv::Mat Rt, std::vector<cv::Point2d>& pixels);
  // TS.cpp:332-344 — skew terms b, c included.
  cv::Point2d project(cv::Mat P);
  // TS.h:39-57 — closed-form back-projection to the unit sphere.
  cv::Point3d get_unit_sphere_coordinate(cv::Point2d pixel, cv::Mat transform = cv::Mat::eye(3, 3, cv::CV_64F));
  // TS.h:58-69 — summed Euclidean reprojection error of one board under (R, t).
  double ReprojectError(const std::vector<cv::Point2d>& pixels, const std::vector<cv::Point3d>& worlds,
                        cv::Mat R, cv::Mat t);
  // TS.cpp:284-306 — CV_32FC1 lookup tables of a pinhole (fx, fy, cx, cy) view, on the GPU.
  void undistort(double fx, double fy, double cx, double cy, cv::Size img_size, cv::Mat& mapx, cv::Mat& mapy);
  // TS.cpp:308-326 — the tables of the fronto-parallel board image of frame `index` (false
  // if the frame has no board, where the reference returns an empty image).
  bool undistort_chessboard_maps(int index, cv::Size chessboard, double chessboard_size, cv::Mat& mapx,
                                 cv::Mat& mapy);
#ifdef TSCM_USE_OPENCV
  // TS.cpp:308-330 — tables as above, then cv::remap (OpenCV's own).
  cv::Mat undistort_chessboard(cv::Mat src, int index, cv::Size chessboard, double chessboard_size);
#endif
  // The cold-start part of calibrate() on its own (TS.cpp:41-52): default intrinsics,
  // estimate_focal, estimate_extrinsic.  False when the focal estimate fails (TS.cpp:50).
  bool initial_guess(const std::vector<std::vector<cv::Point2d>>& pixels, std::vector<bool> has_chessboard,
                     const std::vector<cv::Point3d>& worlds, const cv::Size img_size,
                     const cv::Size chessboard_num);

  double cx() { return cx_; }
  double cy() { return cy_; }
  double fx() { return fx_; }
  double fy() { return fy_; }
  double xi() { return xi_; }
  double lamda() { return lamda_; }
  double alpha() { return alpha_; }
  double b() { return b_; }
  double c() { return c_; }
  std::vector<cv::Mat> Rt() { return Rt_; }
  cv::Mat Rt(int id) { return Rt_[id]; }
  void setRt(std::vector<cv::Mat> Rts) { Rt_ = Rts; }
  void setPixels(std::vector<std::vector<cv::Point2d>> pixels) { pixels_ = pixels; }
  std::vector<std::vector<cv::Point2d>> pixels() { return pixels_; }
  const std::vector<std::vector<cv::Point2d>>& pixels_ref() const { return pixels_; }   // (no copy; not in the reference)
  std::vector<bool> has_chessboard() { return has_chessboard_; }
  bool has_chessboard(int id) { return has_chessboard_[id]; }
  void setHasChessboard(std::vector<bool> has_chessboard) { has_chessboard_ = has_chessboard; }

  // Solver knobs the reference hard-codes (TS.cpp:271-274); exposed for tests.
  tscm_options& options() { return options_; }
  const tscm_summary& last_summary() const { return summary_; }
  int device = -1;

 private:
  bool refinement(const std::vector<std::vector<cv::Point2d>>& pixels, const std::vector<cv::Point3d>& worlds);

  bool has_init_guess_;
  double cx_, cy_, fx_, fy_, xi_, lamda_, alpha_, b_, c_;
  std::vector<cv::Mat> Rt_;
  std::vector<std::vector<double>> rt_;
  std::vector<std::vector<cv::Point2d>> pixels_;
  std::vector<double> intrinsic_;
  std::vector<bool> has_chessboard_;
  tscm_options options_;
  tscm_summary summary_;
};
