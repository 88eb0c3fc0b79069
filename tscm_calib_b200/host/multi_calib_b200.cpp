// multi_calib_b200.cpp — see multi_calib_b200.h.
#include "multi_calib_b200.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>

namespace {

cv::Mat vec3(double a, double b, double c) {
  cv::Mat m(3, 1);
  m.at<double>(0, 0) = a; m.at<double>(1, 0) = b; m.at<double>(2, 0) = c;
  return m;
}

std::vector<double> pack_rt(const cv::Mat& R, const cv::Mat& t) {
  cv::Mat r;
  cv::Rodrigues(R, r);
  return {r.at<double>(0, 0), r.at<double>(1, 0), r.at<double>(2, 0),
          t.at<double>(0, 0), t.at<double>(1, 0), t.at<double>(2, 0)};
}

// [r1 r2 t] -> (R, t) with single-precision r1, r2 as multi_calib.h:130-137 (cv::Vec3f).
void split_homography(const cv::Mat& Rt, cv::Mat& R, cv::Mat& t) {
  const float r1[3] = {(float)Rt.at<double>(0, 0), (float)Rt.at<double>(1, 0), (float)Rt.at<double>(2, 0)};
  const float r2[3] = {(float)Rt.at<double>(0, 1), (float)Rt.at<double>(1, 1), (float)Rt.at<double>(2, 1)};
  const float r3[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
  R = cv::Mat(3, 3);
  for (int k = 0; k < 3; ++k) { R.at<double>(k, 0) = r1[k]; R.at<double>(k, 1) = r2[k]; R.at<double>(k, 2) = r3[k]; }
  t = vec3(Rt.at<double>(0, 2), Rt.at<double>(1, 2), Rt.at<double>(2, 2));
}

}  // namespace

MultiCalib_camera::MultiCalib_camera(double cx, double cy, double fx, double fy, double xi, double lamda,
                                     double alpha, double b, double c, cv::Mat R, cv::Mat t,
                                     std::vector<bool> has_chessboard,
                                     std::vector<std::vector<cv::Point2d>> pixel_coordinates) {
  rt_ = pack_rt(R, t);
  R_ = R; t_ = t;
  intrinsic_ = {fx, fy, cx, cy, xi, lamda, alpha, b, c};
  has_chessboard_ = has_chessboard;
  pixel_coordinates_ = pixel_coordinates;
  is_initial_ = true;
  // R_, t_ stay the matrices that were handed in (multi_calib.h:20-21): the pose-graph chain
  // composes them as they are, not their re-orthogonalised Rodrigues round trip
  intrinsic_matrix_ = cv::Mat(1, 9);
  for (int k = 0; k < 9; ++k) intrinsic_matrix_.at<double>(0, k) = intrinsic_[k];
}

void MultiCalib_camera::update_param() {
  cv::Rodrigues(vec3(rt_[0], rt_[1], rt_[2]), R_);
  t_ = vec3(rt_[3], rt_[4], rt_[5]);
  intrinsic_matrix_ = cv::Mat(1, 9);
  for (int k = 0; k < 9; ++k) intrinsic_matrix_.at<double>(0, k) = intrinsic_[k];
}

MultiCalib_chessboard::MultiCalib_chessboard(cv::Mat R, cv::Mat t) {
  rt_ = pack_rt(R, t);
  R_ = R; t_ = t;
  is_initial_ = true;
}

void MultiCalib_chessboard::update_param() {
  cv::Rodrigues(vec3(rt_[0], rt_[1], rt_[2]), R_);
  t_ = vec3(rt_[3], rt_[4], rt_[5]);
}

void MultiCalib::init_options() {
  tscm_options_init(&options_);          // Ceres defaults: 50 iterations (multi_calib.cpp:209-212)
  options_.verbose = 1;                  // summary.BriefReport() (multi_calib.cpp:218)
  std::memset(&summary_, 0, sizeof(summary_));
}

// Pose-graph initialisation (multi_calib.cpp:6-153): camera i is chained to camera i-1 through
// every board both see, the candidate with the smallest summed reprojection error wins;
// board poses likewise over the cameras that see them.
MultiCalib::MultiCalib(std::vector<TripleSphereCamera> cameras, const std::vector<cv::Point3d>& worlds) {
  init_options();
  worlds_ = worlds;
  const int camera_num = (int)cameras.size();
  const int board_num = (int)cameras[0].has_chessboard().size();
  cameras_.resize(camera_num);
  chessboards_.resize(board_num);
  for (int i = 0; i < camera_num; ++i) {
    cv::Mat R, t;
    if (i == 0) {
      R = cv::Mat::eye(3, 3);            // camera 0 is the reference frame (multi_calib.cpp:21-22)
      t = cv::Mat::zeros(3, 1);
    } else {
      if (!cameras_[i - 1].is_initial()) {
        std::cout << "[tscm] cameras must be given in adjacent order" << std::endl;
        return;
      }
      std::vector<cv::Mat> Rs, ts;
      cv::Mat Rk = cameras_[i - 1].R(), tk = cameras_[i - 1].t();
      for (int j = 0; j < board_num; ++j) {
        if (!cameras[i - 1].has_chessboard(j) || !cameras[i].has_chessboard(j)) continue;
        cv::Mat Ri, ti, Rp, tp;
        split_homography(cameras[i].Rt(j), Ri, ti);
        split_homography(cameras[i - 1].Rt(j), Rp, tp);
        cv::Mat R_ik = Ri * Rp.t();
        cv::Mat t_ik = ti - R_ik * tp;
        Rs.push_back(R_ik * Rk);
        ts.push_back(R_ik * tk + t_ik);
      }
      if (Rs.empty()) {
        std::cout << "[tscm] cameras " << i - 1 << " and " << i << " share no board" << std::endl;
        return;                          // (the reference indexes Rs[-1] here: multi_calib.cpp:51,86)
      }
      double best = 1e10; int best_id = 0;
      for (size_t c = 0; c < Rs.size(); ++c) {
        double error = 0;
        for (int k = 0; k < board_num; ++k) {
          if (!cameras[i - 1].has_chessboard(k) || !cameras[i].has_chessboard(k)) continue;
          cv::Mat Ri, ti, Rp, tp;
          split_homography(cameras[i].Rt(k), Ri, ti);
          cv::Mat R_ki = Rk * Rs[c].t();
          cv::Mat t_ki = tk - R_ki * ts[c];
          error += cameras[i - 1].ReprojectError(cameras[i - 1].pixels()[k], worlds, R_ki * Ri, R_ki * ti + t_ki);
          split_homography(cameras[i - 1].Rt(k), Rp, tp);
          cv::Mat R_ik = Rs[c] * Rk.t();
          cv::Mat t_ik = ts[c] - R_ik * tk;
          error += cameras[i].ReprojectError(cameras[i].pixels()[k], worlds, R_ik * Rp, R_ik * tp + t_ik);
        }
        if (error < best) { best = error; best_id = (int)c; }
      }
      R = Rs[best_id]; t = ts[best_id];
    }
    cameras_[i] = MultiCalib_camera(cameras[i].cx(), cameras[i].cy(), cameras[i].fx(), cameras[i].fy(),
                                    cameras[i].xi(), cameras[i].lamda(), cameras[i].alpha(), cameras[i].b(),
                                    cameras[i].c(), R, t, cameras[i].has_chessboard(), cameras[i].pixels());
  }
  for (int i = 0; i < board_num; ++i) {
    std::vector<int> ids;
    for (int j = 0; j < camera_num; ++j) if (cameras[j].has_chessboard(i)) ids.push_back(j);
    if (ids.empty()) continue;           // multi_calib.cpp:102
    std::vector<cv::Mat> Rs, ts;
    for (int id : ids) {
      cv::Mat Rb, tb;
      split_homography(cameras[id].Rt(i), Rb, tb);
      cv::Mat Rc = cameras_[id].R(), tc = cameras_[id].t();
      Rs.push_back(Rc.t() * Rb);
      ts.push_back(Rc.t() * (tb - tc));
    }
    int best_id = 0;
    if (ids.size() > 1) {
      double best = 1e10;
      for (size_t c = 0; c < Rs.size(); ++c) {
        double error = 0;
        for (int id : ids) {
          cv::Mat Rc = cameras_[id].R(), tc = cameras_[id].t();
          error += cameras[id].ReprojectError(cameras[id].pixels()[i], worlds, Rc * Rs[c], Rc * ts[c] + tc);
        }
        if (error < best) { best = error; best_id = (int)c; }
      }
    }
    chessboards_[i] = MultiCalib_chessboard(Rs[best_id], ts[best_id]);
  }
}

// Replaces multi_calib.cpp:155-284.
void MultiCalib::calibrate() {
  const int C = (int)cameras_.size(), B = (int)chessboards_.size(), K = (int)worlds_.size();
  std::vector<double> board_xy(2 * (size_t)K);
  for (int j = 0; j < K; ++j) { board_xy[2 * j] = worlds_[j].x; board_xy[2 * j + 1] = worlds_[j].y; }
  // Residual blocks exist for (camera m, board i) iff the board is initialised and the
  // camera detected it (pixels[i] non-empty): multi_calib.cpp:162-169.  Boards that end up
  // with no view are not part of the problem.
  std::vector<int> frame_id(B, -1);
  std::vector<int32_t> view_camera, view_frame;
  std::vector<char> used(B, 0);
  for (int m = 0; m < C; ++m) {
    const auto& px = cameras_[m].pixels_ref();
    for (int i = 0; i < B; ++i)
      if (chessboards_[i].is_initial() && (int)px[i].size() == K) used[i] = 1;
  }
  int F = 0;
  for (int i = 0; i < B; ++i) if (used[i]) frame_id[i] = F++;
  for (int m = 0; m < C; ++m) {
    const auto& px = cameras_[m].pixels_ref();
    for (int i = 0; i < B; ++i) {
      if (!chessboards_[i].is_initial() || (int)px[i].size() != K) continue;
      view_camera.push_back(m);
      view_frame.push_back(frame_id[i]);
    }
  }
  if (F == 0 || view_camera.empty()) { std::cout << "[tscm] nothing to calibrate" << std::endl; return; }
  // the detections are packed into page-locked staging memory (tscm_host_alloc), so that the
  // upload inside tscm_solve() runs at the full PCIe rate; pageable memory if that fails
  const size_t obs_doubles = view_camera.size() * 2 * (size_t)K;
  double* pinned = static_cast<double*>(tscm_host_alloc(obs_doubles * sizeof(double)));
  std::vector<double> pageable;
  if (!pinned) pageable.resize(obs_doubles);
  double* obs = pinned ? pinned : pageable.data();
  {
    size_t o = 0;
    for (int m = 0; m < C; ++m) {
      const auto& px = cameras_[m].pixels_ref();
      for (int i = 0; i < B; ++i) {
        if (!chessboards_[i].is_initial() || (int)px[i].size() != K) continue;
        std::memcpy(obs + o, px[i].data(), sizeof(double) * 2 * K);   // cv::Point2d = {x, y}
        o += 2 * (size_t)K;
      }
    }
  }
  std::vector<double> intr(9 * (size_t)C), cam_rt(6 * (size_t)C), board_rt(6 * (size_t)F);
  for (int m = 0; m < C; ++m) {
    std::memcpy(&intr[9 * m], cameras_[m].intrinsic_.data(), sizeof(double) * 9);
    std::memcpy(&cam_rt[6 * m], cameras_[m].rt_.data(), sizeof(double) * 6);
  }
  for (int i = 0; i < B; ++i)
    if (frame_id[i] >= 0) std::memcpy(&board_rt[6 * frame_id[i]], chessboards_[i].rt_.data(), sizeof(double) * 6);

  tscm_problem p;
  p.num_cameras = C; p.num_frames = F; p.corners_per_board = K; p.num_views = (int)view_camera.size();
  p.board_xy = board_xy.data(); p.view_camera = view_camera.data(); p.view_frame = view_frame.data();
  p.obs_xy = obs;
  p.fixed_camera = 0;                     // SetParameterBlockConstant(cameras_[0].rt_), multi_calib.cpp:186
  std::memset(&summary_, 0, sizeof(summary_));
  tscm_options opt = options_;
  if (quiet) opt.verbose = 0;
  const int rc = tscm_solve(&p, &opt, intr.data(), cam_rt.data(), board_rt.data(), &summary_, device);
  tscm_host_free(pinned);
  if (rc != TSCM_OK) { std::cout << "[tscm] calibrate failed: " << tscm_last_error() << std::endl; return; }
  for (int m = 0; m < C; ++m) {
    std::memcpy(cameras_[m].intrinsic_.data(), &intr[9 * m], sizeof(double) * 9);
    std::memcpy(cameras_[m].rt_.data(), &cam_rt[6 * m], sizeof(double) * 6);
    if (cameras_[m].is_initial()) cameras_[m].update_param();
  }
  for (int i = 0; i < B; ++i) {
    if (frame_id[i] >= 0) std::memcpy(chessboards_[i].rt_.data(), &board_rt[6 * frame_id[i]], sizeof(double) * 6);
    if (chessboards_[i].is_initial()) chessboards_[i].update_param();
  }
  // mean Euclidean reprojection error per camera and overall (multi_calib.cpp:235-283)
  camera_reprojection_error.assign(C, 0.0);
  double error_sum = 0; long total = 0;
  for (int m = 0; m < C; ++m) {
    const auto& px = cameras_[m].pixels_ref();
    cv::Mat Rc = cameras_[m].R(), tc = cameras_[m].t();
    double error = 0; long cnt = 0;
    for (int i = 0; i < B; ++i) {
      if (!chessboards_[i].is_initial()) continue;
      cv::Mat Rb = chessboards_[i].R(), tb = chessboards_[i].t();
      for (size_t j = 0; j < px[i].size(); ++j) {
        double w[3], q[3];
        for (int r = 0; r < 3; ++r)
          w[r] = Rb.at<double>(r, 0) * worlds_[j].x + Rb.at<double>(r, 1) * worlds_[j].y +
                 Rb.at<double>(r, 2) * worlds_[j].z + tb.at<double>(r, 0);
        for (int r = 0; r < 3; ++r)
          q[r] = Rc.at<double>(r, 0) * w[0] + Rc.at<double>(r, 1) * w[1] + Rc.at<double>(r, 2) * w[2] + tc.at<double>(r, 0);
        const double r2 = q[0] * q[0] + q[1] * q[1];
        const double d1 = std::sqrt(r2 + q[2] * q[2]);
        const double z1 = q[2] + cameras_[m].xi() * d1;
        const double d2 = std::sqrt(r2 + z1 * z1);
        const double z2 = z1 + cameras_[m].lamda() * d2;
        const double d3 = std::sqrt(r2 + z2 * z2);
        const double den = z2 + cameras_[m].alpha() / (1 - cameras_[m].alpha()) * d3;
        const double u = cameras_[m].fx() * q[0] / den + cameras_[m].b() * q[1] / den + cameras_[m].cx();
        const double v = cameras_[m].c() * q[0] / den + cameras_[m].fy() * q[1] / den + cameras_[m].cy();
        error += std::sqrt((px[i][j].x - u) * (px[i][j].x - u) + (px[i][j].y - v) * (px[i][j].y - v));
        ++cnt; ++total;
      }
    }
    error_sum += error;
    camera_reprojection_error[m] = cnt ? error / cnt : 0.0;
    if (!quiet) std::cout << "camera_" << m << " reprojection error: " << camera_reprojection_error[m] << std::endl;
  }
  average_reprojection_error = total ? error_sum / total : 0.0;
  if (!quiet) std::cout << "average reproject error: " << average_reprojection_error << std::endl;
}

// cv::FileStorage(YAML) dialect of main.cpp:305-319: "%YAML:1.0", one !!opencv-matrix per key.
bool MultiCalib::write_yaml(const std::string& filename) {
  FILE* f = std::fopen(filename.c_str(), "w");
  if (!f) return false;
  auto emit = [&](const char* name, int rows, int cols, const double* v) {
    std::fprintf(f, "%s: !!opencv-matrix\n   rows: %d\n   cols: %d\n   dt: d\n   data: [ ", name, rows, cols);
    for (int k = 0; k < rows * cols; ++k) {
      char buf[40];
      if (v[k] == std::floor(v[k]) && std::fabs(v[k]) < 1e15) std::snprintf(buf, sizeof(buf), "%.0f.", v[k]);
      else std::snprintf(buf, sizeof(buf), "%.16e", v[k]);
      std::fprintf(f, "%s%s", buf, k + 1 < rows * cols ? ((k % 2 == 1) ? ",\n       " : ", ") : " ]\n");
    }
  };
  std::fprintf(f, "%%YAML:1.0\n---\n");
  for (size_t i = 0; i < cameras_.size(); ++i) {
    char key[16];
    std::snprintf(key, sizeof(key), "cam%d", (int)i);
    emit(key, 1, 9, cameras_[i].intrinsic_.data());
    std::snprintf(key, sizeof(key), "Twc%d", (int)i);
    cv::Mat R = cameras_[i].R(), t = cameras_[i].t();
    double T[12];
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) T[4 * r + c] = R.at<double>(r, c);
      T[4 * r + 3] = t.at<double>(r, 0);
    }
    emit(key, 3, 4, T);
  }
  std::fclose(f);
  return true;
}
