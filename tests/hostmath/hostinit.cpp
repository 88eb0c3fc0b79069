// Test-only C wrappers around the C++ drop-in adapter (tscm_calib_b200/host/ts_camera.*), so
// that pytest can drive TripleSphereCamera's cold-start initialisation (TS.cpp:36-52,110-203),
// the cv shim's solveZ / solvePnPRansac, the full calibrate() and the remap tables through
// ctypes.  Not part of the product library.
#include <chrono>
#include <cstring>
#include <vector>

#include "../../tscm_calib_b200/host/multi_calib_b200.h"
#include "../../tscm_calib_b200/host/ts_camera.h"
#include "../../oracle/cv_calib3d_port.h"   // the restated cv::SVD::solveZ / cv::solvePnPRansac (oracle side)

namespace {
std::vector<std::vector<cv::Point2d>> unpack(const double* px, const unsigned char* has, int F, int K) {
  std::vector<std::vector<cv::Point2d>> pixels(F);
  for (int i = 0; i < F; ++i) {
    if (!has[i]) continue;                      // empty vector = no detection (main.cpp:33-37)
    pixels[i].resize(K);
    std::memcpy((void*)pixels[i].data(), px + (size_t)i * K * 2, sizeof(double) * 2 * K);
  }
  return pixels;
}
std::vector<cv::Point3d> board(int W, int H, double square) {   // main.cpp:12-18
  std::vector<cv::Point3d> w;
  for (int j = 0; j < W * H; ++j) w.push_back(cv::Point3d((j % W) * square, (j / W) * square, 0.0));
  return w;
}
void export_state(TripleSphereCamera& cam, int F, const unsigned char* has, double* intr9, double* Rt) {
  const double v[9] = {cam.fx(), cam.fy(), cam.cx(), cam.cy(), cam.xi(), cam.lamda(), cam.alpha(), cam.b(), cam.c()};
  std::memcpy(intr9, v, sizeof(v));
  for (int i = 0; i < F; ++i) {
    if (!has[i]) continue;
    cv::Mat M = cam.Rt(i);
    for (int k = 0; k < 9; ++k) Rt[(size_t)i * 9 + k] = M.at<double>(k / 3, k % 3);
  }
}
}  // namespace

extern "C" {

void hostinit_solve_z(const double* A, int rows, int cols, double* z) {
  cv::Mat M(rows, cols), out;
  for (int i = 0; i < rows * cols; ++i) M.at<double>(i / cols, i % cols) = A[i];
  cv::SVD::solveZ(M, out);
  for (int i = 0; i < cols; ++i) z[i] = out.at<double>(i, 0);
}

// cv::Rodrigues of the shim: dir 0 = 3-vector -> 3x3, dir 1 = 3x3 -> 3-vector.
void hostinit_rodrigues(const double* in, int dir, double* out) {
  if (dir == 0) {
    cv::Mat v(3, 1), R;
    for (int k = 0; k < 3; ++k) v.at<double>(k, 0) = in[k];
    cv::Rodrigues(v, R);
    for (int k = 0; k < 9; ++k) out[k] = R.at<double>(k / 3, k % 3);
  } else {
    cv::Mat R(3, 3), v;
    for (int k = 0; k < 9; ++k) R.at<double>(k / 3, k % 3) = in[k];
    cv::Rodrigues(R, v);
    for (int k = 0; k < 3; ++k) out[k] = v.at<double>(k, 0);
  }
}

int hostinit_solve_pnp(const double* obj_xyz, const double* img_xy, int n, double* rvec, double* tvec) {
  std::vector<cv::Point3d> o(n);
  std::vector<cv::Point2d> p(n);
  for (int i = 0; i < n; ++i) { o[i] = cv::Point3d(obj_xyz[3 * i], obj_xyz[3 * i + 1], obj_xyz[3 * i + 2]); p[i] = cv::Point2d(img_xy[2 * i], img_xy[2 * i + 1]); }
  cv::Mat r, t;
  if (!cv::solvePnPRansac(o, p, cv::Mat::eye(3, 3, cv::CV_64F), cv::Mat::zeros(4, 0, cv::CV_64F), r, t)) return 1;
  for (int k = 0; k < 3; ++k) { rvec[k] = r.at<double>(k, 0); tvec[k] = t.at<double>(k, 0); }
  return 0;
}

// TS.cpp:36-52 from scratch (guess7 == NULL) or from given intrinsics.
int hostinit_initial_guess(const double* px, const unsigned char* has, int F, int W, int H, double square,
                           int img_w, int img_h, const double* guess7, double* intr9, double* Rt) {
  TripleSphereCamera cam = guess7 ? TripleSphereCamera(guess7[0], guess7[1], guess7[2], guess7[3], guess7[4], guess7[5], guess7[6])
                                  : TripleSphereCamera();
  std::vector<bool> hb(F);
  for (int i = 0; i < F; ++i) hb[i] = has[i] != 0;
  const bool ok = cam.initial_guess(unpack(px, has, F, W * H), hb, board(W, H, square), cv::Size(img_w, img_h), cv::Size(W, H));
  export_state(cam, F, has, intr9, Rt);
  return ok ? 0 : 1;
}

// The whole TripleSphereCamera::calibrate (TS.cpp:30-108): cold start + refinement on the GPU.
// summary4 = {termination, iterations, initial cost, final cost}.
int hostinit_calibrate(const double* px, const unsigned char* has, int F, int W, int H, double square,
                       int img_w, int img_h, const double* guess7, double* intr9, double* Rt,
                       double* summary4) {
  TripleSphereCamera cam = guess7 ? TripleSphereCamera(guess7[0], guess7[1], guess7[2], guess7[3], guess7[4], guess7[5], guess7[6])
                                  : TripleSphereCamera();
  cam.device = 0;
  std::vector<bool> hb(F);
  for (int i = 0; i < F; ++i) hb[i] = has[i] != 0;
  const bool ok = cam.calibrate(unpack(px, has, F, W * H), hb, board(W, H, square), cv::Size(img_w, img_h), cv::Size(W, H));
  export_state(cam, F, has, intr9, Rt);
  const tscm_summary& s = cam.last_summary();
  summary4[0] = s.termination_type; summary4[1] = s.num_iterations; summary4[2] = s.initial_cost; summary4[3] = s.final_cost;
  return ok ? 0 : 1;
}

// The plain-double projection family of the unchanged TS interface through the
// adapter: project (TS.cpp:332-344), Reproject (TS.cpp:227-245), ReprojectError (TS.h:58-69 /
// TS.cpp:205-225), get_unit_sphere_coordinate (TS.h:39-57).
int hostinit_project(const double* intr9, const double* pts, int n, double* uv) {
  TripleSphereCamera cam(intr9[0], intr9[1], intr9[2], intr9[3], intr9[4], intr9[5], intr9[6]);
  for (int i = 0; i < n; ++i) {
    cv::Mat P(3, 1);
    for (int k = 0; k < 3; ++k) P.at<double>(k, 0) = pts[3 * i + k];
    const cv::Point2d q = cam.project(P);
    uv[2 * i] = q.x; uv[2 * i + 1] = q.y;
  }
  return 0;
}
int hostinit_reproject(const double* intr9, const double* Rt9, const double* worlds, int n, double* uv) {
  TripleSphereCamera cam(intr9[0], intr9[1], intr9[2], intr9[3], intr9[4], intr9[5], intr9[6]);
  cv::Mat M(3, 3);
  for (int k = 0; k < 9; ++k) M.at<double>(k / 3, k % 3) = Rt9[k];
  std::vector<cv::Point3d> w(n);
  for (int i = 0; i < n; ++i) w[i] = cv::Point3d(worlds[3 * i], worlds[3 * i + 1], worlds[3 * i + 2]);
  std::vector<cv::Point2d> px;
  cam.Reproject(w, M, px);
  if ((int)px.size() != n) return 1;
  for (int i = 0; i < n; ++i) { uv[2 * i] = px[i].x; uv[2 * i + 1] = px[i].y; }
  return 0;
}
double hostinit_reproject_error(const double* intr9, const double* pixels, const double* worlds, int n,
                                const double* R9, const double* t3) {
  TripleSphereCamera cam(intr9[0], intr9[1], intr9[2], intr9[3], intr9[4], intr9[5], intr9[6]);
  cv::Mat R(3, 3), t(3, 1);
  for (int k = 0; k < 9; ++k) R.at<double>(k / 3, k % 3) = R9[k];
  for (int k = 0; k < 3; ++k) t.at<double>(k, 0) = t3[k];
  std::vector<cv::Point3d> w(n);
  std::vector<cv::Point2d> px(n);
  for (int i = 0; i < n; ++i) {
    w[i] = cv::Point3d(worlds[3 * i], worlds[3 * i + 1], worlds[3 * i + 2]);
    px[i] = cv::Point2d(pixels[2 * i], pixels[2 * i + 1]);
  }
  return cam.ReprojectError(px, w, R, t);
}
int hostinit_unit_sphere(const double* intr9, const double* pixels, int n, double* xyz) {
  TripleSphereCamera cam(intr9[0], intr9[1], intr9[2], intr9[3], intr9[4], intr9[5], intr9[6]);
  for (int i = 0; i < n; ++i) {
    const cv::Point3d p = cam.get_unit_sphere_coordinate(cv::Point2d(pixels[2 * i], pixels[2 * i + 1]));
    xyz[3 * i] = p.x; xyz[3 * i + 1] = p.y; xyz[3 * i + 2] = p.z;
  }
  return 0;
}

// TS.cpp:284-306 and 308-326 through the adapter.
int hostinit_undistort(const double* intr9, double fx, double fy, double cx, double cy, int w, int h,
                       float* mapx, float* mapy) {
  TripleSphereCamera cam(intr9[0], intr9[1], intr9[2], intr9[3], intr9[4], intr9[5], intr9[6]);
  cam.device = 0;
  cv::Mat mx, my;
  cam.undistort(fx, fy, cx, cy, cv::Size(w, h), mx, my);
  if (mx.empty() || mx.rows != h || mx.cols != w) return 1;
  std::memcpy(mapx, mx.ptr<float>(), sizeof(float) * (size_t)w * h);
  std::memcpy(mapy, my.ptr<float>(), sizeof(float) * (size_t)w * h);
  return 0;
}

int hostinit_undistort_chessboard(const double* intr9, const double* Rt9, int W, int H, double square,
                                  float* mapx, float* mapy, int* out_w, int* out_h) {
  TripleSphereCamera cam(intr9[0], intr9[1], intr9[2], intr9[3], intr9[4], intr9[5], intr9[6]);
  cam.device = 0;
  cv::Mat M(3, 3);
  for (int k = 0; k < 9; ++k) M.at<double>(k / 3, k % 3) = Rt9[k];
  cam.setRt(std::vector<cv::Mat>{M});
  cam.setHasChessboard(std::vector<bool>{true});
  cv::Mat mx, my;
  if (!cam.undistort_chessboard_maps(0, cv::Size(W, H), square, mx, my)) return 1;
  *out_w = mx.cols; *out_h = mx.rows;
  std::memcpy(mapx, mx.ptr<float>(), sizeof(float) * (size_t)mx.cols * mx.rows);
  std::memcpy(mapy, my.ptr<float>(), sizeof(float) * (size_t)mx.cols * mx.rows);
  return 0;
}

// MultiCalib's pose-graph initialisation (multi_calib.cpp:6-153) from per-camera mono results:
// px [C][F][K][2], has [C][F], intr [C][9], Rt [C][F][9] (each camera's [r1 r2 t] board poses).
// Outputs per camera / per board: R (9), t (3), rt_ (6); board_init[F] = is_initial().
int hostinit_pose_graph(int C, int F, int W, int H, double square, const double* px, const unsigned char* has,
                        const double* intr, const double* Rt, double* cam_R, double* cam_t, double* cam_rt,
                        double* board_R, double* board_t, double* board_rt, unsigned char* board_init) {
  const int K = W * H;
  std::vector<TripleSphereCamera> cams;
  for (int m = 0; m < C; ++m) {
    const double* in = intr + 9 * m;
    TripleSphereCamera cam(in[0], in[1], in[2], in[3], in[4], in[5], in[6]);
    std::vector<bool> hb(F);
    std::vector<cv::Mat> Rts(F);
    for (int i = 0; i < F; ++i) {
      hb[i] = has[(size_t)m * F + i] != 0;
      cv::Mat M(3, 3);
      for (int k = 0; k < 9; ++k) M.at<double>(k / 3, k % 3) = Rt[((size_t)m * F + i) * 9 + k];
      Rts[i] = M;
    }
    cam.setRt(Rts);
    cam.setHasChessboard(hb);
    cam.setPixels(unpack(px + (size_t)m * F * K * 2, has + (size_t)m * F, F, K));
    cams.push_back(cam);
  }
  MultiCalib calib(cams, board(W, H, square));
  if ((int)calib.cameras_.size() != C || (int)calib.chessboards_.size() != F) return 1;
  for (int m = 0; m < C; ++m) {
    if (!calib.cameras_[m].is_initial()) return 2;
    cv::Mat R = calib.cameras_[m].R(), t = calib.cameras_[m].t();
    for (int k = 0; k < 9; ++k) cam_R[9 * m + k] = R.at<double>(k / 3, k % 3);
    for (int k = 0; k < 3; ++k) cam_t[3 * m + k] = t.at<double>(k, 0);
    for (int k = 0; k < 6; ++k) cam_rt[6 * m + k] = calib.cameras_[m].rt_[k];
  }
  for (int i = 0; i < F; ++i) {
    board_init[i] = calib.chessboards_[i].is_initial() ? 1 : 0;
    if (!board_init[i]) continue;
    cv::Mat R = calib.chessboards_[i].R(), t = calib.chessboards_[i].t();
    for (int k = 0; k < 9; ++k) board_R[9 * i + k] = R.at<double>(k / 3, k % 3);
    for (int k = 0; k < 3; ++k) board_t[3 * i + k] = t.at<double>(k, 0);
    for (int k = 0; k < 6; ++k) board_rt[6 * i + k] = calib.chessboards_[i].rt_[k];
  }
  return 0;
}

// The reference's whole flow from corners only (main.cpp:57-129,283-289): per camera a cold-start
// mono calibration (TS.cpp:30-108, refinement on the GPU), the pose graph (multi_calib.cpp:6-153),
// the joint refinement (multi_calib.cpp:155-283, on the GPU).
// summary5 = {mono calibrations that converged, termination, iterations, final cost, mean error}.
// stage_s (optional) = seconds of {mono calibrations, pose graph, joint refinement}.
int hostinit_full_pipeline_timed(int C, int F, int W, int H, double square, int img_w, int img_h, const double* px,
                                 const unsigned char* has, double* intr, double* cam_rt, double* board_rt,
                                 double* summary5, double* stage_s) {
  using clk = std::chrono::steady_clock;
  auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
  const auto t0 = clk::now();
  const int K = W * H;
  const std::vector<cv::Point3d> worlds = board(W, H, square);
  std::vector<TripleSphereCamera> cams;
  int converged = 0;
  for (int m = 0; m < C; ++m) {
    TripleSphereCamera cam;
    cam.device = 0;
    std::vector<bool> hb(F);
    for (int i = 0; i < F; ++i) hb[i] = has[(size_t)m * F + i] != 0;
    if (cam.calibrate(unpack(px + (size_t)m * F * K * 2, has + (size_t)m * F, F, K), hb, worlds,
                      cv::Size(img_w, img_h), cv::Size(W, H)))
      ++converged;
    cams.push_back(cam);
  }
  const auto t1 = clk::now();
  MultiCalib calib(cams, worlds);
  const auto t2 = clk::now();
  calib.device = 0;
  calib.calibrate();
  const auto t3 = clk::now();
  if (stage_s) { stage_s[0] = secs(t0, t1); stage_s[1] = secs(t1, t2); stage_s[2] = secs(t2, t3); }
  for (int m = 0; m < C; ++m) {
    std::memcpy(intr + 9 * m, calib.cameras_[m].intrinsic_.data(), 72);
    std::memcpy(cam_rt + 6 * m, calib.cameras_[m].rt_.data(), 48);
  }
  for (int i = 0; i < F; ++i)
    if (calib.chessboards_[i].is_initial()) std::memcpy(board_rt + 6 * i, calib.chessboards_[i].rt_.data(), 48);
  const tscm_summary& s = calib.last_summary();
  summary5[0] = converged; summary5[1] = s.termination_type; summary5[2] = s.num_iterations;
  summary5[3] = s.final_cost; summary5[4] = calib.average_reprojection_error;
  return 0;
}
int hostinit_full_pipeline(int C, int F, int W, int H, double square, int img_w, int img_h, const double* px,
                           const unsigned char* has, double* intr, double* cam_rt, double* board_rt,
                           double* summary5) {
  return hostinit_full_pipeline_timed(C, F, W, H, square, img_w, img_h, px, has, intr, cam_rt, board_rt, summary5,
                                      nullptr);
}

}  // extern "C"
