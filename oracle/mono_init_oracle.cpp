// mono_init_oracle.cpp — TEST INFRASTRUCTURE.  CPU restatement of the mono cold start of
// TripleSphereCamera::calibrate, /root/reference/TS.cpp:36-52:
//   estimate_focal      TS.cpp:110-168   circle fit per board row through cv::SVD::solveZ
//   estimate_extrinsic  TS.cpp:170-203   unit-sphere lift (TS.h:39-57), facing rotation,
//                                        cv::solvePnPRansac on the normalised plane, [r1 r2 t]
// on flat arrays, with the two OpenCV calls restated in cv_calib3d_port.h (pinned to golden
// vectors of the real OpenCV 4.13: tests/golden/mono_init.npz, tests/test_mono_init.py).
// Only tests/ and bench legs may call this; the product path is csrc/tscm_monoinit.cu.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "cv_calib3d_port.h"

namespace {

struct Intr { double fx, fy, cx, cy, xi, lamda, alpha, b, c; };

// TS.h:39-57
void unit_sphere(const Intr& I, double px, double py, const double* T, double* o) {
  const double x0 = px - I.cx, y0 = py - I.cy;
  const double det = I.fx * I.fy - I.b * I.c;
  const double mx = (I.fy * x0 - I.b * y0) / det, my = (-I.c * x0 + I.fx * y0) / det;
  const double k = I.alpha / (1 - I.alpha);
  const double r2 = mx * mx + my * my;
  const double gamma = (k + std::sqrt(1 + (1 - k * k) * r2)) / (r2 + 1);
  const double gk = gamma - k;
  const double eta = I.lamda * gk + std::sqrt((gk * gk - 1) * I.lamda * I.lamda + 1);
  const double mz = eta * gk;
  const double ml = mz - I.lamda;
  const double mu = I.xi * ml + std::sqrt(I.xi * I.xi * (ml * ml - 1) + 1);
  const double v[3] = {mu * eta * gamma * mx, mu * eta * gamma * my, mu * ml - I.xi};
  for (int r = 0; r < 3; ++r) o[r] = T[3 * r] * v[0] + T[3 * r + 1] * v[1] + T[3 * r + 2] * v[2];
}

// One board row: TS.cpp:127-158.  false = row skipped.
bool focal_from_row(const double* row_xy, int count, double cx, double cy, double* focal) {
  cv::Mat design(count, 4);
  for (int j = 0; j < count; ++j) {
    const double x = row_xy[2 * j] - cx, y = row_xy[2 * j + 1] - cy;
    design.at<double>(j, 0) = x;
    design.at<double>(j, 1) = y;
    design.at<double>(j, 2) = 0.5;
    design.at<double>(j, 3) = -0.5 * (x * x + y * y);
  }
  cv::Mat z;
  cv::SVD::solveZ(design, z);
  const double c[4] = {z.at<double>(0), z.at<double>(1), z.at<double>(2), z.at<double>(3)};
  const double t = c[0] * c[0] + c[1] * c[1] + c[2] * c[3];
  if (t < 0) return false;
  const double d = std::sqrt(1 / t);
  const double nx = c[0] * d, ny = c[1] * d;
  const double ob = nx * nx + ny * ny;
  if (ob > 0.95) return false;
  *focal = std::fabs(c[2] * d / std::sqrt(1 - ob));
  return true;
}

}  // namespace

extern "C" {

// pixels [F][K][2], has [F], worlds [K][3].  has_init_guess != 0: intr9 is input (TS.cpp:41);
// otherwise cx, cy, xi, lamda, alpha are set as TS.cpp:43-47 and the focal length estimated.
// Outputs: intr9, mono_rt [F][9] (zero where no pose), frame_ok [F], rows_used.
// Returns 0, or 1 when the focal estimate failed (fx == 0: TS.cpp:50 returns false).
int tscm_oracle_mono_init(int F, int W, int H, int img_w, int img_h, const double* worlds,
                          const uint8_t* has, const double* pixels, int has_init_guess, double* intr9,
                          double* mono_rt, uint8_t* frame_ok, int32_t* rows_used) {
  const int K = W * H;
  Intr I;
  if (has_init_guess) {
    I = Intr{intr9[0], intr9[1], intr9[2], intr9[3], intr9[4], intr9[5], intr9[6], intr9[7], intr9[8]};
  } else {
    I = Intr{0, 0, img_w / 2 - 0.5, img_h / 2 - 0.5, 0.0, 0.0, 0.5, 0.0, 0.0};   // integer halves
    double sum = 0;
    int used = 0;
    for (int k = 0; k < F; ++k) {
      if (!has[k]) continue;                                  // pixels[k].size() == 0
      for (int r = 0; r < H; ++r) {
        double f;
        if (!focal_from_row(pixels + ((size_t)k * K + (size_t)r * W) * 2, W, I.cx, I.cy, &f)) continue;
        sum += f;
        ++used;
      }
    }
    if (rows_used) *rows_used = used;
    I.fx = I.fy = used > 0 ? sum / used : 0.0;
  }
  const double out[9] = {I.fx, I.fy, I.cx, I.cy, I.xi, I.lamda, I.alpha, I.b, I.c};
  std::memcpy(intr9, out, sizeof(out));
  std::memset(mono_rt, 0, sizeof(double) * 9 * (size_t)F);
  for (int k = 0; k < F; ++k) frame_ok[k] = 0;
  if (!has_init_guess && I.fx == 0) return 1;
  std::vector<cv::Point3d> obj(K);
  for (int j = 0; j < K; ++j) obj[j] = cv::Point3d(worlds[3 * j], worlds[3 * j + 1], worlds[3 * j + 2]);
  const double eye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  const cv::Mat identity = cv::Mat::eye(3, 3, cv::CV_64F);
  for (int k = 0; k < F; ++k) {
    if (!has[k]) continue;
    const double* px = pixels + (size_t)k * K * 2;
    const size_t centre = (size_t)K / 2 - W / 2 - 1;                               // TS.cpp:177
    double p[3];
    unit_sphere(I, px[2 * centre], px[2 * centre + 1], eye, p);
    const double az = std::atan2(p[0], p[2]), el = std::asin(p[1]);               // TS.cpp:178-179
    const double R1[9] = {std::cos(az), 0, -std::sin(az), 0, 1, 0, std::sin(az), 0, std::cos(az)};
    const double R2[9] = {1, 0, 0, 0, std::cos(el), -std::sin(el), 0, std::sin(el), std::cos(el)};
    double T[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        double s = 0.0;
        for (int q = 0; q < 3; ++q) s += R2[3 * r + q] * R1[3 * q + c];
        T[3 * r + c] = s;
      }
    std::vector<cv::Point2d> plane(K);
    for (int i = 0; i < K; ++i) {
      double ray[3];
      unit_sphere(I, px[2 * i], px[2 * i + 1], T, ray);
      plane[i] = cv::Point2d(ray[0] / ray[2], ray[1] / ray[2]);
    }
    cv::Mat rvec, tvec, pose;
    const bool found = cv::solvePnPRansac(obj, plane, identity, cv::Mat::zeros(4, 0, cv::CV_64F), rvec, tvec);
    if (!found || rvec.empty() || tvec.empty()) continue;      // (OpenCV would raise here)
    cv::Rodrigues(rvec, pose);
    double* M = mono_rt + 9 * (size_t)k;
    // Rt = transform.t() * Rt; tvec = transform.t() * tvec; third column <- tvec (TS.cpp:195-200)
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 2; ++c) {
        double s = 0.0;
        for (int q = 0; q < 3; ++q) s += T[3 * q + r] * pose.at<double>(q, c);
        M[3 * r + c] = s;
      }
      double s = 0.0;
      for (int q = 0; q < 3; ++q) s += T[3 * q + r] * tvec.at<double>(q, 0);
      M[3 * r + 2] = s;
    }
    frame_ok[k] = 1;
  }
  return 0;
}

// the two restated OpenCV calls on their own, for the golden-vector tests
void tscm_oracle_solve_z(const double* A, int rows, int cols, double* z) {
  cv::Mat M(rows, cols), out;
  for (int i = 0; i < rows * cols; ++i) M.at<double>(i / cols, i % cols) = A[i];
  cv::SVD::solveZ(M, out);
  for (int i = 0; i < cols; ++i) z[i] = out.at<double>(i, 0);
}
int tscm_oracle_solve_pnp(const double* obj_xyz, const double* img_xy, int n, double* rvec, double* tvec) {
  std::vector<cv::Point3d> o(n);
  std::vector<cv::Point2d> p(n);
  for (int i = 0; i < n; ++i) {
    o[i] = cv::Point3d(obj_xyz[3 * i], obj_xyz[3 * i + 1], obj_xyz[3 * i + 2]);
    p[i] = cv::Point2d(img_xy[2 * i], img_xy[2 * i + 1]);
  }
  cv::Mat r, t;
  if (!cv::solvePnPRansac(o, p, cv::Mat::eye(3, 3, cv::CV_64F), cv::Mat::zeros(4, 0, cv::CV_64F), r, t)) return 1;
  if (r.empty() || t.empty()) return 1;
  for (int k = 0; k < 3; ++k) { rvec[k] = r.at<double>(k, 0); tvec[k] = t.at<double>(k, 0); }
  return 0;
}

}  // extern "C"
