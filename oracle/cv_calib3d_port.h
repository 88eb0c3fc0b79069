// cv_calib3d_port.h — TEST INFRASTRUCTURE (oracle).  CPU restatement of the two OpenCV calls the
// reference's mono cold start delegates to:
//   cv::SVD::solveZ        /root/reference/TS.cpp:142-143   (estimate_focal, circle fits)
//   cv::solvePnPRansac     /root/reference/TS.cpp:193       (estimate_extrinsic)
// OpenCV is a third-party dependency of the reference that is absent from /root/reference and
// from this image's C++ toolchain (find_package(OpenCV), CMakeLists.txt); the Python wheel
// (cv2 4.13) IS present, so both restatements are pinned against golden vectors of the real
// library (tests/golden/mono_init.npz, cv_shim.npz; make_golden_init.py, make_golden_cvshim.py).
// The product path (csrc/tscm_monoinit.cu) never includes this file; only oracle/ and tests/ do.
#pragma once

#include <algorithm>
#include <cmath>
#include <vector>

#include "../tscm_calib_b200/host/cv_compat.h"   // the cv::Mat / Point type shim only

#ifndef TSCM_USE_OPENCV
namespace cv {

// cv::SVD::solveZ: unit vector z minimising |A z| (right singular vector of the smallest
// singular value; sign arbitrary, as in OpenCV).  One-sided Jacobi (Hestenes) SVD, the same
// family of algorithm as OpenCV's JacobiSVDImpl_.  Used by estimate_focal (TS.cpp:142-143).
struct SVD {
  static void solveZ(const Mat& A, Mat& z) {
    const int m = A.rows, n = A.cols;
    std::vector<double> W((size_t)m * n), V((size_t)n * n, 0.0);
    for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) W[(size_t)i * n + j] = A.at<double>(i, j);
    for (int j = 0; j < n; ++j) V[(size_t)j * n + j] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
      bool rotated = false;
      for (int p = 0; p < n - 1; ++p)
        for (int q = p + 1; q < n; ++q) {
          double a = 0, b = 0, g = 0;
          for (int i = 0; i < m; ++i) {
            const double wp = W[(size_t)i * n + p], wq = W[(size_t)i * n + q];
            a += wp * wp; b += wq * wq; g += wp * wq;
          }
          if (std::fabs(g) <= 1e-300 || std::fabs(g) <= 2.220446049250313e-16 * std::sqrt(a * b)) continue;
          rotated = true;
          const double zeta = (b - a) / (2.0 * g);
          const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
          const double c = 1.0 / std::sqrt(1.0 + t * t), sn = c * t;
          for (int i = 0; i < m; ++i) {
            double& wp = W[(size_t)i * n + p]; double& wq = W[(size_t)i * n + q];
            const double x = wp, y = wq;
            wp = c * x - sn * y; wq = sn * x + c * y;
          }
          for (int i = 0; i < n; ++i) {
            double& vp = V[(size_t)i * n + p]; double& vq = V[(size_t)i * n + q];
            const double x = vp, y = vq;
            vp = c * x - sn * y; vq = sn * x + c * y;
          }
        }
      if (!rotated) break;
    }
    int best = 0; double best_norm = -1.0;
    for (int j = 0; j < n; ++j) {
      double s = 0;
      for (int i = 0; i < m; ++i) s += W[(size_t)i * n + j] * W[(size_t)i * n + j];
      if (best_norm < 0 || s < best_norm) { best_norm = s; best = j; }
    }
    z = Mat(n, 1);
    double nz = 0;
    for (int i = 0; i < n; ++i) nz += V[(size_t)i * n + best] * V[(size_t)i * n + best];
    nz = std::sqrt(nz);
    for (int i = 0; i < n; ++i) z.at<double>(i, 0) = V[(size_t)i * n + best] / nz;
  }
};

namespace detail {
// 6x6 (or smaller) symmetric positive-definite solve, in place; false if not SPD.
inline bool chol_solve(double* A, double* b, int n) {
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0)) return false;
    d = std::sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double v = A[i * n + j];
      for (int k = 0; k < j; ++k) v -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = v / d;
    }
  }
  for (int i = 0; i < n; ++i) { double v = b[i]; for (int k = 0; k < i; ++k) v -= A[i * n + k] * b[k]; b[i] = v / A[i * n + i]; }
  for (int i = n - 1; i >= 0; --i) { double v = b[i]; for (int k = i + 1; k < n; ++k) v -= A[k * n + i] * b[k]; b[i] = v / A[i * n + i]; }
  return true;
}
}  // namespace detail

// solvePnP(SOLVEPNP_ITERATIVE) for a PLANAR target (every objectPoints[i].z equal).
// OpenCV's iterative planar path = homography initialisation + Levenberg-Marquardt on the
// reprojection error; restated here as a normalised DLT homography + damped Gauss-Newton
// run to convergence, i.e. the same minimiser (OpenCV stops at 20 iterations / FLT_EPSILON).
inline bool solvePnPPlanarIterative(const std::vector<Point3d>& obj, const std::vector<Point2d>& img, const Mat& K,
                                    Mat& rvec, Mat& tvec) {
  const int n = (int)obj.size();
  if (n < 4 || (int)img.size() != n) return false;
  const double kfx = K.at<double>(0, 0), kfy = K.at<double>(1, 1), kcx = K.at<double>(0, 2), kcy = K.at<double>(1, 2);
  std::vector<double> x(n), y(n), X(n), Y(n);
  double mx = 0, my = 0, mX = 0, mY = 0;
  for (int i = 0; i < n; ++i) {
    x[i] = (img[i].x - kcx) / kfx; y[i] = (img[i].y - kcy) / kfy; X[i] = obj[i].x; Y[i] = obj[i].y;
    mx += x[i]; my += y[i]; mX += X[i]; mY += Y[i];
  }
  mx /= n; my /= n; mX /= n; mY /= n;
  double sx = 0, sX = 0;
  for (int i = 0; i < n; ++i) {
    sx += std::sqrt((x[i] - mx) * (x[i] - mx) + (y[i] - my) * (y[i] - my));
    sX += std::sqrt((X[i] - mX) * (X[i] - mX) + (Y[i] - mY) * (Y[i] - mY));
  }
  if (!(sx > 0) || !(sX > 0)) return false;
  sx = std::sqrt(2.0) * n / sx; sX = std::sqrt(2.0) * n / sX;      // Hartley normalisation
  Mat A(2 * n, 9);
  for (int i = 0; i < n; ++i) {
    const double u = (x[i] - mx) * sx, v = (y[i] - my) * sx, a = (X[i] - mX) * sX, b = (Y[i] - mY) * sX;
    const double r0[9] = {a, b, 1, 0, 0, 0, -u * a, -u * b, -u};
    const double r1[9] = {0, 0, 0, a, b, 1, -v * a, -v * b, -v};
    for (int k = 0; k < 9; ++k) { A.at<double>(2 * i, k) = r0[k]; A.at<double>(2 * i + 1, k) = r1[k]; }
  }
  Mat hz;
  SVD::solveZ(A, hz);
  double Hn[9];
  for (int k = 0; k < 9; ++k) Hn[k] = hz.at<double>(k, 0);
  // H = T_img^-1 * Hn * T_obj
  const double Ti[9] = {1 / sx, 0, mx, 0, 1 / sx, my, 0, 0, 1};
  const double To[9] = {sX, 0, -mX * sX, 0, sX, -mY * sX, 0, 0, 1};
  double T1[9], H[9];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { double s = 0; for (int k = 0; k < 3; ++k) s += Ti[3 * r + k] * Hn[3 * k + c]; T1[3 * r + c] = s; }
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { double s = 0; for (int k = 0; k < 3; ++k) s += T1[3 * r + k] * To[3 * k + c]; H[3 * r + c] = s; }
  // pose from H = s [r1 r2 t]; the board must lie in front of the camera
  const double n1 = std::sqrt(H[0] * H[0] + H[3] * H[3] + H[6] * H[6]);
  const double n2 = std::sqrt(H[1] * H[1] + H[4] * H[4] + H[7] * H[7]);
  if (!(n1 > 0) || !(n2 > 0)) return false;
  double sc = 2.0 / (n1 + n2);
  const double z0 = obj[0].z;
  if ((H[6] * mX + H[7] * mY + H[8]) * sc < 0) sc = -sc;           // depth of the board centre > 0
  double r1[3] = {H[0] * sc, H[3] * sc, H[6] * sc}, r2[3] = {H[1] * sc, H[4] * sc, H[7] * sc};
  double t[3] = {H[2] * sc, H[5] * sc, H[8] * sc};
  // orthonormalise (r1, r2) symmetrically, r3 = r1 x r2
  {
    const double a1 = std::sqrt(r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2]);
    const double a2 = std::sqrt(r2[0] * r2[0] + r2[1] * r2[1] + r2[2] * r2[2]);
    for (int k = 0; k < 3; ++k) { r1[k] /= a1; r2[k] /= a2; }
    double s[3], d[3];
    for (int k = 0; k < 3; ++k) { s[k] = r1[k] + r2[k]; d[k] = r1[k] - r2[k]; }
    const double ns = std::sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
    double dd = (d[0] * s[0] + d[1] * s[1] + d[2] * s[2]) / (ns * ns);
    for (int k = 0; k < 3; ++k) d[k] -= dd * s[k];
    const double nd = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    const double h = std::sqrt(0.5);
    for (int k = 0; k < 3; ++k) { r1[k] = h * (s[k] / ns + d[k] / nd); r2[k] = h * (s[k] / ns - d[k] / nd); }
  }
  double R[9] = {r1[0], r2[0], r1[1] * r2[2] - r1[2] * r2[1],
                 r1[1], r2[1], r1[2] * r2[0] - r1[0] * r2[2],
                 r1[2], r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
  // the homography was fitted for z = z0 == 0; a constant offset moves t along r3
  for (int k = 0; k < 3; ++k) t[k] -= R[3 * k + 2] * z0;
  // damped Gauss-Newton on sum |(Px/Pz, Py/Pz) - (x, y)|^2, left-multiplicative rotation update
  auto cost_at = [&](const double* Rm, const double* tm) {
    double c = 0;
    for (int i = 0; i < n; ++i) {
      const double p[3] = {obj[i].x, obj[i].y, obj[i].z};
      double P[3];
      for (int r = 0; r < 3; ++r) P[r] = Rm[3 * r] * p[0] + Rm[3 * r + 1] * p[1] + Rm[3 * r + 2] * p[2] + tm[r];
      const double eu = P[0] / P[2] - x[i], ev = P[1] / P[2] - y[i];
      c += eu * eu + ev * ev;
    }
    return c;
  };
  double lambda = 1e-6, cost = cost_at(R, t);
  for (int it = 0; it < 200; ++it) {
    double JtJ[36] = {0}, Jtr[6] = {0};
    for (int i = 0; i < n; ++i) {
      const double p[3] = {obj[i].x, obj[i].y, obj[i].z};
      double q[3], P[3];
      for (int r = 0; r < 3; ++r) { q[r] = R[3 * r] * p[0] + R[3 * r + 1] * p[1] + R[3 * r + 2] * p[2]; P[r] = q[r] + t[r]; }
      const double iz = 1.0 / P[2], u = P[0] * iz, v = P[1] * iz;
      // dP/d(omega) = -[q]x, dP/dt = I
      const double dP[3][6] = {{0, q[2], -q[1], 1, 0, 0}, {-q[2], 0, q[0], 0, 1, 0}, {q[1], -q[0], 0, 0, 0, 1}};
      double Ju[6], Jv[6];
      for (int k = 0; k < 6; ++k) { Ju[k] = iz * (dP[0][k] - u * dP[2][k]); Jv[k] = iz * (dP[1][k] - v * dP[2][k]); }
      const double eu = u - x[i], ev = v - y[i];
      for (int a = 0; a < 6; ++a) {
        Jtr[a] += Ju[a] * eu + Jv[a] * ev;
        for (int b = 0; b <= a; ++b) JtJ[a * 6 + b] += Ju[a] * Ju[b] + Jv[a] * Jv[b];
      }
    }
    for (int a = 0; a < 6; ++a) for (int b = a + 1; b < 6; ++b) JtJ[a * 6 + b] = JtJ[b * 6 + a];
    bool improved = false;
    double step_norm = 0;
    for (int tries = 0; tries < 30 && !improved; ++tries) {
      double Am[36], d[6];
      for (int k = 0; k < 36; ++k) Am[k] = JtJ[k];
      for (int k = 0; k < 6; ++k) { Am[k * 6 + k] *= 1.0 + lambda; d[k] = -Jtr[k]; }
      if (!detail::chol_solve(Am, d, 6)) { lambda *= 10; continue; }
      Mat w(3, 1), dR;
      for (int k = 0; k < 3; ++k) w.at<double>(k, 0) = d[k];
      Rodrigues(w, dR);
      double Rn[9], tn[3];
      for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { double s = 0; for (int k = 0; k < 3; ++k) s += dR.at<double>(r, k) * R[3 * k + c]; Rn[3 * r + c] = s; }
      for (int k = 0; k < 3; ++k) tn[k] = t[k] + d[3 + k];
      const double cn = cost_at(Rn, tn);
      if (cn <= cost) {
        step_norm = 0; double pn = 0;
        for (int k = 0; k < 3; ++k) { step_norm += d[k] * d[k] + d[3 + k] * d[3 + k] / (1.0 + t[k] * t[k]); pn += 1.0; }
        for (int k = 0; k < 9; ++k) R[k] = Rn[k];
        for (int k = 0; k < 3; ++k) t[k] = tn[k];
        improved = true; cost = cn; lambda = std::max(lambda * 0.1, 1e-12);
        (void)pn;
      } else {
        lambda *= 10;
      }
    }
    if (!improved || step_norm < 1e-28) break;
  }
  Mat Rm(3, 3);
  for (int k = 0; k < 9; ++k) Rm.at<double>(k / 3, k % 3) = R[k];
  Rodrigues(Rm, rvec);
  tvec = Mat(3, 1);
  for (int k = 0; k < 3; ++k) tvec.at<double>(k, 0) = t[k];
  return true;
}

// cv::solvePnPRansac(objectPoints, imagePoints, cameraMatrix, distCoeffs, rvec, tvec) with
// OpenCV's defaults (SOLVEPNP_ITERATIVE, reprojectionError = 8.0) as the reference calls it
// (TS.cpp:193): planar board, normalised image points, identity camera matrix.  In those
// units the 8.0 threshold only removes corners whose back-projection failed (NaN: pixel
// outside the model's domain under the current guess, TS.h:47) or blew up near the
// horizon; OpenCV then returns solvePnP(ITERATIVE) on the inliers.  Restated as consensus
// re-fitting: fit the finite points, drop those beyond the threshold, re-fit until the set
// is stable.  The random minimal-sample stage of RANSAC is not restated (it only selects
// the inlier set).  distCoeffs must be empty.
inline bool solvePnPRansac(const std::vector<Point3d>& obj, const std::vector<Point2d>& img, const Mat& K,
                           const Mat& /*distCoeffs*/, Mat& rvec, Mat& tvec) {
  const size_t n = obj.size();
  if (img.size() != n) return false;
  const double kfx = K.at<double>(0, 0), kfy = K.at<double>(1, 1), kcx = K.at<double>(0, 2), kcy = K.at<double>(1, 2);
  std::vector<char> use(n);
  for (size_t i = 0; i < n; ++i) use[i] = std::isfinite(img[i].x) && std::isfinite(img[i].y);
  for (int round = 0; round < 8; ++round) {
    std::vector<Point3d> o;
    std::vector<Point2d> p;
    for (size_t i = 0; i < n; ++i) if (use[i]) { o.push_back(obj[i]); p.push_back(img[i]); }
    if (!solvePnPPlanarIterative(o, p, K, rvec, tvec)) return false;
    Mat R;
    Rodrigues(rvec, R);
    bool changed = false;
    for (size_t i = 0; i < n; ++i) {
      if (!use[i]) continue;
      double P[3];
      for (int r = 0; r < 3; ++r)
        P[r] = R.at<double>(r, 0) * obj[i].x + R.at<double>(r, 1) * obj[i].y + R.at<double>(r, 2) * obj[i].z + tvec.at<double>(r, 0);
      const double eu = kfx * P[0] / P[2] + kcx - img[i].x, ev = kfy * P[1] / P[2] + kcy - img[i].y;
      if (!(eu * eu + ev * ev <= 8.0 * 8.0)) { use[i] = 0; changed = true; }
    }
    if (!changed) return true;
  }
  return true;
}

}  // namespace cv
#endif  // TSCM_USE_OPENCV
