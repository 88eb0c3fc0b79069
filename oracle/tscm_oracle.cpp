// tscm_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY (see tscm_oracle.h).
//
// PARITY UNPINNED: there is no reference test, golden vector or runnable
// reference/Ceres build for this path (tscm_oracle.h explains).  What follows
// restates
//   * the reference's residual functors           TS.h:100-131, multi_calib.h:146-195
//   * its problem construction and solver options TS.cpp:251-274, multi_calib.cpp:162-212
//   * its accuracy read-out                       multi_calib.cpp:235-283
//   * ceres::AngleAxisRotatePoint, Jet autodiff, TrustRegionMinimizer,
//     LevenbergMarquardtStrategy, SchurEliminator, DenseSchurComplementSolver,
//     LossFunction/Corrector (Ceres 1.14-2.2 published behaviour; Ceres is a
//     third-party dependency that is not vendored in /root/reference and not
//     installed in this image).
//
// Build: see oracle/Makefile (g++ -O3 -march=x86-64-v3 -pthread).

#include "tscm_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>

#include <functional>
#include <thread>

namespace {

// Static-partition parallel for over [0, n): fn(thread_index, begin, end).
// (std::thread instead of OpenMP: the image's default $CXX has no libgomp spec.)
template <typename Fn>
void ParallelFor(int n, int num_threads, Fn fn) {
  num_threads = std::max(1, std::min(num_threads, n));
  if (num_threads == 1) { fn(0, 0, n); return; }
  std::vector<std::thread> pool;
  for (int t = 0; t < num_threads; ++t) {
    const int b = (int)((long long)n * t / num_threads), e = (int)((long long)n * (t + 1) / num_threads);
    pool.emplace_back([=]() { fn(t, b, e); });
  }
  for (auto& th : pool) th.join();
}

// ---------------------------------------------------------------------------
// Dual numbers (what ceres::Jet<double, N> is): value + N partials.
// ---------------------------------------------------------------------------
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0.0) { for (int i = 0; i < N; ++i) v[i] = 0.0; }
  explicit Jet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; }
  Jet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; v[k] = 1.0; }
};
template <int N> inline Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <int N> inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  // (f/g)' = (f' - (f/g) g') / g, evaluated the way jet.h does.
  Jet<N> h; const double gi = 1.0 / g.a; const double q = f.a * gi; h.a = q;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - q * g.v[i]) * gi; return h; }
template <int N> inline Jet<N>& operator+=(Jet<N>& f, const Jet<N>& g) { f = f + g; return f; }
template <int N> inline Jet<N> sqrt(const Jet<N>& f) {
  Jet<N> h; const double t = std::sqrt(f.a); const double k = 1.0 / (2.0 * t); h.a = t;
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * k; return h; }
template <int N> inline Jet<N> sin(const Jet<N>& f) {
  Jet<N> h; h.a = std::sin(f.a); const double c = std::cos(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i]; return h; }
template <int N> inline Jet<N> cos(const Jet<N>& f) {
  Jet<N> h; h.a = std::cos(f.a); const double s = -std::sin(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = s * f.v[i]; return h; }
template <int N> inline bool operator>(const Jet<N>& f, const Jet<N>& g) { return f.a > g.a; }
using std::sqrt; using std::sin; using std::cos;

// ---------------------------------------------------------------------------
// ceres::AngleAxisRotatePoint (ceres/rotation.h): Rodrigues' formula away from
// zero, first-order Taylor expansion R = I + hat(w) for theta^2 <= DBL_EPSILON.
// Call sites: TS.h:112, multi_calib.h:158,164.
// ---------------------------------------------------------------------------
template <typename T>
inline void AngleAxisRotatePoint(const T aa[3], const T pt[3], T out[3]) {
  const T theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (theta2 > T(std::numeric_limits<double>::epsilon())) {
    const T theta = sqrt(theta2);
    const T costheta = cos(theta);
    const T sintheta = sin(theta);
    const T theta_inverse = T(1.0) / theta;
    const T w[3] = {aa[0] * theta_inverse, aa[1] * theta_inverse, aa[2] * theta_inverse};
    const T w_cross_pt[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2],
                             w[0] * pt[1] - w[1] * pt[0]};
    const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (T(1.0) - costheta);
    out[0] = pt[0] * costheta + w_cross_pt[0] * sintheta + w[0] * tmp;
    out[1] = pt[1] * costheta + w_cross_pt[1] * sintheta + w[1] * tmp;
    out[2] = pt[2] * costheta + w_cross_pt[2] * sintheta + w[2] * tmp;
  } else {
    const T w_cross_pt[3] = {aa[1] * pt[2] - aa[2] * pt[1], aa[2] * pt[0] - aa[0] * pt[2],
                             aa[0] * pt[1] - aa[1] * pt[0]};
    out[0] = pt[0] + w_cross_pt[0];
    out[1] = pt[1] + w_cross_pt[1];
    out[2] = pt[2] + w_cross_pt[2];
  }
}

// ---------------------------------------------------------------------------
// TS projection tail shared by both functors: TS.h:117-129 == multi_calib.h:169-193.
// Operation order kept as written in the reference.
// ---------------------------------------------------------------------------
template <typename T>
inline void TsResidualFromCameraPoint(const T P[3], const T* intrinsic_, double obs_x,
                                      double obs_y, T* residuals) {
  const T one = T(1.0);
  const T d1 = sqrt(P[0] * P[0] + P[1] * P[1] + P[2] * P[2]);
  const T d2 = sqrt(P[0] * P[0] + P[1] * P[1] +
                    (P[2] + intrinsic_[4] * d1) * (P[2] + intrinsic_[4] * d1));
  const T d3 = sqrt(P[0] * P[0] + P[1] * P[1] +
                    (P[2] + intrinsic_[4] * d1 + intrinsic_[5] * d2) *
                        (P[2] + intrinsic_[4] * d1 + intrinsic_[5] * d2));
  const T ksai = P[2] + intrinsic_[4] * d1 + intrinsic_[5] * d2 +
                 intrinsic_[6] / (one - intrinsic_[6]) * d3;
  const T pixel_x = intrinsic_[0] * P[0] / ksai + intrinsic_[2];
  const T pixel_y = intrinsic_[1] * P[1] / ksai + intrinsic_[3];
  residuals[0] = T(obs_x) - pixel_x;
  residuals[1] = T(obs_y) - pixel_y;
}

// MultiCalib::ReprojectionError::operator()  multi_calib.h:146-195
template <typename T>
inline void RigResidual(const T* camera_rt_, const T* chessboard_rt_, const T* intrinsic_,
                        double board_x, double board_y, double obs_x, double obs_y,
                        T* residuals) {
  T chessboard_p[3] = {T(board_x), T(board_y), T(0.0)};          // :153-156 (z forced to 0)
  T world_p[3];
  AngleAxisRotatePoint(chessboard_rt_, chessboard_p, world_p);   // :158
  world_p[0] += chessboard_rt_[3];
  world_p[1] += chessboard_rt_[4];
  world_p[2] += chessboard_rt_[5];
  T camera_p[3];
  AngleAxisRotatePoint(camera_rt_, world_p, camera_p);           // :164
  camera_p[0] += camera_rt_[3];
  camera_p[1] += camera_rt_[4];
  camera_p[2] += camera_rt_[5];
  TsResidualFromCameraPoint(camera_p, intrinsic_, obs_x, obs_y, residuals);  // :169-193
}

// TripleSphereCamera::ReprojectionError::operator()  TS.h:100-131
template <typename T>
inline void MonoResidual(const T* intrinsic_, const T* rt_, double board_x, double board_y,
                         double obs_x, double obs_y, T* residuals) {
  T p[3] = {T(board_x), T(board_y), T(0.0)};                     // :107-109
  T P[3];
  AngleAxisRotatePoint(rt_, p, P);                               // :112
  P[0] += rt_[3];
  P[1] += rt_[4];
  P[2] += rt_[5];
  TsResidualFromCameraPoint(P, intrinsic_, obs_x, obs_y, residuals);  // :117-129
}

// ---------------------------------------------------------------------------
// ceres::LossFunction / Corrector (loss_function.cc, corrector.cc).  The
// reference uses NULL; Huber/Cauchy exist for BASELINE.json config 5.
// ---------------------------------------------------------------------------
inline void LossEvaluate(int type, double a, double s, double rho[3]) {
  const double b = a * a;
  if (type == TSCM_LOSS_HUBER) {
    if (s > b) {
      const double r = std::sqrt(s);
      rho[0] = 2.0 * a * r - b;
      rho[1] = std::max(std::numeric_limits<double>::min(), a / r);
      rho[2] = -rho[1] / (2.0 * s);
    } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
  } else if (type == TSCM_LOSS_CAUCHY) {
    const double c = 1.0 / b;
    const double sum = 1.0 + s * c;
    const double inv = 1.0 / sum;
    rho[0] = b * std::log(sum);
    rho[1] = std::max(std::numeric_limits<double>::min(), inv);
    rho[2] = -c * (inv * inv);
  } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
}

struct Corrector {
  double sqrt_rho1, residual_scaling, alpha_sq_norm;
  Corrector(double sq_norm, const double rho[3]) {
    sqrt_rho1 = std::sqrt(rho[1]);
    if (sq_norm == 0.0 || rho[2] <= 0.0) {
      residual_scaling = sqrt_rho1; alpha_sq_norm = 0.0; return;
    }
    const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
    const double alpha = 1.0 - std::sqrt(D);
    residual_scaling = sqrt_rho1 / (1.0 - alpha);
    alpha_sq_norm = alpha / sq_norm;
  }
};

// ---------------------------------------------------------------------------
// Small dense linear algebra (what Eigen does for Ceres here).
// ---------------------------------------------------------------------------
// In-place lower Cholesky of the n x n row-major SPD matrix A (leading dim ld).
// Returns false when a pivot is not positive (Eigen::LLT info() != Success).
inline bool CholeskyLower(double* A, int n, int ld) {
  for (int j = 0; j < n; ++j) {
    double d = A[j * ld + j];
    for (int k = 0; k < j; ++k) d -= A[j * ld + k] * A[j * ld + k];
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    A[j * ld + j] = d;
    const double inv = 1.0 / d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * ld + j];
      for (int k = 0; k < j; ++k) s -= A[i * ld + k] * A[j * ld + k];
      A[i * ld + j] = s * inv;
    }
  }
  return true;
}
inline void CholeskySolve(const double* L, int n, int ld, double* b) {
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= L[i * ld + k] * b[k];
    b[i] = s / L[i * ld + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= L[k * ld + i] * b[k];
    b[i] = s / L[i * ld + i];
  }
}
// InvertPSDMatrix<6>(assume_full_rank = true): LLT solve against the identity.
inline bool Invert6(const double* M, double* inv) {
  double L[36];
  std::memcpy(L, M, sizeof(L));
  if (!CholeskyLower(L, 6, 6)) {
    for (int i = 0; i < 36; ++i) inv[i] = std::numeric_limits<double>::quiet_NaN();
    return false;
  }
  for (int c = 0; c < 6; ++c) {
    double e[6] = {0, 0, 0, 0, 0, 0};
    e[c] = 1.0;
    CholeskySolve(L, 6, 6, e);
    for (int r = 0; r < 6; ++r) inv[r * 6 + c] = e[r];
  }
  return true;
}

// ---------------------------------------------------------------------------
// Problem layout: parameter vector x = [board poses (6 each, the Schur e-blocks)
// | per camera: rt(6) unless constant, intrinsic(9)] — Ceres' automatic Schur
// ordering for this problem (poses form the maximal independent set), with the
// constant block removed (multi_calib.cpp:186).
// ---------------------------------------------------------------------------
constexpr int kJ = 21;  // 6 + 6 + 9 columns of AutoDiffCostFunction<...,2,6,6,9>

struct Layout {
  int C = 0, F = 0, K = 0, V = 0, fixed = -1;
  int n_e = 0, n_c = 0, n = 0;
  std::vector<int> cam_off;     // offset of camera block inside the reduced system
  std::vector<int> cam_rt_sz;   // 6 or 0
  std::vector<std::vector<int>> frame_views;
  bool ok = false;
};

Layout MakeLayout(const tscm_problem* p) {
  Layout L;
  if (!p || p->num_cameras <= 0 || p->num_frames <= 0 || p->corners_per_board <= 0 ||
      p->num_views < 0 || !p->board_xy || (p->num_views > 0 && (!p->view_camera ||
      !p->view_frame || !p->obs_xy)))
    return L;
  L.C = p->num_cameras; L.F = p->num_frames; L.K = p->corners_per_board; L.V = p->num_views;
  L.fixed = p->fixed_camera;
  L.n_e = 6 * L.F;
  L.cam_off.resize(L.C); L.cam_rt_sz.resize(L.C);
  int off = 0;
  for (int c = 0; c < L.C; ++c) {
    L.cam_off[c] = off;
    L.cam_rt_sz[c] = (c == L.fixed) ? 0 : 6;
    off += L.cam_rt_sz[c] + 9;
  }
  L.n_c = off;
  L.n = L.n_e + L.n_c;
  L.frame_views.assign(L.F, {});
  for (int v = 0; v < L.V; ++v) {
    const int m = p->view_camera[v], i = p->view_frame[v];
    if (m < 0 || m >= L.C || i < 0 || i >= L.F) return L;
    L.frame_views[i].push_back(v);
  }
  L.ok = true;
  return L;
}

void PackX(const Layout& L, const double* intr, const double* cam_rt, const double* board_rt,
           std::vector<double>& x) {
  x.assign(L.n, 0.0);
  for (int i = 0; i < L.F; ++i) for (int k = 0; k < 6; ++k) x[6 * i + k] = board_rt[6 * i + k];
  for (int c = 0; c < L.C; ++c) {
    double* xc = &x[L.n_e + L.cam_off[c]];
    if (L.cam_rt_sz[c]) for (int k = 0; k < 6; ++k) xc[k] = cam_rt[6 * c + k];
    for (int k = 0; k < 9; ++k) xc[L.cam_rt_sz[c] + k] = intr[9 * c + k];
  }
}
void UnpackX(const Layout& L, const std::vector<double>& x, double* intr, double* cam_rt,
             double* board_rt) {
  for (int i = 0; i < L.F; ++i) for (int k = 0; k < 6; ++k) board_rt[6 * i + k] = x[6 * i + k];
  for (int c = 0; c < L.C; ++c) {
    const double* xc = &x[L.n_e + L.cam_off[c]];
    if (L.cam_rt_sz[c]) for (int k = 0; k < 6; ++k) cam_rt[6 * c + k] = xc[k];
    for (int k = 0; k < 9; ++k) intr[9 * c + k] = xc[L.cam_rt_sz[c] + k];
  }
}

struct Evaluator {
  const tscm_problem* p;
  const Layout& L;
  const double* fixed_cam_rt;  // constant block values (cam_rt of the fixed camera)
  int loss_type; double loss_scale;
  int num_threads;

  // Camera parameter pointers for camera c at x.
  void CameraParams(const std::vector<double>& x, int c, double rt[6], double intr[9]) const {
    const double* xc = &x[L.n_e + L.cam_off[c]];
    if (L.cam_rt_sz[c]) for (int k = 0; k < 6; ++k) rt[k] = xc[k];
    else for (int k = 0; k < 6; ++k) rt[k] = fixed_cam_rt[k];
    for (int k = 0; k < 9; ++k) intr[k] = xc[L.cam_rt_sz[c] + k];
  }

  // Residuals + (loss-corrected) Jacobian of one view; jac may be null.
  // Returns the view's cost contribution sum 1/2 rho(s).
  double EvalView(const std::vector<double>& x, int v, double* res, double* jac) const {
    const int m = p->view_camera[v], i = p->view_frame[v];
    double crt[6], cin[9];
    CameraParams(x, m, crt, cin);
    const double* brt = &x[6 * i];
    const double* obs = p->obs_xy + (size_t)v * L.K * 2;
    double cost = 0.0;
    if (jac) {
      Jet<kJ> jc[6], jb[6], ji[9];
      for (int k = 0; k < 6; ++k) jc[k] = Jet<kJ>(crt[k], k);
      for (int k = 0; k < 6; ++k) jb[k] = Jet<kJ>(brt[k], 6 + k);
      for (int k = 0; k < 9; ++k) ji[k] = Jet<kJ>(cin[k], 12 + k);
      for (int j = 0; j < L.K; ++j) {
        Jet<kJ> r[2];
        RigResidual(jc, jb, ji, p->board_xy[2 * j], p->board_xy[2 * j + 1], obs[2 * j],
                    obs[2 * j + 1], r);
        double* Jr = jac + (size_t)j * 2 * kJ;
        for (int a = 0; a < 2; ++a) {
          res[2 * j + a] = r[a].a;
          for (int k = 0; k < kJ; ++k) Jr[a * kJ + k] = r[a].v[k];
          // Constant parameter block: Ceres never forms these columns.
          if (!L.cam_rt_sz[m]) for (int k = 0; k < 6; ++k) Jr[a * kJ + k] = 0.0;
        }
        cost += ApplyLoss(res + 2 * j, Jr);
      }
    } else {
      for (int j = 0; j < L.K; ++j) {
        double r[2];
        RigResidual(crt, brt, cin, p->board_xy[2 * j], p->board_xy[2 * j + 1], obs[2 * j],
                    obs[2 * j + 1], r);
        if (res) { res[2 * j] = r[0]; res[2 * j + 1] = r[1]; }
        cost += ApplyLoss(r, nullptr);
      }
    }
    return cost;
  }

  // ResidualBlock::Evaluate: cost = 1/2 rho(s); correct Jacobian, then residuals.
  double ApplyLoss(double* r, double* Jr) const {
    const double s = r[0] * r[0] + r[1] * r[1];
    if (loss_type == TSCM_LOSS_NONE) return 0.5 * s;
    double rho[3];
    LossEvaluate(loss_type, loss_scale, s, rho);
    Corrector corr(s, rho);
    if (Jr) {
      if (corr.alpha_sq_norm == 0.0) {
        for (int k = 0; k < 2 * kJ; ++k) Jr[k] *= corr.sqrt_rho1;
      } else {
        for (int k = 0; k < kJ; ++k) {
          const double rtj = r[0] * Jr[k] + r[1] * Jr[kJ + k];
          Jr[k] = corr.sqrt_rho1 * (Jr[k] - corr.alpha_sq_norm * r[0] * rtj);
          Jr[kJ + k] = corr.sqrt_rho1 * (Jr[kJ + k] - corr.alpha_sq_norm * r[1] * rtj);
        }
      }
    }
    r[0] *= corr.residual_scaling;
    r[1] *= corr.residual_scaling;
    return 0.5 * rho[0];
  }

  // Evaluator::Evaluate(x, &cost, residuals, gradient, jacobian).
  void Evaluate(const std::vector<double>& x, double* cost, std::vector<double>* residuals,
                std::vector<double>* gradient, std::vector<double>* jacobian) const {
    const size_t per_view_r = (size_t)L.K * 2, per_view_j = (size_t)L.K * 2 * kJ;
    if (residuals) residuals->resize(per_view_r * L.V);
    if (jacobian) jacobian->resize(per_view_j * L.V);
    std::vector<double> view_cost(L.V, 0.0);
ParallelFor(L.V, num_threads, [&](int, int vb, int ve) {
      for (int v = vb; v < ve; ++v) {
        double* r = residuals ? residuals->data() + per_view_r * v : nullptr;
        double* J = jacobian ? jacobian->data() + per_view_j * v : nullptr;
        view_cost[v] = EvalView(x, v, r, J);
      }
    });
    // Summed in program order (frame-major after Ceres' Schur reordering).
    double c = 0.0;
    for (int i = 0; i < L.F; ++i) for (int v : L.frame_views[i]) c += view_cost[v];
    *cost = c;
    if (gradient && jacobian && residuals) {
      gradient->assign(L.n, 0.0);
      JtVec(*jacobian, *residuals, *gradient);
    }
  }

  // out += J^T vec (vec has 2N entries).  Frames are dealt to the threads (disjoint pose
  // segments of `out`); the camera part goes through per-thread accumulators that are
  // added in thread order.
  void JtVec(const std::vector<double>& jac, const std::vector<double>& vec,
             std::vector<double>& out) const {
    const int nt = std::max(1, num_threads);
    std::vector<std::vector<double>> cam(nt, std::vector<double>(L.n_c, 0.0));
    ParallelFor(L.F, nt, [&](int t, int fb, int fe) {
      double* cacc = cam[t].data();
      for (int i = fb; i < fe; ++i) {
        for (int v : L.frame_views[i]) {
          const int m = p->view_camera[v];
          const int coff = L.cam_off[m], rts = L.cam_rt_sz[m];
          for (int j = 0; j < L.K; ++j) {
            const double* Jr = &jac[((size_t)v * L.K + j) * 2 * kJ];
            for (int a = 0; a < 2; ++a) {
              const double rv = vec[((size_t)v * L.K + j) * 2 + a];
              const double* row = Jr + a * kJ;
              if (rts) for (int k = 0; k < 6; ++k) cacc[coff + k] += row[k] * rv;
              for (int k = 0; k < 6; ++k) out[6 * i + k] += row[6 + k] * rv;
              for (int k = 0; k < 9; ++k) cacc[coff + rts + k] += row[12 + k] * rv;
            }
          }
        }
      }
    });
    for (int t = 0; t < nt; ++t)
      for (int k = 0; k < L.n_c; ++k) out[L.n_e + k] += cam[t][k];
  }

  // SparseMatrix::SquaredColumnNorm
  void SquaredColumnNorm(const std::vector<double>& jac, std::vector<double>& out) const {
    out.assign(L.n, 0.0);
    const int nt = std::max(1, num_threads);
    std::vector<std::vector<double>> cam(nt, std::vector<double>(L.n_c, 0.0));
    ParallelFor(L.F, nt, [&](int t, int fb, int fe) {
      double* cacc = cam[t].data();
      for (int i = fb; i < fe; ++i) {
        for (int v : L.frame_views[i]) {
          const int m = p->view_camera[v];
          const int coff = L.cam_off[m], rts = L.cam_rt_sz[m];
          for (int j = 0; j < L.K; ++j) {
            const double* Jr = &jac[((size_t)v * L.K + j) * 2 * kJ];
            for (int a = 0; a < 2; ++a) {
              const double* row = Jr + a * kJ;
              if (rts) for (int k = 0; k < 6; ++k) cacc[coff + k] += row[k] * row[k];
              for (int k = 0; k < 6; ++k) out[6 * i + k] += row[6 + k] * row[6 + k];
              for (int k = 0; k < 9; ++k) cacc[coff + rts + k] += row[12 + k] * row[12 + k];
            }
          }
        }
      }
    });
    for (int t = 0; t < nt; ++t)
      for (int k = 0; k < L.n_c; ++k) out[L.n_e + k] += cam[t][k];
  }

  // SparseMatrix::ScaleColumns
  void ScaleColumns(std::vector<double>& jac, const std::vector<double>& scale) const {
ParallelFor(L.V, num_threads, [&](int, int vb, int ve) {
      for (int v = vb; v < ve; ++v) {
        const int m = p->view_camera[v], i = p->view_frame[v];
        const int coff = L.n_e + L.cam_off[m], rts = L.cam_rt_sz[m];
        for (int j = 0; j < L.K; ++j) {
          double* Jr = &jac[((size_t)v * L.K + j) * 2 * kJ];
          for (int a = 0; a < 2; ++a) {
            double* row = Jr + a * kJ;
            if (rts) for (int k = 0; k < 6; ++k) row[k] *= scale[coff + k];
            for (int k = 0; k < 6; ++k) row[6 + k] *= scale[6 * i + k];
            for (int k = 0; k < 9; ++k) row[12 + k] *= scale[coff + rts + k];
          }
        }
      }
    });
  }

  // model_residuals = J * step; returns -(m . (r + m/2))  (ComputeTrustRegionStep).
  double ModelCostChange(const std::vector<double>& jac, const std::vector<double>& res,
                         const std::vector<double>& step) const {
    const int nt = std::max(1, num_threads);
    std::vector<double> part(nt, 0.0);
    ParallelFor(L.F, nt, [&](int t, int fb, int fe) {
      double acc = 0.0;
      for (int i = fb; i < fe; ++i) {
        for (int v : L.frame_views[i]) {
          const int m = p->view_camera[v];
          const int coff = L.n_e + L.cam_off[m], rts = L.cam_rt_sz[m];
          for (int j = 0; j < L.K; ++j) {
            const double* Jr = &jac[((size_t)v * L.K + j) * 2 * kJ];
            for (int a = 0; a < 2; ++a) {
              const double* row = Jr + a * kJ;
              double mr = 0.0;
              if (rts) for (int k = 0; k < 6; ++k) mr += row[k] * step[coff + k];
              for (int k = 0; k < 6; ++k) mr += row[6 + k] * step[6 * i + k];
              for (int k = 0; k < 9; ++k) mr += row[12 + k] * step[coff + rts + k];
              const double r = res[((size_t)v * L.K + j) * 2 + a];
              acc += mr * (r + mr / 2.0);
            }
          }
        }
      }
      part[t] = acc;
    });
    double acc = 0.0;
    for (int t = 0; t < nt; ++t) acc += part[t];
    return -acc;
  }
};

// ---------------------------------------------------------------------------
// SchurEliminator::Eliminate + DenseSchurComplementSolver + BackSubstitute on
// the (column-scaled) Jacobian `jac`, right-hand side `res`, LM diagonal D.
// Solves (J^T J + D^T D) y = J^T res.  Returns false on LLT failure.
// If lhs_out/rhs_out are given, the reduced system is also copied out.
// ---------------------------------------------------------------------------
bool SchurSolve(const Evaluator& E, const std::vector<double>& jac, const std::vector<double>& res,
                const std::vector<double>& D, std::vector<double>& y, double* lhs_out,
                double* rhs_out) {
  const Layout& L = E.L;
  const tscm_problem* p = E.p;
  const int nc = L.n_c;
  int nthreads = std::max(1, E.num_threads);
  std::vector<std::vector<double>> lhs_t(nthreads), rhs_t(nthreads);
  for (int t = 0; t < nthreads; ++t) { lhs_t[t].assign((size_t)nc * nc, 0.0); rhs_t[t].assign(nc, 0.0); }
  std::vector<double> inv_ete_all((size_t)L.F * 36), g_all((size_t)L.F * 6);
  std::vector<int> ok_t(nthreads, 1);

  ParallelFor(L.F, nthreads, [&](int t, int fb, int fe) {
  for (int i = fb; i < fe; ++i) {
    double* lhs = lhs_t[t].data();
    double* rhs = rhs_t[t].data();
    const std::vector<int>& views = L.frame_views[i];
    double ete[36] = {0}, g[6] = {0};
    for (int k = 0; k < 6; ++k) ete[k * 6 + k] = D[6 * i + k] * D[6 * i + k];
    std::vector<double> buf(views.size() * 6 * 15, 0.0);  // E^T F per view (6 x fsz)
    for (size_t vi = 0; vi < views.size(); ++vi) {
      const int v = views[vi];
      const int m = p->view_camera[v];
      const int off = L.cam_off[m], rts = L.cam_rt_sz[m], fsz = rts + 9;
      double* EtF = &buf[vi * 90];
      for (int j = 0; j < L.K; ++j) {
        const double* Jr = &jac[((size_t)v * L.K + j) * 2 * kJ];
        for (int a = 0; a < 2; ++a) {
          const double* row = Jr + a * kJ;
          const double b = res[((size_t)v * L.K + j) * 2 + a];
          const double* e = row + 6;
          double f[15];
          for (int k = 0; k < rts; ++k) f[k] = row[k];
          for (int k = 0; k < 9; ++k) f[rts + k] = row[12 + k];
          for (int r = 0; r < 6; ++r) {
            g[r] += e[r] * b;
            for (int c = 0; c < 6; ++c) ete[r * 6 + c] += e[r] * e[c];
            for (int c = 0; c < fsz; ++c) EtF[r * 15 + c] += e[r] * f[c];
          }
          // F^T F on the block diagonal and F^T b.
          for (int r = 0; r < fsz; ++r) {
            rhs[off + r] += f[r] * b;
            for (int c = 0; c < fsz; ++c) lhs[(size_t)(off + r) * nc + off + c] += f[r] * f[c];
          }
        }
      }
    }
    double inv[36];
    if (!Invert6(ete, inv)) ok_t[t] = 0;
    std::memcpy(&inv_ete_all[(size_t)i * 36], inv, sizeof(inv));
    std::memcpy(&g_all[(size_t)i * 6], g, sizeof(g));
    double inv_g[6];
    for (int r = 0; r < 6; ++r) { double s = 0; for (int c = 0; c < 6; ++c) s += inv[r * 6 + c] * g[c]; inv_g[r] = s; }
    for (size_t vi = 0; vi < views.size(); ++vi) {
      const int m1 = p->view_camera[views[vi]];
      const int off1 = L.cam_off[m1], f1 = L.cam_rt_sz[m1] + 9;
      const double* B1 = &buf[vi * 90];
      // b1^T inv(ete)
      double BtI[15 * 6];
      for (int c = 0; c < f1; ++c) for (int r = 0; r < 6; ++r) {
        double s = 0; for (int k = 0; k < 6; ++k) s += B1[k * 15 + c] * inv[k * 6 + r];
        BtI[c * 6 + r] = s;
      }
      for (int c = 0; c < f1; ++c) {
        double s = 0; for (int k = 0; k < 6; ++k) s += B1[k * 15 + c] * inv_g[k];
        rhs[off1 + c] -= s;
      }
      for (size_t vj = 0; vj < views.size(); ++vj) {
        const int m2 = p->view_camera[views[vj]];
        const int off2 = L.cam_off[m2], f2 = L.cam_rt_sz[m2] + 9;
        const double* B2 = &buf[vj * 90];
        for (int r = 0; r < f1; ++r) for (int c = 0; c < f2; ++c) {
          double s = 0; for (int k = 0; k < 6; ++k) s += BtI[r * 6 + k] * B2[k * 15 + c];
          lhs[(size_t)(off1 + r) * nc + off2 + c] -= s;
        }
      }
    }
  }
  });
  bool ok_all = true;
  for (int t = 0; t < nthreads; ++t) ok_all = ok_all && ok_t[t];
  std::vector<double> lhs((size_t)nc * nc, 0.0), rhs(nc, 0.0);
  for (int t = 0; t < nthreads; ++t) {
    for (size_t k = 0; k < lhs.size(); ++k) lhs[k] += lhs_t[t][k];
    for (int k = 0; k < nc; ++k) rhs[k] += rhs_t[t][k];
  }
  for (int k = 0; k < nc; ++k) lhs[(size_t)k * nc + k] += D[L.n_e + k] * D[L.n_e + k];
  if (lhs_out) std::memcpy(lhs_out, lhs.data(), lhs.size() * sizeof(double));
  if (rhs_out) std::memcpy(rhs_out, rhs.data(), rhs.size() * sizeof(double));

  y.assign(L.n, 0.0);
  if (!ok_all) { for (double& t : y) t = std::numeric_limits<double>::quiet_NaN(); return true; }
  // DenseSchurComplementSolver::SolveReducedLinearSystem — Eigen LLT.
  if (!CholeskyLower(lhs.data(), nc, nc)) return false;
  CholeskySolve(lhs.data(), nc, nc, rhs.data());
  for (int k = 0; k < nc; ++k) y[L.n_e + k] = rhs[k];
  // BackSubstitute: y_e = inv(ete) (E^T b - E^T F y_f).
  ParallelFor(L.F, nthreads, [&](int, int fb, int fe) {
  for (int i = fb; i < fe; ++i) {
    double t6[6];
    for (int r = 0; r < 6; ++r) t6[r] = g_all[(size_t)i * 6 + r];
    for (int v : L.frame_views[i]) {
      const int m = p->view_camera[v];
      const int coff = L.n_e + L.cam_off[m], rts = L.cam_rt_sz[m];
      for (int j = 0; j < L.K; ++j) {
        const double* Jr = &jac[((size_t)v * L.K + j) * 2 * kJ];
        for (int a = 0; a < 2; ++a) {
          const double* row = Jr + a * kJ;
          double fy = 0.0;
          for (int k = 0; k < rts; ++k) fy += row[k] * y[coff + k];
          for (int k = 0; k < 9; ++k) fy += row[12 + k] * y[coff + rts + k];
          for (int r = 0; r < 6; ++r) t6[r] -= row[6 + r] * fy;
        }
      }
    }
    const double* inv = &inv_ete_all[(size_t)i * 36];
    for (int r = 0; r < 6; ++r) {
      double s = 0; for (int c = 0; c < 6; ++c) s += inv[r * 6 + c] * t6[c];
      y[6 * i + r] = s;
    }
  }
  });
  return true;
}

double Norm(const std::vector<double>& a) {
  double s = 0; for (double v : a) s += v * v; return std::sqrt(s);
}

void Trace(tscm_summary* s, int k, double cost, double radius, double gmax, double step_norm,
           int flags) {
  if (!s || k >= s->trace_capacity) return;
  if (s->trace_cost) s->trace_cost[k] = cost;
  if (s->trace_radius) s->trace_radius[k] = radius;
  if (s->trace_gradient_max_norm) s->trace_gradient_max_norm[k] = gmax;
  if (s->trace_step_norm) s->trace_step_norm[k] = step_norm;
  if (s->trace_step_flags) s->trace_step_flags[k] = flags;
}

const char* TerminationName(int t) {
  return t == TSCM_CONVERGENCE ? "CONVERGENCE" : t == TSCM_NO_CONVERGENCE ? "NO_CONVERGENCE" : "FAILURE";
}

}  // namespace

extern "C" {

int tscm_oracle_reduced_size(const tscm_problem* problem) {
  Layout L = MakeLayout(problem);
  return L.ok ? L.n_c : -1;
}

int tscm_oracle_eval_jacobian(const tscm_problem* problem, const double* intrinsics,
                              const double* cam_rt, const double* board_rt,
                              double* residuals, double* jacobian, double* cost) {
  Layout L = MakeLayout(problem);
  if (!L.ok) return TSCM_ERR_INVALID_ARGUMENT;
  std::vector<double> x;
  PackX(L, intrinsics, cam_rt, board_rt, x);
  const double zero6[6] = {0, 0, 0, 0, 0, 0};
  Evaluator E{problem, L, L.fixed >= 0 ? cam_rt + 6 * L.fixed : zero6, TSCM_LOSS_NONE, 1.0, 1};
  std::vector<double> r, J;
  double c;
  E.Evaluate(x, &c, &r, nullptr, &J);
  if (residuals) std::memcpy(residuals, r.data(), r.size() * sizeof(double));
  if (jacobian) std::memcpy(jacobian, J.data(), J.size() * sizeof(double));
  if (cost) *cost = c;
  return TSCM_OK;
}

int tscm_oracle_eval_jacobian_mono(const tscm_problem* problem, const double* intrinsics,
                                   const double* board_rt, double* residuals,
                                   double* jacobian, double* cost) {
  Layout L = MakeLayout(problem);
  if (!L.ok || L.C != 1) return TSCM_ERR_INVALID_ARGUMENT;
  constexpr int N = 15;  // AutoDiffCostFunction<...,2,9,6>  TS.cpp:261-264
  double c = 0.0;
  for (int v = 0; v < L.V; ++v) {
    const int i = problem->view_frame[v];
    Jet<N> ji[9], jr[6];
    for (int k = 0; k < 9; ++k) ji[k] = Jet<N>(intrinsics[k], k);
    for (int k = 0; k < 6; ++k) jr[k] = Jet<N>(board_rt[6 * i + k], 9 + k);
    const double* obs = problem->obs_xy + (size_t)v * L.K * 2;
    for (int j = 0; j < L.K; ++j) {
      Jet<N> r[2];
      MonoResidual(ji, jr, problem->board_xy[2 * j], problem->board_xy[2 * j + 1], obs[2 * j],
                   obs[2 * j + 1], r);
      const size_t o = (size_t)v * L.K + j;
      for (int a = 0; a < 2; ++a) {
        if (residuals) residuals[o * 2 + a] = r[a].a;
        if (jacobian) for (int k = 0; k < N; ++k) jacobian[(o * 2 + a) * N + k] = r[a].v[k];
        c += 0.5 * r[a].a * r[a].a;
      }
    }
  }
  if (cost) *cost = c;
  return TSCM_OK;
}

int tscm_oracle_reduced_system(const tscm_problem* problem, const tscm_options* options,
                               const double* intrinsics, const double* cam_rt,
                               const double* board_rt, double radius, double* lhs,
                               double* rhs) {
  Layout L = MakeLayout(problem);
  if (!L.ok || !options) return TSCM_ERR_INVALID_ARGUMENT;
  std::vector<double> x;
  PackX(L, intrinsics, cam_rt, board_rt, x);
  const double zero6[6] = {0, 0, 0, 0, 0, 0};
  Evaluator E{problem, L, L.fixed >= 0 ? cam_rt + 6 * L.fixed : zero6, options->loss_type,
              options->loss_scale, 1};
  std::vector<double> r, J, scale, diag, D(L.n), y;
  double c;
  E.Evaluate(x, &c, &r, nullptr, &J);
  E.SquaredColumnNorm(J, scale);
  for (int k = 0; k < L.n; ++k)
    scale[k] = options->jacobi_scaling ? 1.0 / (1.0 + std::sqrt(scale[k])) : 1.0;
  E.ScaleColumns(J, scale);
  E.SquaredColumnNorm(J, diag);
  for (int k = 0; k < L.n; ++k) {
    diag[k] = std::min(std::max(diag[k], options->min_lm_diagonal), options->max_lm_diagonal);
    D[k] = std::sqrt(diag[k] / radius);
  }
  SchurSolve(E, J, r, D, y, lhs, rhs);
  return TSCM_OK;
}

// TrustRegionMinimizer::Minimize with LevenbergMarquardtStrategy and
// DENSE_SCHUR, monotonic steps, no inner iterations, no bounds, no manifolds.
int tscm_oracle_solve(const tscm_problem* problem, const tscm_options* options,
                      double* intrinsics, double* cam_rt, double* board_rt,
                      tscm_summary* summary, int num_threads) {
  Layout L = MakeLayout(problem);
  if (!L.ok || !options || !intrinsics || !cam_rt || !board_rt)
    return TSCM_ERR_INVALID_ARGUMENT;
  const tscm_options& o = *options;
  const double zero6[6] = {0, 0, 0, 0, 0, 0};
  std::vector<double> fixed_rt(6, 0.0);
  if (L.fixed >= 0) for (int k = 0; k < 6; ++k) fixed_rt[k] = cam_rt[6 * L.fixed + k];
  (void)zero6;
  Evaluator E{problem, L, fixed_rt.data(), o.loss_type, o.loss_scale, std::max(1, num_threads)};

  std::vector<double> x, candidate_x, residuals, gradient, jacobian, scale(L.n, 1.0),
      diagonal(L.n), lm_diagonal(L.n), step(L.n), delta(L.n), tmp;
  PackX(L, intrinsics, cam_rt, board_rt, x);
  std::vector<double> best_x = x;

  double x_cost = 0, candidate_cost = 0, model_cost_change = 0;
  double radius = o.initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int num_consecutive_invalid_steps = 0;
  int termination = TSCM_NO_CONVERGENCE;
  int num_successful = 0, num_unsuccessful = 0, recorded = 0;
  double minimum_cost = std::numeric_limits<double>::max();
  double initial_cost = 0, final_cost = 0;
  bool atleast_one_successful_step = false;

  struct Iter { int iteration; bool valid, successful; double cost, gmax, gnorm, step_norm, radius; };
  Iter it{};

  // EvaluateGradientAndJacobian
  auto evaluate_gradient_and_jacobian = [&]() {
    E.Evaluate(x, &x_cost, &residuals, &gradient, &jacobian);
    it.cost = x_cost;
    if (o.jacobi_scaling) {
      if (it.iteration == 0) {
        E.SquaredColumnNorm(jacobian, scale);
        for (int k = 0; k < L.n; ++k) scale[k] = 1.0 / (1.0 + std::sqrt(scale[k]));
      }
      E.ScaleColumns(jacobian, scale);
    }
    // |Plus(x, -gradient) - x| with Plus(x, d) = x + d (no manifolds, no bounds).
    double gmax = 0, g2 = 0;
    for (int k = 0; k < L.n; ++k) {
      const double projected = x[k] + (-gradient[k]);
      const double d = x[k] - projected;
      gmax = std::max(gmax, std::fabs(d));
      g2 += d * d;
    }
    it.gmax = gmax;
    it.gnorm = std::sqrt(g2);
  };

  // IterationZero
  it.iteration = 0; it.valid = false; it.successful = false; it.step_norm = 0;
  evaluate_gradient_and_jacobian();
  initial_cost = x_cost;
  it.valid = true; it.successful = true;
  double x_norm = Norm(x);
  double reference_cost = x_cost;  // TrustRegionStepEvaluator (monotonic)

  bool returned = false;  // a `return` inside the loop body (no Finalize)
  for (;;) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (it.successful) {
      ++num_successful;
      if (x_cost < minimum_cost) { minimum_cost = x_cost; best_x = x; }
    } else {
      ++num_unsuccessful;
    }
    it.radius = radius;
    Trace(summary, recorded, it.cost, it.radius, it.gmax, it.step_norm,
          (it.valid ? 1 : 0) | (it.successful ? 2 : 0));
    final_cost = recorded == 0 ? it.cost : std::min(final_cost, it.cost);
    ++recorded;
    if (it.iteration >= o.max_num_iterations) { termination = TSCM_NO_CONVERGENCE; break; }
    if (!o.disable_tolerances) {
      if (it.successful && it.gmax <= o.gradient_tolerance) { termination = TSCM_CONVERGENCE; break; }
      if (!(it.radius > o.min_trust_region_radius)) { termination = TSCM_CONVERGENCE; break; }
    }

    const double previous_gmax = it.gmax, previous_gnorm = it.gnorm;
    const int next_iteration = it.iteration + 1;
    it = Iter{};
    it.iteration = next_iteration;

    // ComputeTrustRegionStep -> LevenbergMarquardtStrategy::ComputeStep
    if (!reuse_diagonal) {
      E.SquaredColumnNorm(jacobian, diagonal);
      for (int k = 0; k < L.n; ++k)
        diagonal[k] = std::min(std::max(diagonal[k], o.min_lm_diagonal), o.max_lm_diagonal);
    }
    for (int k = 0; k < L.n; ++k) lm_diagonal[k] = std::sqrt(diagonal[k] / radius);
    bool solver_ok = SchurSolve(E, jacobian, residuals, lm_diagonal, step, nullptr, nullptr);
    if (solver_ok) {
      for (int k = 0; k < L.n; ++k) if (!std::isfinite(step[k])) { solver_ok = false; break; }
    }
    if (solver_ok) for (int k = 0; k < L.n; ++k) step[k] = -step[k];
    reuse_diagonal = true;

    it.valid = false;
    if (solver_ok) {
      model_cost_change = E.ModelCostChange(jacobian, residuals, step);
      it.valid = model_cost_change > 0.0;
      if (it.valid) {
        for (int k = 0; k < L.n; ++k) delta[k] = step[k] * scale[k];
        num_consecutive_invalid_steps = 0;
      }
    }
    if (!it.valid) {
      // HandleInvalidStep
      if (++num_consecutive_invalid_steps >= o.max_num_consecutive_invalid_steps) {
        termination = TSCM_FAILURE; returned = true; break;
      }
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      it.cost = x_cost; it.gmax = previous_gmax; it.gnorm = previous_gnorm; it.step_norm = 0.0;
      it.successful = false;
      continue;
    }

    // ComputeCandidatePointAndEvaluateCost
    candidate_x.resize(L.n);
    for (int k = 0; k < L.n; ++k) candidate_x[k] = x[k] + delta[k];
    E.Evaluate(candidate_x, &candidate_cost, nullptr, nullptr, nullptr);
    if (!std::isfinite(candidate_cost)) candidate_cost = std::numeric_limits<double>::max();

    // ParameterToleranceReached
    {
      double s2 = 0;
      for (int k = 0; k < L.n; ++k) { const double d = x[k] - candidate_x[k]; s2 += d * d; }
      it.step_norm = std::sqrt(s2);
      const double tol = o.parameter_tolerance * (x_norm + o.parameter_tolerance);
      const bool armed = !o.parameter_tolerance_needs_successful_step || atleast_one_successful_step;
      if (!o.disable_tolerances && armed && it.step_norm <= tol) {
        termination = TSCM_CONVERGENCE; returned = true; break;
      }
    }
    // FunctionToleranceReached
    {
      const double cost_change = x_cost - candidate_cost;
      // Ceres >= 2.1 gates this test on atleast_one_successful_step as well
      const bool armed = !o.parameter_tolerance_needs_successful_step || atleast_one_successful_step;
      if (!o.disable_tolerances && armed && std::fabs(cost_change) <= o.function_tolerance * x_cost) {
        termination = TSCM_CONVERGENCE; returned = true; break;
      }
    }
    // IsStepSuccessful (TrustRegionStepEvaluator::StepQuality, monotonic)
    double relative_decrease;
    if (candidate_cost >= std::numeric_limits<double>::max()) {
      relative_decrease = std::numeric_limits<double>::lowest();
    } else {
      const double rd = (x_cost - candidate_cost) / model_cost_change;
      const double hrd = (reference_cost - candidate_cost) / model_cost_change;
      relative_decrease = std::max(rd, hrd);
    }
    if (relative_decrease > o.min_relative_decrease) {
      // HandleSuccessfulStep
      x = candidate_x;
      x_norm = Norm(x);
      evaluate_gradient_and_jacobian();
      it.valid = true; it.successful = true;
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * relative_decrease - 1.0, 3));
      radius = std::min(o.max_trust_region_radius, radius);
      decrease_factor = 2.0; reuse_diagonal = false;
      reference_cost = x_cost;
      atleast_one_successful_step = true;
    } else {
      it.successful = false;
      it.cost = candidate_cost;
      it.gmax = previous_gmax; it.gnorm = previous_gnorm;
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
    }
  }
  (void)returned;

  // Parameters are written back only from Finalize after successful steps.
  UnpackX(L, best_x, intrinsics, cam_rt, board_rt);
  if (summary) {
    summary->termination_type = termination;
    summary->num_iterations = recorded;
    summary->num_successful_steps = num_successful;
    summary->num_unsuccessful_steps = num_unsuccessful;
    summary->initial_cost = initial_cost;
    summary->final_cost = final_cost;
    summary->final_radius = radius;
  }
  if (o.verbose) {
    std::printf("Ceres Solver Report: Iterations: %d, Initial cost: %e, Final cost: %e, Termination: %s\n",
                num_successful + num_unsuccessful, initial_cost, final_cost,
                TerminationName(termination));
  }
  return TSCM_OK;
}

// cv::Rodrigues(rvec -> R).
static void RodriguesToMatrix(const double r[3], double R[9]) {
  const double theta = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  if (theta < DBL_EPSILON) {
    for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.0 : 0.0;
    return;
  }
  const double c = std::cos(theta), s = std::sin(theta), c1 = 1.0 - c, it = 1.0 / theta;
  const double x = r[0] * it, y = r[1] * it, z = r[2] * it;
  R[0] = c + c1 * x * x;     R[1] = c1 * x * y - s * z; R[2] = c1 * x * z + s * y;
  R[3] = c1 * x * y + s * z; R[4] = c + c1 * y * y;     R[5] = c1 * y * z - s * x;
  R[6] = c1 * x * z - s * y; R[7] = c1 * y * z + s * x; R[8] = c + c1 * z * z;
}

void tscm_oracle_project(const double* in, const double* pts, int n, double* uv) {
  // TripleSphereCamera::project  TS.cpp:332-344
  const double fx = in[0], fy = in[1], cx = in[2], cy = in[3], xi = in[4], lamda = in[5],
               alpha = in[6], b = in[7], c = in[8];
  for (int k = 0; k < n; ++k) {
    const double X = pts[3 * k], Y = pts[3 * k + 1], Z = pts[3 * k + 2];
    const double d1 = std::sqrt(X * X + Y * Y + Z * Z);
    const double d2 = std::sqrt(X * X + Y * Y + std::pow(Z + xi * d1, 2));
    const double d3 = std::sqrt(X * X + Y * Y + std::pow(Z + xi * d1 + lamda * d2, 2));
    const double ksai = Z + xi * d1 + lamda * d2 + alpha / (1 - alpha) * d3;
    uv[2 * k] = fx * X / ksai + b * Y / ksai + cx;
    uv[2 * k + 1] = c * X / ksai + fy * Y / ksai + cy;
  }
}

int tscm_oracle_reprojection_error(const tscm_problem* problem, const double* intrinsics,
                                   const double* cam_rt, const double* board_rt,
                                   double* per_camera, double* overall, double* rms) {
  // multi_calib.cpp:221-283: update_param() (Rodrigues) then mean Euclidean error.
  Layout L = MakeLayout(problem);
  if (!L.ok) return TSCM_ERR_INVALID_ARGUMENT;
  std::vector<double> err(L.C, 0.0), cnt(L.C, 0.0);
  double sum = 0, sq = 0; long total = 0;
  for (int v = 0; v < L.V; ++v) {
    const int m = problem->view_camera[v], i = problem->view_frame[v];
    double Rb[9], Rc[9];
    RodriguesToMatrix(board_rt + 6 * i, Rb);
    RodriguesToMatrix(cam_rt + 6 * m, Rc);
    const double* tb = board_rt + 6 * i + 3;
    const double* tc = cam_rt + 6 * m + 3;
    const double* obs = problem->obs_xy + (size_t)v * L.K * 2;
    for (int j = 0; j < L.K; ++j) {
      const double wx = problem->board_xy[2 * j], wy = problem->board_xy[2 * j + 1], wz = 0.0;
      double pw[3], pc[3], uv[2];
      for (int r = 0; r < 3; ++r) pw[r] = Rb[3 * r] * wx + Rb[3 * r + 1] * wy + Rb[3 * r + 2] * wz + tb[r];
      for (int r = 0; r < 3; ++r) pc[r] = Rc[3 * r] * pw[0] + Rc[3 * r + 1] * pw[1] + Rc[3 * r + 2] * pw[2] + tc[r];
      tscm_oracle_project(intrinsics + 9 * m, pc, 1, uv);
      const double dx = obs[2 * j] - uv[0], dy = obs[2 * j + 1] - uv[1];
      const double e = std::sqrt(dx * dx + dy * dy);
      err[m] += e; cnt[m] += 1.0; sum += e; sq += dx * dx + dy * dy; ++total;
    }
  }
  if (per_camera) for (int m = 0; m < L.C; ++m) per_camera[m] = cnt[m] > 0 ? err[m] / cnt[m] : 0.0;
  if (overall) *overall = total ? sum / total : 0.0;
  if (rms) *rms = total ? std::sqrt(sq / total) : 0.0;
  return TSCM_OK;
}

}  // extern "C"
