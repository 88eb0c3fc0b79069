// tscm_math.cuh — per-observation TS projection, residual and hand-derived
// analytic Jacobian, plus the small dense helpers of the Schur step.
//
// Everything here is __host__ __device__ so that tests/ can compile the very
// same arithmetic with g++ and compare it to the CPU oracle without a GPU.
// The shipped library only ever calls it from kernels (tscm_kernels.cu).
//
// Replaces the Jet<double,21> evaluation Ceres performs on
//   MultiCalib::ReprojectionError::operator()   /root/reference/multi_calib.h:146-195
//   TripleSphereCamera::ReprojectionError       /root/reference/TS.h:100-131
// (formula sheet: SURVEY.md Appendix B).
#pragma once

#include <cfloat>
#include <cmath>

#if defined(__CUDACC__)
#define TSCM_HD __host__ __device__ __forceinline__
#define TSCM_UNROLL _Pragma("unroll")
#else
#define TSCM_HD inline
#define TSCM_UNROLL
#endif

namespace tscm {

// Column order of the per-observation Jacobian row used on the device:
//   B 0..5   chessboard_rt  {w_f(3), t_f(3)}     (Schur e-block)
//   C 6..11  camera_rt      {w_c(3), t_c(3)}
//   I 12..18 intrinsic      {fx fy cx cy xi lambda alpha}  (b, c columns are
//            identically zero: TS.h:122-125 / multi_calib.h:175-178, dropped)
//   r 19     the residual itself (Gram of [J | r] gives J^T J, J^T r and r^T r)
constexpr int kNB = 6, kNC = 6, kNI = 7;
constexpr int kCols = kNB + kNC + kNI + 1;  // 20

// The 20x20 Gram matrix of [J | r] over a view's corners is kept in two records (doubles).
// Per-VIEW record (what the Schur elimination of the view's frame needs):
//   BB  0..20    6x6 upper triangle, row-major         -> V
//   BC  21..56   6x6  [b][c]                           -> W (camera_rt part)
//   BI  57..104  6x8  [b][i], i = 7 intrinsics then r  -> W (intrinsic part), g_e
//   105          padding
// Per-CAMERA record (summed over all views of a camera; partial sums per (CTA, camera)):
//   CC  0..20    6x6 upper triangle                    -> U
//   CI  21..68   6x8  [c][i]                           -> U, g_c
//   II  69..104  8x8 upper triangle                    -> U, g_c, r^T r
//   105          cost  sum 1/2 rho(s)
//   106          sum sqrt(s)  (mean Euclidean reprojection error read-out,
//                multi_calib.cpp:273; s taken BEFORE the loss correction)
constexpr int kOffBB = 0, kOffBC = 21, kOffBI = 57;
constexpr int kViewStride = 106;
constexpr int kCamCC = 0, kCamCI = 21, kCamII = 69, kCamCost = 105, kCamErr = 106;
constexpr int kCamRec = 107;

TSCM_HD constexpr int tri6(int i, int j) { return i * 6 - (i * (i - 1)) / 2 + (j - i); }   // i <= j
TSCM_HD constexpr int tri8(int i, int j) { return i * 8 - (i * (i - 1)) / 2 + (j - i); }

// Rotation matrix of an angle-axis vector and its partial derivatives,
// matching ceres::AngleAxisRotatePoint: Rodrigues' formula for theta^2 > eps,
// R = I + hat(w) (so dR/dw_b = hat(e_b)) otherwise.
// R is row-major; dR[b] = dR/dw_b row-major.
TSCM_HD void rotation_with_derivative(const double w[3], double R[9], double dR[3][9]) {
  const double theta2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  if (theta2 > DBL_EPSILON) {
    const double theta = sqrt(theta2);
    const double c = cos(theta), s = sin(theta);
    const double it = 1.0 / theta;
    const double sh = sin(0.5 * theta);
    const double c1 = 2.0 * sh * sh;  // 1 - cos(theta) without cancellation
    const double k[3] = {w[0] * it, w[1] * it, w[2] * it};
    // R = c I + s [k]x + c1 k k^T
    R[0] = c + c1 * k[0] * k[0];        R[1] = c1 * k[0] * k[1] - s * k[2]; R[2] = c1 * k[0] * k[2] + s * k[1];
    R[3] = c1 * k[0] * k[1] + s * k[2]; R[4] = c + c1 * k[1] * k[1];        R[5] = c1 * k[1] * k[2] - s * k[0];
    R[6] = c1 * k[0] * k[2] - s * k[1]; R[7] = c1 * k[1] * k[2] + s * k[0]; R[8] = c + c1 * k[2] * k[2];
    // d theta / d w_b = k_b ;  d k_i / d w_b = (delta_ib - k_i k_b) / theta
    TSCM_UNROLL
    for (int b = 0; b < 3; ++b) {
      double dk[3];
      TSCM_UNROLL
      for (int i = 0; i < 3; ++i) dk[i] = ((i == b ? 1.0 : 0.0) - k[i] * k[b]) * it;
      const double dc = -s * k[b];   // d cos
      const double ds = c * k[b];    // d sin
      const double dc1 = s * k[b];   // d (1 - cos)
      double* D = dR[b];
      // d(c I)
      D[0] = dc; D[4] = dc; D[8] = dc;
      D[1] = D[2] = D[3] = D[5] = D[6] = D[7] = 0.0;
      // d(s [k]x) = ds [k]x + s [dk]x
      D[1] += -(ds * k[2] + s * dk[2]); D[2] += (ds * k[1] + s * dk[1]);
      D[3] += (ds * k[2] + s * dk[2]);  D[5] += -(ds * k[0] + s * dk[0]);
      D[6] += -(ds * k[1] + s * dk[1]); D[7] += (ds * k[0] + s * dk[0]);
      // d(c1 k k^T) = dc1 k k^T + c1 (dk k^T + k dk^T)
      TSCM_UNROLL
      for (int i = 0; i < 3; ++i) {
        TSCM_UNROLL
        for (int j = 0; j < 3; ++j)
          D[3 * i + j] += dc1 * k[i] * k[j] + c1 * (dk[i] * k[j] + k[i] * dk[j]);
      }
    }
  } else {
    R[0] = 1.0;   R[1] = -w[2]; R[2] = w[1];
    R[3] = w[2];  R[4] = 1.0;   R[5] = -w[0];
    R[6] = -w[1]; R[7] = w[0];  R[8] = 1.0;
    TSCM_UNROLL
    for (int b = 0; b < 3; ++b) {
      TSCM_UNROLL
      for (int i = 0; i < 9; ++i) dR[b][i] = 0.0;
    }
    dR[0][5] = -1.0; dR[0][7] = 1.0;   // hat(e_x)
    dR[1][2] = 1.0;  dR[1][6] = -1.0;  // hat(e_y)
    dR[2][1] = -1.0; dR[2][3] = 1.0;   // hat(e_z)
  }
}

// Per-camera constants of one evaluation point (built once per evaluation).
struct CamConst {
  double R[9];       // reference -> camera rotation
  double t[3];
  double dR[3][9];   // dR/dw_b
  double fx, fy, cx, cy, xi, lam, k, dk;  // k = alpha/(1-alpha), dk = 1/(1-alpha)^2
  int free_rt;       // 0 for the constant block (multi_calib.cpp:186)
  int pad_;
};

TSCM_HD void make_cam_const(const double rt[6], const double intr[9], int free_rt, CamConst& c) {
  rotation_with_derivative(rt, c.R, c.dR);
  c.t[0] = rt[3]; c.t[1] = rt[4]; c.t[2] = rt[5];
  c.fx = intr[0]; c.fy = intr[1]; c.cx = intr[2]; c.cy = intr[3];
  c.xi = intr[4]; c.lam = intr[5];
  const double om = 1.0 - intr[6];
  c.k = intr[6] / om;
  c.dk = 1.0 / (om * om);
  c.free_rt = free_rt;
  c.pad_ = 0;
}

// Per-view constants (board pose of the view's frame).  The board point is
// (X, Y, 0), so only the first two columns of R_f and dR_f/dw_b are needed.
struct FrameConst {
  double r1[3], r2[3], t[3];   // R_f e_x, R_f e_y, t_f
  double d1[3][3], d2[3][3];   // d1[b] = dR_f/dw_b e_x, d2[b] = dR_f/dw_b e_y
};

TSCM_HD void make_frame_const(const double rt[6], FrameConst& f) {
  double R[9], dR[3][9];
  rotation_with_derivative(rt, R, dR);
  TSCM_UNROLL
  for (int i = 0; i < 3; ++i) {
    f.r1[i] = R[3 * i]; f.r2[i] = R[3 * i + 1]; f.t[i] = rt[3 + i];
    TSCM_UNROLL
    for (int b = 0; b < 3; ++b) { f.d1[b][i] = dR[b][3 * i]; f.d2[b][i] = dR[b][3 * i + 1]; }
  }
}

// One observation: residual rows and Jacobian rows in the column order above.
struct ObsRow {
  double Ju[kCols], Jv[kCols];   // [19] holds the residual
};

// Loss (ceres::HuberLoss / CauchyLoss + Corrector); type 0 = none.
TSCM_HD void loss_rho(int type, double a, double s, double rho[3]) {
  const double b = a * a;
  if (type == 1) {
    if (s > b) {
      const double r = sqrt(s);
      rho[0] = 2.0 * a * r - b;
      rho[1] = fmax(DBL_MIN, a / r);
      rho[2] = -rho[1] / (2.0 * s);
    } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
  } else if (type == 2) {
    const double c = 1.0 / b;
    const double sum = 1.0 + s * c;
    const double inv = 1.0 / sum;
    rho[0] = b * log(sum);
    rho[1] = fmax(DBL_MIN, inv);
    rho[2] = -c * (inv * inv);
  } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
}

// Forward TS projection of a camera-frame point.  Returns the intermediate
// quantities the Jacobian needs.
struct TsForward {
  double x, y, z, d1, d2, d3, id1, id2, id3, z1, z2, D, iD, mx, my;
};

// d = sqrt(s) and 1/d.  On the device one rsqrt (MUFU seed + Newton, <= 1 ulp) yields both
// (d = s * rsqrt(s), <= 2 ulp) instead of an IEEE sqrt followed by an IEEE division: the three
// sphere distances sit on the projection's critical dependent chain.
TSCM_HD void dist_and_inverse(double s, double& d, double& id) {
#if defined(__CUDA_ARCH__)
  id = rsqrt(s);
  d = s * id;
#else
  d = sqrt(s);
  id = 1.0 / d;
#endif
}

TSCM_HD void ts_forward(const CamConst& c, const double P[3], TsForward& f) {
  f.x = P[0]; f.y = P[1]; f.z = P[2];
  const double rho2 = f.x * f.x + f.y * f.y;
  dist_and_inverse(rho2 + f.z * f.z, f.d1, f.id1);
  f.z1 = f.z + c.xi * f.d1;
  dist_and_inverse(rho2 + f.z1 * f.z1, f.d2, f.id2);
  f.z2 = f.z1 + c.lam * f.d2;
  dist_and_inverse(rho2 + f.z2 * f.z2, f.d3, f.id3);
  f.D = f.z2 + c.k * f.d3;
  f.iD = 1.0 / f.D;
  f.mx = f.x * f.iD;
  f.my = f.y * f.iD;
}

// Camera-frame point of board corner (X, Y, 0) and the world point.
TSCM_HD void view_point(const CamConst& c, const FrameConst& f, double X, double Y,
                        double Pw[3], double P[3]) {
  TSCM_UNROLL
  for (int i = 0; i < 3; ++i) Pw[i] = X * f.r1[i] + Y * f.r2[i] + f.t[i];
  TSCM_UNROLL
  for (int i = 0; i < 3; ++i)
    P[i] = c.R[3 * i] * Pw[0] + c.R[3 * i + 1] * Pw[1] + c.R[3 * i + 2] * Pw[2] + c.t[i];
}

// Residual only (cost-only evaluation).
TSCM_HD void obs_residual(const CamConst& c, const FrameConst& f, double X, double Y,
                          double uo, double vo, double& ru, double& rv) {
  double Pw[3], P[3];
  view_point(c, f, X, Y, Pw, P);
  TsForward t;
  ts_forward(c, P, t);
  ru = uo - (c.fx * t.mx + c.cx);
  rv = vo - (c.fy * t.my + c.cy);
}

// Residual + analytic Jacobian.  NEED_* let a caller that only consumes some
// column groups drop the others at compile time.
template <bool NEED_B, bool NEED_C, bool NEED_I>
TSCM_HD void obs_jacobian(const CamConst& c, const FrameConst& f, double X, double Y,
                          double uo, double vo, ObsRow& o) {
  double Pw[3], P[3];
  view_point(c, f, X, Y, Pw, P);
  TsForward t;
  ts_forward(c, P, t);
  o.Ju[19] = uo - (c.fx * t.mx + c.cx);
  o.Jv[19] = vo - (c.fy * t.my + c.cy);

  // grad D = (e x, e y, h)  (radial symmetry of the three-sphere chain)
  const double id1 = t.id1, id2 = t.id2, id3 = t.id3;
  const double a1 = c.xi * id1;                       // grad z1 = (a1 x, a1 y, 1 + a1 z)
  const double g1 = 1.0 + a1 * t.z;
  const double b1 = (1.0 + t.z1 * a1) * id2;          // grad d2 = (b1 x, b1 y, c1)
  const double c1 = t.z1 * g1 * id2;
  const double a2 = a1 + c.lam * b1;                  // grad z2 = (a2 x, a2 y, g2)
  const double g2 = g1 + c.lam * c1;
  const double b2 = (1.0 + t.z2 * a2) * id3;          // grad d3 = (b2 x, b2 y, c2)
  const double c2 = t.z2 * g2 * id3;
  const double e = a2 + c.k * b2;
  const double h = g2 + c.k * c2;
  // A = d(u,v)/dP
  const double fu = c.fx * t.iD, fv = c.fy * t.iD;
  const double Au[3] = {fu * (1.0 - t.mx * e * t.x), -fu * t.mx * e * t.y, -fu * t.mx * h};
  const double Av[3] = {-fv * t.my * e * t.x, fv * (1.0 - t.my * e * t.y), -fv * t.my * h};

  if (NEED_C) {
    if (c.free_rt) {
      TSCM_UNROLL
      for (int b = 0; b < 3; ++b) {
        const double* D = c.dR[b];
        const double v0 = D[0] * Pw[0] + D[1] * Pw[1] + D[2] * Pw[2];
        const double v1 = D[3] * Pw[0] + D[4] * Pw[1] + D[5] * Pw[2];
        const double v2 = D[6] * Pw[0] + D[7] * Pw[1] + D[8] * Pw[2];
        o.Ju[6 + b] = -(Au[0] * v0 + Au[1] * v1 + Au[2] * v2);
        o.Jv[6 + b] = -(Av[0] * v0 + Av[1] * v1 + Av[2] * v2);
      }
      TSCM_UNROLL
      for (int i = 0; i < 3; ++i) { o.Ju[9 + i] = -Au[i]; o.Jv[9 + i] = -Av[i]; }
    } else {
      TSCM_UNROLL
      for (int i = 0; i < 6; ++i) { o.Ju[6 + i] = 0.0; o.Jv[6 + i] = 0.0; }
    }
  }
  if (NEED_B) {
    double ARu[3], ARv[3];   // A R_c
    TSCM_UNROLL
    for (int j = 0; j < 3; ++j) {
      ARu[j] = Au[0] * c.R[j] + Au[1] * c.R[3 + j] + Au[2] * c.R[6 + j];
      ARv[j] = Av[0] * c.R[j] + Av[1] * c.R[3 + j] + Av[2] * c.R[6 + j];
    }
    TSCM_UNROLL
    for (int b = 0; b < 3; ++b) {
      const double v0 = X * f.d1[b][0] + Y * f.d2[b][0];
      const double v1 = X * f.d1[b][1] + Y * f.d2[b][1];
      const double v2 = X * f.d1[b][2] + Y * f.d2[b][2];
      o.Ju[b] = -(ARu[0] * v0 + ARu[1] * v1 + ARu[2] * v2);
      o.Jv[b] = -(ARv[0] * v0 + ARv[1] * v1 + ARv[2] * v2);
    }
    TSCM_UNROLL
    for (int i = 0; i < 3; ++i) { o.Ju[3 + i] = -ARu[i]; o.Jv[3 + i] = -ARv[i]; }
  }
  if (NEED_I) {
    // dD/dxi, dD/dlambda, dD/dalpha
    const double d2xi = t.z1 * t.d1 * id2;
    const double z2xi = t.d1 + c.lam * d2xi;
    const double Dxi = z2xi + c.k * (t.z2 * z2xi * id3);
    const double Dlam = t.d2 + c.k * (t.z2 * t.d2 * id3);
    const double Dal = t.d3 * c.dk;
    const double ex = fu * t.mx, ey = fv * t.my;   // fx x / D^2, fy y / D^2
    o.Ju[12] = -t.mx; o.Jv[12] = 0.0;
    o.Ju[13] = 0.0;   o.Jv[13] = -t.my;
    o.Ju[14] = -1.0;  o.Jv[14] = 0.0;
    o.Ju[15] = 0.0;   o.Jv[15] = -1.0;
    o.Ju[16] = ex * Dxi;  o.Jv[16] = ey * Dxi;
    o.Ju[17] = ex * Dlam; o.Jv[17] = ey * Dlam;
    o.Ju[18] = ex * Dal;  o.Jv[18] = ey * Dal;
  }
}

// ---------------------------------------------------------------------------
// Moment formulation of the per-view normal-equation blocks.
//
// Camera-frame point of board corner (X, Y, 0):  P = X m1 + Y m2 + T  (per-view m1, m2, T).
// Every extrinsic column of dP/dparams is affine in (X, Y):
//      g_a(X, Y) = X c0[a] + Y c1[a] + c2[a]          a = 0..11  (w_f, t_f, w_c, t_c)
// and the residual Jacobian is  J = [ At G | J_I ],  At = -d(u,v)/dP  (2x3).  Hence
//      J_ext^T J_ext   = sum_{m,n} C_m^T  M[m][n]  C_n ,   M[m][n] = sum_j mu_m mu_n At_j^T At_j
//      J_ext^T [J_I|r] = sum_m     C_m^T  N[m]         ,   N[m]    = sum_j mu_m At_j^T [J_I | r]_j
// with mu = (X, Y, 1).  Per corner only At (6 numbers), the intrinsic rows and r are
// evaluated and 36 + 72 moment sums are updated; the 12x12 / 12x8 blocks are rebuilt once
// per view from the moments (view_blocks_column).
// ---------------------------------------------------------------------------
struct ViewConst {
  double m1[3], m2[3], T[3];
};

TSCM_HD void make_view_const(const CamConst& c, const FrameConst& f, ViewConst& v) {
  TSCM_UNROLL
  for (int i = 0; i < 3; ++i) {
    v.m1[i] = c.R[3 * i] * f.r1[0] + c.R[3 * i + 1] * f.r1[1] + c.R[3 * i + 2] * f.r1[2];
    v.m2[i] = c.R[3 * i] * f.r2[0] + c.R[3 * i + 1] * f.r2[1] + c.R[3 * i + 2] * f.r2[2];
    v.T[i] = c.R[3 * i] * f.t[0] + c.R[3 * i + 1] * f.t[1] + c.R[3 * i + 2] * f.t[2] + c.t[i];
  }
}

// Compact per-observation row: At rows (au, av), intrinsic rows + residual (ju, jv; [7] = r).
struct ObsCompact {
  double au[3], av[3];
  double ju[8], jv[8];
};

TSCM_HD void obs_compact(const CamConst& c, const ViewConst& vc, double X, double Y, double uo,
                         double vo, ObsCompact& o) {
  double P[3];
  TSCM_UNROLL
  for (int i = 0; i < 3; ++i) P[i] = X * vc.m1[i] + Y * vc.m2[i] + vc.T[i];
  TsForward t;
  ts_forward(c, P, t);
  o.ju[7] = uo - (c.fx * t.mx + c.cx);
  o.jv[7] = vo - (c.fy * t.my + c.cy);
  const double a1 = c.xi * t.id1;
  const double g1 = 1.0 + a1 * t.z;
  const double b1 = (1.0 + t.z1 * a1) * t.id2;
  const double c1 = t.z1 * g1 * t.id2;
  const double a2 = a1 + c.lam * b1;
  const double g2 = g1 + c.lam * c1;
  const double b2 = (1.0 + t.z2 * a2) * t.id3;
  const double c2 = t.z2 * g2 * t.id3;
  const double e = a2 + c.k * b2;
  const double h = g2 + c.k * c2;
  const double fu = c.fx * t.iD, fv = c.fy * t.iD;
  // At = -A
  o.au[0] = -fu * (1.0 - t.mx * e * t.x); o.au[1] = fu * t.mx * e * t.y; o.au[2] = fu * t.mx * h;
  o.av[0] = fv * t.my * e * t.x; o.av[1] = -fv * (1.0 - t.my * e * t.y); o.av[2] = fv * t.my * h;
  const double d2xi = t.z1 * t.d1 * t.id2;
  const double z2xi = t.d1 + c.lam * d2xi;
  const double Dxi = z2xi + c.k * (t.z2 * z2xi * t.id3);
  const double Dlam = t.d2 + c.k * (t.z2 * t.d2 * t.id3);
  const double Dal = t.d3 * c.dk;
  const double ex = fu * t.mx, ey = fv * t.my;
  o.ju[0] = -t.mx; o.jv[0] = 0.0;
  o.ju[1] = 0.0;   o.jv[1] = -t.my;
  o.ju[2] = -1.0;  o.jv[2] = 0.0;
  o.ju[3] = 0.0;   o.jv[3] = -1.0;
  o.ju[4] = ex * Dxi;  o.jv[4] = ey * Dxi;
  o.ju[5] = ex * Dlam; o.jv[5] = ey * Dlam;
  o.ju[6] = ex * Dal;  o.jv[6] = ey * Dal;
}

// Loss correction of a compact row (Huber / Cauchy have rho'' <= 0, so Ceres' Corrector
// reduces to a uniform scaling by sqrt(rho') of J and r — corrector.cc, alpha == 0 branch;
// structural zeros of the rows are preserved).  Returns 1/2 rho(s); *err = sqrt(s) (raw).
TSCM_HD double obs_compact_loss(int loss_type, double loss_scale, ObsCompact& o, double* err,
                                bool want_err) {
  const double ru = o.ju[7], rv = o.jv[7];
  const double s = ru * ru + rv * rv;
  *err = want_err ? sqrt(s) : 0.0;
  if (loss_type == 0) return 0.5 * s;
  double rho[3];
  loss_rho(loss_type, loss_scale, s, rho);
  const double w = sqrt(rho[1]);
  TSCM_UNROLL
  for (int k = 0; k < 3; ++k) { o.au[k] *= w; o.av[k] *= w; }
  TSCM_UNROLL
  for (int k = 0; k < 8; ++k) { o.ju[k] *= w; o.jv[k] *= w; }
  return 0.5 * rho[0];
}

// The three coefficient vectors of extrinsic column a (0..2 w_f, 3..5 t_f, 6..8 w_c, 9..11 t_c):
// out[m*3 + i] = c_m[a][i].
TSCM_HD void view_column_vectors(const CamConst& c, const FrameConst& f, int a, double out[9]) {
  TSCM_UNROLL
  for (int i = 0; i < 9; ++i) out[i] = 0.0;
  if (a < 3) {            // w_f[b]:  X R_c d1[b] + Y R_c d2[b]
    for (int i = 0; i < 3; ++i) {
      out[i] = c.R[3 * i] * f.d1[a][0] + c.R[3 * i + 1] * f.d1[a][1] + c.R[3 * i + 2] * f.d1[a][2];
      out[3 + i] = c.R[3 * i] * f.d2[a][0] + c.R[3 * i + 1] * f.d2[a][1] + c.R[3 * i + 2] * f.d2[a][2];
    }
  } else if (a < 6) {     // t_f[k]:  R_c[:, k]
    for (int i = 0; i < 3; ++i) out[6 + i] = c.R[3 * i + (a - 3)];
  } else if (c.free_rt) {
    if (a < 9) {          // w_c[b]:  dR_c[b] (X r1 + Y r2 + t_f)
      const double* D = c.dR[a - 6];
      for (int i = 0; i < 3; ++i) {
        out[i] = D[3 * i] * f.r1[0] + D[3 * i + 1] * f.r1[1] + D[3 * i + 2] * f.r1[2];
        out[3 + i] = D[3 * i] * f.r2[0] + D[3 * i + 1] * f.r2[1] + D[3 * i + 2] * f.r2[2];
        out[6 + i] = D[3 * i] * f.t[0] + D[3 * i + 1] * f.t[1] + D[3 * i + 2] * f.t[2];
      }
    } else {              // t_c[k]:  e_k
      out[6 + (a - 9)] = 1.0;
    }
  }
}

// Moment layouts.  M: 6 (m<=n pairs: 00 01 02 11 12 22) x 6 (sym 3x3: 00 01 02 11 12 22) = 36;
// N: 3 (m) x 3 (k) x 8 (i) = 72.
TSCM_HD constexpr int mom_pair(int m, int n) {   // m <= n
  return m == 0 ? n : (m == 1 ? 2 + n : 5);
}
TSCM_HD double sym3(const double* q, int i, int j) {   // q: 00 01 02 11 12 22
  const int a = i < j ? i : j, b = i < j ? j : i;
  return q[a == 0 ? b : (a == 1 ? 2 + b : 5)];
}

// Column b of the per-view blocks from the moments.  cols: [12][9] coefficient vectors of
// all columns (view_column_vectors), ld = stride between columns.
//   outE[a] = E[a][b] for a = 0..11    (use a <= b: upper triangle of J_ext^T J_ext)
//   outX[i] = X[b][i] for i = 0..7     (row b of J_ext^T [J_I | r])
//
// view_blocks_column_g is the same computation with the coefficient vectors behind an
// accessor col(a, k) (k = 0..8), for layouts that are not [12][ld] in memory (k_eval5 keeps
// them lane-interleaved in shared memory).
template <typename ColLoad, typename MomLoad>
TSCM_HD void view_blocks_column_g(ColLoad col, int b, MomLoad mom, double* outE, double* outX) {
  double cb[9];
  TSCM_UNROLL
  for (int k = 0; k < 9; ++k) cb[k] = col(b, k);
  double h[3][3];
  TSCM_UNROLL
  for (int m = 0; m < 3; ++m) {
    TSCM_UNROLL
    for (int i = 0; i < 3; ++i) h[m][i] = 0.0;
    TSCM_UNROLL
    for (int n = 0; n < 3; ++n) {
      const int pr = m <= n ? mom_pair(m, n) : mom_pair(n, m);
      double q[6];
      TSCM_UNROLL
      for (int e = 0; e < 6; ++e) q[e] = mom(pr * 6 + e);
      TSCM_UNROLL
      for (int i = 0; i < 3; ++i)
        h[m][i] += sym3(q, i, 0) * cb[3 * n] + sym3(q, i, 1) * cb[3 * n + 1] + sym3(q, i, 2) * cb[3 * n + 2];
    }
  }
  TSCM_UNROLL
  for (int a = 0; a < 12; ++a) {
    double s = 0.0;
    TSCM_UNROLL
    for (int m = 0; m < 3; ++m)
      s += col(a, 3 * m) * h[m][0] + col(a, 3 * m + 1) * h[m][1] + col(a, 3 * m + 2) * h[m][2];
    outE[a] = s;
  }
  TSCM_UNROLL
  for (int i = 0; i < 8; ++i) {
    double s = 0.0;
    TSCM_UNROLL
    for (int m = 0; m < 3; ++m)
      s += cb[3 * m] * mom(36 + (m * 3 + 0) * 8 + i) + cb[3 * m + 1] * mom(36 + (m * 3 + 1) * 8 + i) +
           cb[3 * m + 2] * mom(36 + (m * 3 + 2) * 8 + i);
    outX[i] = s;
  }
}

template <typename MomLoad>
TSCM_HD void view_blocks_column(const double* cols, int ld, int b, MomLoad mom, double* outE,
                                double* outX) {
  const double* cb = cols + b * ld;
  // h_m = sum_n M[m][n] c_n[b]
  double h[3][3];
  TSCM_UNROLL
  for (int m = 0; m < 3; ++m) {
    TSCM_UNROLL
    for (int i = 0; i < 3; ++i) h[m][i] = 0.0;
    TSCM_UNROLL
    for (int n = 0; n < 3; ++n) {
      const int pr = m <= n ? mom_pair(m, n) : mom_pair(n, m);
      double q[6];
      TSCM_UNROLL
      for (int e = 0; e < 6; ++e) q[e] = mom(pr * 6 + e);
      TSCM_UNROLL
      for (int i = 0; i < 3; ++i)
        h[m][i] += sym3(q, i, 0) * cb[3 * n] + sym3(q, i, 1) * cb[3 * n + 1] + sym3(q, i, 2) * cb[3 * n + 2];
    }
  }
  // all 12 rows are evaluated (fixed trip count: the 12 dot products interleave instead of
  // running as one dependent chain after another); the caller uses rows a <= b only
  TSCM_UNROLL
  for (int a = 0; a < 12; ++a) {
    const double* ca = cols + a * ld;
    double s = 0.0;
    TSCM_UNROLL
    for (int m = 0; m < 3; ++m)
      s += ca[3 * m] * h[m][0] + ca[3 * m + 1] * h[m][1] + ca[3 * m + 2] * h[m][2];
    outE[a] = s;
  }
  TSCM_UNROLL
  for (int i = 0; i < 8; ++i) {
    double s = 0.0;
    TSCM_UNROLL
    for (int m = 0; m < 3; ++m)
      s += cb[3 * m] * mom(36 + (m * 3 + 0) * 8 + i) + cb[3 * m + 1] * mom(36 + (m * 3 + 1) * 8 + i) +
           cb[3 * m + 2] * mom(36 + (m * 3 + 2) * 8 + i);
    outX[i] = s;
  }
}

// Loss correction of one observation's rows (ResidualBlock::Evaluate order:
// Jacobian first with the uncorrected residual, then the residual).  Returns
// 1/2 rho(s); *err receives sqrt(s) of the raw residual.
template <int LO, int HI>
TSCM_HD double obs_apply_loss(int loss_type, double loss_scale, ObsRow& o, double* err) {
  const double ru = o.Ju[19], rv = o.Jv[19];
  const double s = ru * ru + rv * rv;
  *err = sqrt(s);
  if (loss_type == 0) return 0.5 * s;
  double rho[3];
  loss_rho(loss_type, loss_scale, s, rho);
  const double sqrt_rho1 = sqrt(rho[1]);
  double residual_scaling, alpha_sq_norm;
  if (s == 0.0 || rho[2] <= 0.0) {
    residual_scaling = sqrt_rho1; alpha_sq_norm = 0.0;
  } else {
    const double D = 1.0 + 2.0 * s * rho[2] / rho[1];
    const double alpha = 1.0 - sqrt(D);
    residual_scaling = sqrt_rho1 / (1.0 - alpha);
    alpha_sq_norm = alpha / s;
  }
  if (alpha_sq_norm == 0.0) {
    TSCM_UNROLL
    for (int k = LO; k < HI; ++k) { o.Ju[k] *= sqrt_rho1; o.Jv[k] *= sqrt_rho1; }
  } else {
    TSCM_UNROLL
    for (int k = LO; k < HI; ++k) {
      const double rtj = ru * o.Ju[k] + rv * o.Jv[k];
      o.Ju[k] = sqrt_rho1 * (o.Ju[k] - alpha_sq_norm * ru * rtj);
      o.Jv[k] = sqrt_rho1 * (o.Jv[k] - alpha_sq_norm * rv * rtj);
    }
  }
  o.Ju[19] = ru * residual_scaling;
  o.Jv[19] = rv * residual_scaling;
  return 0.5 * rho[0];
}

// ---------------------------------------------------------------------------
// 6x6 SPD helpers for the Schur step (InvertPSDMatrix<6> of Ceres'
// SchurEliminator).  M is a full row-major 6x6; on return its lower triangle
// holds the Cholesky factor L with the RECIPROCAL of the diagonal stored on the
// diagonal (every later use multiplies instead of dividing).  A non-positive
// pivot produces NaN/Inf that propagate into the step, which the LM loop then
// rejects as an invalid step, as Ceres does for a failed linear solve.
// ---------------------------------------------------------------------------
TSCM_HD void chol6(double M[36]) {
  TSCM_UNROLL
  for (int j = 0; j < 6; ++j) {
    double d = M[j * 6 + j];
    TSCM_UNROLL
    for (int k = 0; k < j; ++k) d -= M[j * 6 + k] * M[j * 6 + k];
#if defined(__CUDA_ARCH__)
    const double inv = rsqrt(d);
#else
    const double inv = 1.0 / sqrt(d);
#endif
    M[j * 6 + j] = inv;
    TSCM_UNROLL
    for (int i = j + 1; i < 6; ++i) {
      double s = M[i * 6 + j];
      TSCM_UNROLL
      for (int k = 0; k < j; ++k) s -= M[i * 6 + k] * M[j * 6 + k];
      M[i * 6 + j] = s * inv;
    }
  }
}
// Solve L L^T x = b in place (L as produced by chol6).
TSCM_HD void chol6_solve(const double L[36], double b[6]) {
  TSCM_UNROLL
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
    TSCM_UNROLL
    for (int k = 0; k < i; ++k) s -= L[i * 6 + k] * b[k];
    b[i] = s * L[i * 6 + i];
  }
  TSCM_UNROLL
  for (int i = 5; i >= 0; --i) {
    double s = b[i];
    TSCM_UNROLL
    for (int k = i + 1; k < 6; ++k) s -= L[k * 6 + i] * b[k];
    b[i] = s * L[i * 6 + i];
  }
}

// chol6_solve with the factor stored packed (row i at i(i+1)/2, reciprocal diagonal), e.g. in
// shared memory: the 36-double register copy of the factor does not have to stay live.
TSCM_HD void chol6_solve_packed(const double* Lp, double b[6]) {
  TSCM_UNROLL
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
    TSCM_UNROLL
    for (int k = 0; k < i; ++k) s -= Lp[(i * (i + 1)) / 2 + k] * b[k];
    b[i] = s * Lp[(i * (i + 1)) / 2 + i];
  }
  TSCM_UNROLL
  for (int i = 5; i >= 0; --i) {
    double s = b[i];
    TSCM_UNROLL
    for (int k = i + 1; k < 6; ++k) s -= Lp[(k * (k + 1)) / 2 + i] * b[k];
    b[i] = s * Lp[(i * (i + 1)) / 2 + i];
  }
}

}  // namespace tscm
