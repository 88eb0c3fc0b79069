/*
 * tscm_oracle.h — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may link or call this.  The product path
 * (tscm_calib_b200/csrc, libtscm_b200.so) never does.
 *
 * PARITY UNPINNED: the reference (imuncle/TSCM_Calib) ships no tests, golden
 * vectors or known-answer fixtures for this path, and its numerical core is the
 * un-vendored, un-pinned third-party Ceres Solver (find_package(Ceres REQUIRED),
 * /root/reference/CMakeLists.txt:7), which together with Eigen and OpenCV is
 * absent from this image, so neither the reference nor Ceres can be run here.
 * This oracle restates (a) the reference's own residual functors and problem
 * construction, citing file:line, and (b) Ceres' published TRUST_REGION /
 * LEVENBERG_MARQUARDT / DENSE_SCHUR algorithm with the defaults the reference
 * leaves in force.  It is pinned only against an independent numpy/mpmath
 * restatement (oracle/numpy_ref.py, tests/golden/).
 */
#ifndef TSCM_ORACLE_H_
#define TSCM_ORACLE_H_

#include "../include/tscm.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Full Ceres-semantics solve.  num_threads = 1 is faithful to the reference
 * (Solver::Options::num_threads is never set: TS.cpp:271-274,
 * multi_calib.cpp:209-212); > 1 evaluates/eliminates frames with OpenMP. */
int tscm_oracle_solve(const tscm_problem* problem, const tscm_options* options,
                      double* intrinsics, double* cam_rt, double* board_rt,
                      tscm_summary* summary, int num_threads);

/* Residuals (N x 2) and autodiff Jacobian (N x 2 x 21; camera_rt 6,
 * chessboard_rt 6, intrinsic 9 — multi_calib.cpp:177-180) of the rig functor. */
int tscm_oracle_eval_jacobian(const tscm_problem* problem, const double* intrinsics,
                              const double* cam_rt, const double* board_rt,
                              double* residuals, double* jacobian, double* cost);

/* Mono functor (TS.h:100-131, AutoDiffCostFunction<...,2,9,6> TS.cpp:261-264):
 * residuals N x 2, Jacobian N x 2 x 15 (intrinsic 9, rt 6).  problem must have
 * num_cameras == 1. */
int tscm_oracle_eval_jacobian_mono(const tscm_problem* problem, const double* intrinsics,
                                   const double* board_rt, double* residuals,
                                   double* jacobian, double* cost);

/* Reduced camera system of one LM step at `radius`, Jacobi scaling computed at
 * this point (as at iteration 0).  Ceres layout: per camera [rt(6) unless
 * fixed][intrinsic(9)], n = 9C + 6(C - [fixed>=0]).  lhs row-major n x n. */
int tscm_oracle_reduced_size(const tscm_problem* problem);
int tscm_oracle_reduced_system(const tscm_problem* problem, const tscm_options* options,
                               const double* intrinsics, const double* cam_rt,
                               const double* board_rt, double radius, double* lhs,
                               double* rhs);

/* The reference's accuracy read-out, multi_calib.cpp:235-283. */
int tscm_oracle_reprojection_error(const tscm_problem* problem, const double* intrinsics,
                                   const double* cam_rt, const double* board_rt,
                                   double* per_camera, double* overall, double* rms);

/* Plain-double TS projection of camera-frame points (TS.cpp:332-344), skew
 * terms included.  pts n x 3 -> uv n x 2. */
void tscm_oracle_project(const double* intrinsic9, const double* pts, int n, double* uv);

#ifdef __cplusplus
}
#endif
#endif
