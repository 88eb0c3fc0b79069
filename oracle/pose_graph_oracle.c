/*
 * pose_graph_oracle.c — TEST INFRASTRUCTURE.  CPU restatement of MultiCalib's pose-graph
 * initialisation, /root/reference/multi_calib.cpp:6-153, in plain C (no FMA contraction:
 * compiled with -ffp-contract=off, so every double operation is one correctly rounded
 * instruction in the reference's evaluation order).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this; the
 * product path (tscm_pose_graph_init, csrc/tscm_posegraph.cuh) never does.
 *
 * Pinned against tests/golden/pose_graph.npz, a numpy transcription of the same constructor
 * written independently (tests/golden/make_golden_posegraph.py): same candidates chosen,
 * poses to 1e-12 / 1e-9 mm (tests/test_pose_graph.py).
 *
 * Conventions
 *   intr     [C][9]        {fx, fy, cx, cy, xi, lambda, alpha, b, c}       (TS.cpp:53-61)
 *   has      [C][B]        TripleSphereCamera::has_chessboard(j)
 *   mono_rt  [C][B][9]     TripleSphereCamera::Rt(j), row-major 3x3 = [r1 r2 t]  (TS.cpp:195-201)
 *   pixels   [C][B][K][2]  TripleSphereCamera::pixels()[j]
 *   worlds   [K][3]
 *   poses out: 12 doubles = R row-major (9) | t (3)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* multi_calib.h:130-137 — r1, r2 are narrowed to cv::Vec3f and r3 = r1 x r2 is a float product */
static void rt_to_R_t(const double* Rt, double* R, double* t) {
  const float r1[3] = {(float)Rt[0], (float)Rt[3], (float)Rt[6]};
  const float r2[3] = {(float)Rt[1], (float)Rt[4], (float)Rt[7]};
  const float r3[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2],
                       r1[0] * r2[1] - r1[1] * r2[0]};
  for (int k = 0; k < 3; ++k) {
    R[3 * k + 0] = r1[k];
    R[3 * k + 1] = r2[k];
    R[3 * k + 2] = r3[k];
  }
  t[0] = Rt[2];
  t[1] = Rt[5];
  t[2] = Rt[8];
}

/* cv::Mat products of 3x3 / 3x1 doubles: s = 0; s += a(i,k) * b(k,j), k ascending */
static void mat33(const double* a, const double* b, double* o) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += a[3 * i + k] * b[3 * k + j];
      o[3 * i + j] = s;
    }
}
static void mat33_bt(const double* a, const double* b, double* o) { /* a * b^T */
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += a[3 * i + k] * b[3 * j + k];
      o[3 * i + j] = s;
    }
}
static void mat33_at(const double* a, const double* b, double* o) { /* a^T * b */
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += a[3 * k + i] * b[3 * k + j];
      o[3 * i + j] = s;
    }
}
static void mat31(const double* a, const double* v, double* o) {
  for (int i = 0; i < 3; ++i) {
    double s = 0.0;
    for (int k = 0; k < 3; ++k) s += a[3 * i + k] * v[k];
    o[i] = s;
  }
}
static void mat31_at(const double* a, const double* v, double* o) { /* a^T * v */
  for (int i = 0; i < 3; ++i) {
    double s = 0.0;
    for (int k = 0; k < 3; ++k) s += a[3 * k + i] * v[k];
    o[i] = s;
  }
}

/* TS.h:58-69 with project() of TS.cpp:332-344: the SUM of the Euclidean pixel errors */
static double reproject_error(const double* in, const double* px, const double* worlds, int K,
                              const double* R, const double* t) {
  const double fx = in[0], fy = in[1], cx = in[2], cy = in[3], xi = in[4], lamda = in[5], alpha = in[6],
               b = in[7], c = in[8];
  double error = 0;
  for (int i = 0; i < K; ++i) {
    const double* w = worlds + 3 * i;
    double P[3];
    mat31(R, w, P);
    const double X = P[0] + t[0], Y = P[1] + t[1], Z = P[2] + t[2];
    const double d1 = sqrt(X * X + Y * Y + Z * Z);
    const double z1 = Z + xi * d1;
    const double d2 = sqrt(X * X + Y * Y + z1 * z1);
    const double z2 = Z + xi * d1 + lamda * d2;
    const double d3 = sqrt(X * X + Y * Y + z2 * z2);
    const double ksai = Z + xi * d1 + lamda * d2 + alpha / (1 - alpha) * d3;
    const double qx = fx * X / ksai + b * Y / ksai + cx;
    const double qy = c * X / ksai + fy * Y / ksai + cy;
    error += sqrt((px[2 * i] - qx) * (px[2 * i] - qx) + (px[2 * i + 1] - qy) * (px[2 * i + 1] - qy));
  }
  return error;
}

/* multi_calib.cpp:53-81: the score of one candidate pose (Rs | ts) of camera i — the summed
 * reprojection error of camera i-1 and of camera i over every board both detected. */
static double pair_candidate_error(int i, int B, int K, const double* worlds, const double* intr,
                                   const uint8_t* has, const double* mono_rt, const double* pixels,
                                   const double* camk, const double* cand) {
  const size_t PK = (size_t)K * 2;
  const double *Rk = camk, *tk = camk + 9, *Rs = cand, *ts = cand + 9;
  const uint8_t *ha = has + (size_t)(i - 1) * B, *hb = has + (size_t)i * B;
  double error = 0;
  for (int k = 0; k < B; ++k) {
    if (!ha[k] || !hb[k]) continue;
    double Ri[9], ti[3], Rp[9], tp[3], R_ki[9], t_ki[3], R_ik[9], t_ik[3], R[9], t[3], v[3];
    rt_to_R_t(mono_rt + ((size_t)i * B + k) * 9, Ri, ti);
    mat33_bt(Rk, Rs, R_ki);
    mat31(R_ki, ts, v);
    for (int r = 0; r < 3; ++r) t_ki[r] = tk[r] - v[r];
    mat33(R_ki, Ri, R);
    mat31(R_ki, ti, v);
    for (int r = 0; r < 3; ++r) t[r] = v[r] + t_ki[r];
    error += reproject_error(intr + 9 * (i - 1), pixels + ((size_t)(i - 1) * B + k) * PK, worlds, K, R, t);
    rt_to_R_t(mono_rt + ((size_t)(i - 1) * B + k) * 9, Rp, tp);
    mat33_bt(Rs, Rk, R_ik);
    mat31(R_ik, tk, v);
    for (int r = 0; r < 3; ++r) t_ik[r] = ts[r] - v[r];
    mat33(R_ik, Rp, R);
    mat31(R_ik, tp, v);
    for (int r = 0; r < 3; ++r) t[r] = v[r] + t_ik[r];
    error += reproject_error(intr + 9 * i, pixels + ((size_t)i * B + k) * PK, worlds, K, R, t);
  }
  return error;
}

/* multi_calib.cpp:36-47: the candidate pose of camera i chained through board j */
static void pair_candidate(int i, int j, int B, const double* mono_rt, const double* camk, double* cand) {
  double Ri[9], ti[3], Rp[9], tp[3], R_ik[9], t_ik[3], v[3];
  rt_to_R_t(mono_rt + ((size_t)i * B + j) * 9, Ri, ti);
  rt_to_R_t(mono_rt + ((size_t)(i - 1) * B + j) * 9, Rp, tp);
  mat33_bt(Ri, Rp, R_ik);
  mat31(R_ik, tp, v);
  for (int r = 0; r < 3; ++r) t_ik[r] = ti[r] - v[r];
  mat33(R_ik, camk, cand);
  mat31(R_ik, camk + 9, v);
  for (int r = 0; r < 3; ++r) cand[9 + r] = v[r] + t_ik[r];
}

/* Score of the single candidate of camera i built from board j, given the pose (R | t) already
 * chosen for camera i-1: for spot checks at sizes where the full n^2 loop takes minutes. */
double tscm_oracle_pose_pair_error(int i, int j, int B, int K, const double* worlds, const double* intr,
                                   const uint8_t* has, const double* mono_rt, const double* pixels,
                                   const double* camk_pose) {
  double cand[12];
  pair_candidate(i, j, B, mono_rt, camk_pose, cand);
  return pair_candidate_error(i, B, K, worlds, intr, has, mono_rt, pixels, camk_pose, cand);
}

/*
 * Returns 0, or 2 when two adjacent cameras share no board (the reference indexes Rs[-1] there,
 * multi_calib.cpp:51,86) or when no candidate scores below the 1e10 start value.
 * cand_err [C][B] (optional): summed error of the candidate built from board j, NaN elsewhere.
 * board_err [B][C] (optional): summed error of the candidate of camera j, NaN elsewhere.
 */
int tscm_oracle_pose_graph(int C, int B, int K, const double* worlds, const double* intr,
                           const uint8_t* has, const double* mono_rt, const double* pixels,
                           double* camera_pose, double* board_pose, uint8_t* board_init,
                           int32_t* camera_choice, int32_t* board_choice, double* cand_err,
                           double* board_err) {
  const size_t PK = (size_t)K * 2;
  if (cand_err) for (size_t k = 0; k < (size_t)C * B; ++k) cand_err[k] = NAN;
  if (board_err) for (size_t k = 0; k < (size_t)C * B; ++k) board_err[k] = NAN;
  double* cand = (double*)malloc((size_t)(B > C ? B : C) * 12 * sizeof(double));
  int* cand_board = (int*)malloc((size_t)(B > C ? B : C) * sizeof(int));
  int rc = 0;
  for (int i = 0; i < C && rc == 0; ++i) {
    double* Ri_out = camera_pose + 12 * i;
    if (i == 0) { /* multi_calib.cpp:18-23 */
      const double eye[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
      memcpy(Ri_out, eye, sizeof(eye));
      if (camera_choice) camera_choice[0] = -1;
      continue;
    }
    const double* Rk = camera_pose + 12 * (i - 1);
    const uint8_t *ha = has + (size_t)(i - 1) * B, *hb = has + (size_t)i * B;
    int n = 0;
    for (int j = 0; j < B; ++j) { /* multi_calib.cpp:29-48 */
      if (!ha[j] || !hb[j]) continue;
      pair_candidate(i, j, B, mono_rt, Rk, cand + 12 * n);
      cand_board[n++] = j;
    }
    double min_error = 1e10;
    int min_id = -1;
    for (int c = 0; c < n; ++c) { /* multi_calib.cpp:50-88 */
      const double error = pair_candidate_error(i, B, K, worlds, intr, has, mono_rt, pixels, Rk, cand + 12 * c);
      if (cand_err) cand_err[(size_t)i * B + cand_board[c]] = error;
      if (error < min_error) {
        min_error = error;
        min_id = c;
      }
    }
    if (min_id < 0) { rc = 2; break; }
    memcpy(Ri_out, cand + 12 * min_id, 12 * sizeof(double));
    if (camera_choice) camera_choice[i] = cand_board[min_id];
  }
  for (int i = 0; i < B && rc == 0; ++i) { /* multi_calib.cpp:95-152 */
    board_init[i] = 0;
    if (board_choice) board_choice[i] = -1;
    int n = 0;
    for (int j = 0; j < C; ++j) {
      if (!has[(size_t)j * B + i]) continue;
      double Rb[9], tb[3], d[3];
      const double *Rc = camera_pose + 12 * j, *tc = Rc + 9;
      rt_to_R_t(mono_rt + ((size_t)j * B + i) * 9, Rb, tb);
      mat33_at(Rc, Rb, cand + 12 * n);
      for (int r = 0; r < 3; ++r) d[r] = tb[r] - tc[r];
      mat31_at(Rc, d, cand + 12 * n + 9);
      cand_board[n++] = j;
    }
    if (n == 0) continue;
    int min_id = 0;
    if (n > 1) {
      double min_error = 1e10;
      min_id = -1;
      for (int c = 0; c < n; ++c) {
        const double *Rs = cand + 12 * c, *ts = Rs + 9;
        double error = 0;
        for (int k = 0; k < n; ++k) {
          const int m = cand_board[k];
          const double *Rc = camera_pose + 12 * m, *tc = Rc + 9;
          double R[9], t[3], v[3];
          mat33(Rc, Rs, R);
          mat31(Rc, ts, v);
          for (int r = 0; r < 3; ++r) t[r] = v[r] + tc[r];
          error += reproject_error(intr + 9 * m, pixels + ((size_t)m * B + i) * PK, worlds, K, R, t);
        }
        if (board_err) board_err[(size_t)i * C + cand_board[c]] = error;
        if (error < min_error) {
          min_error = error;
          min_id = c;
        }
      }
      if (min_id < 0) { rc = 2; break; }
    }
    memcpy(board_pose + 12 * i, cand + 12 * min_id, 12 * sizeof(double));
    board_init[i] = 1;
    if (board_choice) board_choice[i] = cand_board[min_id];
  }
  free(cand);
  free(cand_board);
  return rc;
}
