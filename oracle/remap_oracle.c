/*
 * remap_oracle.c — CPU restatement of the reference's remap-table loops.
 *
 * TEST INFRASTRUCTURE ONLY (see tscm_oracle.h): used by tests/ to check the CUDA kernel
 * k_remap_tables, never by the product path.
 *
 * Follows, operation by operation and in the same evaluation order,
 *   TripleSphereCamera::undistort             /root/reference/TS.cpp:284-306
 *   TripleSphereCamera::undistort_chessboard  /root/reference/TS.cpp:308-330 (mapx/mapy loops)
 *   TripleSphereCamera::project               /root/reference/TS.cpp:332-344
 *   TScamera::project (validity cut-off)      /root/reference/EpipolarRectify/rectify.cpp:22-36
 *   Remap::init_remap (one block per call)    /root/reference/EpipolarRectify/rectify.cpp:86-199
 * Compiled with -ffp-contract=off so that every operation is one IEEE double rounding
 * (std::pow(x, 2) is exactly x*x).  cv::Mat's 3x3 * 3x1 product is restated as a
 * left-to-right row sum.
 *
 * Pinning: tests/golden/remap_tables.npz holds tables produced by an independent numpy
 * float64 transcription of the same loops (tests/golden/make_golden_remap.py) on the
 * reference's own fixture EpipolarRectify/calib.yaml; this file must match them bit for bit.
 */
#include <math.h>
#include <stdint.h>

#include "../include/tscm.h"

static void project_ts(const double* in, double X, double Y, double Z, double w2, double* u, double* v) {
  const double fx = in[0], fy = in[1], cx = in[2], cy = in[3];
  const double xi = in[4], lamda = in[5], alpha = in[6], b = in[7], c = in[8];
  const double d1 = sqrt(X * X + Y * Y + Z * Z);
  if (w2 > 0.0 && Z <= -w2 * d1) { *u = -1.0; *v = -1.0; return; }       /* rectify.cpp:28 */
  const double t1 = Z + xi * d1;
  const double d2 = sqrt(X * X + Y * Y + t1 * t1);
  const double t2 = Z + xi * d1 + lamda * d2;
  const double d3 = sqrt(X * X + Y * Y + t2 * t2);
  const double ksai = Z + xi * d1 + lamda * d2 + alpha / (1 - alpha) * d3;
  *u = fx * X / ksai + b * Y / ksai + cx;
  *v = c * X / ksai + fy * Y / ksai + cy;
}

int tscm_oracle_remap_tables(const tscm_remap_job* jobs, int32_t num_jobs, int32_t map_width,
                             int32_t map_height, float* mapx, float* mapy) {
  for (int k = 0; k < num_jobs; ++k) {
    const tscm_remap_job* J = &jobs[k];
    if (J->row0 < 0 || J->col0 < 0 || J->row0 + J->height > map_height || J->col0 + J->width > map_width)
      return 1;
    const double* M = J->matrix;
    for (int i = 0; i < J->height; ++i)
      for (int j = 0; j < J->width; ++j) {
        const double x = (j - J->ray_cx) / J->ray_fx;
        const double y = (i - J->ray_cy) / J->ray_fy;
        const double z = 1.0;
        const double X = M[0] * x + M[1] * y + M[2] * z;
        const double Y = M[3] * x + M[4] * y + M[5] * z;
        const double Z = M[6] * x + M[7] * y + M[8] * z;
        double u, v;
        project_ts(J->intrinsics, X, Y, Z, J->cutoff_w2, &u, &v);
        const long o = (long)(J->row0 + i) * map_width + J->col0 + j;
        mapx[o] = (float)(J->offset_x != 0.0 ? u + J->offset_x : u);
        mapy[o] = (float)(J->offset_y != 0.0 ? v + J->offset_y : v);
      }
  }
  return 0;
}
