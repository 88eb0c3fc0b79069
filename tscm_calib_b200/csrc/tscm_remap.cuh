// tscm_remap.cuh — per-pixel remap tables through the Triple-Sphere projection
// (SURVEY.md §8f #4).
//
// Replaces the nested pixel loops of
//   TripleSphereCamera::undistort            /root/reference/TS.cpp:284-306
//   TripleSphereCamera::undistort_chessboard /root/reference/TS.cpp:308-330 (table part)
//   Remap::init_remap                        /root/reference/EpipolarRectify/rectify.cpp:86-199
// which are all the same operation: ray of output pixel (i, j) -> 3x3 matrix -> TS
// projection with skew (TS.cpp:332-344 / rectify.cpp:22-36) -> + mosaic offset -> float.
//
// Byte/bit contract: the tables are CV_32FC1; every double operation is issued as a single
// correctly-rounded instruction in the reference's evaluation order (__dmul_rn / __dadd_rn
// keep ptxas from contracting them into FMAs), so the stored floats are BIT-IDENTICAL to the
// scalar loops compiled without contraction (oracle/remap_oracle.c).
//
// One thread per output pixel, lanes along a row: every warp stores two 128-byte rows
// (mapx, mapy).  8 B written per pixel, nothing read but the job record: the kernel is bound
// by FP64 sqrt/div issue, not HBM (3 sqrt + 4 div per pixel with zero skew).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tscm.h"

namespace tscm {

// A job in device form (tscm_remap_job with the prefix sum of pixels before it).
struct RemapJobDev {
  double intr[9];
  double M[9];
  double ray_fx, ray_fy, ray_cx, ray_cy;
  double offset_x, offset_y;
  double cutoff_w2;
  int32_t width, height, row0, col0;
  int64_t first_pixel;   // pixels of all earlier jobs
};

constexpr int kRemapMaxJobs = 64;

struct RemapBatch {
  RemapJobDev job[kRemapMaxJobs];
  int32_t num_jobs;
  int32_t map_width;
  int64_t total_pixels;
};

__global__ void __launch_bounds__(256)
k_remap_tables(const RemapBatch* __restrict__ batch, float* __restrict__ mapx, float* __restrict__ mapy) {
  __shared__ int64_t s_first[kRemapMaxJobs + 1];
  const int nj = batch->num_jobs;
  for (int k = threadIdx.x; k <= nj; k += blockDim.x)
    s_first[k] = k < nj ? batch->job[k].first_pixel : batch->total_pixels;
  __syncthreads();
  const int64_t total = batch->total_pixels;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int jb = 0, ratio_job = -1;
  double ratio = 0.0;                                // alpha/(1-alpha) of the current job
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += stride) {
    while (p >= s_first[jb + 1]) ++jb;              // jobs are visited in increasing order
    const RemapJobDev& J = batch->job[jb];
    const int64_t q = p - s_first[jb];
    const int i = (int)(q / J.width), j = (int)(q - (int64_t)i * J.width);
    // ray of the output pixel: ((j-cx)/fx, (i-cy)/fy, 1)   TS.cpp:293-295, rectify.cpp:98
    const double x = __ddiv_rn(__dsub_rn((double)j, J.ray_cx), J.ray_fx);
    const double y = __ddiv_rn(__dsub_rn((double)i, J.ray_cy), J.ray_fy);
    // P = M * (x, y, 1), rows summed left to right
    const double X = __dadd_rn(__dadd_rn(__dmul_rn(J.M[0], x), __dmul_rn(J.M[1], y)), J.M[2]);
    const double Y = __dadd_rn(__dadd_rn(__dmul_rn(J.M[3], x), __dmul_rn(J.M[4], y)), J.M[5]);
    const double Z = __dadd_rn(__dadd_rn(__dmul_rn(J.M[6], x), __dmul_rn(J.M[7], y)), J.M[8]);
    // TS projection, TS.cpp:334-342
    const double fx = J.intr[0], fy = J.intr[1], cx = J.intr[2], cy = J.intr[3];
    const double xi = J.intr[4], lamda = J.intr[5], alpha = J.intr[6], b = J.intr[7], c = J.intr[8];
    const double r2 = __dadd_rn(__dmul_rn(X, X), __dmul_rn(Y, Y));
    const double d1 = __dsqrt_rn(__dadd_rn(r2, __dmul_rn(Z, Z)));
    double px, py;
    if (J.cutoff_w2 > 0.0 && Z <= __dmul_rn(-J.cutoff_w2, d1)) {   // rectify.cpp:28
      px = -1.0; py = -1.0;
    } else {
      const double z1 = __dadd_rn(Z, __dmul_rn(xi, d1));
      const double d2 = __dsqrt_rn(__dadd_rn(r2, __dmul_rn(z1, z1)));
      const double z2 = __dadd_rn(z1, __dmul_rn(lamda, d2));
      const double d3 = __dsqrt_rn(__dadd_rn(r2, __dmul_rn(z2, z2)));
      if (jb != ratio_job) { ratio = __ddiv_rn(alpha, __dsub_rn(1.0, alpha)); ratio_job = jb; }
      const double ksai = __dadd_rn(z2, __dmul_rn(ratio, d3));
      // the skew terms b*Y/ksai, c*X/ksai: with b = c = 0 (every calibration the reference
      // writes: TS.h:122-125 never moves them) the quotient is a signed zero, which a
      // multiplication produces bit-identically for finite non-zero ksai — 2 of the 4
      // FP64 divisions by ksai disappear
      const bool plain = ksai != 0.0 && isfinite(ksai);
      const double bY = __dmul_rn(b, Y), cX = __dmul_rn(c, X);
      const double su = (bY == 0.0 && plain) ? __dmul_rn(bY, ksai) : __ddiv_rn(bY, ksai);
      const double sv = (cX == 0.0 && plain) ? __dmul_rn(cX, ksai) : __ddiv_rn(cX, ksai);
      px = __dadd_rn(__dadd_rn(__ddiv_rn(__dmul_rn(fx, X), ksai), su), cx);
      py = __dadd_rn(__dadd_rn(sv, __ddiv_rn(__dmul_rn(fy, Y), ksai)), cy);
    }
    const size_t o = (size_t)(J.row0 + i) * batch->map_width + J.col0 + j;
    // mosaic offsets (rectify.cpp:113-114 etc.); no add at all where the reference has none
    mapx[o] = (float)(J.offset_x != 0.0 ? __dadd_rn(px, J.offset_x) : px);
    mapy[o] = (float)(J.offset_y != 0.0 ? __dadd_rn(py, J.offset_y) : py);
  }
}

}  // namespace tscm
