// tscm_posegraph.cuh — candidate scoring of MultiCalib's pose-graph initialisation
// (SURVEY.md §8f #2).
//
// Replaces the two exhaustive scoring loops of the reference constructor
//   camera chain   /root/reference/multi_calib.cpp:50-88    every candidate pose of camera i (one
//                  per board it shares with camera i-1) is scored by the summed reprojection
//                  error of BOTH cameras over ALL shared boards: n^2 x 2 x K projections
//   board poses    /root/reference/multi_calib.cpp:128-149  every candidate (one per camera that
//                  sees the board) scored over all cameras that see it
// with TripleSphereCamera::ReprojectError (TS.h:58-69) / project (TS.cpp:332-344) inside.
// At BASELINE config 3 (5,000 shared boards per adjacent pair) the chain is 7 x 4.4e9 TS
// projections: minutes of scalar host code, tens of milliseconds per pair here.
//
// Bit contract: every double operation is issued as ONE correctly rounded instruction in the
// reference's evaluation order (__dmul_rn / __dadd_rn keep ptxas from contracting to FMAs;
// sqrt and division are IEEE), one thread owns one (candidate, board, camera) error so its K
// corner errors are added in corner order, and k_pg_sum adds the per-board errors of a
// candidate in board order — the candidate errors are therefore BIT-IDENTICAL to a scalar
// loop compiled without contraction (oracle/pose_graph_oracle.c), and so is the arg-min.
//
// k_pg_pair_score: CTA = (128 candidates) x (a tile of shared boards).  The tile's base poses
//   and pixels are staged in shared memory once and read as warp broadcasts; thread =
//   candidate, its two relative transforms live in registers.  Bound: FP64 issue (4 sqrt + 2
//   div per projection).  The (board, side) x candidate error matrix goes to HBM with
//   coalesced 256-byte rows and is read once by k_pg_sum.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace tscm {

constexpr int kPgThreads = 128;
#ifndef PG_MIN_CTAS
#define PG_MIN_CTAS 6      // CTAs per SM the register allocation of k_pg_pair_score aims at (4: 281 ms, 5: 261, 6: 247 at config 3)
#endif

// out = a * b for 3x3 row-major doubles: s = 0; s += a(i,k) * b(k,j), k ascending (the cv::Mat
// product of the reference; the leading 0 + keeps the sign of a zero sum identical)
__device__ __forceinline__ void pg_mat33(const double* a, const double* b, double* o) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      o[3 * i + j] = __dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(a[3 * i], b[j])), __dmul_rn(a[3 * i + 1], b[3 + j])),
                               __dmul_rn(a[3 * i + 2], b[6 + j]));
}
__device__ __forceinline__ void pg_mat31(const double* a, const double* v, double* o) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
    o[i] = __dadd_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(a[3 * i], v[0])), __dmul_rn(a[3 * i + 1], v[1])),
                     __dmul_rn(a[3 * i + 2], v[2]));
}
// pose (R | t) = T o base:  R = T.R * base.R,  t = T.R * base.t + T.t
__device__ __forceinline__ void pg_compose(const double* T, const double* base, double* pose) {
  pg_mat33(T, base, pose);
  double v[3];
  pg_mat31(T, base + 9, v);
#pragma unroll
  for (int r = 0; r < 3; ++r) pose[9 + r] = __dadd_rn(v[r], T[9 + r]);
}

struct PgIntr {
  double fx, fy, cx, cy, xi, lamda, ratio, b, c;   // ratio = alpha / (1 - alpha)
};
__host__ __device__ inline PgIntr pg_intr(const double* in) {
  PgIntr I;
  I.fx = in[0]; I.fy = in[1]; I.cx = in[2]; I.cy = in[3]; I.xi = in[4]; I.lamda = in[5];
  I.ratio = in[6] / (1 - in[6]);
  I.b = in[7]; I.c = in[8];
  return I;
}

// One corner: P = R w + t (TS.h:63-64), project (TS.cpp:332-344), Euclidean pixel error.
__device__ __forceinline__ double pg_corner_error(const PgIntr& I, const double* pose, const double* w,
                                                  double ox, double oy) {
  double P[3];
  pg_mat31(pose, w, P);
  const double X = __dadd_rn(P[0], pose[9]), Y = __dadd_rn(P[1], pose[10]), Z = __dadd_rn(P[2], pose[11]);
  const double r2 = __dadd_rn(__dmul_rn(X, X), __dmul_rn(Y, Y));
  const double d1 = __dsqrt_rn(__dadd_rn(r2, __dmul_rn(Z, Z)));
  const double z1 = __dadd_rn(Z, __dmul_rn(I.xi, d1));
  const double d2 = __dsqrt_rn(__dadd_rn(r2, __dmul_rn(z1, z1)));
  const double z2 = __dadd_rn(z1, __dmul_rn(I.lamda, d2));
  const double d3 = __dsqrt_rn(__dadd_rn(r2, __dmul_rn(z2, z2)));
  const double ksai = __dadd_rn(z2, __dmul_rn(I.ratio, d3));
  // zero skew terms as in k_remap_tables: 0 / ksai is the signed zero 0 * ksai for finite ksai != 0
  const bool plain = ksai != 0.0 && isfinite(ksai);
  const double bY = __dmul_rn(I.b, Y), cX = __dmul_rn(I.c, X);
  const double su = (bY == 0.0 && plain) ? __dmul_rn(bY, ksai) : __ddiv_rn(bY, ksai);
  const double sv = (cX == 0.0 && plain) ? __dmul_rn(cX, ksai) : __ddiv_rn(cX, ksai);
  const double qx = __dadd_rn(__dadd_rn(__ddiv_rn(__dmul_rn(I.fx, X), ksai), su), I.cx);
  const double qy = __dadd_rn(__dadd_rn(sv, __ddiv_rn(__dmul_rn(I.fy, Y), ksai)), I.cy);
  const double dx = __dsub_rn(ox, qx), dy = __dsub_rn(oy, qy);
  return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

// Sum over the K corners in corner order; PG_ILP independent chains per trip for latency.
#ifndef PG_ILP
#define PG_ILP 2
#endif
template <typename PixelLoad>
__device__ __forceinline__ double pg_reproject_error(const PgIntr& I, const double* pose, const double* worlds,
                                                     int K, PixelLoad pixel) {
  double e = 0.0;
  int j = 0;
  for (; j + PG_ILP <= K; j += PG_ILP) {
    double ej[PG_ILP];
#pragma unroll
    for (int u = 0; u < PG_ILP; ++u) {
      const double2 p = pixel(j + u);
      ej[u] = pg_corner_error(I, pose, worlds + 3 * (j + u), p.x, p.y);
    }
#pragma unroll
    for (int u = 0; u < PG_ILP; ++u) e = __dadd_rn(e, ej[u]);
  }
  for (; j < K; ++j) {
    const double2 p = pixel(j);
    e = __dadd_rn(e, pg_corner_error(I, pose, worlds + 3 * j, p.x, p.y));
  }
  return e;
}

struct PgPairArgs {
  const double* pixels;     // [C][B][K][2] as handed in
  const int32_t* shared;    // [n] board index of every shared board, ascending
  const double* base;       // [n][2][12]: side 0 = (Ri | ti) of camera i, side 1 = (Rp | tp) of camera i-1
  const double* cand;       // [nc][2][12]: T1 = (R_ki | t_ki), T2 = (R_ik | t_ik) per candidate
  const double* worlds;     // [K][3]
  double* E;                // [n][2][nc]
  PgIntr intr[2];           // side 0 is scored through camera i-1, side 1 through camera i
  int64_t cam_stride;       // B * K * 2
  int32_t cam[2];           // camera of side 0 / side 1
  int32_t n, nc, K, tile;   // tile = shared boards per CTA
};

__host__ __device__ inline size_t pg_pair_smem_bytes(int K, int tile) {
  return ((((size_t)3 * K + 1) & ~(size_t)1) + (size_t)tile * 2 * (12 + 2 * (size_t)K)) * sizeof(double);
}

__global__ void __launch_bounds__(kPgThreads, PG_MIN_CTAS)
k_pg_pair_score(PgPairArgs A) {
  extern __shared__ __align__(16) double pg_smem[];
  double* s_w = pg_smem;                          // worlds
  double* s_f = pg_smem + ((3 * A.K + 1) & ~1);   // [tile][2][12 + 2K], 16-byte aligned rows
  const int fstride = 12 + 2 * A.K;
  const int f0 = blockIdx.y * A.tile;
  const int nf = min(A.tile, A.n - f0);
  for (int k = threadIdx.x; k < 3 * A.K; k += blockDim.x) s_w[k] = A.worlds[k];
  for (int q = threadIdx.x; q < nf * 2 * fstride; q += blockDim.x) {
    const int fs = q / fstride, e = q - fs * fstride;      // fs = frame-in-tile * 2 + side
    const int f = f0 + (fs >> 1), side = fs & 1;
    double v;
    if (e < 12) v = A.base[((size_t)f * 2 + side) * 12 + e];
    else v = A.pixels[(size_t)A.cam[side] * A.cam_stride + (size_t)A.shared[f] * (2 * A.K) + (e - 12)];
    s_f[q] = v;
  }
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= A.nc) return;
  double T[2][12];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int e = 0; e < 12; ++e) T[s][e] = A.cand[((size_t)c * 2 + s) * 12 + e];
  for (int f = 0; f < nf; ++f) {
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      const double* fr = s_f + (f * 2 + side) * fstride;
      double pose[12];
      pg_compose(T[side], fr, pose);
      const double2* px = reinterpret_cast<const double2*>(fr + 12);
      const double e = pg_reproject_error(A.intr[side], pose, s_w, A.K, [&](int j) { return px[j]; });
      A.E[((size_t)(f0 + f) * 2 + side) * A.nc + c] = e;
    }
  }
}

// error[c] = sum over the shared boards, in board order, of (error of camera i-1) then (error of
// camera i): the `error += e` sequence of multi_calib.cpp:57-80.
__global__ void k_pg_sum(const double* __restrict__ E, int rows, int nc, double* __restrict__ err) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  double s = 0.0;
  for (int k = 0; k < rows; ++k) s = __dadd_rn(s, E[(size_t)k * nc + c]);
  err[c] = s;
}

struct PgBoardArgs {
  const double* pixels;     // [C][B][K][2]
  const uint8_t* has;       // [C][B]
  const double* cam_pose;   // [C][12]
  const double* cand;       // [B][C][12]: candidate of board i built from camera j
  const double* worlds;     // [K][3]
  const double* intr;       // [C][9]
  double* E;                // [B][C][C]: error of candidate j of board i seen through camera k
  int32_t C, B, K;
};

// thread = (board i, candidate camera j, scoring camera k): multi_calib.cpp:134-141
__global__ void __launch_bounds__(kPgThreads)
k_pg_board_score(PgBoardArgs A) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)A.B * A.C * A.C;
  if (t >= total) return;
  const int k = (int)(t % A.C), j = (int)((t / A.C) % A.C), i = (int)(t / ((int64_t)A.C * A.C));
  if (!A.has[(size_t)j * A.B + i] || !A.has[(size_t)k * A.B + i]) return;
  double Tc[12], cd[12], pose[12];
#pragma unroll
  for (int e = 0; e < 12; ++e) { Tc[e] = A.cam_pose[12 * k + e]; cd[e] = A.cand[((size_t)i * A.C + j) * 12 + e]; }
  pg_compose(Tc, cd, pose);
  const PgIntr I = pg_intr(A.intr + 9 * k);
  const double2* px = reinterpret_cast<const double2*>(A.pixels + ((size_t)k * A.B + i) * (2 * (size_t)A.K));
  A.E[t] = pg_reproject_error(I, pose, A.worlds, A.K, [&](int c) { return __ldg(px + c); });
}

// ---- host side: the O(n) algebra around the scoring (multi_calib.cpp:26-48, 114-127) ------
// Plain double / float statements in the reference's order.  The host compiler targets baseline
// x86-64 (no FMA instructions), so nothing contracts here either.
namespace pg_host {

// Rt_to_R_t, multi_calib.h:130-137: r1, r2 narrowed to cv::Vec3f, r3 = r1.cross(r2) in float
inline void split(const double* Rt, double* pose) {
  const float r1[3] = {(float)Rt[0], (float)Rt[3], (float)Rt[6]};
  const float r2[3] = {(float)Rt[1], (float)Rt[4], (float)Rt[7]};
  const float r3[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
  for (int k = 0; k < 3; ++k) { pose[3 * k] = r1[k]; pose[3 * k + 1] = r2[k]; pose[3 * k + 2] = r3[k]; }
  pose[9] = Rt[2]; pose[10] = Rt[5]; pose[11] = Rt[8];
}
// o = op(a) * op(b), accumulated from 0 with k ascending like a cv::Mat product
inline void mul33(const double* a, bool at, const double* b, bool bt, double* o) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += (at ? a[3 * k + i] : a[3 * i + k]) * (bt ? b[3 * j + k] : b[3 * k + j]);
      o[3 * i + j] = s;
    }
}
inline void mul31(const double* a, bool at, const double* v, double* o) {
  for (int i = 0; i < 3; ++i) {
    double s = 0.0;
    for (int k = 0; k < 3; ++k) s += (at ? a[3 * k + i] : a[3 * i + k]) * v[k];
    o[i] = s;
  }
}
// candidate pose of camera i through one shared board (multi_calib.cpp:36-47)
inline void chain_candidate(const double* Pi, const double* Pp, const double* camk, double* cand) {
  double R_ik[9], t_ik[3], v[3];
  mul33(Pi, false, Pp, true, R_ik);                 // chess_R_i * chess_R_k.t()
  mul31(R_ik, false, Pp + 9, v);
  for (int r = 0; r < 3; ++r) t_ik[r] = Pi[9 + r] - v[r];
  mul33(R_ik, false, camk, false, cand);            // R_ik * cameras_[i-1].R()
  mul31(R_ik, false, camk + 9, v);
  for (int r = 0; r < 3; ++r) cand[9 + r] = v[r] + t_ik[r];
}
// the two relative transforms a candidate is scored with (multi_calib.cpp:63-66 and 73-76)
inline void chain_transforms(const double* camk, const double* cand, double* T) {
  double v[3];
  mul33(camk, false, cand, true, T);                // R_ki = camera_R_k * Rs[j].t()
  mul31(T, false, cand + 9, v);
  for (int r = 0; r < 3; ++r) T[9 + r] = camk[9 + r] - v[r];
  mul33(cand, false, camk, true, T + 12);           // R_ik = Rs[j] * camera_R_k.t()
  mul31(T + 12, false, camk + 9, v);
  for (int r = 0; r < 3; ++r) T[12 + 9 + r] = cand[9 + r] - v[r];
}
// candidate pose of a board through one camera that sees it (multi_calib.cpp:107-112, 121-126)
inline void board_candidate(const double* cam, const double* Pb, double* cand) {
  double d[3];
  mul33(cam, true, Pb, false, cand);                // camera_R.t() * chess_R
  for (int r = 0; r < 3; ++r) d[r] = Pb[9 + r] - cam[9 + r];
  mul31(cam, true, d, cand + 9);                    // camera_R.t() * (chess_t - camera_t)
}

}  // namespace pg_host

}  // namespace tscm
