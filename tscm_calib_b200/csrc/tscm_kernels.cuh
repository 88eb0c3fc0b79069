// tscm_kernels.cuh — sm_100a FP64 kernels of the calibration solve.
//
// Stage map (reference stage each replaces):
//   k_prep_cams, k_eval5,
//   k_view_blocks (tscm_eval5.cuh)  K1+K2  Jet evaluation of multi_calib.h:146-195 for every
//                                     residual block + SchurEliminator's E^T E / E^T F /
//                                     F^T F / E^T b / F^T b products (multi_calib.cpp:162-207)
//   k_post_eval                K2     per-camera F^T F, F^T b, cost; gradient max-norm
//   k_schur_frames + k_schur_update,
//   k_schur2, k_pair_* (+ tscm_schur_pairs.cuh),
//   k_reduce_s                 K3     SchurEliminator::Eliminate chunk loop
//   k_solve (tscm_solve.cuh)   K4     DenseSchurComplementSolver (Eigen LLT) on the reduced
//                                     camera system
//   k_backsub                  K4     SchurEliminator::BackSubstitute + candidate point
//   k_init, k_decide (or the
//   tail of k_post_eval)       K5     TrustRegionMinimizer / LevenbergMarquardtStrategy
//                                     control flow (accept/reject, radius, termination)
#pragma once

#include <cstdint>

#include "tscm_math.cuh"

namespace tscm {

// ---------------------------------------------------------------------------
// Device-resident LM state (one per solver).  Every kernel of the iteration
// graph reads it; only k_solve / k_post_eval / k_decide / k_init write it.
// ---------------------------------------------------------------------------
struct LmOptions {
  int max_num_iterations;
  double function_tolerance, gradient_tolerance, parameter_tolerance;
  double initial_radius, max_radius, min_radius;
  double min_relative_decrease, min_lm_diagonal, max_lm_diagonal;
  int max_num_consecutive_invalid_steps;
  int jacobi_scaling;
  int loss_type;
  double loss_scale;
  int ptol_needs_success;
  int disable_tolerances;
};

struct LmState {
  int done;
  int termination;
  int iteration;            // index of the last finalized iteration
  int recorded;             // entries written to the trace
  int cur;                  // which of the two parameter/Gram buffers holds x
  int num_successful, num_unsuccessful, num_consecutive_invalid;
  int atleast_one_successful_step;
  int solve_ok;             // reduced-system Cholesky succeeded (k_solve)
  int pad0, pad1;
  double radius, decrease_factor;
  double x_cost, x_norm, gmax;
  double initial_cost, final_cost, minimum_cost;
  double model_cost_change, candidate_cost, step_norm;  // last step (diagnostics)
  // camera-side partial sums of the current step (k_solve)
  double cam_lin, cam_quad, cam_dn2, cam_xn2;
};

// Comm record: what must be globally summed after an evaluation
//   [0, C*kCamRec)            per-camera U / g_c / cost / err sums
//   +0 frame lin, +1 frame quad, +2 frame |delta|^2, +3 frame |x+|^2
constexpr int kCommExtra = 4;

struct Trace {
  double* cost; double* radius; double* gmax; double* step_norm; int* flags;
  int capacity;
};

// Programmatic dependent launch (CUDA graphs keep the edges): every kernel of the LM-iteration
// graph lets its successor's CTAs become resident as soon as all of its own have started, and
// waits for its predecessor's memory before it reads anything an earlier kernel produced.  The
// semantics are those of plain stream order; what overlaps is launch latency and drain.
__device__ __forceinline__ void pdl_entry() {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// Static problem description on the device.
struct DeviceProblem {
  int C, F, K, V, Vpad, fixed_camera;
  int NL;                    // live reduced-system size: sum_c (free_rt ? 13 : 7)
  int Q;                     // NL (NL + 1) / 2
  const double2* obsT;       // [K][Vpad] observed pixels, corner-major
  const double* board_xy;    // [K][2]
  const int* view_camera;    // [V]
  const int* view_frame;     // [V]
  const int* frame_ptr;      // [F+1] CSR over frames
  const int* frame_views;    // [V] view ids grouped by frame (camera order)
  const int* cam_view_begin; // [C+1] views of camera c are [begin[c], begin[c+1])
  const int* live_off;       // [C+1] offset of camera c in the live reduced system
  const short* live_cam;     // [NL]
  const short* live_kk;      // [NL] index inside the camera's [rt 6 | intr 7] block
  const short* q_i;          // [Q] row of packed upper entry q
  const short* q_j;          // [Q] col
  // camera-partial slots written by the evaluation kernel: CTA b (views 32b..32b+31)
  // owns slots [blk_slot[b], blk_slot[b+1]), one per distinct camera among its views;
  // views are camera-major, so camera c's slots are [cam_slot_begin[c], cam_slot_begin[c+1])
  const int* blk_slot;       // [nblk+1]
  const int* cam_slot_begin; // [C+1]
  int nslot;
};

// One parameter set (there are two: current / candidate).
struct ParamSet {
  double* intr;      // [C][9]
  double* cam_rt;    // [C][6]
  double* board_rt;  // [F][6]
  CamConst* cam;     // [C] derived constants (k_prep_cams)
  double* G;         // [V][kViewStride] per-view records (BB | BC | BI)
  double* cam_part;  // [nslot][kCamRec] camera-block partial sums of the evaluation kernel
  double* comm;      // [C*kCamRec + kCommExtra] (+ gmax stored separately)
  double* gmax;      // [1]
};

// ---------------------------------------------------------------------------
// K1: per-camera constants
// ---------------------------------------------------------------------------
__global__ void k_prep_cams(DeviceProblem P, ParamSet ps0, ParamSet ps1, const LmState* st,
                            int which /*0 = current, 1 = candidate, 2/3 = absolute*/) {
  const int sel = which >= 2 ? which - 2 : (st->cur ^ which);
  if (which < 2 && st->done) return;
  const ParamSet& ps = sel ? ps1 : ps0;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.C) return;
  CamConst cc;
  make_cam_const(ps.cam_rt + 6 * c, ps.intr + 9 * c, c != P.fixed_camera, cc);
  ps.cam[c] = cc;
}

// ---------------------------------------------------------------------------
// K1+K2: residual + analytic Jacobian + normal-equation blocks live in tscm_eval5.cuh
// (k_eval5 + k_view_blocks).  Shared here: the structural-zero pattern of the intrinsic
// block and the size of the per-view frame-constant record.
// ---------------------------------------------------------------------------
// The intrinsic block of a residual row has structural zeros (u does not depend on fy, cy;
// v not on fx, cx — TS.h:124-125), which also survive the loss scaling.
__device__ __forceinline__ constexpr bool e3_live(int row, int icol) {
  // icol: 0 fx, 1 fy, 2 cx, 3 cy, 4 xi, 5 lambda, 6 alpha, 7 residual
  return row == 0 ? !(icol == 1 || icol == 3) : !(icol == 0 || icol == 2);
}
constexpr int kFcElems = 27;     // doubles of a FrameConst

// Inspection kernel: one thread per observation, full residual + Jacobian rows
// in the reference's column order (camera_rt 6, chessboard_rt 6, intrinsic 9).
__global__ void k_eval_rows(DeviceProblem P, ParamSet ps, LmOptions opt, double* residuals,
                            double* jacobian) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)P.V * P.K) return;
  const int v = (int)(idx / P.K), j = (int)(idx % P.K);
  const CamConst cc = ps.cam[P.view_camera[v]];
  FrameConst fc;
  make_frame_const(ps.board_rt + 6 * P.view_frame[v], fc);
  const double2 uv = P.obsT[(size_t)j * P.Vpad + v];
  ObsRow o;
  obs_jacobian<true, true, true>(cc, fc, P.board_xy[2 * j], P.board_xy[2 * j + 1], uv.x, uv.y, o);
  double err;
  obs_apply_loss<0, 19>(opt.loss_type, opt.loss_scale, o, &err);
  if (residuals) { residuals[2 * idx] = o.Ju[19]; residuals[2 * idx + 1] = o.Jv[19]; }
  if (jacobian) {
    double* Ju = jacobian + idx * 42;
    double* Jv = Ju + 21;
    for (int k = 0; k < 6; ++k) { Ju[k] = o.Ju[6 + k]; Jv[k] = o.Jv[6 + k]; }
    for (int k = 0; k < 6; ++k) { Ju[6 + k] = o.Ju[k]; Jv[6 + k] = o.Jv[k]; }
    for (int k = 0; k < 7; ++k) { Ju[12 + k] = o.Ju[12 + k]; Jv[12 + k] = o.Jv[12 + k]; }
    Ju[19] = Ju[20] = Jv[19] = Jv[20] = 0.0;
  }
}

// Observation transpose [V][K] (reference layout, cv::Point2d) -> [K][Vpad].
__global__ void k_transpose_obs(const double2* __restrict__ in, double2* __restrict__ out, int V,
                                int K, int Vpad) {
  __shared__ double2 tile[32][33];
  const int v0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int v = v0 + r, j = j0 + threadIdx.x;
    if (v < V && j < K) tile[r][threadIdx.x] = in[(size_t)v * K + j];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int j = j0 + r, v = v0 + threadIdx.x;
    if (v < V && j < K) out[(size_t)j * Vpad + v] = tile[threadIdx.x][r];
  }
}

// Block-wide deterministic sum / max helpers (blockDim.x a multiple of 32, <= 1024; s_red
// holds >= 33 doubles).  Warp shuffles, one shared-memory hop, two barriers: the fixed
// butterfly order makes the result independent of scheduling.
template <typename Op>
__device__ __forceinline__ double block_reduce(double v, double* s_red, Op op, double identity) {
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double w = lane < nw ? s_red[lane] : identity;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w = op(w, __shfl_xor_sync(0xffffffffu, w, o));
    if (lane == 0) s_red[32] = w;
  }
  __syncthreads();
  const double r = s_red[32];
  __syncthreads();
  return r;
}
__device__ __forceinline__ double block_sum(double v, double* s_red) {
  return block_reduce(v, s_red, [](double a, double b) { return a + b; }, 0.0);
}
__device__ __forceinline__ double block_max(double v, double* s_red) {
  return block_reduce(v, s_red, [](double a, double b) { return fmax(a, b); }, 0.0);
}

// ---------------------------------------------------------------------------
// Jacobi scaling (iteration 0): scale = 1 / (1 + sqrt(sum J_col^2)).
// ---------------------------------------------------------------------------
__global__ void k_jacobi_scale(DeviceProblem P, ParamSet ps0, ParamSet ps1, const LmState* st,
                               LmOptions opt, double* __restrict__ scale_e /*[F][6]*/,
                               double* __restrict__ scale_c /*[C][13]*/) {
  const ParamSet& ps = st->cur ? ps1 : ps0;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < P.F * 6) {
    const int f = idx / 6, b = idx % 6;
    double d = 0.0;
    for (int p = P.frame_ptr[f]; p < P.frame_ptr[f + 1]; ++p)
      d += ps.G[(size_t)P.frame_views[p] * kViewStride + kOffBB + tri6(b, b)];
    scale_e[idx] = opt.jacobi_scaling ? 1.0 / (1.0 + sqrt(d)) : 1.0;
  } else if (idx < P.F * 6 + P.C * 13) {
    const int k = idx - P.F * 6;
    const int c = k / 13, kk = k % 13;
    const double* U = ps.comm + c * kCamRec;   // globally summed camera record
    const double d = kk < 6 ? U[kCamCC + tri6(kk, kk)] : U[kCamII + tri8(kk - 6, kk - 6)];
    scale_c[k] = opt.jacobi_scaling ? 1.0 / (1.0 + sqrt(d)) : 1.0;
  }
}

// ---------------------------------------------------------------------------
// K3: Schur elimination of the frame poses onto the reduced camera system.
// A CTA walks a contiguous range of frames in batches of kSchurFB (one warp per
// frame): V_f + D_e^2 -> Cholesky -> Y = V^-1 W_s, z = V^-1 g_s staged in shared
// memory; then every thread updates the 4x4 register tile(s) of S it owns with
//   S_ij -= sum_r W_s[r][i] Y[r][j]      (r over the 6 rows of the batch's frames)
// Tiles are enumerated row-major over the upper block triangle, so the lanes of
// a warp share the W operand (broadcast) and read consecutive Y operands.
// Per-frame record saved for the back-substitution (SoA over frames):
//   [0,21) L (row-major lower, reciprocal diagonal), [21,27) z, [27,48) V_s upper,
//   [48,54) g_s
// ---------------------------------------------------------------------------
constexpr int kSchurFB = 8;
constexpr int kFrameRec = 54;

struct SchurArgs {
  const double* scale_e;   // [F][6]
  const double* scale_c;   // [C][13]
  double* frame_rec;       // [kFrameRec][Fpad]
  int Fpad;
  double* Spart;           // [nblk][Q]
  double* rpart;           // [nblk][NL]
  int frames_per_block;
  double radius_override;  // > 0: use instead of st->radius (inspection)
  const short* tile_bi;    // [ntiles] block row of tile t
  const short* tile_bj;    // [ntiles] block col (bj >= bi)
  int ntiles;
  int NLp;                 // NL rounded up to a multiple of 4
};

// ---------------------------------------------------------------------------
// K3, pipelined form (used when two staging buffers fit in shared memory):
// kSchurFB PRODUCER warps (one frame each: gather V / g_e, damp, 6x6 Cholesky,
// Y = V^-1 W_s, z) fill staging buffer (b+1)&1 while the CONSUMER warps fold
// buffer b&1 into their 4x4 register tiles of S; one __syncthreads per batch.
// The per-frame column list (view, in-camera index, reduced column) is
// precomputed on the host, so a producer issues all its global loads — the 27
// V/g entries over the frame's views and the raw W columns — back to back
// instead of through chains of dependent index loads.
// ---------------------------------------------------------------------------
struct Schur2Args {
  SchurArgs a;
  const int* col_ptr;      // [F+1] first column descriptor of frame f
  const int* col_src;      // [ncol] view * 16 + kk
  const short* col_g;      // [ncol] reduced (live) column
};

constexpr int kS2ColsPerLane = 4;   // raw W columns a lane keeps in registers per round

__global__ void __launch_bounds__(640)
k_schur2(DeviceProblem P, ParamSet ps0, ParamSet ps1, const LmState* st, LmOptions opt,
         Schur2Args B) {
  pdl_entry();
  if (st->done) return;
  const SchurArgs& A = B.a;
  const ParamSet& ps = st->cur ? ps1 : ps0;
  const double radius = A.radius_override > 0.0 ? A.radius_override : st->radius;
  extern __shared__ __align__(32) double s_mem[];
  const int NL = P.NL, NLp = A.NLp;
  const int buf_doubles = 2 * kSchurFB * 6 * NLp + kSchurFB * 6;   // Ws | Ys | zs
  double* scr = s_mem + 2 * buf_doubles;                            // [FB][64]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int f_begin = blockIdx.x * A.frames_per_block;
  const int f_end = min(P.F, f_begin + A.frames_per_block);
  const int nbatch = (f_end - f_begin + kSchurFB - 1) / kSchurFB;

  // One loop, ONE barrier site for both roles (CUDA allows __syncthreads() in conditional code
  // only if the whole block takes the same branch: compute-sanitizer synccheck flagged the two
  // role-private loops of round 1, profiles/r02_sanitizer_synccheck_robust.log).
  const bool producer = warp < kSchurFB;
  double* my = scr + (producer ? warp : 0) * 64;
  const int ct = tid - kSchurFB * 32;            // consumer thread index (negative: producer)
  const int tbi = ct < A.ntiles ? A.tile_bi[ct] : -1;
  const int tbj = ct < A.ntiles ? A.tile_bj[ct] : 0;
  double acc[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) acc[e] = 0.0;
  double racc = 0.0;
  for (int b = 0; b <= nbatch; ++b) {
    if (producer) {
      // ------------------------------- producer -------------------------------------
      if (b < nbatch) {
        double* Ws = s_mem + (b & 1) * buf_doubles;
        double* Ys = Ws + kSchurFB * 6 * NLp;
        double* zs = Ys + kSchurFB * 6 * NLp;
        // clear this warp's 12 staging rows (columns of cameras that do not see the frame)
        for (int i = lane; i < 6 * NLp; i += 32) {
          Ws[warp * 6 * NLp + i] = 0.0;
          Ys[warp * 6 * NLp + i] = 0.0;
        }
        if (lane < 6) zs[warp * 6 + lane] = 0.0;
        const int f = f_begin + b * kSchurFB + warp;
        if (f < f_end) {
          const int p0 = P.frame_ptr[f], nv = P.frame_ptr[f + 1] - p0;
          const int c0 = B.col_ptr[f], ncols = B.col_ptr[f + 1] - c0;
          // independent loads first: view ids, column descriptors
          int vid = lane < nv ? P.frame_views[p0 + lane] : 0;
          int csrc[kS2ColsPerLane], cg[kS2ColsPerLane];
#pragma unroll
          for (int r = 0; r < kS2ColsPerLane; ++r) {
            const int col = lane + 32 * r;
            csrc[r] = col < ncols ? B.col_src[c0 + col] : -1;
            cg[r] = col < ncols ? B.col_g[c0 + col] : 0;
          }
          double se[6];
#pragma unroll
          for (int i = 0; i < 6; ++i) se[i] = A.scale_e[f * 6 + i];
          // second level: V / g_e entries and raw W columns
          double raw[kS2ColsPerLane][6];
#pragma unroll
          for (int r = 0; r < kS2ColsPerLane; ++r) {
            if (csrc[r] >= 0) {
              const int kk = csrc[r] & 15;
              const double* Gv = ps.G + (size_t)(csrc[r] >> 4) * kViewStride;
#pragma unroll
              for (int q = 0; q < 6; ++q)
                raw[r][q] = kk < 6 ? Gv[kOffBC + q * 6 + kk] : Gv[kOffBI + q * 8 + (kk - 6)];
            } else {
#pragma unroll
              for (int q = 0; q < 6; ++q) raw[r][q] = 0.0;
            }
          }
          {
            const int e = lane < 21 ? kOffBB + lane : kOffBI + (min(lane, 26) - 21) * 8 + 7;
            double s0 = 0.0, s1 = 0.0;
            int p = 0;
            for (; p + 1 < nv; p += 2) {
              const int va = __shfl_sync(0xffffffffu, vid, p), vb = __shfl_sync(0xffffffffu, vid, p + 1);
              s0 += ps.G[(size_t)va * kViewStride + e];
              s1 += ps.G[(size_t)vb * kViewStride + e];
            }
            if (p < nv) {
              const int va = __shfl_sync(0xffffffffu, vid, p);
              s0 += ps.G[(size_t)va * kViewStride + e];
            }
            if (lane < 27) my[lane] = s0 + s1;
          }
          __syncwarp();
          double M[36], gs[6];
#pragma unroll
          for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int j = i; j < 6; ++j) {
              const double v = se[i] * se[j] * my[tri6(i, j)];
              M[i * 6 + j] = v;
              M[j * 6 + i] = v;
            }
            gs[i] = se[i] * my[21 + i];
          }
          __syncwarp();
          if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 6; ++i) {
#pragma unroll
              for (int j = i; j < 6; ++j) my[27 + tri6(i, j)] = M[i * 6 + j];
              my[48 + i] = gs[i];
            }
          }
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const double d = fmin(fmax(M[i * 6 + i], opt.min_lm_diagonal), opt.max_lm_diagonal);
            const double D = sqrt(d / radius);
            M[i * 6 + i] += D * D;
          }
          chol6(M);
          double z[6];
#pragma unroll
          for (int i = 0; i < 6; ++i) z[i] = gs[i];
          chol6_solve(M, z);
          if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 6; ++i) {
#pragma unroll
              for (int j = 0; j <= i; ++j) my[(i * (i + 1)) / 2 + j] = M[i * 6 + j];
              my[21 + i] = z[i];
              zs[warp * 6 + i] = z[i];
            }
          }
          __syncwarp();
          for (int i = lane; i < kFrameRec; i += 32) A.frame_rec[(size_t)i * A.Fpad + f] = my[i];
          // scaled W columns and Y = (V + D^2)^-1 W_s
#pragma unroll
          for (int r = 0; r < kS2ColsPerLane; ++r) {
            if (csrc[r] >= 0) {
              const int kk = csrc[r] & 15;
              const int m = P.view_camera[csrc[r] >> 4];
              const double sc = A.scale_c[m * 13 + kk];
              double w[6];
#pragma unroll
              for (int q = 0; q < 6; ++q) w[q] = se[q] * raw[r][q] * sc;
#pragma unroll
              for (int q = 0; q < 6; ++q) Ws[(warp * 6 + q) * NLp + cg[r]] = w[q];
              chol6_solve(M, w);
#pragma unroll
              for (int q = 0; q < 6; ++q) Ys[(warp * 6 + q) * NLp + cg[r]] = w[q];
            }
          }
          // frames with more live columns than a lane holds in registers (C > 9)
          for (int col = lane + 32 * kS2ColsPerLane; col < ncols; col += 32) {
            const int src = B.col_src[c0 + col], g = B.col_g[c0 + col];
            const int kk = src & 15;
            const int m = P.view_camera[src >> 4];
            const double sc = A.scale_c[m * 13 + kk];
            const double* Gv = ps.G + (size_t)(src >> 4) * kViewStride;
            double w[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) {
              const double rv = kk < 6 ? Gv[kOffBC + q * 6 + kk] : Gv[kOffBI + q * 8 + (kk - 6)];
              w[q] = se[q] * rv * sc;
            }
#pragma unroll
            for (int q = 0; q < 6; ++q) Ws[(warp * 6 + q) * NLp + g] = w[q];
            chol6_solve(M, w);
#pragma unroll
            for (int q = 0; q < 6; ++q) Ys[(warp * 6 + q) * NLp + g] = w[q];
          }
        }
      }
    } else {
      // ------------------------------- consumer -------------------------------------
      if (b > 0) {
        const double* Ws = s_mem + ((b - 1) & 1) * buf_doubles;
        const double* Ys = Ws + kSchurFB * 6 * NLp;
        const double* zs = Ys + kSchurFB * 6 * NLp;
        if (tbi >= 0) {
          const double* wp = Ws + 4 * tbi;
          const double* yp = Ys + 4 * tbj;
  #pragma unroll 4
          for (int r = 0; r < kSchurFB * 6; ++r) {
            const double4 w4 = *reinterpret_cast<const double4*>(wp + r * NLp);
            const double4 y4 = *reinterpret_cast<const double4*>(yp + r * NLp);
            const double w[4] = {w4.x, w4.y, w4.z, w4.w};
            const double y[4] = {y4.x, y4.y, y4.z, y4.w};
  #pragma unroll
            for (int a = 0; a < 4; ++a)
  #pragma unroll
              for (int c = 0; c < 4; ++c) acc[a * 4 + c] = fma(-w[a], y[c], acc[a * 4 + c]);
          }
        }
        if (ct < NL) {
          double a = racc;
          for (int r = 0; r < kSchurFB * 6; ++r) a = fma(-Ws[r * NLp + ct], zs[r], a);
          racc = a;
        }
      }
    }
    __syncwarp();          // lane-divergent column loops reconverge before the aligned CTA barrier
    __syncthreads();
  }
  if (producer) return;
  double* Sp = A.Spart + (size_t)blockIdx.x * P.Q;
  if (tbi >= 0) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int i = 4 * tbi + a, j = 4 * tbj + c;
        if (i <= j && j < NL) Sp[i * NL - (i * (i - 1)) / 2 + (j - i)] = acc[a * 4 + c];
      }
    }
  }
  if (ct < NL) A.rpart[(size_t)blockIdx.x * NL + ct] = racc;
}

// ---------------------------------------------------------------------------
// K3, split form (default for dense-ish visibility): the per-frame work and the
// accumulation of S are separate kernels.
//  k_schur_frames: one warp per frame with the whole GPU's occupancy hiding the
//    dependent-load and FP64 latency chains (gather V / g_e, damp, 6x6 Cholesky,
//    Y = V^-1 W_s, z); W_s, Y (6 x NLp rows per frame) and z go to global memory.
//  k_schur_update: each CTA streams a contiguous range of frames through shared
//    memory with TMA bulk copies (cp.async.bulk + mbarrier, two stages of 8 frames)
//    and folds them into the 4x4 register tiles of S it owns.
// ---------------------------------------------------------------------------
// Column permutation of the materialised W_s / Y rows: the two low doubles of every 4-column
// tile first, then the two high doubles, so that the update kernel's 16-byte operand loads
// are contiguous across the lanes of a warp (no shared-memory bank conflicts).
__host__ __device__ __forceinline__ int schur_perm(int col, int NLp) {
  return ((col & 2) ? (NLp >> 1) : 0) + ((col >> 2) << 1) + (col & 1);
}

struct SchurSplitArgs {
  SchurArgs a;
  const int* col_ptr;      // [F+1]
  const int* col_src;      // [ncol] view * 16 + kk
  const short* col_g;      // [ncol] reduced (live) column
  const short* col_sidx;   // [ncol] camera * 13 + kk (index into scale_c)
  double* Wg;              // [F][6][NLp] scaled W rows, columns permuted (schur_perm)
  double* Yg;              // [F][6][NLp] (V + D^2)^-1 W_s
  double* zg;              // [Fpad8][6]   (V + D^2)^-1 g_s
  // sparse-visibility form (k_schur_pairs): per-VIEW 6 x 16 blocks instead of dense rows
  double* Wv;              // [V][6][16] scaled W_s block of the view's camera, column c at c
  double* Yv;              // [V][6][16] (V + D^2)^-1 W_s, column c at pair_pos(c); z at pair_pos(13)
  double* fact;            // [F][32] packed Cholesky factor (21, reciprocal diagonal) | z (6): the
                           // per-frame result k_pair_blocks applies to every view of the frame
};

// Column c (0..13) of a view's Y block sits at c + (c >= 7): two 8-double halves of 7 columns
// each, so that a lane of k_schur_pairs reads its 7 operands with aligned 16-byte loads.
__host__ __device__ __forceinline__ int pair_pos(int c) { return c + (c >= 7 ? 1 : 0); }

#ifndef TSCM_SF_MINB
#define TSCM_SF_MINB 2
#endif
#ifndef TSCM_SF_WARPS
#define TSCM_SF_WARPS 8      // frames (warps) per CTA
#endif
constexpr int kSfWarps = TSCM_SF_WARPS;
__global__ void __launch_bounds__(32 * kSfWarps, TSCM_SF_MINB)
k_schur_frames(DeviceProblem P, ParamSet ps0, ParamSet ps1, const LmState* st, LmOptions opt,
               SchurSplitArgs B) {
  pdl_entry();
  if (st->done) return;
  __shared__ double s_scr[kSfWarps][64];
  const SchurArgs& A = B.a;
  const ParamSet& ps = st->cur ? ps1 : ps0;
  const double radius = A.radius_override > 0.0 ? A.radius_override : st->radius;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.x * kSfWarps + warp;
  if (f >= P.F) return;
  const int NLp = A.NLp;
  double* my = s_scr[warp];
  double* Wf = B.Wg + (size_t)f * 6 * NLp;
  double* Yf = B.Yg + (size_t)f * 6 * NLp;
  const int p0 = P.frame_ptr[f], nv = P.frame_ptr[f + 1] - p0;
  const int c0 = B.col_ptr[f], ncols = B.col_ptr[f + 1] - c0;
  // cameras that do not see the frame contribute zero columns
  if (ncols < P.NL) {
    for (int i = lane; i < 6 * NLp; i += 32) { Wf[i] = 0.0; Yf[i] = 0.0; }
    __syncwarp();
  } else if (NLp > P.NL) {
    for (int q = lane; q < 6 * (NLp - P.NL); q += 32) {
      const int r = q / (NLp - P.NL), c = schur_perm(P.NL + q % (NLp - P.NL), NLp);
      Wf[r * NLp + c] = 0.0; Yf[r * NLp + c] = 0.0;
    }
  }
  const int vid = lane < nv ? P.frame_views[p0 + lane] : 0;
  double se[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) se[i] = A.scale_e[f * 6 + i];
  {
    const int e = lane < 21 ? kOffBB + lane : kOffBI + (min(lane, 26) - 21) * 8 + 7;
    double s0 = 0.0, s1 = 0.0;
    int p = 0;
    for (; p + 1 < nv; p += 2) {
      const int va = __shfl_sync(0xffffffffu, vid, p), vb = __shfl_sync(0xffffffffu, vid, p + 1);
      s0 += ps.G[(size_t)va * kViewStride + e];
      s1 += ps.G[(size_t)vb * kViewStride + e];
    }
    if (p < nv) s0 += ps.G[(size_t)__shfl_sync(0xffffffffu, vid, p) * kViewStride + e];
    if (lane < 27) my[lane] = s0 + s1;
  }
  __syncwarp();
  double M[36], gs[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int j = i; j < 6; ++j) {
      const double v = se[i] * se[j] * my[tri6(i, j)];
      M[i * 6 + j] = v;
      M[j * 6 + i] = v;
    }
    gs[i] = se[i] * my[21 + i];
  }
  __syncwarp();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
      for (int j = i; j < 6; ++j) my[27 + tri6(i, j)] = M[i * 6 + j];
      my[48 + i] = gs[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const double d = fmin(fmax(M[i * 6 + i], opt.min_lm_diagonal), opt.max_lm_diagonal);
    const double D = sqrt(d / radius);
    M[i * 6 + i] += D * D;
  }
  chol6(M);   // a failed pivot yields NaNs that the LM loop turns into an invalid step
  double z[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) z[i] = gs[i];
  chol6_solve(M, z);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
      for (int j = 0; j <= i; ++j) my[(i * (i + 1)) / 2 + j] = M[i * 6 + j];
      my[21 + i] = z[i];
      B.zg[(size_t)f * 6 + i] = z[i];
    }
  }
  __syncwarp();
  for (int i = lane; i < kFrameRec; i += 32) A.frame_rec[(size_t)i * A.Fpad + f] = my[i];
  // columns of W_s and Y: four per lane and round, with every global load of a round issued
  // before the first use (two dependent levels: descriptors, then records).  The Cholesky
  // factor is read from shared memory (packed, written above).
  constexpr int kCols = 4;
  for (int base = 0; base < ncols; base += 32 * kCols) {
    int csrc[kCols], cg[kCols];
    double csc[kCols];
#pragma unroll
    for (int r = 0; r < kCols; ++r) {
      const int col = base + lane + 32 * r;
      const bool ok = col < ncols;
      csrc[r] = ok ? B.col_src[c0 + col] : -1;
      cg[r] = ok ? schur_perm(B.col_g[c0 + col], NLp) : 0;
      csc[r] = ok ? A.scale_c[B.col_sidx[c0 + col]] : 0.0;
    }
    double raw[kCols][6];
#pragma unroll
    for (int r = 0; r < kCols; ++r) {
      const int kk = csrc[r] & 15;
      const double* Gv = ps.G + (size_t)(csrc[r] >= 0 ? (csrc[r] >> 4) : 0) * kViewStride;
#pragma unroll
      for (int q = 0; q < 6; ++q)
        raw[r][q] = csrc[r] >= 0 ? (kk < 6 ? Gv[kOffBC + q * 6 + kk] : Gv[kOffBI + q * 8 + (kk - 6)]) : 0.0;
    }
#pragma unroll
    for (int r = 0; r < kCols; ++r) {
      if (csrc[r] >= 0) {
        double w[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) w[q] = se[q] * raw[r][q] * csc[r];
#pragma unroll
        for (int q = 0; q < 6; ++q) Wf[q * NLp + cg[r]] = w[q];
        chol6_solve_packed(my, w);
#pragma unroll
        for (int q = 0; q < 6; ++q) Yf[q * NLp + cg[r]] = w[q];
      }
    }
  }
}

// Per-frame part of the sparse-visibility Schur form: V_s + D^2 -> Cholesky -> z, frame record
// and packed factor.  EIGHT lanes per frame (four frames per warp): the chain frame -> views ->
// 27 sums -> 6 x 6 factorisation is pure latency, so what counts is frames in flight per warp
// instruction (k_schur_frames, one warp per frame, spent 138 us on 40,000 frames here).  Same
// arithmetic and summation order as k_schur_frames.
__global__ void __launch_bounds__(256, 2)
k_pair_frames(DeviceProblem P, ParamSet ps0, ParamSet ps1, const LmState* st, LmOptions opt,
              SchurSplitArgs B) {
  pdl_entry();
  if (st->done) return;
  __shared__ double s_rec[32][kFrameRec + 2];
  const SchurArgs& A = B.a;
  const ParamSet& ps = st->cur ? ps1 : ps0;
  const double radius = A.radius_override > 0.0 ? A.radius_override : st->radius;
  const int lane = threadIdx.x & 31, g = lane & 7, grp = threadIdx.x >> 3;   // grp: 0..31 in the CTA
  const int f = blockIdx.x * 32 + grp;
  const bool active = f < P.F;
  const int fc = active ? f : P.F - 1;
  const int p0 = P.frame_ptr[fc], nv = P.frame_ptr[fc + 1] - p0;
  // lane g sums entries g, g + 8, g + 16, g + 24 (< 27) of [BB 21 | g_e 6] over the frame's views
  int eidx[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int e = min(g + 8 * j, 26);
    eidx[j] = e < 21 ? kOffBB + e : kOffBI + (e - 21) * 8 + 7;
  }
  double s0[4] = {0.0, 0.0, 0.0, 0.0}, s1[4] = {0.0, 0.0, 0.0, 0.0};
  int p = 0;
  for (; p + 1 < nv; p += 2) {
    const size_t va = (size_t)P.frame_views[p0 + p] * kViewStride, vb = (size_t)P.frame_views[p0 + p + 1] * kViewStride;
#pragma unroll
    for (int j = 0; j < 4; ++j) { s0[j] += ps.G[va + eidx[j]]; s1[j] += ps.G[vb + eidx[j]]; }
  }
  if (p < nv) {
    const size_t va = (size_t)P.frame_views[p0 + p] * kViewStride;
#pragma unroll
    for (int j = 0; j < 4; ++j) s0[j] += ps.G[va + eidx[j]];
  }
  double my[27];
  const int gbase = lane & ~7;
#pragma unroll
  for (int e = 0; e < 27; ++e) my[e] = __shfl_sync(0xffffffffu, s0[e >> 3] + s1[e >> 3], gbase + (e & 7));
  double se[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) se[i] = A.scale_e[fc * 6 + i];
  double M[36], gs[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int j = i; j < 6; ++j) {
      const double v = se[i] * se[j] * my[tri6(i, j)];
      M[i * 6 + j] = v;
      M[j * 6 + i] = v;
    }
    gs[i] = se[i] * my[21 + i];
  }
  double* rec = s_rec[grp];
  if (g == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
      for (int j = i; j < 6; ++j) rec[27 + tri6(i, j)] = M[i * 6 + j];
      rec[48 + i] = gs[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const double d = fmin(fmax(M[i * 6 + i], opt.min_lm_diagonal), opt.max_lm_diagonal);
    const double D = sqrt(d / radius);
    M[i * 6 + i] += D * D;
  }
  chol6(M);   // a failed pivot yields NaNs that the LM loop turns into an invalid step
  double z[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) z[i] = gs[i];
  chol6_solve(M, z);
  if (g == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
      for (int j = 0; j <= i; ++j) rec[(i * (i + 1)) / 2 + j] = M[i * 6 + j];
      rec[21 + i] = z[i];
    }
  }
  __syncwarp();
  if (active) {
    for (int i = g; i < kFrameRec; i += 8) A.frame_rec[(size_t)i * A.Fpad + f] = rec[i];
    for (int i = g; i < 27; i += 8) B.fact[(size_t)f * 32 + i] = rec[i];
  }
}

// Per-view blocks of the sparse-visibility Schur form: thread = (view, block column 0..15).
// Columns 0..12 are the camera's [rt 6 | intr 7] columns of the scaled W_s (a fixed camera's rt
// columns stay zero), Y = (V + D^2)^-1 W_s through the frame's packed factor, column 13 of Y
// carries z; every thread stores its whole column so that each 768-byte block is rewritten
// completely (full 128-byte rows, padding included).
__global__ void __launch_bounds__(256)
k_pair_blocks(DeviceProblem P, ParamSet ps0, ParamSet ps1, const LmState* st, SchurSplitArgs B) {
  pdl_entry();
  if (st->done) return;
  const ParamSet& ps = st->cur ? ps1 : ps0;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = t >> 4, c = t & 15;
  if (v >= P.V) return;
  const int m = P.view_camera[v], f = P.view_frame[v];
  const bool free_rt = P.live_off[m + 1] - P.live_off[m] == 13;
  const bool live = c < 13 && (free_rt || c >= 6);
  const double* Lp = B.fact + (size_t)f * 32;
  double w[6];
#pragma unroll
  for (int q = 0; q < 6; ++q) w[q] = 0.0;
  if (live) {
    const double csc = B.a.scale_c[m * 13 + c];
    const double* Gv = ps.G + (size_t)v * kViewStride;
    const double* se = B.a.scale_e + (size_t)f * 6;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const double raw = c < 6 ? Gv[kOffBC + q * 6 + c] : Gv[kOffBI + q * 8 + (c - 6)];
      w[q] = se[q] * raw * csc;
    }
  }
  double* Wb = B.Wv + (size_t)v * 96 + c;
#pragma unroll
  for (int q = 0; q < 6; ++q) Wb[q * 16] = w[q];
  if (live) {
    double L[21];
#pragma unroll
    for (int i = 0; i < 21; ++i) L[i] = Lp[i];
    chol6_solve_packed(L, w);
  } else if (c == 13) {
#pragma unroll
    for (int q = 0; q < 6; ++q) w[q] = Lp[21 + q];
  }
  // positions: columns 0..13 at pair_pos(c); the two padding slots (7, 15) take threads 14, 15
  const int pos = c < 14 ? pair_pos(c) : (c == 14 ? 7 : 15);
  double* Yb = B.Yv + (size_t)v * 96 + pos;
#pragma unroll
  for (int q = 0; q < 6; ++q) Yb[q * 16] = w[q];
}

// --- TMA bulk copy + mbarrier helpers (sm_90+/sm_100a) -----------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes,
                                             unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

template <int T>   // 4x4 tiles per thread
__global__ void __launch_bounds__(T == 1 ? 512 : 768)
k_schur_update(DeviceProblem P, const LmState* st, SchurSplitArgs B) {
  pdl_entry();
  if (st->done) return;
  const SchurArgs& A = B.a;
  extern __shared__ __align__(128) double s_mem[];
  const int NL = P.NL, NLp = A.NLp, NT = blockDim.x;
  const int rows_per_stage = kSchurFB * 6;
  const int stage_doubles = 2 * rows_per_stage * NLp + rows_per_stage;   // Ws | Ys | zs
  __shared__ __align__(8) unsigned long long s_bar[2];
  const int tid = threadIdx.x;
  int tbi[T], tbj[T];
  double acc[T][16];
#pragma unroll
  for (int k = 0; k < T; ++k) {
    const int t = tid + k * NT;
    tbi[k] = t < A.ntiles ? A.tile_bi[t] : -1;
    tbj[k] = t < A.ntiles ? A.tile_bj[t] : 0;
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[k][e] = 0.0;
  }
  double racc = 0.0;
  const int f_begin = blockIdx.x * A.frames_per_block;
  const int f_end = min(P.F, f_begin + A.frames_per_block);
  const int nchunk = (f_end - f_begin + kSchurFB - 1) / kSchurFB;
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int c) {
    const int f0 = f_begin + c * kSchurFB;
    const int nf = min(kSchurFB, f_end - f0);
    double* dst = s_mem + (c & 1) * stage_doubles;
    const unsigned wbytes = (unsigned)(nf * 6 * NLp * sizeof(double));
    const unsigned zbytes = (unsigned)(nf * 6 * sizeof(double));
    mbar_expect_tx(&s_bar[c & 1], 2 * wbytes + zbytes);
    tma_bulk_g2s(dst, B.Wg + (size_t)f0 * 6 * NLp, wbytes, &s_bar[c & 1]);
    tma_bulk_g2s(dst + rows_per_stage * NLp, B.Yg + (size_t)f0 * 6 * NLp, wbytes, &s_bar[c & 1]);
    tma_bulk_g2s(dst + 2 * rows_per_stage * NLp, B.zg + (size_t)f0 * 6, zbytes, &s_bar[c & 1]);
  };
  if (tid == 0 && nchunk > 0) issue(0);
  for (int c = 0; c < nchunk; ++c) {
    if (tid == 0 && c + 1 < nchunk) issue(c + 1);     // stage (c+1)&1 was released by the barrier below
    mbar_wait(&s_bar[c & 1], (c >> 1) & 1);
    const double* Ws = s_mem + (c & 1) * stage_doubles;
    const double* Ys = Ws + rows_per_stage * NLp;
    const double* zs = Ys + rows_per_stage * NLp;
    const int rows = min(kSchurFB, f_end - (f_begin + c * kSchurFB)) * 6;
#pragma unroll
    for (int k = 0; k < T; ++k) {
      if (tbi[k] >= 0) {
        // operands of tile (bi, bj) in the permuted row: low pair at 2 b, high pair at NLp/2 + 2 b
        const double* wl = Ws + 2 * tbi[k];
        const double* yl = Ys + 2 * tbj[k];
        const int hi = NLp >> 1;
        for (int r0 = 0; r0 < rows; r0 += 3) {     // rows is a multiple of 6
          double2 wa[3], wb[3], ya[3], yb[3];
#pragma unroll
          for (int u = 0; u < 3; ++u) {
            const int o = (r0 + u) * NLp;
            wa[u] = *reinterpret_cast<const double2*>(wl + o);
            wb[u] = *reinterpret_cast<const double2*>(wl + o + hi);
            ya[u] = *reinterpret_cast<const double2*>(yl + o);
            yb[u] = *reinterpret_cast<const double2*>(yl + o + hi);
          }
#pragma unroll
          for (int u = 0; u < 3; ++u) {
            const double w[4] = {wa[u].x, wa[u].y, wb[u].x, wb[u].y};
            const double y[4] = {ya[u].x, ya[u].y, yb[u].x, yb[u].y};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
              for (int b = 0; b < 4; ++b) acc[k][a * 4 + b] = fma(-w[a], y[b], acc[k][a * 4 + b]);
          }
        }
      }
    }
    if (tid < NL) {
      double a = racc;
      const int pc = schur_perm(tid, NLp);
      for (int r = 0; r < rows; ++r) a = fma(-Ws[r * NLp + pc], zs[r], a);
      racc = a;
    }
    __syncthreads();
  }
  double* Sp = A.Spart + (size_t)blockIdx.x * P.Q;
#pragma unroll
  for (int k = 0; k < T; ++k) {
    if (tbi[k] < 0) continue;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int i = 4 * tbi[k] + a, j = 4 * tbj[k] + b;
        if (i <= j && j < NL) Sp[i * NL - (i * (i - 1)) / 2 + (j - i)] = acc[k][a * 4 + b];
      }
    }
  }
  if (tid < NL) A.rpart[(size_t)blockIdx.x * NL + tid] = racc;
}

// U(i, j) of camera record `U` for in-camera indices a <= b (0..12 over [rt 6 | intr 7]).
__device__ __forceinline__ double cam_block(const double* U, int a, int b) {
  if (b < 6) return U[kCamCC + tri6(a, b)];
  if (a < 6) return U[kCamCI + a * 8 + (b - 6)];
  return U[kCamII + tri8(a - 6, b - 6)];
}
__device__ __forceinline__ double cam_grad(const double* U, int a) {
  return a < 6 ? U[kCamCI + a * 8 + 7] : U[kCamII + tri8(a - 6, 7)];
}

// ---------------------------------------------------------------------------
// The reduced camera system handed to k_solve: the AUGMENTED matrix [[lhs, rhs], [rhs^T, .]]
//   lhs = sum_f(-W^T V^-1 W) [Schur partials] + U_s + D_c^2,   rhs = sum_f(-W^T z) + g_s
// as 4x4 tiles of the lower block triangle, tile (bi, bj) at tile_id(bi, bj) * 16, entry
// (a, b) of a tile at a * 4 + b — exactly the 16 registers a k_solve thread owns, so that its
// load is eight 16-byte loads from one 128-byte line.  Entries outside the system (identity
// padding, the upper part of diagonal tiles) are never written and stay zero; k_solve
// overrides them.  The assembly (camera blocks U_s, the LM diagonal D_c^2 = clamp(diag) /
// radius, the gradient) happens HERE, in the thread that finishes an entry, because this
// kernel has 155 mostly idle CTAs while k_solve is one latency-bound CTA.
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ int solve_tile_id(int bi, int bj, int nbk) {   // bi >= bj, column-major
  return bj * nbk - (bj * (bj - 1)) / 2 + (bi - bj);
}
struct AssembleArgs {
  const double* scale_c;     // [C][13]
  double radius_override;    // > 0: use instead of st->radius (inspection)
  int nbk;                   // ceil((NL + 1) / 4)
  int add_cam;               // 0: raw sums only (NCCL fallback: the camera terms follow the all-reduce)
};
// position of linear entry i (< Q: packed upper S, then rhs) in the tile-major system
__device__ __forceinline__ int assemble_position(const DeviceProblem& P, const AssembleArgs& A, int i, int* row,
                                                 int* col) {
  int r, c;
  if (i < P.Q) { c = P.q_i[i]; r = P.q_j[i]; } else { c = i - P.Q; r = P.NL; }
  *row = r; *col = c;
  return solve_tile_id(r >> 2, c >> 2, A.nbk) * 16 + (r & 3) * 4 + (c & 3);
}
// U_s + D_c^2 (or the scaled gradient for the rhs row) of entry (r, c), c <= r
__device__ __forceinline__ double assemble_cam_term(const DeviceProblem& P, const ParamSet& ps, const LmOptions& opt,
                                                    const AssembleArgs& A, double radius, int r, int c) {
  const int cc = P.live_cam[c], kc = P.live_kk[c];
  const double* U = ps.comm + cc * kCamRec;
  const double sc_c = A.scale_c[cc * 13 + kc];
  if (r == P.NL) return sc_c * cam_grad(U, kc);
  if (P.live_cam[r] != cc) return 0.0;
  const int kr = P.live_kk[r];
  const double us = sc_c * A.scale_c[cc * 13 + kr] * cam_block(U, kc, kr);
  if (r != c) return us;
  const double d = fmin(fmax(us, opt.min_lm_diagonal), opt.max_lm_diagonal);
  const double D = sqrt(d / radius);
  return us + D * D;
}

// Sum the per-CTA partials.  A CTA owns 32 consecutive entries; its 8 warps each sum every 8th
// partial (4 loads in flight per thread, 256-byte coalesced rows) and the 8 warp sums are added
// in a fixed order: ~5 dependent L2 round trips instead of the 37 of a one-thread-per-entry sum.
constexpr int kReduceThreads = 256;
__global__ void __launch_bounds__(kReduceThreads)
k_reduce_s(DeviceProblem P, ParamSet ps0, ParamSet ps1, const LmState* st, LmOptions opt,
           const double* __restrict__ Spart, const double* __restrict__ rpart, int nblk,
           double* __restrict__ out, AssembleArgs A) {
  pdl_entry();
  if (st->done) return;
  __shared__ double s_part[8][33];
  const int e = threadIdx.x & 31, part = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + e;
  const bool ok = i < P.Q + P.NL;
  // the finishing thread's camera term is requested first: its loads overlap the partial sums
  int pos = 0;
  double cam = 0.0;
  if (part == 0 && ok) {
    int r, c;
    pos = assemble_position(P, A, i, &r, &c);
    if (A.add_cam)
      cam = assemble_cam_term(P, st->cur ? ps1 : ps0, opt, A, A.radius_override > 0.0 ? A.radius_override : st->radius, r, c);
  }
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (ok) {
    const double* src = i < P.Q ? Spart + i : rpart + (i - P.Q);
    const size_t stride = i < P.Q ? P.Q : P.NL;
    int b = part;
    for (; b + 24 < nblk; b += 32) {
      s0 += src[(size_t)b * stride];
      s1 += src[(size_t)(b + 8) * stride];
      s2 += src[(size_t)(b + 16) * stride];
      s3 += src[(size_t)(b + 24) * stride];
    }
    for (; b < nblk; b += 8) s0 += src[(size_t)b * stride];
  }
  s_part[part][e] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (part == 0 && ok) {
    double t = s_part[0][e];
#pragma unroll
    for (int q = 1; q < 8; ++q) t += s_part[q][e];
    out[pos] = t + cam;
  }
}

// NCCL fallback: the camera terms are added once, after the all-reduce of the raw sums.
__global__ void k_add_cam_terms(DeviceProblem P, ParamSet ps0, ParamSet ps1, const LmState* st, LmOptions opt,
                                double* __restrict__ out, AssembleArgs A) {
  pdl_entry();
  if (st->done) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.Q + P.NL) return;
  int r, c;
  const int pos = assemble_position(P, A, i, &r, &c);
  out[pos] += assemble_cam_term(P, st->cur ? ps1 : ps0, opt, A, A.radius_override > 0.0 ? A.radius_override : st->radius, r, c);
}

// 1/d to ~1 ulp: hardware seed + two Newton steps (no IEEE division sequence on
// the factorisation's critical path).
__device__ __forceinline__ double fast_rcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  x = fma(x, fma(-d, x, 1.0), x);
  x = fma(x, fma(-d, x, 1.0), x);
  return x;
}

// ---------------------------------------------------------------------------
// K4: back-substitution y_e = z - (V + D^2)^-1 W_s y_c, candidate board poses
// and the frame-side sums of the model cost change.  One warp per frame.
// ---------------------------------------------------------------------------
constexpr int kBacksubThreads = 256;
constexpr int kBacksubFrames = 32;      // frames per CTA: 8 lanes per frame, then lane = frame

// Round 1 ran one warp per frame and left the 6x6 solve, the 54 strided loads of the frame record
// and the model-cost sums to lane 0 (20 us at 5,000 frames, FP64 pipe 12 % busy).  Now a CTA owns
// 32 frames: (1) eight lanes per frame form w = W_s y_c (all loads of a frame in flight at once),
// (2) warp 0 takes over with lane = frame — coalesced frame records, 32 solves side by side,
// shuffle sums — and writes one partial per CTA (157 instead of 625 for k_post_eval to add).
__global__ void __launch_bounds__(kBacksubThreads)
k_backsub(DeviceProblem P, ParamSet ps0, ParamSet ps1, const LmState* st, SchurArgs A,
          const double* __restrict__ y_c, double* __restrict__ part /*[4][nblk]*/, int nblk,
          const double* __restrict__ Wg /* materialised, permuted W_s rows or nullptr */) {
  __shared__ double s_yp[224];
  __shared__ double s_w[kBacksubFrames][7];
  pdl_entry();
  if (st->done) return;
  if ((int)blockIdx.x == nblk) {
    // extra block: derived constants of the candidate cameras, whose parameters k_solve has
    // just written (what k_prep_cams does for an explicitly uploaded point)
    const ParamSet& pcand = st->cur ? ps0 : ps1;
    const int c = threadIdx.x;
    if (c < P.C) {
      CamConst cc;
      make_cam_const(pcand.cam_rt + 6 * c, pcand.intr + 9 * c, c != P.fixed_camera, cc);
      pcand.cam[c] = cc;
    }
    return;
  }
  if (Wg) {
    for (int i = threadIdx.x; i < A.NLp; i += kBacksubThreads) s_yp[i] = 0.0;
    __syncthreads();
    for (int i = threadIdx.x; i < P.NL; i += kBacksubThreads) s_yp[schur_perm(i, A.NLp)] = y_c[i];
    __syncthreads();
  }
  const int sel = st->cur;
  const ParamSet& ps = sel ? ps1 : ps0;
  const ParamSet& pc = sel ? ps0 : ps1;
  {
    // ---- w = W_s y_c, eight lanes per frame ---------------------------------------------------
    const int grp = threadIdx.x >> 3, g = threadIdx.x & 7;
    const int f = blockIdx.x * kBacksubFrames + grp;
    double w[6] = {0, 0, 0, 0, 0, 0};
    if (f < P.F) {
      if (Wg) {
        const double* Wf = Wg + (size_t)f * 6 * A.NLp;
#pragma unroll 4
        for (int c = g; c < A.NLp; c += 8) {
          const double yv = s_yp[c];
#pragma unroll
          for (int r = 0; r < 6; ++r) w[r] = fma(Wf[r * A.NLp + c], yv, w[r]);
        }
      } else {
        const int p0 = P.frame_ptr[f], nv = P.frame_ptr[f + 1] - p0;
        for (int p = 0; p < nv; ++p) {
          const int v = P.frame_views[p0 + p];
          const int m = P.view_camera[v];
          const int o0 = P.live_off[m], n = P.live_off[m + 1] - o0;
          const double* Gv = ps.G + (size_t)v * kViewStride;
          for (int k = g; k < n; k += 8) {
            const int kk = n == 13 ? k : k + 6;
            const double f_sc = A.scale_c[m * 13 + kk] * y_c[o0 + k];
#pragma unroll
            for (int r = 0; r < 6; ++r) {
              const double raw = kk < 6 ? Gv[kOffBC + r * 6 + kk] : Gv[kOffBI + r * 8 + (kk - 6)];
              w[r] += raw * f_sc;
            }
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) {
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) w[r] += __shfl_xor_sync(0xffffffffu, w[r], o);
    }
    if (g == 0) {
#pragma unroll
      for (int r = 0; r < 6; ++r) s_w[grp][r] = w[r];
    }
  }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  // ---- y_e = z - (V + D^2)^-1 w, candidate pose, model-cost sums: lane = frame ------------------
  const int lane = threadIdx.x;
  const int f = blockIdx.x * kBacksubFrames + lane;
  double lin = 0.0, quad = 0.0, dn2 = 0.0, xn2 = 0.0;
  if (f < P.F) {
    double se[6], w[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      se[i] = A.scale_e[f * 6 + i];
      w[i] = Wg ? s_w[lane][i] : s_w[lane][i] * se[i];     // the materialised rows are already scaled
    }
    double Lm[36], z[6], Vs[21], gs[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
      for (int j = 0; j <= i; ++j) Lm[i * 6 + j] = A.frame_rec[(size_t)((i * (i + 1)) / 2 + j) * A.Fpad + f];
      z[i] = A.frame_rec[(size_t)(21 + i) * A.Fpad + f];
      gs[i] = A.frame_rec[(size_t)(48 + i) * A.Fpad + f];
    }
#pragma unroll
    for (int i = 0; i < 21; ++i) Vs[i] = A.frame_rec[(size_t)(27 + i) * A.Fpad + f];
    double t[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) t[i] = w[i];
    chol6_solve(Lm, t);
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) y[i] = z[i] - t[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      lin += y[i] * gs[i];
      double row = 0.0;
#pragma unroll
      for (int j = 0; j < 6; ++j) row += (i <= j ? Vs[tri6(i, j)] : Vs[tri6(j, i)]) * y[j];
      quad += y[i] * (row + 2.0 * w[i]);
      const double delta = -y[i] * se[i];
      const double xn = ps.board_rt[f * 6 + i] + delta;
      pc.board_rt[f * 6 + i] = xn;
      dn2 += delta * delta;
      xn2 += xn * xn;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lin += __shfl_xor_sync(0xffffffffu, lin, o);
    quad += __shfl_xor_sync(0xffffffffu, quad, o);
    dn2 += __shfl_xor_sync(0xffffffffu, dn2, o);
    xn2 += __shfl_xor_sync(0xffffffffu, xn2, o);
  }
  if (lane == 0) {
    part[blockIdx.x] = lin; part[nblk + blockIdx.x] = quad;
    part[2 * nblk + blockIdx.x] = dn2; part[3 * nblk + blockIdx.x] = xn2;
  }
}


// ---------------------------------------------------------------------------
// K5: trust-region bookkeeping.  The comm record of an evaluated parameter set holds the
// per-camera sums and, at [C*kCamRec + 0..3], frame lin / quad / |delta|^2 / |x+|^2;
// gmax[0] is the frame gradient max-norm (all written by k_post_eval).
// ---------------------------------------------------------------------------
// Camera part of |x - (x - g)|_inf and |x_c|^2 for a parameter set whose comm
// record is globally summed.
// Camera-side inputs of the bookkeeping for a parameter set whose comm record is globally summed:
// total cost, camera part of |x - (x - g)|_inf and |x_c|^2.  Block-wide (every thread of the CTA
// calls it; s_red holds >= 33 doubles): the serial loops over global memory that one thread ran
// here in round 1 were most of k_post_eval's 20 us.
struct CameraSummary { double cost, gmax, xn2; };
__device__ inline CameraSummary camera_summary(const DeviceProblem& P, const ParamSet& ps, double* s_red) {
  double c = 0.0, gm = 0.0, s = 0.0;
  for (int idx = threadIdx.x; idx < P.C * 15; idx += blockDim.x) {
    const int cam = idx / 15, k = idx % 15;
    const double* U = ps.comm + cam * kCamRec;
    if (k == 0) c += U[kCamCost];
    if (k < 6 && cam == P.fixed_camera) continue;
    const double x = k < 6 ? ps.cam_rt[cam * 6 + k] : ps.intr[cam * 9 + (k - 6)];
    const double g = k < 13 ? cam_grad(U, k) : 0.0;
    const double projected = x + (-g);
    gm = fmax(gm, fabs(x - projected));
    s += x * x;
  }
  CameraSummary r;
  r.cost = block_sum(c, s_red);
  r.gmax = block_max(gm, s_red);
  r.xn2 = block_sum(s, s_red);
  return r;
}

__device__ inline void trace_push(Trace tr, LmState& st, double cost, double gmax,
                                  double step_norm, int flags) {
  const int k = st.recorded;
  if (k < tr.capacity) {
    tr.cost[k] = cost; tr.radius[k] = st.radius; tr.gmax[k] = gmax;
    tr.step_norm[k] = step_norm; tr.flags[k] = flags;
  }
  st.final_cost = k == 0 ? cost : fmin(st.final_cost, cost);
  st.recorded = k + 1;
}

// FinalizeIterationAndCheckIfMinimizerCanContinue
__device__ inline void finalize_iteration(LmState& st, const LmOptions& opt, Trace tr,
                                          int iteration, bool valid, bool successful,
                                          double cost, double step_norm) {
  if (successful) {
    st.num_successful++;
    if (st.x_cost < st.minimum_cost) st.minimum_cost = st.x_cost;
  } else {
    st.num_unsuccessful++;
  }
  trace_push(tr, st, cost, st.gmax, step_norm, (valid ? 1 : 0) | (successful ? 2 : 0));
  st.iteration = iteration;
  if (iteration >= opt.max_num_iterations) { st.termination = 1; st.done = 1; return; }
  if (!opt.disable_tolerances) {
    if (successful && st.gmax <= opt.gradient_tolerance) { st.termination = 0; st.done = 1; return; }
    if (!(st.radius > opt.min_radius)) { st.termination = 0; st.done = 1; return; }
  }
}

// IterationZero: consumes the evaluation of the initial point (set `cur`).  One CTA.
__global__ void k_init(DeviceProblem P, ParamSet ps0, ParamSet ps1, LmState* stg, LmOptions opt,
                       Trace tr) {
  __shared__ double s_red[40];
  const ParamSet& ps = stg->cur ? ps1 : ps0;
  const CameraSummary cs = camera_summary(P, ps, s_red);
  if (threadIdx.x != 0) return;
  LmState st = *stg;                      // the state lives in registers while it is worked on
  st.x_cost = cs.cost;
  st.initial_cost = st.x_cost;
  st.x_norm = sqrt(cs.xn2 + ps.comm[P.C * kCamRec + 3]);
  st.gmax = fmax(cs.gmax, ps.gmax[0]);
  st.radius = opt.initial_radius;
  st.decrease_factor = 2.0;
  st.minimum_cost = DBL_MAX;
  st.num_successful = st.num_unsuccessful = st.num_consecutive_invalid = 0;
  st.atleast_one_successful_step = 0;
  st.recorded = 0;
  st.done = 0;
  st.termination = 1;
  st.solve_ok = 1;
  finalize_iteration(st, opt, tr, 0, true, true, st.x_cost, 0.0);
  *stg = st;
}

// TrustRegionMinimizer step bookkeeping for the candidate just evaluated (one thread; `cs` is the
// block-wide camera_summary of the candidate set).
__device__ inline void decide_step(const DeviceProblem& P, const ParamSet& ps0, const ParamSet& ps1,
                                   LmState* stg, const LmOptions& opt, Trace tr, const CameraSummary& cs) {
  LmState st = *stg;                      // one vector load instead of ~40 dependent global accesses
  if (st.done) return;
  const ParamSet& pc = st.cur ? ps0 : ps1;   // candidate
  const double* extra = pc.comm + P.C * kCamRec;
  const double lin = st.cam_lin + extra[0];
  const double quad = st.cam_quad + extra[1];
  const double dn2 = st.cam_dn2 + extra[2];
  const double xn2c = st.cam_xn2 + extra[3];
  const double frame_gmax = pc.gmax[0];
  const double model_cost_change = lin - 0.5 * quad;
  st.model_cost_change = model_cost_change;
  const int iteration = st.iteration + 1;
  const bool solver_ok = st.solve_ok && isfinite(dn2) && isfinite(model_cost_change);
  const bool valid = solver_ok && model_cost_change > 0.0;
  if (!valid) {
    // HandleInvalidStep
    if (++st.num_consecutive_invalid >= opt.max_num_consecutive_invalid_steps) {
      st.termination = 2; st.done = 1; *stg = st; return;
    }
    st.radius = st.radius / st.decrease_factor;
    st.decrease_factor *= 2.0;
    finalize_iteration(st, opt, tr, iteration, false, false, st.x_cost, 0.0);
    *stg = st;
    return;
  }
  st.num_consecutive_invalid = 0;
  double candidate_cost = cs.cost;
  if (!isfinite(candidate_cost)) candidate_cost = DBL_MAX;
  st.candidate_cost = candidate_cost;
  const double step_norm = sqrt(dn2);
  st.step_norm = step_norm;
  if (!opt.disable_tolerances) {
    // Ceres <= 2.0 tests both tolerances on every valid step; >= 2.1 only once a step has been
    // successful (tscm_options.parameter_tolerance_needs_successful_step)
    const bool armed = !opt.ptol_needs_success || st.atleast_one_successful_step;
    // ParameterToleranceReached
    if (armed && step_norm <= opt.parameter_tolerance * (st.x_norm + opt.parameter_tolerance)) {
      st.termination = 0; st.done = 1; *stg = st; return;
    }
    // FunctionToleranceReached
    if (armed && fabs(st.x_cost - candidate_cost) <= opt.function_tolerance * st.x_cost) {
      st.termination = 0; st.done = 1; *stg = st; return;
    }
  }
  // IsStepSuccessful (monotonic TrustRegionStepEvaluator)
  const double relative_decrease =
      candidate_cost >= DBL_MAX ? -DBL_MAX : (st.x_cost - candidate_cost) / model_cost_change;
  if (relative_decrease > opt.min_relative_decrease) {
    // HandleSuccessfulStep: x = candidate, gradient/Jacobian are already there
    st.cur ^= 1;
    st.x_cost = candidate_cost;
    st.x_norm = sqrt(xn2c);
    st.gmax = fmax(cs.gmax, frame_gmax);
    const double t = 2.0 * relative_decrease - 1.0;
    st.radius = st.radius / fmax(1.0 / 3.0, 1.0 - t * t * t);
    st.radius = fmin(opt.max_radius, st.radius);
    st.decrease_factor = 2.0;
    st.atleast_one_successful_step = 1;
    finalize_iteration(st, opt, tr, iteration, true, true, st.x_cost, step_norm);
  } else {
    st.radius = st.radius / st.decrease_factor;
    st.decrease_factor *= 2.0;
    finalize_iteration(st, opt, tr, iteration, true, false, candidate_cost, step_norm);
  }
  *stg = st;
}

constexpr int kDecideThreads = 128;
__global__ void __launch_bounds__(kDecideThreads)
k_decide(DeviceProblem P, ParamSet ps0, ParamSet ps1, LmState* st, LmOptions opt, Trace tr) {
  __shared__ double s_red[40];
  pdl_entry();
  if (st->done) return;
  const CameraSummary cs = camera_summary(P, st->cur ? ps0 : ps1, s_red);
  if (threadIdx.x == 0) decide_step(P, ps0, ps1, st, opt, tr, cs);
}

// ---------------------------------------------------------------------------
// Post-evaluation kernel.
// ---------------------------------------------------------------------------
constexpr int kPostSplit = 4;      // blocks per camera
struct PostArgs {
  double* gmax_part;       // [fg_nblk]
  double* xn2_part;        // [fg_nblk]
  const double* bs_part;   // [4][bs_nblk]
  double* cam_sum_part;    // [C][kPostSplit][kCamRec]
  int bs_nblk, fg_nblk;
  int lanes_per_frame;     // 8 / 16 / 32 >= the largest number of views of a frame
  unsigned int* ticket;
  int initial;             // evaluation of the initial point (no step quantities)
  int decide;              // 1: single GPU, run decide_step in the tail
};

constexpr int kPostThreads = 512;

// Blocks [0, kPostSplit C): camera c's partial slots (written by k_view_blocks, one per (tile,
// camera)), a quarter of them per block, summed in a fixed order.  Blocks behind them: frame
// gradient max-norm |x - (x - g)|_inf and |x_f|^2 — a group of lanes per frame fetches the
// gradient entries of all its views at once (round 1 walked the views one dependent load after
// the other: most of that kernel's 24 us).  The LAST block to finish (threadfence + ticket) adds
// the camera quarters into the comm record, folds the frame partials and the step scalars in
// and — on a single GPU — runs the accept/reject decision itself.  Which block runs the tail
// varies; the arithmetic it performs does not, so results stay deterministic.
__global__ void __launch_bounds__(kPostThreads)
k_post_eval(DeviceProblem P, ParamSet ps0, ParamSet ps1, LmState* st, int which, LmOptions opt,
            Trace tr, PostArgs A) {
  __shared__ double s_red[kPostThreads];
  __shared__ int s_last;
  pdl_entry();
  if (which < 2 && st->done) return;
  const int sel = which >= 2 ? which - 2 : (st->cur ^ which);
  const ParamSet& ps = sel ? ps1 : ps0;
  const int t = threadIdx.x;
  const int ncam_blk = P.C * kPostSplit;
  if ((int)blockIdx.x < ncam_blk) {
    // 4 thread groups of 128 take every 4th slot of this block's quarter; 4 loads in flight
    const int c = blockIdx.x / kPostSplit, quarter = blockIdx.x % kPostSplit;
    const int e = t & 127, grp = t >> 7;
    const int b0 = P.cam_slot_begin[c], b1 = P.cam_slot_begin[c + 1];
    const int chunk = (b1 - b0 + kPostSplit - 1) / kPostSplit;
    const int se = min(b1, b0 + (quarter + 1) * chunk);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (e < kCamRec) {
      const double* part = ps.cam_part + e;
      int sl = b0 + quarter * chunk + grp;
      for (; sl + 12 < se; sl += 16) {
        s0 += part[(size_t)sl * kCamRec];
        s1 += part[(size_t)(sl + 4) * kCamRec];
        s2 += part[(size_t)(sl + 8) * kCamRec];
        s3 += part[(size_t)(sl + 12) * kCamRec];
      }
      for (; sl < se; sl += 4) s0 += part[(size_t)sl * kCamRec];
    }
    s_red[t] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (t < kCamRec)
      A.cam_sum_part[(size_t)blockIdx.x * kCamRec + t] = (s_red[t] + s_red[t + 128]) + (s_red[t + 256] + s_red[t + 384]);
  } else {
    const int fb = blockIdx.x - ncam_blk;
    const int lpf = A.lanes_per_frame, fpc = kPostThreads / lpf;
    const int f = fb * fpc + t / lpf, j = t % lpf;
    double g[6] = {0, 0, 0, 0, 0, 0};
    if (f < P.F) {
      const int p0 = P.frame_ptr[f], nv = P.frame_ptr[f + 1] - p0;
      if (j < nv) {
        const double* Gv = ps.G + (size_t)P.frame_views[p0 + j] * kViewStride + kOffBI + 7;
#pragma unroll
        for (int b = 0; b < 6; ++b) g[b] = Gv[b * 8];
      }
    }
    for (int o = lpf >> 1; o > 0; o >>= 1) {
#pragma unroll
      for (int b = 0; b < 6; ++b) g[b] += __shfl_xor_sync(0xffffffffu, g[b], o);
    }
    double gm = 0.0, xn2 = 0.0;
    if (f < P.F && j < 6) {
      const double gb = j == 0 ? g[0] : (j == 1 ? g[1] : (j == 2 ? g[2] : (j == 3 ? g[3] : (j == 4 ? g[4] : g[5]))));
      const double x = ps.board_rt[f * 6 + j];
      const double projected = x + (-gb);
      gm = fabs(x - projected);
      xn2 = x * x;
    }
    const double m = block_max(gm, s_red);
    const double s = block_sum(xn2, s_red);
    if (t == 0) { A.gmax_part[fb] = m; A.xn2_part[fb] = s; }
  }
  // ---- last block: combine ----------------------------------------------------------
  __threadfence();
  __syncthreads();
  if (t == 0) {
    const unsigned int n = atomicAdd(A.ticket, 1u);
    s_last = (n == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int i = t; i < P.C * kCamRec; i += kPostThreads) {
    const int c = i / kCamRec, e = i % kCamRec;
    const volatile double* q = A.cam_sum_part + (size_t)c * kPostSplit * kCamRec + e;
    ps.comm[i] = (q[0] + q[kCamRec]) + (q[2 * kCamRec] + q[3 * kCamRec]);
  }
  double* extra = ps.comm + P.C * kCamRec;
  for (int k = 0; k < 4; ++k) {
    double s = 0.0;
    if (!A.initial) {
      for (int i = t; i < A.bs_nblk; i += kPostThreads) s += A.bs_part[k * A.bs_nblk + i];
    } else if (k == 3) {
      const volatile double* xp = A.xn2_part;
      for (int i = t; i < A.fg_nblk; i += kPostThreads) s += xp[i];
    }
    const double tot = block_sum(s, s_red);
    if (t == 0) extra[k] = tot;
  }
  {
    double m = 0.0;
    const volatile double* gp = A.gmax_part;
    for (int i = t; i < A.fg_nblk; i += kPostThreads) m = fmax(m, gp[i]);
    const double tot = block_max(m, s_red);
    if (t == 0) ps.gmax[0] = tot;
  }
  __syncthreads();
  if (t == 0) *A.ticket = 0u;
  if (A.decide && !st->done) {
    // the candidate's comm record is complete (written above by this block)
    const CameraSummary cs = camera_summary(P, ps, s_red);
    if (t == 0) decide_step(P, ps0, ps1, st, opt, tr, cs);
  }
}

// NCCL fallback of the evaluation exchange: the record to all-reduce lives in the parameter
// set the DEVICE selects (st->cur), so it is staged through a fixed buffer.  (All-reducing both
// sets in place would sum the current point's already-global record a second time — harmless
// after an accepted step, which replaces it, but wrong after a REJECTED one.)
// dir 0: ps[sel].comm, gmax -> stage[0..n), stage[n]; dir 1: back.
__global__ void k_comm_stage(ParamSet ps0, ParamSet ps1, const LmState* st, int which, int n,
                             double* __restrict__ stage, int dir) {
  pdl_entry();
  const int sel = which >= 2 ? which - 2 : (st->cur ^ which);
  const ParamSet& ps = sel ? ps1 : ps0;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    if (dir == 0) stage[i] = ps.comm[i]; else ps.comm[i] = stage[i];
  } else if (i == n) {
    if (dir == 0) stage[n] = ps.gmax[0]; else ps.gmax[0] = stage[n];
  }
}

// DFMA throughput probe (the FP64 roofline denominator is measured, not assumed).
__global__ void k_dfma_peak(double* out, int iters, double m) {
  double a0 = threadIdx.x, a1 = 1.0, a2 = 2.0, a3 = 3.0, a4 = 4.0, a5 = 5.0, a6 = 6.0, a7 = 7.0;
  const double c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

}  // namespace tscm
