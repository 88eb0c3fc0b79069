// tscm_solve.cuh — K4: the reduced camera system, factored and solved by one CTA.
// Replaces DenseSchurComplementSolver (Eigen LLT) of the Ceres solve behind
// /root/reference/multi_calib.cpp:210,216 and TS.cpp:272,278.  The system arrives assembled
// (k_reduce_s: Schur sums + U_s + D_c^2, rhs) as 4x4 tiles, see solve_tile_id().
//
// Measured facts behind the design (B200, tools/ubench.cu, profiles/r02_ubench.txt): dependent
// DFMA 8.4 cycles, rcp.approx.f64 17.7, LDS 29, SHFL 27, __syncthreads 16-44 (64-512 threads),
// an mbarrier hop (store + arrive -> try_wait + load) 161-175 cycles; straight-line code that
// runs once costs an instruction-cache miss per 128 bytes.  The round-1 kernel (matrix packed
// in shared memory, three CTA barriers and a masked lanes-along-rows update per 4 columns)
// needed 118 k cycles at NL = 98: assemble 10.7 k, LDL^T 80.9 k, back-substitution 16 k, tail
// 10.4 k.  A first register-tile version with one mbarrier per step and kind (look-ahead) spent
// 41 k cycles in the factorisation — two 170-cycle hops per step — and 12-18 k in 33-57 KB of
// unrolled assembly code.  Hence:
//
//  * the augmented matrix [[lhs, rhs], [rhs^T, 1]] (the extra row carries the forward
//    substitution through the factorisation), padded with identity to a multiple of 4, lives in
//    REGISTERS: thread t owns 4x4 tile t (+ NT) of the lower block triangle, column-major;
//  * square-root-free blocked LDL^T, one block column per step s and ONE __syncthreads per step:
//      before the barrier  every owner of a tile (i, s) factors the 4x4 pivot block (its own
//                          up-to-date copy, see below; cubic-step reciprocals, no divisions) and
//                          eliminates its four panel rows -> panel P(i, s) in shared memory,
//                          column-major per step so that the loads of a warp are contiguous; the
//                          owner of tile (s+1, s+1) publishes that tile as it is before update s;
//      after the barrier   tile (i, j), j > s: A -= P(i, s) D^-1 P(j, s)^T, 64 FMAs on registers;
//                          the owners of column s + 1 apply the same update to their private copy
//                          of the next pivot block (40 more FMAs) — which is what saves the second
//                          barrier of a step;
//  * back-substitution by ONE warp, lane = row (r_k = y_k - sum P[j][k] x_j in registers), the
//    four residuals of a block fetched by shuffles and multiplied with the block's explicit
//    upper-triangular inverse (prepared by the pivot owner off the critical path): four
//    independent dot products instead of a 10-deep substitution chain, no barrier;
//  * meanwhile the other warps build the small tables of the tail (scaled gradient, scaled
//    camera blocks) behind a named barrier of their own.
// Same positive-definiteness test (d_j > 0 for every live column) as Eigen's LLT; a failed test
// makes the LM loop treat the step as invalid.
#pragma once

#include "tscm_kernels.cuh"

namespace tscm {

struct SolveDims {
  int nbk;        // 4x4 block rows of the augmented, padded matrix: ceil((NL + 1) / 4)
  int ntile;      // nbk (nbk + 1) / 2
  int T;          // tiles per thread (1: NL <= 123, else 4)
  int NT;         // threads
  size_t smem;    // dynamic shared memory
};

__host__ __device__ inline int solve_pan_off(int s, int nbk) {      // doubles
  return 16 * (s * (nbk - 1) - (s * (s - 1)) / 2);
}
// pivot record in shared memory: g = 1/d (4) | w10 w20 w21 w30 w31 w32 (unscaled in-block factor)
// | N (10): upper-triangular inverse of the block's P^T, row-major | l10 l20 l21 l30 l31 l32
constexpr int kPivRec = 26;
__device__ __forceinline__ int piv_widx(int a, int b) { return 4 + (a * (a - 1)) / 2 + b; }   // a > b

inline SolveDims solve_dims(int NL, int C) {
  SolveDims d;
  d.nbk = (NL + 1 + 3) / 4;
  d.ntile = d.nbk * (d.nbk + 1) / 2;
  d.T = d.ntile <= 480 ? 1 : 4;
  d.NT = std::max(128, ((d.ntile + d.T - 1) / d.T + 31) / 32 * 32) + 32;    // + the pivot warp
  const size_t doubles = (size_t)solve_pan_off(d.nbk, d.nbk)     // panels
                         + (size_t)d.nbk * 10                    // published pivot tiles
                         + (size_t)d.nbk * kPivRec               // pivot records
                         + (size_t)3 * (NL + 4)                  // x, gsv, sc
                         + (size_t)C * kCamRec                   // camera records
                         + (size_t)C * 169                       // scaled camera blocks U_s
                         + (size_t)4 * 32 + 4;                   // reduction scratch
  d.smem = doubles * sizeof(double) + (size_t)2 * (NL + 4) * sizeof(short) + 16;
  return d;
}

// 1/d to ~1 ulp: hardware seed (2^-23) + one cubic step x (1 + e + e^2), e = 1 - d x:
// three dependent FMAs after the seed
__device__ __forceinline__ double rcp_cubic(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  const double e = fma(-d, x, 1.0);
  return fma(x, fma(e, e, e), x);
}

// LDL^T of a 4x4 block given its lower triangle t = {00, 10, 11, 20, 21, 22, 30, 31, 32, 33}.
// l: unit-lower factor, i: 1 / d, w: unscaled lower entries (w_ab = l_ab d_b), d: pivots.
// The pivots go through squares (d1 = t11 - t10^2 / d0 ...) so that the serial chain is one
// reciprocal and one FMA per column.
struct Piv4 {
  double l10, l20, l21, l30, l31, l32, i0, i1, i2, i3, w21, w31, w32, d0, d1, d2, d3;
};
__device__ __forceinline__ void ldlt4(const double* t, Piv4& f) {
  const double t00 = t[0], t10 = t[1], t11 = t[2], t20 = t[3], t21 = t[4], t22 = t[5], t30 = t[6],
               t31 = t[7], t32 = t[8], t33 = t[9];
  f.d0 = t00; f.i0 = rcp_cubic(f.d0);
  f.d1 = fma(-(t10 * t10), f.i0, t11); f.i1 = rcp_cubic(f.d1);
  f.l10 = t10 * f.i0; f.l20 = t20 * f.i0; f.l30 = t30 * f.i0;
  f.w21 = fma(-t20, f.l10, t21); f.w31 = fma(-t30, f.l10, t31);
  f.d2 = fma(-(f.w21 * f.w21), f.i1, fma(-t20, f.l20, t22)); f.i2 = rcp_cubic(f.d2);
  f.l21 = f.w21 * f.i1; f.l31 = f.w31 * f.i1;
  f.w32 = fma(-f.w31, f.l21, fma(-t30, f.l20, t32));
  f.d3 = fma(-(f.w32 * f.w32), f.i2, fma(-f.w31, f.l31, fma(-t30, f.l30, t33))); f.i3 = rcp_cubic(f.d3);
  f.l32 = f.w32 * f.i2;
}

template <int T, int RMAX>   // tiles per thread; rows per lane of the back-substitution warp
__global__ void __launch_bounds__(T == 1 ? 512 : 448)
k_solve(DeviceProblem P, ParamSet ps0, ParamSet ps1, LmState* st, const double* __restrict__ M /*tiles*/,
        const double* __restrict__ scale_c, double* __restrict__ y_c /*[NL]*/, int prof) {
  pdl_entry();
  if (st->done) return;
  long long tk0 = clock64(), tk1 = 0, tk2 = 0, tk3 = 0;
  const int sel = st->cur;
  const ParamSet& ps = sel ? ps1 : ps0;     // current x
  const ParamSet& pc = sel ? ps0 : ps1;     // candidate
  extern __shared__ __align__(16) double s_mem[];
  const int NL = P.NL, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, NT = blockDim.x;
  const int nbk = (NL + 1 + 3) >> 2, ntile = (nbk * (nbk + 1)) >> 1;
  double* Lp = s_mem;                                // panels, see solve_pan_off
  double* Pt = Lp + solve_pan_off(nbk, nbk);         // [nbk][10] published pivot tiles
  double* piv = Pt + nbk * 10;                       // [nbk][kPivRec]
  double* x = piv + nbk * kPivRec;                   // [NL + 4] solution y
  double* gsv = x + NL + 4;                          // [NL + 4] scaled gradient
  double* sc = gsv + NL + 4;                         // [NL + 4] Jacobi scale
  double* s_comm = sc + NL + 4;                      // [C][kCamRec] camera records of the current point
  double* s_U = s_comm + P.C * kCamRec;              // [C][13][13] scaled camera blocks sc_i sc_j U_ij
  double* s_red = s_U + P.C * 169;                   // [32][4] + 4
  short* s_cam = reinterpret_cast<short*>(s_red + 4 * 32 + 4);   // [NL + 4]
  short* s_kk = s_cam + NL + 4;                                  // [NL + 4]
  __shared__ int s_ok;

  // ---- tile ownership and load --------------------------------------------------------------
  int bi[T], bj[T];
  bool valid[T];
  double acc[T][16];
#pragma unroll
  for (int k = 0; k < T; ++k) {
    const int id = tid + k * (NT - 32);
    valid[k] = id < ntile && tid < NT - 32;
    bi[k] = bj[k] = -1;
    if (valid[k]) {
      const double2* src = reinterpret_cast<const double2*>(M + (size_t)id * 16);
#pragma unroll
      for (int q = 0; q < 8; ++q) { const double2 v = src[q]; acc[k][2 * q] = v.x; acc[k][2 * q + 1] = v.y; }
      // column-major over the lower block triangle: column j starts at j nbk - j (j - 1) / 2
      const float D = 2.0f * nbk + 1.0f;
      int j = (int)((D - sqrtf(D * D - 8.0f * (float)id)) * 0.5f);
      j = max(0, min(j, nbk - 1));
      while (j + 1 < nbk && (j + 1) * nbk - ((j + 1) * j) / 2 <= id) ++j;
      while (j * nbk - (j * (j - 1)) / 2 > id) --j;
      bj[k] = j;
      bi[k] = j + (id - (j * nbk - (j * (j - 1)) / 2));
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[k][e] = 0.0;
    }
  }
  if (tid == 0) s_ok = 1;
  // identity padding; the augmented row keeps the rhs that k_reduce_s assembled
#pragma unroll
  for (int k = 0; k < T; ++k) {
    if (!valid[k] || 4 * bi[k] + 3 < NL) continue;       // only the last block row
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int r = 4 * bi[k] + (e >> 2), c = 4 * bj[k] + (e & 3);
      if (c >= NL) acc[k][e] = r == c ? 1.0 : 0.0;
      else if (r > NL) acc[k][e] = 0.0;
    }
  }
  // tiles are owned by threads [0, NTt); the last warp is the PIVOT warp: it owns no tile and
  // factors the next 4x4 pivot block while the tile owners update
  const int NTt = NT - 32;
  const bool pivot_warp = tid >= NTt;
  if (tid == 0) {                                    // tile 0 = (0, 0)
    double* t = Pt;
    t[0] = acc[0][0];
    t[1] = acc[0][4]; t[2] = acc[0][5];
    t[3] = acc[0][8]; t[4] = acc[0][9]; t[5] = acc[0][10];
    t[6] = acc[0][12]; t[7] = acc[0][13]; t[8] = acc[0][14]; t[9] = acc[0][15];
  }
  __syncthreads();
  // pivot record of block s: g = 1/d (4) | w10 w20 w21 w30 w31 w32 | N (10, filled after the loop)
  // | l10 l20 l21 l30 l31 l32 (unit-lower factor for the panel solve)
  auto factor_block = [&](int s, const double* t10) {
    Piv4 f;
    ldlt4(t10, f);
    if (lane == 0) {
      double* pr = piv + kPivRec * s;
      pr[0] = f.i0; pr[1] = f.i1; pr[2] = f.i2; pr[3] = f.i3;
      pr[4] = t10[1]; pr[5] = t10[3]; pr[6] = f.w21; pr[7] = t10[6]; pr[8] = f.w31; pr[9] = f.w32;
      pr[20] = f.l10; pr[21] = f.l20; pr[22] = f.l21; pr[23] = f.l30; pr[24] = f.l31; pr[25] = f.l32;
      const int j0 = 4 * s;
      const bool ok = (j0 >= NL || f.d0 > 0.0) && (j0 + 1 >= NL || f.d1 > 0.0) &&
                      (j0 + 2 >= NL || f.d2 > 0.0) && (j0 + 3 >= NL || f.d3 > 0.0);
      if (!ok) s_ok = 0;
    }
  };
  if (pivot_warp) {
    double t10[10];
#pragma unroll
    for (int e = 0; e < 10; ++e) t10[e] = Pt[e];
    factor_block(0, t10);
  }
  __syncthreads();
  tk1 = clock64();
  // ---- blocked LDL^T: two barriers per block column ---------------------------------------------
  //   P(s): owners of tiles (i, s) eliminate their four panel rows with the factors of block s
  //         -> panel s in shared memory; the owner of tile (s+1, s+1) publishes it
  //   U(s): every tile (i, j), j > s: A -= P(i, s) D^-1 P(j, s)^T (64 FMAs on registers), WHILE
  //         the pivot warp applies the same update to the published block and factors it
  // The serial reciprocal chain of a step (~45 dependent instructions, ~450 cycles for a single
  // warp) thus runs beside the updates instead of after them.
  // Element (row, col c) of panel s sits at c * nrows + (a >> 1) * (nrows / 2) + 2 * b + (a & 1),
  // row = 4 (s + 1 + b) + a: a warp's 16-byte loads of consecutive blocks are contiguous.
  for (int s = 0; s < nbk; ++s) {
    const int nrows = 4 * (nbk - 1 - s);
    double* pan = Lp + solve_pan_off(s, nbk);
    // -- P(s) --------------------------------------------------------------------------------------
#pragma unroll
    for (int k = 0; k < T; ++k) {
      if (!valid[k]) continue;
      if (bj[k] == s && bi[k] > s) {
        const double* pr = piv + kPivRec * s + 20;
        const double l10 = pr[0], l20 = pr[1], l21 = pr[2], l30 = pr[3], l31 = pr[4], l32 = pr[5];
        const int b2 = 2 * (bi[k] - s - 1);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          double p[2][4];
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const double* row = &acc[k][(2 * h + a) * 4];
            const double p0 = row[0];
            const double p1 = fma(-p0, l10, row[1]);
            const double p2 = fma(-p1, l21, fma(-p0, l20, row[2]));
            const double p3 = fma(-p2, l32, fma(-p1, l31, fma(-p0, l30, row[3])));
            p[a][0] = p0; p[a][1] = p1; p[a][2] = p2; p[a][3] = p3;
          }
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<double2*>(pan + c * nrows + h * (nrows >> 1) + b2) = make_double2(p[0][c], p[1][c]);
        }
      } else if (bj[k] == s + 1 && bi[k] == s + 1) {
        // the next pivot block as it is BEFORE update s (the pivot warp applies update s itself)
        double* t = Pt + 10 * (s + 1);
        t[0] = acc[k][0];
        t[1] = acc[k][4]; t[2] = acc[k][5];
        t[3] = acc[k][8]; t[4] = acc[k][9]; t[5] = acc[k][10];
        t[6] = acc[k][12]; t[7] = acc[k][13]; t[8] = acc[k][14]; t[9] = acc[k][15];
      }
    }
    __syncthreads();
    if (s + 1 == nbk) break;
    // -- U(s) --------------------------------------------------------------------------------------
    const double g0 = piv[kPivRec * s], g1 = piv[kPivRec * s + 1], g2 = piv[kPivRec * s + 2],
                 g3 = piv[kPivRec * s + 3];
    if (pivot_warp) {
      double t10[10];
#pragma unroll
      for (int e = 0; e < 10; ++e) t10[e] = Pt[10 * (s + 1) + e];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double2 b01 = *reinterpret_cast<const double2*>(pan + c * nrows);               // block s + 1
        const double2 b23 = *reinterpret_cast<const double2*>(pan + c * nrows + (nrows >> 1));
        const double g = c == 0 ? g0 : (c == 1 ? g1 : (c == 2 ? g2 : g3));
        const double fb[4] = {b01.x, b01.y, b23.x, b23.y};
        const double q[4] = {fb[0] * g, fb[1] * g, fb[2] * g, fb[3] * g};
        t10[0] = fma(-fb[0], q[0], t10[0]);
        t10[1] = fma(-fb[1], q[0], t10[1]); t10[2] = fma(-fb[1], q[1], t10[2]);
        t10[3] = fma(-fb[2], q[0], t10[3]); t10[4] = fma(-fb[2], q[1], t10[4]); t10[5] = fma(-fb[2], q[2], t10[5]);
        t10[6] = fma(-fb[3], q[0], t10[6]); t10[7] = fma(-fb[3], q[1], t10[7]); t10[8] = fma(-fb[3], q[2], t10[8]);
        t10[9] = fma(-fb[3], q[3], t10[9]);
      }
      factor_block(s + 1, t10);
    } else {
#pragma unroll
      for (int k = 0; k < T; ++k) {
        if (!valid[k] || bj[k] <= s) continue;
        const double* pi = pan + 2 * (bi[k] - s - 1);
        const double* pj = pan + 2 * (bj[k] - s - 1);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const double2 a01 = *reinterpret_cast<const double2*>(pi + c * nrows);
          const double2 a23 = *reinterpret_cast<const double2*>(pi + c * nrows + (nrows >> 1));
          const double2 b01 = *reinterpret_cast<const double2*>(pj + c * nrows);
          const double2 b23 = *reinterpret_cast<const double2*>(pj + c * nrows + (nrows >> 1));
          const double g = c == 0 ? g0 : (c == 1 ? g1 : (c == 2 ? g2 : g3));
          const double fa[4] = {a01.x, a01.y, a23.x, a23.y};
          const double q[4] = {b01.x * g, b01.y * g, b23.x * g, b23.y * g};
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[k][a * 4 + b] = fma(-fa[a], q[b], acc[k][a * 4 + b]);
        }
      }
    }
    __syncthreads();
  }
  // explicit inverse N of every block's P^T (upper triangular) for the back-substitution: one
  // thread per block; columns that belong to the augmented row / padding are masked
  for (int s = tid; s < nbk; s += NT) {
    double* pr = piv + kPivRec * s;
    const double i0 = pr[0], i1 = pr[1], i2 = pr[2], i3 = pr[3];
    const double w10 = pr[4], w20 = pr[5], w21 = pr[6], w30 = pr[7], w31 = pr[8], w32 = pr[9];
    const int nbc = NL - 4 * s;          // live columns of this block (>= 4: all)
    double n[10];
    n[9] = i3;
    n[7] = i2; n[8] = -i2 * (w32 * n[9]);
    n[4] = i1; n[5] = -i1 * (w21 * n[7]); n[6] = -i1 * fma(w21, n[8], w31 * n[9]);
    n[0] = i0; n[1] = -i0 * (w10 * n[4]); n[2] = -i0 * fma(w10, n[5], w20 * n[7]);
    n[3] = -i0 * fma(w10, n[6], fma(w20, n[8], w30 * n[9]));
    if (nbc < 4) { n[3] = n[6] = n[8] = n[9] = 0.0; }
    if (nbc < 3) { n[2] = n[5] = n[7] = 0.0; }
    if (nbc < 2) { n[1] = n[4] = 0.0; }
#pragma unroll
    for (int e = 0; e < 10; ++e) pr[10 + e] = n[e];
  }
  __syncthreads();
  tk2 = clock64();
  if (warp == 0) {
    // ---- back-substitution: x_k = (y_k - sum_{j > k} P[j][k] x_j) / d_k, lane = row -------------
    const int bl = nbk - 1;                     // block of the augmented row NL
    double r[RMAX];
    const double* colb[RMAX];   // column k of the panels, addressed by block row: colb + 2 jb (+ half)
    int half[RMAX];
#pragma unroll
    for (int m = 0; m < RMAX; ++m) {
      const int k = lane + 32 * m;
      r[m] = 0.0;
      colb[m] = Lp;
      half[m] = 0;
      if (k < NL) {
        const int s = k >> 2;
        const int nr = 4 * (nbk - 1 - s);
        colb[m] = Lp + solve_pan_off(s, nbk) + (k & 3) * nr - 2 * (s + 1);
        half[m] = nr >> 1;
        if (s < bl) {
          const int rr = NL - 4 * (s + 1);     // row NL inside panel s
          r[m] = Lp[solve_pan_off(s, nbk) + (k & 3) * nr + ((rr >> 1) & 1) * (nr >> 1) + 2 * (rr >> 2) + (rr & 1)];
        }
        else r[m] = piv[kPivRec * bl + piv_widx(NL & 3, k & 3)];
      }
    }
    for (int jb = bl; jb >= 0; --jb) {
      const int j0 = 4 * jb;
      if (j0 >= NL) continue;
      // operands of the fold first: they do not depend on this block's solution
      double2 p01[RMAX], p23[RMAX];
#pragma unroll
      for (int m = 0; m < RMAX; ++m) {
        const int k = lane + 32 * m;
        p01[m] = p23[m] = make_double2(0.0, 0.0);
        if (k < j0) {        // rows above the block (j0 <= NL - 1 < 32 RMAX)
          const double* col = colb[m] + 2 * jb;
          p01[m] = *reinterpret_cast<const double2*>(col);
          p23[m] = *reinterpret_cast<const double2*>(col + half[m]);
        }
      }
      const double* nn = piv + kPivRec * jb + 10;
      const double n0 = nn[0], n1 = nn[1], n2 = nn[2], n3 = nn[3], n4 = nn[4], n5 = nn[5], n6 = nn[6],
                   n7 = nn[7], n8 = nn[8], n9 = nn[9];
      const int msel = j0 >> 5, l0 = j0 & 31;
      double rv = r[0];
#pragma unroll
      for (int m = 1; m < RMAX; ++m) rv = msel == m ? r[m] : rv;
      const double r0 = __shfl_sync(0xffffffffu, rv, l0), r1 = __shfl_sync(0xffffffffu, rv, l0 + 1),
                   r2 = __shfl_sync(0xffffffffu, rv, l0 + 2), r3 = __shfl_sync(0xffffffffu, rv, l0 + 3);
      const double x0 = fma(n0, r0, n1 * r1) + fma(n2, r2, n3 * r3);
      const double x1 = fma(n4, r1, n5 * r2) + n6 * r3;
      const double x2 = fma(n7, r2, n8 * r3);
      const double x3 = n9 * r3;
      if (lane == 0) {
        const int nbc = min(4, NL - j0);
        x[j0] = x0;
        if (nbc > 1) x[j0 + 1] = x1;
        if (nbc > 2) x[j0 + 2] = x2;
        if (nbc > 3) x[j0 + 3] = x3;
      }
#pragma unroll
      for (int m = 0; m < RMAX; ++m)
        r[m] -= fma(p01[m].x, x0, p01[m].y * x1) + fma(p23[m].x, x2, p23[m].y * x3);
    }
  } else {
    // ---- meanwhile: the small tables of the tail (rolled loops, named barrier 1) -------------------
    const int t1 = tid - 32, n1 = NT - 32;
    for (int i = t1; i < NL; i += n1) {
      const int c = P.live_cam[i], kk = P.live_kk[i];
      s_cam[i] = (short)c; s_kk[i] = (short)kk;
      sc[i] = scale_c[c * 13 + kk];
    }
    for (int i = t1; i < P.C * kCamRec; i += n1) s_comm[i] = ps.comm[i];
    __syncwarp();
    asm volatile("bar.sync 1, %0;" ::"r"(n1) : "memory");
    for (int k = t1; k < NL; k += n1) gsv[k] = sc[k] * cam_grad(s_comm + s_cam[k] * kCamRec, s_kk[k]);
    for (int q = t1; q < P.C * 169; q += n1) {
      const int c = q / 169, a = (q % 169) / 13, b = q % 13;
      const int o0 = P.live_off[c], n = P.live_off[c + 1] - o0, sh = 13 - n;   // sh = 6: fixed camera
      double v = 0.0;
      if (a >= sh && b >= sh)
        v = sc[o0 + a - sh] * sc[o0 + b - sh] * (a <= b ? cam_block(s_comm + c * kCamRec, a, b)
                                                         : cam_block(s_comm + c * kCamRec, b, a));
      s_U[q] = v;
    }
  }
  __syncthreads();
  tk3 = clock64();
  // ---- y_c, candidate camera parameters, camera-side partial sums ----------------------------
  double lin = 0.0, dn2 = 0.0, quad = 0.0;
  for (int i = tid; i < NL; i += NT) {
    const double y = x[i];
    y_c[i] = y;
    lin += y * gsv[i];
    const double delta = -y * sc[i];
    dn2 += delta * delta;
    // quad: y^T U_s y restricted to this row
    const int ci = s_cam[i];
    const int o0 = P.live_off[ci], n = P.live_off[ci + 1] - o0;
    const double* Urow = s_U + ci * 169 + s_kk[i] * 13 + (13 - n);
    double row = 0.0;
    for (int t = 0; t < n; ++t) row += Urow[t] * x[o0 + t];
    quad += y * row;
  }
  // candidate camera parameters (b, c intrinsics are carried unchanged); their derived
  // constants are prepared by the extra block of k_backsub, the next kernel in the stream
  double xn2 = 0.0;
  for (int idx = tid; idx < P.C * 15; idx += NT) {
    const int c = idx / 15, k = idx % 15;   // k < 6: rt, else intrinsic k - 6
    const bool free_rt = (c != P.fixed_camera);
    const double xv = k < 6 ? ps.cam_rt[c * 6 + k] : ps.intr[c * 9 + (k - 6)];
    double xn = xv;
    if (k < 6) {
      if (free_rt) xn = xv + (-x[P.live_off[c] + k] * sc[P.live_off[c] + k]);
    } else if (k - 6 < 7) {
      const int li = P.live_off[c] + (free_rt ? 6 : 0) + (k - 6);
      xn = xv + (-x[li] * sc[li]);
    }
    if (k < 6) pc.cam_rt[c * 6 + k] = xn; else pc.intr[c * 9 + (k - 6)] = xn;
    if (k >= 6 || free_rt) xn2 += xn * xn;
  }
  // four sums at once: warp butterflies, one shared-memory hop (fixed order: deterministic)
  double v4[4] = {lin, quad, dn2, xn2};
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v4[q] += __shfl_xor_sync(0xffffffffu, v4[q], o);
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) s_red[warp * 4 + q] = v4[q];
  }
  __syncthreads();
  if (tid == 0) {
    double t4[4] = {0.0, 0.0, 0.0, 0.0};
    const int nw = NT >> 5;
    for (int w = 0; w < nw; ++w)
#pragma unroll
      for (int q = 0; q < 4; ++q) t4[q] += s_red[w * 4 + q];
    st->cam_lin = t4[0]; st->cam_quad = t4[1]; st->cam_dn2 = t4[2]; st->cam_xn2 = t4[3];
    st->solve_ok = s_ok;
    if (prof)
      printf("k_solve cycles: load %lld  ldlt %lld  backsub|tables %lld  tail %lld\n", tk1 - tk0, tk2 - tk1,
             tk3 - tk2, clock64() - tk3);
  }
}

}  // namespace tscm
