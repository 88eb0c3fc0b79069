// tscm_eval5.cuh — k_eval5 + k_view_blocks: persistent, mbarrier-pipelined form of the
// residual + analytic Jacobian + normal-equation pass (same outputs as k_eval4: per-view
// records ps.G and the per-(tile, camera) partial records ps.cam_part; replaces the Jet
// evaluation of multi_calib.h:146-195 / TS.h:100-131 and SchurEliminator's block products).
//
// Why (ncu, profiles/r01_k_eval4.txt and the intermediate persistent builds of this round):
// k_eval4 runs one 12-warp CTA per SM whose life is one third prologue/epilogue with an idle
// FP64 pipe; its four consumer warps (one per SM sub-partition, ~86 instructions per corner)
// are a serial bottleneck; and every attempt to keep the per-view "blocks from moments"
// epilogue inside a warp-specialised CTA made four latency-bound epilogue warps (and their
// 20-50 KB of code in a 32 KB instruction cache) the critical path.  So the pass is split:
//
// k_eval5 — one 16-warp CTA per SM stays resident and walks over tiles of 32 views (lane = view):
//   warps 0-7   CONSUMERS  two per sub-partition, eight moment slices of 19-25 accumulators:
//                 0: M, weights X^2 XY X (+cost)      1: M, weights Y^2 Y 1 (+err)
//                 2+2m+r: N[m] restricted to residual row r (u/v) + a third of the row's
//                         intrinsic Gram (u and v partial sums are added by k_view_blocks)
//   warps 8-15  PRODUCERS  two per sub-partition: projection, At, intrinsic rows, loss;
//                          producer 0 also prepares the frame constants of the NEXT tile
//   The row ring holds kE5Depth groups of 8 corners (one row per producer); every group has
//   a full and an empty mbarrier (elected-lane arrive after __syncwarp, 8 arrivals each), so
//   a warp only waits for the group it needs — no CTA-wide barrier after initialisation.  At the end of a tile every consumer
//   stores its sums straight to the moment buffer mom[tile][188][32] (coalesced 256-byte
//   rows; it stays in the 126 MB L2 for the next kernel).  Registers are re-partitioned with
//   setmaxnreg (512 threads x 128 at launch: consumers 104, producers 152).
//
// k_view_blocks — CTA = tile, 12 warps = the 12 extrinsic columns, lane = view: rebuilds the
//   12x12 / 12x8 blocks from the moments (view_blocks_column arithmetic), writes the per-view
//   record through a shared-memory transpose and the camera partial record of every camera
//   run of the tile (thread = record entry, fixed summation order).  Massively parallel and
//   latency-tolerant (several CTAs per SM) where the in-kernel epilogue was not.
#pragma once

#include "tscm_kernels.cuh"

namespace tscm {

// corners folded per consumer loop trip: the eight consumer loops and the producer loop
// share a 32 KB instruction cache (unroll 4 measured 25 % slower than 2 for that reason)
#ifndef TSCM_E5_UNROLL
#define TSCM_E5_UNROLL 2
#endif
constexpr int kE5Unroll = TSCM_E5_UNROLL;
constexpr int kE5Consumers = 8;
constexpr int kE5Producers = 8;
constexpr int kE5Threads = 32 * (kE5Consumers + kE5Producers);   // 512
#ifndef TSCM_E5_DEPTH
#define TSCM_E5_DEPTH 4
#endif
constexpr int kE5Depth = TSCM_E5_DEPTH;                      // ring slots per producer warp
// published row = 10 double2 per lane (structural zeros of the intrinsic rows are never stored):
//   u group  0 (au0 au1) | 1 (au2 ju0) | 2 (ju1 ju2) | 3 (ju3 ju4) | 4 (ju5 1/2rho)
//   v group  5 (av0 av1) | 6 (av2 jv0) | 7 (jv1 jv2) | 8 (jv3 jv4) | 9 (jv5 sqrt(s))
// ju/jv indexed by live column (e5_live_col); a consumer slice reads one group with five
// 16-byte loads
constexpr int kE5Elems = 20;
constexpr int kE5Slot = kE5Elems * 32;           // doubles per slot: [20][32]
// moment buffer of a tile: M 36 | Nu [9][6] | Nv [9][6] | IIu 21 | IIv 21 | cost | err, x 32 lanes
constexpr int kE5OffNu = 36, kE5OffNv = 90, kE5OffIIu = 144, kE5OffIIv = 165, kE5OffCost = 186,
              kE5OffErr = 187;
constexpr int kE5MomEntries = 188;
constexpr int kE5CamStage = kCamII;              // CC 21 | CI 48 staged per view by k_view_blocks
constexpr int kE5RecLd = kViewStride + 1;        // odd stride: conflict-free staging rows
constexpr int kE5Bars = 2 * kE5Depth + 4;
constexpr int kVbThreads = 384;

__host__ __device__ inline size_t e5_smem_bytes(int K, int C) {
  const size_t kpad = ((size_t)K + 7) / 8 * 8;
  size_t d = 0;
  d += kE5Bars;                                         // mbarriers (8 B each)
  d += (size_t)kE5Producers * kE5Depth * kE5Slot;       // row ring
  d += 2 * kFcElems * 32;                               // frame constants, double-buffered
  d += 5 * kpad + (kpad & 1);                           // board moments table
  return d * sizeof(double) + (size_t)C * sizeof(CamConst);
}
// all lanes have finished their shared-memory accesses -> one elected arrival
__device__ __forceinline__ void warp_arrive(unsigned long long* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}
__device__ __forceinline__ double warp_sum_xor(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// Live intrinsic columns of residual row r (0 = u, 1 = v; see e3_live): li = 0..5 ->
// u: fx cx xi lambda alpha r     v: fy cy xi lambda alpha r
__host__ __device__ constexpr int e5_live_col(int r, int li) { return li < 2 ? 2 * li + r : li + 2; }
__host__ __device__ constexpr int e5_live_idx(int i) { return i < 4 ? i / 2 : i - 2; }
__host__ __device__ constexpr int e5_live_tri(int a, int b) { return tri6(e5_live_idx(a), e5_live_idx(b)); }

// One corner folded into consumer slice W (row layout: see kE5Elems).
template <int W>
__device__ __forceinline__ void e5_consume(const double* __restrict__ row, int lane,
                                           const double* __restrict__ mu, double* __restrict__ acc) {
  const double2* __restrict__ row2 = reinterpret_cast<const double2*>(row) + lane;
  if (W < 2) {
    const double2 u0 = row2[0 * 32], u1 = row2[1 * 32], v0 = row2[5 * 32], v1 = row2[6 * 32];
    const double2 ce = row2[(W == 0 ? 4 : 9) * 32];      // .y = 1/2 rho (W = 0) or sqrt(s) (W = 1)
    const double au[3] = {u0.x, u0.y, u1.x}, av[3] = {v0.x, v0.y, v1.x};
    double q[6];
    q[0] = fma(av[0], av[0], au[0] * au[0]); q[1] = fma(av[0], av[1], au[0] * au[1]);
    q[2] = fma(av[0], av[2], au[0] * au[2]); q[3] = fma(av[1], av[1], au[1] * au[1]);
    q[4] = fma(av[1], av[2], au[1] * au[2]); q[5] = fma(av[2], av[2], au[2] * au[2]);
    // pairs 00 01 02 11 12 22 <-> weights X^2, XY, X, Y^2, Y, 1   (mu = X, Y, X^2, XY, Y^2)
    if (W == 0) {
      const double w[3] = {mu[2], mu[3], mu[0]};
#pragma unroll
      for (int pr = 0; pr < 3; ++pr)
#pragma unroll
        for (int e = 0; e < 6; ++e) acc[pr * 6 + e] = fma(w[pr], q[e], acc[pr * 6 + e]);
    } else {
      const double w[2] = {mu[4], mu[1]};
#pragma unroll
      for (int pr = 0; pr < 2; ++pr)
#pragma unroll
        for (int e = 0; e < 6; ++e) acc[pr * 6 + e] = fma(w[pr], q[e], acc[pr * 6 + e]);
#pragma unroll
      for (int e = 0; e < 6; ++e) acc[12 + e] += q[e];
    }
    acc[18] += ce.y;
  } else {
    constexpr int m = (W - 2) / 2, r = (W - 2) % 2;
    const double2 p0 = row2[(5 * r + 0) * 32], p1 = row2[(5 * r + 1) * 32], p2 = row2[(5 * r + 2) * 32],
                  p3 = row2[(5 * r + 3) * 32], p4 = row2[(5 * r + 4) * 32];
    double a[3] = {p0.x, p0.y, p1.x};
    const double j[6] = {p1.y, p2.x, p2.y, p3.x, p3.y, p4.x};
    if (m < 2) {
      const double s = mu[m];      // X or Y
#pragma unroll
      for (int k = 0; k < 3; ++k) a[k] *= s;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int li = 0; li < 6; ++li) acc[k * 6 + li] = fma(a[k], j[li], acc[k * 6 + li]);
    // this slice's third of the row's 21 live intrinsic pairs
#pragma unroll
    for (int la = 0; la < 6; ++la)
#pragma unroll
      for (int lb = la; lb < 6; ++lb) {
        const int e = tri6(la, lb);
        if (e >= 7 * m && e < 7 * m + 7) acc[18 + e - 7 * m] = fma(j[la], j[lb], acc[18 + e - 7 * m]);
      }
  }
}

// 1/sqrt(s) and 1/d to ~1 ulp from the hardware seeds (2^-23) + two Newton steps, without
// the library routines' special-case call: a call site is a scheduling barrier, and the two
// interleaved projection chains below must stay in one basic block.  Arguments here are
// squared distances / projection denominators of points in front of the rig (never 0,
// denormal or infinite for a valid evaluation; NaN/Inf propagate as they must).
__device__ __forceinline__ double e5_rsqrt(double s) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
  double hy = 0.5 * y;
  y = fma(hy, fma(-s * y, y, 1.0), y);
  hy = 0.5 * y;
  y = fma(hy, fma(-s * y, y, 1.0), y);
  return y;
}
__device__ __forceinline__ double e5_rcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  x = fma(x, fma(-d, x, 1.0), x);
  x = fma(x, fma(-d, x, 1.0), x);
  return x;
}

// obs_compact for N corners of one view at once (the arithmetic of tscm_math.cuh's
// obs_compact / ts_forward statement by statement, each statement over all N corners):
// N independent dependent-chains for the scheduler to interleave.
template <int N>
__device__ __forceinline__ void e5_obs_compact_n(const CamConst& c, const ViewConst& vc, const double* X,
                                                 const double* Y, const double* uo, const double* vo,
                                                 ObsCompact* o) {
  double x[N], y[N], z[N], rho2[N], d1[N], id1[N], z1[N], d2[N], id2[N], z2[N], d3[N], id3[N], D[N], iD[N],
      mx[N], my[N];
#define E5_ALL for (int n = 0; n < N; ++n)
#pragma unroll
  E5_ALL { x[n] = X[n] * vc.m1[0] + Y[n] * vc.m2[0] + vc.T[0]; y[n] = X[n] * vc.m1[1] + Y[n] * vc.m2[1] + vc.T[1];
           z[n] = X[n] * vc.m1[2] + Y[n] * vc.m2[2] + vc.T[2]; }
#pragma unroll
  E5_ALL rho2[n] = x[n] * x[n] + y[n] * y[n];
#pragma unroll
  E5_ALL { const double s = rho2[n] + z[n] * z[n]; id1[n] = e5_rsqrt(s); d1[n] = s * id1[n]; }
#pragma unroll
  E5_ALL z1[n] = z[n] + c.xi * d1[n];
#pragma unroll
  E5_ALL { const double s = rho2[n] + z1[n] * z1[n]; id2[n] = e5_rsqrt(s); d2[n] = s * id2[n]; }
#pragma unroll
  E5_ALL z2[n] = z1[n] + c.lam * d2[n];
#pragma unroll
  E5_ALL { const double s = rho2[n] + z2[n] * z2[n]; id3[n] = e5_rsqrt(s); d3[n] = s * id3[n]; }
#pragma unroll
  E5_ALL { D[n] = z2[n] + c.k * d3[n]; iD[n] = e5_rcp(D[n]); mx[n] = x[n] * iD[n]; my[n] = y[n] * iD[n]; }
#pragma unroll
  E5_ALL {
    o[n].ju[7] = uo[n] - (c.fx * mx[n] + c.cx);
    o[n].jv[7] = vo[n] - (c.fy * my[n] + c.cy);
    const double a1 = c.xi * id1[n];
    const double g1 = 1.0 + a1 * z[n];
    const double b1 = (1.0 + z1[n] * a1) * id2[n];
    const double c1 = z1[n] * g1 * id2[n];
    const double a2 = a1 + c.lam * b1;
    const double g2 = g1 + c.lam * c1;
    const double b2 = (1.0 + z2[n] * a2) * id3[n];
    const double c2 = z2[n] * g2 * id3[n];
    const double e = a2 + c.k * b2;
    const double h = g2 + c.k * c2;
    const double fu = c.fx * iD[n], fv = c.fy * iD[n];
    o[n].au[0] = -fu * (1.0 - mx[n] * e * x[n]); o[n].au[1] = fu * mx[n] * e * y[n]; o[n].au[2] = fu * mx[n] * h;
    o[n].av[0] = fv * my[n] * e * x[n]; o[n].av[1] = -fv * (1.0 - my[n] * e * y[n]); o[n].av[2] = fv * my[n] * h;
    const double d2xi = z1[n] * d1[n] * id2[n];
    const double z2xi = d1[n] + c.lam * d2xi;
    const double Dxi = z2xi + c.k * (z2[n] * z2xi * id3[n]);
    const double Dlam = d2[n] + c.k * (z2[n] * d2[n] * id3[n]);
    const double Dal = d3[n] * c.dk;
    const double ex = fu * mx[n], ey = fv * my[n];
    o[n].ju[0] = -mx[n]; o[n].jv[0] = 0.0;
    o[n].ju[1] = 0.0;    o[n].jv[1] = -my[n];
    o[n].ju[2] = -1.0;   o[n].jv[2] = 0.0;
    o[n].ju[3] = 0.0;    o[n].jv[3] = -1.0;
    o[n].ju[4] = ex * Dxi;  o[n].jv[4] = ey * Dxi;
    o[n].ju[5] = ex * Dlam; o[n].jv[5] = ey * Dlam;
    o[n].ju[6] = ex * Dal;  o[n].jv[6] = ey * Dal;
  }
#undef E5_ALL
}

// Cold paths kept out of line: the producer loop has to share a 32 KB instruction cache with
// eight consumer loops.
__device__ __noinline__ void e5_frame_const_to(const double* __restrict__ rt, double* __restrict__ s_dst,
                                               double* __restrict__ g_dst, int lane) {
  FrameConst fc;
  make_frame_const(rt, fc);
  const double* fp = reinterpret_cast<const double*>(&fc);
#pragma unroll
  for (int q = 0; q < kFcElems; ++q) { s_dst[q * 32 + lane] = fp[q]; g_dst[q * 32 + lane] = fp[q]; }
}
// Robust-loss weights of one observation (obs_compact_loss without touching the row):
// returns {1/2 rho(s), sqrt(rho'(s)), sqrt(s)}.
struct E5Loss { double half_rho, w, err; };
__device__ __noinline__ E5Loss e5_loss_cold(int loss_type, double loss_scale, double ru, double rv) {
  const double s = ru * ru + rv * rv;
  E5Loss r;
  r.err = sqrt(s);
  if (loss_type == 0) { r.half_rho = 0.5 * s; r.w = 1.0; return r; }
  double rho[3];
  loss_rho(loss_type, loss_scale, s, rho);
  r.w = sqrt(rho[1]);
  r.half_rho = 0.5 * rho[0];
  return r;
}
// loss_type == 0 inline (1/2 s); robust losses and the reprojection read-out out of line
__device__ __forceinline__ double e5_loss(int loss_type, double loss_scale, ObsCompact& o, double* err,
                                          bool want_err) {
  const double ru = o.ju[7], rv = o.jv[7];
  if (loss_type == 0 && !want_err) { *err = 0.0; return 0.5 * (ru * ru + rv * rv); }
  const E5Loss l = e5_loss_cold(loss_type, loss_scale, ru, rv);
  *err = want_err ? l.err : 0.0;
  if (loss_type != 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { o.au[k] *= l.w; o.av[k] *= l.w; }
#pragma unroll
    for (int k = 0; k < 8; ++k) { o.ju[k] *= l.w; o.jv[k] *= l.w; }
  }
  return l.half_rho;
}

// Work items of k_eval5.  A tile of 32 views is one item — except the tiles of the LAST, partial
// round of the persistent grid (1,250 tiles on 148 SMs: 8 full rounds + 66 tiles), which are cut
// into `split` corner ranges so that the tail round costs 1 / split of a tile time; their partial
// moments land in consecutive slots that k_view_blocks adds up.  item < n_whole: tile = item.
struct E5Items {
  int n_whole;    // tiles processed whole (a multiple of the grid size)
  int split;      // parts per remaining tile (1, 2 or 4)
  int nitems;     // n_whole + (ntiles - n_whole) * split; moment slot of an item = its index
};
inline E5Items e5_items(int ntiles, int grid, int ngroups) {
  E5Items I;
  I.n_whole = ntiles / grid * grid;
  const int rem = ntiles - I.n_whole;
  I.split = 1;
  if (rem > 0) {
    // rounds the remainder costs, in tile times: ceil(rem * split / grid) / split
    double best = 1.0;
    for (int sp = 2; sp <= 4 && sp <= ngroups; sp *= 2) {
      const double cost = (double)((rem * sp + grid - 1) / grid) / sp;
      if (cost < best - 1e-9) { best = cost; I.split = sp; }
    }
  }
  I.nitems = I.n_whole + rem * I.split;
  return I;
}
struct E5Item { int tile, g0, g1; };
__device__ __forceinline__ E5Item e5_item(const E5Items& I, int item, int ngroups) {
  E5Item it;
  if (item < I.n_whole) { it.tile = item; it.g0 = 0; it.g1 = ngroups; return it; }
  const int r = item - I.n_whole, p = r % I.split;
  it.tile = I.n_whole + r / I.split;
  it.g0 = p * ngroups / I.split;
  it.g1 = (p + 1) * ngroups / I.split;
  return it;
}

template <int W>
__device__ __forceinline__ void e5_consumer(const double* __restrict__ s_ring,
                                            const double* __restrict__ s_mu, double* __restrict__ mom_g,
                                            unsigned long long* full, unsigned long long* empty,
                                            int lane, E5Items items, int ngroups) {
  constexpr int NA = W < 2 ? 19 : 25;
  double acc[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) acc[i] = 0.0;
  unsigned k = 0;
  for (int item = blockIdx.x; item < items.nitems; item += gridDim.x) {
    const E5Item it = e5_item(items, item, ngroups);
    for (int g = it.g0; g < it.g1; ++g, ++k) {
      const unsigned slot = k % kE5Depth, ph = (k / kE5Depth) & 1u;
      const double* mu = s_mu + 5 * g * kE5Producers;
      mbar_wait(full + slot, ph);            // all eight rows of the group are published
#pragma unroll kE5Unroll
      for (int o = 0; o < kE5Producers; ++o)
        e5_consume<W>(s_ring + (size_t)(o * kE5Depth + slot) * kE5Slot, lane, mu + 5 * o, acc);
      warp_arrive(empty + slot, lane);
    }
    // the item's sums go straight to its slot of the moment buffer (256-byte rows, lane = view)
    double* dst = mom_g + (size_t)item * (kE5MomEntries * 32) + lane;
    if (W < 2) {
#pragma unroll
      for (int i = 0; i < 18; ++i) { dst[(18 * W + i) * 32] = acc[i]; acc[i] = 0.0; }
      dst[(W == 0 ? kE5OffCost : kE5OffErr) * 32] = acc[18];
      acc[18] = 0.0;
    } else {
      constexpr int m = (W - 2) / 2, r = (W - 2) % 2;
      constexpr int offN = (r ? kE5OffNv : kE5OffNu) + m * 18, offII = (r ? kE5OffIIv : kE5OffIIu) + 7 * m;
#pragma unroll
      for (int i = 0; i < 18; ++i) { dst[(offN + i) * 32] = acc[i]; acc[i] = 0.0; }
#pragma unroll
      for (int i = 0; i < 7; ++i) { dst[(offII + i) * 32] = acc[18 + i]; acc[18 + i] = 0.0; }
    }
  }
}

__global__ void __launch_bounds__(kE5Threads, 1)
k_eval5(DeviceProblem P, ParamSet ps0, ParamSet ps1, const LmState* st, int which, LmOptions opt,
        int want_err, double* __restrict__ mom_g, double* __restrict__ fc_g, E5Items items) {
  pdl_entry();
  if (which < 2 && st->done) return;
  const int sel = which >= 2 ? which - 2 : (st->cur ^ which);
  const ParamSet& ps = sel ? ps1 : ps0;
  extern __shared__ __align__(16) double s_mem[];
  const int Kpad = (P.K + 7) / 8 * 8;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(s_mem);
  unsigned long long* full = bars;                                   // [D] group published (8 producers)
  unsigned long long* empty = bars + kE5Depth;                       // [D] group consumed (8 consumers)
  unsigned long long* fc_full = empty + kE5Depth;                    // [2]
  unsigned long long* fc_empty = fc_full + 2;                        // [2]
  double* s_ring = s_mem + kE5Bars;
  double* s_fc = s_ring + kE5Producers * kE5Depth * kE5Slot;         // [2][27][32]
  double* s_mu = s_fc + 2 * kFcElems * 32;                           // [Kpad][5]
  CamConst* s_cam = reinterpret_cast<CamConst*>(s_mu + 5 * Kpad + (Kpad & 1));

  for (int j = threadIdx.x; j < Kpad; j += blockDim.x) {
    const double X = j < P.K ? P.board_xy[2 * j] : 0.0, Y = j < P.K ? P.board_xy[2 * j + 1] : 0.0;
    s_mu[5 * j] = X; s_mu[5 * j + 1] = Y; s_mu[5 * j + 2] = X * X; s_mu[5 * j + 3] = X * Y; s_mu[5 * j + 4] = Y * Y;
  }
  {
    const int n = P.C * (int)(sizeof(CamConst) / sizeof(double));
    const double* src = reinterpret_cast<const double*>(ps.cam);
    double* dst = reinterpret_cast<double*>(s_cam);
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kE5Depth; ++i) { mbar_init(full + i, kE5Producers); mbar_init(empty + i, kE5Consumers); }
    mbar_init(fc_full, 1); mbar_init(fc_full + 1, 1);
    mbar_init(fc_empty, kE5Producers); mbar_init(fc_empty + 1, kE5Producers);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = (P.V + 31) / 32;
  const int ngroups = (P.K + kE5Producers - 1) / kE5Producers;

  if (warp < kE5Consumers) {
    // ------------------------------ consumers -----------------------------------
    asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
#define TSCM_E5_CONSUMER(W) e5_consumer<W>(s_ring, s_mu, mom_g, full, empty, lane, items, ngroups)
    // sub-partition = warp % 4: one M slice or one row-u slice next to one row-v slice
    switch (warp) {
      case 0: TSCM_E5_CONSUMER(0); break;
      case 1: TSCM_E5_CONSUMER(1); break;
      case 2: TSCM_E5_CONSUMER(2); break;
      case 3: TSCM_E5_CONSUMER(3); break;
      case 4: TSCM_E5_CONSUMER(4); break;
      case 5: TSCM_E5_CONSUMER(5); break;
      case 6: TSCM_E5_CONSUMER(6); break;
      default: TSCM_E5_CONSUMER(7); break;
    }
#undef TSCM_E5_CONSUMER
  } else {
    // ------------------------------ producers -----------------------------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    const int p = warp - kE5Consumers;
    double* my_ring = s_ring + (size_t)p * kE5Depth * kE5Slot;
    unsigned k = 0;
    int it = 0;
    // frame constants of a tile: computed once by producer 0 (one tile ahead), shared with
    // the other producers through shared memory and with k_view_blocks through fc_g
    auto publish_fc = [&](int tile, int buf, unsigned use) {
      // use = how many times this buffer has been filled before; every producer has read the
      // previous content (fc_empty) before it is overwritten
      mbar_wait(fc_empty + buf, (use & 1u) ^ 1u);
      const int v = min(tile * 32 + lane, P.V - 1);
      e5_frame_const_to(ps.board_rt + 6 * P.view_frame[v], s_fc + buf * (kFcElems * 32),
                        fc_g + (size_t)tile * (kFcElems * 32), lane);
      warp_arrive(fc_full + buf, lane);
    };
    if (p == 0 && (int)blockIdx.x < items.nitems) publish_fc(e5_item(items, blockIdx.x, ngroups).tile, 0, 0u);
    // Two corners (groups g, g+1) are evaluated together: the projection is one long
    // dependent FP64 chain (3 rsqrt + a reciprocal), two independent chains interleave in
    // one warp.  The observations of the NEXT pair (also across a tile boundary) are
    // requested before the current pair is evaluated.
    auto fetch = [&](int tile, int j) {
      const int vv = tile * 32 + lane;
      return (j < P.K && tile < ntiles && vv < P.V) ? P.obsT[(size_t)j * P.Vpad + vv] : make_double2(0.0, 0.0);
    };
    auto store_row = [&](double* mine, const ObsCompact& o, double half_rho, double err, bool ok) {
      double2* m2 = reinterpret_cast<double2*>(mine) + lane;
      const double ju[6] = {o.ju[e5_live_col(0, 0)], o.ju[e5_live_col(0, 1)], o.ju[e5_live_col(0, 2)],
                            o.ju[e5_live_col(0, 3)], o.ju[e5_live_col(0, 4)], o.ju[e5_live_col(0, 5)]};
      const double jv[6] = {o.jv[e5_live_col(1, 0)], o.jv[e5_live_col(1, 1)], o.jv[e5_live_col(1, 2)],
                            o.jv[e5_live_col(1, 3)], o.jv[e5_live_col(1, 4)], o.jv[e5_live_col(1, 5)]};
      double2 pr[10] = {make_double2(o.au[0], o.au[1]), make_double2(o.au[2], ju[0]), make_double2(ju[1], ju[2]),
                        make_double2(ju[3], ju[4]), make_double2(ju[5], half_rho),
                        make_double2(o.av[0], o.av[1]), make_double2(o.av[2], jv[0]), make_double2(jv[1], jv[2]),
                        make_double2(jv[3], jv[4]), make_double2(jv[5], err)};
      if (!__all_sync(0xffffffffu, ok)) {      // last tile / padding corner: zero rows
#pragma unroll
        for (int q = 0; q < 10; ++q) if (!ok) pr[q] = make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int q = 0; q < 10; ++q) m2[q * 32] = pr[q];
    };
    E5Item cur = e5_item(items, min((int)blockIdx.x, items.nitems - 1), ngroups);
    double2 uvA = fetch(cur.tile, cur.g0 * kE5Producers + p), uvB = fetch(cur.tile, (cur.g0 + 1) * kE5Producers + p);
    for (int item = blockIdx.x; item < items.nitems; item += gridDim.x, ++it) {
      cur = e5_item(items, item, ngroups);
      const int tile = cur.tile;
      const bool has_next = item + (int)gridDim.x < items.nitems;
      const E5Item nxt = e5_item(items, has_next ? item + (int)gridDim.x : item, ngroups);
      const int v0 = tile * 32 + lane;
      const bool valid = v0 < P.V;
      const int v = valid ? v0 : P.V - 1;
      const CamConst& cc = s_cam[P.view_camera[v]];
      ViewConst vc;
      mbar_wait(fc_full + (it & 1), (unsigned)(it >> 1) & 1u);
      {
        FrameConst fc;
        double* fp = reinterpret_cast<double*>(&fc);
        const double* src = s_fc + (it & 1) * (kFcElems * 32);
#pragma unroll
        for (int e = 0; e < kFcElems; ++e) fp[e] = src[e * 32 + lane];
        make_view_const(cc, fc, vc);
        warp_arrive(fc_empty + (it & 1), lane);
      }
      for (int g = cur.g0; g < cur.g1; g += 2) {
        const bool two = g + 1 < cur.g1;
        const int j0 = g * kE5Producers + p, j1 = j0 + kE5Producers;
        const double2 uv0 = uvA, uv1 = two ? uvB : make_double2(0.0, 0.0);
        // the observations of the next pair — of this item or of the next one — are requested now
        const int nt = has_next ? nxt.tile : ntiles;          // ntiles: fetch() returns zeros
        const int nj = nxt.g0 * kE5Producers + p;
        uvA = g + 2 < cur.g1 ? fetch(tile, j0 + 2 * kE5Producers) : fetch(nt, nj);
        if (g + 2 < cur.g1) { if (g + 3 < cur.g1) uvB = fetch(tile, j1 + 2 * kE5Producers); }
        else uvB = fetch(nt, nj + kE5Producers);
        const bool ok0 = j0 < P.K && valid, ok1 = two && j1 < P.K && valid;
        const int c0 = min(j0, P.K - 1), c1 = min(j1, P.K - 1);
        ObsCompact o[2];
        double e0, e1, r0, r1 = 0.0;
        const bool we = (want_err & 1) != 0;
        {
          // both evaluated unconditionally (clamped inputs; the second one is a dummy in the
          // odd last group) so that the two chains share one basic block; rows of padding
          // corners / views are zeroed on the way out
          const double Xs[2] = {s_mu[5 * c0], s_mu[5 * c1]}, Ys[2] = {s_mu[5 * c0 + 1], s_mu[5 * c1 + 1]};
          const double us[2] = {uv0.x, uv1.x}, vs[2] = {uv0.y, uv1.y};
          e5_obs_compact_n<2>(cc, vc, Xs, Ys, us, vs, o);
          r0 = e5_loss(opt.loss_type, opt.loss_scale, o[0], &e0, we);
          r1 = e5_loss(opt.loss_type, opt.loss_scale, o[1], &e1, we);
        }
        // the rows are complete in registers before their slots are claimed
        {
          const unsigned slot = k % kE5Depth, ph = (k / kE5Depth) & 1u;
          mbar_wait(empty + slot, ph ^ 1u);
          store_row(my_ring + slot * kE5Slot, o[0], r0, e0, ok0);
          warp_arrive(full + slot, lane);
          ++k;
        }
        if (two) {
          const unsigned slot = k % kE5Depth, ph = (k / kE5Depth) & 1u;
          mbar_wait(empty + slot, ph ^ 1u);
          store_row(my_ring + slot * kE5Slot, o[1], r1, e1, ok1);
          warp_arrive(full + slot, lane);
          ++k;
        }
        if (p == 0 && g == cur.g0 && has_next) publish_fc(nxt.tile, (it + 1) & 1, (unsigned)(it + 1) >> 1);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// k_view_blocks: per-view blocks from the moments.  CTA = tile of 32 views, warp = extrinsic
// column b (0..2 w_f, 3..5 t_f, 6..8 w_c, 9..11 t_c), lane = view.  The tile's moments and
// frame constants arrive by TMA bulk copies (three instructions, one mbarrier); once every
// warp has folded them into h_b / X[b][:] the M | N area is re-used as the staging of the
// per-view records, which leave with one TMA bulk store.  Because a warp owns one column, the
// zero pattern of its coefficient vector (w_f: c_2 = 0; t_f, t_c: c_0 = c_1 = 0) is a
// warp-uniform compile-time range [QLO, QHI) of live coefficients: 40 % fewer FMAs and
// shared-memory loads than the dense 9-term products.  Camera-part sums (CC, CI, II, cost,
// err) are reduced over the views of each camera run with warp shuffles (fixed order).
// ---------------------------------------------------------------------------
constexpr int kVbMN = kE5OffIIu;                       // 144 entries: M | Nu | Nv
constexpr int kVbTail = kE5MomEntries - kE5OffIIu;     // 44 entries: IIu | IIv | cost | err
constexpr int kVbAlias = kVbMN * 32;                   // doubles: M | N, later records [32][106]
static_assert(kVbAlias >= 32 * kViewStride, "record staging must fit into the M | N area");

__host__ __device__ inline size_t vb_smem_bytes() {
  return (size_t)(2 + kVbAlias + kVbTail * 32 + 108 * 32 + kFcElems * 32) * sizeof(double) + 72 * sizeof(short);
}

// h_b = M c_b and X[b][:] = c_b^T N for a column whose live coefficients are [QLO, QHI)
template <int QLO, int QHI>
__device__ __forceinline__ void vb_fold(const double* __restrict__ mom_l, const double* cb, double* h,
                                        double* ox) {
#pragma unroll
  for (int q = 0; q < 9; ++q) h[q] = 0.0;
#pragma unroll
  for (int m = 0; m < 3; ++m)
#pragma unroll
    for (int n = m; n < 3; ++n) {
      const bool use_n = 3 * n >= QLO && 3 * n < QHI;            // h_m += M[m][n] c_n
      const bool use_m = m != n && 3 * m >= QLO && 3 * m < QHI;  // h_n += M[n][m] c_m
      if (!use_n && !use_m) continue;
      double q6[6];
#pragma unroll
      for (int t = 0; t < 6; ++t) q6[t] = mom_l[(mom_pair(m, n) * 6 + t) * 32];
      if (use_n) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
          h[3 * m + i] += sym3(q6, i, 0) * cb[3 * n] + sym3(q6, i, 1) * cb[3 * n + 1] + sym3(q6, i, 2) * cb[3 * n + 2];
      }
      if (use_m) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
          h[3 * n + i] += sym3(q6, i, 0) * cb[3 * m] + sym3(q6, i, 1) * cb[3 * m + 1] + sym3(q6, i, 2) * cb[3 * m + 2];
      }
    }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    // fx fy cx cy come from one residual row only, xi lambda alpha r from both
    const int li = e5_live_idx(i);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int q = QLO; q < QHI; ++q) {
      double nv = 0.0;
      if (e3_live(0, i)) nv = mom_l[(kE5OffNu + q * 6 + li) * 32];
      if (e3_live(1, i)) nv += mom_l[(kE5OffNv + q * 6 + li) * 32];
      if (q % 3 == 0) s0 = fma(cb[q], nv, s0);
      else if (q % 3 == 1) s1 = fma(cb[q], nv, s1);
      else s2 = fma(cb[q], nv, s2);
    }
    ox[i] = (s0 + s1) + s2;
  }
}
// c_a . h over the live coefficients [QLO, QHI) of row a
template <int QLO, int QHI>
__device__ __forceinline__ double vb_dot(const double* __restrict__ ca, const double* h) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
  for (int q = QLO; q < QHI; ++q) {
    if (q % 3 == 0) s0 = fma(ca[q * 32], h[q], s0);
    else if (q % 3 == 1) s1 = fma(ca[q * 32], h[q], s1);
    else s2 = fma(ca[q * 32], h[q], s2);
  }
  return (s0 + s1) + s2;
}

__global__ void __launch_bounds__(kVbThreads, 2)
k_view_blocks(DeviceProblem P, ParamSet ps0, ParamSet ps1, const LmState* st, int which,
              const double* __restrict__ mom_g, const double* __restrict__ fc_g, E5Items items) {
  pdl_entry();
  if (which < 2 && st->done) return;
  const int sel = which >= 2 ? which - 2 : (st->cur ^ which);
  const ParamSet& ps = sel ? ps1 : ps0;
  extern __shared__ __align__(128) double s_mem[];
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(s_mem);
  double* s_mn = s_mem + 2;                            // [144][32] M | Nu | Nv, later:
  double* s_recs = s_mn;                               //   [32][106] per-view records
  double* s_tail = s_mn + kVbAlias;                    // [44][32] IIu | IIv | cost | err
  double* s_colv = s_tail + kVbTail * 32;              // [108][32]
  double* s_fc = s_colv + 108 * 32;                    // [27][32]
  short* s_iisrc = reinterpret_cast<short*>(s_fc + kFcElems * 32);   // [36][2]
  // the split tiles of the last k_eval5 round (a little more work: their parts are added below)
  // go first, not into the tail of this grid
  const int ntiles_all = (P.V + 31) / 32;
  const int tile = ((int)blockIdx.x + items.n_whole) % ntiles_all;
  const int b = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int slot = tile < items.n_whole ? tile : items.n_whole + (tile - items.n_whole) * items.split;
    const double* gm = mom_g + (size_t)slot * (kE5MomEntries * 32);
    mbar_expect_tx(bar, (unsigned)((kE5MomEntries + kFcElems) * 32 * sizeof(double)));
    tma_bulk_g2s(s_mn, gm, kVbMN * 32 * sizeof(double), bar);
    tma_bulk_g2s(s_tail, gm + kVbMN * 32, kVbTail * 32 * sizeof(double), bar);
    tma_bulk_g2s(s_fc, fc_g + (size_t)tile * (kFcElems * 32), kFcElems * 32 * sizeof(double), bar);
  }
  const int v0 = tile * 32 + lane;
  const bool valid = v0 < P.V;
  const int my_cam = P.view_camera[valid ? v0 : P.V - 1];
  if (b == 1 && lane < 8) {
    // source entries (tail area) of intrinsic Gram entry (a <= c): u part, v part, -1 = none
    const int a = lane;
    for (int c = a; c < 8; ++c) {
      const int t = e5_live_tri(a, c);
      s_iisrc[2 * tri8(a, c)] = (short)(e3_live(0, a) && e3_live(0, c) ? t : -1);
      s_iisrc[2 * tri8(a, c) + 1] = (short)(e3_live(1, a) && e3_live(1, c) ? 21 + t : -1);
    }
  }
  // camera constants this column needs (global, L2-resident) are requested before the wait
  // for the bulk copies: R_c for w_f / t_f, dR_c/dw_(b-6) for w_c
  const CamConst& cam = ps.cam[my_cam];
  double Rm[9];
  const double* Rsrc = (b >= 6 && b < 9) ? cam.dR[b - 6] : cam.R;
#pragma unroll
  for (int q = 0; q < 9; ++q) Rm[q] = Rsrc[q];
  const int free_rt = cam.free_rt;
  const int slot0 = P.blk_slot[tile];
  __syncthreads();          // barrier initialised before anyone polls it
  mbar_wait(bar, 0);
  if (tile >= items.n_whole && items.split > 1) {
    // a tile of the last round: its moments arrived as `split` partial sums over corner ranges
    const double* gm = mom_g + (size_t)(items.n_whole + (tile - items.n_whole) * items.split) * (kE5MomEntries * 32);
    constexpr int kN = kE5MomEntries * 32, kPer = (kN + kVbThreads - 1) / kVbThreads;
    for (int q = 1; q < items.split; ++q) {
      double v[kPer];                           // all loads of a part in flight at once
#pragma unroll
      for (int u = 0; u < kPer; ++u) {
        const int i = threadIdx.x + u * kVbThreads;
        v[u] = i < kN ? gm[(size_t)q * kN + i] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < kPer; ++u) {
        const int i = threadIdx.x + u * kVbThreads;
        if (i < kN) s_mn[i] += v[u];
      }
    }
    __syncthreads();
  }
  double cb[9], h[9], ox[8];
  {
    // view_column_vectors(cam, fc, b) with the frame constants read from shared memory at
    // their dynamic offsets (no register array is indexed dynamically).  FrameConst layout:
    // r1 0..2 | r2 3..5 | t 6..8 | d1[a][j] 9+3a+j | d2[a][j] 18+3a+j
    const double* f = s_fc + lane;
#pragma unroll
    for (int q = 0; q < 9; ++q) cb[q] = 0.0;
    if (b < 3) {                     // w_f[b]:  X R_c d1[b] + Y R_c d2[b]
      const double d1x = f[(9 + 3 * b) * 32], d1y = f[(10 + 3 * b) * 32], d1z = f[(11 + 3 * b) * 32];
      const double d2x = f[(18 + 3 * b) * 32], d2y = f[(19 + 3 * b) * 32], d2z = f[(20 + 3 * b) * 32];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        cb[i] = Rm[3 * i] * d1x + Rm[3 * i + 1] * d1y + Rm[3 * i + 2] * d1z;
        cb[3 + i] = Rm[3 * i] * d2x + Rm[3 * i + 1] * d2y + Rm[3 * i + 2] * d2z;
      }
    } else if (b < 6) {              // t_f[k]:  R_c[:, k]
      const int k = b - 3;
#pragma unroll
      for (int i = 0; i < 3; ++i) cb[6 + i] = k == 0 ? Rm[3 * i] : (k == 1 ? Rm[3 * i + 1] : Rm[3 * i + 2]);
    } else if (free_rt) {
      if (b < 9) {                   // w_c[k]:  dR_c[k] (X r1 + Y r2 + t_f)
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          const double x = f[(3 * m) * 32], y = f[(3 * m + 1) * 32], z = f[(3 * m + 2) * 32];
#pragma unroll
          for (int i = 0; i < 3; ++i) cb[3 * m + i] = Rm[3 * i] * x + Rm[3 * i + 1] * y + Rm[3 * i + 2] * z;
        }
      } else {                       // t_c[k]:  e_k
#pragma unroll
        for (int i = 0; i < 3; ++i) cb[6 + i] = (i == b - 9) ? 1.0 : 0.0;
      }
    }
#pragma unroll
    for (int q = 0; q < 9; ++q) s_colv[(b * 9 + q) * 32 + lane] = cb[q];
  }
  {
    const double* mom_l = s_mn + lane;
    if (b < 3) vb_fold<0, 6>(mom_l, cb, h, ox);
    else if (b >= 6 && b < 9) vb_fold<0, 9>(mom_l, cb, h, ox);
    else vb_fold<6, 9>(mom_l, cb, h, ox);
  }
  __syncthreads();          // M | N consumed by every warp, column vectors published
  double* rec = s_recs + lane * kViewStride;
  const double* colv_l = s_colv + lane;
  // camera runs of the tile (views are camera-major: a run = consecutive views of one camera)
  const unsigned all_valid = __ballot_sync(0xffffffffu, valid);
  if (b < 6) {
    // BI | BB: rows a <= b < 6
#pragma unroll
    for (int i = 0; i < 8; ++i) rec[kOffBI + b * 8 + i] = ox[i];
#pragma unroll 1
    for (int a = 0; a <= b; ++a) {
      const double* ca = colv_l + a * (9 * 32);
      const double sum = a < 3 ? vb_dot<0, 6>(ca, h) : vb_dot<6, 9>(ca, h);
      rec[kOffBB + a * 6 - (a * (a - 1)) / 2 + (b - a)] = sum;
    }
    if (b == 0) rec[kViewStride - 1] = 0.0;
    // II | cost | err: 38 record entries, 7 per warp; u and v partial sums are added here
    unsigned remaining = all_valid;
    int slot = slot0;
    while (remaining) {
      const int cam = __shfl_sync(0xffffffffu, my_cam, __ffs(remaining) - 1);
      const bool member = valid && my_cam == cam;
      remaining &= ~__ballot_sync(0xffffffffu, member);
      double* dst = ps.cam_part + (size_t)slot * kCamRec + kCamII;
#pragma unroll 1
      for (int q = 7 * b; q < min(7 * b + 7, 38); ++q) {
        int iu, iv = -1;
        if (q < 36) { iu = s_iisrc[2 * q]; iv = s_iisrc[2 * q + 1]; }
        else iu = q + 6;                                  // cost 42, err 43 of the tail area
        double x = 0.0;
        if (member) {
          if (iu >= 0) x = s_tail[iu * 32 + lane];
          if (iv >= 0) x += s_tail[iv * 32 + lane];
        }
        const double sum = warp_sum_xor(x);
        if (lane == 0) dst[q] = sum;
      }
      ++slot;
    }
  } else {
    // BC: rows a < 6;  CC: rows 6 <= a <= b;  CI: X[b][:]
    const int cc = b - 6;
#pragma unroll 1
    for (int a = 0; a < 6; ++a) {
      const double* ca = colv_l + a * (9 * 32);
      rec[kOffBC + a * 6 + cc] = a < 3 ? vb_dot<0, 6>(ca, h) : vb_dot<6, 9>(ca, h);
    }
    double ccv[6];
#pragma unroll
    for (int a = 6; a < 12; ++a) {
      const double* ca = colv_l + a * (9 * 32);
      ccv[a - 6] = a > b ? 0.0 : (a < 9 ? vb_dot<0, 9>(ca, h) : vb_dot<6, 9>(ca, h));
    }
    unsigned remaining = all_valid;
    int slot = slot0;
    while (remaining) {
      const int cam = __shfl_sync(0xffffffffu, my_cam, __ffs(remaining) - 1);
      const bool member = valid && my_cam == cam;
      remaining &= ~__ballot_sync(0xffffffffu, member);
      double* dst = ps.cam_part + (size_t)slot * kCamRec;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        if (a <= cc) {      // warp-uniform
          const double sum = warp_sum_xor(member ? ccv[a] : 0.0);
          if (lane == 0) dst[kCamCC + a * 6 - (a * (a - 1)) / 2 + (cc - a)] = sum;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const double sum = warp_sum_xor(member ? ox[i] : 0.0);
        if (lane == 0) dst[kCamCI + cc * 8 + i] = sum;
      }
      ++slot;
    }
  }
  // records -> global with one bulk store
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nv = min(32, P.V - tile * 32);
    double* dst = ps.G + (size_t)tile * 32 * kViewStride;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                 "r"(smem_u32(s_recs)), "r"((unsigned)(nv * kViewStride * sizeof(double)))
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

}  // namespace tscm
