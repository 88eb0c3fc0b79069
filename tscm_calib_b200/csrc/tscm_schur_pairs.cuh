// tscm_schur_pairs.cuh — Schur complement update for SPARSE visibility (ring rigs, BASELINE
// config 4: 16 cameras, every frame seen by ~6 of them).
//
// Replaces (Ceres) SchurEliminator::Eliminate's outer products  S -= sum_f W_f^T (V_f + D^2)^-1 W_f
// where W_f only has columns for the cameras that see frame f.  The dense-row kernels
// (k_schur / k_schur_update) multiply full NL-wide rows, i.e. mostly zeros at 39 % visibility
// and NL = 202.  Here the product is organised by CAMERA PAIR: block (a, b) of S is
//     S_ab = - sum over frames seen by both a and b of  W_a^T Y_b ,   Y = (V + D^2)^-1 W
// a 13 x 13 block with a 6-deep inner dimension per frame.  k_schur_frames (compact form)
// leaves one 6 x 16 W block and one 6 x 16 Y block per VIEW in HBM (768 B each, z as an extra
// Y column so that rhs_a = -sum W_a^T z falls out of the diagonal pairs).
//
// k_schur_pairs2: one WARP per work item = (pair, <= kPairChunk consecutive common frames).
//   The two blocks of every entry stream through a per-warp 4-stage shared-memory ring with
//   cp.async (16 B per lane, 3 per entry); the half-warps take alternate entries and a lane =
//   (row pair of W_a^T, half of the Y columns) keeps 14 accumulators in registers for the whole
//   item.  No atomics: the item's 13 x 14 partial goes to its own slot.
// k_reduce_pairs: CTA per pair sums the pair's item partials in item order (4 thread groups,
//   4 loads in flight) and scatters into the packed upper S / rhs that k_reduce_s and k_solve
//   consume.  Deterministic: fixed item boundaries, fixed summation order.
//
// Bound: FP64 pipe (42 DFMA per 1.5 KB streamed = 3.6 flop/B is below the machine balance of
// 5.3 flop/B only nominally: most blocks are re-read from L2 by the other pairs of the frame).
#pragma once

#include "tscm_kernels.cuh"

namespace tscm {

#ifndef TSCM_PAIR_CHUNK
#define TSCM_PAIR_CHUNK 128
#endif
constexpr int kPairChunk = TSCM_PAIR_CHUNK;    // entries per work item
constexpr int kPairWarps = 16;
constexpr int kPairPart = 208;     // 13 rows x 16 positions per item partial

struct PairArgs {
  const double* Wv;        // [V][96]
  const double* Yv;        // [V][96]
  const int2* ent;         // [nent] (view of camera a, view of camera b) per common frame
  const int2* item_range;  // [nitems] entry range [x, y) of an item; items are ordered by the
                           // first frame they touch, so that the warps running at the same time
                           // work on the same stretch of frames and share its blocks through L2
  int nitems;
  double* part;            // [nitems][kPairPart]
  // reduction
  const int* pair_item;    // [npairs + 1] range of a pair in pair_items
  const int* pair_items;   // item ids of every pair, in frame order
  const short* pair_a;     // [npairs]
  const short* pair_b;     // [npairs]
  const short* live_off;   // [C + 1] first live column of a camera
  int npairs;
  double* Sout;            // [Q] packed upper
  double* rout;            // [NL]
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Two entries per step: the half-warps take entry 2s and 2s + 1, a lane owns TWO rows of
// W_a^T (one 16-byte load) x 7 columns = 14 accumulators, i.e. 5 LDS.128 per 14 DFMA.  (The
// first version — one entry per step, one row x 7 columns per lane, 5 loads per 7 DFMA — was
// bound by the shared-memory instruction queue: mio_throttle 22 %, LSU pipe 66-88 % busy,
// FP64 pipe 36 %, 262 us against 200 us on 40,000 frames of config 4; it is in the history.)
constexpr int kPair2Stages = 4;
constexpr size_t kPair2Smem = (size_t)kPairWarps * kPair2Stages * 384 * sizeof(double);   // 192 KB

__global__ void __launch_bounds__(kPairWarps * 32, 1)
k_schur_pairs2(const LmState* st, PairArgs A) {
  pdl_entry();
  if (st->done) return;
  extern __shared__ __align__(128) double s_ring[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* ring = s_ring + (size_t)warp * kPair2Stages * 384;
  const int half = lane >> 4, hl = lane & 15, rp = hl >> 1, h = hl & 1;
  const int nwarps = gridDim.x * kPairWarps;
  for (int item = blockIdx.x * kPairWarps + warp; item < A.nitems; item += nwarps) {
    const int2 range = A.item_range[item];
    const int e0 = range.x, n = range.y - range.x, nsteps = (n + 1) >> 1;
    double acc0[7], acc1[7];
#pragma unroll
    for (int q = 0; q < 7; ++q) { acc0[q] = 0.0; acc1[q] = 0.0; }
    int2 pcur = lane < n ? A.ent[e0 + lane] : make_int2(0, 0);
    int2 pnxt = 32 + lane < n ? A.ent[e0 + 32 + lane] : make_int2(0, 0);
    int pblock = 0;
    auto produce = [&](int s) {          // entries 2s, 2s + 1 -> stage s % kPair2Stages
      const int e = 2 * s;
      if (e < n) {
        if ((e >> 5) != pblock) {        // warp-uniform (2s and 2s + 1 share a 32-entry block)
          pcur = pnxt;
          pblock = e >> 5;
          const int q = (pblock + 1) * 32 + lane;
          pnxt = q < n ? A.ent[e0 + q] : make_int2(0, 0);
        }
        double* dst = ring + (s % kPair2Stages) * 384;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (e + u < n) {               // warp-uniform
            const int va = __shfl_sync(0xffffffffu, pcur.x, (e + u) & 31);
            const int vb = __shfl_sync(0xffffffffu, pcur.y, (e + u) & 31);
            const double* wsrc = A.Wv + (size_t)va * 96;
            const double* ysrc = A.Yv + (size_t)vb * 96;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              const int c = lane + 32 * r;
              cp_async16(dst + u * 192 + 2 * c, c < 48 ? wsrc + 2 * c : ysrc + 2 * (c - 48));
            }
          }
        }
      }
      cp_async_commit();
    };
    __syncwarp();
#pragma unroll
    for (int s = 0; s < kPair2Stages - 1; ++s) produce(s);
    for (int s = 0; s < nsteps; ++s) {
      produce(s + kPair2Stages - 1);
      cp_async_wait<kPair2Stages - 1>();
      __syncwarp();
      if (2 * s + half < n) {            // the odd tail: only the first half-warp has an entry
        const double* Wb = ring + (s % kPair2Stages) * 384 + half * 192;
        const double2* W2 = reinterpret_cast<const double2*>(Wb + 2 * rp);
        const double2* Yb = reinterpret_cast<const double2*>(Wb + 96 + h * 8);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const double2 w = W2[k * 8];
          const double2 y0 = Yb[k * 8 + 0], y1 = Yb[k * 8 + 1], y2 = Yb[k * 8 + 2], y3 = Yb[k * 8 + 3];
          acc0[0] = fma(w.x, y0.x, acc0[0]); acc1[0] = fma(w.y, y0.x, acc1[0]);
          acc0[1] = fma(w.x, y0.y, acc0[1]); acc1[1] = fma(w.y, y0.y, acc1[1]);
          acc0[2] = fma(w.x, y1.x, acc0[2]); acc1[2] = fma(w.y, y1.x, acc1[2]);
          acc0[3] = fma(w.x, y1.y, acc0[3]); acc1[3] = fma(w.y, y1.y, acc1[3]);
          acc0[4] = fma(w.x, y2.x, acc0[4]); acc1[4] = fma(w.y, y2.x, acc1[4]);
          acc0[5] = fma(w.x, y2.y, acc0[5]); acc1[5] = fma(w.y, y2.y, acc1[5]);
          acc0[6] = fma(w.x, y3.x, acc0[6]); acc1[6] = fma(w.y, y3.x, acc1[6]);
        }
      }
      __syncwarp();
    }
    cp_async_wait<0>();
    // even entries (lower half-warp) + odd entries (upper half-warp), fixed order
#pragma unroll
    for (int q = 0; q < 7; ++q) {
      acc0[q] += __shfl_xor_sync(0xffffffffu, acc0[q], 16);
      acc1[q] += __shfl_xor_sync(0xffffffffu, acc1[q], 16);
    }
    if (half == 0 && rp < 7) {
      double* out = A.part + (size_t)item * kPairPart + (2 * rp) * 16 + h * 8;
#pragma unroll
      for (int q = 0; q < 7; ++q) out[q] = acc0[q];
      if (rp < 6) {
#pragma unroll
        for (int q = 0; q < 7; ++q) out[16 + q] = acc1[q];
      }
    }
  }
}

constexpr int kPairReduceGroups = 4;

__global__ void __launch_bounds__(kPairReduceGroups * kPairPart)
k_reduce_pairs(DeviceProblem P, const LmState* st, PairArgs A) {
  pdl_entry();
  if (st->done) return;
  __shared__ double s_sum[kPairReduceGroups][kPairPart];
  const int pr = blockIdx.x;
  const int t = threadIdx.x % kPairPart, grp = threadIdx.x / kPairPart;
  const int it0 = A.pair_item[pr], it1 = A.pair_item[pr + 1];
  const double* src = A.part + t;
  const int* ids = A.pair_items;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int it = it0 + grp;
  for (; it + 3 * kPairReduceGroups < it1; it += 4 * kPairReduceGroups) {
    const int i0 = ids[it], i1 = ids[it + kPairReduceGroups], i2 = ids[it + 2 * kPairReduceGroups],
              i3 = ids[it + 3 * kPairReduceGroups];
    s0 += src[(size_t)i0 * kPairPart];
    s1 += src[(size_t)i1 * kPairPart];
    s2 += src[(size_t)i2 * kPairPart];
    s3 += src[(size_t)i3 * kPairPart];
  }
  for (; it < it1; it += kPairReduceGroups) s0 += src[(size_t)ids[it] * kPairPart];
  s_sum[grp][t] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (grp != 0) return;
  const double total = (s_sum[0][t] + s_sum[1][t]) + (s_sum[2][t] + s_sum[3][t]);
  const int i = t >> 4, pos = t & 15;
  if (pos == 7 || pos == 15) return;                  // padding of the two column halves
  const int c = pos > 7 ? pos - 1 : pos;              // inverse of pair_pos
  const int a = A.pair_a[pr], b = A.pair_b[pr];
  const int na = A.live_off[a + 1] - A.live_off[a], nb = A.live_off[b + 1] - A.live_off[b];
  const int li = na == 13 ? i : i - 6;                // the fixed camera keeps its 7 intrinsics only
  if (li < 0) return;
  const int gi = A.live_off[a] + li;
  if (c == 13) {                                      // W_a^T z
    if (a == b) A.rout[gi] = -total;
    return;
  }
  const int lj = nb == 13 ? c : c - 6;
  if (lj < 0) return;
  const int gj = A.live_off[b] + lj;
  if (gi > gj) return;                                // lower triangle of a diagonal pair
  A.Sout[gi * P.NL - (gi * (gi - 1)) / 2 + (gj - gi)] = -total;
}

}  // namespace tscm
