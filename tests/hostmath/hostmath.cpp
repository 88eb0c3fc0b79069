// Test-only host build of tscm_calib_b200/csrc/tscm_math.cuh: the same
// per-observation arithmetic the CUDA kernels run, compiled with g++ so the
// analytic Jacobian can be checked against the oracle's Jets without a GPU.
// Not part of the product library.
#include "../../tscm_calib_b200/csrc/tscm_math.cuh"
#include "../../include/tscm.h"
#include <cstddef>

extern "C" int hostmath_eval_jacobian(const tscm_problem* p, const double* intr, const double* cam_rt,
                                      const double* board_rt, int loss_type, double loss_scale,
                                      double* residuals, double* jacobian, double* cost) {
  using namespace tscm;
  const int K = p->corners_per_board;
  double total = 0.0;
  for (int v = 0; v < p->num_views; ++v) {
    const int m = p->view_camera[v], i = p->view_frame[v];
    CamConst cc; FrameConst fc;
    make_cam_const(cam_rt + 6 * m, intr + 9 * m, m != p->fixed_camera, cc);
    make_frame_const(board_rt + 6 * i, fc);
    for (int j = 0; j < K; ++j) {
      const size_t o = (size_t)v * K + j;
      ObsRow row;
      obs_jacobian<true, true, true>(cc, fc, p->board_xy[2 * j], p->board_xy[2 * j + 1],
                                     p->obs_xy[2 * o], p->obs_xy[2 * o + 1], row);
      double err;
      total += obs_apply_loss<0, 19>(loss_type, loss_scale, row, &err);
      double ru, rv;
      obs_residual(cc, fc, p->board_xy[2 * j], p->board_xy[2 * j + 1], p->obs_xy[2 * o],
                   p->obs_xy[2 * o + 1], ru, rv);
      if (loss_type == 0 && (ru != row.Ju[19] || rv != row.Jv[19])) return 100;
      residuals[2 * o] = row.Ju[19]; residuals[2 * o + 1] = row.Jv[19];
      double* Ju = jacobian + (o * 2) * 21; double* Jv = Ju + 21;
      for (int k = 0; k < 6; ++k) { Ju[k] = row.Ju[6 + k]; Jv[k] = row.Jv[6 + k]; }        // camera_rt
      for (int k = 0; k < 6; ++k) { Ju[6 + k] = row.Ju[k]; Jv[6 + k] = row.Jv[k]; }        // chessboard_rt
      for (int k = 0; k < 7; ++k) { Ju[12 + k] = row.Ju[12 + k]; Jv[12 + k] = row.Jv[12 + k]; }
      Ju[19] = Ju[20] = Jv[19] = Jv[20] = 0.0;
    }
  }
  *cost = total;
  return 0;
}


// Per-view blocks through the moment formulation (what the evaluation kernel does) and
// through the direct Gram of the full Jacobian rows; both written as [V][210] in the order
// BB 21 | BC 36 | BI 48 | CC 21 | CI 48 | II 36.
extern "C" int hostmath_view_blocks(const tscm_problem* p, const double* intr, const double* cam_rt,
                                    const double* board_rt, int loss_type, double loss_scale,
                                    double* via_moments, double* direct) {
  using namespace tscm;
  const int K = p->corners_per_board;
  for (int v = 0; v < p->num_views; ++v) {
    const int m = p->view_camera[v], i = p->view_frame[v];
    CamConst cc; FrameConst fc; ViewConst vc;
    make_cam_const(cam_rt + 6 * m, intr + 9 * m, m != p->fixed_camera, cc);
    make_frame_const(board_rt + 6 * i, fc);
    make_view_const(cc, fc, vc);
    double mom[108] = {0}, II[36] = {0};
    double G[20][20] = {{0}};
    for (int j = 0; j < K; ++j) {
      const size_t o = (size_t)v * K + j;
      const double X = p->board_xy[2 * j], Y = p->board_xy[2 * j + 1];
      ObsCompact oc;
      obs_compact(cc, vc, X, Y, p->obs_xy[2 * o], p->obs_xy[2 * o + 1], oc);
      double err;
      obs_compact_loss(loss_type, loss_scale, oc, &err, true);
      const double mu[3] = {X, Y, 1.0};
      double q[6];
      int e = 0;
      for (int a = 0; a < 3; ++a) for (int b = a; b < 3; ++b) q[e++] = oc.au[a] * oc.au[b] + oc.av[a] * oc.av[b];
      for (int mm = 0; mm < 3; ++mm) for (int nn = mm; nn < 3; ++nn)
        for (int k = 0; k < 6; ++k) mom[mom_pair(mm, nn) * 6 + k] += mu[mm] * mu[nn] * q[k];
      for (int mm = 0; mm < 3; ++mm) for (int k = 0; k < 3; ++k) for (int ii = 0; ii < 8; ++ii)
        mom[36 + (mm * 3 + k) * 8 + ii] += mu[mm] * (oc.au[k] * oc.ju[ii] + oc.av[k] * oc.jv[ii]);
      for (int a = 0; a < 8; ++a) for (int b = a; b < 8; ++b)
        II[tri8(a, b)] += oc.ju[a] * oc.ju[b] + oc.jv[a] * oc.jv[b];
      // direct
      ObsRow row;
      obs_jacobian<true, true, true>(cc, fc, X, Y, p->obs_xy[2 * o], p->obs_xy[2 * o + 1], row);
      obs_apply_loss<0, 19>(loss_type, loss_scale, row, &err);
      for (int a = 0; a < 20; ++a) for (int b = 0; b < 20; ++b)
        G[a][b] += row.Ju[a] * row.Ju[b] + row.Jv[a] * row.Jv[b];
    }
    double cols[12][9];
    for (int a = 0; a < 12; ++a) view_column_vectors(cc, fc, a, cols[a]);
    double E[12][12], X[12][8];
    for (int b = 0; b < 12; ++b) {
      double oe[12], ox[8];
      view_blocks_column(&cols[0][0], 9, b, [&](int k) { return mom[k]; }, oe, ox);
      for (int a = 0; a <= b; ++a) E[a][b] = oe[a];
      for (int ii = 0; ii < 8; ++ii) X[b][ii] = ox[ii];
    }
    double* out = via_moments + (size_t)v * 210;
    double* ref = direct + (size_t)v * 210;
    int k = 0;
    for (int a = 0; a < 6; ++a) for (int b = a; b < 6; ++b) { out[k] = E[a][b]; ref[k++] = G[a][b]; }
    for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) { out[k] = E[a][6 + b]; ref[k++] = G[a][6 + b]; }
    for (int a = 0; a < 6; ++a) for (int b = 0; b < 8; ++b) { out[k] = X[a][b]; ref[k++] = G[a][12 + b]; }
    for (int a = 0; a < 6; ++a) for (int b = a; b < 6; ++b) { out[k] = E[6 + a][6 + b]; ref[k++] = G[6 + a][6 + b]; }
    for (int a = 0; a < 6; ++a) for (int b = 0; b < 8; ++b) { out[k] = X[6 + a][b]; ref[k++] = G[6 + a][12 + b]; }
    for (int a = 0; a < 8; ++a) for (int b = a; b < 8; ++b) { out[k] = II[tri8(a, b)]; ref[k++] = G[12 + a][12 + b]; }
  }
  return 0;
}


// Work lists of the per-camera-pair Schur update (csrc/tscm_pair_lists.h), for the CPU tests.
#include "../../tscm_calib_b200/csrc/tscm_pair_lists.h"
extern "C" int hostmath_pair_lists(int C, int F, int V, const int* view_camera, const int* view_frame, int chunk,
                                   int cap_ent, int cap_items, int* sizes /*[2]: entries, items*/,
                                   int* ent /*[cap_ent][2]*/, int* item_range /*[cap_items][2]*/,
                                   int* pair_item /*[npairs + 1]*/, int* pair_items /*[cap_items]*/) {
  const tscm::PairLists L = tscm::build_pair_lists(C, F, V, view_camera, view_frame, chunk);
  sizes[0] = (int)L.ent.size();
  sizes[1] = (int)L.item_range.size();
  if ((int)L.ent.size() > cap_ent || (int)L.item_range.size() > cap_items) return 1;
  for (size_t k = 0; k < L.ent.size(); ++k) { ent[2 * k] = L.ent[k].view_a; ent[2 * k + 1] = L.ent[k].view_b; }
  for (size_t k = 0; k < L.item_range.size(); ++k) { item_range[2 * k] = L.item_range[k].begin; item_range[2 * k + 1] = L.item_range[k].end; }
  for (size_t k = 0; k < L.pair_item.size(); ++k) pair_item[k] = L.pair_item[k];
  for (size_t k = 0; k < L.pair_items.size(); ++k) pair_items[k] = L.pair_items[k];
  return 0;
}
