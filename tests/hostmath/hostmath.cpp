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
