"""Generates the committed golden fixtures in tests/golden/.

The reference has no tests or golden vectors for this path and cannot be built
or imported here (C++ needing Ceres/Eigen/OpenCV), so these vectors come from
the INDEPENDENT numpy restatement (oracle/numpy_ref.py: vectorised autodiff,
dense normal equations, no Schur complement) and from 50-digit mpmath central
differences — not from the C++ oracle they are used to pin.  PARITY UNPINNED
with respect to Ceres itself.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tscm_calib_b200 import synth  # noqa: E402
from oracle import numpy_ref as nr  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def save_solve(name, sp, **kw):
    P = nr.Problem.from_arrays(sp.problem)
    i, c, b, S = nr.solve(P, sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt, **kw)
    r, J = nr.evaluate(P, sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        num_cameras=P.C, num_frames=P.F, board_xy=P.board_xy, view_camera=P.view_camera,
        view_frame=P.view_frame, obs_xy=P.obs_xy, fixed_camera=P.fixed_camera,
        init_intrinsics=sp.init_intrinsics, init_cam_rt=sp.init_cam_rt,
        init_board_rt=sp.init_board_rt,
        options=np.array([kw.get("max_num_iterations", 50), kw.get("loss_type", 0),
                          kw.get("loss_scale", 1.0)], dtype=np.float64),
        final_intrinsics=i, final_cam_rt=c, final_board_rt=b,
        cost=np.array(S.cost), radius=np.array(S.radius),
        gradient_max_norm=np.array(S.gradient_max_norm), step_norm=np.array(S.step_norm),
        flags=np.array(S.flags), termination=S.termination,
        initial_residuals=r, initial_jacobian_sample=J[:: max(1, len(J) // 64)],
        jacobian_sample_stride=max(1, len(J) // 64))
    print(name, S.termination, S.num_iterations, "cost", S.cost[0], "->", S.cost[-1])


def save_mp_jacobian():
    sp = synth.generate(num_cameras=3, num_frames=6, board=(5, 4), rig="calib", seed=77)
    p = sp.problem
    rng = np.random.default_rng(5)
    rows = []
    for _ in range(24):
        v = int(rng.integers(0, p.num_views))
        j = int(rng.integers(0, p.corners_per_board))
        m, i = int(p.view_camera[v]), int(p.view_frame[v])
        crt = sp.init_cam_rt[m].copy()
        if len(rows) % 6 == 5:
            sp.init_board_rt[i, :3] = 0.0  # exercise the theta^2 <= eps Taylor branch
        r, J = nr.mp_jacobian(crt, sp.init_board_rt[i], sp.init_intrinsics[m], p.board_xy[j, 0],
                              p.board_xy[j, 1], p.obs_xy[v, j, 0], p.obs_xy[v, j, 1])
        rows.append(dict(cam_rt=crt, board_rt=sp.init_board_rt[i].copy(),
                         intr=sp.init_intrinsics[m].copy(), board=p.board_xy[j].copy(),
                         obs=p.obs_xy[v, j].copy(), r=r, J=J))
    np.savez_compressed(os.path.join(OUT, "jacobian_mpmath.npz"),
                        **{k: np.stack([row[k] for row in rows]) for k in rows[0]})
    print("jacobian_mpmath", len(rows))


if __name__ == "__main__":
    save_solve("mono_cfg1", synth.config(1), max_num_iterations=100)
    save_solve("rig3_small", synth.generate(num_cameras=3, num_frames=30, board=(11, 8),
                                            rig="calib", seed=21))
    save_solve("rig4_huber", synth.generate(num_cameras=4, num_frames=30, board=(11, 8),
                                            rig="calib", seed=22, outlier_fraction=0.05),
               loss_type=1, loss_scale=1.0)
    save_solve("rig4_cauchy", synth.generate(num_cameras=4, num_frames=30, board=(11, 8),
                                             rig="calib", seed=23, outlier_fraction=0.05),
               loss_type=2, loss_scale=1.0)
    save_mp_jacobian()
