"""The plain-double projection family of the unchanged TS interface through
the C++ adapter (host/ts_camera.cpp) against the oracle's restatement of TS.cpp:332-344:
project, Reproject (TS.cpp:227-245: the 3x3 [r1 r2 t] matrix applied to (x, y, 1)),
ReprojectError (TS.h:58-69: summed Euclidean error of one board) and the closed-form
back-projection get_unit_sphere_coordinate (TS.h:39-57)."""
import ctypes as C

import numpy as np

from tscm_calib_b200 import synth

dp = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(dp)


def _setup(hostinit):
    hostinit.hostinit_reproject_error.restype = C.c_double
    rng = np.random.default_rng(7)
    intr = synth.CALIB_INTRINSICS[1].copy()
    # camera-frame points over a fisheye field of view (up to ~100 degrees off axis)
    ang = rng.uniform(0, np.deg2rad(100), 300)
    az = rng.uniform(0, 2 * np.pi, 300)
    rad = rng.uniform(300, 900, 300)
    pts = np.stack([rad * np.sin(ang) * np.cos(az), rad * np.sin(ang) * np.sin(az), rad * np.cos(ang)], axis=1)
    return intr, np.ascontiguousarray(pts)


def test_project_matches_oracle(hostinit, oracle):
    intr, pts = _setup(hostinit)
    uv = np.zeros((len(pts), 2))
    assert hostinit.hostinit_project(_p(intr), _p(pts), len(pts), _p(uv)) == 0
    uv0 = oracle.project(intr, pts)
    np.testing.assert_allclose(uv, uv0, rtol=0, atol=1e-10)


def test_reproject_and_reproject_error_match_oracle(hostinit, oracle):
    intr, _ = _setup(hostinit)
    sp = synth.config(1)
    board = np.concatenate([sp.problem.board_xy, np.zeros((sp.problem.corners_per_board, 1))], axis=1)
    board = np.ascontiguousarray(board)
    n = len(board)
    for f in range(5):
        R = synth.rodrigues(sp.gt_board_rt[f, :3])
        t = sp.gt_board_rt[f, 3:]
        Rt = np.ascontiguousarray(np.stack([R[:, 0], R[:, 1], t], axis=1))        # [r1 r2 t], TS.cpp:232-233
        uv = np.zeros((n, 2))
        assert hostinit.hostinit_reproject(_p(intr), _p(Rt), _p(board), n, _p(uv)) == 0
        P = board @ R.T + t                                                        # z = 0: R p + t = Rt (x, y, 1)
        uv0 = oracle.project(intr, np.ascontiguousarray(P))
        np.testing.assert_allclose(uv, uv0, rtol=0, atol=1e-9)
        px = np.ascontiguousarray(uv0 + np.random.default_rng(f).normal(0, 0.3, uv0.shape))
        err = hostinit.hostinit_reproject_error(_p(intr), _p(px), _p(board), n, _p(np.ascontiguousarray(R)),
                                                _p(np.ascontiguousarray(t)))
        err0 = float(np.sum(np.linalg.norm(px - uv0, axis=1)))
        assert abs(err - err0) <= 1e-9 * max(1.0, err0)


def test_unit_sphere_back_projection_inverts_project(hostinit, oracle):
    intr, pts = _setup(hostinit)
    uv = oracle.project(intr, pts)
    xyz = np.zeros((len(pts), 3))
    assert hostinit.hostinit_unit_sphere(_p(intr), _p(np.ascontiguousarray(uv)), len(uv), _p(xyz)) == 0
    np.testing.assert_allclose(np.linalg.norm(xyz, axis=1), 1.0, atol=1e-12)
    rays = pts / np.linalg.norm(pts, axis=1, keepdims=True)
    np.testing.assert_allclose(xyz, rays, rtol=0, atol=1e-9)
