"""MultiCalib's pose-graph initialisation (multi_calib.cpp:6-153) in the C++
drop-in adapter against a numpy transcription of the reference constructor
(tests/golden/make_golden_posegraph.py): the same candidates must win."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from tscm_calib_b200 import synth

G = np.load(os.path.join(ROOT, "tests", "golden", "pose_graph.npz"))
dp = C.POINTER(C.c_double)
up = C.POINTER(C.c_ubyte)


def _d(a):
    return a.ctypes.data_as(dp)


def run(lib, has=None, Rt=None):
    has = np.ascontiguousarray(G["has"] if has is None else has, dtype=np.uint8)
    Rt = np.ascontiguousarray(G["Rt"] if Rt is None else Rt, dtype=np.float64)
    Cn, F = has.shape
    W, H = (int(v) for v in G["board"])
    px = np.ascontiguousarray(G["pixels"], dtype=np.float64)
    intr = np.ascontiguousarray(G["intrinsics"], dtype=np.float64)
    cam_R, cam_t, cam_rt = np.zeros((Cn, 3, 3)), np.zeros((Cn, 3)), np.zeros((Cn, 6))
    b_R, b_t, b_rt = np.zeros((F, 3, 3)), np.zeros((F, 3)), np.zeros((F, 6))
    b_init = np.zeros(F, dtype=np.uint8)
    rc = lib.hostinit_pose_graph(Cn, F, W, H, C.c_double(float(G["square"])), _d(px), has.ctypes.data_as(up),
                                 _d(intr), _d(Rt), _d(cam_R), _d(cam_t), _d(cam_rt), _d(b_R), _d(b_t), _d(b_rt),
                                 b_init.ctypes.data_as(up))
    return rc, cam_R, cam_t, cam_rt, b_R, b_t, b_rt, b_init


def test_pose_graph_selects_the_reference_candidates(hostinit):
    rc, cam_R, cam_t, cam_rt, b_R, b_t, b_rt, b_init = run(hostinit)
    assert rc == 0
    np.testing.assert_array_equal(b_init, G["board_init"])
    # identical candidate choice => identical compositions up to the order of a few FP64 sums
    np.testing.assert_allclose(cam_R, G["cam_R"], atol=1e-12)
    np.testing.assert_allclose(cam_t, G["cam_t"], atol=1e-9)
    np.testing.assert_allclose(b_R, G["board_R"], atol=1e-12)
    np.testing.assert_allclose(b_t, G["board_t"], atol=1e-9)
    # camera 0 is the reference frame (multi_calib.cpp:21-22)
    assert np.all(cam_R[0] == np.eye(3)) and np.all(cam_t[0] == 0) and np.all(cam_rt[0] == 0)
    # rt_ = Rodrigues(R), t (multi_calib.h:16-18)
    for R, t, rt in list(zip(cam_R, cam_t, cam_rt)) + list(zip(b_R, b_t, b_rt)):
        np.testing.assert_allclose(synth.rodrigues(rt[:3]), R, atol=2e-7)    # R comes from float r1, r2
        np.testing.assert_array_equal(rt[3:], t)
    # and the chain lands near the generating rig (mono poses carry 2 mrad / 1 mm of error)
    np.testing.assert_allclose(cam_t, G["gt_cam_rt"][:, 3:], atol=10.0)
    np.testing.assert_allclose(cam_R, synth.rodrigues(G["gt_cam_rt"][:, :3]), atol=2e-2)


def test_frame_seen_by_no_camera_stays_uninitialised(hostinit):
    """multi_calib.cpp:102: such a board is skipped (and later excluded from the solve)."""
    has = G["has"].copy()
    has[:, 7] = 0
    rc, *_, b_init = run(hostinit, has=has)
    assert rc == 0 and b_init[7] == 0 and b_init.sum() == has.any(axis=0).sum()


def test_cameras_without_a_common_board_are_reported(hostinit):
    """The reference indexes Rs[-1] here (multi_calib.cpp:51,86); the adapter stops with a message."""
    has = G["has"].copy()
    shared = (has[1] & has[2]).astype(bool)
    has[2, shared] = 0
    rc, *_ = run(hostinit, has=has)
    assert rc == 2


@pytest.mark.gpu
def test_whole_flow_from_corners_only_on_gpu(hostinit):
    """The reference's main.cpp flow with nothing but detected corners: cold-start mono
    calibration of every camera (host init + GPU refinement), pose graph, joint GPU refinement.
    No oracle at this level (it needs OpenCV + Ceres end to end): the result is checked against
    the generating rig and the noise floor."""
    sp = synth.config(2, num_frames=80)
    p = sp.problem
    Cn, F, K = p.num_cameras, p.num_frames, p.corners_per_board
    has = np.ascontiguousarray(sp.visible, dtype=np.uint8)
    px = np.zeros((Cn, F, K, 2))
    px[p.view_camera, p.view_frame] = p.obs_xy
    intr, cam_rt, board_rt, summ = np.zeros((Cn, 9)), np.zeros((Cn, 6)), np.zeros((F, 6)), np.zeros(5)
    rc = hostinit.hostinit_full_pipeline(Cn, F, 11, 8, C.c_double(45.0), 1280, 1080, _d(px), has.ctypes.data_as(up),
                                         _d(intr), _d(cam_rt), _d(board_rt), _d(summ))
    assert rc == 0
    rms = np.sqrt(2 * summ[3] / p.num_observations)
    print(f"mono calibrations converged: {int(summ[0])}/{Cn}; joint solve: {int(summ[2])} iterations, "
          f"rms {rms:.4f} px, mean error {summ[4]:.4f} px")
    assert int(summ[1]) == 0                                   # CONVERGENCE
    assert 0.12 < rms < 0.16                                   # 0.1 px noise per axis
    assert np.all(cam_rt[0] == 0)                              # camera 0 stays the reference frame
    np.testing.assert_allclose(cam_rt[:, 3:], sp.gt_cam_rt[:, 3:], atol=3.0)          # mm
    np.testing.assert_allclose(synth.rodrigues(cam_rt[:, :3]), synth.rodrigues(sp.gt_cam_rt[:, :3]), atol=5e-3)
    np.testing.assert_allclose(intr[:, 2:4], sp.gt_intrinsics[:, 2:4], atol=1.0)      # principal points, px
