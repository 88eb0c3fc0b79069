"""MultiCalib's pose-graph initialisation (multi_calib.cpp:6-153).

CPU: the C oracle (oracle/pose_graph_oracle.c) against a numpy transcription of the reference
constructor (tests/golden/make_golden_posegraph.py) — same candidates, same scores.
GPU: tscm_pose_graph_init() (candidate scoring in CUDA) against the oracle, BIT for bit on every
candidate's summed error, through the C-ABI and through the C++ drop-in adapter."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from tscm_calib_b200 import capi, synth

G = np.load(os.path.join(ROOT, "tests", "golden", "pose_graph.npz"))
dp = C.POINTER(C.c_double)
up = C.POINTER(C.c_ubyte)


def _d(a):
    return a.ctypes.data_as(dp)


def golden_worlds():
    W, H = (int(v) for v in G["board"])
    xy = synth.make_board(W, H, float(G["square"]))
    return np.concatenate([xy, np.zeros((W * H, 1))], axis=1)


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


# ------------------------------------------------------------------------------------------
# CPU: the oracle is pinned to the numpy transcription
# ------------------------------------------------------------------------------------------
def test_oracle_matches_the_numpy_transcription(oracle):
    r = oracle.pose_graph(golden_worlds(), G["intrinsics"], G["has"], G["Rt"], G["pixels"])
    assert r.rc == 0
    np.testing.assert_array_equal(r.board_init, G["board_init"])
    np.testing.assert_array_equal(r.camera_choice, G["cam_choice"])
    np.testing.assert_array_equal(r.board_choice, G["board_choice"])
    np.testing.assert_allclose(r.camera_pose[:, :9].reshape(-1, 3, 3), G["cam_R"], atol=1e-12)
    np.testing.assert_allclose(r.camera_pose[:, 9:], G["cam_t"], atol=1e-9)
    np.testing.assert_allclose(r.board_pose[:, :9].reshape(-1, 3, 3), G["board_R"], atol=1e-12)
    np.testing.assert_allclose(r.board_pose[:, 9:], G["board_t"], atol=1e-9)
    # every candidate's score, not only the winners (numpy sums with other association: 1e-12)
    for mine, ref in ((r.camera_candidate_error, G["cam_err"]), (r.board_candidate_error, G["board_err"])):
        np.testing.assert_array_equal(np.isnan(mine), np.isnan(ref))
        m = ~np.isnan(ref)
        np.testing.assert_allclose(mine[m], ref[m], rtol=1e-12)
    assert m.sum() > 30
    # camera 0 is the reference frame (multi_calib.cpp:21-22)
    np.testing.assert_array_equal(r.camera_pose[0], [1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0])


def test_oracle_single_candidate_entry_equals_the_full_loop(oracle):
    r = oracle.pose_graph(golden_worlds(), G["intrinsics"], G["has"], G["Rt"], G["pixels"])
    for i in range(1, G["has"].shape[0]):
        for j in np.flatnonzero(G["has"][i - 1] & G["has"][i])[:5]:
            e = oracle.pose_pair_error(i, j, golden_worlds(), G["intrinsics"], G["has"], G["Rt"], G["pixels"],
                                       r.camera_pose[i - 1])
            assert e == r.camera_candidate_error[i, j]


def test_oracle_frame_seen_by_no_camera_stays_uninitialised(oracle):
    """multi_calib.cpp:102: such a board is skipped (and later excluded from the solve)."""
    has = G["has"].copy()
    has[:, 7] = 0
    r = oracle.pose_graph(golden_worlds(), G["intrinsics"], has, G["Rt"], G["pixels"])
    assert r.rc == 0 and r.board_init[7] == 0 and r.board_init.sum() == has.any(axis=0).sum()


def test_oracle_reports_cameras_without_a_common_board(oracle):
    has = G["has"].copy()
    has[2, (has[1] & has[2]).astype(bool)] = 0
    assert oracle.pose_graph(golden_worlds(), G["intrinsics"], has, G["Rt"], G["pixels"]).rc == 2


def test_pose_graph_without_a_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.TscmError, match="no CPU fallback"):
        capi.pose_graph_init(golden_worlds(), G["intrinsics"], G["has"], G["Rt"], G["pixels"])


# ------------------------------------------------------------------------------------------
# GPU: scoring kernels against the oracle, bit for bit
# ------------------------------------------------------------------------------------------
def assert_bit_identical(r, r0):
    np.testing.assert_array_equal(r.camera_choice, r0.camera_choice)
    np.testing.assert_array_equal(r.board_choice, r0.board_choice)
    np.testing.assert_array_equal(r.board_init, r0.board_init)
    # NaN marks "no candidate" in both; everything else must agree in every bit
    np.testing.assert_array_equal(r.camera_candidate_error.view(np.int64) * ~np.isnan(r.camera_candidate_error),
                                  r0.camera_candidate_error.view(np.int64) * ~np.isnan(r0.camera_candidate_error))
    np.testing.assert_array_equal(np.isnan(r.camera_candidate_error), np.isnan(r0.camera_candidate_error))
    np.testing.assert_array_equal(r.board_candidate_error.view(np.int64) * ~np.isnan(r.board_candidate_error),
                                  r0.board_candidate_error.view(np.int64) * ~np.isnan(r0.board_candidate_error))
    np.testing.assert_array_equal(np.isnan(r.board_candidate_error), np.isnan(r0.board_candidate_error))
    np.testing.assert_array_equal(r.camera_pose, r0.camera_pose)
    np.testing.assert_array_equal(r.board_pose[r0.board_init == 1], r0.board_pose[r0.board_init == 1])


@pytest.mark.gpu
def test_gpu_scores_equal_the_oracle_bit_for_bit_on_the_golden_case(oracle):
    args = (golden_worlds(), G["intrinsics"], G["has"], G["Rt"], G["pixels"])
    r = capi.pose_graph_init(*args)
    r0 = oracle.pose_graph(*args)
    assert_bit_identical(r, r0)
    np.testing.assert_array_equal(r.camera_choice, G["cam_choice"])
    np.testing.assert_array_equal(r.board_choice, G["board_choice"])
    assert r.projections > 0 and r.kernel_ms > 0


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,frames,kw", [
    (3, 300, dict(dense=False, rig="ring")),       # 8-camera ring, all-or-nothing masks
    (3, 257, dict()),                              # dense: every pair shares every board, ragged tile
    (5, 120, dict()),                              # 5 % outlier corners: large errors in the sums
    (2, 64, dict(board=(5, 3))),                   # K = 15: odd corner count (staging alignment, tail corner)
    (1, 40, dict()),                               # one camera: no chain, single-candidate boards
])
def test_gpu_scores_equal_the_oracle_bit_for_bit(oracle, cfg, frames, kw):
    sp = synth.config(cfg, num_frames=frames, **kw)
    args = synth.mono_results(sp, seed=7)
    r = capi.pose_graph_init(*args)
    r0 = oracle.pose_graph(*args)
    assert r0.rc == 0
    assert_bit_identical(r, r0)


@pytest.mark.gpu
def test_gpu_skew_terms_and_degenerate_inputs(oracle):
    """Non-zero skew b, c (the division path of the projection); a wildly wrong mono pose (its
    candidate must lose, and the frame's huge error enters every other sum identically); a NaN mono
    pose (it enters EVERY candidate's sum, NaN < 1e10 is false, nobody wins: the reference would
    index Rs[-1], the oracle returns 2 and the library TSCM_ERR_NO_CANDIDATE)."""
    sp = synth.config(2, num_frames=48)
    worlds, intr, has, Rt, px = synth.mono_results(sp, seed=3)
    intr[:, 7], intr[:, 8] = 0.013, -0.007
    r, r0 = capi.pose_graph_init(worlds, intr, has, Rt, px), oracle.pose_graph(worlds, intr, has, Rt, px)
    assert_bit_identical(r, r0)
    j = int(np.flatnonzero(has[0] & has[1])[0])
    Rt2 = Rt.copy()
    Rt2[1, j, :, 2] += 1e7
    r, r0 = capi.pose_graph_init(worlds, intr, has, Rt2, px), oracle.pose_graph(worlds, intr, has, Rt2, px)
    assert r0.rc == 0 and r0.camera_choice[1] != j
    assert_bit_identical(r, r0)
    Rt2 = Rt.copy()
    Rt2[1, j] = np.nan
    assert oracle.pose_graph(worlds, intr, has, Rt2, px).rc == 2
    with pytest.raises(capi.TscmError, match="none of the"):
        capi.pose_graph_init(worlds, intr, has, Rt2, px)


@pytest.mark.gpu
def test_gpu_error_codes(oracle):
    has = G["has"].copy()
    has[2, (has[1] & has[2]).astype(bool)] = 0
    with pytest.raises(capi.TscmError, match="share no board"):
        capi.pose_graph_init(golden_worlds(), G["intrinsics"], has, G["Rt"], G["pixels"])
    has = G["has"].copy()
    has[:, 7] = 0
    r = capi.pose_graph_init(golden_worlds(), G["intrinsics"], has, G["Rt"], G["pixels"])
    assert r.board_init[7] == 0 and r.board_init.sum() == has.any(axis=0).sum() and r.board_choice[7] == -1


@pytest.mark.gpu
def test_gpu_config3_full_size_spot_checks(oracle):
    """BASELINE config 3 as named (8 cameras x 5,000 boards, every adjacent pair shares all of
    them): 7 x 5,000^2 x 2 x 88 = 3.1e10 TS projections.  The full oracle loop would take a quarter
    of an hour, so the winner and a random sample of candidates per camera are re-scored by the
    oracle's single-candidate entry — bit for bit — and the chain must land on the generating rig."""
    sp = synth.config(3)
    worlds, intr, has, Rt, px = synth.mono_results(sp, seed=11)
    r = capi.pose_graph_init(worlds, intr, has, Rt, px)
    Cn, F = has.shape
    assert r.projections >= (Cn - 1) * F * F * 2 * 88
    print(f"pose graph at config 3: {r.projections / 1e9:.1f} G projections in {r.kernel_ms:.1f} ms of kernels "
          f"= {r.projections / r.kernel_ms / 1e6:.1f} G projections/s")
    rng = np.random.default_rng(5)
    for i in range(1, Cn):
        row = r.camera_candidate_error[i]
        assert not np.isnan(row).any()
        assert r.camera_choice[i] == int(np.argmin(row))         # first minimum wins (strict <)
        for j in [r.camera_choice[i]] + list(rng.integers(0, F, 6)):
            e = oracle.pose_pair_error(i, int(j), worlds, intr, has, Rt, px, r.camera_pose[i - 1])
            assert e == row[j], (i, j, e, row[j])
    assert r.board_init.all()
    np.testing.assert_allclose(r.camera_pose[:, 9:], sp.gt_cam_rt[:, 3:], atol=10.0)
    np.testing.assert_allclose(r.camera_pose[:, :9].reshape(-1, 3, 3), synth.rodrigues(sp.gt_cam_rt[:, :3]), atol=2e-2)


# ------------------------------------------------------------------------------------------
# GPU: through the C++ drop-in adapter (MultiCalib's constructor)
# ------------------------------------------------------------------------------------------
@pytest.mark.gpu
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


@pytest.mark.gpu
def test_frame_seen_by_no_camera_stays_uninitialised(hostinit):
    """multi_calib.cpp:102: such a board is skipped (and later excluded from the solve)."""
    has = G["has"].copy()
    has[:, 7] = 0
    rc, *_, b_init = run(hostinit, has=has)
    assert rc == 0 and b_init[7] == 0 and b_init.sum() == has.any(axis=0).sum()


@pytest.mark.gpu
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
