"""Mono cold-start initialisation: TripleSphereCamera::calibrate up to the
refinement (TS.cpp:36-52), estimate_focal (TS.cpp:110-168) and estimate_extrinsic
(TS.cpp:170-203).

CPU: the oracle (oracle/mono_init_oracle.cpp + cv_calib3d_port.h) against golden vectors produced
with the real OpenCV at the two calls the reference delegates to it (make_golden_init.py).
GPU: tscm_mono_init() (csrc/tscm_monoinit.cu) against the oracle — same statements, no FMA
contraction on either side, so only libm's sin/cos/atan2/asin rounding separates them — against
the OpenCV golden vectors, and through the C++ drop-in adapter."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from tscm_calib_b200 import capi, synth

G = np.load(os.path.join(ROOT, "tests", "golden", "mono_init.npz"))
dp = C.POINTER(C.c_double)


def _d(a):
    return a.ctypes.data_as(dp)


def run_guess(lib, fn, guess7=None, extra=None):
    px = np.ascontiguousarray(G["pixels"], dtype=np.float64)
    has = np.ascontiguousarray(G["has"], dtype=np.uint8)
    F = len(has)
    W, H = (int(v) for v in G["board"])
    iw, ih = (int(v) for v in G["image"])
    intr, Rt = np.zeros(9), np.zeros((F, 3, 3))
    g7 = None if guess7 is None else np.ascontiguousarray(guess7, dtype=np.float64)
    args = [_d(px), has.ctypes.data_as(C.POINTER(C.c_ubyte)), F, W, H, C.c_double(float(G["square"])), iw, ih,
            _d(g7) if g7 is not None else None, _d(intr), _d(Rt)]
    if extra is not None:
        args.append(_d(extra))
    rc = getattr(lib, fn)(*args)
    return rc, intr, Rt


def unit_sphere(intr, px):
    """TS.h:39-57 (vectorised over corners)."""
    fx, fy, cx, cy, xi, lam, alpha, b, c = intr
    x, y = px[..., 0] - cx, px[..., 1] - cy
    mx, my = (fy * x - b * y) / (fx * fy - b * c), (-c * x + fx * y) / (fx * fy - b * c)
    k = alpha / (1 - alpha)
    r2 = mx * mx + my * my
    with np.errstate(invalid="ignore"):
        gamma = (k + np.sqrt(1 + (1 - k * k) * r2)) / (r2 + 1)
        eta = lam * (gamma - k) + np.sqrt(((gamma - k) * (gamma - k) - 1) * lam * lam + 1)
        mz = eta * (gamma - k)
        mu = xi * (mz - lam) + np.sqrt(xi * xi * ((mz - lam) * (mz - lam) - 1) + 1)
    return np.stack([mu * eta * gamma * mx, mu * eta * gamma * my, mu * (mz - lam) - xi], axis=-1)


def pnp_cost(intr, k, Rt):
    """The objective solvePnPRansac minimises for frame k at TS.cpp:193: squared reprojection
    error on the normalised plane of the rotated view (TS.cpp:177-192), finite corners only."""
    px, W, K = G["pixels"][k], int(G["board"][0]), 54
    p = unit_sphere(intr, px[K // 2 - W // 2 - 1])
    a, b = np.arctan2(p[0], p[2]), np.arcsin(p[1])
    R1 = np.array([[np.cos(a), 0, -np.sin(a)], [0, 1, 0], [np.sin(a), 0, np.cos(a)]])
    R2 = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
    T = R2 @ R1
    q = unit_sphere(intr, px) @ T.T
    pn = q[:, :2] / q[:, 2:3]
    ok = np.isfinite(pn).all(axis=1)
    M = T @ Rt
    P = G["worlds"][:, :2] @ M[:, :2].T + M[:, 2]
    return float((((P[:, :2] / P[:, 2:3]) - pn)[ok] ** 2).sum())


def check_poses(intr, Rt, Rt0):
    """Frame by frame: the same pose as OpenCV's (to its LM stopping tolerance), or — where
    OpenCV's RANSAC + 20-iteration LM stopped in the mirrored planar-pose minimum — a strictly
    lower value of the same objective.  Returns the number of frames in the second class."""
    better = 0
    for k in np.flatnonzero(G["has"] == 1):
        rot = np.abs(Rt[k][:, :2] - Rt0[k][:, :2]).max()
        tr = np.linalg.norm(Rt[k][:, 2] - Rt0[k][:, 2]) / np.linalg.norm(Rt0[k][:, 2])
        if rot < 2e-5 and tr < 2e-5:
            continue
        mine, theirs = pnp_cost(intr, k, Rt[k]), pnp_cost(intr, k, Rt0[k])
        assert mine < 0.5 * theirs, (k, rot, tr, mine, theirs)
        better += 1
    return better


def oracle_guess(oracle, guess7=None):
    r = oracle.mono_init(G["board"], G["image"], G["worlds"], G["has"], G["pixels"], guess=guess7)
    return r.rc, r.intrinsics, r.Rt


def test_solve_z_matches_opencv(oracle):
    z = oracle.solve_z(G["solvez_A"])
    z0 = G["solvez_z"]
    if np.dot(z, z0) < 0:          # the sign of a singular vector is arbitrary
        z = -z
    assert abs(np.linalg.norm(z) - 1) < 1e-14
    np.testing.assert_allclose(z, z0, rtol=0, atol=1e-12)


def test_planar_pnp_recovers_exact_pose(oracle):
    rng = np.random.default_rng(9)
    obj = np.array([[(j % 9) * 45.0, (j // 9) * 45.0, 0.0] for j in range(54)])
    for _ in range(10):
        rv = rng.normal(0, 0.5, 3)
        t = np.array([rng.uniform(-300, 100), rng.uniform(-200, 100), rng.uniform(300, 900)])
        P = obj @ synth.rodrigues(rv).T + t
        img = np.ascontiguousarray(P[:, :2] / P[:, 2:3])
        rc, r, tt = oracle.solve_pnp(obj, img)
        assert rc == 0
        np.testing.assert_allclose(synth.rodrigues(r), synth.rodrigues(rv), atol=1e-10)
        np.testing.assert_allclose(tt, t, rtol=1e-10)


def check_cold(rc, intr, Rt):
    assert rc == 0
    g = G["intr_cold"]
    assert intr[2] == 639.5 and intr[3] == 539.5 and list(intr[4:]) == [0.0, 0.0, 0.5, 0.0, 0.0]
    assert abs(intr[0] - g[0]) <= 1e-9 * g[0] and intr[1] == intr[0]
    # OpenCV's LM stops at 20 iterations / FLT_EPSILON; the shim iterates to convergence
    assert check_poses(G["intr_cold"], Rt, G["Rt_cold"]) <= 3      # of 40 frames
    assert np.all(Rt[G["has"] == 0] == 0)


def check_warm(rc, intr, Rt):
    assert int((G["inliers_warm"][G["has"] == 1] < 54).sum()) >= 1
    assert rc == 0
    np.testing.assert_array_equal(intr[:7], G["guess7"])
    assert check_poses(np.concatenate([G["guess7"], [0.0, 0.0]]), Rt, G["Rt_warm"]) <= 3


def test_cold_start_matches_opencv_golden(oracle):
    """No initial guess: cx, cy from the image size, xi = lamda = 0, alpha = 0.5 (TS.cpp:43-47),
    focal from the circle fits, poses from PnP on unit-sphere-normalised corners."""
    check_cold(*oracle_guess(oracle))


def test_warm_start_with_rejected_corners_matches_opencv_golden(oracle):
    """7-argument constructor: intrinsics are kept, only estimate_extrinsic runs (TS.cpp:41,52).
    Under the perturbed guess some corners back-project to NaN (TS.h:47); OpenCV's RANSAC
    rejects them and so must the restatement."""
    check_warm(*oracle_guess(oracle, G["guess7"]))


def test_focal_failure_returns_false(oracle):
    """All frames without a board: estimate_focal finds no row, fx stays 0, calibrate returns
    false before touching the solver (TS.cpp:50)."""
    r = oracle.mono_init((9, 6), (1280, 1080), G["worlds"], np.zeros(3, dtype=np.uint8), np.zeros((3, 54, 2)))
    assert r.rc == 1 and r.intrinsics[0] == 0.0


def test_mono_init_without_a_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.TscmError, match="no CPU fallback"):
        capi.mono_init(G["board"], G["image"], G["worlds"], G["has"], G["pixels"])


# ------------------------------------------------------------------------------------------
# GPU: tscm_mono_init against the oracle and the OpenCV golden vectors
# ------------------------------------------------------------------------------------------
def assert_same_init(r, r0, rot_tol=2e-8, t_rtol=2e-8, stable=None):
    """GPU result against the oracle: identical frame decisions, focal length to 1e-13 (no
    transcendental function on that path), poses to 2e-8: both sides run the Gauss-Newton until no
    step lowers the objective any more, and a least-squares minimum is only defined to
    sqrt(machine epsilon) by its objective — a 1-ulp difference in a sine (CUDA libm vs glibc)
    picks another point of that flat bottom (measured: <= 5e-10)."""
    np.testing.assert_array_equal(r.frame_ok, r0.frame_ok)
    assert r.rows_used == r0.rows_used
    np.testing.assert_allclose(r.intrinsics, r0.intrinsics, rtol=1e-13, atol=0)
    keep = np.ones(len(r.frame_ok), dtype=bool) if stable is None else stable
    np.testing.assert_allclose(r.Rt[keep][:, :, :2], r0.Rt[keep][:, :, :2], rtol=0, atol=rot_tol)
    scale = np.maximum(np.linalg.norm(r0.Rt[:, :, 2], axis=1, keepdims=True), 1.0)
    np.testing.assert_allclose((r.Rt[:, :, 2] / scale)[keep], (r0.Rt[:, :, 2] / scale)[keep], rtol=0, atol=t_rtol)


def stable_frames(oracle, board, image, worlds, has, px, guess=None):
    """Frames whose pose the ORACLE ITSELF reproduces when every pixel coordinate is moved by one
    unit in the last place.  With outlier corners (config 5) a few planar-pose problems have two
    nearly equal minima (the mirrored pose) or a stalled last Gauss-Newton step, and the restated
    solver lands on either side for a 1e-16 perturbation — no implementation can be compared
    tighter than the algorithm's own conditioning there.  Returns a boolean mask."""
    base = oracle.mono_init(board, image, worlds, has, px, guess=guess)
    ok = np.ones(len(has), dtype=bool)
    for seed in (1, 2, 3):
        rng = np.random.default_rng(seed)
        moved = px * (1 + rng.integers(-1, 2, px.shape) * 1.1e-16)
        other = oracle.mono_init(board, image, worlds, has, moved, guess=guess if guess is not None else None)
        ok &= np.abs(other.Rt - base.Rt)[:, :, :2].max(axis=(1, 2)) < 1e-7
    return ok


@pytest.mark.gpu
def test_gpu_cold_and_warm_start_match_oracle_and_opencv_golden(oracle):
    r = capi.mono_init(G["board"], G["image"], G["worlds"], G["has"], G["pixels"])
    r0 = oracle.mono_init(G["board"], G["image"], G["worlds"], G["has"], G["pixels"])
    assert r0.rc == 0
    assert_same_init(r, r0)
    check_cold(0, r.intrinsics, r.Rt)
    r = capi.mono_init(G["board"], G["image"], G["worlds"], G["has"], G["pixels"], guess=G["guess7"])
    r0 = oracle.mono_init(G["board"], G["image"], G["worlds"], G["has"], G["pixels"], guess=G["guess7"])
    assert_same_init(r, r0)
    check_warm(0, r.intrinsics, r.Rt)
    assert r.kernel_ms > 0


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,frames,kw", [(1, 40, {}), (2, 120, {}), (3, 400, dict(dense=False, rig="ring")),
                                           (5, 100, {})])
def test_gpu_mono_init_matches_oracle_per_camera(oracle, cfg, frames, kw):
    """Every camera of a synthetic rig, cold start: masked frames (no corners), outlier corners
    (config 5), boards near the edge of the fisheye field."""
    sp = synth.config(cfg, num_frames=frames, **kw)
    worlds, intr, has, Rt, px = synth.mono_results(sp)
    W = 9 if cfg == 1 else 11
    H = worlds.shape[0] // W
    for m in range(has.shape[0]):
        r = capi.mono_init((W, H), (1280, 1080), worlds, has[m], px[m])
        r0 = oracle.mono_init((W, H), (1280, 1080), worlds, has[m], px[m])
        assert r0.rc == 0 and r0.frame_ok.sum() >= 0.9 * has[m].sum()
        stable = stable_frames(oracle, (W, H), (1280, 1080), worlds, has[m], px[m])
        assert stable.sum() >= 0.9 * len(stable), (cfg, m, int(stable.sum()))
        if cfg != 5:
            assert stable.all(), (cfg, m, int(stable.sum()))
        # the flat bottom is 1e-8 wide on clean data, 4e-8 with outlier corners (oracle against itself)
        tol = 2e-7 if cfg == 5 else 2e-8
        assert_same_init(r, r0, rot_tol=tol, t_rtol=tol, stable=stable)


@pytest.mark.gpu
def test_gpu_warp_and_thread_forms_agree(oracle, monkeypatch):
    """The warp-per-frame kernel (default) against the thread-per-frame kernel that keeps the
    oracle's statement order (TSCM_MI_FORM=thread): same frame decisions, same focal length bit for
    bit (the focal kernel is shared), poses to the width of the flat bottom."""
    args = (G["board"], G["image"], G["worlds"], G["has"], G["pixels"])
    for guess in (None, G["guess7"]):
        monkeypatch.delenv("TSCM_MI_FORM", raising=False)
        rw = capi.mono_init(*args, guess=guess)
        monkeypatch.setenv("TSCM_MI_FORM", "thread")
        rt = capi.mono_init(*args, guess=guess)
        monkeypatch.delenv("TSCM_MI_FORM", raising=False)
        np.testing.assert_array_equal(rw.intrinsics, rt.intrinsics)
        assert_same_init(rw, rt)
        assert_same_init(rt, oracle.mono_init(*args, guess=guess))


@pytest.mark.gpu
def test_gpu_focal_failure_and_bad_arguments():
    r = capi.mono_init((9, 6), (1280, 1080), G["worlds"], np.zeros(3, dtype=np.uint8), np.zeros((3, 54, 2)))
    assert r.intrinsics[0] == 0.0 and r.rows_used == 0 and not r.frame_ok.any() and not r.Rt.any()
    with pytest.raises(capi.TscmError, match="bad mono-init sizes"):
        capi.mono_init((65, 1), (1280, 1080), np.zeros((65, 3)), np.ones(1, dtype=np.uint8), np.zeros((1, 65, 2)))


@pytest.mark.gpu
def test_gpu_mono_init_config3_size(oracle):
    """One camera of BASELINE config 3: 5,000 frames x 88 corners in one call, every frame
    compared with the oracle; all frames must get a pose."""
    sp = synth.config(3)
    worlds, intr, has, Rt, px = synth.mono_results(sp)
    r = capi.mono_init((11, 8), (1280, 1080), worlds, has[0], px[0])
    assert r.frame_ok.all() and r.rows_used > 0.3 * 8 * 5000     # oblique rows are skipped (TS.cpp:151)
    print(f"mono init of 5,000 frames: {r.kernel_ms:.1f} ms of kernels, focal {r.intrinsics[0]:.2f}")
    r0 = oracle.mono_init((11, 8), (1280, 1080), worlds, has[0], px[0])
    assert_same_init(r, r0)


@pytest.mark.gpu
def test_adapter_initial_guess_on_gpu_matches_opencv_golden(hostinit):
    """The drop-in TripleSphereCamera::initial_guess (cold and 7-argument warm start) through the
    C++ adapter."""
    check_cold(*run_guess(hostinit, "hostinit_initial_guess"))
    check_warm(*run_guess(hostinit, "hostinit_initial_guess", guess7=G["guess7"]))
    px = np.zeros((3, 54, 2))
    has = np.zeros(3, dtype=np.uint8)
    intr, Rt = np.zeros(9), np.zeros((3, 3, 3))
    rc = hostinit.hostinit_initial_guess(_d(px), has.ctypes.data_as(C.POINTER(C.c_ubyte)), 3, 9, 6,
                                         C.c_double(45.0), 1280, 1080, None, _d(intr), _d(Rt))
    assert rc == 1 and intr[0] == 0.0


@pytest.mark.gpu
def test_cold_start_calibrate_on_gpu_matches_oracle(hostinit, oracle):
    """The whole TripleSphereCamera::calibrate from nothing but corners: cold start on the host,
    refinement on the B200.  The oracle solves from the same initial guess (exported before the
    refinement) with the reference's options (100 iterations, TS.cpp:274)."""
    rc0, intr0, Rt0 = run_guess(hostinit, "hostinit_initial_guess")
    summ = np.zeros(4)
    rc, intr, Rt = run_guess(hostinit, "hostinit_calibrate", extra=summ)
    has = G["has"] == 1
    K = 54
    # the parameter packing of TS.cpp:62-73 (single-precision r1, r2) restated for the oracle
    board_rt = np.zeros((int(has.sum()), 6))
    for n, i in enumerate(np.flatnonzero(has)):
        r1 = Rt0[i][:, 0].astype(np.float32)
        r2 = Rt0[i][:, 1].astype(np.float32)
        r3 = np.cross(r1, r2).astype(np.float32)
        R = np.stack([r1, r2, r3], axis=1).astype(np.float64)
        board_rt[n, :3] = synth.rotation_to_rvec(R)
        board_rt[n, 3:] = Rt0[i][:, 2]
    F = int(has.sum())
    p = capi.ProblemArrays(G["worlds"][:, :2].copy(), np.zeros(F, np.int32), np.arange(F, dtype=np.int32),
                           G["pixels"][has].reshape(F * K, 2).copy(), 1, F, 0)
    opt = capi.default_options(max_num_iterations=100)
    a0, _, c0, s0 = oracle.solve(p, intr0.reshape(1, 9), np.zeros((1, 6)), board_rt, opt)
    assert capi.TERMINATION[int(summ[0])] == s0.termination and int(summ[1]) == s0.num_iterations
    assert abs(summ[2] - s0.initial_cost) <= 1e-7 * s0.initial_cost
    assert abs(summ[3] - s0.final_cost) <= 1e-9 * s0.final_cost
    assert (rc == 0) == (s0.termination == "CONVERGENCE")
    np.testing.assert_allclose(intr, a0[0], rtol=1e-7, atol=1e-9)
    for n, i in enumerate(np.flatnonzero(has)):
        np.testing.assert_allclose(Rt[i][:, :2], synth.rodrigues(c0[n, :3])[:, :2], atol=1e-8)
        np.testing.assert_allclose(Rt[i][:, 2], c0[n, 3:], rtol=1e-7, atol=1e-7)
    # and the calibration actually lands on the generating camera (0.1 px noise)
    rms = np.sqrt(2 * summ[3] / (F * K))
    print(f"cold start: focal {intr0[0]:.2f} -> fx {intr[0]:.3f} (truth {synth.CALIB_INTRINSICS[0, 0]:.3f}), "
          f"{int(summ[1])} iterations, rms {rms:.4f} px")
    assert rms < 0.2
    # (f, xi, lamda, alpha) are nearly degenerate in the TS model: the principal point is pinned
    # by the data, the focal length only loosely
    np.testing.assert_allclose(intr[2:4], synth.CALIB_INTRINSICS[0, 2:4], atol=1.0)
    np.testing.assert_allclose(intr[:2], synth.CALIB_INTRINSICS[0, :2], rtol=0.1)
