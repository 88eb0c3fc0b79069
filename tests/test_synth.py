"""Synthetic generator and frame sharding (host-side logic, CPU)."""
import numpy as np

from tscm_calib_b200 import synth


def test_rodrigues_roundtrip():
    rng = np.random.default_rng(0)
    r = rng.normal(0, 1.0, (200, 3))
    R = synth.rodrigues(r)
    np.testing.assert_allclose(np.einsum("nij,nkj->nik", R, R), np.broadcast_to(np.eye(3), R.shape), atol=1e-12)
    r2 = synth.rotation_to_rvec(R)
    np.testing.assert_allclose(synth.rodrigues(r2), R, atol=1e-12)
    assert np.all(synth.rodrigues(np.zeros(3)) == np.eye(3))


def test_config_shapes_and_ordering():
    sp = synth.config(2)
    p = sp.problem
    assert p.num_cameras == 4 and p.num_frames == 200 and p.corners_per_board == 88
    assert np.all(sp.visible.any(axis=0))                       # multi_calib.cpp:102
    key = p.view_camera.astype(np.int64) * p.num_frames + p.view_frame
    assert np.all(np.diff(key) > 0)                             # camera-major, frames increasing
    assert np.all(sp.init_cam_rt[0] == 0) and np.all(sp.gt_cam_rt[0] == 0)  # multi_calib.cpp:21-22
    assert np.all(sp.init_intrinsics[:, 7:] == 0)
    assert p.obs_xy.min() > -2 and p.obs_xy[..., 0].max() < synth.IMAGE_W + 2


def test_generator_is_deterministic():
    a, b = synth.config(1), synth.config(1)
    np.testing.assert_array_equal(a.problem.obs_xy, b.problem.obs_xy)
    np.testing.assert_array_equal(a.init_board_rt, b.init_board_rt)


def test_dense_config3_geometry_small():
    sp = synth.config(3, num_frames=64)
    assert sp.visible.all() and sp.problem.num_views == 8 * 64
    assert sp.num_observations == 8 * 64 * 88


def test_outliers_fraction():
    sp = synth.config(5)
    uv, _, _ = synth._compose_views(sp.gt_intrinsics, sp.gt_cam_rt, sp.gt_board_rt, sp.problem.board_xy)
    clean = uv[sp.problem.view_camera, sp.problem.view_frame]
    d = np.linalg.norm(sp.problem.obs_xy - clean, axis=-1)
    frac = (d > 2.0).mean()
    assert 0.03 < frac < 0.07


def test_shard_frames_partitions_everything():
    sp = synth.config(2)
    p = sp.problem
    for world in (2, 3, 4, 8):
        seen_frames, views = [], 0
        for r in range(world):
            local, frames = synth.shard_frames(sp, r, world)
            seen_frames.append(frames)
            views += local.num_views
            assert local.num_frames == len(frames) and local.num_cameras == p.num_cameras
            if local.num_views:
                assert local.view_frame.min() >= 0 and local.view_frame.max() < local.num_frames
                assert len(np.unique(local.view_frame)) == local.num_frames
            sel = np.isin(p.view_frame, frames)
            np.testing.assert_array_equal(local.obs_xy, p.obs_xy[sel])
        assert views == p.num_views
        np.testing.assert_array_equal(np.concatenate(seen_frames), np.arange(p.num_frames))
        sizes = [len(f) for f in seen_frames]
        assert max(sizes) - min(sizes) <= 0.25 * p.num_frames / world + 2


def test_batched_generation_concatenates_frame_batches():
    """config_batched: frames drawn batch by batch for one common rig (what the config-4
    stress run and the weak-scaling bench use); serial and parallel generation agree, views
    stay camera-major / frame-minor, and a rank's share is a subset of the same frames."""
    sp = synth.config_batched(4, 500, batch=200, processes=1)
    par = synth.config_batched(4, 500, batch=200, processes=2)
    p = sp.problem
    np.testing.assert_array_equal(p.obs_xy, par.problem.obs_xy)
    assert p.num_frames == 500 and sp.visible.shape == (16, 500) and sp.init_board_rt.shape == (500, 6)
    key = p.view_camera.astype(np.int64) * p.num_frames + p.view_frame
    assert np.all(np.diff(key) > 0)
    assert len(np.unique(p.view_frame)) == 500                  # every frame is seen by a camera
    cams, frames = np.nonzero(sp.visible)
    np.testing.assert_array_equal(cams, p.view_camera)
    np.testing.assert_array_equal(frames, p.view_frame)
    # the first batch is config(4) with frame_seed 0; the rig is common to all batches
    first = synth.config(4, num_frames=200, frame_seed=0)
    np.testing.assert_array_equal(sp.gt_board_rt[:200], first.gt_board_rt)
    np.testing.assert_array_equal(sp.gt_intrinsics, first.gt_intrinsics)
    # ranks of a strong-scaling run take alternate batches: together, the same frames
    r0 = synth.config_batched(4, 500, batch=200, rank=0, world=2)
    r1 = synth.config_batched(4, 500, batch=200, rank=1, world=2)
    assert r0.problem.num_frames == 300 and r1.problem.num_frames == 200
    np.testing.assert_array_equal(r0.gt_board_rt, np.concatenate([sp.gt_board_rt[:200], sp.gt_board_rt[400:]]))
    np.testing.assert_array_equal(r1.gt_board_rt, sp.gt_board_rt[200:400])
    assert r0.problem.num_observations + r1.problem.num_observations == p.num_observations
