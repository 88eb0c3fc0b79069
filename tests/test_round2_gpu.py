"""Round-2 GPU tests: the cached one-shot tscm_solve(), single-process multi-GPU
(tscm_options.num_gpus), multi-process ranks over peer memory, both Ceres-version settings of
the tolerance gating, and config 3 with per-frame visibility masks at its full size.
Everything goes through the C-ABI; the oracle is only the checker."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tscm_calib_b200 import capi, synth
from conftest import ROOT, rel_err

pytestmark = pytest.mark.gpu


def device_count():
    import torch
    return torch.cuda.device_count()


def assert_same(res, ref, params, ref_params, cost_rtol=1e-9, param_rtol=1e-7):
    assert res.termination == ref.termination
    assert res.num_iterations == ref.num_iterations
    np.testing.assert_allclose(res.cost, ref.cost, rtol=cost_rtol)
    for x, y in zip(params, ref_params):
        np.testing.assert_allclose(x, y, rtol=param_rtol, atol=1e-9)


def test_one_shot_solve_reuses_the_cached_solver_bit_identically():
    """tscm_solve() keeps the solver of the last problem structure: a second call on the same
    structure (same or new observations) must give exactly what a fresh solver gives."""
    capi.cache_release()
    sp = synth.config(2)
    init = (sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
    opt = capi.default_options()
    a0, b0, c0, s0 = capi.solve(sp.problem, *init, opt)          # builds the solver
    a1, b1, c1, s1 = capi.solve(sp.problem, *init, opt)          # cached
    np.testing.assert_array_equal(s0.cost, s1.cost)
    for x, y in ((a0, a1), (b0, b1), (c0, c1)):
        np.testing.assert_array_equal(x, y)
    # new observations, other options, same structure: cached solver vs a resident one
    sp2 = synth.generate(num_cameras=4, num_frames=200, board=(11, 8), rig="calib", seed=2, noise_px=0.3)
    assert np.array_equal(sp2.problem.view_frame, sp.problem.view_frame)
    opt2 = capi.default_options(max_num_iterations=7, loss_type="huber", loss_scale=1.0)
    init2 = (sp2.init_intrinsics, sp2.init_cam_rt, sp2.init_board_rt)
    a2, b2, c2, s2 = capi.solve(sp2.problem, *init2, opt2)
    a3, b3, c3, s3 = capi.solve_resident(sp2.problem, *init2, opt2)
    np.testing.assert_array_equal(s2.cost, s3.cost)
    for x, y in ((a2, a3), (b2, b3), (c2, c3)):
        np.testing.assert_array_equal(x, y)
    # a different structure evicts it; caching disabled still works
    spm = synth.config(1)
    capi.solve(spm.problem, spm.init_intrinsics, spm.init_cam_rt, spm.init_board_rt, capi.default_options(max_num_iterations=100))
    capi.cache_configure(0)
    a4, b4, c4, s4 = capi.solve(sp.problem, *init, opt)
    capi.cache_configure(1)
    np.testing.assert_array_equal(s0.cost, s4.cost)
    capi.cache_release()


@pytest.mark.parametrize("needs_success", [0, 1])
def test_both_ceres_tolerance_gatings_match_the_oracle(oracle, needs_success):
    """parameter_tolerance_needs_successful_step = 0 (Ceres <= 2.0) / 1 (>= 2.1: parameter AND
    function tolerance only after a successful step).  A start at the optimum makes the gating
    decide the iteration count."""
    sp = synth.config(2, num_frames=40)
    opt = capi.default_options(parameter_tolerance_needs_successful_step=needs_success)
    init = (sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
    a, b, c, s = capi.solve(sp.problem, *init, opt)
    a0, b0, c0, s0 = oracle.solve(sp.problem, *init, opt)
    assert_same(s, s0, (a, b, c), (a0, b0, c0))
    # restart at the solution: the first step is tiny
    a2, b2, c2, s2 = capi.solve(sp.problem, a, b, c, opt)
    a3, b3, c3, s3 = oracle.solve(sp.problem, a, b, c, opt)
    assert s2.termination == s3.termination and s2.num_iterations == s3.num_iterations


@pytest.mark.parametrize("form", ["auto", "rows", "pairs"])
def test_masked_ring_sample_matches_the_oracle(oracle, form):
    """8-camera outward ring with all-or-nothing visibility masks (config 3's named shape) at a
    size the oracle solves in seconds, under every Schur form.  The start is
    hard enough for rejected steps, which the trace must reproduce too."""
    sp = synth.config(3, num_frames=150, dense=False, rig="ring")
    opt = capi.default_options(max_num_iterations=25)
    init = (sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
    a, b, c, s = capi.solve_resident(sp.problem, *init, opt, schur_form=form)
    a0, b0, c0, s0 = oracle.solve(sp.problem, *init, opt, num_threads=4)
    np.testing.assert_array_equal(s.step_flags, s0.step_flags)
    rejected = (s.step_flags & 2) == 0
    assert np.any(rejected)                         # at least one rejected step
    # a rejected candidate lies far outside the basin (cost 1e12-1e13 against 1e7: points where
    # the projection denominators nearly vanish), so its cost is conditioned 1e5 times worse
    # than the accepted ones: 1e-5 there, the usual 1e-9 on every accepted iteration
    np.testing.assert_allclose(s.cost[rejected], s0.cost[rejected], rtol=1e-5)
    cost, cost0 = s.cost.copy(), s0.cost.copy()
    cost[rejected] = cost0[rejected] = 0.0
    s.cost, s0.cost = cost, cost0
    assert_same(s, s0, (a, b, c), (a0, b0, c0))


@pytest.mark.parametrize("rig,min_fill,other", [("array", 0.95, "pairs"), ("ring", 0.3, "rows")])
def test_config3_with_visibility_masks_full_size(rig, min_fill, other):
    """BASELINE config 3 as named: 8 cameras, 5,000 frames, per-frame all-or-nothing visibility
    masks (main.cpp:33-37) — a forward array (nearly every camera sees every frame) and an outward
    ring (every frame seen by ~3 of 8 cameras).  The oracle cannot solve this size in seconds, so
    size-independent properties: convergence, monotone accepted costs, RMS at the noise floor, a
    second run bit-identical, and another Schur form agreeing to round-off."""
    sp = synth.config(3, dense=False, rig=rig)
    vis = sp.visible.mean()
    assert min_fill <= vis < 1.0, vis
    opt = capi.default_options(max_num_iterations=200)
    init = (sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
    s = capi.Solver(sp.problem, opt)
    s.set_parameters(*init)
    r1 = s.run()
    p1 = s.get_parameters()
    per_cam, overall, rms = s.reprojection_error()
    s.set_parameters(*init)
    r2 = s.run()
    p2 = s.get_parameters()
    s.close()
    assert r1.termination == "CONVERGENCE", (r1.num_iterations, r1.cost[-5:], rms)
    ok = (r1.step_flags & 2) != 0
    assert np.all(np.diff(r1.cost[ok]) < 0)
    assert 0.13 < rms < 0.15, rms                      # 0.1 px noise per axis -> 0.141 px
    np.testing.assert_array_equal(r1.cost, r2.cost)
    for x, y in zip(p1, p2):
        np.testing.assert_array_equal(x, y)
    a3, b3, c3, r3 = capi.solve_resident(sp.problem, *init, opt, schur_form=other)
    assert r3.num_iterations == r1.num_iterations
    np.testing.assert_allclose(r3.cost, r1.cost, rtol=1e-9)
    for x, y in zip((a3, b3, c3), p1):
        np.testing.assert_allclose(x, y, rtol=1e-7, atol=1e-9)
    # the principal points are recovered to a fraction of a pixel (focal lengths trade off against
    # the board distances and are only determined to a few percent)
    assert np.max(np.abs(p1[0][:, 2:4] - sp.gt_intrinsics[:, 2:4])) < 0.05


@pytest.mark.parametrize("cfg,frames,ngpu", [(2, 200, 2), (3, 600, 2), (4, 1200, 2), (3, 800, 4), (3, 1600, 8)])
def test_single_process_multi_gpu_matches_one_gpu(cfg, frames, ngpu):
    """tscm_options.num_gpus: one host thread, N devices, frames sharded inside the library."""
    if device_count() < ngpu:
        pytest.skip(f"needs {ngpu} GPUs")
    capi.cache_release()
    sp = synth.config(cfg, num_frames=frames)
    init = (sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
    a0, b0, c0, s0 = capi.solve(sp.problem, *init, capi.default_options())
    a1, b1, c1, s1 = capi.solve(sp.problem, *init, capi.default_options(num_gpus=ngpu))
    a2, b2, c2, s2 = capi.solve(sp.problem, *init, capi.default_options(num_gpus=ngpu))    # cached group
    assert_same(s1, s0, (a1, b1, c1), (a0, b0, c0))
    np.testing.assert_array_equal(s1.cost, s2.cost)
    # the accuracy read-out of a group is global (it was num_ranks times too large)
    g = capi.Solver(sp.problem, capi.default_options(num_gpus=ngpu))
    g.set_parameters(a1, b1, c1)
    per_g, overall_g, rms_g = g.reprojection_error()
    g.close()
    one = capi.Solver(sp.problem, capi.default_options())
    one.set_parameters(a1, b1, c1)
    per_1, overall_1, rms_1 = one.reprojection_error()
    one.close()
    np.testing.assert_allclose(per_g, per_1, rtol=0, atol=1e-9)
    assert abs(overall_g - overall_1) < 1e-9 and abs(rms_g - rms_1) < 1e-9
    capi.cache_release()


def test_host_adapter_on_two_gpus_reproduces_the_one_gpu_yaml(tmp_path):
    """`host_demo --gpus 2`: MultiCalib::calibrate() with options().num_gpus = 2."""
    if device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from test_host_adapter import dump_problem, read_result
    from tscm_calib_b200 import build as tbuild
    demo = tbuild.build_host()
    sp = synth.config(2)
    prob = str(tmp_path / "problem.bin")
    dump_problem(prob, sp)
    out = {}
    for n in (1, 2):
        res, yml = str(tmp_path / f"r{n}.bin"), str(tmp_path / f"c{n}.yaml")
        r = subprocess.run([demo, prob, res, yml, "--gpus", str(n)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        out[n] = read_result(res, sp.problem.num_cameras, sp.problem.num_frames)
    assert out[1][0] == out[2][0]
    for k in (2, 3, 4):
        np.testing.assert_allclose(out[2][k], out[1][k], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(out[2][5], out[1][5], rtol=0, atol=1e-8)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_process_ranks_match_one_gpu(world):
    """One process per GPU (torchrun), frames sharded by the host, exchanges over CUDA-IPC peer
    memory: tools/dist_parity.py compares every rank's solve with the 1-GPU solve."""
    if device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29511 + world),
           os.path.join(ROOT, "tools", "dist_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0 and "DIST PARITY PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
