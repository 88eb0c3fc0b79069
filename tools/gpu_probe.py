"""One GPU-box visit worth of diagnostics for the host/solve rewrite: k_solve cycle counts,
one-shot tscm_solve() phase timings (cold and cached), creation time at config 3 / config 4."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tscm_calib_b200 import capi, synth

what = sys.argv[1:] or ["solve_cycles", "oneshot", "cfg4_create"]
if "solve_cycles" in what:
    capi.set_debug(2)
    for cfg, frames in ((3, 96), (4, 160)):
        sp = synth.config(cfg, num_frames=frames)
        opt = capi.default_options(max_num_iterations=2)
        a, b, c, s = capi.solve(sp.problem, sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt, opt)
        print("cfg", cfg, s.termination, s.cost, flush=True)
    capi.set_debug(0)
if "oneshot" in what:
    capi.set_debug(1)
    sp = synth.config(3)
    obs = capi.pinned_array(sp.problem.obs_xy.shape)
    obs[...] = sp.problem.obs_xy
    hp = capi.ProblemArrays(sp.problem.board_xy, sp.problem.view_camera, sp.problem.view_frame, obs,
                            sp.problem.num_cameras, sp.problem.num_frames, sp.problem.fixed_camera)
    opt = capi.default_options(max_num_iterations=20, disable_tolerances=1)
    for tag, prob in (("pinned", hp), ("pageable", sp.problem)):
        capi.cache_release()
        for rep in range(3):
            t = time.perf_counter()
            capi.solve(prob, sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt, opt)
            dt = time.perf_counter() - t
            print(f"one-shot {tag} rep {rep}: {dt*1e3:.2f} ms -> {20/dt:.0f} it/s", flush=True)
    capi.set_debug(0)
if "cfg4_create" in what:
    capi.set_debug(1)
    capi.cache_release()
    t = time.perf_counter()
    sp = synth.config(4, num_frames=int(os.environ.get("CFG4_FRAMES", "100000")))
    print(f"synth config 4: {time.perf_counter()-t:.1f} s, views {sp.problem.num_views}", flush=True)
    opt = capi.default_options(max_num_iterations=5, disable_tolerances=1)
    for rep in range(2):
        t = time.perf_counter()
        s = capi.Solver(sp.problem, opt); t1 = time.perf_counter()
        s.set_parameters(sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt); t2 = time.perf_counter()
        r = s.run(); t3 = time.perf_counter()
        s.close(); t4 = time.perf_counter()
        print(f"cfg4 rep {rep}: create {1e3*(t1-t):.1f} set {1e3*(t2-t1):.1f} run(5 it) {1e3*(t3-t2):.1f} destroy {1e3*(t4-t3):.1f} ms", flush=True)
if "group" in what:
    import torch
    n = min(int(os.environ.get("GROUP_GPUS", "2")), torch.cuda.device_count())
    capi.cache_release()
    sp = synth.config(3)
    obs = capi.pinned_array(sp.problem.obs_xy.shape)
    obs[...] = sp.problem.obs_xy
    hp = capi.ProblemArrays(sp.problem.board_xy, sp.problem.view_camera, sp.problem.view_frame, obs,
                            sp.problem.num_cameras, sp.problem.num_frames, sp.problem.fixed_camera)
    init = (sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
    for ng in (1, n):
        opt = capi.default_options(max_num_iterations=20, disable_tolerances=1, num_gpus=ng)
        capi.set_debug(1)
        for rep in range(3):
            t = time.perf_counter()
            capi.solve(hp, *init, opt)
            print(f"num_gpus={ng} one-shot rep {rep}: {(time.perf_counter()-t)*1e3:.2f} ms", flush=True)
        capi.set_debug(0)
        s = capi.Solver(hp, capi.default_options(max_num_iterations=200, disable_tolerances=1, num_gpus=ng))
        s.set_parameters(*init)
        s.time_stage(4, 8)
        s.set_parameters(*init)
        print(f"num_gpus={ng}: graph replay {s.time_stage(4, 40)*1e3:.1f} us per LM iteration", flush=True)
        for k in range(3):
            t = time.perf_counter(); s.set_observations(obs); t1 = time.perf_counter()
            s.set_parameters(*init); t2 = time.perf_counter(); r = s.run(); t3 = time.perf_counter()
            s.get_parameters(); t4 = time.perf_counter()
            print(f"   resident: obs {1e3*(t1-t):.2f} set {1e3*(t2-t1):.2f} run(200) {1e3*(t3-t2):.2f} get {1e3*(t4-t3):.2f} ms", flush=True)
        s.close()
        capi.cache_release()
if "schur_forms" in what:
    for frames in (625, 1250, 2500, 5000):
        sp = synth.config(3, num_frames=frames)
        for form in ("rows", "fused", "pairs"):
            try:
                s = capi.Solver(sp.problem, capi.default_options(max_num_iterations=100, disable_tolerances=1), schur_form=form)
            except capi.TscmError as e:
                print(frames, form, "unsupported"); continue
            s.set_parameters(sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
            s.time_stage(1, 3)
            t1 = s.time_stage(1, 20)
            s.set_parameters(sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
            s.time_stage(4, 5)
            s.set_parameters(sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
            t4 = s.time_stage(4, 30)
            st = {n: s.time_stage(i, 10) for i, n in ((6, "eval5"), (7, "vblocks"), (5, "evalpass"), (2, "solve"), (3, "backsub"))}
            print(f"frames {frames:5d} {form:6s}: schur stage {t1*1e3:6.1f} us, iteration {t4*1e3:6.1f} us  " +
                  " ".join(f"{k} {v*1e3:.1f}" for k, v in st.items()), flush=True)
            s.close()
