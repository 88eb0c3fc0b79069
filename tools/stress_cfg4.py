"""BASELINE config 4 (stress: 16-camera ring, 100,000 frames x 88 corners with visibility
masks, ~55 M observations = ~110 M residuals) at FULL size on one B200.

There is no oracle at this size (the CPU port needs minutes per iteration), so the run is
checked through size-independent properties: monotone cost over accepted steps, RMS at the
noise floor, recovery of the generating cameras, and bit-identical repeat solves.  Timing:
whole LM iterations replayed from the CUDA graph (tscm_solver_time_stage 4), CUDA events.

  python tools/stress_cfg4.py [--frames 100000] [--out gpurun_out/stress_cfg4.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tscm_calib_b200 import capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=100000)
    ap.add_argument("--processes", type=int, default=min(10, os.cpu_count() or 1))
    ap.add_argument("--timed-iterations", type=int, default=10)
    ap.add_argument("--batch", type=int, default=0,
                    help="frames per generated batch (batches are dealt round-robin to the ranks); "
                         "default: about 10,000, adjusted so that every rank gets the same number")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "stress_cfg4.json"))
    a = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        # strong scaling: the 10,000-frame batches of the same problem are dealt to the ranks
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if a.batch <= 0:
        per_rank = max(1, round(a.frames / (10000.0 * world)))       # batches per rank
        a.batch = -(-a.frames // (per_rank * world))
    t0 = time.time()
    sp = synth.config_batched(4, a.frames, batch=a.batch, processes=max(1, a.processes // world),
                              rank=rank, world=world)
    p = sp.problem
    t_gen = time.time() - t0
    N = p.num_observations
    if world > 1:
        tn = torch.tensor([N], dtype=torch.int64, device="cuda")
        dist.all_reduce(tn)
        N = int(tn.item())
    print(f"generated {p.num_cameras} cameras x {p.num_frames} frames: {p.num_views} views, {N} observations "
          f"({N * 16 / 1e6:.0f} MB) in {t_gen:.1f} s", flush=True)

    t0 = time.time()
    s = capi.Solver(p, capi.default_options(), device=local)
    if world > 1:
        capi.attach_ranks(s, rank, world)
    t_create = time.time() - t0
    res = {"workload": "config4: 16-camera ring x %d frames x 88 corners, visibility masks" % a.frames,
           "n_gpus": world, "scaling": "strong", "frames_this_rank": int(p.num_frames), "batch": int(a.batch),
           "num_views_this_rank": int(p.num_views), "observations": int(N), "residuals": int(2 * N),
           "visible_fraction": float(p.num_views) / (p.num_cameras * p.num_frames),
           "reduced_size": int(s.reduced_size()), "generate_s": t_gen, "solver_create_s": t_create}

    def solve():
        s.set_parameters(sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
        t = time.time()
        r = s.run()
        return r, time.time() - t, s.get_parameters()

    r1, wall1, x1 = solve()
    r2, wall2, x2 = solve()
    acc = r1.cost[(r1.step_flags & 2) != 0]
    per, overall, rms = s.reprojection_error() if world == 1 else (None, float("nan"), float(np.sqrt(2 * r1.final_cost / N)))
    res.update({
        "termination": r1.termination, "iterations": int(r1.num_iterations),
        "successful_steps": int(r1.num_successful_steps), "initial_cost": float(r1.initial_cost),
        "final_cost": float(r1.final_cost), "rms_px": float(rms), "mean_reprojection_error_px": float(overall),
        "solve_wall_s": wall2,
        "cost_monotone_over_accepted_steps": bool(np.all(np.diff(acc) <= 0)),
        "repeat_solve_bit_identical": bool(all(np.array_equal(u, v) for u, v in zip(x1, x2)) and
                                           np.array_equal(r1.cost, r2.cost)),
        "max_abs_intrinsic_error_vs_truth": float(np.abs(x1[0][:, :7] - sp.gt_intrinsics[:, :7]).max()),
        "max_abs_cam_translation_error_mm": float(np.abs(x1[1][:, 3:] - sp.gt_cam_rt[:, 3:]).max()),
    })
    # timing: fixed-iteration mode from the initial point
    s.set_options(capi.default_options(max_num_iterations=a.timed_iterations + 2, disable_tolerances=1))
    s.set_parameters(sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
    s.time_stage(4, 3)
    s.set_parameters(sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
    ms = s.time_stage(4, a.timed_iterations)
    stages = {}
    if world == 1:
        for name, st in (("evaluation_pass", 5), ("schur", 1), ("reduced_solve", 2), ("backsub", 3)):
            stages[name] = s.time_stage(st, 5)
    else:
        tm = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)       # device-timed, max over ranks
        ms = float(tm.item())
    res.update({"ms_per_lm_iteration": ms, "lm_iterations_per_sec": 1e3 / ms,
                "gobs_per_sec": N / ms / 1e6, "stage_ms": stages})
    if world == 1:
        res["fp64_tflops_algorithmic_1060_flop_per_obs"] = N * 1060.0 / (stages["evaluation_pass"] * 1e-3) / 1e12
    ok = (res["cost_monotone_over_accepted_steps"] and res["repeat_solve_bit_identical"] and
          res["termination"] == "CONVERGENCE" and 0.12 < res["rms_px"] < 0.16)
    res["properties_ok"] = bool(ok)
    if rank == 0:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)
        print(json.dumps(res))
    s.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
