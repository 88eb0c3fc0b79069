#!/usr/bin/env python
"""bench.py — LM iterations/s and residual+Jacobian Gobs/s of the calibration solve.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--scaling weak|strong]

A "step" is ONE Levenberg-Marquardt iteration (Schur elimination, reduced solve,
back-substitution, residual+Jacobian+normal-equation pass at the candidate point,
accept/reject) over BASELINE.json config 3: 8 cameras x 5,000 frames x 88 corners =
3.52 M observations per GPU, synthetic, generated through the TS model.

Default scaling is "weak": every rank owns 5,000 frames (its own frames, one common
rig), the reduced camera system is all-reduced over NCCL each iteration; `value` is
the whole-job residual+Jacobian throughput in Gobs/s = observations on all ranks x
LM iterations / s.  `--scaling strong` shards the 5,000 frames of ONE config-3
problem over the ranks instead.

The JSON line also carries:
  lm_iterations_per_sec   the other half of BASELINE.json's metric
  roofline                the dominant kernel (k_eval) against measured HBM peak, plus
                          roofline_fp64 against the measured DFMA peak (the kernel is
                          FP64-pipe bound)
  cpu_baseline            the CPU oracle (Ceres-semantics port) timed on this box
  e2e                     the same metric through tscm_solve() with host buffers

`--impl reference` times the CPU restatement of the reference's Ceres path (the
reference itself cannot be built: Ceres/Eigen/OpenCV are absent) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = "config3: 8 cameras x 5000 frames x 88 corners (3.52M observations), dense visibility"
METRIC = "LM residual+Jacobian throughput at 3.52M corner observations per GPU"
UNIT = "Gobs/s"
ALG_BYTES_PER_OBS_EVAL = 24.0   # J-pass with per-view blocks materialised
ALG_FLOPS_PER_OBS_EVAL = 1060.0  # forward 130 + Jacobian 250 + outer products 680


def captured_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture
    (profiles/r*_traffic.json, written by tools/make_profile_summary.py); None if absent."""
    import glob
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json"))):
        try:
            with open(path) as f:
                d = json.load(f)
            if kernel in d:
                best = d[kernel]["dram_bytes"]
        except Exception:
            pass
    return best


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 8:
                for name, val in zip(names, r[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(rank: int, world: int, scaling: str, frames: int):
    from tscm_calib_b200 import synth
    if scaling == "strong" and world > 1:
        sp = synth.config(3, num_frames=frames)
        local, fr = synth.shard_frames(sp, rank, world)
        return (local, sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt[fr],
                sp.num_observations)
    sp = synth.config(3, num_frames=frames, frame_seed=rank)
    return (sp.problem, sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt,
            sp.num_observations * world)


def fixed_iteration_options(n):
    from tscm_calib_b200 import capi
    return capi.default_options(max_num_iterations=int(n), disable_tolerances=1)


def time_oracle(problem, intr, cam_rt, board_rt, iterations, threads):
    from oracle import oracle
    opt = fixed_iteration_options(iterations)
    t0 = time.perf_counter()
    _, _, _, s = oracle.solve(problem, intr, cam_rt, board_rt, opt, num_threads=threads)
    dt = time.perf_counter() - t0
    assert s.num_iterations == iterations + 1, s
    return dt


def run_reference(args):
    """The reference arm: CPU restatement of the reference's Ceres path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tscm_calib_b200 import synth
    threads = os.cpu_count() or 1
    frames = args.reference_frames
    sp = synth.config(3, num_frames=frames)
    if args.warmup > 0:
        time_oracle(sp.problem, sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt,
                    min(args.warmup, 1), threads)
    dt = time_oracle(sp.problem, sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt, args.steps,
                     threads)
    gobs = sp.num_observations * args.steps / dt / 1e9
    sample = (f"{frames} of 5000 frames of config 3 ({sp.num_observations} observations), "
              f"{args.steps} LM iterations incl. iteration-0 evaluation, {threads} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": gobs, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "lm_iterations_per_sec_at_sample": args.steps / dt,
        "lm_iterations_per_sec_at_3.52M": gobs * 1e9 / 3.52e6,
        "cpu_baseline": {"value": gobs, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": gobs, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist
    from tscm_calib_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the calibration solve has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    capi.load_library()

    problem, intr, cam_rt, board_rt, total_obs = make_workload(rank, world, args.scaling, args.frames)
    total_iters = args.warmup + args.steps
    opt = fixed_iteration_options(total_iters + 8)
    solver = capi.Solver(problem, opt, device=local_rank)
    if world > 1:
        capi.attach_ranks(solver, rank, world)
    solver.set_parameters(intr, cam_rt, board_rt)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up, then EXACTLY `steps` LM iterations, device-timed ------------------
    if args.warmup > 0:
        solver.time_stage(4, args.warmup)
        solver.set_parameters(intr, cam_rt, board_rt)     # restart from the same initial point
    barrier()
    launches0 = solver.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_iter = solver.time_stage(4, args.steps)      # CUDA events on the solver's stream
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = solver.launch_count() - launches0
    t = torch.tensor([ms_iter], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_iter = float(t.item())
    it_per_s = 1e3 / ms_iter
    value = total_obs * it_per_s / 1e9

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_iter, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "observations_total": int(total_obs),
                   "observations_per_gpu": int(problem.num_observations),
                   "step": "one LM iteration (Schur + reduced solve + back-substitution + "
                           "residual/Jacobian/normal-equation pass + accept/reject)",
                   "l2": "per-step working set (observations 56 MB, moment buffer 60 MB, per-view "
                         "records 2 x 34 MB, Schur rows 48 MB) exceeds the 126 MB L2; no explicit flush",
                   "parallelism": (f"frames sharded over {world} GPU(s); the reduced camera system and "
                                   "the evaluation record are exchanged by "
                                   + ("this library's kernels over NVLink peer memory (CUDA IPC)"
                                      if os.environ.get("TSCM_P2P", "1") != "0" and world <= 8
                                      else "NCCL all-reduce")) if world > 1 else "single GPU"},
        "lm_iterations_per_sec": it_per_s,
        "gpu_launches": int(launches),
        "clocks": clocks,
    }

    # ---- roofline of the dominant kernel, timed alone with CUDA events ----------------------
    # The J-pass (residual + analytic Jacobian + normal-equation blocks) is two kernels:
    # k_eval5 (persistent, per-view moments; the dominant kernel of the iteration) and
    # k_view_blocks (per-view blocks from the moments).  TSCM_EVAL_VARIANT=4 selects the
    # older single-kernel form k_eval4, for which stage 7 is empty.
    # (every rank runs the same stage sequence: the set-up in front of a timed stage
    # contains collectives; only rank 0 reports)
    variant = os.environ.get("TSCM_EVAL_VARIANT", "5")
    main_kernel = "k_eval5" if variant not in ("3", "4") else "k_eval" + variant
    solver.time_stage(0, 3)
    ms_pass = solver.time_stage(0, 20)
    ms_main = solver.time_stage(6, 20)
    ms_blocks = solver.time_stage(7, 20) if main_kernel == "k_eval5" else 0.0
    stages = {}
    for sid, name in ((5, "evaluation_pass"), (1, "schur"), (2, "reduced_solve"), (3, "backsub")):
        solver.time_stage(sid, 2)
        stages[name] = solver.time_stage(sid, 10)
    if rank == 0:
        n_obs = problem.num_observations
        hbm_peak, how = measured_peaks()
        achieved = n_obs * ALG_BYTES_PER_OBS_EVAL / (ms_main * 1e-3) / 1e9
        line["roofline"] = {
            "kernel": main_kernel, "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
            "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": captured_traffic(main_kernel),
            "peak_source": how, "ms_per_launch": ms_main,
            "algorithmic_bytes_per_observation": ALG_BYTES_PER_OBS_EVAL,
            "note": f"{main_kernel} is FP64-pipe bound (arithmetic intensity ~44 flop/B vs machine "
                    "balance ~5.3 flop/B); see roofline_fp64.  traffic = DRAM bytes of one launch "
                    "from the committed ncu capture (profiles/)",
        }
        fp64_peak = capi.device_fp64_peak(local_rank)
        tf = n_obs * ALG_FLOPS_PER_OBS_EVAL / (ms_pass * 1e-3) / 1e12
        line["roofline_fp64"] = {
            "kernel": main_kernel + ("+k_view_blocks" if ms_blocks else ""), "bound": "fp64",
            "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak,
            "peak_source": "measured in this run (DFMA microbenchmark, tscm_device_fp64_peak)",
            "algorithmic_flops_per_observation": ALG_FLOPS_PER_OBS_EVAL,
            "ms_per_pass": ms_pass, "ms_main_kernel": ms_main, "ms_view_blocks": ms_blocks,
            "note": "SURVEY 8(d)'s 1,060 flop/observation J-pass (projection, Jacobian, normal-equation "
                    "blocks) over the time of the whole pass (both kernels)",
        }
        line["stage_ms"] = stages

    # ---- end to end through the public C-ABI with HOST buffers ------------------------------
    # The call sequence a reference adapter makes for one calibration on a resident solver:
    # upload observations (pinned host memory) and initial parameters, run the LM loop, read
    # the parameters and the summary back.  Solver/communicator creation is outside the timed
    # region (one per process); the one-shot tscm_solve(), which also pays allocation and
    # teardown, is reported beside it at N = 1.
    solver.close()
    e2e_iters = args.steps
    pinned = torch.empty(problem.obs_xy.shape, dtype=torch.float64).pin_memory()
    pinned.copy_(torch.from_numpy(problem.obs_xy))
    host_problem = capi.ProblemArrays(problem.board_xy, problem.view_camera, problem.view_frame,
                                      pinned.numpy(), problem.num_cameras, problem.num_frames,
                                      problem.fixed_camera)
    e_opt = fixed_iteration_options(e2e_iters)
    s2 = capi.Solver(host_problem, e_opt, device=local_rank)
    if world > 1:
        capi.attach_ranks(s2, rank, world)

    def one_call():
        s2.set_observations(host_problem.obs_xy)
        s2.set_parameters(intr, cam_rt, board_rt)
        r = s2.run()
        return r, s2.get_parameters()

    one_call()                                   # warm-up call
    barrier()
    t0 = time.perf_counter()
    res, (a, b, c) = one_call()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    s2.close()
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    assert res.num_iterations == e2e_iters + 1
    h2d = problem.obs_xy.nbytes + intr.nbytes + cam_rt.nbytes + board_rt.nbytes
    d2h = a.nbytes + b.nbytes + c.nbytes + 5 * 8 * (e2e_iters + 1)
    line["e2e"] = {
        "value": total_obs * e2e_iters / dt / 1e9, "unit": UNIT,
        "h2d_bytes_per_step": h2d / e2e_iters, "d2h_bytes_per_step": d2h / e2e_iters,
        "lm_iterations_per_sec": e2e_iters / dt, "seconds_per_call": dt,
        "what": f"one calibration on a resident solver = {e2e_iters} LM iterations: H2D of the "
                "observations (pinned host memory) and initial parameters, on-device transposition, "
                "iteration zero, the iterations, D2H of parameters and trace; inputs are uploaded "
                "once per call, bytes are per call / steps; max over ranks",
    }
    if world == 1:
        capi.solve(host_problem, intr, cam_rt, board_rt, e_opt, device=local_rank)   # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        capi.solve(host_problem, intr, cam_rt, board_rt, e_opt, device=local_rank)
        dt1 = time.perf_counter() - t0
        line["e2e"]["one_shot_tscm_solve"] = {
            "lm_iterations_per_sec": e2e_iters / dt1, "seconds_per_call": dt1,
            "what": "tscm_solve(): additionally creates and destroys the solver (cudaMalloc/Free of "
                    "~0.4 GB, index tables, CUDA graph capture) inside the timed region"}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from tscm_calib_b200 import synth
        threads = os.cpu_count() or 1
        frames = args.cpu_frames
        sp = synth.config(3, num_frames=frames)
        iters = args.cpu_iterations      # ~10-20 s of CPU work on 16 host threads
        dt = time_oracle(sp.problem, sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt, iters,
                         threads)
        line["cpu_baseline"] = {
            "value": sp.num_observations * iters / dt / 1e9, "unit": UNIT, "cores": threads,
            "kind": "port",
            "sample": f"{frames} of 5000 frames of config 3 ({sp.num_observations} observations), "
                      f"{iters} LM iterations incl. iteration-0 evaluation, Ceres-semantics oracle "
                      f"(dual-number autodiff, dense Schur), {threads} threads, {dt:.1f} s",
            "lm_iterations_per_sec_at_3.52M": sp.num_observations * iters / dt / 3.52e6,
        }
        # the reference never sets Solver::Options::num_threads (Ceres default 1): the same
        # port on ONE thread is the faithful picture of what a reference user runs today
        sp1 = synth.config(3, num_frames=max(1, frames // 10))
        it1 = max(2, iters // 10)
        dt1 = time_oracle(sp1.problem, sp1.init_intrinsics, sp1.init_cam_rt, sp1.init_board_rt, it1, 1)
        line["cpu_baseline_single_thread"] = {
            "value": sp1.num_observations * it1 / dt1 / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{sp1.problem.num_frames} of 5000 frames of config 3 ({sp1.num_observations} "
                      f"observations), {it1} LM iterations, 1 thread (Ceres' default num_threads, which the "
                      f"reference leaves in force), {dt1:.1f} s",
            "lm_iterations_per_sec_at_3.52M": sp1.num_observations * it1 / dt1 / 3.52e6,
        }
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--frames", type=int, default=5000, help="frames of config 3 (per GPU if weak)")
    ap.add_argument("--cpu-iterations", type=int, default=30, help="cpu_baseline LM iterations")
    ap.add_argument("--cpu-frames", type=int, default=5000, help="cpu_baseline sample size (frames)")
    ap.add_argument("--reference-frames", type=int, default=5000, help="--impl reference sample size (frames)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_b200(args)


if __name__ == "__main__":
    main()
