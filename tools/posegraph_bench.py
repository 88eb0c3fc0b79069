"""Pose-graph initialisation at BASELINE config 3's size: tscm_pose_graph_init()
on the GPU (kernel time from CUDA events, whole call from the host clock, host buffers in and
out) beside the CPU oracle timed on a bounded sample of candidates.

    python tools/posegraph_bench.py [--frames 5000] [--cfg 3] [--cpu-candidates 24] [--out file.json]
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
    ap.add_argument("--cfg", type=int, default=3)
    ap.add_argument("--frames", type=int, default=5000)
    ap.add_argument("--ring", action="store_true", help="config 3 as the masked 8-camera ring")
    ap.add_argument("--cpu-candidates", type=int, default=24)
    ap.add_argument("--repeats", type=int, default=3)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    kw = dict(dense=False, rig="ring") if a.ring else {}
    sp = synth.config(a.cfg, num_frames=a.frames, **kw)
    args = synth.mono_results(sp, seed=11)
    worlds, intr, has, Rt, px = args
    Cn, F = has.shape
    K = worlds.shape[0]
    capi.pose_graph_init(*args)                                   # module load, first-touch
    wall, kern = [], []
    for _ in range(a.repeats):
        t0 = time.perf_counter()
        r = capi.pose_graph_init(*args)
        wall.append(time.perf_counter() - t0)
        kern.append(r.kernel_ms)
    out = {
        "what": "tscm_pose_graph_init: multi_calib.cpp:6-153 with the candidate scoring on the GPU",
        "workload": f"config {a.cfg}{' ring' if a.ring else ''}: {Cn} cameras x {F} boards x {K} corners, "
                    f"shared boards per adjacent pair {[int((has[i - 1] & has[i]).sum()) for i in range(1, Cn)]}",
        "projections": int(r.projections),
        "kernel_ms": float(np.median(kern)),
        "call_s": float(np.median(wall)),
        "gprojections_per_s_kernel": r.projections / np.median(kern) / 1e6,
        "gprojections_per_s_call": r.projections / np.median(wall) / 1e9,
        "h2d_bytes": int(px.nbytes + Rt.nbytes),
    }
    if a.cpu_candidates > 0 and Cn > 1:
        from oracle import oracle
        rng = np.random.default_rng(1)
        n1 = int((has[0] & has[1]).sum())
        js = rng.choice(np.flatnonzero(has[0] & has[1]), size=min(a.cpu_candidates, n1), replace=False)
        t0 = time.perf_counter()
        same = True
        for j in js:
            e = oracle.pose_pair_error(1, int(j), worlds, intr, has, Rt, px, r.camera_pose[0])
            same &= bool(e == r.camera_candidate_error[1, j])
        dt = time.perf_counter() - t0
        proj = len(js) * n1 * 2 * K
        out["cpu_baseline"] = {
            "kind": "port", "cores": 1,
            "sample": f"{len(js)} of the {n1} candidates of camera 1 (each scored over all {n1} shared boards)",
            "gprojections_per_s": proj / dt / 1e9,
            "seconds_extrapolated_full_job": r.projections / (proj / dt),
            "bit_identical_to_gpu": same,
        }
        out["speedup_call_vs_cpu_1_thread"] = out["gprojections_per_s_call"] / out["cpu_baseline"]["gprojections_per_s"]
    line = json.dumps(out)
    print(line)
    if a.out:
        with open(a.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
