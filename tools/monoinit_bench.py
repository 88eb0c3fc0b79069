"""Mono cold start at BASELINE config 3's size: tscm_mono_init() for one camera
(5,000 frames x 88 corners) on the GPU beside the CPU oracle on the same frames.

    python tools/monoinit_bench.py [--frames 5000] [--out file.json]
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
    ap.add_argument("--frames", type=int, default=5000)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    sp = synth.config(3, num_frames=a.frames)
    worlds, intr, has, Rt, px = synth.mono_results(sp)
    args = ((11, 8), (1280, 1080), worlds, has[0], px[0])
    capi.mono_init(*args)
    wall, kern = [], []
    for _ in range(3):
        t0 = time.perf_counter()
        r = capi.mono_init(*args)
        wall.append(time.perf_counter() - t0)
        kern.append(r.kernel_ms)
    from oracle import oracle
    t0 = time.perf_counter()
    r0 = oracle.mono_init(*args)
    cpu = time.perf_counter() - t0
    out = {
        "what": "tscm_mono_init: TS.cpp:36-52, 110-203 (focal from circle fits, poses from planar PnP) for one camera",
        "workload": f"{a.frames} frames x 88 corners, cold start",
        "kernel_ms": float(np.median(kern)), "call_s": float(np.median(wall)),
        "frames_per_s_call": a.frames / float(np.median(wall)),
        "focal": float(r.intrinsics[0]), "rows_used": int(r.rows_used), "frames_ok": int(r.frame_ok.sum()),
        "cpu_baseline": {"kind": "port", "cores": 1, "sample": "the same frames", "seconds": cpu,
                         "frames_per_s": a.frames / cpu},
        "max_pose_difference_vs_oracle": float(np.abs(r.Rt - r0.Rt)[:, :, :2].max()),
        "speedup_call_vs_cpu_1_thread": cpu / float(np.median(wall)),
    }
    line = json.dumps(out)
    print(line)
    if a.out:
        with open(a.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
