#!/bin/bash
TSCM_PROF=1 python - <<'PY' 2>&1 | grep -v k_solve | tail -40
import os, sys, time
sys.path.insert(0, os.getcwd())
from tscm_calib_b200 import capi, synth
sp = synth.config(3)
opt = capi.default_options(max_num_iterations=5, disable_tolerances=1)
for rep in range(2):
    t = time.perf_counter()
    s = capi.Solver(sp.problem, opt); t1 = time.perf_counter()
    s.set_parameters(sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt); t2 = time.perf_counter()
    s.run(); t3 = time.perf_counter()
    s.get_parameters(); t4 = time.perf_counter()
    s.close(); t5 = time.perf_counter()
    print(f"rep {rep}: create {1e3*(t1-t):.1f} set {1e3*(t2-t1):.1f} run {1e3*(t3-t2):.1f} get {1e3*(t4-t3):.1f} destroy {1e3*(t5-t4):.1f} ms", file=sys.stderr)
PY
