#!/bin/bash
TSCM_PROF=2 python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from tscm_calib_b200 import capi, synth
sp = synth.config(3, num_frames=int(os.environ.get("FRAMES", "5000")))
opt = capi.default_options(max_num_iterations=1, disable_tolerances=1)
s = capi.Solver(sp.problem, opt)
s.set_parameters(sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
print(s.run())
PY
