#!/bin/bash
# Fast visit: stage timings only (no tests).
python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from tscm_calib_b200 import capi, synth
sp = synth.config(3)
opt = capi.default_options(max_num_iterations=1000000, disable_tolerances=1)
s = capi.Solver(sp.problem, opt)
s.set_parameters(sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
for stage, name in ((0, "k_eval"), (5, "eval pass"), (1, "schur+reduce"), (2, "solve"), (3, "backsub")):
    s.time_stage(stage, 3)
    print(f" stage {name}: {s.time_stage(stage, 20)*1e3:.1f} us")
s.time_stage(4, 5)
ms = s.time_stage(4, 50)
print(f" LM iteration: {ms*1e3:.1f} us -> {1000.0 / ms:.1f} it/s")
PY
