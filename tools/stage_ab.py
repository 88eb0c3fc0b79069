"""Iteration and Schur-stage time of config 3 for the library named by TSCM_LIB_PATH (A/B builds)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tscm_calib_b200 import capi, synth
from bench import fixed_iteration_options
sp = synth.config(3)
init = (sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
s = capi.Solver(sp.problem, fixed_iteration_options(200))
for rep in range(2):
    s.set_parameters(*init); s.time_stage(4, 5); s.set_parameters(*init)
    it = s.time_stage(4, 50)
    s.time_stage(1, 3); sch = s.time_stage(1, 20)
    print(os.environ.get("TSCM_LIB_PATH", "default"), "iteration us", round(it * 1e3, 2), "schur us", round(sch * 1e3, 2))
s.close()
