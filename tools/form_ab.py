"""Iteration time of config 3 on the masked outward ring for every Schur form (auto = the library's choice)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tscm_calib_b200 import capi, synth
from bench import fixed_iteration_options
for rig in ("ring", "array"):
    sp = synth.config(3, dense=False, rig=rig)
    init = (sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
    for form in ("auto", "rows", "fused", "pairs"):
        try:
            s = capi.Solver(sp.problem, fixed_iteration_options(200), schur_form=form)
        except Exception as e:
            print(rig, form, "unavailable:", str(e)[:80]); continue
        s.set_parameters(*init); s.time_stage(4, 5); s.set_parameters(*init)
        it = s.time_stage(4, 30)
        s.time_stage(1, 3); sch = s.time_stage(1, 10)
        print(rig, form, "iteration us", round(it * 1e3, 1), "schur us", round(sch * 1e3, 1), "fill", round(float(sp.visible.mean()), 3))
        s.close()
