"""Iteration time for every Schur form (auto = the library's choice) on sparse-visibility problems of
several sizes: the data behind choose_schur_form()."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tscm_calib_b200 import capi, synth
from bench import fixed_iteration_options
cases = [("ring8", dict(idx=3, dense=False, rig="ring"), (150, 600, 2000, 5000)),
         ("cfg5", dict(idx=5), (200, 1000)),
         ("ring16", dict(idx=4), (1000, 5000))]
for name, kw, sizes in cases:
    for F in sizes:
        k = dict(kw); idx = k.pop("idx")
        sp = synth.config(idx, num_frames=F, **k)
        init = (sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
        out = []
        for form in ("auto", "rows", "fused", "pairs"):
            try:
                s = capi.Solver(sp.problem, fixed_iteration_options(200), schur_form=form)
            except Exception as e:
                out.append(f"{form} n/a"); continue
            s.set_parameters(*init); s.time_stage(4, 5); s.set_parameters(*init)
            it = s.time_stage(4, 30)
            out.append(f"{form} {it * 1e3:.1f}")
            s.close()
        print(name, F, "fill", round(float(sp.visible.mean()), 2), "|", " | ".join(out))
