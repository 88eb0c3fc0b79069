import os, sys, json
import numpy as np
sys.path.insert(0, "/root/repo")
from tscm_calib_b200 import capi, synth
from bench import fixed_iteration_options
sp = synth.config(3)
init = (sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
res = {}
for mode in ("0", "0e", "0", "0e"):
    os.environ["TSCM_SCHUR_U8"] = mode[0]
    os.environ["TSCM_SCHUR_EVEN"] = "1" if mode.endswith("e") else "0"
    s = capi.Solver(sp.problem, fixed_iteration_options(100))
    s.set_parameters(*init); s.time_stage(4, 5); s.set_parameters(*init)
    it = s.time_stage(4, 50)
    s.time_stage(1, 3); sch = s.time_stage(1, 20)
    s.close()
    s = capi.Solver(sp.problem, capi.default_options())
    s.set_parameters(*init); r = s.run(); s.close()
    res.setdefault(mode, []).append((round(it*1e3,2), round(sch*1e3,2), r.num_iterations, r.final_cost))
    print(mode, res[mode][-1])
print("final cost rounded vs even:", res["0"][0][3], res["0e"][0][3])
for cfg, fr in ((2, 200), (5, 100)):
    out = []
    os.environ["TSCM_SCHUR_EVEN"] = "0"
    for mode in ("0", "1"):
        os.environ["TSCM_SCHUR_EVEN"] = mode
        q = synth.config(cfg, num_frames=fr)
        a, b, c, r = capi.solve_resident(q.problem, q.init_intrinsics, q.init_cam_rt, q.init_board_rt, capi.default_options())
        out.append(r.cost.copy())
    print(cfg, "max relative cost difference rounded vs even:", float(np.max(np.abs(out[0] / out[1] - 1))), len(out[0]))
