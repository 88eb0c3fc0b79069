"""Small solves for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): every kernel of the LM
iteration graph runs at least twice, on sizes a sanitizer finishes in a minute.
  python tools/sanitize_case.py dense|pairs|mono|robust|masks|group|init
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tscm_calib_b200 import capi, synth

case = sys.argv[1] if len(sys.argv) > 1 else "dense"
if case == "init":       # the callers of the solve: mono cold start, pose graph, remap tables
    sp = synth.config(2, num_frames=60)
    worlds, intr, has, Rt, px = synth.mono_results(sp)
    for m in range(2):
        r = capi.mono_init((11, 8), (1280, 1080), worlds, has[m], px[m])
        assert r.frame_ok.sum() == has[m].sum() and r.intrinsics[0] > 0
    g = capi.pose_graph_init(worlds, intr, has, Rt, px)
    assert g.board_init.sum() == has.any(axis=0).sum() and np.all(g.camera_choice[1:] >= 0)
    job = capi.remap_job(intr[0], np.eye(3), (200.0, 200.0, 160.0, 120.0), (320, 240))
    mx, my, _ = capi.remap_tables([job], (320, 240))
    assert np.isfinite(mx).all()
    print(case, "ok", r.kernel_ms, g.kernel_ms)
    sys.exit(0)
if case == "dense":      # config 3 shape: k_eval5, k_view_blocks, k_schur_frames, k_schur_update, k_solve, ...
    sp, opt = synth.config(3, num_frames=96), capi.default_options(max_num_iterations=4)
elif case == "pairs":    # 16-camera ring, sparse visibility: k_pair_frames, k_pair_blocks, k_schur_pairs2, k_reduce_pairs
    sp, opt = synth.config(4, num_frames=160), capi.default_options(max_num_iterations=4)
elif case == "mono":
    sp, opt = synth.config(1), capi.default_options(max_num_iterations=4)
elif case == "masks":    # 8-camera ring with visibility masks, fused Schur kernel (k_schur2), split tail tiles
    sp, opt = synth.config(3, num_frames=150, dense=False, rig="ring"), capi.default_options(max_num_iterations=4)
elif case == "group":    # tscm_options.num_gpus = 2: mailbox exchanges between two devices of one process
    sp, opt = synth.config(3, num_frames=96), capi.default_options(max_num_iterations=4, num_gpus=2)
else:                    # robust loss, ragged visibility (config 5)
    sp, opt = synth.config(5, num_frames=60), capi.default_options(max_num_iterations=4, loss_type="huber", loss_scale=1.0)
if case == "group":
    # under a sanitizer kernels run 10-100x slower: give the exchanges time
    h = capi.Solver(sp.problem, opt)
    h.set_exchange_timeout(600.0)
    h.set_parameters(sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
    s = h.run()
    h.close()
else:
    a, b, c, s = capi.solve(sp.problem, sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt, opt)
print(case, s.termination, s.num_iterations, s.cost)
assert np.all(np.isfinite(s.cost)) and s.final_cost < s.initial_cost     # (the trace also holds rejected candidates)
