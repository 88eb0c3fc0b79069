"""Multi-GPU parity: frames sharded over the ranks must reproduce the single-GPU solve
(and therefore the oracle) — same iteration count, per-iteration cost to 1e-9 relative."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from tscm_calib_b200 import capi, synth

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
# config 4 = 16-camera ring, sparse visibility: the per-camera-pair Schur path on every rank
CASES = ((2, 200), (3, 600), (4, 2000))
if os.environ.get("DIST_CFGS"):
    CASES = tuple(c for c in CASES if str(c[0]) in os.environ["DIST_CFGS"].split(","))
for cfg, frames in CASES:
    sp = synth.config(cfg, num_frames=frames)
    opt = capi.default_options()
    prob, fr = synth.shard_frames(sp, rank, world)
    s = capi.Solver(prob, opt, device=local)
    capi.attach_ranks(s, rank, world)
    s.set_parameters(sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt[fr])
    res = s.run()
    intr, cam_rt, board_rt = s.get_parameters()
    per_cam, overall, rms = s.reprojection_error()      # global sums / global counts on every rank
    s.close()
    if rank == 0:
        a, b, c, ref = capi.solve(sp.problem, sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt, opt, device=local)
        one = capi.Solver(sp.problem, opt, device=local)
        one.set_parameters(a, b, c)
        per1, overall1, rms1 = one.reprojection_error()
        one.close()
        readout = max(float(np.max(np.abs(per_cam - per1))), abs(overall - overall1), abs(rms - rms1))
        same_iters = res.num_iterations == ref.num_iterations and res.termination == ref.termination
        n = min(len(res.cost), len(ref.cost))
        cost_rel = float(np.max(np.abs(res.cost[:n] - ref.cost[:n]) / ref.cost[:n]))
        p_rel = max(float(np.max(np.abs(intr - a) / (np.abs(a) + 1e-9))),
                    float(np.max(np.abs(cam_rt - b) / (np.abs(b) + 1e-9))),
                    float(np.max(np.abs(board_rt - c[fr]) / (np.abs(c[fr]) + 1e-9))))
        good = same_iters and cost_rel < 1e-9 and p_rel < 1e-7 and readout < 1e-9
        ok = ok and good
        print(f"cfg{cfg} world={world}: iters {res.num_iterations}/{ref.num_iterations} {res.termination} "
              f"cost_rel {cost_rel:.2e} param_rel {p_rel:.2e} read-out diff {readout:.1e} px -> {'OK' if good else 'MISMATCH'}", flush=True)
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("DIST PARITY", "PASS" if ok else "FAIL")
