"""Stage times (CUDA events) of config 3 with --frames F on this rank count: what the exchanges cost is
the difference between N ranks on F frames and one rank on F / N frames.
    python tools/stage_split.py --frames 2500                       # one GPU, the shard's size
    torchrun --nproc-per-node 2 tools/stage_split.py --frames 5000  # two ranks"""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from tscm_calib_b200 import capi, synth
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import fixed_iteration_options

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=5000)
a = ap.parse_args()
world, rank, lr = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
sp = synth.config(3, num_frames=a.frames)
problem, fr = synth.shard_frames(sp, rank, world) if world > 1 else (sp.problem, np.arange(a.frames))
s = capi.Solver(problem, fixed_iteration_options(200), device=lr)
if world > 1:
    capi.attach_ranks(s, rank, world)
s.set_parameters(sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt[fr])
s.time_stage(4, 5)
out = {}
for sid, name in ((4, "iteration"), (5, "evaluation_pass"), (6, "k_eval5"), (7, "k_view_blocks"), (1, "schur"), (2, "reduced_solve"), (3, "backsub")):
    s.time_stage(sid, 3)
    out[name] = round(1e3 * s.time_stage(sid, 20), 2)
if rank == 0:
    print(json.dumps({"n_gpus": world, "frames": a.frames, "frames_this_rank": int(problem.num_frames), "us": out}))
s.close()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
