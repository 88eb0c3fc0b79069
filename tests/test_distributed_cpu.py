"""N > 1 host-side logic on CPU (gloo, world_size 2): frame sharding + the sums the GPUs
all-reduce.  Each rank evaluates ITS frames with the oracle, the per-camera blocks and the
cost are summed across ranks with torch.distributed and must equal the unsharded problem —
the identity the multi-GPU path relies on.  Also exercises the 128-byte
unique-id broadcast the GPU path uses to bootstrap its NCCL communicator."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tscm_calib_b200 import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _camera_blocks(problem, intr, cam_rt, board_rt):
    """Per-camera J_c^T J_c (15x15), J_c^T r (15) and the cost from the oracle's Jacobian."""
    from oracle import oracle
    r, J, cost = oracle.eval_jacobian(problem, intr, cam_rt, board_rt)
    C, K = problem.num_cameras, problem.corners_per_board
    cam = np.repeat(problem.view_camera, K)
    Jc = np.concatenate([J[:, :, 0:6], J[:, :, 12:21]], axis=2)       # camera_rt | intrinsic
    U, g = np.zeros((C, 15, 15)), np.zeros((C, 15))
    for m in range(C):
        A = Jc[cam == m].reshape(-1, 15)
        U[m] = A.T @ A
        g[m] = A.T @ r[cam == m].reshape(-1)
    return U, g, cost


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sp = synth.config(2, num_frames=40)
    local, frames = synth.shard_frames(sp, rank, world)
    U, g, cost = _camera_blocks(local, sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt[frames])
    t = torch.from_numpy(np.concatenate([U.ravel(), g.ravel(), [cost, float(local.num_views)]]))
    dist.all_reduce(t)
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.arange(128, dtype=torch.uint8)
    dist.broadcast(uid, 0)
    if rank == 0:
        np.save(out, t.numpy())
    assert uid.tolist() == list(range(128))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_sharded_sums_equal_the_unsharded_problem(tmp_path, world):
    out = str(tmp_path / "sum.npy")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = np.load(out)
    sp = synth.config(2, num_frames=40)
    U, g, cost = _camera_blocks(sp.problem, sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
    ref = np.concatenate([U.ravel(), g.ravel(), [cost, float(sp.problem.num_views)]])
    scale = np.abs(ref).max()
    np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-12 * scale)
    assert got[-1] == sp.problem.num_views
