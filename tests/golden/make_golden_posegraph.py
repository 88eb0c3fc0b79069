"""Generates tests/golden/pose_graph.npz: golden vectors for MultiCalib's pose-graph
initialisation (/root/reference/multi_calib.cpp:6-153).

A numpy transcription of the constructor — candidate poses of camera i chained through every
board it shares with camera i-1, scored by the summed reprojection error of both cameras
(TS.h:58-69), board poses likewise — including the single-precision r1, r2, r1 x r2 of
Rt_to_R_t (multi_calib.h:130-137, cv::Vec3f).  Input: BASELINE config 2's rig (calib.yaml
cameras) with per-camera mono results = ground-truth board poses in the camera frame plus a
small per-(camera, frame) error, so that the candidates differ and the selection matters.

Run from the repo root:  python tests/golden/make_golden_posegraph.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tscm_calib_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def project(intr, P):
    """TS.cpp:332-344."""
    fx, fy, cx, cy, xi, lam, alpha, b, c = intr
    X, Y, Z = P[:, 0], P[:, 1], P[:, 2]
    d1 = np.sqrt(X * X + Y * Y + Z * Z)
    d2 = np.sqrt(X * X + Y * Y + (Z + xi * d1) ** 2)
    d3 = np.sqrt(X * X + Y * Y + (Z + xi * d1 + lam * d2) ** 2)
    k = Z + xi * d1 + lam * d2 + alpha / (1 - alpha) * d3
    return np.stack([fx * X / k + b * Y / k + cx, c * X / k + fy * Y / k + cy], axis=1)


def reproject_error(intr, pixels, worlds, R, t):
    """TS.h:58-69: SUM of Euclidean errors."""
    q = project(intr, worlds @ R.T + t)
    return float(np.sqrt(((pixels - q) ** 2).sum(axis=1)).sum())


def rt_to_R_t(Rt):
    """multi_calib.h:130-137 with cv::Vec3f arithmetic."""
    r1, r2 = Rt[:, 0].astype(np.float32), Rt[:, 1].astype(np.float32)
    r3 = np.array([r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]],
                  dtype=np.float32)
    return np.stack([r1, r2, r3], axis=1).astype(np.float64), Rt[:, 2].copy()


def pose_graph(intr, Rt, has, pixels, worlds):
    C, F = has.shape
    cam_R, cam_t = np.zeros((C, 3, 3)), np.zeros((C, 3))
    # candidate scores and the winners, for the scoring kernels' parity tests
    cam_choice, cam_err = np.full(C, -1, dtype=np.int32), np.full((C, F), np.nan)
    board_choice, board_err = np.full(F, -1, dtype=np.int32), np.full((F, C), np.nan)
    for i in range(C):
        if i == 0:
            cam_R[0], cam_t[0] = np.eye(3), 0.0
            continue
        Rk, tk = cam_R[i - 1], cam_t[i - 1]
        Rs, ts, src = [], [], []
        for j in range(F):
            if not (has[i - 1, j] and has[i, j]):
                continue
            src.append(j)
            Ri, ti = rt_to_R_t(Rt[i, j])
            Rp, tp = rt_to_R_t(Rt[i - 1, j])
            R_ik = Ri @ Rp.T
            t_ik = ti - R_ik @ tp
            Rs.append(R_ik @ Rk)
            ts.append(R_ik @ tk + t_ik)
        best, best_id = 1e10, -1
        for c in range(len(Rs)):
            err = 0.0
            for k in range(F):
                if not (has[i - 1, k] and has[i, k]):
                    continue
                Ri, ti = rt_to_R_t(Rt[i, k])
                R_ki = Rk @ Rs[c].T
                t_ki = tk - R_ki @ ts[c]
                err += reproject_error(intr[i - 1], pixels[i - 1, k], worlds, R_ki @ Ri, R_ki @ ti + t_ki)
                Rp, tp = rt_to_R_t(Rt[i - 1, k])
                R_ik = Rs[c] @ Rk.T
                t_ik = ts[c] - R_ik @ tk
                err += reproject_error(intr[i], pixels[i, k], worlds, R_ik @ Rp, R_ik @ tp + t_ik)
            cam_err[i, src[c]] = err
            if err < best:
                best, best_id = err, c
        cam_R[i], cam_t[i] = Rs[best_id], ts[best_id]
        cam_choice[i] = src[best_id]
    board_R, board_t, board_init = np.zeros((F, 3, 3)), np.zeros((F, 3)), np.zeros(F, dtype=np.uint8)
    for i in range(F):
        ids = [j for j in range(C) if has[j, i]]
        if not ids:
            continue
        Rs, ts = [], []
        for j in ids:
            Rb, tb = rt_to_R_t(Rt[j, i])
            Rs.append(cam_R[j].T @ Rb)
            ts.append(cam_R[j].T @ (tb - cam_t[j]))
        best_id = 0
        if len(ids) > 1:
            best = 1e10
            for c in range(len(Rs)):
                err = 0.0
                for j in ids:
                    err += reproject_error(intr[j], pixels[j, i], worlds, cam_R[j] @ Rs[c], cam_R[j] @ ts[c] + cam_t[j])
                board_err[i, ids[c]] = err
                if err < best:
                    best, best_id = err, c
        board_R[i], board_t[i], board_init[i] = Rs[best_id], ts[best_id], 1
        board_choice[i] = ids[best_id]
    return cam_R, cam_t, board_R, board_t, board_init, dict(cam_choice=cam_choice, cam_err=cam_err,
                                                            board_choice=board_choice, board_err=board_err)


def main():
    sp = synth.config(2, num_frames=40)
    p = sp.problem
    C, F, K = p.num_cameras, p.num_frames, p.corners_per_board
    has = sp.visible.astype(np.uint8)
    pixels = np.zeros((C, F, K, 2))
    pixels[p.view_camera, p.view_frame] = p.obs_xy
    worlds = np.concatenate([p.board_xy, np.zeros((K, 1))], axis=1)
    rng = np.random.default_rng(42)
    Rc, Rb = synth.rodrigues(sp.gt_cam_rt[:, :3]), synth.rodrigues(sp.gt_board_rt[:, :3])
    Rt = np.zeros((C, F, 3, 3))
    for m in range(C):
        for i in range(F):
            if not has[m, i]:
                continue
            R = synth.rodrigues(rng.normal(0, 2e-3, 3)) @ Rc[m] @ Rb[i]
            t = Rc[m] @ sp.gt_board_rt[i, 3:] + sp.gt_cam_rt[m, 3:] + rng.normal(0, 1.0, 3)
            Rt[m, i] = np.stack([R[:, 0], R[:, 1], t], axis=1)
    intr = sp.gt_intrinsics.copy()
    cam_R, cam_t, board_R, board_t, board_init, scores = pose_graph(intr, Rt, has, pixels, worlds)
    err_t = np.abs(cam_t - sp.gt_cam_rt[:, 3:]).max()
    print("cameras: max |t - truth| =", err_t, "mm; boards initialised:", int(board_init.sum()), "of", F)
    np.savez_compressed(os.path.join(OUT, "pose_graph.npz"), board=np.array([11, 8]), square=45.0, has=has,
                        pixels=pixels, intrinsics=intr, Rt=Rt, cam_R=cam_R, cam_t=cam_t, board_R=board_R,
                        board_t=board_t, board_init=board_init, gt_cam_rt=sp.gt_cam_rt, gt_board_rt=sp.gt_board_rt, **scores)


if __name__ == "__main__":
    main()
