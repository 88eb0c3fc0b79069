"""Generates tests/golden/mono_init.npz: golden vectors for the mono cold-start
initialisation, produced with the REAL OpenCV (cv2 wheel in this image)
at the two calls the reference delegates to it:

  cv::SVD::solveZ       /root/reference/TS.cpp:142-143  (estimate_focal)
  cv::solvePnPRansac    /root/reference/TS.cpp:193      (estimate_extrinsic)

Everything around those calls (TS.cpp:36-52, 110-203 and TS.h:39-57) is transcribed below in
numpy in the reference's order.  The C++ adapter (tscm_calib_b200/host/ts_camera.cpp with the
shim's own solveZ / planar PnP) must reproduce these numbers to solver tolerance.

Run from the repo root:  python tests/golden/make_golden_init.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tscm_calib_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def solve_z(P):
    """cv::SVD::solveZ: last row of vt."""
    _, _, vt = cv2.SVDecomp(P)
    return vt[-1].copy()


def unit_sphere(intr, px, T=np.eye(3)):
    """TS.h:39-57."""
    fx, fy, cx, cy, xi, lam, alpha, b, c = intr
    x, y = px[0] - cx, px[1] - cy
    mx = (fy * x - b * y) / (fx * fy - b * c)
    my = (-c * x + fx * y) / (fx * fy - b * c)
    k = alpha / (1 - alpha)
    r2 = mx * mx + my * my
    gamma = (k + np.sqrt(1 + (1 - k * k) * r2)) / (r2 + 1)
    eta = lam * (gamma - k) + np.sqrt(((gamma - k) * (gamma - k) - 1) * lam * lam + 1)
    mz = eta * (gamma - k)
    mu = xi * (mz - lam) + np.sqrt(xi * xi * ((mz - lam) * (mz - lam) - 1) + 1)
    p = np.array([mu * eta * gamma * mx, mu * eta * gamma * my, mu * (mz - lam) - xi])
    return T @ p


def estimate_focal(pixels, has, cx, cy, W, H):
    """TS.cpp:110-168."""
    focal, total = 0.0, 0
    for k in range(len(pixels)):
        if not has[k]:
            continue
        pc = pixels[k] - np.array([cx, cy])
        for i in range(H):
            row = pc[i * W:(i + 1) * W]
            x, y = row[:, 0], row[:, 1]
            P = np.stack([x, y, np.full(W, 0.5), -0.5 * (x * x + y * y)], axis=1)
            c1, c2, c3, c4 = solve_z(P)
            t = c1 * c1 + c2 * c2 + c3 * c4
            if t < 0:
                continue
            d = np.sqrt(1 / t)
            nx, ny = c1 * d, c2 * d
            if nx * nx + ny * ny > 0.95:
                continue
            nz = np.sqrt(1 - nx * nx - ny * ny)
            focal += abs(c3 * d / nz)
            total += 1
    return (focal / total if total else 0.0), total


def estimate_extrinsic(intr, pixels, has, worlds, W):
    """TS.cpp:170-203 with the real cv2.solvePnPRansac (default arguments)."""
    K = len(worlds)
    out = np.zeros((len(pixels), 3, 3))
    inliers = np.zeros(len(pixels), dtype=np.int64)
    for k in range(len(pixels)):
        if not has[k]:
            continue
        p = unit_sphere(intr, pixels[k][K // 2 - W // 2 - 1])
        a, b = np.arctan2(p[0], p[2]), np.arcsin(p[1])
        R1 = np.array([[np.cos(a), 0, -np.sin(a)], [0, 1, 0], [np.sin(a), 0, np.cos(a)]])
        R2 = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
        T = R2 @ R1
        pn = np.zeros((K, 2))
        for i in range(K):
            q = unit_sphere(intr, pixels[k][i], T)
            pn[i] = q[:2] / q[2]
        ok, rvec, tvec, inl = cv2.solvePnPRansac(worlds.reshape(-1, 1, 3), pn.reshape(-1, 1, 2), np.eye(3), None)
        # corners whose back-projection is NaN under the current guess (TS.h:47) are rejected by
        # RANSAC; every finite corner must be an inlier for the fixture to be unambiguous
        assert ok and len(inl) == int(np.isfinite(pn).all(axis=1).sum()), (k, ok, len(inl))
        inliers[k] = len(inl)
        Rm, _ = cv2.Rodrigues(rvec)
        Rt = T.T @ Rm
        Rt[:, 2] = (T.T @ tvec).reshape(3)
        out[k] = Rt
    return out, inliers


def main():
    sp = synth.config(1)                      # 1 camera, 40 poses, 9x6 board, calib.yaml cam0
    p = sp.problem
    W, H, K = 9, 6, 54
    F = p.num_frames + 2                      # plus two frames without a detection
    has = np.ones(F, dtype=np.uint8)
    has[[5, 17]] = 0
    pixels = np.zeros((F, K, 2))
    pixels[has == 1] = p.obs_xy.reshape(p.num_frames, K, 2)
    worlds = np.concatenate([p.board_xy.reshape(K, 2), np.zeros((K, 1))], axis=1)
    img = (1280, 1080)
    # cold start: TS.cpp:43-47
    cx, cy = img[0] // 2 - 0.5, img[1] // 2 - 0.5
    focal, rows_used = estimate_focal(pixels, has, cx, cy, W, H)
    intr_cold = np.array([focal, focal, cx, cy, 0.0, 0.0, 0.5, 0.0, 0.0])
    Rt_cold, inl_cold = estimate_extrinsic(intr_cold, pixels, has, worlds, W)
    # warm start from the 7-argument constructor: extrinsics only (TS.cpp:52)
    guess7 = sp.init_intrinsics[0, :7].copy()
    Rt_warm, inl_warm = estimate_extrinsic(np.concatenate([guess7, [0.0, 0.0]]), pixels, has, worlds, W)
    # stand-alone vector for solveZ
    rng = np.random.default_rng(3)
    A = rng.standard_normal((11, 4)) @ np.diag([3.0, 1.0, 0.2, 1e-3])
    z = solve_z(A)
    print("frames with rejected (NaN) corners: cold", int((inl_cold[has == 1] < K).sum()), "warm",
          int((inl_warm[has == 1] < K).sum()))
    print("focal", focal, "rows used", rows_used, "of", int(has.sum()) * H, "| ground truth fx",
          synth.CALIB_INTRINSICS[0, 0])
    np.savez_compressed(os.path.join(OUT, "mono_init.npz"), pixels=pixels, has=has, board=np.array([W, H]),
                        square=45.0, image=np.array(img), worlds=worlds, intr_cold=intr_cold, Rt_cold=Rt_cold,
                        guess7=guess7, Rt_warm=Rt_warm, inliers_cold=inl_cold, inliers_warm=inl_warm, solvez_A=A, solvez_z=z)


if __name__ == "__main__":
    main()
