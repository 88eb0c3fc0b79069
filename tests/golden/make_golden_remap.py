"""Generates tests/golden/remap_tables.npz: golden CV_32FC1 remap tables for §8(f) #4.

An INDEPENDENT numpy float64 transcription of the reference's loops (numpy element-wise
arithmetic rounds every operation once and never contracts to FMA, exactly the arithmetic
of the scalar C++ loops):
  TripleSphereCamera::undistort            /root/reference/TS.cpp:284-306
  TripleSphereCamera::undistort_chessboard /root/reference/TS.cpp:308-330
  Remap::init_remap                        /root/reference/EpipolarRectify/rectify.cpp:86-199
on the reference's own fixture EpipolarRectify/calib.yaml (read here with cv2.FileStorage when
/root/reference is present; the values equal tscm_calib_b200.synth.CALIB_*).  Neither
oracle/remap_oracle.c nor the CUDA kernel is involved.  cv::Mat's 3x3 product is transcribed as
a left-to-right row sum (OpenCV's gemm order for 3x3 is not observable here: no OpenCV C++).

The tables are large (4 x 400x1600 + 2 x 1280x1080 floats), so the fixture stores a SHA-256 of
every table plus a strided sample.

Run from the repo root:  python tests/golden/make_golden_remap.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tscm_calib_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SAMPLE = 37       # stride of the stored sample (flattened table)


def calib_yaml():
    path = "/root/reference/EpipolarRectify/calib.yaml"
    cams, twcs = synth.CALIB_INTRINSICS.copy(), synth.CALIB_TWC.copy()
    if os.path.exists(path):
        import cv2
        fs = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
        for m in range(4):
            assert np.array_equal(fs.getNode(f"cam{m}").mat().reshape(9), cams[m])
            assert np.array_equal(fs.getNode(f"Twc{m}").mat(), twcs[m])
        fs.release()
    return cams, twcs


def project(intr, X, Y, Z, w2=0.0):
    """TS.cpp:332-344 / rectify.cpp:22-36, element-wise."""
    fx, fy, cx, cy, xi, lam, alpha, b, c = (float(v) for v in intr)
    d1 = np.sqrt(X * X + Y * Y + Z * Z)
    t1 = Z + xi * d1
    d2 = np.sqrt(X * X + Y * Y + t1 * t1)
    t2 = Z + xi * d1 + lam * d2
    d3 = np.sqrt(X * X + Y * Y + t2 * t2)
    ksai = Z + xi * d1 + lam * d2 + alpha / (1 - alpha) * d3
    u = fx * X / ksai + b * Y / ksai + cx
    v = c * X / ksai + fy * Y / ksai + cy
    if w2 > 0:
        bad = Z <= -w2 * d1
        u = np.where(bad, -1.0, u)
        v = np.where(bad, -1.0, v)
    return u, v


def grid(width, height, fx, fy, cx, cy):
    j = np.arange(width, dtype=np.float64)[None, :].repeat(height, 0)
    i = np.arange(height, dtype=np.float64)[:, None].repeat(width, 1)
    return (j - cx) / fx, (i - cy) / fy


def apply(M, x, y):
    M = np.asarray(M, dtype=np.float64).reshape(3, 3)
    one = np.ones_like(x)
    return tuple(M[r, 0] * x + M[r, 1] * y + M[r, 2] * one for r in range(3))


def undistort(intr, fx, fy, cx, cy, size):
    x, y = grid(size[0], size[1], fx, fy, cx, cy)
    u, v = project(intr, x, y, np.ones_like(x))
    return u.astype(np.float32), v.astype(np.float32)


def chessboard(intr, Rt, board, square):
    size = (int((board[0] + 1) * square), int((board[1] + 1) * square))
    x, y = grid(size[0], size[1], 1.0, 1.0, square, square)
    u, v = project(intr, *apply(Rt, x, y))
    return u.astype(np.float32), v.astype(np.float32)


def normalize(v):
    n = np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    return v if n == 0 else v / n


def calc_R(t1, t2):
    x = normalize(np.array([t2[0] - t1[0], t2[1] - t1[1], t2[2] - t1[2]]))
    z = normalize(np.array([-x[2], 0.0, x[0]]))
    y = normalize(np.array([-z[2] * x[1] + z[1] * x[2], z[2] * x[0] - z[0] * x[2], -z[1] * x[0] + z[0] * x[1]]))
    return np.stack([x, y, z], axis=1)


def matmul_t(A, B):
    return np.array([[A[0, i] * B[0, j] + A[1, i] * B[1, j] + A[2, i] * B[2, j] for j in range(3)]
                     for i in range(3)])


def init_remap(cams, twcs):
    R = [t[:, :3] for t in twcs]
    tv = [t[:, 3] for t in twcs]
    S = 400
    maps = {k: np.zeros((4 * S, S), dtype=np.float32) for k in ("left_mapx", "left_mapy", "right_mapx", "right_mapy")}
    x, y = grid(S, S, 200.0, 200.0, 200.0, 200.0)
    F, Rt, Re, L = 0, 1, 2, 3
    off = {F: (0, 0), Rt: (1280, 0), Re: (0, 1080), L: (1280, 1080)}
    for blk, (a, b) in enumerate([(F, Rt), (Rt, Re), (Re, L), (L, F)]):
        Rab = calc_R(tv[a], tv[b])
        for cam, side in ((a, "left"), (b, "right")):
            u, v = project(cams[cam], *apply(matmul_t(R[cam], Rab), x, y), w2=0.42399)
            ox, oy = off[cam]
            maps[side + "_mapx"][blk * S:(blk + 1) * S] = (u + ox if ox else u).astype(np.float32)
            maps[side + "_mapy"][blk * S:(blk + 1) * S] = (v + oy if oy else v).astype(np.float32)
    return maps


def main():
    cams, twcs = calib_yaml()
    tables = {}
    tables["undistort_x"], tables["undistort_y"] = undistort(cams[0], 300.0, 300.0, 639.5, 539.5, (1280, 1080))
    # a frame pose in [r1 r2 t] form: board 500 mm ahead, tilted
    rv = np.array([0.2, -0.3, 0.1])
    Rm = synth.rodrigues(rv)
    Rt = np.stack([Rm[:, 0], Rm[:, 1], np.array([-220.0, -160.0, 500.0])], axis=1)
    tables["board_x"], tables["board_y"] = chessboard(cams[1], Rt, (11, 8), 45.0)
    tables.update(init_remap(cams, twcs))
    # a block that looks 100 degrees off-axis so that the validity cut-off of rectify.cpp:28 fires,
    # placed inside a larger table with a mosaic offset
    Mc = synth.rodrigues(np.array([0.0, np.deg2rad(100.0), 0.0]))
    x, y = grid(256, 192, 100.0, 100.0, 127.5, 95.5)
    u, v = project(cams[2], *apply(Mc, x, y), w2=0.42399)
    cut_x, cut_y = np.full((300, 320), 7.0, np.float32), np.full((300, 320), 7.0, np.float32)
    cut_x[50:242, 32:288] = (u + 1280).astype(np.float32)
    cut_y[50:242, 32:288] = v.astype(np.float32)
    tables["cutoff_x"], tables["cutoff_y"] = cut_x, cut_y
    out = {"board_Rt": Rt, "cutoff_M": Mc, "sample_stride": SAMPLE}
    for k, t in tables.items():
        out[k + "_sha256"] = hashlib.sha256(np.ascontiguousarray(t).tobytes()).hexdigest()
        out[k + "_shape"] = np.array(t.shape)
        out[k + "_sample"] = t.reshape(-1)[::SAMPLE].copy()
        print(k, t.shape, out[k + "_sha256"][:16], "invalid(-1):", int((t == -1).sum()))
    np.savez_compressed(os.path.join(OUT, "remap_tables.npz"), **out)


if __name__ == "__main__":
    main()
