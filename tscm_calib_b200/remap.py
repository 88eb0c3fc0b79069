"""Host-side job builders for tscm_remap_tables() (include/tscm.h).

Each function mirrors one call site of the reference and only PACKS jobs — the per-pixel
work is the CUDA kernel k_remap_tables (csrc/tscm_remap.cuh); there is no CPU fallback.

  undistort_jobs        TripleSphereCamera::undistort            /root/reference/TS.cpp:284-306
  chessboard_jobs       TripleSphereCamera::undistort_chessboard /root/reference/TS.cpp:308-330
  epipolar_jobs         Remap::Remap + Remap::init_remap         /root/reference/EpipolarRectify/rectify.cpp:52-199
"""
from __future__ import annotations

import math

import numpy as np

from . import capi

IDENTITY = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0)
W2_CUTOFF = 0.42399          # rectify.cpp:7


def undistort_jobs(intrinsics, fx, fy, cx, cy, img_size):
    """TS.cpp:284-306: pinhole (fx, fy, cx, cy) output image of img_size = (width, height)."""
    return [capi.remap_job(intrinsics, IDENTITY, (fx, fy, cx, cy), img_size)], tuple(img_size)


def chessboard_jobs(intrinsics, Rt, chessboard, chessboard_size):
    """TS.cpp:308-330: fronto-parallel board image; Rt is the 3x3 [r1 r2 t] of the frame,
    chessboard = (width, height) inner-corner counts, chessboard_size the square in mm.
    img_size = ((w+1)*size, (h+1)*size) truncated to int like cv::Size (TS.cpp:313)."""
    size = (int((chessboard[0] + 1) * chessboard_size), int((chessboard[1] + 1) * chessboard_size))
    ray = (1.0, 1.0, float(chessboard_size), float(chessboard_size))   # P = (j - size, i - size, 1)
    return [capi.remap_job(intrinsics, np.asarray(Rt, dtype=np.float64).reshape(9), ray, size)], size


def _normalize(v):
    """rectify.cpp:215-222."""
    n = math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    return list(v) if n == 0 else [v[0] / n, v[1] / n, v[2] / n]


def calc_R(t1, t2):
    """rectify.cpp:235-250: rectifying rotation of a camera pair from the two centres."""
    x = _normalize([t2[0] - t1[0], t2[1] - t1[1], t2[2] - t1[2]])
    z = _normalize([-x[2], 0.0, x[0]])
    y = _normalize([-z[2] * x[1] + z[1] * x[2], z[2] * x[0] - z[0] * x[2], -z[1] * x[0] + z[0] * x[1]])
    return [[x[0], y[0], z[0]], [x[1], y[1], z[1]], [x[2], y[2], z[2]]]


def _matmul_t(A, B):
    """A.t() * B for 3x3 lists, rows summed left to right."""
    return [[A[0][i] * B[0][j] + A[1][i] * B[1][j] + A[2][i] * B[2][j] for j in range(3)] for i in range(3)]


def epipolar_jobs(cams, Twcs, image_size=(400, 400), focal=200.0, centre=200.0,
                  mosaic=(1280.0, 1080.0)):
    """rectify.cpp:52-199.  cams = [front, right, rear, left] 9-vectors (cam0..3 of the
    calibration YAML), Twcs the matching 3x4 [R|t].  Returns (left_jobs, right_jobs,
    map_size): four 400x400 blocks stacked vertically in each 400x1600 table; the source
    image is the 2x2 mosaic front | right / rear | left (offsets +1280 / +1080)."""
    T = [np.asarray(t, dtype=np.float64).reshape(3, 4) for t in Twcs]
    R = [[[float(t[r, c]) for c in range(3)] for r in range(3)] for t in T]
    tv = [[float(t[r, 3]) for r in range(3)] for t in T]
    FRONT, RIGHT, REAR, LEFT = 0, 1, 2, 3
    offs = {FRONT: (0.0, 0.0), RIGHT: (mosaic[0], 0.0), REAR: (0.0, mosaic[1]), LEFT: (mosaic[0], mosaic[1])}
    pairs = [(FRONT, RIGHT), (RIGHT, REAR), (REAR, LEFT), (LEFT, FRONT)]   # rectify.cpp:88-91
    W, H = image_size
    ray = (focal, focal, centre, centre)
    left, right = [], []
    for blk, (a, b) in enumerate(pairs):
        Rab = calc_R(tv[a], tv[b])
        for cam, dst in ((a, left), (b, right)):
            M = _matmul_t(R[cam], Rab)
            dst.append(capi.remap_job(cams[cam], [v for row in M for v in row], ray, (W, H),
                                      origin=(blk * H, 0), offset=offs[cam], cutoff_w2=W2_CUTOFF))
    return left, right, (W, 4 * H)
