"""ctypes binding of the CPU oracle (oracle/libtscm_oracle.so).

TEST INFRASTRUCTURE ONLY — see oracle/tscm_oracle.h.  Importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
never from the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from tscm_calib_b200.capi import (ProblemArrays, SummaryBuffers, TscmOptions, TscmProblem,
                                  TscmRemapJob, TscmSummary, _dp, c_double_p, default_options)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtscm_oracle.so")
_lib = None


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("tscm_oracle.cpp", "tscm_oracle.h", "remap_oracle.c", "pose_graph_oracle.c", "mono_init_oracle.cpp", "cv_calib3d_port.h",
                                            os.path.join("..", "include", "tscm.h"))]
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B", "libtscm_oracle.so"], check=True,
                   stdout=subprocess.DEVNULL)
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER
    lib.tscm_oracle_solve.argtypes = [P(TscmProblem), P(TscmOptions), c_double_p, c_double_p,
                                      c_double_p, P(TscmSummary), C.c_int]
    lib.tscm_oracle_solve.restype = C.c_int
    lib.tscm_oracle_eval_jacobian.argtypes = [P(TscmProblem), c_double_p, c_double_p, c_double_p,
                                              c_double_p, c_double_p, c_double_p]
    lib.tscm_oracle_eval_jacobian.restype = C.c_int
    lib.tscm_oracle_eval_jacobian_mono.argtypes = [P(TscmProblem), c_double_p, c_double_p,
                                                   c_double_p, c_double_p, c_double_p]
    lib.tscm_oracle_eval_jacobian_mono.restype = C.c_int
    lib.tscm_oracle_reduced_size.argtypes = [P(TscmProblem)]
    lib.tscm_oracle_reduced_size.restype = C.c_int
    lib.tscm_oracle_reduced_system.argtypes = [P(TscmProblem), P(TscmOptions), c_double_p,
                                               c_double_p, c_double_p, C.c_double, c_double_p,
                                               c_double_p]
    lib.tscm_oracle_reduced_system.restype = C.c_int
    lib.tscm_oracle_reprojection_error.argtypes = [P(TscmProblem), c_double_p, c_double_p,
                                                   c_double_p, c_double_p, c_double_p, c_double_p]
    lib.tscm_oracle_reprojection_error.restype = C.c_int
    lib.tscm_oracle_project.argtypes = [c_double_p, c_double_p, C.c_int, c_double_p]
    lib.tscm_oracle_project.restype = None
    lib.tscm_oracle_remap_tables.argtypes = [P(TscmRemapJob), C.c_int32, C.c_int32, C.c_int32,
                                             P(C.c_float), P(C.c_float)]
    lib.tscm_oracle_remap_tables.restype = C.c_int
    u8p, i32p = P(C.c_uint8), P(C.c_int32)
    lib.tscm_oracle_pose_graph.argtypes = [C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, u8p,
                                           c_double_p, c_double_p, c_double_p, c_double_p, u8p,
                                           i32p, i32p, c_double_p, c_double_p]
    lib.tscm_oracle_pose_graph.restype = C.c_int
    lib.tscm_oracle_pose_pair_error.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, u8p,
                                                c_double_p, c_double_p, c_double_p]
    lib.tscm_oracle_pose_pair_error.restype = C.c_double
    lib.tscm_oracle_mono_init.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, u8p, c_double_p,
                                          C.c_int, c_double_p, c_double_p, u8p, i32p]
    lib.tscm_oracle_mono_init.restype = C.c_int
    lib.tscm_oracle_solve_z.argtypes = [c_double_p, C.c_int, C.c_int, c_double_p]
    lib.tscm_oracle_solve_z.restype = None
    lib.tscm_oracle_solve_pnp.argtypes = [c_double_p, c_double_p, C.c_int, c_double_p, c_double_p]
    lib.tscm_oracle_solve_pnp.restype = C.c_int
    _lib = lib
    return lib


def _params(problem, intrinsics, cam_rt, board_rt):
    a = np.array(intrinsics, dtype=np.float64, order="C").reshape(problem.num_cameras, 9)
    b = np.array(cam_rt, dtype=np.float64, order="C").reshape(problem.num_cameras, 6)
    c = np.array(board_rt, dtype=np.float64, order="C").reshape(problem.num_frames, 6)
    return a, b, c


def solve(problem: ProblemArrays, intrinsics, cam_rt, board_rt, options=None, num_threads=1):
    lib = load()
    options = options or default_options()
    a, b, c = _params(problem, intrinsics, cam_rt, board_rt)
    buf = SummaryBuffers(options.max_num_iterations + 2)
    rc = lib.tscm_oracle_solve(C.byref(problem.c), C.byref(options), _dp(a), _dp(b), _dp(c),
                               C.byref(buf.c), num_threads)
    if rc != 0:
        raise RuntimeError(f"oracle solve failed rc={rc}")
    return a, b, c, buf.result()


def eval_jacobian(problem: ProblemArrays, intrinsics, cam_rt, board_rt):
    lib = load()
    a, b, c = _params(problem, intrinsics, cam_rt, board_rt)
    N = problem.num_observations
    r, J, cost = np.zeros((N, 2)), np.zeros((N, 2, 21)), C.c_double()
    rc = lib.tscm_oracle_eval_jacobian(C.byref(problem.c), _dp(a), _dp(b), _dp(c), _dp(r), _dp(J),
                                       C.byref(cost))
    assert rc == 0
    return r, J, cost.value


def eval_jacobian_mono(problem: ProblemArrays, intrinsics, board_rt):
    lib = load()
    a = np.array(intrinsics, dtype=np.float64, order="C").reshape(9)
    c = np.array(board_rt, dtype=np.float64, order="C").reshape(problem.num_frames, 6)
    N = problem.num_observations
    r, J, cost = np.zeros((N, 2)), np.zeros((N, 2, 15)), C.c_double()
    rc = lib.tscm_oracle_eval_jacobian_mono(C.byref(problem.c), _dp(a), _dp(c), _dp(r), _dp(J),
                                            C.byref(cost))
    assert rc == 0
    return r, J, cost.value


def reduced_system(problem: ProblemArrays, intrinsics, cam_rt, board_rt, radius, options=None):
    lib = load()
    options = options or default_options()
    a, b, c = _params(problem, intrinsics, cam_rt, board_rt)
    n = lib.tscm_oracle_reduced_size(C.byref(problem.c))
    lhs, rhs = np.zeros((n, n)), np.zeros(n)
    rc = lib.tscm_oracle_reduced_system(C.byref(problem.c), C.byref(options), _dp(a), _dp(b),
                                        _dp(c), radius, _dp(lhs), _dp(rhs))
    assert rc == 0
    return lhs, rhs


def live_reduced_index(problem: ProblemArrays):
    """Indices of the oracle's reduced system (9 intrinsics per camera, Ceres
    layout) that the CUDA library keeps (it drops the structurally-zero b, c
    columns: 7 intrinsics per camera)."""
    idx, off = [], 0
    for m in range(problem.num_cameras):
        rts = 0 if m == problem.fixed_camera else 6
        idx += list(range(off, off + rts + 7))
        off += rts + 9
    return np.array(idx)


def reprojection_error(problem: ProblemArrays, intrinsics, cam_rt, board_rt):
    lib = load()
    a, b, c = _params(problem, intrinsics, cam_rt, board_rt)
    per = np.zeros(problem.num_cameras)
    overall, rms = C.c_double(), C.c_double()
    rc = lib.tscm_oracle_reprojection_error(C.byref(problem.c), _dp(a), _dp(b), _dp(c), _dp(per),
                                            C.byref(overall), C.byref(rms))
    assert rc == 0
    return per, overall.value, rms.value


def project(intrinsic9, pts):
    lib = load()
    a = np.array(intrinsic9, dtype=np.float64, order="C").reshape(9)
    p = np.array(pts, dtype=np.float64, order="C").reshape(-1, 3)
    uv = np.zeros((p.shape[0], 2))
    lib.tscm_oracle_project(_dp(a), _dp(p), p.shape[0], _dp(uv))
    return uv


def remap_tables(jobs, map_size):
    """CPU restatement of the reference's remap-table loops (remap_oracle.c)."""
    lib = load()
    W, H = int(map_size[0]), int(map_size[1])
    mapx, mapy = np.zeros((H, W), dtype=np.float32), np.zeros((H, W), dtype=np.float32)
    arr = (TscmRemapJob * len(jobs))(*jobs)
    rc = lib.tscm_oracle_remap_tables(arr, len(jobs), W, H, mapx.ctypes.data_as(C.POINTER(C.c_float)),
                                      mapy.ctypes.data_as(C.POINTER(C.c_float)))
    assert rc == 0
    return mapx, mapy


class PoseGraph:
    """Result of a pose-graph initialisation: poses as 12 doubles = R row-major | t."""

    def __init__(self, C_, B):
        self.rc = 0
        self.camera_pose = np.zeros((C_, 12))
        self.board_pose = np.zeros((B, 12))
        self.board_init = np.zeros(B, dtype=np.uint8)
        self.camera_choice = np.full(C_, -1, dtype=np.int32)
        self.board_choice = np.full(B, -1, dtype=np.int32)
        self.camera_candidate_error = np.full((C_, B), np.nan)
        self.board_candidate_error = np.full((B, C_), np.nan)


def pose_graph(worlds, intrinsics, has, mono_rt, pixels):
    """CPU restatement of MultiCalib's constructor (multi_calib.cpp:6-153), pose_graph_oracle.c.
    has [C][B], mono_rt [C][B][3][3], pixels [C][B][K][2], worlds [K][3]."""
    lib = load()
    has = np.ascontiguousarray(has, dtype=np.uint8)
    C_, B = has.shape
    worlds = np.ascontiguousarray(worlds, dtype=np.float64).reshape(-1, 3)
    K = worlds.shape[0]
    intr = np.ascontiguousarray(intrinsics, dtype=np.float64).reshape(C_, 9)
    rt = np.ascontiguousarray(mono_rt, dtype=np.float64).reshape(C_, B, 9)
    px = np.ascontiguousarray(pixels, dtype=np.float64).reshape(C_, B, K, 2)
    r = PoseGraph(C_, B)
    u8p, i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_int32)
    r.rc = lib.tscm_oracle_pose_graph(C_, B, K, _dp(worlds), _dp(intr), has.ctypes.data_as(u8p), _dp(rt),
                                      _dp(px), _dp(r.camera_pose), _dp(r.board_pose),
                                      r.board_init.ctypes.data_as(u8p), r.camera_choice.ctypes.data_as(i32p),
                                      r.board_choice.ctypes.data_as(i32p), _dp(r.camera_candidate_error),
                                      _dp(r.board_candidate_error))
    return r


def pose_pair_error(i, j, worlds, intrinsics, has, mono_rt, pixels, prev_camera_pose):
    """Score of ONE chain candidate (camera i through board j) given the pose chosen for camera
    i-1 (multi_calib.cpp:36-47, 53-81): spot checks of the scoring kernel at full size."""
    lib = load()
    has = np.ascontiguousarray(has, dtype=np.uint8)
    C_, B = has.shape
    worlds = np.ascontiguousarray(worlds, dtype=np.float64).reshape(-1, 3)
    K = worlds.shape[0]
    intr = np.ascontiguousarray(intrinsics, dtype=np.float64).reshape(C_, 9)
    rt = np.ascontiguousarray(mono_rt, dtype=np.float64).reshape(C_, B, 9)
    px = np.ascontiguousarray(pixels, dtype=np.float64).reshape(C_, B, K, 2)
    pose = np.ascontiguousarray(prev_camera_pose, dtype=np.float64).reshape(12)
    return lib.tscm_oracle_pose_pair_error(int(i), int(j), B, K, _dp(worlds), _dp(intr),
                                           has.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(rt), _dp(px), _dp(pose))


class MonoInit:
    """Result of the mono cold start: intrinsics [9], Rt [F][3][3], frame_ok [F]."""

    def __init__(self, F):
        self.rc = 0
        self.intrinsics = np.zeros(9)
        self.Rt = np.zeros((F, 3, 3))
        self.frame_ok = np.zeros(F, dtype=np.uint8)
        self.rows_used = 0
        self.kernel_ms = 0.0


def mono_init(board, image, worlds, has, pixels, guess=None):
    """CPU restatement of TripleSphereCamera::calibrate up to the refinement (TS.cpp:36-52,
    110-203), mono_init_oracle.cpp.  board = (W, H), image = (w, h), pixels [F][K][2];
    guess: 7 or 9 intrinsics of the has_init_guess_ path, None for the cold start."""
    lib = load()
    has = np.ascontiguousarray(has, dtype=np.uint8)
    F = has.shape[0]
    W, H = int(board[0]), int(board[1])
    worlds = np.ascontiguousarray(worlds, dtype=np.float64).reshape(W * H, 3)
    px = np.ascontiguousarray(pixels, dtype=np.float64).reshape(F, W * H, 2)
    r = MonoInit(F)
    if guess is not None:
        r.intrinsics[:len(guess)] = guess
    rows = C.c_int32(0)
    u8p = C.POINTER(C.c_uint8)
    r.rc = lib.tscm_oracle_mono_init(F, W, H, int(image[0]), int(image[1]), _dp(worlds), has.ctypes.data_as(u8p),
                                     _dp(px), 0 if guess is None else 1, _dp(r.intrinsics), _dp(r.Rt),
                                     r.frame_ok.ctypes.data_as(u8p), C.byref(rows))
    r.rows_used = rows.value
    return r


def solve_z(A):
    lib = load()
    A = np.ascontiguousarray(A, dtype=np.float64)
    z = np.zeros(A.shape[1])
    lib.tscm_oracle_solve_z(_dp(A), A.shape[0], A.shape[1], _dp(z))
    return z


def solve_pnp(obj, img):
    lib = load()
    obj, img = np.ascontiguousarray(obj, dtype=np.float64), np.ascontiguousarray(img, dtype=np.float64)
    r, t = np.zeros(3), np.zeros(3)
    rc = lib.tscm_oracle_solve_pnp(_dp(obj), _dp(img), obj.shape[0], _dp(r), _dp(t))
    return rc, r, t
