import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    """A golden fixture as (ProblemArrays, dict of arrays)."""
    from tscm_calib_b200.capi import ProblemArrays
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    p = ProblemArrays(z["board_xy"], z["view_camera"], z["view_frame"], z["obs_xy"],
                      int(z["num_cameras"]), int(z["num_frames"]), int(z["fixed_camera"]))
    return p, z


def golden_options(z, **kw):
    from tscm_calib_b200 import capi
    o = z["options"]
    return capi.default_options(max_num_iterations=int(o[0]), loss_type=int(o[1]),
                                loss_scale=float(o[2]), **kw)


@pytest.fixture(scope="session")
def hostmath():
    """Host build of csrc/tscm_math.cuh (the kernels' per-observation arithmetic)."""
    src = os.path.join(ROOT, "tests", "hostmath", "hostmath.cpp")
    out = os.path.join(ROOT, "tests", "hostmath", "libhostmath.so")
    deps = [src, os.path.join(ROOT, "tscm_calib_b200", "csrc", "tscm_math.cuh"),
            os.path.join(ROOT, "tscm_calib_b200", "csrc", "tscm_pair_lists.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out, src], check=True)
    return ctypes.CDLL(out)


@pytest.fixture(scope="session")
def hostinit():
    """C wrappers (tests/hostmath/hostinit.cpp) around the C++ TripleSphereCamera adapter."""
    return build_hostinit()


def build_hostinit():
    from tscm_calib_b200 import build as tbuild
    tbuild.build_cuda()
    here = os.path.join(ROOT, "tests", "hostmath")
    host = os.path.join(ROOT, "tscm_calib_b200", "host")
    out = os.path.join(here, "libhostinit.so")
    srcs = [os.path.join(here, "hostinit.cpp"), os.path.join(host, "ts_camera.cpp"),
            os.path.join(host, "multi_calib_b200.cpp")]
    deps = srcs + [os.path.join(host, "ts_camera.h"), os.path.join(host, "cv_compat.h"),
                   os.path.join(host, "multi_calib_b200.h"),
                   os.path.join(ROOT, "oracle", "cv_calib3d_port.h"),
                   os.path.join(ROOT, "tscm_calib_b200", "libtscm_b200.so"),
                   os.path.join(ROOT, "include", "tscm.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        env = dict(os.environ)
        env.pop("CXX", None)
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out] + srcs +
                       ["-L" + os.path.join(ROOT, "tscm_calib_b200"), "-ltscm_b200",
                        "-Wl,-rpath," + os.path.join(ROOT, "tscm_calib_b200")], check=True, env=env)
    return ctypes.CDLL(out)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    o.load()
    return o


def rel_err(a, b, floor=1e-12):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / (np.abs(b) + floor))) if a.size else 0.0
