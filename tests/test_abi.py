"""C-ABI surface: the library loads, exports every symbol include/tscm.h declares,
and the structs agree between C and the ctypes mirror.  No compute calls."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from tscm_calib_b200 import capi
from tscm_calib_b200 import build as tbuild


@pytest.fixture(scope="module")
def lib():
    tbuild.build_cuda()
    return capi.load_library()


def test_library_is_in_tree():
    tbuild.build_cuda()
    assert os.path.dirname(capi.LIB_PATH) == os.path.join(ROOT, "tscm_calib_b200")
    assert os.path.exists(capi.LIB_PATH)


def test_header_symbols_all_exported(lib):
    header = open(os.path.join(ROOT, "include", "tscm.h")).read()
    declared = set(re.findall(r"\b(tscm_[a-z_0-9]+)\s*\(", header))
    declared -= {"tscm_problem", "tscm_options", "tscm_summary", "tscm_solver"}
    assert declared == set(capi.EXPORTED_SYMBOLS), declared ^ set(capi.EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_options_init_matches_ceres_defaults(lib):
    o = capi.TscmOptions()
    lib.tscm_options_init(C.byref(o))
    d = capi.default_options()
    for name, _ in capi.TscmOptions._fields_:
        assert getattr(o, name) == getattr(d, name), name
    # ceres::Solver::Options defaults the reference leaves in force (multi_calib.cpp:209-212)
    assert o.max_num_iterations == 50 and o.function_tolerance == 1e-6
    assert o.gradient_tolerance == 1e-10 and o.parameter_tolerance == 1e-8
    assert o.initial_trust_region_radius == 1e4 and o.min_relative_decrease == 1e-3
    assert o.min_lm_diagonal == 1e-6 and o.max_lm_diagonal == 1e32
    assert o.jacobi_scaling == 1 and o.loss_type == 0


def test_struct_sizes():
    # 4 int32 + 4 pointers + int32 (+pad)
    assert C.sizeof(capi.TscmProblem) == 56
    # the init entry points (sizes of the C structs as gcc lays them out: include/tscm.h)
    assert C.sizeof(capi.TscmPoseGraphProblem) == 56 and C.sizeof(capi.TscmPoseGraphResult) == 72
    assert C.sizeof(capi.TscmMonoInitProblem) == 56 and C.sizeof(capi.TscmMonoInitResult) == 104
    assert C.sizeof(capi.TscmSummary) == 16 + 3 * 8 + 8 + 5 * 8


def test_version_and_error_strings(lib):
    assert b"sm_100a" in lib.tscm_version()
    assert isinstance(lib.tscm_last_error(), bytes)


def test_invalid_problem_is_rejected_before_touching_the_device(lib):
    import numpy as np
    board = np.zeros((4, 2))
    # views not camera-major
    p = capi.ProblemArrays(board, [1, 0], [0, 0], np.zeros((2, 4, 2)), 2, 1)
    h = C.c_void_p()
    rc = lib.tscm_solver_create(C.byref(p.c), C.byref(capi.default_options()), 0, C.byref(h))
    assert rc == 1 and b"camera-major" in lib.tscm_last_error()
    # a frame nobody sees
    p = capi.ProblemArrays(board, [0], [0], np.zeros((1, 4, 2)), 1, 2)
    rc = lib.tscm_solver_create(C.byref(p.c), C.byref(capi.default_options()), 0, C.byref(h))
    assert rc == 1 and b"seen by no camera" in lib.tscm_last_error()
    # empty problem
    p = capi.ProblemArrays(board, [], [], np.zeros((0, 4, 2)), 1, 1)
    rc = lib.tscm_solver_create(C.byref(p.c), C.byref(capi.default_options()), 0, C.byref(h))
    assert rc == 1


def test_no_cpu_fallback_without_a_device(lib):
    """On a box without a GPU the solve must fail loudly, not fall back."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from tscm_calib_b200 import synth
    sp = synth.generate(num_cameras=1, num_frames=3, board=(4, 3), seed=1)
    with pytest.raises(capi.TscmError) as e:
        capi.solve(sp.problem, sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tscm_calib_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                for pat in (r"^\s*(from|import)\s+oracle", r"libtscm_oracle", r"tscm_oracle_",
                            r"#include\s+\".*oracle"):
                    assert not re.search(pat, text, flags=re.M), (pat, os.path.join(dirpath, f))
