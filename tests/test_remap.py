"""Remap tables: TS.cpp:284-330 and rectify.cpp:86-199 as one batched
CUDA kernel behind tscm_remap_tables().  CV_32FC1 tables are byte data: the bar is BIT-EXACT.

CPU: oracle/remap_oracle.c against the committed golden tables (independent numpy
transcription, tests/golden/make_golden_remap.py).  GPU: the kernel through the C-ABI against
the oracle and the golden hashes."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

from conftest import ROOT
from tscm_calib_b200 import capi, remap, synth

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "remap_tables.npz"))
CAMS, TWCS = synth.CALIB_INTRINSICS, synth.CALIB_TWC


def sha(t):
    return hashlib.sha256(np.ascontiguousarray(t).tobytes()).hexdigest()


def check_golden(name, t):
    assert tuple(GOLD[name + "_shape"]) == t.shape
    stride = int(GOLD["sample_stride"])
    np.testing.assert_array_equal(t.reshape(-1)[::stride].view(np.uint32), GOLD[name + "_sample"].view(np.uint32))
    assert sha(t) == str(GOLD[name + "_sha256"]), name


def cutoff_job():
    return capi.remap_job(CAMS[2], GOLD["cutoff_M"], (100.0, 100.0, 127.5, 95.5), (256, 192),
                          origin=(50, 32), offset=(1280.0, 0.0), cutoff_w2=remap.W2_CUTOFF)


def golden_cases():
    """(golden names, jobs, map size, initial fill) of every committed table pair."""
    jobs, size = remap.undistort_jobs(CAMS[0], 300.0, 300.0, 639.5, 539.5, (1280, 1080))
    yield ("undistort_x", "undistort_y"), jobs, size, 0.0
    jobs, size = remap.chessboard_jobs(CAMS[1], GOLD["board_Rt"], (11, 8), 45.0)
    yield ("board_x", "board_y"), jobs, size, 0.0
    left, right, size = remap.epipolar_jobs(CAMS, TWCS)
    yield ("left_mapx", "left_mapy"), left, size, 0.0
    yield ("right_mapx", "right_mapy"), right, size, 0.0
    yield ("cutoff_x", "cutoff_y"), [cutoff_job()], (320, 300), 7.0


def oracle_tables(oracle, jobs, size, fill):
    lib = oracle.load()
    x = np.full((size[1], size[0]), fill, np.float32)
    y = x.copy()
    arr = (capi.TscmRemapJob * len(jobs))(*jobs)
    fp = C.POINTER(C.c_float)
    assert lib.tscm_oracle_remap_tables(arr, len(jobs), size[0], size[1], x.ctypes.data_as(fp), y.ctypes.data_as(fp)) == 0
    return x, y


def test_oracle_matches_golden_tables(oracle):
    for names, jobs, size, fill in golden_cases():
        x, y = oracle_tables(oracle, jobs, size, fill)
        check_golden(names[0], x)
        check_golden(names[1], y)


def test_epipolar_jobs_follow_init_remap_layout():
    """rectify.cpp:93-198: block k of the left table is camera k (front, right, rear, left), of the
    right table camera k+1; mosaic offsets +1280 / +1080; all with the validity cut-off."""
    left, right, size = remap.epipolar_jobs(CAMS, TWCS)
    assert size == (400, 1600) and len(left) == len(right) == 4
    offs = [(0, 0), (1280, 0), (0, 1080), (1280, 1080)]
    for k in range(4):
        for job, cam in ((left[k], k), (right[k], (k + 1) % 4)):
            assert (job.row0, job.col0, job.width, job.height) == (400 * k, 0, 400, 400)
            assert (job.offset_x, job.offset_y) == offs[cam]
            assert list(job.intrinsics) == list(CAMS[cam]) and job.cutoff_w2 == 0.42399
            M = np.array(job.matrix).reshape(3, 3)
            np.testing.assert_allclose(M @ M.T, np.eye(3), atol=1e-12)      # a rotation


def test_remap_without_gpu_fails_loudly():
    """No CPU fallback: without a device the call reports TSCM_ERR_NO_DEVICE."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    jobs, size = remap.undistort_jobs(CAMS[0], 300.0, 300.0, 639.5, 539.5, (64, 48))
    with pytest.raises(capi.TscmError, match="tscm error 3"):
        capi.remap_tables(jobs, size)


def test_remap_rejects_bad_jobs():
    bad = capi.remap_job(CAMS[0], remap.IDENTITY, (1.0, 1.0, 0.0, 0.0), (64, 64), origin=(10, 0))
    with pytest.raises(capi.TscmError, match="does not fit"):
        capi.remap_tables([bad], (64, 64))
    with pytest.raises(capi.TscmError, match="limit"):
        capi.remap_tables([cutoff_job()] * 65, (320, 300))


@pytest.mark.gpu
def test_gpu_tables_are_bit_identical_to_oracle_and_golden(oracle):
    for names, jobs, size, fill in golden_cases():
        x0, y0 = oracle_tables(oracle, jobs, size, fill)
        x = np.full((size[1], size[0]), fill, np.float32)
        y = x.copy()
        capi.remap_tables(jobs, size, device=0, mapx=x, mapy=y)
        np.testing.assert_array_equal(x.view(np.uint32), x0.view(np.uint32))
        np.testing.assert_array_equal(y.view(np.uint32), y0.view(np.uint32))
        check_golden(names[0], x)
        check_golden(names[1], y)


@pytest.mark.gpu
def test_gpu_large_table_with_skew_matches_oracle(oracle):
    """4096 x 3072 table (12.6 M pixels, 100 MB of output), non-zero skew b, c and an odd width
    tail: bit-exact against the scalar restatement, and the kernel time is reported."""
    intr = CAMS[3].copy()
    intr[7], intr[8] = 0.7, -0.4
    M = synth.rodrigues(np.array([0.05, -0.4, 0.02]))
    jobs = [capi.remap_job(intr, M, (900.0, 910.0, 2047.5, 1535.5), (4093, 3072), cutoff_w2=remap.W2_CUTOFF)]
    x0, y0 = oracle_tables(oracle, jobs, (4096, 3072), 0.0)
    x, y, ms = capi.remap_tables(jobs, (4096, 3072), device=0)
    np.testing.assert_array_equal(x.view(np.uint32), x0.view(np.uint32))
    np.testing.assert_array_equal(y.view(np.uint32), y0.view(np.uint32))
    assert (x[:, 4093:] == 0).all()
    print(f"k_remap_tables: {4093 * 3072 / 1e6:.1f} Mpixel in {ms * 1e3:.1f} us "
          f"({4093 * 3072 * 8 / ms / 1e6:.1f} GB/s of table writes)")


def test_oracle_matches_numpy_transcription_on_random_jobs(oracle):
    """Beyond the committed tables: random cameras (with skew), rotations, pixel grids, offsets
    and cut-offs — the scalar C restatement and the vectorised numpy transcription of the
    reference loops (tests/golden/make_golden_remap.py) agree bit for bit."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_remap",
                                                  os.path.join(ROOT, "tests", "golden", "make_golden_remap.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(77)
    for n in range(12):
        intr = CAMS[n % 4].copy()
        intr[:7] *= 1 + 0.05 * rng.standard_normal(7)
        if n % 2:
            intr[7], intr[8] = rng.normal(0, 0.8, 2)            # skew b, c
        M = synth.rodrigues(rng.normal(0, [0.3, 1.2, 0.3][n % 3], 3))
        w, h = int(rng.integers(17, 90)), int(rng.integers(9, 70))
        ray = (float(rng.uniform(40, 400)), float(rng.uniform(40, 400)), float(rng.uniform(0, w)), float(rng.uniform(0, h)))
        off = (float(rng.choice([0.0, 1280.0])), float(rng.choice([0.0, 1080.0])))
        w2 = float(rng.choice([0.0, remap.W2_CUTOFF]))
        job = capi.remap_job(intr, M, ray, (w, h), offset=off, cutoff_w2=w2)
        x0, y0 = oracle_tables(oracle, [job], (w, h), 0.0)
        gx, gy = ref.grid(w, h, *ray)
        u, v = ref.project(intr, *ref.apply(M, gx, gy), w2=w2)
        x1 = (u + off[0] if off[0] else u).astype(np.float32)
        y1 = (v + off[1] if off[1] else v).astype(np.float32)
        np.testing.assert_array_equal(x0.view(np.uint32), x1.view(np.uint32))
        np.testing.assert_array_equal(y0.view(np.uint32), y1.view(np.uint32))
