"""The C++ drop-in adapters (tscm_calib_b200/host): MultiCalib / TripleSphereCamera shaped
classes above the C-ABI and the YAML writer of main.cpp:293-319."""
import os
import struct
import subprocess
import tempfile

import numpy as np
import pytest

from tscm_calib_b200 import build as tbuild
from tscm_calib_b200 import capi, synth


def dump_problem(path, sp):
    p = sp.problem
    C, F, K = p.num_cameras, p.num_frames, p.corners_per_board
    with open(path, "wb") as f:
        f.write(struct.pack("<iii", C, F, K))
        f.write(p.board_xy.astype("<f8").tobytes())
        f.write(sp.visible.astype(np.uint8).tobytes())
        f.write(p.obs_xy.astype("<f8").tobytes())            # camera-major views, as the file wants
        f.write(sp.init_intrinsics.astype("<f8").tobytes())
        f.write(sp.init_cam_rt.astype("<f8").tobytes())
        f.write(sp.init_board_rt.astype("<f8").tobytes())


def read_result(path, C, F):
    raw = open(path, "rb").read()
    head = struct.unpack("<4i", raw[:16])
    vals = np.frombuffer(raw[16:], dtype="<f8")
    o = 3
    intr = vals[o:o + 9 * C].reshape(C, 9); o += 9 * C
    cam_rt = vals[o:o + 6 * C].reshape(C, 6); o += 6 * C
    board_rt = vals[o:o + 6 * F].reshape(F, 6); o += 6 * F
    per_cam = vals[o:o + C]
    return head, vals[:3], intr, cam_rt, board_rt, per_cam


def run_demo(sp, yaml_only=False):
    demo = tbuild.build_host()
    d = tempfile.mkdtemp()
    prob, res, yml = (os.path.join(d, n) for n in ("problem.bin", "result.bin", "calib.yaml"))
    dump_problem(prob, sp)
    cmd = [demo, prob, res, yml] + (["--yaml-only"] if yaml_only else [])
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    return res, yml, out.stdout


def test_yaml_output_is_opencv_filestorage_compatible():
    """`EpipolarRectify` reads the file with cv::FileStorage (rectify.cpp:262-270)."""
    cv2 = pytest.importorskip("cv2")
    sp = synth.config(2, num_frames=12)
    res, yml, _ = run_demo(sp, yaml_only=True)
    text = open(yml).read()
    assert text.startswith("%YAML:1.0\n---\ncam0: !!opencv-matrix\n   rows: 1\n   cols: 9\n   dt: d\n")
    fs = cv2.FileStorage(yml, cv2.FILE_STORAGE_READ)
    for m in range(4):
        cam = fs.getNode(f"cam{m}").mat()
        T = fs.getNode(f"Twc{m}").mat()
        assert cam.shape == (1, 9) and T.shape == (3, 4)
        np.testing.assert_allclose(cam[0], sp.init_intrinsics[m], rtol=1e-15)
        R = synth.rodrigues(sp.init_cam_rt[m, :3])
        np.testing.assert_allclose(T[:, :3], R, atol=1e-12)
        np.testing.assert_allclose(T[:, 3], sp.init_cam_rt[m, 3:], rtol=1e-15)
    np.testing.assert_array_equal(fs.getNode("Twc0").mat(), np.eye(3, 4))   # calib.yaml:12-16
    fs.release()


@pytest.mark.gpu
def test_multicalib_adapter_matches_oracle(oracle):
    sp = synth.config(2, num_frames=60)
    res, yml, stdout = run_demo(sp)
    C, F = sp.problem.num_cameras, sp.problem.num_frames
    head, costs, intr, cam_rt, board_rt, per_cam = read_result(res, C, F)
    a0, b0, c0, s0 = oracle.solve(sp.problem, sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt,
                                  capi.default_options())
    assert capi.TERMINATION[head[0]] == s0.termination and head[1] == s0.num_iterations
    assert abs(costs[0] - s0.initial_cost) <= 1e-9 * s0.initial_cost
    assert abs(costs[1] - s0.final_cost) <= 1e-9 * s0.final_cost
    np.testing.assert_allclose(intr, a0, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(cam_rt, b0, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(board_rt, c0, rtol=1e-7, atol=1e-9)
    per0, overall0, _ = oracle.reprojection_error(sp.problem, a0, b0, c0)
    np.testing.assert_allclose(per_cam, per0, rtol=0, atol=1e-8)
    assert abs(costs[2] - overall0) < 1e-8
    # the reference's console contract: BriefReport line + per-camera read-out
    assert "Ceres Solver Report: Iterations:" in stdout and "Termination: CONVERGENCE" in stdout
    assert "average reproject error" in stdout
