"""Synthetic calibration problems generated through the TS model.

Host-side input generation for tests and bench.py — numpy only, seeded.  The
ground-truth parameter magnitudes are those of the reference's only numeric
fixture, /root/reference/EpipolarRectify/calib.yaml:1-69 (cam0..3 intrinsics and
Twc0..3), copied here as data.  Board: W x H inner corners, 45 mm squares
(main.cpp:190-191), p_b(j) = ((j mod W)*45, (j div W)*45, 0) (main.cpp:12-18).
Image 1280 x 1080 (rectify.cpp:115,142).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .capi import ProblemArrays

IMAGE_W, IMAGE_H = 1280, 1080
SQUARE_MM = 45.0

# EpipolarRectify/calib.yaml cam0..cam3: fx fy cx cy xi lambda alpha b c
CALIB_INTRINSICS = np.array([
    [4.3129641731951233e+02, 4.3077528857601646e+02, 6.4653015901902177e+02,
     5.2120451427825685e+02, -2.7125775332873053e-01, -8.7861849854000834e-02,
     5.6023435889162265e-01, 0., 0.],
    [4.3366730337860304e+02, 4.3377366718652252e+02, 6.5043289767844408e+02,
     5.3217610796339648e+02, -2.5567341708788405e-01, -8.0998645840408265e-02,
     5.6043293184809229e-01, 0., 0.],
    [4.4342294254852777e+02, 4.4269548663571004e+02, 6.5012232252239130e+02,
     5.1864631548858017e+02, -2.3275919129762454e-01, -8.7007852953879805e-02,
     5.6302432477866149e-01, 0., 0.],
    [4.3725205336966712e+02, 4.3738251105641092e+02, 6.4148306394889755e+02,
     5.5309342913742341e+02, -2.6287894613485679e-01, -8.5693153628330507e-02,
     5.6177801764159951e-01, 0., 0.],
])
# EpipolarRectify/calib.yaml Twc0..Twc3 (3x4 [R|t], reference -> camera)
CALIB_TWC = np.array([
    [[1., 0., 0., 0.], [0., 1., 0., 0.], [0., 0., 1., 0.]],
    [[5.0160892202284401e-03, -1.2446352011191332e-02, 9.9990995953163087e-01, 3.1111069091426958e+02],
     [-4.9652802236104215e-02, 9.9868604260337857e-01, 1.2680202652341772e-02, -3.2581972269830493e+00],
     [-9.9875394271013351e-01, -4.9711936502369769e-02, 4.3915020377227887e-03, -3.0250006677005149e+02]],
    [[-9.9912757728632307e-01, -4.1543088687854141e-02, 4.2727143873527804e-03, -4.5684542332524316e+00],
     [-4.1723107975334954e-02, 9.9738123070453755e-01, -5.9075061567300073e-02, -3.5570993658832819e+01],
     [-1.8073646121761103e-03, -5.9201794065508177e-02, -9.9824439943962817e-01, -6.1759896466830685e+02]],
    [[-1.0309658021738319e-02, -5.9375126415255344e-02, -9.9818250100602735e-01, -3.0116931471069142e+02],
     [5.1486531371731037e-02, 9.9687992136190828e-01, -5.9829419793145176e-02, -2.9127716209064634e+01],
     [9.9862047247129027e-01, -5.2009775510466122e-02, -7.2204717690991169e-03, -3.0393000777920400e+02]],
])


# ----------------------------------------------------------------------------
# geometry helpers (vectorised)
# ----------------------------------------------------------------------------
def rodrigues(rvec: np.ndarray) -> np.ndarray:
    """angle-axis (...,3) -> rotation matrices (...,3,3)."""
    rvec = np.asarray(rvec, dtype=np.float64)
    theta = np.linalg.norm(rvec, axis=-1)
    small = theta < 1e-300
    k = rvec / np.where(small, 1.0, theta)[..., None]
    c, s = np.cos(theta)[..., None, None], np.sin(theta)[..., None, None]
    K = np.zeros(rvec.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -k[..., 2], k[..., 1]
    K[..., 1, 0], K[..., 1, 2] = k[..., 2], -k[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -k[..., 1], k[..., 0]
    eye = np.broadcast_to(np.eye(3), K.shape)
    kk = k[..., :, None] * k[..., None, :]
    R = c * eye + (1 - c) * kk + s * K
    R[small] = np.eye(3)
    return R


def rotation_to_rvec(R: np.ndarray) -> np.ndarray:
    """rotation matrices (...,3,3) -> angle-axis (...,3) (robust near pi)."""
    R = np.asarray(R, dtype=np.float64)
    out = np.zeros(R.shape[:-2] + (3,))
    flat_R = R.reshape(-1, 3, 3)
    flat_o = out.reshape(-1, 3)
    for n, M in enumerate(flat_R):
        # quaternion route: stable for every angle
        t = np.trace(M)
        if t > 0:
            s = np.sqrt(t + 1.0) * 2
            q = np.array([0.25 * s, (M[2, 1] - M[1, 2]) / s, (M[0, 2] - M[2, 0]) / s,
                          (M[1, 0] - M[0, 1]) / s])
        else:
            i = int(np.argmax(np.diag(M)))
            j, k = (i + 1) % 3, (i + 2) % 3
            s = np.sqrt(1.0 + M[i, i] - M[j, j] - M[k, k]) * 2
            q = np.zeros(4)
            q[0] = (M[k, j] - M[j, k]) / s
            q[1 + i] = 0.25 * s
            q[1 + j] = (M[j, i] + M[i, j]) / s
            q[1 + k] = (M[k, i] + M[i, k]) / s
        if q[0] < 0:
            q = -q
        vn = np.linalg.norm(q[1:])
        if vn < 1e-300:
            flat_o[n] = 0.0
        else:
            flat_o[n] = q[1:] / vn * (2.0 * np.arctan2(vn, q[0]))
    return out


def ts_project(intr: np.ndarray, P: np.ndarray):
    """TS projection (TS.cpp:332-344, b=c terms included) of camera-frame points
    P (...,3) with intrinsics broadcastable (...,9).  Returns (uv (...,2), D)."""
    X, Y, Z = P[..., 0], P[..., 1], P[..., 2]
    fx, fy, cx, cy, xi, lam, al, b, c = [intr[..., k] for k in range(9)]
    rho2 = X * X + Y * Y
    d1 = np.sqrt(rho2 + Z * Z)
    z1 = Z + xi * d1
    d2 = np.sqrt(rho2 + z1 * z1)
    z2 = z1 + lam * d2
    d3 = np.sqrt(rho2 + z2 * z2)
    D = z2 + al / (1 - al) * d3
    u = fx * X / D + b * Y / D + cx
    v = c * X / D + fy * Y / D + cy
    return np.stack([u, v], axis=-1), D


def make_board(width: int, height: int, square: float = SQUARE_MM) -> np.ndarray:
    j = np.arange(width * height)
    return np.stack([(j % width) * square, (j // width) * square], axis=-1).astype(np.float64)


def _random_unit_in_cone(rng, n, half_angle):
    """n unit vectors within `half_angle` (rad) of +z, uniform on the cap."""
    cos_t = rng.uniform(np.cos(half_angle), 1.0, n)
    sin_t = np.sqrt(1 - cos_t * cos_t)
    phi = rng.uniform(0, 2 * np.pi, n)
    return np.stack([sin_t * np.cos(phi), sin_t * np.sin(phi), cos_t], axis=-1)


def _rotation_z_to(v):
    """Rotation matrices taking +z to unit vectors v (n,3)."""
    z = np.array([0.0, 0.0, 1.0])
    axis = np.cross(np.broadcast_to(z, v.shape), v)
    s = np.linalg.norm(axis, axis=-1)
    c = v[..., 2]
    ang = np.arctan2(s, c)
    axis = np.where(s[..., None] < 1e-12, np.array([1.0, 0, 0]), axis / np.maximum(s, 1e-300)[..., None])
    return rodrigues(axis * ang[..., None])


@dataclass
class SyntheticProblem:
    problem: ProblemArrays
    gt_intrinsics: np.ndarray     # C x 9
    gt_cam_rt: np.ndarray         # C x 6
    gt_board_rt: np.ndarray       # F x 6
    init_intrinsics: np.ndarray
    init_cam_rt: np.ndarray
    init_board_rt: np.ndarray
    visible: np.ndarray           # C x F bool
    name: str = ""

    @property
    def num_observations(self):
        return self.problem.num_observations


def rig_from_calib_yaml(num_cameras: int):
    """Cameras 0..3 of calib.yaml (surround rig, OmniVidar-style)."""
    assert 1 <= num_cameras <= 4
    intr = CALIB_INTRINSICS[:num_cameras].copy()
    rt = np.zeros((num_cameras, 6))
    for m in range(num_cameras):
        rt[m, :3] = rotation_to_rvec(CALIB_TWC[m, :, :3])
        rt[m, 3:] = CALIB_TWC[m, :, 3]
    return intr, rt


def rig_forward_array(num_cameras: int, rng, spacing=70.0):
    """Forward-facing camera array (all cameras see every frame -> dense
    visibility, N = C*F*K exactly; SURVEY.md §8d 'dense variant')."""
    intr = np.stack([CALIB_INTRINSICS[m % 4] * (1 + 0.01 * rng.standard_normal(9))
                     for m in range(num_cameras)])
    intr[:, 7:] = 0.0
    cols = int(np.ceil(np.sqrt(num_cameras * 2)))
    rt = np.zeros((num_cameras, 6))
    for m in range(1, num_cameras):
        r, c = divmod(m, cols)
        centre = np.array([c * spacing, r * spacing, 0.0]) + rng.normal(0, 3.0, 3)
        rvec = rng.normal(0, 0.05, 3)
        R = rodrigues(rvec)
        rt[m, :3] = rvec
        rt[m, 3:] = -R @ centre          # P_cam = R (P_ref - centre)
    return intr, rt


def rig_ring(num_cameras: int, rng, radius=400.0):
    """Outward-facing ring (partial visibility, per-frame masks)."""
    intr = np.stack([CALIB_INTRINSICS[m % 4] * (1 + 0.01 * rng.standard_normal(9))
                     for m in range(num_cameras)])
    intr[:, 7:] = 0.0
    rt = np.zeros((num_cameras, 6))
    # camera 0 is the reference frame (rt == 0 exactly); others are placed on a
    # ring around a centre 'radius' behind camera 0.
    centre0 = np.array([0.0, 0.0, -radius])
    for m in range(1, num_cameras):
        yaw = 2 * np.pi * m / num_cameras
        Ry = rodrigues(np.array([0.0, yaw, 0.0]))          # cam m axes in ref frame
        Rpert = rodrigues(rng.normal(0, 0.03, 3))
        R_ref_from_cam = Ry @ Rpert
        pos = centre0 + R_ref_from_cam @ np.array([0.0, 0.0, radius]) + rng.normal(0, 5.0, 3)
        R = R_ref_from_cam.T                                # ref -> cam
        rt[m, :3] = rotation_to_rvec(R)
        rt[m, 3:] = -R @ pos
    return intr, rt


def _compose_views(intr, cam_rt, board_rt, board_xy):
    """Project every (camera, frame, corner): uv (C,F,K,2), D (C,F,K), incidence cos."""
    Rb = rodrigues(board_rt[:, :3])                            # F,3,3
    pb = np.concatenate([board_xy, np.zeros((board_xy.shape[0], 1))], axis=1)  # K,3
    Pw = np.einsum("fij,kj->fki", Rb, pb) + board_rt[:, None, 3:]              # F,K,3
    Rc = rodrigues(cam_rt[:, :3])                              # C,3,3
    Pc = np.einsum("cij,fkj->cfki", Rc, Pw) + cam_rt[:, None, None, 3:]        # C,F,K,3
    uv, D = ts_project(intr[:, None, None, :], Pc)
    cosang = Pc[..., 2] / np.linalg.norm(Pc, axis=-1)
    return uv, D, cosang


def generate(num_cameras: int, num_frames: int, board=(11, 8), rig="calib", dense=False,
             seed=0, noise_px=0.1, outlier_fraction=0.0, perturb=1.0, max_incidence_deg=110.0,
             margin_px=2.0, name="", frame_seed=0) -> SyntheticProblem:
    """Build one synthetic calibration problem.

    rig: 'calib' (calib.yaml cameras, <=4), 'ring' (outward ring, masks), 'array'
    (forward array, use with dense=True).  dense=True keeps only frames seen by
    every camera.  perturb scales the initial-value perturbation of §8d
    (0 = start at ground truth).  `seed` fixes the rig (cameras, their initial
    values); `frame_seed` varies the frames only, so ranks of a weak-scaling run
    can each draw their own frames for one common rig.
    """
    rig_rng = np.random.default_rng(np.random.PCG64(0x7C5C0000 + seed))
    rng = np.random.default_rng(np.random.PCG64([0x7C5C0000 + seed, 1 + frame_seed]))
    board_xy = make_board(*board)
    K = board_xy.shape[0]
    if rig == "calib":
        intr, cam_rt = rig_from_calib_yaml(num_cameras)
    elif rig == "ring":
        intr, cam_rt = rig_ring(num_cameras, rig_rng)
    elif rig == "array":
        intr, cam_rt = rig_forward_array(num_cameras, rig_rng)
    else:
        raise ValueError(rig)
    C = num_cameras
    Rc = rodrigues(cam_rt[:, :3])
    cos_max = np.cos(np.deg2rad(max_incidence_deg))
    centre_b = np.array([(board[0] - 1) * SQUARE_MM / 2, (board[1] - 1) * SQUARE_MM / 2, 0.0])

    board_rt_list, vis_list = [], []
    need = num_frames
    guard = 0
    while need > 0:
        guard += 1
        if guard > 200:
            raise RuntimeError("frame sampler did not converge; relax the rig geometry")
        n = max(64, int(need * (3.0 if dense else 1.6)))
        anchor = rng.integers(0, C, n) if not dense else np.zeros(n, dtype=np.int64)
        cone = np.deg2rad(25.0 if dense else 60.0)
        direction = _random_unit_in_cone(rng, n, cone)                 # in anchor camera frame
        rng_mm = rng.uniform(500.0 if dense else 300.0, 900.0, n)
        centre_cam = direction * rng_mm[:, None]
        # board normal within 45 deg of the viewing ray, random in-plane roll
        tilt = _random_unit_in_cone(rng, n, np.deg2rad(45.0))
        R_view = _rotation_z_to(direction)
        normal = np.einsum("nij,nj->ni", R_view, tilt)                 # board +z in camera frame
        R_n = _rotation_z_to(normal)
        roll = rng.uniform(0, 2 * np.pi, n)
        R_roll = rodrigues(np.stack([np.zeros(n), np.zeros(n), roll], axis=-1))
        R_bc = np.einsum("nij,njk->nik", R_n, R_roll)                  # board -> anchor camera
        t_bc = centre_cam - np.einsum("nij,j->ni", R_bc, centre_b)
        # board -> reference: P_ref = Rc^T (P_cam - tc)
        Ra = Rc[anchor]
        ta = cam_rt[anchor, 3:]
        R_br = np.einsum("nji,njk->nik", Ra, R_bc)
        t_br = np.einsum("nji,nj->ni", Ra, t_bc - ta)
        brt = np.concatenate([rotation_to_rvec(R_br), t_br], axis=1)
        uv, D, cosang = _compose_views(intr, cam_rt, brt, board_xy)
        inside = ((uv[..., 0] >= margin_px) & (uv[..., 0] <= IMAGE_W - 1 - margin_px) &
                  (uv[..., 1] >= margin_px) & (uv[..., 1] <= IMAGE_H - 1 - margin_px) &
                  (D > 0) & (cosang > cos_max))
        vis = inside.all(axis=-1)                                      # C, n (all-or-nothing)
        keep = vis.all(axis=0) if dense else vis.any(axis=0)
        idx = np.nonzero(keep)[0][:need]
        board_rt_list.append(brt[idx])
        vis_list.append(vis[:, idx])
        need -= len(idx)
    gt_board_rt = np.concatenate(board_rt_list, axis=0)
    visible = np.concatenate(vis_list, axis=1)
    F = num_frames

    uv, _, _ = _compose_views(intr, cam_rt, gt_board_rt, board_xy)     # C,F,K,2
    obs = uv + noise_px * rng.standard_normal(uv.shape)
    if outlier_fraction > 0:
        out = rng.random(uv.shape[:-1]) < outlier_fraction
        rand_px = np.stack([rng.uniform(0, IMAGE_W - 1, uv.shape[:-1]),
                            rng.uniform(0, IMAGE_H - 1, uv.shape[:-1])], axis=-1)
        obs = np.where(out[..., None], rand_px, obs)

    cams, frames = np.nonzero(visible)                                  # camera-major order
    view_camera = cams.astype(np.int32)
    view_frame = frames.astype(np.int32)
    obs_xy = obs[cams, frames]                                          # V,K,2

    e = rig_rng.standard_normal
    init_intr = intr.copy()
    init_intr[:, 0:2] *= 1 + perturb * 0.02 * e((C, 2))
    init_intr[:, 2:4] += perturb * 3.0 * e((C, 2))
    init_intr[:, 4] += perturb * 0.05 * e(C)
    init_intr[:, 5] += perturb * 0.03 * e(C)
    init_intr[:, 6] += perturb * 0.03 * e(C)
    init_cam_rt = cam_rt.copy()
    init_cam_rt[:, :3] += perturb * 0.02 * e((C, 3))
    init_cam_rt[:, 3:] += perturb * 5.0 * e((C, 3))
    init_cam_rt[0] = cam_rt[0]                 # camera 0 is the (constant) reference
    e = rng.standard_normal
    init_board_rt = gt_board_rt.copy()
    init_board_rt[:, :3] += perturb * 0.02 * e((F, 3))
    init_board_rt[:, 3:] += perturb * 5.0 * e((F, 3))

    prob = ProblemArrays(board_xy, view_camera, view_frame, obs_xy, C, F, fixed_camera=0)
    return SyntheticProblem(prob, intr, cam_rt, gt_board_rt, init_intr, init_cam_rt,
                            init_board_rt, visible, name)


# BASELINE.json configs --------------------------------------------------------
def config(idx: int, **kw) -> SyntheticProblem:
    """BASELINE.json `configs[idx-1]` (1-based like SURVEY.md §8)."""
    if idx == 1:    # single fisheye camera, 40 poses of a 9x6 board
        d = dict(num_cameras=1, num_frames=40, board=(9, 6), rig="calib", seed=1, name="cfg1-mono")
    elif idx == 2:  # 4-camera surround rig, 200 frames, 11x8 board
        d = dict(num_cameras=4, num_frames=200, board=(11, 8), rig="calib", seed=2, name="cfg2-rig4")
    elif idx == 3:  # 8 cameras x 5000 frames x 88 corners, dense: N = 3.52 M
        d = dict(num_cameras=8, num_frames=5000, board=(11, 8), rig="array", dense=True, seed=3,
                 name="cfg3-rig8-dense")
    elif idx == 4:  # stress: 16 cameras, ring, masks (frames scaled by caller)
        d = dict(num_cameras=16, num_frames=100000, board=(11, 8), rig="ring", seed=4,
                 name="cfg4-stress16")
    elif idx == 5:  # 4-camera rig with 5 % outliers (robust loss)
        d = dict(num_cameras=4, num_frames=200, board=(11, 8), rig="calib", seed=5,
                 outlier_fraction=0.05, name="cfg5-robust")
    else:
        raise ValueError(idx)
    d.update(kw)
    return generate(**d)


def shard_frames(sp: SyntheticProblem, rank: int, world: int):
    """Contiguous frame shard balanced by valid-observation count.
    Returns (ProblemArrays for this rank, frame index array)."""
    vis_per_frame = sp.visible.sum(axis=0)
    csum = np.cumsum(vis_per_frame)
    total = csum[-1]
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(csum, total * r / world, side="left")) + 1)
    bounds.append(sp.problem.num_frames)
    for r in range(1, len(bounds)):
        bounds[r] = max(bounds[r], bounds[r - 1])
    lo, hi = bounds[rank], bounds[rank + 1]
    p = sp.problem
    sel = (p.view_frame >= lo) & (p.view_frame < hi)
    local = ProblemArrays(p.board_xy, p.view_camera[sel], p.view_frame[sel] - lo, p.obs_xy[sel],
                          p.num_cameras, hi - lo, fixed_camera=p.fixed_camera)
    return local, np.arange(lo, hi)


def concat_frames(parts) -> SyntheticProblem:
    """One problem from several frame batches of the SAME rig (same `seed`, different
    `frame_seed`): frames are renumbered batch after batch and the views re-sorted
    camera-major / frame-minor as tscm_problem requires."""
    first = parts[0]
    C = first.problem.num_cameras
    cams, frames, obs, off = [], [], [], 0
    for sp in parts:
        assert np.array_equal(sp.gt_intrinsics, first.gt_intrinsics) and np.array_equal(sp.gt_cam_rt, first.gt_cam_rt)
        cams.append(sp.problem.view_camera)
        frames.append(sp.problem.view_frame + off)
        obs.append(sp.problem.obs_xy)
        off += sp.problem.num_frames
    cams, frames, obs = np.concatenate(cams), np.concatenate(frames), np.concatenate(obs, axis=0)
    order = np.argsort(cams, kind="stable")          # frames already increase within a camera
    prob = ProblemArrays(first.problem.board_xy, cams[order], frames[order], obs[order], C, off,
                         fixed_camera=first.problem.fixed_camera)
    return SyntheticProblem(prob, first.gt_intrinsics, first.gt_cam_rt,
                            np.concatenate([sp.gt_board_rt for sp in parts], axis=0),
                            first.init_intrinsics, first.init_cam_rt,
                            np.concatenate([sp.init_board_rt for sp in parts], axis=0),
                            np.concatenate([sp.visible for sp in parts], axis=1), first.name)


def _config_batch(args):
    """Worker of config_batched: plain arrays (a ProblemArrays holds ctypes pointers and
    cannot cross a process boundary)."""
    idx, kw = args
    sp = config(idx, **kw)
    p = sp.problem
    return (p.board_xy, p.view_camera, p.view_frame, p.obs_xy, p.num_cameras, p.num_frames,
            p.fixed_camera, sp.gt_intrinsics, sp.gt_cam_rt, sp.gt_board_rt, sp.init_intrinsics,
            sp.init_cam_rt, sp.init_board_rt, sp.visible, sp.name)


def _unpack_batch(t) -> SyntheticProblem:
    return SyntheticProblem(ProblemArrays(t[0], t[1], t[2], t[3], t[4], t[5], fixed_camera=t[6]), *t[7:])


def config_batched(idx: int, num_frames: int, batch: int = 10000, processes: int = 1, rank: int = 0,
                   world: int = 1, **kw) -> SyntheticProblem:
    """config(idx) with `num_frames` frames drawn in batches of `batch` (bounded host memory:
    config 4 at its full 100,000 frames would need tens of GB in one piece), optionally in
    parallel processes.  With world > 1 a rank keeps batches rank, rank + world, ... only: the
    union over the ranks is the same set of frames as the single-rank problem."""
    jobs, left, k = [], num_frames, 0
    while left > 0:
        n = min(batch, left)
        if k % world == rank:
            jobs.append((idx, dict(kw, num_frames=n, frame_seed=k)))
        left -= n
        k += 1
    if processes > 1 and len(jobs) > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(processes, len(jobs))) as pool:
            parts = [_unpack_batch(t) for t in pool.map(_config_batch, jobs)]
    else:
        parts = [_unpack_batch(_config_batch(j)) for j in jobs]
    return parts[0] if len(parts) == 1 else concat_frames(parts)


def mono_results(sp: SyntheticProblem, seed: int = 42, rot_sigma: float = 2e-3, t_sigma: float = 1.0):
    """Per-camera mono calibration results as MultiCalib's constructor consumes them
    (multi_calib.cpp:6-153): for every (camera, frame) with a detection the 3x3 [r1 r2 t] board
    pose in the camera frame = ground truth plus a small error, so that the pose-graph candidates
    differ and the selection matters.  Returns (worlds [K][3], intrinsics [C][9], has [C][F] uint8,
    Rt [C][F][3][3], pixels [C][F][K][2])."""
    p = sp.problem
    C, F, K = p.num_cameras, p.num_frames, p.corners_per_board
    has = sp.visible.astype(np.uint8)
    pixels = np.zeros((C, F, K, 2))
    pixels[p.view_camera, p.view_frame] = p.obs_xy
    worlds = np.concatenate([p.board_xy, np.zeros((K, 1))], axis=1)
    rng = np.random.default_rng(seed)
    Rc, Rb = rodrigues(sp.gt_cam_rt[:, :3]), rodrigues(sp.gt_board_rt[:, :3])
    dR = rodrigues(rng.normal(0, rot_sigma, (C, F, 3)))
    R = dR @ (Rc[:, None] @ Rb[None, :])
    t = np.einsum("cij,fj->cfi", Rc, sp.gt_board_rt[:, 3:]) + sp.gt_cam_rt[:, None, 3:] + rng.normal(0, t_sigma, (C, F, 3))
    Rt = np.stack([R[..., 0], R[..., 1], t], axis=-1) * has[..., None, None]
    return worlds, sp.gt_intrinsics.copy(), has, Rt, pixels
