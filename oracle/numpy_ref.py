"""Independent numpy restatement of the calibration solve — the oracle's second opinion.

TEST INFRASTRUCTURE ONLY (see oracle/tscm_oracle.h).  PARITY UNPINNED: nothing in
/root/reference pins this path (no tests, no golden vectors; Ceres is absent).

Written separately from oracle/tscm_oracle.cpp, in a different style, so that a
slip in one is caught by the other:
  * forward-mode autodiff vectorised over all observations (value [N], grad [N,21])
    instead of per-residual Jets;
  * DENSE normal equations (no Schur complement) solved by Cholesky;
  * model cost change evaluated literally as -(J s).(r + J s / 2).
It follows the same published Ceres semantics and the
reference functors multi_calib.h:146-195 / TS.h:100-131.
Only practical for the small configs (1, 2, 5).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.linalg
import scipy.sparse as sp

DBL_EPS = np.finfo(np.float64).eps
DBL_MAX = np.finfo(np.float64).max
DBL_MIN = np.finfo(np.float64).tiny


class Dual:
    """value: [N]; grad: [N, P] (or None for a constant)."""
    __slots__ = ("a", "v")

    def __init__(self, a, v=None):
        self.a = np.asarray(a, dtype=np.float64)
        self.v = v

    @staticmethod
    def lift(x):
        return x if isinstance(x, Dual) else Dual(x)

    def _g(self, like):
        return self.v if self.v is not None else 0.0

    def __add__(self, o):
        o = Dual.lift(o)
        v = None if self.v is None and o.v is None else self._g(o) + o._g(self)
        return Dual(self.a + o.a, v)

    __radd__ = __add__

    def __sub__(self, o):
        o = Dual.lift(o)
        v = None if self.v is None and o.v is None else self._g(o) - o._g(self)
        return Dual(self.a - o.a, v)

    def __rsub__(self, o):
        return Dual.lift(o) - self

    def __mul__(self, o):
        o = Dual.lift(o)
        if self.v is None and o.v is None:
            return Dual(self.a * o.a)
        v = 0.0
        if o.v is not None:
            v = v + np.asarray(self.a)[..., None] * o.v
        if self.v is not None:
            v = v + self.v * np.asarray(o.a)[..., None]
        return Dual(self.a * o.a, v)

    __rmul__ = __mul__

    def __truediv__(self, o):
        o = Dual.lift(o)
        q = self.a / o.a
        if self.v is None and o.v is None:
            return Dual(q)
        num = self._g(o) - np.asarray(q)[..., None] * o._g(self)
        return Dual(q, num / np.asarray(o.a)[..., None])

    def __rtruediv__(self, o):
        return Dual.lift(o) / self


def dsqrt(x: Dual) -> Dual:
    t = np.sqrt(x.a)
    return Dual(t, None if x.v is None else x.v / (2.0 * t)[..., None])


def dsin(x: Dual) -> Dual:
    return Dual(np.sin(x.a), None if x.v is None else np.cos(x.a)[..., None] * x.v)


def dcos(x: Dual) -> Dual:
    return Dual(np.cos(x.a), None if x.v is None else -np.sin(x.a)[..., None] * x.v)


def dwhere(mask, x: Dual, y: Dual) -> Dual:
    a = np.where(mask, x.a, y.a)
    if x.v is None and y.v is None:
        return Dual(a)
    xv = x.v if x.v is not None else np.zeros_like(y.v)
    yv = y.v if y.v is not None else np.zeros_like(x.v)
    return Dual(a, np.where(mask[..., None], xv, yv))


def angle_axis_rotate(aa, pt):
    """ceres::AngleAxisRotatePoint, both branches evaluated and selected per element."""
    theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2]
    big = theta2.a > DBL_EPS
    # guarded theta for the Rodrigues branch (avoid 0/0 where the Taylor branch is selected)
    safe_t2 = dwhere(big, theta2, Dual(np.ones_like(theta2.a)))
    theta = dsqrt(safe_t2)
    c, s = dcos(theta), dsin(theta)
    inv = 1.0 / theta
    w = [aa[0] * inv, aa[1] * inv, aa[2] * inv]
    wxp = [w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2], w[0] * pt[1] - w[1] * pt[0]]
    tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (1.0 - c)
    rod = [pt[i] * c + wxp[i] * s + w[i] * tmp for i in range(3)]
    axp = [aa[1] * pt[2] - aa[2] * pt[1], aa[2] * pt[0] - aa[0] * pt[2], aa[0] * pt[1] - aa[1] * pt[0]]
    tay = [pt[i] + axp[i] for i in range(3)]
    return [dwhere(big, rod[i], tay[i]) for i in range(3)]


def rig_residual(cam_rt, board_rt, intr, bx, by, ox, oy):
    """multi_calib.h:146-195 on Dual vectors; cam_rt/board_rt lists of 6, intr list of 9."""
    p = [Dual(bx), Dual(by), Dual(np.zeros_like(bx))]
    pw = angle_axis_rotate(board_rt[:3], p)
    pw = [pw[i] + board_rt[3 + i] for i in range(3)]
    pc = angle_axis_rotate(cam_rt[:3], pw)
    pc = [pc[i] + cam_rt[3 + i] for i in range(3)]
    X, Y, Z = pc
    d1 = dsqrt(X * X + Y * Y + Z * Z)
    z1 = Z + intr[4] * d1
    d2 = dsqrt(X * X + Y * Y + z1 * z1)
    z2 = Z + intr[4] * d1 + intr[5] * d2
    d3 = dsqrt(X * X + Y * Y + z2 * z2)
    ksai = Z + intr[4] * d1 + intr[5] * d2 + intr[6] / (1.0 - intr[6]) * d3
    px = intr[0] * X / ksai + intr[2]
    py = intr[1] * Y / ksai + intr[3]
    return Dual(ox) - px, Dual(oy) - py


@dataclass
class Problem:
    C: int
    F: int
    K: int
    board_xy: np.ndarray
    view_camera: np.ndarray
    view_frame: np.ndarray
    obs_xy: np.ndarray
    fixed_camera: int = 0

    @staticmethod
    def from_arrays(pa):
        return Problem(pa.num_cameras, pa.num_frames, pa.corners_per_board, pa.board_xy.copy(),
                       pa.view_camera.copy(), pa.view_frame.copy(), pa.obs_xy.copy(), pa.fixed_camera)


def evaluate(P: Problem, intr, cam_rt, board_rt, want_jacobian=True):
    """Residuals [N,2] and autodiff Jacobian [N,2,21] (camera_rt 6, chessboard_rt 6, intrinsic 9)."""
    V, K = len(P.view_camera), P.K
    cam = np.repeat(P.view_camera, K)
    frm = np.repeat(P.view_frame, K)
    bx = np.tile(P.board_xy[:, 0], V)
    by = np.tile(P.board_xy[:, 1], V)
    ox = P.obs_xy.reshape(-1, 2)[:, 0]
    oy = P.obs_xy.reshape(-1, 2)[:, 1]
    N = V * K

    def seed(vals, col):
        if not want_jacobian:
            return Dual(vals)
        g = np.zeros((N, 21))
        g[:, col] = 1.0
        return Dual(vals, g)

    crt = [seed(cam_rt[cam, k], k) for k in range(6)]
    brt = [seed(board_rt[frm, k], 6 + k) for k in range(6)]
    it = [seed(intr[cam, k], 12 + k) for k in range(9)]
    ru, rv = rig_residual(crt, brt, it, bx, by, ox, oy)
    r = np.stack([ru.a, rv.a], axis=1)
    if not want_jacobian:
        return r, None
    J = np.stack([ru.v, rv.v], axis=1)
    if P.fixed_camera >= 0:
        J[cam == P.fixed_camera, :, 0:6] = 0.0
    return r, J


def loss_rho(kind, a, s):
    b = a * a
    if kind == 1:
        r = np.sqrt(np.maximum(s, DBL_MIN))
        out = s > b
        rho0 = np.where(out, 2.0 * a * r - b, s)
        rho1 = np.where(out, np.maximum(DBL_MIN, a / r), 1.0)
        rho2 = np.where(out, -rho1 / (2.0 * np.where(out, s, 1.0)), 0.0)
    elif kind == 2:
        c = 1.0 / b
        summ = 1.0 + s * c
        inv = 1.0 / summ
        rho0, rho1, rho2 = b * np.log(summ), np.maximum(DBL_MIN, inv), -c * inv * inv
    else:
        rho0, rho1, rho2 = s, np.ones_like(s), np.zeros_like(s)
    return rho0, rho1, rho2


def apply_loss(kind, a, r, J):
    """Corrector (corrector.cc).  Returns cost, corrected r, corrected J."""
    s = (r * r).sum(axis=1)
    if kind == 0:
        return 0.5 * s.sum(), r, J
    rho0, rho1, rho2 = loss_rho(kind, a, s)
    sq = np.sqrt(rho1)
    simple = (s == 0.0) | (rho2 <= 0.0)
    D = 1.0 + 2.0 * s * rho2 / rho1
    alpha = np.where(simple, 0.0, 1.0 - np.sqrt(np.where(simple, 1.0, D)))
    res_scale = np.where(simple, sq, sq / (1.0 - alpha))
    asn = np.where(simple, 0.0, alpha / np.where(s == 0, 1.0, s))
    Jc = None
    if J is not None:
        rtj = np.einsum("na,nak->nk", r, J)
        Jc = sq[:, None, None] * (J - asn[:, None, None] * r[:, :, None] * rtj[:, None, :])
    rc = r * res_scale[:, None]
    return 0.5 * rho0.sum(), rc, Jc


class Layout:
    """x = [board poses | per camera: rt (unless fixed), intrinsic(9)] (Ceres' Schur order)."""

    def __init__(self, P: Problem):
        self.P = P
        self.n_e = 6 * P.F
        off, self.cam_off, self.rt_sz = 0, [], []
        for c in range(P.C):
            self.cam_off.append(off)
            self.rt_sz.append(0 if c == P.fixed_camera else 6)
            off += self.rt_sz[-1] + 9
        self.n_c = off
        self.n = self.n_e + self.n_c

    def pack(self, intr, cam_rt, board_rt):
        x = np.zeros(self.n)
        x[:self.n_e] = board_rt.reshape(-1)
        for c in range(self.P.C):
            o = self.n_e + self.cam_off[c]
            if self.rt_sz[c]:
                x[o:o + 6] = cam_rt[c]
            x[o + self.rt_sz[c]:o + self.rt_sz[c] + 9] = intr[c]
        return x

    def unpack(self, x, cam_rt_fixed):
        P = self.P
        board_rt = x[:self.n_e].reshape(P.F, 6).copy()
        intr, cam_rt = np.zeros((P.C, 9)), cam_rt_fixed.copy()
        for c in range(P.C):
            o = self.n_e + self.cam_off[c]
            if self.rt_sz[c]:
                cam_rt[c] = x[o:o + 6]
            intr[c] = x[o + self.rt_sz[c]:o + self.rt_sz[c] + 9]
        return intr, cam_rt, board_rt

    def column_index(self):
        """[N, 21] global column of every local Jacobian column (-1 for the constant block)."""
        P = self.P
        cam = np.repeat(P.view_camera, P.K)
        frm = np.repeat(P.view_frame, P.K)
        N = len(cam)
        idx = np.full((N, 21), -1, dtype=np.int64)
        cam_off = np.array(self.cam_off)[cam] + self.n_e
        rts = np.array(self.rt_sz)[cam]
        for k in range(6):
            idx[:, k] = np.where(rts > 0, cam_off + k, -1)
            idx[:, 6 + k] = 6 * frm + k
        for k in range(9):
            idx[:, 12 + k] = cam_off + rts + k
        return idx


@dataclass
class Summary:
    termination: str = "NO_CONVERGENCE"
    cost: list = field(default_factory=list)
    radius: list = field(default_factory=list)
    gradient_max_norm: list = field(default_factory=list)
    step_norm: list = field(default_factory=list)
    flags: list = field(default_factory=list)
    num_successful_steps: int = 0
    num_unsuccessful_steps: int = 0
    initial_cost: float = 0.0
    final_cost: float = 0.0

    @property
    def num_iterations(self):
        return len(self.cost)


def solve(P: Problem, intr, cam_rt, board_rt, max_num_iterations=50, function_tolerance=1e-6,
          gradient_tolerance=1e-10, parameter_tolerance=1e-8, initial_radius=1e4, max_radius=1e16,
          min_radius=1e-32, min_relative_decrease=1e-3, min_lm_diagonal=1e-6, max_lm_diagonal=1e32,
          max_invalid=5, jacobi_scaling=True, loss_type=0, loss_scale=1.0,
          ptol_needs_success=False):
    """Ceres TRUST_REGION / LEVENBERG_MARQUARDT with dense normal equations."""
    L = Layout(P)
    col = L.column_index()
    rows = np.repeat(np.arange(col.shape[0] * 2).reshape(-1, 2), 21, axis=1).reshape(-1, 2, 21) \
        if False else None
    N = col.shape[0]
    row_u = np.repeat(2 * np.arange(N), 21).reshape(N, 21)
    valid = col >= 0

    def sparse_J(J):
        data = np.concatenate([J[:, 0, :][valid], J[:, 1, :][valid]])
        r = np.concatenate([row_u[valid], row_u[valid] + 1])
        c = np.concatenate([col[valid], col[valid]])
        return sp.csr_matrix((data, (r, c)), shape=(2 * N, L.n))

    def eval_full(x):
        i, c, b = L.unpack(x, cam_rt)
        r, J = evaluate(P, i, c, b, True)
        cost, rc, Jc = apply_loss(loss_type, loss_scale, r, J)
        return cost, rc.reshape(-1), sparse_J(Jc)

    def eval_cost(x):
        i, c, b = L.unpack(x, cam_rt)
        r, _ = evaluate(P, i, c, b, False)
        cost, _, _ = apply_loss(loss_type, loss_scale, r, None)
        return cost if np.isfinite(cost) else DBL_MAX

    S = Summary()
    x = L.pack(intr, cam_rt, board_rt)
    x_cost, res, Jm = eval_full(x)
    g = Jm.T @ res
    if jacobi_scaling:
        scale = 1.0 / (1.0 + np.sqrt(np.asarray(Jm.multiply(Jm).sum(axis=0)).ravel()))
    else:
        scale = np.ones(L.n)
    Js = Jm @ sp.diags(scale)
    gmax = np.abs(x - (x + (-g))).max()
    x_norm = np.linalg.norm(x)
    radius, decrease = initial_radius, 2.0
    S.initial_cost = x_cost
    iteration, nci, one_success = 0, 0, False
    it = dict(cost=x_cost, gmax=gmax, step_norm=0.0, valid=True, success=True)
    best_x = x.copy()
    while True:
        # finalize
        if it["success"]:
            S.num_successful_steps += 1
            best_x = x.copy()
        else:
            S.num_unsuccessful_steps += 1
        S.cost.append(it["cost"]); S.radius.append(radius); S.gradient_max_norm.append(it["gmax"])
        S.step_norm.append(it["step_norm"]); S.flags.append(int(it["valid"]) | 2 * int(it["success"]))
        if iteration >= max_num_iterations:
            S.termination = "NO_CONVERGENCE"; break
        if it["success"] and it["gmax"] <= gradient_tolerance:
            S.termination = "CONVERGENCE"; break
        if not radius > min_radius:
            S.termination = "CONVERGENCE"; break
        iteration += 1
        prev_gmax = it["gmax"]
        # LM step on dense normal equations
        H = (Js.T @ Js).toarray()
        diag = np.clip(np.diag(H), min_lm_diagonal, max_lm_diagonal)
        D = np.sqrt(diag / radius)
        A = H + np.diag(D * D)
        rhs = Js.T @ res
        ok = True
        try:
            cf = scipy.linalg.cho_factor(A, lower=True, check_finite=False)
            y = scipy.linalg.cho_solve(cf, rhs, check_finite=False)
            ok = bool(np.all(np.isfinite(y)))
        except np.linalg.LinAlgError:
            ok = False
        valid_step = False
        if ok:
            step = -y
            m = Js @ step
            model_cost_change = -float(m @ (res + m / 2.0))
            valid_step = model_cost_change > 0.0
        if not valid_step:
            nci += 1
            if nci >= max_invalid:
                S.termination = "FAILURE"; break
            radius /= decrease; decrease *= 2.0
            it = dict(cost=x_cost, gmax=prev_gmax, step_norm=0.0, valid=False, success=False)
            continue
        nci = 0
        delta = step * scale
        xc = x + delta
        cand = eval_cost(xc)
        step_norm = np.linalg.norm(x - xc)
        if (not ptol_needs_success or one_success) and \
                step_norm <= parameter_tolerance * (x_norm + parameter_tolerance):
            S.termination = "CONVERGENCE"; break
        if (not ptol_needs_success or one_success) and abs(x_cost - cand) <= function_tolerance * x_cost:
            S.termination = "CONVERGENCE"; break
        rho = -DBL_MAX if cand >= DBL_MAX else (x_cost - cand) / model_cost_change
        if rho > min_relative_decrease:
            x = xc
            x_norm = np.linalg.norm(x)
            x_cost, res, Jm = eval_full(x)
            g = Jm.T @ res
            Js = Jm @ sp.diags(scale)
            gmax = np.abs(x - (x + (-g))).max()
            radius = min(max_radius, radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3))
            decrease = 2.0
            one_success = True
            it = dict(cost=x_cost, gmax=gmax, step_norm=step_norm, valid=True, success=True)
        else:
            radius /= decrease; decrease *= 2.0
            it = dict(cost=cand, gmax=prev_gmax, step_norm=step_norm, valid=True, success=False)
    S.final_cost = min(S.cost) if S.cost else S.initial_cost
    i, c, b = L.unpack(best_x, cam_rt)
    return i, c, b, S


# --- 50-digit reference for single observations (mpmath) ----------------------
def mp_residual(cam_rt, board_rt, intr, bx, by, ox, oy):
    import mpmath as mp

    def rot(aa, p):
        t2 = aa[0] ** 2 + aa[1] ** 2 + aa[2] ** 2
        if t2 > mp.mpf(DBL_EPS):
            t = mp.sqrt(t2)
            w = [a / t for a in aa]
            c, s = mp.cos(t), mp.sin(t)
            wxp = [w[1] * p[2] - w[2] * p[1], w[2] * p[0] - w[0] * p[2], w[0] * p[1] - w[1] * p[0]]
            d = (w[0] * p[0] + w[1] * p[1] + w[2] * p[2]) * (1 - c)
            return [p[i] * c + wxp[i] * s + w[i] * d for i in range(3)]
        axp = [aa[1] * p[2] - aa[2] * p[1], aa[2] * p[0] - aa[0] * p[2], aa[0] * p[1] - aa[1] * p[0]]
        return [p[i] + axp[i] for i in range(3)]

    p = [mp.mpf(bx), mp.mpf(by), mp.mpf(0)]
    pw = rot(board_rt[:3], p)
    pw = [pw[i] + board_rt[3 + i] for i in range(3)]
    pc = rot(cam_rt[:3], pw)
    X, Y, Z = [pc[i] + cam_rt[3 + i] for i in range(3)]
    d1 = mp.sqrt(X * X + Y * Y + Z * Z)
    z1 = Z + intr[4] * d1
    d2 = mp.sqrt(X * X + Y * Y + z1 * z1)
    z2 = z1 + intr[5] * d2
    d3 = mp.sqrt(X * X + Y * Y + z2 * z2)
    D = z2 + intr[6] / (1 - intr[6]) * d3
    return [mp.mpf(ox) - (intr[0] * X / D + intr[2]), mp.mpf(oy) - (intr[1] * Y / D + intr[3])]


def mp_jacobian(cam_rt, board_rt, intr, bx, by, ox, oy, digits=50):
    """Central differences in `digits`-digit arithmetic: [2, 21]."""
    import mpmath as mp
    mp.mp.dps = digits
    params = [mp.mpf(float(v)) for v in list(cam_rt) + list(board_rt) + list(intr)]
    h = mp.mpf(10) ** (-digits // 3)
    J = np.zeros((2, 21))

    def f(pv):
        return mp_residual(pv[0:6], pv[6:12], pv[12:21], bx, by, ox, oy)

    for k in range(21):
        pp, pm = list(params), list(params)
        pp[k] += h
        pm[k] -= h
        fp, fm = f(pp), f(pm)
        for a in range(2):
            J[a, k] = float((fp[a] - fm[a]) / (2 * h))
    r = f(params)
    return np.array([float(r[0]), float(r[1])]), J
