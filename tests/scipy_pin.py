"""A third-party optimiser as the pin of the OPTIMUM: scipy.optimize.least_squares
(trust-region reflective, exact dense sub-problems, x_scale='jac') minimises the same sum of squares
from the same start.  Its Jacobian is scipy's own complex-step differentiation of the residual
function below — a plain-numpy restatement of /root/reference/multi_calib.h:146-195 written for
this test, sharing no code with the oracle, the numpy reference or the kernels.  Neither the LM
path nor the iteration counts are comparable (different trust-region algorithm); the minimum
and the minimiser are."""
import numpy as np
from scipy.optimize import least_squares


def _rotate(aa, p):
    """ceres::AngleAxisRotatePoint; aa (..., 3), p (..., 3); complex-step safe (no abs / norm)."""
    th2 = np.sum(aa * aa, axis=-1, keepdims=True)
    small = np.real(th2) <= np.finfo(float).eps
    th = np.sqrt(np.where(small, 1.0, th2))
    k = aa / th
    c, s = np.cos(th), np.sin(th)
    rod = p * c + np.cross(k, p) * s + k * np.sum(k * p, axis=-1, keepdims=True) * (1.0 - c)
    return np.where(small, p + np.cross(aa, p), rod)


class RigProblem:
    """x = [7 live intrinsics per camera | rt of every camera but the fixed one | board poses]."""

    def __init__(self, problem, intr, cam_rt, board_rt):
        self.p = problem
        self.C, self.F, self.K = problem.num_cameras, problem.num_frames, problem.corners_per_board
        self.intr0 = np.array(intr, dtype=float).reshape(self.C, 9)
        self.cam_rt0 = np.array(cam_rt, dtype=float).reshape(self.C, 6)
        self.board_rt0 = np.array(board_rt, dtype=float).reshape(self.F, 6)
        self.free = [m for m in range(self.C) if m != problem.fixed_camera]
        self.vc = np.asarray(problem.view_camera)
        self.vf = np.asarray(problem.view_frame)
        self.board = np.concatenate([np.asarray(problem.board_xy).reshape(self.K, 2), np.zeros((self.K, 1))], axis=1)
        self.obs = np.asarray(problem.obs_xy).reshape(-1, self.K, 2)

    def pack(self):
        return np.concatenate([self.intr0[:, :7].ravel(), self.cam_rt0[self.free].ravel(), self.board_rt0.ravel()])

    def unpack(self, x):
        C, F = self.C, self.F
        intr = np.array(self.intr0, dtype=x.dtype)
        intr[:, :7] = x[:7 * C].reshape(C, 7)
        cam_rt = np.array(self.cam_rt0, dtype=x.dtype)
        cam_rt[self.free] = x[7 * C:7 * C + 6 * len(self.free)].reshape(-1, 6)
        board_rt = x[7 * C + 6 * len(self.free):].reshape(F, 6)
        return intr, cam_rt, board_rt

    def residuals(self, x):
        intr, cam_rt, board_rt = self.unpack(x)
        b = board_rt[self.vf][:, None, :]                      # (V, 1, 6)
        c = cam_rt[self.vc][:, None, :]
        k = intr[self.vc][:, None, :]
        pw = _rotate(b[..., :3], self.board[None, :, :] + 0 * b[..., :3]) + b[..., 3:]
        pc = _rotate(c[..., :3], pw) + c[..., 3:]
        X, Y, Z = pc[..., 0], pc[..., 1], pc[..., 2]
        d1 = np.sqrt(X * X + Y * Y + Z * Z)
        z1 = Z + k[..., 4] * d1
        d2 = np.sqrt(X * X + Y * Y + z1 * z1)
        z2 = z1 + k[..., 5] * d2
        d3 = np.sqrt(X * X + Y * Y + z2 * z2)
        ksai = z2 + k[..., 6] / (1.0 - k[..., 6]) * d3
        ru = self.obs[..., 0] - (k[..., 0] * X / ksai + k[..., 2])
        rv = self.obs[..., 1] - (k[..., 1] * Y / ksai + k[..., 3])
        return np.stack([ru, rv], axis=-1).ravel()


def scipy_optimum(problem, intr, cam_rt, board_rt):
    rp = RigProblem(problem, intr, cam_rt, board_rt)
    res = least_squares(rp.residuals, rp.pack(), jac="cs", method="trf", tr_solver="exact", x_scale="jac",
                        ftol=1e-15, xtol=1e-15, gtol=1e-15, max_nfev=400)
    a, b, c = rp.unpack(res.x)
    return res.cost, a, b, c, res
