"""Golden Jacobians from SYMBOLIC differentiation (sympy 1.14, evaluated with 50-digit mpmath) of
the reference's rig residual functor, /root/reference/multi_calib.h:146-195, including
ceres::AngleAxisRotatePoint with its first-order branch for theta^2 <= DBL_EPSILON.

Nothing here was derived by hand or shares code with the oracle, the numpy reference or the
kernels: the functor is restated as a sympy expression and differentiated by sympy.  The points
are those of tests/golden/jacobian_mpmath.npz (50-digit central differences), so the two
goldens also pin each other.

    python tests/golden/make_golden_sympy.py     ->  tests/golden/jacobian_sympy.npz
"""
import os

import mpmath
import numpy as np
import sympy as sp

HERE = os.path.dirname(os.path.abspath(__file__))
mpmath.mp.dps = 50


def rotate(aa, pt, taylor):
    """ceres::AngleAxisRotatePoint (rotation.h): Rodrigues' formula, or pt + aa x pt near zero."""
    w = sp.Matrix(aa)
    p = sp.Matrix(pt)
    if taylor:
        return p + w.cross(p)
    theta = sp.sqrt(w.dot(w))
    k = w / theta
    c, s = sp.cos(theta), sp.sin(theta)
    return p * c + k.cross(p) * s + k * (k.dot(p)) * (1 - c)


def residual_expr(taylor_cam, taylor_board):
    crt = sp.symbols("c0:6")          # camera_rt_   (multi_calib.h:147)
    brt = sp.symbols("b0:6")          # chessbaord_rt_
    K = sp.symbols("k0:9")            # intrinsic_: fx fy cx cy xi lambda alpha b c
    bx, by, ox, oy = sp.symbols("bx by ox oy")
    pw = rotate(brt[:3], (bx, by, 0), taylor_board) + sp.Matrix(brt[3:])        # :153-161
    pc = rotate(crt[:3], list(pw), taylor_cam) + sp.Matrix(crt[3:])              # :163-167
    X, Y, Z = pc
    d1 = sp.sqrt(X * X + Y * Y + Z * Z)                                           # :170
    d2 = sp.sqrt(X * X + Y * Y + (Z + K[4] * d1) ** 2)                            # :171
    d3 = sp.sqrt(X * X + Y * Y + (Z + K[4] * d1 + K[5] * d2) ** 2)                # :172
    ksai = Z + K[4] * d1 + K[5] * d2 + K[6] / (1 - K[6]) * d3                     # :173
    r = sp.Matrix([ox - (K[0] * X / ksai + K[2]), oy - (K[1] * Y / ksai + K[3])])  # :177-178,192-193
    params = list(crt) + list(brt) + list(K)
    J = r.jacobian(params)
    args = params + [bx, by, ox, oy]
    return sp.lambdify(args, r, "mpmath"), sp.lambdify(args, J, "mpmath")


def main():
    z = np.load(os.path.join(HERE, "jacobian_mpmath.npz"))
    n = len(z["r"])
    cache = {}
    R = np.zeros((n, 2))
    J = np.zeros((n, 2, 21))
    for k in range(n):
        tc = bool(np.dot(z["cam_rt"][k, :3], z["cam_rt"][k, :3]) <= np.finfo(float).eps)
        tb = bool(np.dot(z["board_rt"][k, :3], z["board_rt"][k, :3]) <= np.finfo(float).eps)
        if (tc, tb) not in cache:
            cache[(tc, tb)] = residual_expr(tc, tb)
        fr, fJ = cache[(tc, tb)]
        vals = [mpmath.mpf(float(v)) for v in np.concatenate([z["cam_rt"][k], z["board_rt"][k], z["intr"][k],
                                                              z["board"][k], z["obs"][k]])]
        r = fr(*vals)
        Jm = fJ(*vals)
        R[k] = [float(r[0]), float(r[1])]
        for i in range(2):
            for j in range(21):
                J[k, i, j] = float(Jm[i, j])
    out = os.path.join(HERE, "jacobian_sympy.npz")
    np.savez_compressed(out, cam_rt=z["cam_rt"], board_rt=z["board_rt"], intr=z["intr"], board=z["board"],
                        obs=z["obs"], r=R, J=J)
    d = np.max(np.abs(J - z["J"]) / np.maximum(np.abs(z["J"]), 1.0))
    print(f"{n} points, {sum(1 for k in cache)} branch combinations; symbolic vs 50-digit central differences: "
          f"max scaled Jacobian difference {d:.2e}, residual difference {np.max(np.abs(R - z['r'])):.2e}")


if __name__ == "__main__":
    main()
