"""Pins that the builder did not write: scipy's trust-region least squares reaches the optimum the
oracle (and, on the GPU box, the CUDA solve) reaches — cost to 1e-9 relative, parameters to 1e-7 —
and scipy's own complex-step Jacobian of an independent restatement of the functor certifies the
solution of a larger problem as a stationary point."""
import numpy as np
import pytest

from scipy_pin import RigProblem, scipy_optimum
from tscm_calib_b200 import capi, synth

# run to the bottom: the defaults (function tolerance 1e-6) stop ~1e-6 above the minimum
TIGHT = dict(max_num_iterations=200, function_tolerance=1e-16, parameter_tolerance=1e-14, gradient_tolerance=1e-13)
# problems whose minimum is sharp enough for two different trust-region methods to land on the same
# point (a 24-frame rig has a valley so flat that both still creep after 200 iterations)
CASES = {"cfg1-mono": lambda: synth.config(1), "cfg2-rig4-60frames": lambda: synth.config(2, num_frames=60)}


def check_same_optimum(name, solve):
    sp = CASES[name]()
    init = (sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
    cost, a0, b0, c0, res = scipy_optimum(sp.problem, *init)
    assert res.status > 0, res.message
    a, b, c, s = solve(sp.problem, *init, capi.default_options(**TIGHT))
    assert abs(s.final_cost - cost) <= 1e-9 * cost, (s.final_cost, cost)
    # The minimum VALUE agrees to 1e-9; the minimiser to 1e-5 only: scipy stops on the cost (its ftol /
    # xtol tests bottom out near sqrt(eps)), which leaves ~1e-6 of slack along the flat focal-length /
    # board-distance direction (a 1e-4 px change of fx moves the cost by 1e-12 relative).  Entries that
    # are nearly zero (lambda of the mono fixture is -4e-7) are compared absolutely.  The sharper
    # statement about the minimiser is the stationarity test below.
    for x, y in ((a, a0), (b, b0), (c, c0)):
        np.testing.assert_allclose(x, y, rtol=1e-5, atol=1e-6)
    # and the optimum is a real one: RMS at the noise floor of the synthetic data (0.1 px per axis)
    rms = np.sqrt(2 * cost / sp.num_observations)
    assert 0.12 < rms < 0.16, rms


def check_stationary(solve):
    """config 3 shape (8 cameras, 40 frames): at the solver's solution every column of scipy's
    complex-step Jacobian is orthogonal to the residual to 1e-7 (|J_j . r| <= 1e-7 |J_j| |r|)."""
    from scipy.optimize._numdiff import approx_derivative
    sp = synth.config(3, num_frames=40)
    a, b, c, s = solve(sp.problem, sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt, capi.default_options(**TIGHT))
    assert s.termination == "CONVERGENCE"
    rp = RigProblem(sp.problem, a, b, c)
    x = rp.pack()
    r = rp.residuals(x)
    assert abs(0.5 * r @ r - s.final_cost) <= 1e-9 * s.final_cost          # the same objective
    J = approx_derivative(rp.residuals, x, method="cs")
    cosine = np.abs(J.T @ r) / (np.linalg.norm(J, axis=0) * np.linalg.norm(r))
    assert cosine.max() < 1e-7, cosine.max()


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_reaches_scipys_optimum(oracle, name):
    check_same_optimum(name, lambda p, a, b, c, o: oracle.solve(p, a, b, c, o))


def test_oracle_solution_is_stationary_for_scipys_jacobian(oracle):
    check_stationary(lambda p, a, b, c, o: oracle.solve(p, a, b, c, o, num_threads=4))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_gpu_solve_reaches_scipys_optimum(name):
    check_same_optimum(name, lambda p, a, b, c, o: capi.solve(p, a, b, c, o))


@pytest.mark.gpu
def test_gpu_solution_is_stationary_for_scipys_jacobian():
    check_stationary(lambda p, a, b, c, o: capi.solve(p, a, b, c, o))
