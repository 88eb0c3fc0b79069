"""First-light diagnostic for the GPU box: prints parity numbers instead of asserting,
so one gpurun call yields as much information as possible."""
import os, sys, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tscm_calib_b200 import capi, synth
from oracle import oracle


def section(name):
    print(f"\n=== {name} ===", flush=True)


def jacobian_check(sp):
    s = capi.Solver(sp.problem, capi.default_options())
    s.set_parameters(sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
    r, J, cost = s.eval_jacobian()
    r0, J0, c0 = oracle.eval_jacobian(sp.problem, sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
    print("residual max abs diff", np.abs(r - r0).max(), "J max abs diff", np.abs(J - J0).max(),
          "J rel", (np.abs(J - J0) / (np.abs(J0) + 1e-3)).max(), "cost", cost, c0, abs(cost - c0) / c0)
    return s


def reduced_check(sp, s):
    radius = 1e4
    lhs, rhs = s.reduced_system(radius)
    lhs0, rhs0 = oracle.reduced_system(sp.problem, sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt, radius)
    idx = oracle.live_reduced_index(sp.problem)
    lhs0, rhs0 = lhs0[np.ix_(idx, idx)], rhs0[idx]
    sc = np.sqrt(np.abs(np.diag(lhs0)))
    print("reduced n", lhs.shape[0], "lhs rel diff", (np.abs(lhs - lhs0) / np.outer(sc, sc)).max(),
          "rhs rel diff", np.abs(rhs - rhs0).max() / np.abs(rhs0).max())


def solve_check(sp, opt, tag):
    t = time.time()
    a, b, c, s = capi.solve(sp.problem, sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt, opt)
    tg = time.time() - t
    t = time.time()
    a0, b0, c0, s0 = oracle.solve(sp.problem, sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt, opt)
    tc = time.time() - t
    print(tag, "gpu", s.termination, s.num_iterations, "oracle", s0.termination, s0.num_iterations,
          f"time gpu {tg:.3f}s cpu {tc:.3f}s")
    n = min(len(s.cost), len(s0.cost))
    print(" cost gpu   ", s.cost[:8])
    print(" cost oracle", s0.cost[:8])
    if n:
        print(" max rel cost diff", (np.abs(s.cost[:n] - s0.cost[:n]) / s0.cost[:n]).max())
        print(" radius gpu", s.radius[:6], "oracle", s0.radius[:6])
        print(" gmax gpu", s.gradient_max_norm[:4], "oracle", s0.gradient_max_norm[:4])
        print(" step gpu", s.step_norm[:4], "oracle", s0.step_norm[:4])
    for name, x, x0 in (("intr", a, a0), ("cam_rt", b, b0), ("board_rt", c, c0)):
        print(f" {name} max rel diff", (np.abs(x - x0) / (np.abs(x0) + 1e-9)).max(), "abs", np.abs(x - x0).max())


def main():
    lib = capi.load_library()
    print(lib.tscm_version().decode())
    for idx in (1, 2):
        sp = synth.config(idx)
        section(f"config {idx} {sp.name} N={sp.num_observations}")
        try:
            s = jacobian_check(sp)
            reduced_check(sp, s)
            s.close()
            opt = capi.default_options(max_num_iterations=100 if idx == 1 else 50)
            solve_check(sp, opt, "solve")
        except Exception:
            traceback.print_exc()
    section("config 5 robust (huber 1px)")
    try:
        sp = synth.config(5)
        solve_check(sp, capi.default_options(loss_type="huber", loss_scale=1.0), "huber")
        solve_check(sp, capi.default_options(loss_type="cauchy", loss_scale=1.0), "cauchy")
    except Exception:
        traceback.print_exc()
    section("timing config 3 (8 cams x 5000 frames x 88)")
    try:
        t = time.time()
        sp = synth.config(3)
        print("generated in", time.time() - t, "N", sp.num_observations)
        opt = capi.default_options(max_num_iterations=1000000, disable_tolerances=1)
        s = capi.Solver(sp.problem, opt)
        s.set_parameters(sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
        for stage, name in ((0, "k_eval"), (5, "eval pass"), (1, "schur+reduce"), (2, "solve"), (3, "backsub")):
            s.time_stage(stage, 3)
            print(f" stage {name}: {s.time_stage(stage, 20):.4f} ms")
        s.time_stage(4, 5)
        ms = s.time_stage(4, 30)
        print(f" LM iteration: {ms:.4f} ms -> {1000.0 / ms:.1f} it/s")
        sol = capi.default_options(max_num_iterations=50)
        s.set_options(sol)
        s.set_parameters(sp.init_intrinsics, sp.init_cam_rt, sp.init_board_rt)
        t = time.time()
        res = s.run()
        print(" full solve", res, "wall", time.time() - t)
        print(" cost", res.cost)
        print(" reproj", s.reprojection_error())
        s.close()
    except Exception:
        traceback.print_exc()


if __name__ == "__main__":
    main()
