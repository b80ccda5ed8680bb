"""BASELINE config 4 as a parity case: tric model (steering-angle and steering-rate constraints), SQP iterations
with warm start by shifting the iterate, 200 closed-loop ticks against the nominal RK4 plant.  Not a behaviour
of the reference (its wrapper does one RTI per tick and never shifts the iterate): the parity
target is the oracle driven through the same protocol.

Protocol per tick: SQP = repeated RTI steps from the current iterate until the step is
below 1e-8 for every instance or MAX_SQP steps were taken; apply u_0 to the plant x+ = phi_RK4(x, u_0); shift the
iterate and the reference window by one stage (last stage repeated); x0 <- x+.

The closed loop is driven by the oracle and the CUDA solver is given the oracle's state and iterate before every
RTI step ("teacher forcing"), so that every one of the ~1000 x B solves is compared on identical inputs.  Letting
both loops run free does not test the solver: with the full-step SQP chattering between the steering-rate bounds
on some instances, rounding-level differences (1e-13) grow by about a decade per tick and the two closed loops
separate after 10-30 ticks whatever the solver (measured: 7 of 24 instances above 1e-9 by tick 31, before the
first QP-iteration-count mismatch)."""
import numpy as np
import pytest

from helpers import instances, parity_report

pytestmark = pytest.mark.gpu

MAX_SQP = 6


def _shift(a):
    """a[:, k] <- a[:, k+1], last stage repeated"""
    out = np.empty_like(a)
    out[:, :-1] = a[:, 1:]
    out[:, -1] = a[:, -1]
    return out


def test_tric_sqp_closed_loop_200_ticks(oracle_mod):
    from nmpc_nav_control_b200.solver import BatchedRtiSolver
    name, B, T = "tric", 24, 200
    spec, x0, yref, _ = instances(name, 700, B)
    tb = spec.codegen_defaults()
    orc = oracle_mod.Oracle(name, tb)
    p0 = np.ascontiguousarray(tb["p"][0])
    s = BatchedRtiSolver(spec, B)
    xo = np.zeros((B, spec.n + 1, spec.nx)); uo = np.zeros((B, spec.n, spec.nu))
    n_rti = iter_mismatch = bad = illcond = 0
    worst = 0.0
    converged_ticks = 0
    traj = np.zeros((T, B, spec.nx)); u_applied = np.zeros((T, B, spec.nu))
    for t in range(T):
        for j in range(MAX_SQP):
            s.set_iterate(xo, uo)                                   # the oracle's iterate, state and references
            out = s.solve_host(x0, yref)
            xprev, uprev = xo.copy(), uo.copy()
            ref = orc.rti_batch(np.ascontiguousarray(x0), np.ascontiguousarray(yref), xo, uo)
            xg, ug = s.get_iterate(B)
            assert (out["status"] == 0).all() and (ref["status"] == 0).all(), (t, j)
            n_rti += B
            iter_mismatch += int((out["qp_iter"] != ref["qp_iter"]).sum())
            lr = ref["lin_res"]
            illcond += int((lr > 1e-10).sum())
            same_path = out["qp_iter"] == ref["qp_iter"]            # a termination flip changes the result by up to ~1e-5
            nbx, ex = parity_report(xg[same_path], xo[same_path], lr[same_path])
            nbu, eu = parity_report(ug[same_path], uo[same_path], lr[same_path])
            bad += nbx + nbu
            worst = max(worst, ex, eu)
            if max(np.abs(xo - xprev).max(), np.abs(uo - uprev).max()) < 1e-8:
                converged_ticks += 1
                break
        traj[t] = x0; u_applied[t] = uo[:, 0]
        x0 = np.stack([orc.discrete_map(x0[i], uo[i, 0], p0, tb["dt"])[0] for i in range(B)])
        xo, uo, yref = _shift(xo), _shift(uo), _shift(yref)
    deg = np.pi / 180.0
    print(f"closed loop tric: {T} ticks x {B} instances, {n_rti} RTI solves compared; qp_iter mismatches {iter_mismatch}, "
          f"solves outside 1e-9 (equal iteration path) {bad}, worst |diff| {worst:.2e}, ill-conditioned solves {illcond}; "
          f"ticks with SQP converged below 1e-8 within {MAX_SQP} steps: {converged_ticks}; "
          f"max |alpha_ref| {np.abs(traj[:, :, 6]).max() / deg:.1f} deg, max |d alpha_ref| {np.abs(u_applied[:, :, 1]).max() / deg:.1f} deg/s")
    # near-degenerate QPs (the oracle's own Newton-solve residual above 1e-10: 346 of 28,800 solves here) are only defined
    # to about cond * eps; helpers.parity_report widens the bound by 20 x that residual, which a handful still exceed
    assert bad <= n_rti // 5000 and worst < 1e-6
    assert iter_mismatch <= max(2, n_rti // 1000)      # termination-test flips are the exception (hard part 5)
    # the steering constraints were exercised and respected along the closed loop
    assert np.abs(traj[:, :, 6]).max() <= 30.0 * deg + 1e-9 and np.abs(u_applied[:, :, 1]).max() <= 120.0 * deg + 1e-9
    assert np.abs(u_applied[:, :, 1]).max() > 100.0 * deg
