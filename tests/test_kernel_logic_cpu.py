"""CPU check of the CUDA solver's per-lane arithmetic: rti_core.cuh compiled by g++ as a lane-by-lane
emulation (tests/host_emul) against the oracle.  The structured Riccati / delta-corrector path of
the kernels is a different factorisation of the same Newton systems as the oracle's dense
square-root Riccati, so agreement is to rounding, not bit-exact; iteration counts must be equal."""
import numpy as np
import pytest

import emul
from helpers import instances, oracle_solve, parity_report


@pytest.mark.parametrize("name,B,start", [("diff", 192, 0), ("omni4", 64, 300), ("tric", 128, 77)])
def test_cold_step_matches_oracle(oracle_mod, name, B, start):
    spec, x0, yref, _ = instances(name, start, B)
    ref = oracle_solve(oracle_mod, name, x0, yref)
    out = emul.emul_rti(name, x0, yref)
    assert (out["qp_iter"] == ref["qp_iter"]).all()
    assert parity_report(out["x"], ref["x"])[0] == 0
    assert parity_report(out["u"], ref["u"])[0] == 0


@pytest.mark.parametrize("name", ["diff", "omni4", "tric"])
def test_warm_steps_and_pose_only_yref(oracle_mod, name):
    """three consecutive RTI steps (iterate carried over, x0 <- x1), pose-only yref in the kernel path"""
    B = 48
    spec, x0, yref3, _ = instances(name, 9000, B, pose_only=True)
    yfull = np.zeros((B, spec.n + 1, spec.ny)); yfull[:, :, :3] = yref3
    xr = ur = xe = ue = None
    x0r, x0e = x0.copy(), x0.copy()
    for step in range(3):
        ref = oracle_solve(oracle_mod, name, x0r, yfull, x=xr, u=ur)
        out = emul.emul_rti(name, x0e, yref3, x=xe, u=ue)
        assert (out["qp_iter"] == ref["qp_iter"]).all(), step
        lr = ref["lin_res"]
        assert (lr > 1e-10).sum() <= 2, step          # ill-conditioned QPs are the exception
        assert parity_report(out["x"], ref["x"], lr)[0] == 0 and parity_report(out["u"], ref["u"], lr)[0] == 0, step
        xr, ur, xe, ue = ref["x"], ref["u"], out["x"], out["u"]
        x0r, x0e = xr[:, 1].copy(), xe[:, 1].copy()


def test_diff_per_instance_terminal_weight(oracle_mod):
    B = 64
    spec, x0, yref, We = instances("diff", 100, B, terminal_hack=True)
    ref = oracle_solve(oracle_mod, "diff", x0, yref, We=We)
    out = emul.emul_rti("diff", x0, yref, We=We)
    assert (out["qp_iter"] == ref["qp_iter"]).all()
    assert parity_report(out["u"], ref["u"])[0] == 0


def test_nondefault_tables(oracle_mod):
    """stage-varying weights / bounds / parameters set through the table setters"""
    rng = np.random.default_rng(5)
    for name in ("diff", "tric", "omni4"):
        B = 32
        spec, x0, yref, _ = instances(name, 4000, B)
        tb = spec.codegen_defaults()
        tb["W"] = tb["W"] * (1.0 + 0.3 * rng.random(tb["W"].shape))
        tb["W"][:, 3:spec.nx] = 0.05 * rng.random((spec.n, spec.nx - 3))     # weights on the velocity states too
        tb["We"] = tb["We"] * 0.5 + 0.1
        tb["lbu"] = tb["lbu"] * (1.0 - 0.2 * rng.random(tb["lbu"].shape))
        tb["ubx"] = tb["ubx"] * (1.0 + 0.1 * rng.random(tb["ubx"].shape))
        tb["p"] = tb["p"] * (1.0 + 0.05 * rng.random(tb["p"].shape))
        yref[:, :, 3:] = 0.1 * rng.standard_normal(yref[:, :, 3:].shape)       # full-width yref
        ref = oracle_solve(oracle_mod, name, x0, yref, tables=tb)
        out = emul.emul_rti(name, x0, yref, tables=tb)
        assert (out["qp_iter"] == ref["qp_iter"]).all(), name
        assert parity_report(out["x"], ref["x"])[0] == 0 and parity_report(out["u"], ref["u"])[0] == 0, name
