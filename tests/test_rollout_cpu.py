"""SURVEY.md 8(f3), CPU part: the plant step and the nearest-path-parameter search of csrc/rollout.cuh (compiled for the
host by tests/host_emul) against oracle/rollout.py and the solver oracle's RK4 map."""
import math

import numpy as np
import pytest

import emul
import pathcases
from nmpc_nav_control_b200.problem import MODELS
from oracle import pathdisc, rollout as orollout


@pytest.mark.parametrize("name", ["diff", "omni4", "tric"])
def test_plant_step_matches_oracle_map(oracle_mod, name):
    spec = MODELS[name]
    o = oracle_mod.Oracle(name, spec.codegen_defaults())
    rng = np.random.default_rng(2)
    B = 40
    x = rng.uniform(-1, 1, (spec.nx, B)); x[2] = rng.uniform(-math.pi, math.pi, B)
    if name == "tric":
        x[4] = rng.uniform(-0.5, 0.5, B); x[6] = rng.uniform(-0.5, 0.5, B)
    u0 = rng.uniform(-1, 1, (spec.nu, B)); nz = 0.05 * rng.standard_normal((spec.nu, B))
    p = np.array(spec.p)
    xp, pose, vel, steer = emul.emul_plant_step(name, x, u0, nz, p, spec.dt)
    for i in range(B):
        xn, _ = o.discrete_map(x[:, i], u0[:, i] + nz[:, i], p, spec.dt)
        assert np.abs(xp[:, i] - xn).max() <= 1e-13, (i, np.abs(xp[:, i] - xn).max())
        ps, vl, stw = orollout.measurements(name, spec, xp[:, i])
        assert np.array_equal(pose[:, i], ps)
        assert vel[0, i] == vl[0] and vel[1, i] == vl[1]
        if vl[2] is not None:
            assert vel[2, i] == vl[2]
        if name == "tric":
            assert steer[i] == stw
    # without noise
    xp2, *_ = emul.emul_plant_step(name, x, u0, None, p, spec.dt)
    xn, _ = o.discrete_map(x[:, 0], u0[:, 0], p, spec.dt)
    assert np.abs(xp2[:, 0] - xn).max() <= 1e-13


def test_nearest_u_matches_restatement_and_is_a_local_minimum():
    paths, pid, u0 = pathcases.cases(seed=12, n_paths=10, B=120)
    rng = np.random.default_rng(0)
    for p, u in zip(pid, u0):
        seg = paths[p]
        path = [pathdisc.Seg(r) for r in seg]
        x, y, _ = pathdisc._pose(path, min(u + 0.07, len(seg)), False)
        px, py = x + rng.uniform(-0.05, 0.05), y + rng.uniform(-0.05, 0.05)
        got = emul.emul_nearest_u(seg, u, px, py, 0.05, 0.5)
        want = orollout.nearest_u(seg, u, px, py, 0.05, 0.5)
        assert abs(got - want) <= 1e-12, (got, want)
        d = lambda su: math.hypot(pathdisc._pose(path, su, False)[0] - px, pathdisc._pose(path, su, False)[1] - py)
        lo, hi = max(u - 0.05, 0.0), min(u + 0.5, len(seg))
        grid = np.linspace(lo, hi, 400)
        assert d(got) <= min(d(g) for g in grid) + 2e-3          # no grid point is noticeably closer
    # degenerate windows
    ln = paths[0]
    assert emul.emul_nearest_u(ln, float(len(ln)), 0.0, 0.0, 0.0, 0.0) == float(len(ln))
    assert emul.emul_nearest_u(ln, -5.0, 0.0, 0.0, 0.0, 1.0) == 0.0
