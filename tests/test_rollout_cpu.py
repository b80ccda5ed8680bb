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


@pytest.mark.parametrize("name", ["diff", "tric"])
def test_whole_tick_chain_on_the_host_matches_oracle_chain(oracle_mod, name):
    """every device function of one closed-loop tick, compiled for the host and chained exactly as
    nmpc_nav_control_b200/rollout.py chains the kernels (nearest parameter -> path discretiser -> controller glue ->
    RTI step -> glue -> plant step), against oracle/rollout.py, free running for a few ticks"""
    spec = MODELS[name]
    B, T = 6, 4
    paths, pid, u0 = pathcases.cases(seed=41, n_paths=3, B=B)
    for p in paths:
        p[:, 1] = np.clip(np.abs(p[:, 1]), 0.2, 0.6)
    u0 = np.minimum(u0, np.array([len(paths[p]) for p in pid]) - 0.5)
    off = np.cumsum([0] + [len(p) for p in paths]).astype(np.int32)
    segs = np.concatenate(paths)
    rng = np.random.default_rng(6)
    start = np.array([pathdisc._pose([pathdisc.Seg(r) for r in paths[p]], u, False) for p, u in zip(pid, u0)])
    pose0 = start + rng.uniform(-0.03, 0.03, (B, 3))
    noise = 0.05 * rng.standard_normal((T, spec.nu, B))
    ctl = orollout.OracleController(None, name)                 # for its tables only (W_e = Q, as the wrappers set it)
    tb = ctl.tb
    # device-side state, SoA
    x_plant = np.zeros((spec.nx, B)); x_plant[:3] = pose0.T
    pose = pose0.T.copy(); vel = np.zeros((3, B)); steer = np.zeros(B); u = u0.copy()
    vref = np.zeros((spec.nv, B)); cmd = np.zeros((3, B))
    xs = np.zeros((B, spec.n + 1, spec.nx)); us = np.zeros((B, spec.n, spec.nu))          # after reset_mpc()
    oracles = [orollout.OracleRollout(oracle_mod, name, paths[pid[i]], pose0[i], u0[i]) for i in range(B)]
    for t in range(T):
        for i in range(B):
            a, b = off[pid[i]], off[pid[i] + 1]
            u[i] = emul.emul_nearest_u(segs[a:b], u[i], pose[0, i], pose[1, i], 0.05, 0.5)
        refs = emul.emul_path_discretize(segs, off, pid, u, spec.dt, spec.n + 1)
        x0bar, yref, We = emul.emul_ctrl_pre(name, pose, vel, steer if name == "tric" else None, refs, None, vref, tb)
        r = emul.emul_rti(name, np.ascontiguousarray(x0bar.T), np.ascontiguousarray(np.moveaxis(yref, -1, 0)), xs, us,
                          We=None if We is None else np.ascontiguousarray(We.T), tables=tb)
        assert (r["qp_status"] == 0).all()
        xs, us = r["x"], r["u"]
        u_first = np.ascontiguousarray(us[:, 0, :].T)
        vref, cmd = emul.emul_ctrl_post(name, np.zeros(B, np.int32), x0bar, u_first, spec.dt, vref, cmd, tb)
        x_plant, pose, vel, steer = emul.emul_plant_step(name, x_plant, u_first, noise[t], np.array(spec.p), spec.dt)
        for i in range(B):
            want, qi = oracles[i].step(noise[t, :, i])
            assert r["qp_iter"][i] == qi, (t, i)
            assert np.abs(cmd[:, i] - np.array(want)).max() <= 1e-9, (t, i, cmd[:, i], want)
            assert np.abs(pose[:, i] - oracles[i].pose).max() <= 1e-9, (t, i)


def test_oracle_sqp_and_shift_protocol(oracle_mod):
    """the SQP / warm-start-shift extension of the oracle chain (the parity target of nmpc_rollout_device with
    sqp_max_iter > 1 and shift = 1): SQP stops on the step norm, the shift moves every stage down by one and keeps the last"""
    from oracle.ctrl import OracleController
    c = OracleController(oracle_mod, "tric")
    refs = [(0.02 * k, 0.01 * k, 0.05) for k in range(c.spec.n + 1)]
    cmd1, qp1 = c.run(np.array([0.0, 0.02, 0.0]), (0.1, 0.0, 0.0), 0.05, refs, sqp_max_iter=6, sqp_tol=1e-8)
    assert 1 < c.sqp_steps <= 6 and qp1 > 0
    x, u = c.x.copy(), c.u.copy()
    c.shift()
    assert np.array_equal(c.x[:-1], x[1:]) and np.array_equal(c.x[-1], x[-1])
    assert np.array_equal(c.u[:-1], u[1:]) and np.array_equal(c.u[-1], u[-1])
    c2 = OracleController(oracle_mod, "tric")
    c2.run(np.array([0.0, 0.02, 0.0]), (0.1, 0.0, 0.0), 0.05, refs)          # the reference's single RTI step
    assert c2.sqp_steps == 1 and np.abs(c2.u - u).max() > 1e-9
