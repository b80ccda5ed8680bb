"""SURVEY.md 8(f1), CPU part: the controller glue of nmpc_nav_control_b200/csrc/ctrl_glue.cuh (compiled for the host by
tests/host_emul) against the wrappers' protocol restated in oracle/ctrl.py - initial state, heading unwrap chain,
reference padding, the diff terminal-weight switch, reference-state integration, inverse kinematics.  Pure data
movement and a handful of additions: the comparison is bit for bit (the emulation is compiled with -ffp-contract=off)."""
import math

import numpy as np
import pytest

import emul
from nmpc_nav_control_b200.problem import MODELS
from oracle.ctrl import OracleController


def _batch(name, B, seed, nref_max=81):
    rng = np.random.default_rng(seed)
    spec = MODELS[name]
    pose = np.stack([rng.uniform(-2, 2, B), rng.uniform(-2, 2, B), rng.uniform(-math.pi, math.pi, B)])
    vel = np.stack([rng.uniform(-1, 1, B), rng.uniform(-0.3, 0.3, B), rng.uniform(-1, 1, B)])
    steer = rng.uniform(-0.5, 0.5, B)
    # reference headings wrapped to (-pi, pi] and crossing the cut on many instances; some lists short, some of length 1,
    # some with the last two poses identical (the terminal-weight switch)
    s = np.arange(nref_max)[:, None] * 0.02
    kap = rng.uniform(-3, 3, B)
    th = pose[2] + rng.uniform(-1, 1, B) + kap * s
    th = np.arctan2(np.sin(th), np.cos(th))
    refs = np.stack([pose[0] + s * np.cos(pose[2]), pose[1] + s * np.sin(pose[2]), th], axis=1)     # [nref_max, 3, B]
    nref = rng.choice([1, 2, 37, 60, 80, nref_max, nref_max + 7], size=B).astype(np.int32)
    dup = rng.random(B) < 0.3
    refs[-1][:, dup] = refs[-2][:, dup]
    vref = rng.uniform(-1, 1, (spec.nv, B))
    return pose, vel, steer, refs, nref, vref


@pytest.mark.parametrize("name", ["diff", "omni4", "tric"])
def test_pre_matches_wrapper_protocol(name):
    spec = MODELS[name]
    B = 96
    pose, vel, steer, refs, nref, vref = _batch(name, B, 5)
    ctl = OracleController(None, name)
    x0bar, yref, We = emul.emul_ctrl_pre(name, pose, vel, steer if name == "tric" else None, refs, nref, vref, ctl.tb)
    n_switch = 0
    for i in range(B):
        ctl.x0[:] = 0.0
        ctl.x0[3 + spec.nv:] = vref[:, i]
        n = min(int(nref[i]), refs.shape[0])
        x0, yr, we = ctl.pre(pose[:, i], vel[:, i], steer[i], [tuple(refs[k, :, i]) for k in range(n)])
        assert np.array_equal(x0bar[:, i], x0), (i, x0bar[:, i], x0)
        assert np.array_equal(yref[:, :, i], yr[:, :3]), i
        if name == "diff":
            assert np.array_equal(We[:, i], we), (i, We[:, i], we)
            n_switch += int(we[0] == 100.0 * spec.Q[0])
        else:
            assert We is None and we is None
    if name == "diff":
        assert 0 < n_switch < B          # both branches of the switch were taken


@pytest.mark.parametrize("name", ["diff", "omni4", "tric"])
def test_nref_none_and_unwrap_chain(name):
    """nref = NULL means every list is full; the unwrapped headings never jump by more than pi between rows"""
    B = 32
    pose, vel, steer, refs, _, vref = _batch(name, B, 9)
    ctl = OracleController(None, name)
    _, yref, _ = emul.emul_ctrl_pre(name, pose, vel, steer if name == "tric" else None, refs, None, vref, ctl.tb)
    chain = np.concatenate([pose[2][None], yref[:, 2, :]], axis=0)
    assert np.abs(np.diff(chain, axis=0)).max() <= math.pi
    assert np.allclose(np.cos(yref[:, 2, :]), np.cos(refs[:, 2, :]), atol=1e-12)     # same angle modulo 2 pi


@pytest.mark.parametrize("name", ["diff", "omni4", "tric"])
def test_post_matches_wrapper_protocol(name):
    spec = MODELS[name]
    B = 64
    rng = np.random.default_rng(3)
    x0bar = rng.uniform(-1, 1, (spec.nx, B)); u0 = rng.uniform(-2, 2, (spec.nu, B))
    status = (rng.random(B) < 0.2).astype(np.int32) * 2              # some failed solves
    vref0 = x0bar[3 + spec.nv:].copy(); cmd0 = rng.uniform(-1, 1, (3, B))
    dt = 0.02
    ctl = OracleController(None, name, dt=dt)
    vref, cmd = emul.emul_ctrl_post(name, status, x0bar, u0, dt, vref0, cmd0, ctl.tb)
    for i in range(B):
        if status[i]:
            assert np.array_equal(vref[:, i], vref0[:, i]) and np.array_equal(cmd[:, i], cmd0[:, i])
            continue
        c, nr = ctl.post(x0bar[:, i], u0[:, i])
        assert np.array_equal(vref[:, i], nr), i
        assert np.array_equal(cmd[:, i], np.array(c)), (i, cmd[:, i], c)
