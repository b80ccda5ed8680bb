"""SURVEY.md 8(f2) on the GPU: `nmpc_path_discretize_device` (BatchedPathDiscretizer) against the golden vectors of the
reference's own compiled discretiser, against the oracle restatement on fresh cases, through size-independent
properties at the bench batch size, and chained into the controller tick without a host hop."""
import ctypes as C
import math
import os

import numpy as np
import pytest
import torch

import pathcases
from helpers import ATOL, RTOL
from nmpc_nav_control_b200 import paths as P
from nmpc_nav_control_b200.problem import MODELS
from oracle import pathdisc
from oracle.ctrl import OracleController

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pathdisc.npz")


def _close(got, want):
    return (np.abs(got - want) <= ATOL + RTOL * np.abs(want)).all()


@pytest.mark.parametrize("hol,period,num", [(0, 0.025, 81), (0, 1.0, 12), (1, 0.025, 81), (1, 1.0, 12)])
def test_matches_reference_golden_vectors(hol, period, num):
    g = np.load(GOLD)
    off = g["offsets"]
    ps = P.PathSet([g["segments"][off[i]:off[i + 1]] for i in range(len(off) - 1)])
    d = P.BatchedPathDiscretizer(period, num, bool(hol))
    out = d.get_next_n_poses(ps, torch.from_numpy(g["path_id"]).cuda(), torch.from_numpy(g["u0"]).cuda())
    got = out.permute(2, 0, 1).cpu().numpy()
    want = g[f"poses_h{hol}_T{period}_n{num}"]
    assert _close(got, want), np.abs(got - want).max()


def test_matches_oracle_on_fresh_cases_and_edges():
    paths, pid, u0 = pathcases.cases(seed=123, n_paths=40, B=768)
    u0[:4] = [0.0, -2.0, 50.0, len(paths[pid[3]]) - 1e-9]           # path start, before it, past its end, a hair before the end
    ps = P.PathSet(paths)
    for hol in (False, True):
        d = P.BatchedPathDiscretizer(0.025, 81, hol)
        got = d.get_next_n_poses(ps, torch.from_numpy(pid).cuda(), torch.from_numpy(u0).cuda()).permute(2, 0, 1).cpu().numpy()
        worst = 0.0
        assert np.isfinite(got).all()
        end = pathdisc._pose([pathdisc.Seg(r) for r in paths[pid[2]]], float(len(paths[pid[2]])), hol)
        assert np.allclose(got[2], np.tile(end, (81, 1)), atol=1e-12)         # a start past the end: the end pose throughout
        for i in range(len(pid)):
            if i in (1, 2):
                continue             # a start outside the path indexes out of range upstream (PathDiscretizer.cpp:23); here: clamped
            want = pathdisc.get_next_n_poses(paths[pid[i]], u0[i], 0.025, 81, hol)
            assert _close(got[i], want), (i, hol, np.abs(got[i] - want).max())
            worst = max(worst, np.abs(got[i] - want).max())
        print(f"path discretiser holonomic={hol}: 768 robots x 81 poses, worst |diff| {worst:.2e}")


def test_full_size_properties():
    """65,536 robots on 512 paths: spacing, padding and determinism (no oracle at this size)"""
    B = 65536
    paths, pid, _ = pathcases.cases(seed=9, n_paths=512, B=B)
    rng = np.random.default_rng(1)
    nseg = np.array([len(paths[p]) for p in pid])
    u0 = rng.uniform(0, 1, B) * nseg
    ps = P.PathSet(paths)
    d = P.BatchedPathDiscretizer(0.025, 81, False)
    tp, tu = torch.from_numpy(pid).cuda(), torch.from_numpy(u0).cuda()
    a = d.get_next_n_poses(ps, tp, tu)
    b = d.get_next_n_poses(ps, tp, tu)
    assert torch.equal(a, b) and torch.isfinite(a).all()
    step = torch.hypot(a[1:, 0] - a[:-1, 0], a[1:, 1] - a[:-1, 1])               # [80, B]
    vmax = torch.from_numpy(np.array([np.abs(paths[p][:, 1]).max() for p in pid])).cuda()
    # one period of travel within the 1 % threshold plus one sub-step; the sub-step is sized with the tangent at the
    # previous sample, so it overshoots where the parameter speed jumps between segments (rare, bounded)
    ratio = step / (vmax * 0.025)
    assert (ratio > 1.11).double().mean() < 1e-3 and ratio.max() < 3.0
    # once a robot's list is padded (zero spacing) it stays at the path's end pose
    ends = torch.from_numpy(np.array([pathdisc._pose([pathdisc.Seg(r) for r in paths[p]], float(len(paths[p])), False) for p in pid[:256]])).cuda()
    padded = step[-1, :256] == 0
    assert padded.any() and torch.allclose(a[-1, :, :256].t()[padded], ends[padded], atol=1e-12)


def test_discretiser_feeds_the_controller_tick_on_device(oracle_mod):
    """path -> reference poses -> controller tick, device-resident end to end, against the same chain of oracles"""
    from nmpc_nav_control_b200.controller import BatchedNavController
    name, B = "diff", 12
    spec = MODELS[name]
    paths, pid, u0 = pathcases.cases(seed=31, n_paths=5, B=B)
    for p in paths:
        p[:, 1] = np.abs(p[:, 1])                                                # forward driving
    ps = P.PathSet(paths)
    d = P.BatchedPathDiscretizer(spec.dt, spec.n + 1, False)
    ctl = BatchedNavController(name, B, dt=spec.dt); ctl.reset_mpc()
    refc = [OracleController(oracle_mod, name) for _ in range(B)]
    start = np.array([pathdisc._pose([pathdisc.Seg(r) for r in paths[p]], u, False) for p, u in zip(pid, u0)])
    pose = start + np.array([0.02, -0.03, 0.05])
    vel = np.tile([0.2, 0.0, 0.0], (B, 1))
    for t in range(3):
        u_t = u0 + 0.01 * t
        refs = d.get_next_n_poses(ps, torch.from_numpy(pid).cuda(), torch.from_numpy(u_t).cuda())
        out = ctl.run(torch.from_numpy(pose.T.copy()).cuda(), torch.from_numpy(vel.T.copy()).cuda(), refs)
        cmd = out["cmd"].cpu().numpy().T
        assert (out["status"].cpu().numpy() == 0).all()
        for i in range(B):
            r = pathdisc.get_next_n_poses(paths[pid[i]], u_t[i], spec.dt, spec.n + 1, False)
            want, _ = refc[i].run(pose[i], vel[i], 0.0, [tuple(q) for q in r])
            assert _close(cmd[i], np.array(want)), (t, i, cmd[i], want)
    ctl.close()


def test_argument_errors():
    from nmpc_nav_control_b200 import _lib
    lib = _lib.load()
    z = torch.zeros(1024, dtype=torch.float64, device="cuda"); zi = torch.zeros(16, dtype=torch.int32, device="cuda")
    p, pi = C.c_void_p(z.data_ptr()), C.c_void_p(zi.data_ptr())
    f = lib.nmpc_path_discretize_device
    assert f(0, 4, None, pi, 1, pi, p, 0.025, 81, 0, p, None) == -1
    assert f(0, 0, p, pi, 1, pi, p, 0.025, 81, 0, p, None) == -1
    assert f(0, 4, p, pi, 1, pi, p, 0.0, 81, 0, p, None) == -1
    assert f(0, 4, p, pi, 1, pi, p, 0.025, 0, 0, p, None) == -1
    assert f(99, 4, p, pi, 1, pi, p, 0.025, 81, 0, p, None) == -1
