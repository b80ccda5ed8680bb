"""SURVEY.md 8(f1) on the GPU: the batched controller tick (`BatchedNavController.run`, nmpc_ctrl_tick_device /
nmpc_ctrl_tick_host) against the wrappers' protocol restated around the oracle (oracle/ctrl.py, itself pinned
against the reference's unmodified wrapper sources by tests/test_acados_dropin.py).

Every robot follows its own path at its own speed; the measured pose advances along the path, so consecutive
ticks are warm-started RTI steps exactly as in NMPCNavControlROS::executeNMPC.  Reference lists are ragged (a short
list every third tick, which also flips the diff terminal-weight switch)."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from helpers import ATOL, RTOL
from nmpc_nav_control_b200.problem import MODELS
from oracle.ctrl import OracleController

pytestmark = pytest.mark.gpu
NREF = 81


def _scenarios(name, B, T, seed):
    """per tick: pose [B,3], vel [B,3], steer [B], refs [B,NREF,3] (wrapped headings), nref [B]"""
    rng = np.random.default_rng(seed)
    pose = np.stack([rng.uniform(-1, 1, B), rng.uniform(-1, 1, B), rng.uniform(-3.1, 3.1, B)], axis=1)
    speed = rng.uniform(0.3, 0.7, B); kap = rng.uniform(-1.0, 1.0, B)
    ticks = []
    i = np.arange(NREF)[None, :]
    for t in range(T):
        ramp = min(1.0, 0.1 * t)
        vel = np.stack([speed * ramp, np.full(B, 0.05 if name == "omni4" else 0.0), 0.1 * kap], axis=1)
        steer = 0.05 * kap
        s = speed[:, None] * 0.025 * i
        th = pose[:, 2:3] + kap[:, None] * s
        th = np.arctan2(np.sin(th), np.cos(th))
        refs = np.stack([pose[:, 0:1] + s * np.cos(pose[:, 2:3] + 0.5 * kap[:, None] * s),
                         pose[:, 1:2] + s * np.sin(pose[:, 2:3] + 0.5 * kap[:, None] * s), th], axis=2)
        nref = np.where((t + np.arange(B)) % 3 == 0, 60, NREF).astype(np.int32)
        ticks.append((pose.copy(), vel, steer, refs, nref))
        ds = speed * 0.025
        pose = np.stack([pose[:, 0] + ds * np.cos(pose[:, 2]), pose[:, 1] + ds * np.sin(pose[:, 2]), pose[:, 2] + kap * ds], axis=1)
        pose[:, 2] = np.arctan2(np.sin(pose[:, 2]), np.cos(pose[:, 2]))
    return ticks


def _soa(a):
    """instance-major host array -> SoA CUDA tensor (instance index fastest)"""
    return torch.from_numpy(np.ascontiguousarray(np.moveaxis(a, 0, -1))).cuda()


@pytest.mark.parametrize("name", ["diff", "omni4", "tric"])
def test_device_tick_matches_wrapper_protocol(oracle_mod, name):
    from nmpc_nav_control_b200.controller import BatchedNavController
    spec = MODELS[name]
    B, T = 24, 10
    ticks = _scenarios(name, B, T, seed=21)
    ctl = BatchedNavController(name, B, dt=spec.dt)
    assert ctl.get_horizon() == 80 and ctl.get_delta_time() == spec.dt
    ctl.reset_mpc()                  # a new goal / path resets the iterate before the first tick (NMPCNavControlROS.cpp:309-326)
    refc = [OracleController(oracle_mod, name) for _ in range(B)]
    worst = 0.0
    out = None
    for t, (pose, vel, steer, refs, nref) in enumerate(ticks):
        if name == "tric":
            ctl.set_steering_wheel_angle(torch.from_numpy(steer).cuda())
        out = ctl.run(_soa(pose), _soa(vel), _soa(refs), torch.from_numpy(nref).cuda(), out=out)
        cmd = out["cmd"].cpu().numpy().T; st = out["status"].cpu().numpy(); it = out["qp_iter"].cpu().numpy()
        assert (st == 0).all(), (t, st)
        for i in range(B):
            want, qi = refc[i].run(pose[i], vel[i], steer[i], [tuple(r) for r in refs[i, :nref[i]]])
            assert it[i] == qi, (t, i, it[i], qi)
            err = np.abs(cmd[i] - np.array(want))
            assert (err <= ATOL + RTOL * np.abs(want)).all(), (t, i, cmd[i], want)
            worst = max(worst, err.max())
        # the carried reference states are the oracle's
        vref = ctl.reference_states()[:, :B].cpu().numpy().T
        want_v = np.stack([c.x0[3 + spec.nv:] for c in refc])
        assert np.abs(vref - want_v).max() <= ATOL + RTOL * np.abs(want_v).max()
    print(f"controller tick {name}: {T} ticks x {B} robots, worst |cmd diff| {worst:.2e}")
    ctl.close()


@pytest.mark.parametrize("name", ["diff", "tric"])
def test_host_tick_equals_device_tick(name):
    from nmpc_nav_control_b200.controller import BatchedNavController
    spec = MODELS[name]
    B, T = 40, 4
    ticks = _scenarios(name, B, T, seed=4)
    a = BatchedNavController(name, B, dt=spec.dt); b = BatchedNavController(name, B, dt=spec.dt)
    a.reset_mpc(); b.reset_mpc()
    for pose, vel, steer, refs, nref in ticks:
        if name == "tric":
            a.set_steering_wheel_angle(steer); b.set_steering_wheel_angle(steer)
        oa = a.run(_soa(pose), _soa(vel), _soa(refs), torch.from_numpy(nref).cuda())
        ob = b.run_host(pose, vel, refs, nref)
        assert np.array_equal(oa["cmd"].cpu().numpy().T, ob["cmd"])
        assert np.array_equal(oa["qp_iter"].cpu().numpy(), ob["qp_iter"]) and (ob["status"] == 0).all()
    a.close(); b.close()


def test_failed_robot_keeps_command_and_state_and_reset():
    from nmpc_nav_control_b200.controller import BatchedNavController
    name, B = "diff", 16
    spec = MODELS[name]
    pose, vel, steer, refs, nref = _scenarios(name, B, 1, seed=8)[0]
    ctl = BatchedNavController(name, B, dt=spec.dt)
    ctl.reset_mpc()
    o1 = ctl.run_host(pose, vel, refs, nref)
    v1 = ctl.reference_states()[:, :B].cpu().numpy().copy()
    assert (o1["status"] == 0).all() and np.abs(v1).max() > 0
    bad = pose.copy(); bad[5, 0] = np.nan
    cmd_in = o1["cmd"].copy()
    o2 = ctl.run_host(bad, vel, refs, nref, out=dict(cmd=cmd_in, status=np.empty(B, np.int32), qp_iter=np.empty(B, np.int32)))
    v2 = ctl.reference_states()[:, :B].cpu().numpy()
    assert o2["status"][5] != 0 and (np.delete(o2["status"], 5) == 0).all()
    assert np.array_equal(o2["cmd"][5], o1["cmd"][5]) and np.array_equal(v2[:, 5], v1[:, 5])       # untouched
    assert np.isfinite(o2["cmd"]).all() and not np.array_equal(np.delete(v2, 5, axis=1), np.delete(v1, 5, axis=1))
    # constructor state again: zero carried states, zero iterate -> the first tick repeats bit for bit
    from nmpc_nav_control_b200 import _lib
    _lib.check(ctl.lib.nmpc_ctrl_reset(ctl.solver._h, None), "nmpc_ctrl_reset"); ctl.reset_mpc()
    o3 = ctl.run_host(pose, vel, refs, nref)
    assert np.array_equal(o3["cmd"], o1["cmd"])
    ctl.close()


def test_tick_argument_errors():
    from nmpc_nav_control_b200 import _lib
    from nmpc_nav_control_b200.controller import BatchedNavController
    ctl = BatchedNavController("tric", 8, dt=0.025)
    lib, h = ctl.lib, ctl.solver._h
    z = torch.zeros(81 * 3 * 8, dtype=torch.float64, device="cuda")
    p = C.c_void_p(z.data_ptr())
    assert lib.nmpc_ctrl_tick_device(h, 8, p, p, None, p, None, 81, 0.025, p, None, None, None) == -1     # tric without steering angle
    assert b"steering" in lib.nmpc_last_error()
    assert lib.nmpc_ctrl_tick_device(h, 8, p, p, p, p, None, 0, 0.025, p, None, None, None) == -1        # no reference pose
    assert lib.nmpc_ctrl_tick_device(h, 9, p, p, p, p, None, 81, 0.025, p, None, None, None) == -4       # capacity
    assert lib.nmpc_ctrl_tick_device(h, 8, p, p, p, p, None, 81, 0.0, p, None, None, None) == -1         # dt
    assert lib.nmpc_ctrl_tick_device(h, 8, None, p, p, p, None, 81, 0.025, p, None, None, None) == -1
    hz = np.zeros((8, 100, 3))
    assert lib.nmpc_ctrl_tick_host(h, 8, C.c_void_p(hz.ctypes.data), C.c_void_p(hz.ctypes.data), C.c_void_p(hz.ctypes.data),
                                   C.c_void_p(hz.ctypes.data), None, 100, 0.025, C.c_void_p(hz.ctypes.data),
                                   C.c_void_p(hz.ctypes.data), None) == -1                               # nref_max > N+1
    ctl.close()
