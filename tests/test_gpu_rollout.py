"""SURVEY.md 8(f3) on the GPU: the device-resident closed loop (nmpc_nav_control_b200/rollout.py: nearest path parameter
-> path discretiser -> controller tick -> plant step, no host hop between ticks) against the same loop of oracles
(oracle/rollout.py), free running, with seeded acceleration noise in the plant."""
import numpy as np
import pytest
import torch

import pathcases
from helpers import ATOL, RTOL
from nmpc_nav_control_b200.problem import MODELS
from oracle import pathdisc
from oracle.rollout import OracleRollout

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["diff", "omni4", "tric"])
def test_closed_loop_matches_oracle_chain(oracle_mod, name):
    from nmpc_nav_control_b200.controller import BatchedNavController
    from nmpc_nav_control_b200.paths import PathSet
    from nmpc_nav_control_b200.rollout import ClosedLoopRollout
    spec = MODELS[name]
    B, T = 10, 12
    paths, pid, u0 = pathcases.cases(seed=41, n_paths=4, B=B)
    for p in paths:
        p[:, 1] = np.clip(np.abs(p[:, 1]), 0.2, 0.6)                       # forward, within the actuator limits
    u0 = np.minimum(u0, np.array([len(paths[p]) for p in pid]) - 0.5)
    rng = np.random.default_rng(6)
    start = np.array([pathdisc._pose([pathdisc.Seg(r) for r in paths[p]], u, False) for p, u in zip(pid, u0)])
    pose0 = start + rng.uniform(-0.03, 0.03, (B, 3))
    noise = 0.05 * rng.standard_normal((T, spec.nu, B))                    # acados_sim_diff.py:151-152: N(0, 0.05) on the accelerations
    ctl = BatchedNavController(name, B, dt=spec.dt)
    ro = ClosedLoopRollout(ctl, PathSet(paths), torch.from_numpy(pid).cuda())
    ro.reset(torch.from_numpy(pose0.T.copy()).cuda(), torch.from_numpy(u0).cuda())
    res = ro.run(T, torch.from_numpy(noise).cuda())
    traj = res["pose"].cpu().numpy(); cmds = res["cmd"].cpu().numpy()
    assert int(res["failed"].sum()) == 0
    worst = 0.0
    for i in range(B):
        o = OracleRollout(oracle_mod, name, paths[pid[i]], pose0[i], u0[i])
        for t in range(T):
            cmd, _ = o.step(noise[t, :, i])
            want = np.array(cmd)
            e1 = np.abs(cmds[t, :, i] - want); e2 = np.abs(traj[t + 1, :, i] - o.pose)
            assert (e1 <= ATOL + RTOL * np.abs(want)).all() and (e2 <= ATOL + RTOL * np.abs(o.pose)).all(), (name, i, t, e1, e2)
            worst = max(worst, e1.max(), e2.max())
    # the robots moved along their paths (not tric: with the model's sin-for-cos defect, scripts/tric/tric_amr_model.py:45,
    # the pose rate is v sin(alpha) and a robot steering straight ahead stands still)
    assert name == "tric" or np.hypot(traj[-1, 0] - traj[0, 0], traj[-1, 1] - traj[0, 1]).max() > 0.005
    print(f"closed loop {name}: {T} ticks x {B} robots, worst |diff| {worst:.2e}")
    ctl.close()


def test_large_rollout_stays_on_device_and_is_repeatable():
    from nmpc_nav_control_b200.controller import BatchedNavController
    from nmpc_nav_control_b200.paths import PathSet
    from nmpc_nav_control_b200.rollout import ClosedLoopRollout
    name, B, T = "diff", 4096, 5
    spec = MODELS[name]
    paths, pid, u0 = pathcases.cases(seed=3, n_paths=64, B=B)
    for p in paths:
        p[:, 1] = np.clip(np.abs(p[:, 1]), 0.2, 0.6)
    start = np.array([pathdisc._pose([pathdisc.Seg(r) for r in paths[p]], u, False) for p, u in zip(pid, u0)])
    ctl = BatchedNavController(name, B, dt=spec.dt)
    ro = ClosedLoopRollout(ctl, PathSet(paths), torch.from_numpy(pid).cuda())
    outs = []
    for _ in range(2):
        ro.reset(torch.from_numpy(start.T.copy()).cuda(), torch.from_numpy(u0).cuda())
        r = ro.run(T)
        outs.append((r["pose"].clone(), r["cmd"].clone()))
        assert int(r["failed"].sum()) == 0 and torch.isfinite(r["pose"]).all()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    v = outs[0][1][-1, 0]
    assert (v.abs() <= 1.0 + 1e-9).all() and v.abs().max() > 0.05                   # commands respect v_max and are not idle
    ctl.close()
