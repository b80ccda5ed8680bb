"""BASELINE config 4 / north-star kernel (4) on the GPU: warm-start shift, SQP to convergence with a device-side
convergence mask, and the closed-loop rollout engine (`nmpc_rollout_device`: T ticks of nearest point -> reference poses ->
controller tick -> plant step -> shift enqueued by ONE call of the C ABI).  Not reference behaviours (the reference does one
RTI step per tick and never shifts: NMPCNavControlROS.cpp:309,316,326 -> NMPCNavControlDiff.cpp:142); the shapes mirrored are
scripts/test_scripts/casadi_sim_diff.py:104-106 (warm start from the previous solution) and acados_sim_diff.py:119-163
(closed loop).  Parity target: the oracle driven through the same protocol (oracle/ctrl.py, oracle/rollout.py)."""
import time

import numpy as np
import pytest
import torch

import pathcases
from helpers import ATOL, RTOL, instances, parity_report
from nmpc_nav_control_b200.problem import MODELS
from oracle import pathdisc
from oracle.rollout import OracleRollout

pytestmark = pytest.mark.gpu


def _soa(a):
    """instance-major [B, K, R] -> device SoA [K, R, B]"""
    return torch.from_numpy(np.ascontiguousarray(np.transpose(a, (1, 2, 0)))).cuda()


def test_shift_device_matches_numpy():
    from nmpc_nav_control_b200.solver import BatchedRtiSolver
    spec = MODELS["omni4"]
    B = 70
    rng = np.random.default_rng(0)
    x = rng.standard_normal((B, spec.n + 1, spec.nx)); u = rng.standard_normal((B, spec.n, spec.nu))
    s = BatchedRtiSolver(spec, 96)
    # external iterate, masked
    dx, du = _soa(x), _soa(u)
    mask = torch.from_numpy((rng.uniform(size=B) < 0.6).astype(np.int32)).cuda()
    s.shift(B, dx, du, mask)
    wx, wu = x.copy(), u.copy()
    m = mask.cpu().numpy().astype(bool)
    wx[m, :-1] = x[m, 1:]; wu[m, :-1] = u[m, 1:]
    assert np.array_equal(dx.cpu().numpy(), np.transpose(wx, (1, 2, 0))) and np.array_equal(du.cpu().numpy(), np.transpose(wu, (1, 2, 0)))
    # the persisted iterate, all instances, twice
    s.set_iterate(x, u)
    s.shift(B); s.shift(B)
    torch.cuda.synchronize()          # the shift runs on torch's stream, the host copies of the iterate on the solver's own
    gx, gu = s.get_iterate(B)
    wx, wu = x.copy(), u.copy()
    for _ in range(2):
        wx[:, :-1] = wx[:, 1:].copy(); wu[:, :-1] = wu[:, 1:].copy()
    assert np.array_equal(gx, wx) and np.array_equal(gu, wu)
    s.close()


@pytest.mark.parametrize("name,B", [("tric", 48), ("diff", 40)])
def test_sqp_to_convergence_matches_oracle_loop(oracle_mod, name, B):
    """per instance: RTI steps until the step inf-norm <= tol; the device keeps converged instances out of later passes"""
    from nmpc_nav_control_b200.solver import BatchedRtiSolver
    spec, x0, yref, _ = instances(name, 300, B)
    MAXI, TOL = 8, 1e-8
    orc = oracle_mod.Oracle(name, spec.codegen_defaults())
    xo = np.zeros((B, spec.n + 1, spec.nx)); uo = np.zeros((B, spec.n, spec.nu))
    steps = np.zeros(B, dtype=int); qps = np.zeros(B, dtype=int); last = np.zeros(B)
    for i in range(B):
        for _ in range(MAXI):
            r = orc.rti(x0[i], yref[i], xo[i], uo[i])
            assert r["status"] == 0
            last[i] = max(np.abs(r["x"] - xo[i]).max(), np.abs(r["u"] - uo[i]).max())
            xo[i], uo[i] = r["x"], r["u"]
            steps[i] += 1; qps[i] += r["qp_iter"]
            if last[i] <= TOL:
                break
    s = BatchedRtiSolver(spec, B)
    s.reset()
    out = s.sqp_solve_device(_soa(x0[:, None, :])[0].contiguous(), _soa(yref), MAXI, TOL)
    torch.cuda.synchronize()
    xg, ug = s.get_iterate(B)
    st = out["status"].cpu().numpy(); it = out["sqp_iter"].cpu().numpy(); qp = out["qp_iter"].cpu().numpy()
    assert (st == 0).all()
    # an instance whose deciding step norm sits within rounding of the tolerance may take one step more or less
    border = np.abs(last - TOL) < 1e-11
    same = it == steps
    assert (same | border).all(), (it, steps)
    assert (qp[same] == qps[same]).all()
    nbx, ex = parity_report(xg[same], xo[same]); nbu, eu = parity_report(ug[same], uo[same])
    print(f"SQP {name}: {B} instances, steps min/mean/max {steps.min()}/{steps.mean():.2f}/{steps.max()}, converged {int((last <= TOL).sum())}, "
          f"worst |diff| {max(ex, eu):.2e}, outside 1e-9: {nbx + nbu}")
    # the SQP loop is a chain of solves, each starting from the previous one's result: a rounding-level difference of one
    # step is carried (and, where the active set chatters, amplified) by the following ones, so the 1e-9 bound of a single
    # solve is asserted for (nearly) all instances and 1e-6 for every one (tests/test_gpu_closed_loop.py has the per-solve,
    # teacher-forced comparison of the same protocol)
    assert steps.max() > 1
    assert nbx + nbu <= max(1, B // 16) and max(ex, eu) < 1e-6
    s.close()


def _setup(name, B, seed, n_paths):
    from nmpc_nav_control_b200.controller import BatchedNavController
    from nmpc_nav_control_b200.paths import PathSet
    from nmpc_nav_control_b200.rollout import ClosedLoopRollout
    spec = MODELS[name]
    paths, pid, u0 = pathcases.cases(seed=seed, n_paths=n_paths, B=B)
    for p in paths:
        p[:, 1] = np.clip(np.abs(p[:, 1]), 0.2, 0.6)
    u0 = np.minimum(u0, np.array([len(paths[p]) for p in pid]) - 0.5)
    rng = np.random.default_rng(seed + 1)
    start = np.array([pathdisc._pose([pathdisc.Seg(r) for r in paths[p]], u, False) for p, u in zip(pid, u0)])
    pose0 = start + rng.uniform(-0.03, 0.03, (B, 3))
    ctl = BatchedNavController(name, B, dt=spec.dt)
    ro = ClosedLoopRollout(ctl, PathSet(paths), torch.from_numpy(pid).cuda())
    return spec, paths, pid, u0, pose0, ctl, ro, rng


@pytest.mark.parametrize("name", ["diff", "omni4", "tric"])
def test_rollout_engine_equals_the_tick_by_tick_calls(name):
    """RTI, no shift: one nmpc_rollout_device call = the same C-ABI calls issued tick by tick from Python, bit for bit"""
    B, T = 96, 6
    spec, paths, pid, u0, pose0, ctl, ro, rng = _setup(name, B, 21, 8)
    noise = torch.from_numpy(0.05 * rng.standard_normal((T, spec.nu, B))).cuda()
    p0, uu = torch.from_numpy(pose0.T.copy()).cuda(), torch.from_numpy(u0).cuda()
    ro.reset(p0, uu)
    a = ro.run(T, noise)
    a = {k: v.clone() for k, v in a.items()}
    ro.reset(p0, uu)
    b = ro.run_engine(T, noise)
    torch.cuda.synchronize()
    assert torch.equal(a["pose"], b["pose"]) and torch.equal(a["cmd"], b["cmd"])
    assert int(a["failed"].sum()) == 0 and int(b["failed"].sum()) == 0
    ctl.close()


@pytest.mark.parametrize("name", ["tric", "diff"])
def test_rollout_engine_sqp_and_shift_match_the_oracle_chain(oracle_mod, name):
    """SQP to convergence + warm-start shift inside the engine, free running, against oracle/rollout.py with the same options"""
    B, T, MAXI, TOL = 8, 8, 4, 1e-8
    spec, paths, pid, u0, pose0, ctl, ro, rng = _setup(name, B, 33, 4)
    noise = 0.05 * rng.standard_normal((T, spec.nu, B))
    ro.reset(torch.from_numpy(pose0.T.copy()).cuda(), torch.from_numpy(u0).cuda())
    res = ro.run_engine(T, torch.from_numpy(noise).cuda(), sqp_max_iter=MAXI, sqp_tol=TOL, shift=True)
    traj = res["pose"].cpu().numpy(); cmds = res["cmd"].cpu().numpy()
    assert int(res["failed"].sum()) == 0
    worst, nsteps = 0.0, 0
    for i in range(B):
        o = OracleRollout(oracle_mod, name, paths[pid[i]], pose0[i], u0[i])
        for t in range(T):
            cmd, _ = o.step(noise[t, :, i], sqp_max_iter=MAXI, sqp_tol=TOL, shift=True)
            nsteps += o.c.sqp_steps
            want = np.array(cmd)
            e1 = np.abs(cmds[t, :, i] - want); e2 = np.abs(traj[t + 1, :, i] - o.pose)
            worst = max(worst, e1.max(), e2.max())
            # the full-step SQP amplifies rounding differences along a free-running closed loop (tests/test_gpu_closed_loop.py):
            # 1e-9 on the first ticks, 1e-7 on all
            lim = 1.0 if t < 3 else 100.0
            assert (e1 <= lim * (ATOL + RTOL * np.abs(want))).all() and (e2 <= lim * (ATOL + RTOL * np.abs(o.pose))).all(), (name, i, t, e1, e2)
    print(f"rollout engine {name}: SQP + shift, {T} ticks x {B} robots, {nsteps / (B * T):.2f} RTI steps per tick, worst |diff| {worst:.2e}")
    assert nsteps > B * T                 # the SQP loop did iterate
    ctl.close()


def test_config4_tric_200_ticks_sqp_shift_through_one_call():
    """BASELINE config 4: tric (steering-angle and steering-rate constraints), SQP to convergence, warm-start shift, 200
    closed-loop ticks - one C call, nothing on the host between ticks; repeatable bit for bit, constraints respected"""
    B, T = 192, 200
    spec, paths, pid, u0, pose0, ctl, ro, rng = _setup("tric", B, 55, 16)
    p0, uu = torch.from_numpy(pose0.T.copy()).cuda(), torch.from_numpy(u0).cuda()
    outs = []
    for _ in range(2):
        ro.reset(p0, uu)
        # robots already rolling and steering: at rest with the wheel straight the tric model is locally uncontrollable
        # (every pose rate carries sin(alpha), scripts/tric/tric_amr_model.py:45-49), the zero iterate is stationary and
        # the controller never moves
        ro.x[3].fill_(0.3); ro.x[4].fill_(0.15); ro.x[5].fill_(0.3); ro.x[6].fill_(0.15)
        ro.vel[0].fill_(0.3); ro.steer.fill_(0.15)
        ctl.reference_states()[0, :B].fill_(0.3); ctl.reference_states()[1, :B].fill_(0.15)
        t0 = time.perf_counter()
        r = ro.run_engine(T, None, sqp_max_iter=6, sqp_tol=1e-8, shift=True)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        outs.append({k: v.clone() for k, v in r.items()})
        assert int(r["failed"].sum()) == 0 and torch.isfinite(r["pose"]).all() and torch.isfinite(r["cmd"]).all()
    assert torch.equal(outs[0]["pose"], outs[1]["pose"]) and torch.equal(outs[0]["cmd"], outs[1]["cmd"])
    deg = np.pi / 180.0
    cmd = outs[0]["cmd"].cpu().numpy()                       # tric: (v_ref, alpha_ref, 0)
    assert np.abs(cmd[:, 0]).max() <= 1.0 + 1e-9 and np.abs(cmd[:, 1]).max() <= 30.0 * deg + 1e-9
    dalpha = np.abs(np.diff(cmd[:, 1], axis=0)).max() / spec.dt
    assert dalpha <= 120.0 * deg + 1e-6                      # steering-rate bound along the closed loop
    assert np.abs(cmd[:, 0]).max() > 0.05
    moved = np.hypot(*(outs[0]["pose"][-1, :2] - outs[0]["pose"][0, :2]).cpu().numpy())
    assert moved.max() > 0.05
    x = ro.x.cpu().numpy()
    assert np.abs(x[6]).max() <= 30.0 * deg + 1e-9
    print(f"config 4: tric {B} robots x {T} ticks, SQP + shift in one call: max |v_ref| {np.abs(cmd[:, 0]).max():.3f}, "
          f"max |alpha_ref| {np.abs(cmd[:, 1]).max() / deg:.1f} deg, max steering rate {dalpha / deg:.1f} deg/s; "
          f"{wall:.2f} s for the call = {wall / T * 1e3:.2f} ms per tick")
    ctl.close()
