"""GPU parity: the CUDA path through the C ABI vs the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest
import torch

from helpers import instances, oracle_solve, parity_report
from nmpc_nav_control_b200.problem import MODELS

pytestmark = pytest.mark.gpu


def _solver(name, cap):
    from nmpc_nav_control_b200.solver import BatchedRtiSolver
    return BatchedRtiSolver(name, cap)


def _to_soa(a):  # [B, ...] -> [..., B] contiguous on the GPU
    t = torch.from_numpy(a)
    perm = list(range(1, t.dim())) + [0]
    return t.permute(*perm).contiguous().cuda()


@pytest.mark.parametrize("name,B", [("diff", 1000), ("omni4", 500), ("tric", 777)])
def test_device_batch_matches_oracle(oracle_mod, name, B):
    """ragged batch (not a multiple of 32), cold (reset) iterate, device-resident SoA inputs"""
    spec, x0, yref, _ = instances(name, 0, B)
    ref = oracle_solve(oracle_mod, name, x0, yref)
    s = _solver(name, B)
    s.reset()
    out = s.solve_device(_to_soa(x0), _to_soa(yref), want_stats=True)
    torch.cuda.synchronize()
    x, u = s.get_iterate(B)
    status = out["status"].cpu().numpy(); it = out["qp_iter"].cpu().numpy()
    assert (status == ref["status"]).all()
    mism = int((it != ref["qp_iter"]).sum())
    nbx, ex = parity_report(x, ref["x"]); nbu, eu = parity_report(u, ref["u"])
    print(f"{name}: B={B} qp_iter mismatches={mism} max|dx|={ex:.3e} max|du|={eu:.3e} "
          f"lin_res={out['stats'][5].max().item():.2e}")
    assert mism == 0
    assert nbx == 0 and nbu == 0


@pytest.mark.parametrize("name", ["diff", "omni4", "tric"])
def test_host_call_and_warm_second_step(oracle_mod, name):
    """controller-facing call with host buffers; a second RTI step from the persisted iterate"""
    B = 96
    spec, x0, yref, _ = instances(name, 5000, B, pose_only=True)
    yfull = np.zeros((B, spec.n + 1, spec.ny)); yfull[:, :, :3] = yref
    s = _solver(name, B)
    s.reset()
    out = s.solve_host(x0, yref)
    ref = oracle_solve(oracle_mod, name, x0, yfull)
    assert (out["status"] == 0).all() and (out["qp_iter"] == ref["qp_iter"]).all()
    assert parity_report(out["u0"], ref["u"][:, 0])[0] == 0
    assert parity_report(out["x1"], ref["x"][:, 1])[0] == 0
    # second tick: x0 <- x1 (as the wrapper does, NMPCNavControlDiff.cpp:168-172), same refs
    x0b = out["x1"].copy()
    out2 = s.solve_host(x0b, yref)
    ref2 = oracle_solve(oracle_mod, name, x0b, yfull, x=ref["x"], u=ref["u"])
    assert (out2["qp_iter"] == ref2["qp_iter"]).all()
    assert parity_report(out2["u0"], ref2["u"][:, 0])[0] == 0
    assert parity_report(out2["x1"], ref2["x"][:, 1])[0] == 0


def test_diff_terminal_weight_switch(oracle_mod):
    """per-instance W_e (the diff wrapper's x1/x100 switch, NMPCNavControlDiff.cpp:127-139)"""
    B = 200
    spec, x0, yref, We = instances("diff", 100, B, terminal_hack=True)
    assert We is not None and len(np.unique(We[:, 0])) == 2
    ref = oracle_solve(oracle_mod, "diff", x0, yref, We=We)
    s = _solver("diff", B)
    s.reset()
    out = s.solve_host(x0, yref, We=We)
    assert (out["qp_iter"] == ref["qp_iter"]).all()
    assert parity_report(out["u0"], ref["u"][:, 0])[0] == 0


def test_smoke_case_default_iterate(oracle_mod):
    """the codegen smoke solve (scripts/diff/generate_c_code.py:79-83): x0 = default, yref = 0,
    iterate as created (x_k = x0 default)"""
    for name, spec in MODELS.items():
        x0 = np.array(spec.x0_default)[None, :]
        yref = np.zeros((1, spec.n + 1, spec.ny))
        xi = np.tile(x0, (1, spec.n + 1, 1)).reshape(1, spec.n + 1, spec.nx)
        ref = oracle_solve(oracle_mod, name, x0, yref, x=xi)
        s = _solver(name, 1)
        out = s.solve_host(x0, yref)
        assert out["status"][0] == 0 and out["qp_iter"][0] == ref["qp_iter"][0]
        assert parity_report(out["u0"], ref["u"][:, 0])[0] == 0


def test_full_size_properties():
    """BASELINE config 2 size (65,536 diff instances): size-independent properties"""
    from nmpc_nav_control_b200 import synth
    spec = MODELS["diff"]
    B = 65536
    inst = synth.make_instances(spec, 0, B, device="cuda", pose_only=True)
    x0 = inst["x0"].t().contiguous(); yref = inst["yref"].permute(1, 2, 0).contiguous()
    s = _solver("diff", B)
    s.reset()
    out = s.solve_device(x0, yref)
    torch.cuda.synchronize()
    assert int((out["status"] != 0).sum()) == 0
    dx, du, ld = None, None, None
    import ctypes as C
    px, pu, pl = C.c_void_p(), C.c_void_p(), C.c_int()
    s.lib.nmpc_iterate_device(s._h, C.byref(px), C.byref(pu), C.byref(pl))
    x, u = s.get_iterate(4096)
    # x_0 restored to the measurement; bounds respected (interior point => strictly inside)
    assert np.abs(x[:, 0, :] - inst["x0"][:4096].cpu().numpy()).max() == 0.0
    assert u.min() >= -2.0 - 1e-9 and u.max() <= 2.0 + 1e-9
    assert x[:, 1:, 5:7].min() >= -1.0 - 1e-9 and x[:, 1:, 5:7].max() <= 1.0 + 1e-9
    # sharding invariance: instance i is the same whatever batch it is solved in
    s2 = _solver("diff", 4096)
    s2.reset()
    out2 = s2.solve_device(x0[:, 1000:1000 + 4096].contiguous(), yref[:, :, 1000:1000 + 4096].contiguous())
    torch.cuda.synchronize()
    x2, u2 = s2.get_iterate(4096)
    xa, ua = s.get_iterate(1000 + 4096)
    # (the large batch runs the hybrid K3 schedule, the small one the lane-group kernel alone: same iteration path,
    # different summation order, so equal iteration counts and agreement inside the parity tolerance, not bit for bit)
    assert torch.equal(out2["qp_iter"], out["qp_iter"][1000:1000 + 4096])
    assert parity_report(x2, xa[1000:])[0] == 0 and parity_report(u2, ua[1000:])[0] == 0


@pytest.mark.parametrize("name,B", [("diff", 40000), ("omni4", 3000), ("tric", 30000)])
def test_repeatable_bit_for_bit(name, B):
    """the same batch solved twice gives bit-identical results: the work queue, the hand-over order and the slot an
    instance lands in change from run to run, the arithmetic of an instance must not (a missing warp
    synchronisation or a race on the records would show here)"""
    from nmpc_nav_control_b200 import synth
    spec = MODELS[name]
    inst = synth.make_instances(spec, 123, B, device="cuda", pose_only=True)
    x0 = inst["x0"].t().contiguous(); yref = inst["yref"].permute(1, 2, 0).contiguous()
    s = _solver(name, B)
    res = []
    for _ in range(2):
        s.reset()
        out = s.solve_device(x0, yref, want_stats=True)
        torch.cuda.synchronize()
        x, u = s.get_iterate(B)
        res.append((x, u, out["qp_iter"].cpu().numpy().copy(), out["status"].cpu().numpy().copy()))
    assert (res[0][3] == 0).all()
    assert np.array_equal(res[0][2], res[1][2])
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])


@pytest.mark.parametrize("name", ["diff", "omni4", "tric"])
def test_golden_vectors_two_steps(name):
    """committed fixtures (tests/golden/make_golden.py): no oracle at run time"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"rti_{name}.npz"))
    B = g["x0"].shape[0]
    s = _solver(name, B)
    s.reset()
    out = s.solve_host(np.ascontiguousarray(g["x0"]), np.ascontiguousarray(g["yref"]))
    x, u = s.get_iterate(B)
    assert (out["qp_iter"] == g["qp_iter"]).all() and (out["status"] == g["status"]).all()
    assert parity_report(x, g["x"])[0] == 0 and parity_report(u, g["u"])[0] == 0
    out2 = s.solve_host(np.ascontiguousarray(out["x1"]), np.ascontiguousarray(g["yref"]))
    x2, u2 = s.get_iterate(B)
    assert (out2["qp_iter"] == g["qp_iter2"]).all()
    lr = g["lin_res2"]
    assert parity_report(x2, g["x2"], lr)[0] == 0 and parity_report(u2, g["u2"], lr)[0] == 0


def test_error_paths_and_bad_instances(oracle_mod):
    """argument errors are integer codes (the wrapper turns non-zero into std::runtime_error, NMPCNavControl.cpp:5-23; the
    library never throws or aborts); a NaN in one instance's measurement fails that instance only (acados status 4 = QP
    failure, iterate left untouched) and the rest of the batch solves as if it were not there; max-iter is accepted"""
    import ctypes as C
    from nmpc_nav_control_b200 import _lib
    from nmpc_nav_control_b200.solver import BatchedRtiSolver
    spec = MODELS["diff"]
    B = 300
    _, x0, yref, _ = instances("diff", 40, B)
    s = BatchedRtiSolver(spec, B)
    lib = s.lib
    # capacity / argument errors
    u0 = np.empty((B + 1, spec.nu)); x1 = np.empty((B + 1, spec.nx)); st = np.empty(B + 1, np.int32); it = np.empty(B + 1, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    big = np.zeros((B + 1, spec.nx)); bigy = np.zeros((B + 1, spec.n + 1, spec.ny))
    assert lib.nmpc_rti_solve_host(s._h, B + 1, p(big), p(bigy), spec.ny, None, p(u0), p(x1), p(st), p(it)) == -4      # NMPC_E_CAPACITY
    assert lib.nmpc_rti_solve_host(s._h, 0, p(big), p(bigy), spec.ny, None, p(u0), p(x1), p(st), p(it)) == -1          # NMPC_E_ARG
    assert lib.nmpc_rti_solve_host(s._h, B, p(big), p(bigy), 5, None, p(u0), p(x1), p(st), p(it)) == -1                # nyref must be 3 or ny
    assert len(lib.nmpc_last_error()) > 0
    # one poisoned instance
    ref = oracle_solve(oracle_mod, "diff", x0, yref)
    xb = x0.copy(); xb[17, 2] = np.nan
    s.reset()
    out = s.solve_host(xb, yref)
    assert out["status"][17] == 4 and (np.delete(out["status"], 17) == 0).all()
    keep = np.arange(B) != 17
    assert (out["qp_iter"][keep] == ref["qp_iter"][keep]).all()
    assert parity_report(out["u0"][keep], ref["u"][keep, 0])[0] == 0
    x, u = s.get_iterate(B)
    assert np.abs(x[17, 1:]).max() == 0.0 and np.abs(u[17]).max() == 0.0          # the failed instance's iterate is untouched (reset = zero)
    # iteration limit: max-iter QPs are accepted (acados SQP_RTI takes the QP max-iter solution), count = limit
    s.set_opts(iter_max=4)
    s.reset()
    out = s.solve_host(x0, yref)
    assert (out["status"] == 0).all() and out["qp_iter"].max() == 4 and (out["qp_iter"] == np.minimum(ref["qp_iter"], 4)).all()
    s.close()


@pytest.mark.parametrize("name,B", [("diff", 65536), ("tric", 65536), ("omni4", 32768)])
def test_benchmarked_size_and_schedule_against_the_oracle(oracle_mod, name, B):
    """the bench workload as the bench runs it (default schedule: lockstep sweeps, hand-over, lane-cooperative kernel),
    compared DIRECTLY with the oracle on ALL instances, cold step and the second (warm) step: per-instance QP iteration
    counts and the whole x / u trajectories.  Reports the iteration-count mismatch fraction (SURVEY.md 7 hard part 5:
    termination-test flips at rounding level must be rare, never hidden) and how many instances needed the widened bound
    of helpers.parity_report (oracle's own Newton residual above 1e-10: near-degenerate QPs)."""
    spec, x0, yref, _ = instances(name, 0, B, pose_only=True)
    yfull = np.zeros((B, spec.n + 1, spec.ny)); yfull[:, :, :3] = yref
    s = _solver(name, B)
    s.reset()
    xs, us = None, None
    x0k = x0
    for step in (1, 2):
        out = s.solve_device(_to_soa(x0k), _to_soa(yref))
        torch.cuda.synchronize()
        x, u = s.get_iterate(B)
        ref = oracle_solve(oracle_mod, name, x0k, yfull, x=xs, u=us, fast=False)
        status = out["status"].cpu().numpy(); it = out["qp_iter"].cpu().numpy()
        assert (status == 0).all() and (ref["status"] == 0).all()
        same = it == ref["qp_iter"]
        mism = int((~same).sum())
        lr = ref["lin_res"]
        widened = int((lr > 1e-10).sum())
        nb_strict = parity_report(x[same], ref["x"][same])[0] + parity_report(u[same], ref["u"][same])[0]
        nbx, ex = parity_report(x[same], ref["x"][same], lr[same]); nbu, eu = parity_report(u[same], ref["u"][same], lr[same])
        print(f"{name} B={B} step {step}: qp_iter mismatch fraction {mism / B:.2e} ({mism}), instances with the widened bound "
              f"{widened}, outside the strict 1e-9 bound {nb_strict}, outside the widened bound {nbx + nbu}, worst |diff| {max(ex, eu):.2e}, "
              f"mean qp_iter {it.mean():.3f}")
        # measured (round 2): no iteration-count mismatch in 2 x 163,840 solves; cold step: 0.14 % of the diff QPs are
        # near-degenerate by the oracle's own Newton residual (none for tric), 0.02 % of the instances leave the strict 1e-9
        # bound; warm step: 2.2 % near-degenerate, 0.1 - 0.2 % outside 1e-9, worst 1.5e-7, at most 5 outside the widened bound
        assert mism / B <= 1e-4
        assert widened <= B // 25 and nb_strict <= B // 200
        assert nbx + nbu <= max(2, B // 5000) and max(ex, eu) < 1e-6
        xs, us = ref["x"], ref["u"]
        x0k = ref["x"][:, 1].copy()                       # second tick: x0 <- x1 (NMPCNavControlDiff.cpp:168-172), same references
        s.set_iterate(xs, us)                             # both continue from the oracle's iterate: every solve on identical inputs
    s.close()
