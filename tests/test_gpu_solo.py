"""GPU parity of the block-per-instance K3 mapping (csrc/rti_solo.cuh, the latency path for small batches; omni4 keeps its two
largest per-stage arrays in a global scratch):
against the CPU oracle on the same seeded inputs, and against the lane-cooperative kernel it replaces at those sizes."""
import os

import numpy as np
import pytest
import torch

from helpers import instances, oracle_solve, parity_report

pytestmark = pytest.mark.gpu


def _solver(name, cap, solo_max=None):
    """the schedule is chosen when the solver is created (NMPC_SOLO_MAX: largest batch that goes to the solo kernel)"""
    from nmpc_nav_control_b200.solver import BatchedRtiSolver
    old = os.environ.get("NMPC_SOLO_MAX")
    if solo_max is not None:
        os.environ["NMPC_SOLO_MAX"] = str(solo_max)
    try:
        return BatchedRtiSolver(name, cap)
    finally:
        if solo_max is not None:
            if old is None:
                del os.environ["NMPC_SOLO_MAX"]
            else:
                os.environ["NMPC_SOLO_MAX"] = old


def _to_soa(a):
    t = torch.from_numpy(a)
    perm = list(range(1, t.dim())) + [0]
    return t.permute(*perm).contiguous().cuda()


@pytest.mark.parametrize("name,B,start", [("diff", 1, 0), ("diff", 37, 100), ("diff", 300, 2000), ("tric", 1, 3), ("tric", 150, 500),
                                          ("omni4", 1, 11), ("omni4", 70, 640)])
def test_solo_matches_oracle_cold_and_warm(oracle_mod, name, B, start):
    """cold step and a second (warm) step from the persisted iterate, device-resident inputs, solo kernel forced"""
    spec, x0, yref, _ = instances(name, start, B)
    s = _solver(name, B, solo_max=100000)
    s.reset()
    ref = oracle_solve(oracle_mod, name, x0, yref)
    out = s.solve_device(_to_soa(x0), _to_soa(yref), want_stats=True)
    torch.cuda.synchronize()
    x, u = s.get_iterate(B)
    assert (out["status"].cpu().numpy() == ref["status"]).all()
    assert (out["qp_iter"].cpu().numpy() == ref["qp_iter"]).all()
    nbx, ex = parity_report(x, ref["x"], ref["lin_res"]); nbu, eu = parity_report(u, ref["u"], ref["lin_res"])
    print(f"{name} B={B} cold: max|dx|={ex:.2e} max|du|={eu:.2e}")
    assert nbx == 0 and nbu == 0
    # the statistics row of the solve: residual norms at exit below the tolerances, mu > 0
    st = out["stats"].cpu().numpy()
    assert (st[0] <= 1e-6).all() and (st[1] <= 1e-8).all() and (st[4] > 0).all()
    # warm step: x0 <- x1 of the oracle's iterate
    x0b = ref["x"][:, 1].copy()
    ref2 = oracle_solve(oracle_mod, name, x0b, yref, x=ref["x"], u=ref["u"])
    s.set_iterate(ref["x"], ref["u"])
    out2 = s.solve_device(_to_soa(x0b), _to_soa(yref))
    torch.cuda.synchronize()
    x2, u2 = s.get_iterate(B)
    assert (out2["qp_iter"].cpu().numpy() == ref2["qp_iter"]).all()
    assert parity_report(x2, ref2["x"], ref2["lin_res"])[0] == 0 and parity_report(u2, ref2["u"], ref2["lin_res"])[0] == 0
    s.close()


@pytest.mark.parametrize("name", ["diff", "tric", "omni4"])
def test_solo_equals_cooperative_kernel(name):
    """the two mappings a small batch can take: equal iteration counts, results to rounding; the solo kernel is repeatable bit for bit"""
    B = 96
    spec, x0, yref, _ = instances(name, 7000, B, pose_only=True)
    res = []
    for solo_max in (100000, 0, 100000):
        s = _solver(name, B, solo_max=solo_max)
        s.reset()
        out = s.solve_host(x0, yref)
        res.append({k: v.copy() for k, v in out.items()})
        s.close()
    a, b, c = res
    assert (a["qp_iter"] == b["qp_iter"]).all() and (a["status"] == b["status"]).all()
    assert np.abs(a["u0"] - b["u0"]).max() < 1e-9 and np.abs(a["x1"] - b["x1"]).max() < 1e-9
    assert np.array_equal(a["u0"], c["u0"]) and np.array_equal(a["x1"], c["x1"]) and np.array_equal(a["qp_iter"], c["qp_iter"])


def test_solo_per_instance_terminal_weight_and_nan_isolation(oracle_mod):
    """diff: the per-instance W_e of the controller tick (NMPCNavControlDiff.cpp:126-139) reaches the solo kernel; a NaN instance
    does not disturb its neighbours"""
    name, B = "diff", 12
    spec, x0, yref, _ = instances(name, 900, B)
    tb = spec.codegen_defaults()
    We = np.tile(np.asarray(tb["We"], dtype=np.float64), (B, 1))
    We[::2, :3] *= 100.0
    ref = oracle_solve(oracle_mod, name, x0, yref, We=We)
    s = _solver(name, B, solo_max=100000)
    s.reset()
    out = s.solve_device(_to_soa(x0), _to_soa(yref), We=_to_soa(We))
    torch.cuda.synchronize()
    x, u = s.get_iterate(B)
    assert (out["qp_iter"].cpu().numpy() == ref["qp_iter"]).all()
    assert parity_report(x, ref["x"], ref["lin_res"])[0] == 0 and parity_report(u, ref["u"], ref["lin_res"])[0] == 0
    x0n = x0.copy(); x0n[5, 2] = np.nan
    s.reset()
    out = s.solve_device(_to_soa(x0n), _to_soa(yref), We=_to_soa(We))
    torch.cuda.synchronize()
    st = out["status"].cpu().numpy()
    assert st[5] != 0 and (np.delete(st, 5) == 0).all()
    x2, u2 = s.get_iterate(B)
    keep = np.arange(B) != 5
    assert np.array_equal(x2[keep], x[keep]) and np.array_equal(u2[keep], u[keep])
    s.close()


@pytest.mark.parametrize("name,nyfull", [("diff", False), ("diff", True), ("tric", False), ("omni4", True)])
def test_small_batch_host_path_equals_the_general_host_path(oracle_mod, name, nyfull):
    """nmpc_rti_solve_host stages batches of at most 64 instances through one pinned buffer each way (the ROS drop-in's call,
    NMPCNavControlDiff.cpp:96-169: x0 / yref / W_e in, u_0 / x_1 / status out).  The instances of a 65-instance call (general
    path: one copy and one transposition per argument) solved in calls of 1, 7 and 57 give the same bits, with and without the
    per-instance W_e, with pose-only and full reference rows; and the oracle agrees."""
    B = 65
    spec, x0, yref, _ = instances(name, 4242, B, pose_only=not nyfull)
    tb = spec.codegen_defaults()
    We = np.tile(np.asarray(tb["We"], dtype=np.float64), (B, 1))
    We[::3, :3] *= 100.0
    for use_we in (False, True):
        w = We if use_we else None
        s = _solver(name, B, solo_max=100000)
        s.reset()
        big = {k: v.copy() for k, v in s.solve_host(x0, yref, We=w).items()}
        s.close()
        yfull = np.zeros((B, spec.n + 1, spec.ny)); yfull[:, :, :yref.shape[2]] = yref     # reference rows beyond the pose are zero
        ref = oracle_solve(oracle_mod, name, x0, yfull, We=w)
        assert (big["qp_iter"] == ref["qp_iter"]).all() and (big["status"] == 0).all()
        assert parity_report(big["u0"], ref["u"][:, 0], ref["lin_res"])[0] == 0
        lo = 0
        for n in (1, 7, 57):
            s = _solver(name, n, solo_max=100000)
            s.reset()
            sl = slice(lo, lo + n)
            out = s.solve_host(x0[sl], yref[sl], We=None if w is None else w[sl])
            for k in ("u0", "x1", "status", "qp_iter"):
                assert np.array_equal(out[k], big[k][sl]), (name, use_we, n, k)
            s.close()
            lo += n
