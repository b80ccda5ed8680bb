"""SURVEY.md 8(f2), CPU part.
1. the oracle restatement (oracle/pathdisc.py) against the golden vectors made by the reference's own compiled
   discretiser (tests/golden/pathdisc.npz), and against that library itself when it is present (oracle/_ref);
2. the device code of nmpc_nav_control_b200/csrc/path_disc.cuh, compiled for the host by tests/host_emul, against the
   oracle on the same paths: ragged paths, starts on segment boundaries and at the path's end (padding), reversed
   segments, holonomic headings, both sample-period regimes (10 / 20 points per cycle)."""
import math
import os

import numpy as np
import pytest

import emul
import pathcases
from nmpc_nav_control_b200 import paths as P
from oracle import pathdisc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pathdisc.npz")
CFG = [(0, 0.025, 81), (0, 1.0, 12), (1, 0.025, 81), (1, 1.0, 12)]


def _golden():
    g = np.load(GOLD)
    off = g["offsets"]
    paths = [g["segments"][off[i]:off[i + 1]] for i in range(len(off) - 1)]
    return g, paths


@pytest.mark.parametrize("hol,period,num", CFG)
def test_oracle_matches_reference_golden_vectors(hol, period, num):
    g, paths = _golden()
    want = g[f"poses_h{hol}_T{period}_n{num}"]
    for i, (p, u) in enumerate(zip(g["path_id"], g["u0"])):
        got = pathdisc.get_next_n_poses(paths[p], u, period, num, bool(hol))
        assert np.array_equal(got, want[i]), (i, np.abs(got - want[i]).max())       # same libm, same order: bit for bit


def test_oracle_matches_compiled_reference_on_fresh_cases():
    if not pathdisc.build_ref():
        pytest.skip("oracle/_ref/libpathdisc_ref.so not built (needs the reference tree)")
    paths, pid, u0 = pathcases.cases(seed=77, n_paths=20, B=200)
    for p, u in zip(pid, u0):
        for hol in (False, True):
            assert np.array_equal(pathdisc.get_next_n_poses(paths[p], u, 0.025, 81, hol), pathdisc.ref(paths[p], u, 0.025, 81, hol))


@pytest.mark.parametrize("hol,period,num", CFG)
def test_device_code_matches_reference_golden_vectors(hol, period, num):
    g, _ = _golden()
    want = g[f"poses_h{hol}_T{period}_n{num}"]                     # [B, num, 3]
    got = emul.emul_path_discretize(g["segments"], g["offsets"], g["path_id"], g["u0"], period, num, bool(hol))
    got = np.moveaxis(got, -1, 0)
    # sincos / atan2 come from the same libm here; only the fused evaluation order of the arc differs
    assert np.abs(got - want).max() <= 1e-12, np.abs(got - want).max()


def test_device_code_edge_cases():
    # one pose; a path of one short segment (all padding); start past the end; zero speed; start before the path
    ln = P.line((0, 0), (0.01, 0), 0.5)
    a = emul.emul_path_discretize(ln, [0, 1], [0, 0, 0, 0], [0.0, 0.999, 5.0, -3.0], 0.025, 81)
    o = [pathdisc.get_next_n_poses(ln, u, 0.025, 81) for u in (0.0, 0.999)]
    assert np.array_equal(a[:, :, 0], o[0]) and np.array_equal(a[:, :, 1], o[1])
    assert np.array_equal(a[-1, :, 0], [0.01, 0.0, 0.0])                        # padded with the end pose
    assert np.array_equal(a[:, :, 2], np.tile([0.01, 0.0, 0.0], (81, 1)))       # start past the end: the end pose throughout
    assert np.isfinite(a[:, :, 3]).all() and a[-1, 0, 3] == 0.01                # start before the path: walks in from u < 0
    stop = P.line((0, 0), (1, 0), 0.0)
    z = emul.emul_path_discretize(stop, [0, 1], [0], [0.25], 0.025, 5)          # speed 0: every step emits the current pose
    assert np.array_equal(z[:, :, 0], pathdisc.get_next_n_poses(stop, 0.25, 0.025, 5))
    one = emul.emul_path_discretize(ln, [0, 1], [0], [0.0], 0.025, 1)
    assert one.shape == (1, 3, 1)
    # reversed segment: heading flipped by pi (PathDiscretizer.cpp:80-83)
    back = P.line((0, 0), (1, 0), -0.5)
    b = emul.emul_path_discretize(back, [0, 1], [0], [0.0], 0.025, 3)
    assert np.allclose(b[:, 2, 0], math.pi)


def test_spacing_property():
    """poses are one sample period of travel apart (within the 1 % threshold plus one sub-step) until the path ends"""
    paths, pid, u0 = pathcases.cases(seed=5, n_paths=6, B=64)
    off = np.cumsum([0] + [len(p) for p in paths]).astype(np.int32)
    out = emul.emul_path_discretize(np.concatenate(paths), off, pid, np.zeros(64), 0.025, 40)
    for i in range(64):
        d = np.hypot(np.diff(out[:, 0, i]), np.diff(out[:, 1, i]))
        vmax = np.abs(paths[pid[i]][:, 1]).max(); vmin = np.abs(paths[pid[i]][:, 1]).min()
        moving = d > 0
        assert (d[moving] <= 1.11 * vmax * 0.025 + 1e-12).all() and (d[moving][:-1] >= 0.85 * vmin * 0.025).all(), i


def test_restatement_and_device_code_against_compiled_reference_random_search():
    """hypothesis-driven search over path shapes, start parameters, sample periods and pose counts: the restatement is
    bit-identical to the reference's compiled discretiser, the device code (host build) agrees within 1e-12"""
    if not pathdisc.build_ref():
        pytest.skip("oracle/_ref/libpathdisc_ref.so not built (needs the reference tree)")
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=120, deadline=None, derandomize=True)
    @given(seed=st.integers(0, 10**6), nseg=st.integers(1, 6), frac=st.floats(0.0, 0.999), period=st.sampled_from([0.02, 0.025, 0.1, 1.0, 2.5]),
           num=st.integers(1, 90), hol=st.booleans())
    def check(seed, nseg, frac, period, num, hol):
        rng = np.random.default_rng(seed)
        path = pathcases.random_path(rng, nseg)
        u0 = frac * nseg
        want = pathdisc.ref(path, u0, period, num, hol)
        got = pathdisc.get_next_n_poses(path, u0, period, num, hol)
        assert np.array_equal(got, want)
        dev = emul.emul_path_discretize(path, [0, nseg], [0], [u0], period, num, hol)[:, :, 0]
        assert np.abs(dev - want).max() <= 1e-12
    check()
