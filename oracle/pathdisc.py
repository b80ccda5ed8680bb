"""TEST INFRASTRUCTURE - `PathDiscretizer::getNextNPoses` restated.

`get_next_n_poses` follows src/nmpc_nav_control/PathDiscretizer.cpp:14-63 (getPoseSample :65-86, getVelSample :88-105)
statement by statement in plain Python over the stand-in curve family of oracle/stubs/parametric_trajectories_common
(the reference's TPath is a private, absent dependency).  PINNED against the reference itself: `ref()` loads
oracle/_ref/libpathdisc_ref.so, the reference's UNMODIFIED PathDiscretizer.cpp compiled in place by `make -C oracle ref`,
and tests/test_pathdisc_cpu.py compares the two (and the committed golden vectors tests/golden/pathdisc.npz made by
tests/golden/make_golden_pathdisc.py from that library).  Only tests/ may import this module."""
import ctypes as C
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libpathdisc_ref.so")


def _poly(c, u):
    return ((((c[5] * u + c[4]) * u + c[3]) * u + c[2]) * u + c[1]) * u + c[0]


def _dpoly(c, u):
    return (((5.0 * c[5] * u + 4.0 * c[4]) * u + 3.0 * c[3]) * u + 2.0 * c[2]) * u + c[1]


class Seg:
    """the stand-in TPath (oracle/stubs/parametric_trajectories_common/trajectory_common.h)"""

    def __init__(self, row):
        self.kind, self.vel, self.th0, self.th1 = row[0], row[1], row[2], row[3]
        self.cx, self.cy = list(row[4:10]), list(row[10:16])

    def x(self, u):
        return _poly(self.cx, u) if self.kind == 0.0 else self.cx[0] + self.cx[1] * math.cos(self.cx[2] + self.cx[3] * u)

    def y(self, u):
        return _poly(self.cy, u) if self.kind == 0.0 else self.cy[0] + self.cx[1] * math.sin(self.cx[2] + self.cx[3] * u)

    def dx(self, u):
        return _dpoly(self.cx, u) if self.kind == 0.0 else -self.cx[1] * self.cx[3] * math.sin(self.cx[2] + self.cx[3] * u)

    def dy(self, u):
        return _dpoly(self.cy, u) if self.kind == 0.0 else self.cx[1] * self.cx[3] * math.cos(self.cx[2] + self.cx[3] * u)

    def theta(self, u):
        return math.atan2(self.dy(u), self.dx(u))

    def theta_h(self, u):
        return self.th0 + (self.th1 - self.th0) * u


def _locate(path, sample_u):
    """PathDiscretizer.cpp:67-76 / :90-99"""
    num = math.floor(sample_u)
    u = sample_u - float(num)
    if num >= len(path):
        num, u = len(path) - 1, 1.0
    elif num < 0:
        num, u = 0, 0.0
    return path[num], u


def _pose(path, sample_u, holonomic):
    p, u = _locate(path, sample_u)
    if not holonomic:                                     # :80-83
        th = p.theta(u) if p.vel >= 0 else p.theta(u) + math.pi
    else:
        th = p.theta_h(u)
    return p.x(u), p.y(u), th


def _vel(path, sample_u):
    p, u = _locate(path, sample_u)
    return p.dx(u), p.dy(u)


def get_next_n_poses(segments, nearest_sample_u, sample_period, num_poses, holonomic=False):
    """segments [n_seg, 16]; returns poses [num_poses, 3]"""
    path = [Seg(r) for r in np.asarray(segments, dtype=np.float64).reshape(-1, 16)]
    per_cycle = 20 if sample_period >= 1.0 else 10        # :9-11
    thr = 1e-2
    N = float(len(path))
    vel = abs(path[int(math.floor(nearest_sample_u))].vel)
    goal = vel * sample_period
    rel = goal / per_cycle
    u = nearest_sample_u
    old = _pose(path, nearest_sample_u, holonomic)
    vx, vy = _vel(path, nearest_sample_u)
    step = rel / math.sqrt(vx ** 2 + vy ** 2)
    out = []
    curr = 0.0
    while u < N:                                          # :34-56
        u += step
        u = min(u, N)
        new = _pose(path, u, holonomic)
        curr += math.sqrt((new[0] - old[0]) ** 2 + (new[1] - old[1]) ** 2)
        if (goal - curr) <= thr * goal:
            out.append(new)
            vel = abs(path[int(min(math.floor(u), N - 1))].vel)
            goal = vel * sample_period
            rel = goal / per_cycle
            curr = 0.0
        if num_poses == len(out):
            break
        vx, vy = _vel(path, u)
        step = rel / math.sqrt(vx ** 2 + vy ** 2)
        old = new
    if num_poses > len(out):                              # :58-63
        last = _pose(path, N, holonomic)
        while num_poses > len(out):
            out.append(last)
    return np.array(out, dtype=np.float64)


_ref = None


def build_ref():
    """compile the reference's own discretiser when the reference tree is present (this container)"""
    if os.path.isdir("/root/reference"):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)
    return os.path.exists(REF_SO)


def ref(segments, nearest_sample_u, sample_period, num_poses, holonomic=False):
    """the reference's compiled getNextNPoses (oracle/_ref); raises if the library was never built"""
    global _ref
    if _ref is None:
        _ref = C.CDLL(REF_SO)
        _ref.pathdisc_ref.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p]
    seg = np.ascontiguousarray(segments, dtype=np.float64).reshape(-1, 16)
    out = np.zeros((num_poses, 3))
    n = _ref.pathdisc_ref(seg.ctypes.data, len(seg), float(nearest_sample_u), float(sample_period), int(num_poses), int(holonomic),
                          out.ctypes.data)
    assert n == num_poses
    return out
