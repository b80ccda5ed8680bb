"""TEST INFRASTRUCTURE - the closed loop of nmpc_nav_control_b200/rollout.py restated on the CPU, one robot at a time
(SURVEY.md 8(f3)): nearest path parameter -> oracle/pathdisc.py -> oracle/ctrl.py (around the solver oracle) -> plant
step with the oracle's RK4 map.  The nearest-point search is a stand-in defined by this repository
(csrc/rollout.cuh; the reference's TPathProcessMinDist is private), restated here - that piece is "parity unpinned";
the other three links are pinned against the reference's own sources (see their headers)."""
import numpy as np

from oracle import pathdisc
from oracle.ctrl import OracleController


def nearest_u(segments, u_prev, px, py, back, ahead, ns=24, nt=30):
    path = [pathdisc.Seg(r) for r in np.asarray(segments, dtype=np.float64).reshape(-1, 16)]
    N = float(len(path))

    def d2(su):
        p, u = pathdisc._locate(path, su)
        return (p.x(u) - px) * (p.x(u) - px) + (p.y(u) - py) * (p.y(u) - py)
    lo, hi = max(u_prev - back, 0.0), min(u_prev + ahead, N)
    if not lo < hi:
        return 0.0 if hi < 0.0 else hi
    h = (hi - lo) / ns
    best, dbest = 0, d2(lo)
    for i in range(1, ns + 1):
        d = d2(lo + h * i)
        if d < dbest:
            dbest, best = d, i
    a = lo + h * (best - 1 if best > 0 else 0)
    b = lo + h * (best + 1 if best < ns else ns)
    for _ in range(nt):
        m1, m2 = a + (b - a) / 3.0, b - (b - a) / 3.0
        if d2(m1) <= d2(m2):
            b = m2
        else:
            a = m1
    return 0.5 * (a + b)


def measurements(name, spec, x):
    """pose, twist (v, vn, w), steering angle from the plant state"""
    nv = spec.nv
    a = x[3:3 + nv]
    if name == "diff":
        vel = ((a[1] + a[0]) / 2.0, 0.0, (a[1] - a[0]) / spec.p[0])
    elif name == "omni4":
        L = spec.p[0]
        vel = ((a[0] - a[1] + a[2] - a[3]) / 4.0, (-a[0] - a[1] + a[2] + a[3]) / 4.0, (-a[0] - a[1] - a[2] - a[3]) / (2.0 * L))
    else:
        # yaw rate of the tric model (scripts/tric/tric_amr_model.py, with its sin-for-cos term): v * sin(alpha) / d
        vel = (a[0], 0.0, None)
    return x[:3].copy(), vel, (a[1] if name == "tric" else 0.0)


class OracleRollout:
    def __init__(self, orc, name, segments, pose0, u0, back=0.05, ahead=0.5, holonomic=False):
        self.c = OracleController(orc, name)
        self.name, self.spec = name, self.c.spec
        self.seg = np.asarray(segments, dtype=np.float64).reshape(-1, 16)
        self.x = np.zeros(self.spec.nx); self.x[:3] = pose0
        self.pose, self.vel, self.steer = np.array(pose0, dtype=np.float64), (0.0, 0.0, 0.0), 0.0
        self.u, self.back, self.ahead, self.hol = float(u0), back, ahead, holonomic
        self.p = np.array(self.spec.p, dtype=np.float64)

    def step(self, noise=None, sqp_max_iter=1, sqp_tol=0.0, shift=False):
        """one tick; sqp_max_iter > 1 / shift: SQP to convergence and the warm-start shift of nmpc_rollout_device"""
        s = self.spec
        self.u = nearest_u(self.seg, self.u, self.pose[0], self.pose[1], self.back, self.ahead)
        refs = pathdisc.get_next_n_poses(self.seg, self.u, s.dt, s.n + 1, self.hol)
        cmd, qi = self.c.run(self.pose, self.vel, self.steer, [tuple(r) for r in refs], sqp_max_iter, sqp_tol)
        u0 = self.c.u[0] + (0.0 if noise is None else np.asarray(noise))
        self.x, _ = self.c.o.discrete_map(self.x, u0, self.p, s.dt)
        self.pose, vel, self.steer = measurements(self.name, s, self.x)
        if vel[2] is None:
            vel = (vel[0], 0.0, 0.0)          # run() does not read w for tric (Tric.cpp:96-98)
        self.vel = vel
        if shift:
            self.c.shift()
        return cmd, qi
