"""TEST INFRASTRUCTURE - the reference wrappers' per-tick protocol restated around the oracle.

`OracleController.run` follows NMPCNavControl{Diff,Omni4,Tric}::run line by line
(src/nmpc_nav_control/NMPCNavControlDiff.cpp:82-175, NMPCNavControlOmni4.cpp:91-177, NMPCNavControlTric.cpp:88-181;
unwrapAngle: NMPCNavControl.cpp:26-32; kinematics: Diff.cpp:183-193, Omni4.cpp:185-200), one robot at a time in
plain Python, with the oracle's RTI step in place of `{m}_acados_solve`.

Pinned against the reference itself: tests/test_acados_dropin.py runs the reference's UNMODIFIED wrapper sources
(compiled by tests/conformance/build.sh against this repository's acados-compatible libraries) tick by tick and
compares their commands with this class.  Only tests/ and __graft_entry__.smoke() may import it; the product's
glue is nmpc_nav_control_b200/csrc/ctrl_glue.cuh."""
import math

import numpy as np

from nmpc_nav_control_b200.problem import MODELS


def _unwrap(cur, prev):
    d = cur - prev
    if d > math.pi:
        cur -= 2 * math.pi
    elif d < -math.pi:
        cur += 2 * math.pi
    return cur


class OracleController:
    """NMPCNavControl{Diff,Omni4,Tric}::run restated (Diff.cpp:82-175, Omni4.cpp:91-177, Tric.cpp:88-181)"""

    def __init__(self, orc, name, dt=None):
        self.spec = MODELS[name]
        s = self.spec
        tb = s.codegen_defaults()
        # the wrapper constructors set W_e from W_diag[0..nx-1] = Q, not QN (Diff.cpp:34-41)
        tb["We"] = np.array(s.Q, dtype=np.float64)
        self.tb = tb
        self.orc = orc
        self.name = name
        self.o = orc.Oracle(name, tb) if orc is not None else None
        self.x0 = np.zeros(s.nx)                                               # Diff.cpp:14
        self.x = np.zeros((s.n + 1, s.nx)); self.u = np.zeros((s.n, s.nu))     # after reset_mpc()
        self.dt = s.dt if dt is None else dt

    def reset_mpc(self):
        s = self.spec
        self.x = np.zeros((s.n + 1, s.nx)); self.u = np.zeros((s.n, s.nu))

    def pre(self, pose, vel, steer, refs):
        """everything run() does before the solve: x0 (in place, keeps the carried reference states), yref, W_e"""
        s = self.spec
        x0 = self.x0
        x0[0:3] = pose
        v, vn, w = vel
        if self.name == "diff":
            b = s.p[0]                                   # Diff.cpp:183-187
            x0[3] = v - 0.5 * b * w; x0[4] = v + 0.5 * b * w
        elif self.name == "omni4":
            L = s.p[0]                                   # Omni4.cpp:185-191
            x0[3] = v - vn - 0.5 * L * w; x0[4] = -v - vn - 0.5 * L * w
            x0[5] = v + vn - 0.5 * L * w; x0[6] = -v + vn - 0.5 * L * w
        else:
            x0[3] = v; x0[4] = steer                     # Tric.cpp:96-97
        yref = np.zeros((s.n + 1, s.ny))
        prev = pose[2]
        for i in range(s.n + 1):                         # Diff.cpp:103-118
            if i < len(refs):
                yref[i, 0], yref[i, 1] = refs[i][0], refs[i][1]
                yref[i, 2] = _unwrap(refs[i][2], prev)
                prev = yref[i, 2]
            else:
                yref[i, :3] = yref[i - 1, :3]
        We = None
        if self.name == "diff":                          # terminal-weight switch, Diff.cpp:127-139
            We = np.array(s.Q, dtype=np.float64)
            if (yref[s.n, :3] == yref[s.n - 1, :3]).all():
                We[:3] = 100.0 * np.array(s.Q[:3])
        return x0, yref, We

    def post(self, x0, u0):
        """after the solve: reference states advanced by u_0 dt, inverse kinematics (Diff.cpp:155-166)"""
        s = self.spec
        nv = s.nv
        new_ref = x0[3 + nv:3 + 2 * nv] + np.asarray(u0) * self.dt
        if self.name == "diff":
            cmd = ((new_ref[1] + new_ref[0]) / 2.0, (new_ref[1] - new_ref[0]) / s.p[0], 0.0)
        elif self.name == "omni4":
            L = s.p[0]                                   # Omni4.cpp:193-200
            v1, v2, v3, v4 = new_ref
            cmd = ((v1 - v2 + v3 - v4) / 4.0, (-v1 - v2 + v3 + v4) / 4.0, (-v1 - v2 - v3 - v4) / (2.0 * L))
        else:
            cmd = (new_ref[0], new_ref[1], 0.0)
        return cmd, new_ref

    def run(self, pose, vel, steer, refs, sqp_max_iter=1, sqp_tol=0.0):
        """sqp_max_iter = 1: the reference's single RTI step per tick.  > 1 (BASELINE config 4, not a reference behaviour):
        RTI steps from the current iterate until the inf-norm of the step is <= sqp_tol or sqp_max_iter steps were taken;
        returns the QP iterations summed over the steps and keeps the number of steps in self.sqp_steps"""
        s = self.spec
        x0, yref, We = self.pre(pose, vel, steer, refs)
        qp_total = 0
        self.sqp_steps = 0
        for _ in range(max(1, sqp_max_iter)):
            r = self.o.rti(x0, yref, self.x, self.u, We=We)
            assert r["status"] == 0                      # processAcadosStatus throws, NMPCNavControl.cpp:15-24
            step = max(np.abs(r["x"] - self.x).max(), np.abs(r["u"] - self.u).max())
            self.x, self.u = r["x"], r["u"]
            qp_total += r["qp_iter"]
            self.sqp_steps += 1
            if sqp_max_iter <= 1 or step <= sqp_tol:
                break
        r = dict(qp_iter=qp_total)
        cmd, new_ref = self.post(x0, self.u[0])
        nv = s.nv
        self.x0 = self.x[1].copy()                       # Diff.cpp:168-172
        self.x0[3 + nv:3 + 2 * nv] = new_ref
        return cmd, r["qp_iter"]

    def shift(self):
        """warm-start shift of the iterate: stage k takes stage k+1, the last stage is kept"""
        self.x[:-1] = self.x[1:].copy()
        self.u[:-1] = self.u[1:].copy()
