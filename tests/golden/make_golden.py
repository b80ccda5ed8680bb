"""Generates tests/golden/rti_<model>.npz: seeded inputs (SURVEY.md Appendix D generator) and the
ORACLE's outputs for two consecutive RTI steps.  The reference has no golden vectors and real acados
cannot be run here, so these pin the restatement, not acados.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from helpers import instances          # noqa: E402
from nmpc_nav_control_b200.problem import MODELS   # noqa: E402
from oracle import orc                 # noqa: E402

for name, start, count in (("diff", 1000, 24), ("omni4", 2000, 12), ("tric", 3000, 24)):
    spec, x0, yref, _ = instances(name, start, count)
    o = orc.Oracle(name, spec.codegen_defaults())
    x = np.zeros((count, spec.n + 1, spec.nx)); u = np.zeros((count, spec.n, spec.nu))
    r = o.rti_batch(x0, yref, x, u)
    x1, u1 = x.copy(), u.copy()
    x0b = x[:, 1].copy()
    r2 = o.rti_batch(x0b, yref, x, u)
    np.savez_compressed(os.path.join(HERE, f"rti_{name}.npz"), start=start, x0=x0, yref=yref, x=x1, u=u1,
                        qp_iter=r["qp_iter"], status=r["status"], x2=x, u2=u, qp_iter2=r2["qp_iter"], lin_res2=r2["lin_res"])
    print(name, "iters", r["qp_iter"], r2["qp_iter"])
