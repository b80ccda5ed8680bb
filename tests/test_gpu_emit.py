"""SURVEY.md 8(f4) on the GPU: a library emitted and built for another horizon (YAML with tf_ini = 1.0 -> N = 40, the
counterpart of re-running scripts/generate_acados_libs.py on an edited config/nmpc_nav_control_acados_models.yaml) solves
the N = 40 problems in agreement with the oracle built for that horizon.  Runs in a fresh interpreter because the
horizon is a build-time constant of both the library and the package (include/nmpc_horizon.h)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import yaml

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_emitted_n40_library_matches_the_oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_emit_cpu import _yaml_of_defaults
    from nmpc_nav_control_b200 import emit
    alt = os.path.join(ROOT, "build", "alt_n40")
    os.makedirs(alt, exist_ok=True)
    cfg = _yaml_of_defaults()
    for k in cfg:
        cfg[k]["tf_ini"] = 1.0
    y = os.path.join(alt, "n40.yaml")
    with open(y, "w") as f:
        f.write(yaml.safe_dump(cfg))
    lib = os.path.join(alt, "libnmpc_b200.so")
    srcs = [os.path.join(ROOT, "nmpc_nav_control_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "nmpc_nav_control_b200", "csrc"))]
    fresh = os.path.exists(lib) and all(os.path.getmtime(s) <= os.path.getmtime(lib) for s in srcs)
    assert emit.main([y, "--out-dir", alt] + ([] if fresh else ["--build"])) == 0
    hdr = os.path.join(alt, "include", "nmpc_horizon.h")
    code = r"""
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np, torch
from nmpc_nav_control_b200.problem import MODELS, N_HORIZON
from nmpc_nav_control_b200.solver import BatchedRtiSolver
import helpers
from oracle import orc
assert N_HORIZON == 40 and orc.ORC_N == 40
for name, B in (("diff", 300), ("tric", 200), ("omni4", 100)):
    spec, x0, yref, _ = helpers.instances(name, 10, B)
    ref = helpers.oracle_solve(orc, name, x0, yref)
    s = BatchedRtiSolver(spec, B)
    assert s.lib.nmpc_dims  # the alternate library
    s.reset()
    out = s.solve_host(x0, yref)
    x, u = s.get_iterate(B)
    assert x.shape[1] == 41
    assert (out["status"] == 0).all() and (out["qp_iter"] == ref["qp_iter"]).all(), name
    assert helpers.parity_report(x, ref["x"])[0] == 0 and helpers.parity_report(u, ref["u"])[0] == 0, name
    s.close()
print("n40 gpu parity ok")
""" % (ROOT, os.path.join(ROOT, "tests"))
    env = dict(os.environ, NMPC_HORIZON_H=hdr, NMPC_B200_LIB=lib)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=1200)
    assert r.returncode == 0 and "n40 gpu parity ok" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])
