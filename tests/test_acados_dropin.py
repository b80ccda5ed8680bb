"""Drop-in boundary: the reference's UNMODIFIED C++ solver wrappers compiled against this
repository's acados-compatible headers / libraries (tests/conformance/build.sh) and driven tick by
tick like NMPCNavControlROS::executeNMPC does.

CPU part : headers exist, the wrapper sources compile and link (when /root/reference is present),
           every symbol the wrappers bind is exported.
GPU part : run the prebuilt driver and compare its commands with the oracle driven through the
           same wrapper protocol restated here (NMPCNavControlDiff.cpp:82-175)."""
import math
import os
import subprocess

import numpy as np
import pytest

from helpers import ATOL, RTOL
from nmpc_nav_control_b200.problem import MODELS

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = os.path.join(ROOT, "nmpc_nav_control_b200")
DRIVER = os.path.join(HERE, "conformance", "_build", "conformance_driver")
REF = "/root/reference"

ACADOS_SYMBOLS = ["ocp_nlp_constraints_model_set", "ocp_nlp_cost_model_set", "ocp_nlp_out_get", "ocp_nlp_get"]
SOLVER_SYMBOLS = ["acados_create_capsule", "acados_create", "acados_update_params", "acados_solve", "acados_reset",
                  "acados_free", "acados_free_capsule"]


def _exported(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", lib], check=True, capture_output=True, text=True).stdout
    return {ln.split()[-1] for ln in out.splitlines() if ln.strip()}


def test_shim_libraries_export_the_bound_symbols():
    from nmpc_nav_control_b200 import build
    build.build_all()
    sy = _exported(os.path.join(PKG, "libacados.so"))
    assert set(ACADOS_SYMBOLS) <= sy
    for m in ("diff2amr", "omni4amr", "tric3amr"):
        sy = _exported(os.path.join(PKG, f"libacados_ocp_solver_{m}.so"))
        assert {f"{m}_{s}" for s in SOLVER_SYMBOLS} <= sy, m


def test_headers_on_include_path_and_macros():
    inc = os.path.join(ROOT, "include")
    for h in ("acados/utils/print.h", "acados_c/ocp_nlp_interface.h", "acados_c/external_function_interface.h",
              "acados/ocp_nlp/ocp_nlp_constraints_bgh.h", "acados/ocp_nlp/ocp_nlp_cost_ls.h",
              "blasfeo/include/blasfeo_d_aux.h", "blasfeo/include/blasfeo_d_aux_ext_dep.h"):
        assert os.path.exists(os.path.join(inc, h)), h
    for spec in MODELS.values():
        txt = open(os.path.join(inc, f"acados_solver_{spec.acados_name}.h")).read()
        M = spec.acados_name.upper()
        for key, val in (("NX", spec.nx), ("NU", spec.nu), ("NY", spec.ny), ("NYN", spec.nx),
                         ("NP", spec.np_), ("NBX", spec.nbx), ("NBU", spec.nbu)):
            assert f"#define {M}_{key} {val}\n" in txt, (M, key)
        # the horizon macro follows the emitted header (include/nmpc_horizon.h): evaluate it with the preprocessor
        r = subprocess.run(["g++", "-x", "c++", "-E", "-P", "-I", inc, "-"], input=f'#include "acados_solver_{spec.acados_name}.h"\nHORIZON={M}_N\n',
                           capture_output=True, text=True)
        assert r.returncode == 0 and f"HORIZON={spec.n}" in r.stdout.replace(" ", ""), (M, r.stderr[-500:])


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")
def test_reference_wrappers_compile_and_link_unchanged():
    from nmpc_nav_control_b200 import build
    build.build_all()
    r = subprocess.run(["sh", os.path.join(HERE, "conformance", "build.sh")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    assert os.path.exists(DRIVER)


from oracle.ctrl import OracleController  # noqa: E402  (the wrapper protocol restated around the oracle)


def _scenario(name, n_ticks, seed):
    """pose/vel per tick and a reference list per tick (shorter than N+1 on some ticks: exercises padding)"""
    rng = np.random.default_rng(seed)
    pose = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-3, 3)])
    ticks = []
    speed = rng.uniform(0.3, 0.7)
    kap = rng.uniform(-1.0, 1.0)
    for t in range(n_ticks):
        vel = (speed * min(1.0, 0.1 * t), 0.0 if name != "omni4" else 0.05, 0.1 * kap)
        steer = 0.05 * kap
        n_ref = 81 if t % 3 else 60
        refs = []
        for i in range(n_ref):
            s = speed * 0.025 * i
            th = pose[2] + kap * s
            th = math.atan2(math.sin(th), math.cos(th))
            refs.append((pose[0] + s * math.cos(pose[2] + 0.5 * kap * s), pose[1] + s * math.sin(pose[2] + 0.5 * kap * s), th))
        ticks.append((pose.copy(), vel, steer, refs))
        # the robot advances along the path
        s = speed * 0.025
        pose = np.array([pose[0] + s * math.cos(pose[2]), pose[1] + s * math.sin(pose[2]), pose[2] + kap * s])
    return ticks


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["diff", "omni4", "tric"])
def test_reference_wrapper_closed_loop_matches_oracle(oracle_mod, name, tmp_path):
    if not os.path.exists(DRIVER):
        pytest.skip("conformance driver not built (tests/conformance/build.sh needs the reference tree)")
    ticks = _scenario(name, 12, seed=11)
    inp = tmp_path / "in.txt"
    # the driver reads a fixed number of reference poses per tick; shorter lists are padded here with
    # their last pose, which is exactly what run() does itself (Diff.cpp:113-117)
    n_ref = 81
    with open(inp, "w") as f:
        f.write(f"{len(ticks)} {n_ref}\n")
        for pose, vel, steer, refs in ticks:
            f.write(" ".join(repr(float(v)) for v in (*pose, *vel, steer)) + "\n")
            padded = list(refs) + [refs[-1]] * (n_ref - len(refs))
            for p in padded:
                f.write(" ".join(repr(float(v)) for v in p) + "\n")
    r = subprocess.run([DRIVER, name, str(inp)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("tick")]
    assert len(lines) == len(ticks)
    ctl = OracleController(oracle_mod, name)
    for t, (pose, vel, steer, refs) in enumerate(ticks):
        padded = list(refs) + [refs[-1]] * (n_ref - len(refs))
        cmd, _ = ctl.run(pose, vel, steer, padded)
        parts = lines[t].split()
        assert parts[2] == "ok=1", lines[t]
        got = [float(v) for v in lines[t].split("cmd=")[1].split()]
        cpu_ms = float(parts[3].split("=")[1])
        assert cpu_ms > 0.0
        for a, b in zip(got, cmd):
            assert abs(a - b) <= ATOL + RTOL * abs(b), (t, got, cmd)
