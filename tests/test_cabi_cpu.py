"""The C-ABI library loads here (no GPU) and exports every symbol include/nmpc_b200.h declares;
without a CUDA device the product path fails loudly instead of falling back to a CPU solver."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "nmpc_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nmpc_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from nmpc_nav_control_b200 import _lib, build
    build.build_core()
    names = _declared()
    assert len(names) >= 20
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], check=True, capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    missing = [n for n in names if n not in exported]
    assert not missing, missing
    assert sorted(_lib.SYMBOLS) == names          # the ctypes binding covers the whole header
    _lib.load()


def test_dims_and_default_options_without_device():
    from nmpc_nav_control_b200 import _lib
    from nmpc_nav_control_b200.problem import MODELS
    lib = _lib.load()
    for spec in MODELS.values():
        d = _lib.Dims()
        assert lib.nmpc_dims(spec.model_id, C.byref(d)) == 0
        assert (d.nx, d.nu, d.np, d.ny, d.nyn, d.nbx, d.nbu, d.n) == (spec.nx, spec.nu, spec.np_, spec.ny, spec.nx, spec.nbx, spec.nbu, 80)
    assert lib.nmpc_dims(7, C.byref(_lib.Dims())) == -1
    o = _lib.IpmOpts()
    lib.nmpc_default_opts(C.byref(o))
    # SURVEY.md Appendix B.4
    assert (o.mu0, o.alpha_min, o.res_g_max, o.res_b_max, o.res_d_max, o.res_m_max) == (1.0, 1e-8, 1e-6, 1e-8, 1e-8, 1e-8)
    assert (o.reg_prim, o.lam_min, o.t_min, o.tau_min, o.thr0, o.iter_max, o.cond_pred_corr) == (1e-15, 1e-16, 1e-16, 1e-16, 0.1, 50, 1)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_no_cpu_fallback():
    from nmpc_nav_control_b200 import _lib
    from nmpc_nav_control_b200.solver import BatchedRtiSolver
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.nmpc_create(0, 8, 0, C.byref(h))
    assert rc == -3 and b"no CPU fallback" in lib.nmpc_last_error()
    with pytest.raises(RuntimeError):
        BatchedRtiSolver("diff", 8)
    # the acados-compatible create reports failure the way the wrapper expects (non-zero status)
    shim = C.CDLL(os.path.join(ROOT, "nmpc_nav_control_b200", "libacados_ocp_solver_diff2amr.so"))
    shim.diff2amr_acados_create_capsule.restype = C.c_void_p
    cap = shim.diff2amr_acados_create_capsule()
    assert cap
    assert shim.diff2amr_acados_create(C.c_void_p(cap)) != 0
    shim.diff2amr_acados_free_capsule(C.c_void_p(cap))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_no_cpu_fallback_for_the_stateless_entry_points():
    """the path discretiser and the nearest-parameter search refuse to run without a device, and the
    host mirrors raise instead of computing anything on the CPU"""
    import numpy as np
    from nmpc_nav_control_b200 import _lib
    lib = _lib.load()
    a = np.zeros(64); ai = np.zeros(4, dtype=np.int32)
    p, pi = C.c_void_p(a.ctypes.data), C.c_void_p(ai.ctypes.data)
    assert lib.nmpc_path_discretize_device(0, 1, p, pi, 1, pi, p, 0.025, 4, 0, p, None) == -3
    assert b"no CPU fallback" in lib.nmpc_last_error()
    assert lib.nmpc_path_nearest_device(0, 1, p, pi, 1, pi, p, 0.05, 0.5, p, None) == -3
    from nmpc_nav_control_b200.controller import BatchedNavController
    with pytest.raises(RuntimeError):
        BatchedNavController("diff", 4, dt=0.025)


def test_cpp_example_compiles_links_and_fails_loudly_without_a_device(tmp_path):
    """examples/fleet_tick.cpp: a plain C++14 caller of include/nmpc_b200.h builds against the library; without a CUDA
    device it reports the library's error and exits 2 (no CPU fallback)"""
    from nmpc_nav_control_b200 import build
    build.build_core()
    exe = str(tmp_path / "fleet_tick")
    pkg = os.path.join(ROOT, "nmpc_nav_control_b200")
    r = subprocess.run(["g++", "-std=c++14", "-O2", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "fleet_tick.cpp"), "-L", pkg, "-lnmpc_b200", f"-Wl,-rpath,{pkg}", "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    if not torch.cuda.is_available():
        p = subprocess.run([exe, "8"], capture_output=True, text=True, timeout=120)
        assert p.returncode == 2 and "no CPU fallback" in p.stderr, (p.returncode, p.stderr)


def test_cpp_rollout_example_compiles_and_links(tmp_path):
    """examples/fleet_rollout.cpp: BASELINE config 4 (tric, SQP + warm-start shift, closed loop) through ONE call of
    nmpc_rollout_device from plain C++; without a CUDA device it exits 2 with the library's message"""
    from nmpc_nav_control_b200 import build
    build.build_core()
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("CUDA runtime headers not found")
    exe = str(tmp_path / "fleet_rollout")
    pkg = os.path.join(ROOT, "nmpc_nav_control_b200")
    r = subprocess.run(["g++", "-std=c++14", "-O2", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
                        os.path.join(ROOT, "examples", "fleet_rollout.cpp"), "-L", pkg, "-lnmpc_b200", "-L", os.path.join(cuda, "lib64"),
                        "-lcudart", f"-Wl,-rpath,{pkg}", f"-Wl,-rpath,{os.path.join(cuda, 'lib64')}", "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    if not torch.cuda.is_available():
        p = subprocess.run([exe, "8", "3"], capture_output=True, text=True, timeout=120)
        assert p.returncode == 2 and "no CPU fallback" in p.stderr, (p.returncode, p.stderr)
