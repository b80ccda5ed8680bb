"""TEST INFRASTRUCTURE: ctypes loader of tests/host_emul/libemul.so, the g++ build of the per-lane
CUDA solver logic (nmpc_nav_control_b200/csrc/rti_core.cuh with -DNMPC_HOST_EMUL).  It lets the
kernel arithmetic be compared with the oracle in the GPU-less container.  The product cannot
reach it: it lives under tests/ and nothing in the package imports it."""
import ctypes as C
import os
import subprocess

import numpy as np

from nmpc_nav_control_b200 import _lib
from nmpc_nav_control_b200.problem import MODELS

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emul")
_SRC = os.path.join(_HERE, "emul.cpp")
_SO = os.path.join(_HERE, "libemul.so")
_CSRC = os.path.join(os.path.dirname(_HERE), "..", "nmpc_nav_control_b200", "csrc")


def build(force=False):
    deps = [_SRC] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".cuh")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-DNMPC_HOST_EMUL", "-fPIC", "-shared",
                        "-o", _SO, _SRC], check=True, capture_output=True)
    return C.CDLL(_SO)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def default_opts():
    o = _lib.IpmOpts()
    o.mu0, o.alpha_min = 1.0, 1e-8
    o.res_g_max, o.res_b_max, o.res_d_max, o.res_m_max = 1e-6, 1e-8, 1e-8, 1e-8
    o.reg_prim, o.lam_min, o.t_min, o.tau_min, o.thr0 = 1e-15, 1e-16, 1e-16, 1e-16, 0.1
    o.iter_max, o.cond_pred_corr = 50, 1
    return o


def emul_rti(name, x0, yref, x=None, u=None, We=None, tables=None, opts=None, group=0, hybrid=None):
    """one RTI step of B instances through the emulated kernel logic; returns dict like the oracle helper.
    group = 0: the per-lane K3 of rti_core.cuh; group = G: the lane-group K3 of rti_group.cuh with G lanes per instance;
    hybrid = K: K iterations of the per-lane sweeps, then hand-over of the unfinished instances to the lane-group kernel
    (returns the number handed over as out["resumed"])"""
    lib = build()
    spec = MODELS[name]
    tb = tables or spec.codegen_defaults()
    B = x0.shape[0]
    nyref = yref.shape[2]
    x = np.zeros((B, spec.n + 1, spec.nx)) if x is None else np.array(x, dtype=np.float64, order="C")
    u = np.zeros((B, spec.n, spec.nu)) if u is None else np.array(u, dtype=np.float64, order="C")
    status = np.zeros(B, dtype=np.int32); iters = np.zeros(B, dtype=np.int32)
    stats = np.zeros((B, 8))
    arrs = {k: np.ascontiguousarray(tb[k], dtype=np.float64) for k in ("W", "We", "lbx", "ubx", "lbu", "ubu", "p")}
    o = opts or default_opts()
    x0 = np.ascontiguousarray(x0, dtype=np.float64); yref = np.ascontiguousarray(yref, dtype=np.float64)
    we = None if We is None else np.ascontiguousarray(We, dtype=np.float64)
    fn, lead = (lib.emul_rti, ()) if not group else (lib.emul_rti_group, (C.c_int(group),))
    if hybrid is not None:
        fn, lead = lib.emul_rti_hybrid, (C.c_int(hybrid),)
    rc = fn(C.c_int(spec.model_id), *lead, C.c_int(B), _dp(arrs["W"]), _dp(arrs["We"]), _dp(arrs["lbx"]), _dp(arrs["ubx"]),
                      _dp(arrs["lbu"]), _dp(arrs["ubu"]), _dp(arrs["p"]), C.c_double(tb["dt"]), C.byref(o),
                      _dp(x0), _dp(yref), C.c_int(nyref), None if we is None else _dp(we), _dp(x), _dp(u),
                      status.ctypes.data_as(C.POINTER(C.c_int)), iters.ctypes.data_as(C.POINTER(C.c_int)), _dp(stats))
    assert rc >= 0
    return dict(x=x, u=u, qp_status=status, qp_iter=iters, stats=stats, resumed=rc)
