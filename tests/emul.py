"""TEST INFRASTRUCTURE: ctypes loader of tests/host_emul/libemul.so, the g++ build of the per-lane
CUDA solver logic (nmpc_nav_control_b200/csrc/rti_core.cuh with -DNMPC_HOST_EMUL).  It lets the
kernel arithmetic be compared with the oracle in the GPU-less container.  The product cannot
reach it: it lives under tests/ and nothing in the package imports it."""
import ctypes as C
import os
import subprocess

import numpy as np

from nmpc_nav_control_b200 import _lib
from nmpc_nav_control_b200.problem import DT, MODELS, N_HORIZON

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emul")
_SRC = os.path.join(_HERE, "emul.cpp")
# the horizon is a build parameter (include/nmpc_horizon.h, emitted): an alternate horizon gets its own emulation library
_SO = os.path.join(_HERE, "libemul.so" if N_HORIZON == 80 else f"libemul_n{N_HORIZON}.so")
_CSRC = os.path.join(os.path.dirname(_HERE), "..", "nmpc_nav_control_b200", "csrc")


def build(force=False):
    deps = [_SRC] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".cuh")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-DNMPC_HOST_EMUL", f"-DNMPC_N={N_HORIZON}", f"-DNMPC_DT={DT!r}",
                        "-fPIC", "-shared", "-o", _SO, _SRC], check=True, capture_output=True)
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


def emul_rti(name, x0, yref, x=None, u=None, We=None, tables=None, opts=None, hybrid=None, coop=False, solo=False):
    """one RTI step of B instances through the emulated kernel logic; returns dict like the oracle helper.
    coop = False: the per-lane K3 of rti_core.cuh (the lockstep sweeps); coop = True: the persistent lane-cooperative K3 of
    rti_coop.cuh (G = 4 nv lanes per instance); hybrid = K: K iterations of the per-lane sweeps, then hand-over of the
    unfinished instances to the lane-cooperative kernel (returns the number handed over as out["resumed"])"""
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
    fn, lead = (lib.emul_rti_coop, (C.c_int(0),)) if coop else (lib.emul_rti, ())
    if hybrid is not None:
        fn, lead = lib.emul_rti_hybrid, (C.c_int(hybrid),)
    if solo:          # the block-per-instance mapping of rti_solo.cuh
        fn, lead = lib.emul_rti_solo, ()
    rc = fn(C.c_int(spec.model_id), *lead, C.c_int(B), _dp(arrs["W"]), _dp(arrs["We"]), _dp(arrs["lbx"]), _dp(arrs["ubx"]),
                      _dp(arrs["lbu"]), _dp(arrs["ubu"]), _dp(arrs["p"]), C.c_double(tb["dt"]), C.byref(o),
                      _dp(x0), _dp(yref), C.c_int(nyref), None if we is None else _dp(we), _dp(x), _dp(u),
                      status.ctypes.data_as(C.POINTER(C.c_int)), iters.ctypes.data_as(C.POINTER(C.c_int)), _dp(stats))
    assert rc >= 0
    return dict(x=x, u=u, qp_status=status, qp_iter=iters, stats=stats, resumed=rc)


def emul_ctrl_pre(name, pose, vel, steer, refs, nref, vref, tables):
    """controller glue before the solve (ctrl_glue.cuh CtrlGlue::pre) on SoA arrays: pose [3,B], vel [3,B], steer [B] or
    None, refs [nref_max,3,B], nref [B] or None, vref [nv,B]; returns x0bar [nx,B], yref [N+1,3,B], We [nx,B] (diff) or None"""
    lib = build()
    spec = MODELS[name]
    B = pose.shape[1]
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    pose, vel, refs, vref = c(pose), c(vel), c(refs), c(vref)
    steer = None if steer is None else c(steer)
    nref_a = None if nref is None else np.ascontiguousarray(nref, dtype=np.int32)
    x0bar = np.zeros((spec.nx, B)); yref = np.zeros((spec.n + 1, 3, B))
    We = np.zeros((spec.nx, B)) if name == "diff" else None
    p, W0, Wt = c(tables["p"][0]), c(tables["W"][0]), c(tables["We"])
    lib.emul_ctrl_pre(C.c_int(spec.model_id), C.c_int(B), _dp(pose), _dp(vel), None if steer is None else _dp(steer), _dp(refs),
                      None if nref_a is None else nref_a.ctypes.data_as(C.POINTER(C.c_int)), C.c_int(refs.shape[0]), _dp(vref),
                      _dp(p), _dp(W0), _dp(Wt), _dp(x0bar), _dp(yref), None if We is None else _dp(We))
    return x0bar, yref, We


def emul_ctrl_post(name, status, x0bar, u0, dt, vref, cmd, tables):
    """controller glue after the solve (CtrlGlue::post): status [B], x0bar [nx,B], u0 [nu,B]; updates and returns vref [nv,B],
    cmd [3,B] (instances with a non-zero status keep both)"""
    lib = build()
    spec = MODELS[name]
    B = x0bar.shape[1]
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    st = np.ascontiguousarray(status, dtype=np.int32)
    vref, cmd = c(vref).copy(), c(cmd).copy()
    lib.emul_ctrl_post(C.c_int(spec.model_id), C.c_int(B), st.ctypes.data_as(C.POINTER(C.c_int)), _dp(c(x0bar)), _dp(c(u0)),
                       C.c_double(dt), _dp(c(tables["p"][0])), _dp(vref), _dp(cmd))
    return vref, cmd


def emul_path_discretize(segments, offsets, path_id, u0, period, num_poses, holonomic=False):
    """path discretisation (path_disc.cuh PathDisc::next_poses): returns poses [num_poses, 3, B]"""
    lib = build()
    seg = np.ascontiguousarray(segments, dtype=np.float64); off = np.ascontiguousarray(offsets, dtype=np.int32)
    pid = np.ascontiguousarray(path_id, dtype=np.int32); u0 = np.ascontiguousarray(u0, dtype=np.float64)
    B = len(u0)
    out = np.zeros((num_poses, 3, B))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    lib.emul_path_discretize(C.c_int(B), _dp(seg), ip(off), C.c_int(len(off) - 1), ip(pid), _dp(u0), C.c_double(period),
                             C.c_int(num_poses), C.c_int(int(holonomic)), _dp(out))
    return out


def emul_plant_step(name, x, u0, noise, p, dt):
    """plant step (rollout.cuh Rollout::plant_step) on SoA arrays x [nx,B], u0 [nu,B], noise [nu,B] or None;
    returns x+ [nx,B], pose [3,B], vel [3,B], steer [B]"""
    lib = build()
    spec = MODELS[name]
    B = x.shape[1]
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    xp = c(x).copy(); pose = np.zeros((3, B)); vel = np.zeros((3, B)); steer = np.zeros(B)
    lib.emul_plant_step(C.c_int(spec.model_id), C.c_int(B), _dp(xp), _dp(c(u0)), None if noise is None else _dp(c(noise)),
                        _dp(c(p)), C.c_double(dt), _dp(pose), _dp(vel), _dp(steer))
    return xp, pose, vel, steer


def emul_nearest_u(segments, u_prev, px, py, back, ahead):
    lib = build()
    lib.emul_nearest_u.restype = C.c_double
    seg = np.ascontiguousarray(segments, dtype=np.float64).reshape(-1, 16)
    return lib.emul_nearest_u(_dp(seg), C.c_int(len(seg)), C.c_double(u_prev), C.c_double(px), C.c_double(py), C.c_double(back),
                              C.c_double(ahead))
