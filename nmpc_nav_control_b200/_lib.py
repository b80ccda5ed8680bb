"""ctypes binding of libnmpc_b200.so (the C ABI in include/nmpc_b200.h).

There is no CPU fallback: if the CUDA library is missing or no device is present, loading /
creating a solver raises."""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NMPC_B200_LIB") or os.path.join(PKG, "libnmpc_b200.so")   # override: kernel experiments (tools/gpu)

# every symbol include/nmpc_b200.h declares
SYMBOLS = [
    "nmpc_dims", "nmpc_default_opts", "nmpc_last_error", "nmpc_create", "nmpc_destroy",
    "nmpc_set_weights", "nmpc_set_bounds", "nmpc_set_params", "nmpc_get_tables", "nmpc_set_opts", "nmpc_get_opts",
    "nmpc_iterate_device", "nmpc_reset", "nmpc_reset_async", "nmpc_set_iterate_host", "nmpc_get_iterate_host",
    "nmpc_rti_solve_device", "nmpc_rti_solve_host", "nmpc_last_stats_host", "nmpc_last_timing", "nmpc_last_launches",
    "nmpc_dfma_peak_tflops",
    "nmpc_ctrl_tick_device", "nmpc_ctrl_reset", "nmpc_ctrl_state_device", "nmpc_ctrl_tick_host",
    "nmpc_path_discretize_device", "nmpc_plant_step_device", "nmpc_path_nearest_device",
    "nmpc_shift_device", "nmpc_sqp_solve_device", "nmpc_ctrl_tick_sqp_device", "nmpc_rollout_device",
]


class IpmOpts(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("mu0", "alpha_min", "res_g_max", "res_b_max", "res_d_max", "res_m_max",
                 "reg_prim", "lam_min", "t_min", "tau_min", "thr0")] + \
               [("iter_max", C.c_int), ("cond_pred_corr", C.c_int)]


class RolloutOpts(C.Structure):
    """nmpc_rollout_opts of include/nmpc_b200.h"""
    _fields_ = [("dt", C.c_double), ("back", C.c_double), ("ahead", C.c_double), ("sqp_tol", C.c_double),
                ("is_holonomic", C.c_int), ("sqp_max_iter", C.c_int), ("shift", C.c_int)]


class Dims(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("nx", "nu", "np", "ny", "nyn", "nbx", "nbu", "n")]


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m nmpc_nav_control_b200.build` "
                "(nvcc, sm_100a). This package has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for s in SYMBOLS:
            getattr(lib, s)
        lib.nmpc_last_error.restype = C.c_char_p
        lib.nmpc_dfma_peak_tflops.restype = C.c_double
        lib.nmpc_dfma_peak_tflops.argtypes = [C.c_int, C.c_int]
        lib.nmpc_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        lib.nmpc_destroy.argtypes = [C.c_void_p]
        vp, dp, ip = C.c_void_p, C.c_void_p, C.c_void_p
        lib.nmpc_set_weights.argtypes = [vp, dp, dp]
        lib.nmpc_set_bounds.argtypes = [vp, dp, dp, dp, dp]
        lib.nmpc_set_params.argtypes = [vp, dp]
        lib.nmpc_get_tables.argtypes = [vp, dp, dp, dp, dp, dp, dp, dp]
        lib.nmpc_last_stats_host.argtypes = [vp, C.c_int, dp]
        lib.nmpc_set_opts.argtypes = [vp, C.POINTER(IpmOpts)]
        lib.nmpc_get_opts.argtypes = [vp, C.POINTER(IpmOpts)]
        lib.nmpc_iterate_device.argtypes = [vp, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
        lib.nmpc_reset.argtypes = [vp]
        lib.nmpc_reset_async.argtypes = [vp, vp]
        lib.nmpc_set_iterate_host.argtypes = [vp, C.c_int, dp, dp]
        lib.nmpc_get_iterate_host.argtypes = [vp, C.c_int, dp, dp]
        lib.nmpc_rti_solve_device.argtypes = [vp, C.c_int, dp, dp, C.c_int, dp, dp, dp, C.c_int, ip, ip, dp, vp]
        lib.nmpc_rti_solve_host.argtypes = [vp, C.c_int, dp, dp, C.c_int, dp, dp, dp, ip, ip]
        lib.nmpc_last_timing.argtypes = [vp, C.POINTER(C.c_double)]
        lib.nmpc_last_launches.argtypes = [vp]
        lib.nmpc_dims.argtypes = [C.c_int, C.POINTER(Dims)]
        lib.nmpc_ctrl_tick_device.argtypes = [vp, C.c_int, dp, dp, dp, dp, ip, C.c_int, C.c_double, dp, ip, ip, vp]
        lib.nmpc_path_discretize_device.argtypes = [C.c_int, C.c_int, dp, ip, C.c_int, ip, dp, C.c_double, C.c_int, C.c_int, dp, vp]
        lib.nmpc_plant_step_device.argtypes = [vp, C.c_int, C.c_double, dp, dp, dp, dp, dp, vp]
        lib.nmpc_path_nearest_device.argtypes = [C.c_int, C.c_int, dp, ip, C.c_int, ip, dp, C.c_double, C.c_double, dp, vp]
        lib.nmpc_shift_device.argtypes = [vp, C.c_int, dp, dp, C.c_int, ip, vp]
        lib.nmpc_sqp_solve_device.argtypes = [vp, C.c_int, dp, dp, C.c_int, dp, dp, dp, C.c_int, C.c_int, C.c_double, ip, ip, ip, vp]
        lib.nmpc_ctrl_tick_sqp_device.argtypes = [vp, C.c_int, dp, dp, dp, dp, ip, C.c_int, C.c_double, C.c_int, C.c_double, dp, ip, ip, vp]
        lib.nmpc_rollout_device.argtypes = [vp, C.c_int, C.c_int, C.POINTER(RolloutOpts), dp, ip, C.c_int, ip, dp, dp, dp, dp, dp, dp, dp, dp, ip, vp]
        lib.nmpc_ctrl_reset.argtypes = [vp, vp]
        lib.nmpc_ctrl_state_device.argtypes = [vp, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
        lib.nmpc_ctrl_tick_host.argtypes = [vp, C.c_int, dp, dp, dp, dp, ip, C.c_int, C.c_double, dp, ip, ip]
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().nmpc_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed with code {rc}: {msg}")
