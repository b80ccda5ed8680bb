"""ORACLE (test infrastructure, NOT product code) — ctypes front-end of oracle/liboracle*.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs
import this module.  PARITY UNPINNED: the oracle restates the acados SQP_RTI algorithm from
the reference's model files and SURVEY.md Appendix B; real acados is not available here and
the reference holds no golden vectors.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORC_HIST = 64


def _horizon() -> int:
    """the horizon the checked build was emitted for (nmpc_nav_control_b200.problem reads include/nmpc_horizon.h)"""
    from nmpc_nav_control_b200.problem import N_HORIZON
    return int(N_HORIZON)


ORC_N = _horizon()

_DIMS = {  # name -> (nx, nu, np, nbx, nbu)
    "diff": (7, 2, 2, 2, 2),
    "omni4": (11, 4, 2, 4, 4),
    "tric": (7, 2, 3, 2, 2),
}


class IpmOpts(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("mu0", "alpha_min", "res_g_max", "res_b_max", "res_d_max", "res_m_max",
                 "reg_prim", "lam_min", "t_min", "tau_min", "thr0")] + \
               [(n, C.c_int) for n in ("iter_max", "pred_corr", "cond_pred_corr", "itref_corr_max", "lq_fact")]


class Stats(C.Structure):
    _fields_ = [("status", C.c_int), ("qp_status", C.c_int), ("qp_iter", C.c_int),
                ("itref_solves", C.c_int), ("lq_flags", C.c_int), ("cond_fallbacks", C.c_int),
                ("res", C.c_double * 4), ("mu", C.c_double), ("lin_res_max", C.c_double * 4),
                ("alpha_hist", C.c_double * ORC_HIST), ("mu_hist", C.c_double * ORC_HIST)]


def _names():
    """library file names: the default horizon keeps the plain names, another horizon (an alternate emitted build) gets its own"""
    sfx = "" if ORC_N == 80 else f"_n{ORC_N}"
    return f"liboracle{sfx}.so", f"liboracle_fast{sfx}.so"


def build(force: bool = False) -> None:
    """compile the oracle with the Makefile next to this file (gcc only)"""
    ref, fast = _names()
    need = force or not all(os.path.exists(os.path.join(_HERE, f)) for f in (ref, fast))
    if not need:
        srcs = [os.path.join(_HERE, f) for f in ("orc_api.c", "orc_common.h", "orc_models.h", "orc_rti_core.inc")]
        newest = max(os.path.getmtime(s) for s in srcs)
        need = any(os.path.getmtime(os.path.join(_HERE, f)) < newest for f in (ref, fast))
    if need:
        subprocess.run(["make", "-C", _HERE, "-B", "all", f"ORC_N={ORC_N}", f"REFLIB={ref}", f"FASTLIB={fast}"], check=True, capture_output=True)


_libs = {}


def _lib(fast: bool):
    key = "fast" if fast else "ref"
    if key not in _libs:
        build()
        _libs[key] = C.CDLL(os.path.join(_HERE, _names()[1] if fast else _names()[0]))
    return _libs[key]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def default_opts(tight: bool = False) -> IpmOpts:
    o = IpmOpts()
    _lib(False).orc_default_opts(C.byref(o))
    if tight:   # 'tight' mode of SURVEY.md §7 hard part 1: unique QP optimum, solver-independent
        o.res_g_max = o.res_b_max = o.res_d_max = o.res_m_max = 1e-12
        o.iter_max = 200
    return o


def max_threads() -> int:
    return int(_lib(True).orc_max_threads())


class Oracle:
    """CPU restatement of one model's SQP_RTI step."""

    def __init__(self, model: str, tables: dict, fast: bool = False):
        self.model = model
        self.nx, self.nu, self.np_, self.nbx, self.nbu = _DIMS[model]
        self.ny = self.nx + self.nu
        self.nz = self.ny
        self.nb = self.nbx + self.nbu
        self.lib = _lib(fast)
        self.pfx = f"orc_{model}_"
        n = ORC_N
        # pack the prob struct: dt, W[N][NY], We[NX], lbx, ubx [N][NBX], lbu, ubu [N][NBU], p[N][NP]
        parts = [np.array([tables["dt"]], dtype=np.float64)]
        for key, shape in (("W", (n, self.ny)), ("We", (self.nx,)), ("lbx", (n, self.nbx)), ("ubx", (n, self.nbx)),
                           ("lbu", (n, self.nbu)), ("ubu", (n, self.nbu)), ("p", (n, self.np_))):
            a = np.ascontiguousarray(tables[key], dtype=np.float64)
            assert a.shape == shape, (key, a.shape, shape)
            parts.append(a.ravel())
        self.prob = np.concatenate(parts)
        f = getattr(self.lib, self.pfx + "prob_size"); f.restype = C.c_size_t
        assert f() == self.prob.nbytes, (f(), self.prob.nbytes)
        f = getattr(self.lib, self.pfx + "ws_size"); f.restype = C.c_size_t
        self._ws = np.zeros(f() // 8 + 8, dtype=np.float64)

    # -- single instance, full diagnostics --------------------------------------------------
    def rti(self, x0bar, yref, x, u, opts: IpmOpts | None = None, We=None):
        """x [(N+1),nx], u [N,nu] are copied and returned updated. yref [(N+1), ny]."""
        opts = opts or default_opts()
        x = np.array(x, dtype=np.float64, order="C"); u = np.array(u, dtype=np.float64, order="C")
        x0bar = np.ascontiguousarray(x0bar, dtype=np.float64)
        yref = np.ascontiguousarray(yref, dtype=np.float64)
        assert yref.shape == (ORC_N + 1, self.ny) and x.shape == (ORC_N + 1, self.nx) and u.shape == (ORC_N, self.nu)
        pi = np.zeros((ORC_N, self.nx))
        st = Stats()
        we = None if We is None else np.ascontiguousarray(We, dtype=np.float64)
        f = getattr(self.lib, self.pfx + "rti"); f.restype = C.c_int
        status = f(_dp(self.prob), C.byref(opts), _dp(x0bar), _dp(yref), None if we is None else _dp(we),
                   _dp(x), _dp(u), _dp(pi), C.byref(st), _dp(self._ws))
        return dict(status=status, x=x, u=u, pi=pi, stats=st, qp_iter=st.qp_iter)

    def qp_data(self):
        """QP of the last rti() call: BAbt [N, nz+1, nx], Hd/rq [N+1, nz], dlb/dub [N+1, nb]"""
        n = ORC_N
        BAbt = np.zeros((n, self.nz + 1, self.nx)); Hd = np.zeros((n + 1, self.nz)); rq = np.zeros((n + 1, self.nz))
        dlb = np.zeros((n + 1, self.nb)); dub = np.zeros((n + 1, self.nb))
        getattr(self.lib, self.pfx + "get_qp")(_dp(self._ws), _dp(BAbt), _dp(Hd), _dp(rq), _dp(dlb), _dp(dub))
        return dict(BAbt=BAbt, Hd=Hd, rq=rq, dlb=dlb, dub=dub)

    def qp_sol(self):
        n = ORC_N
        z = np.zeros((n + 1, self.nz)); pi = np.zeros((n, self.nx))
        ll = np.zeros((n + 1, self.nb)); lu = np.zeros((n + 1, self.nb)); tl = np.zeros((n + 1, self.nb)); tu = np.zeros((n + 1, self.nb))
        getattr(self.lib, self.pfx + "get_sol")(_dp(self._ws), _dp(z), _dp(pi), _dp(ll), _dp(lu), _dp(tl), _dp(tu))
        return dict(z=z, pi=pi, lam_lb=ll, lam_ub=lu, t_lb=tl, t_ub=tu)

    def discrete_map(self, x, u, p, h):
        x = np.ascontiguousarray(x, dtype=np.float64); u = np.ascontiguousarray(u, dtype=np.float64)
        p = np.ascontiguousarray(p, dtype=np.float64)
        xn = np.zeros(self.nx); S = np.zeros((self.nx, self.nz))
        getattr(self.lib, self.pfx + "discrete_map")(_dp(x), _dp(u), _dp(p), C.c_double(h), _dp(xn), _dp(S))
        return xn, S

    # -- batch (OpenMP, one solve per core) -------------------------------------------------
    def rti_batch(self, x0bar, yref, x, u, opts: IpmOpts | None = None, We=None, nthreads: int = 0):
        """instance-major arrays: x0bar [B,nx], yref [B,N+1,ny], x [B,N+1,nx], u [B,N,nu] (x,u updated in place)"""
        opts = opts or default_opts()
        B = x0bar.shape[0]
        for a in (x0bar, yref, x, u):
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        assert yref.shape == (B, ORC_N + 1, self.ny) and x.shape == (B, ORC_N + 1, self.nx) and u.shape == (B, ORC_N, self.nu)
        status = np.zeros(B, dtype=np.int32); qp_iter = np.zeros(B, dtype=np.int32); lin_res = np.zeros(B)
        we = None if We is None else np.ascontiguousarray(We, dtype=np.float64)
        f = getattr(self.lib, self.pfx + "rti_batch"); f.restype = C.c_int
        used = f(_dp(self.prob), C.byref(opts), C.c_int(B), _dp(x0bar), _dp(yref), None if we is None else _dp(we),
                 _dp(x), _dp(u), status.ctypes.data_as(C.POINTER(C.c_int)), qp_iter.ctypes.data_as(C.POINTER(C.c_int)),
                 C.c_int(nthreads), _dp(lin_res))
        return dict(status=status, qp_iter=qp_iter, threads=used, lin_res=lin_res)
