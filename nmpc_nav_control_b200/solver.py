"""Python front-end of the batched SQP-RTI solver (torch only for device memory and streams).

`BatchedRtiSolver` mirrors, for B instances at once, the call protocol of the reference's
solver wrapper (src/nmpc_nav_control/NMPCNavControlDiff.cpp):

    ctor  : create capsule, set params / bounds / weights         (Diff.cpp:6-74)
    run() : set x0 (stage-0 lbx=ubx), set yref per stage, [W_e], solve, read u_0 and x_1
                                                                   (Diff.cpp:82-175)
    reset_mpc(): zero the persisted iterate                        (Diff.cpp:177-181)

Everything goes through the C ABI of libnmpc_b200.so (include/nmpc_b200.h).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .problem import ModelSpec, get_model


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


class BatchedRtiSolver:
    def __init__(self, model, max_batch: int, device: int = 0):
        self.spec: ModelSpec = model if isinstance(model, ModelSpec) else get_model(model)
        self.lib = _lib.load()
        self.max_batch = int(max_batch)
        self.device = int(device)
        h = C.c_void_p()
        _lib.check(self.lib.nmpc_create(self.spec.model_id, self.max_batch, self.device, C.byref(h)), "nmpc_create")
        self._h = h
        self.tdev = torch.device("cuda", self.device)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.nmpc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- set-up (constructor part of the reference wrapper) --------------------------------
    def set_tables(self, W=None, We=None, lbx=None, ubx=None, lbu=None, ubu=None, p=None):
        s, n = self.spec, self.spec.n

        def arr(a, shape):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.shape != shape:
                raise ValueError(f"expected shape {shape}, got {a.shape}")
            return a
        W, We = arr(W, (n, s.ny)), arr(We, (s.nx,))
        if W is not None or We is not None:
            _lib.check(self.lib.nmpc_set_weights(self._h, _ptr(W), _ptr(We)), "nmpc_set_weights")
        lbx, ubx = arr(lbx, (n, s.nbx)), arr(ubx, (n, s.nbx))
        lbu, ubu = arr(lbu, (n, s.nbu)), arr(ubu, (n, s.nbu))
        if any(a is not None for a in (lbx, ubx, lbu, ubu)):
            _lib.check(self.lib.nmpc_set_bounds(self._h, _ptr(lbx), _ptr(ubx), _ptr(lbu), _ptr(ubu)), "nmpc_set_bounds")
        p = arr(p, (n, s.np_))
        if p is not None:
            _lib.check(self.lib.nmpc_set_params(self._h, _ptr(p)), "nmpc_set_params")

    def get_opts(self) -> _lib.IpmOpts:
        o = _lib.IpmOpts()
        _lib.check(self.lib.nmpc_get_opts(self._h, C.byref(o)), "nmpc_get_opts")
        return o

    def set_opts(self, **kw):
        o = self.get_opts()
        for k, v in kw.items():
            if not hasattr(o, k):
                raise AttributeError(k)
            setattr(o, k, v)
        _lib.check(self.lib.nmpc_set_opts(self._h, C.byref(o)), "nmpc_set_opts")

    # ---- iterate ---------------------------------------------------------------------------
    def reset(self):
        """zero iterate, as `{m}_acados_reset` does after every new goal / path (synchronous)"""
        _lib.check(self.lib.nmpc_reset(self._h), "nmpc_reset")

    def reset_async(self, stream: "torch.cuda.Stream | None" = None):
        """same, enqueued without host sync on torch's current stream (or `stream`)"""
        st = stream if stream is not None else torch.cuda.current_stream(self.tdev)
        _lib.check(self.lib.nmpc_reset_async(self._h, C.c_void_p(st.cuda_stream)), "nmpc_reset_async")

    def set_iterate(self, x, u):
        x = np.ascontiguousarray(x, dtype=np.float64); u = np.ascontiguousarray(u, dtype=np.float64)
        B = x.shape[0]
        assert x.shape == (B, self.spec.n + 1, self.spec.nx) and u.shape == (B, self.spec.n, self.spec.nu)
        _lib.check(self.lib.nmpc_set_iterate_host(self._h, B, _ptr(x), _ptr(u)), "nmpc_set_iterate_host")

    def get_iterate(self, B: int):
        x = np.empty((B, self.spec.n + 1, self.spec.nx)); u = np.empty((B, self.spec.n, self.spec.nu))
        _lib.check(self.lib.nmpc_get_iterate_host(self._h, B, _ptr(x), _ptr(u)), "nmpc_get_iterate_host")
        return x, u

    # ---- solve: device-resident SoA batch ----------------------------------------------------
    def solve_device(self, x0bar: torch.Tensor, yref: torch.Tensor, We: torch.Tensor | None = None,
                     x: torch.Tensor | None = None, u: torch.Tensor | None = None, want_stats: bool = False,
                     out: dict | None = None, stream: torch.cuda.Stream | None = None):
        """x0bar [nx,B], yref [N+1,nyref,B] (nyref = 3 or ny), We [nx,B]; optional external iterate
        x [N+1,nx,B], u [N,nu,B].  Returns dict(status, qp_iter[, stats]) of CUDA tensors; asynchronous."""
        s = self.spec
        B = x0bar.shape[1]
        nyref = yref.shape[1]
        for t in (x0bar, yref, We, x, u):
            if t is not None:
                assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
        assert x0bar.shape == (s.nx, B) and yref.shape == (s.n + 1, nyref, B)
        if out is None:
            out = dict(status=torch.empty(B, dtype=torch.int32, device=self.tdev),
                       qp_iter=torch.empty(B, dtype=torch.int32, device=self.tdev))
            if want_stats:
                out["stats"] = torch.empty(8, B, dtype=torch.float64, device=self.tdev)
        ld = 0
        if x is not None:
            assert x.shape == (s.n + 1, s.nx, B) and u.shape == (s.n, s.nu, B)
            ld = B
        st = stream if stream is not None else torch.cuda.current_stream(self.tdev)
        _lib.check(self.lib.nmpc_rti_solve_device(
            self._h, B, _ptr(x0bar), _ptr(yref), nyref, _ptr(We), _ptr(x), _ptr(u), ld,
            _ptr(out["status"]), _ptr(out["qp_iter"]), _ptr(out.get("stats")), C.c_void_p(st.cuda_stream)),
            "nmpc_rti_solve_device")
        return out

    # ---- BASELINE config 4: SQP to convergence and the warm-start shift ---------------------------
    def sqp_solve_device(self, x0bar: torch.Tensor, yref: torch.Tensor, max_iter: int, tol: float, We: torch.Tensor | None = None,
                         x: torch.Tensor | None = None, u: torch.Tensor | None = None, out: dict | None = None,
                         stream: torch.cuda.Stream | None = None):
        """per instance: RTI steps until the inf-norm of the step is <= tol (or max_iter); arguments as solve_device.
        Returns dict(status, sqp_iter, qp_iter) of CUDA tensors (qp_iter summed over the steps); asynchronous."""
        s = self.spec
        B = x0bar.shape[1]
        nyref = yref.shape[1]
        for t in (x0bar, yref, We, x, u):
            if t is not None:
                assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
        assert x0bar.shape == (s.nx, B) and yref.shape == (s.n + 1, nyref, B)
        if out is None:
            out = {k: torch.empty(B, dtype=torch.int32, device=self.tdev) for k in ("status", "sqp_iter", "qp_iter")}
        ld = 0
        if x is not None:
            assert x.shape == (s.n + 1, s.nx, B) and u.shape == (s.n, s.nu, B)
            ld = B
        st = stream if stream is not None else torch.cuda.current_stream(self.tdev)
        _lib.check(self.lib.nmpc_sqp_solve_device(
            self._h, B, _ptr(x0bar), _ptr(yref), nyref, _ptr(We), _ptr(x), _ptr(u), ld, int(max_iter), C.c_double(tol),
            _ptr(out["status"]), _ptr(out["sqp_iter"]), _ptr(out["qp_iter"]), C.c_void_p(st.cuda_stream)), "nmpc_sqp_solve_device")
        return out

    def shift(self, B: int, x: torch.Tensor | None = None, u: torch.Tensor | None = None, mask: torch.Tensor | None = None,
              stream: torch.cuda.Stream | None = None):
        """x_k <- x_{k+1}, u_k <- u_{k+1}, last stage kept; x/u None = the persisted iterate;
        mask [B] int32 or None"""
        ld = 0
        if x is not None:
            assert x.is_cuda and u.is_cuda and x.is_contiguous() and u.is_contiguous()
            ld = x.shape[2]
        if mask is not None:
            assert mask.is_cuda and mask.dtype == torch.int32 and mask.shape == (B,)
        st = stream if stream is not None else torch.cuda.current_stream(self.tdev)
        _lib.check(self.lib.nmpc_shift_device(self._h, B, _ptr(x), _ptr(u), ld, _ptr(mask), C.c_void_p(st.cuda_stream)),
                   "nmpc_shift_device")

    # ---- solve: host buffers, instance-major (the controller-facing call) --------------------
    def solve_host(self, x0bar, yref, We=None, out: dict | None = None):
        """x0bar [B,nx], yref [B,N+1,nyref], We [B,nx] host arrays (numpy or pinned torch).
        Returns dict(u0 [B,nu], x1 [B,nx], status [B], qp_iter [B]) as numpy arrays; synchronous."""
        s = self.spec
        x0bar = np.ascontiguousarray(x0bar, dtype=np.float64); yref = np.ascontiguousarray(yref, dtype=np.float64)
        if x0bar.ndim != 2 or x0bar.shape[1] != s.nx:
            raise ValueError(f"x0bar must be [B, {s.nx}], got {x0bar.shape}")
        B = x0bar.shape[0]
        if yref.ndim != 3 or yref.shape[:2] != (B, s.n + 1) or yref.shape[2] not in (3, s.ny):
            raise ValueError(f"yref must be [B, {s.n + 1}, 3 or {s.ny}], got {yref.shape}")
        nyref = yref.shape[2]
        if We is not None:
            We = np.ascontiguousarray(We, dtype=np.float64)
            if We.shape != (B, s.nx):
                raise ValueError(f"We must be [B, {s.nx}], got {We.shape}")
        if out is None:
            out = dict(u0=np.empty((B, s.nu)), x1=np.empty((B, s.nx)),
                       status=np.empty(B, dtype=np.int32), qp_iter=np.empty(B, dtype=np.int32))
        _lib.check(self.lib.nmpc_rti_solve_host(self._h, B, _ptr(x0bar), _ptr(yref), nyref, _ptr(We),
                                                _ptr(out["u0"]), _ptr(out["x1"]), _ptr(out["status"]), _ptr(out["qp_iter"])),
                   "nmpc_rti_solve_host")
        return out

    def last_timing(self):
        ms = (C.c_double * 4)()
        _lib.check(self.lib.nmpc_last_timing(self._h, ms), "nmpc_last_timing")
        return dict(linearize_ms=ms[0], qp_ms=ms[1], step_ms=ms[2], total_ms=ms[3])

    def last_launches(self) -> int:
        return int(self.lib.nmpc_last_launches(self._h))


def dfma_peak_tflops(device: int = 0, iters: int = 20000) -> float:
    return float(_lib.load().nmpc_dfma_peak_tflops(device, iters))
