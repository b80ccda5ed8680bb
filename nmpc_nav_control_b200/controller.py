"""Batched controller tick: `NMPCNavControl{Diff,Omni4,Tric}` for B robots at once.

`BatchedNavController` mirrors the reference's solver-wrapper interface
(include/nmpc_nav_control/NMPCNavControl.h:20-57 and the three subclasses):

    ctor                     : parameters, bounds, W from W_diag, W_e from W_diag[:nx]
                               (src/nmpc_nav_control/NMPCNavControlDiff.cpp:6-74, Omni4.cpp:6-83, Tric.cpp:6-80)
    run(pose, vel, traj_ref) : one tick -> velocity command   (Diff.cpp:82-175, Omni4.cpp:91-177, Tric.cpp:88-181)
    reset_mpc()              : zero the iterate                (Diff.cpp:177-181)
    set_steering_wheel_angle : tric only                       (NMPCNavControlTric.h:67-69)
    get_horizon / get_delta_time                               (NMPCNavControl.h:38-39)

The pre- and post-processing run() does around `{m}_acados_solve` (initial state from the pose and the direct
kinematics of the measured twist, heading unwrap chain, reference padding, the diff terminal-weight switch,
reference-state integration, inverse kinematics) is done on the device by `nmpc_ctrl_tick_device`
(nmpc_nav_control_b200/csrc/ctrl_glue.cuh), so a closed-loop sweep over many robots never leaves the GPU.
Everything goes through the C ABI (include/nmpc_b200.h); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .problem import ModelSpec, get_model
from .solver import BatchedRtiSolver, _ptr


class BatchedNavController:
    def __init__(self, model, max_batch: int, dt: float, W_diag=None, p=None, x_min=None, x_max=None, u_min=None,
                 u_max=None, device: int = 0):
        """dt = 1 / control_freq (NMPCNavControlROS.cpp:82); W_diag [ny] (default Q | R of the yaml), p [np], x_min /
        x_max [nbx], u_min / u_max [nbu] default to the code-generation values (problem.py)."""
        self.spec: ModelSpec = model if isinstance(model, ModelSpec) else get_model(model)
        s = self.spec
        self.dt = float(dt)
        self.solver = BatchedRtiSolver(s, max_batch, device)
        self.lib = self.solver.lib
        self.tdev = self.solver.tdev
        n = s.n
        W_diag = np.asarray(s.Q + s.R if W_diag is None else W_diag, dtype=np.float64)
        if W_diag.shape != (s.ny,):
            raise ValueError(f"W_diag must have {s.ny} entries")
        tile = lambda a, d: np.tile(np.asarray(d if a is None else a, dtype=np.float64), (n, 1))
        # the wrappers set W_e from W_diag[0..nx), not from the yaml's QN (Diff.cpp:34-41)
        self.solver.set_tables(W=np.tile(W_diag, (n, 1)), We=W_diag[:s.nx].copy(), p=tile(p, s.p),
                               lbx=tile(x_min, s.lbx), ubx=tile(x_max, s.ubx), lbu=tile(u_min, s.lbu), ubu=tile(u_max, s.ubu))
        self._steer = None
        self._zero_steer = None
        _lib.check(self.lib.nmpc_ctrl_reset(self.solver._h, None), "nmpc_ctrl_reset")
        torch.cuda.synchronize(self.tdev)

    def close(self):
        self.solver.close()

    def get_horizon(self) -> int:
        return self.spec.n

    def get_delta_time(self) -> float:
        return self.dt

    def set_steering_wheel_angle(self, alpha):
        """tric: measured steering-wheel angle per robot, [B] CUDA tensor or host array (used by the next run)"""
        self._steer = alpha

    def reset_mpc(self):
        self.solver.reset()

    def reference_states(self) -> torch.Tensor:
        """the carried reference states (x0[3+nv:] of the next tick), [nv, max_batch] view of device memory"""
        p = C.c_void_p(); ld = C.c_int()
        _lib.check(self.lib.nmpc_ctrl_state_device(self.solver._h, C.byref(p), C.byref(ld)), "nmpc_ctrl_state_device")
        return _device_view(p.value, (self.spec.nv, ld.value), self.tdev)

    # ---- device tick: SoA CUDA tensors -------------------------------------------------------
    def run(self, pose: torch.Tensor, vel: torch.Tensor, traj_ref: torch.Tensor, nref: torch.Tensor | None = None,
            cmd: torch.Tensor | None = None, out: dict | None = None, stream: torch.cuda.Stream | None = None,
            sqp_max_iter: int = 1, sqp_tol: float = 0.0):
        """pose [3,B], vel [3,B] (v, vn, w), traj_ref [nref_max,3,B], nref [B] int32 or None (= nref_max);
        sqp_max_iter > 1: SQP to convergence instead of the reference's single RTI step (BASELINE config 4);
        returns dict(cmd [3,B], status [B], qp_iter [B]) of CUDA tensors; asynchronous."""
        B = pose.shape[1]
        nref_max = traj_ref.shape[0]
        for t in (pose, vel, traj_ref):
            assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
        assert pose.shape == (3, B) and vel.shape == (3, B) and traj_ref.shape == (nref_max, 3, B)
        if nref is not None:
            assert nref.is_cuda and nref.dtype == torch.int32 and nref.shape == (B,)
        st = stream if stream is not None else torch.cuda.current_stream(self.tdev)
        steer = self._steer
        if steer is None and self.spec.name == "tric":
            if self._zero_steer is None or self._zero_steer.shape[0] < B:      # the wrapper's initial angle, Tric.cpp:14
                self._zero_steer = torch.zeros(self.solver.max_batch, dtype=torch.float64, device=self.tdev)
                torch.cuda.current_stream(self.tdev).synchronize()              # persistent buffer: usable on any stream afterwards
            steer = self._zero_steer[:B]
        if steer is not None and not isinstance(steer, torch.Tensor):
            steer = torch.as_tensor(np.asarray(steer, dtype=np.float64), device=self.tdev)
        if steer is not None:
            assert steer.is_cuda and steer.dtype == torch.float64 and steer.shape == (B,)
        if out is None:
            out = dict(cmd=torch.zeros(3, B, dtype=torch.float64, device=self.tdev) if cmd is None else cmd,
                       status=torch.empty(B, dtype=torch.int32, device=self.tdev),
                       qp_iter=torch.empty(B, dtype=torch.int32, device=self.tdev))
        if sqp_max_iter > 1:
            _lib.check(self.lib.nmpc_ctrl_tick_sqp_device(self.solver._h, B, _ptr(pose), _ptr(vel), _ptr(steer), _ptr(traj_ref),
                                                          _ptr(nref), nref_max, C.c_double(self.dt), int(sqp_max_iter), C.c_double(sqp_tol),
                                                          _ptr(out["cmd"]), _ptr(out["status"]), _ptr(out["qp_iter"]),
                                                          C.c_void_p(st.cuda_stream)), "nmpc_ctrl_tick_sqp_device")
        else:
            _lib.check(self.lib.nmpc_ctrl_tick_device(self.solver._h, B, _ptr(pose), _ptr(vel), _ptr(steer), _ptr(traj_ref),
                                                      _ptr(nref), nref_max, self.dt, _ptr(out["cmd"]), _ptr(out["status"]),
                                                      _ptr(out["qp_iter"]), C.c_void_p(st.cuda_stream)), "nmpc_ctrl_tick_device")
        return out

    # ---- host tick: instance-major numpy arrays ------------------------------------------------
    def run_host(self, pose, vel, traj_ref, nref=None, out: dict | None = None):
        """pose [B,3], vel [B,3], traj_ref [B,nref_max,3] (nref_max <= N+1), nref [B] or None; returns dict(cmd [B,3],
        status [B], qp_iter [B]) numpy arrays; synchronous."""
        pose = np.ascontiguousarray(pose, dtype=np.float64); vel = np.ascontiguousarray(vel, dtype=np.float64)
        traj_ref = np.ascontiguousarray(traj_ref, dtype=np.float64)
        B, nref_max = traj_ref.shape[0], traj_ref.shape[1]
        assert pose.shape == (B, 3) and vel.shape == (B, 3) and traj_ref.shape == (B, nref_max, 3)
        nref = None if nref is None else np.ascontiguousarray(nref, dtype=np.int32)
        steer = None if self._steer is None else np.ascontiguousarray(
            self._steer.cpu().numpy() if isinstance(self._steer, torch.Tensor) else self._steer, dtype=np.float64)
        if steer is None and self.spec.name == "tric":
            steer = np.zeros(B)
        if out is None:
            out = dict(cmd=np.zeros((B, 3)), status=np.empty(B, dtype=np.int32), qp_iter=np.empty(B, dtype=np.int32))
        _lib.check(self.lib.nmpc_ctrl_tick_host(self.solver._h, B, _ptr(pose), _ptr(vel), _ptr(steer), _ptr(traj_ref), _ptr(nref),
                                                nref_max, self.dt, _ptr(out["cmd"]), _ptr(out["status"]), _ptr(out["qp_iter"])),
                   "nmpc_ctrl_tick_host")
        return out


def _device_view(addr: int, shape, dev) -> torch.Tensor:
    """float64 CUDA tensor over memory the library owns (no copy, no ownership)"""
    n = int(np.prod(shape))

    class _Mem:
        __cuda_array_interface__ = dict(shape=(n,), typestr="<f8", data=(addr, False), version=2)
    return torch.as_tensor(_Mem(), device=dev).view(*shape)
