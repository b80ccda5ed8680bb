"""Closed-loop rollout on the device: for B robots, every tick

    nearest path parameter  ->  N+1 reference poses  ->  controller tick (RTI step)  ->  plant step

mirrors `NMPCNavControlROS::processFollowPath` (NMPCNavControlROS.cpp:648-698: processNearestPoint, PathDiscretizer,
executeNMPC -> run) closed through the nominal plant x+ = phi_RK4(x, u_0 + noise), the rollout the reference's
scripts/test_scripts/acados_sim_diff.py:119-163 does for one robot on the host.  All four steps are calls of the C ABI
(include/nmpc_b200.h) on one CUDA stream; nothing is synchronised or copied to the host between ticks.
The nearest-point step is a stand-in (the reference's is a private class, see csrc/rollout.cuh)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .controller import BatchedNavController
from .paths import BatchedPathDiscretizer, PathSet


class ClosedLoopRollout:
    def __init__(self, controller: BatchedNavController, paths: PathSet, path_id: torch.Tensor, back: float = 0.05,
                 ahead: float = 0.5, is_holonomic: bool = False):
        self.ctl, self.paths, self.path_id = controller, paths, path_id
        self.spec = controller.spec
        self.lib = controller.lib
        self.tdev = controller.tdev
        self.B = int(path_id.shape[0])
        self.back, self.ahead = float(back), float(ahead)
        # the node builds PathDiscretizer(dt, N+1, false) every tick (NMPCNavControlROS.cpp:666)
        self.disc = BatchedPathDiscretizer(controller.dt, self.spec.n + 1, is_holonomic, controller.solver.device)
        f64 = dict(dtype=torch.float64, device=self.tdev)
        B = self.B
        self.x = torch.zeros(self.spec.nx, B, **f64)              # plant state
        self.pose = torch.zeros(3, B, **f64); self.vel = torch.zeros(3, B, **f64); self.steer = torch.zeros(B, **f64)
        self.u = torch.zeros(B, **f64)                            # path parameter
        self.refs = torch.empty(self.spec.n + 1, 3, B, **f64)
        self.out = None

    def reset(self, pose0: torch.Tensor, u0: torch.Tensor):
        """robots at rest at pose0 [3,B], path parameters u0 [B]; constructor state of the controllers, zero iterate"""
        self.x.zero_(); self.x[:3].copy_(pose0)
        self.pose.copy_(pose0); self.vel.zero_(); self.steer.zero_(); self.u.copy_(u0)
        _lib.check(self.lib.nmpc_ctrl_reset(self.ctl.solver._h, C.c_void_p(torch.cuda.current_stream(self.tdev).cuda_stream)),
                   "nmpc_ctrl_reset")
        self.ctl.solver.reset_async()

    def step(self, noise: torch.Tensor | None = None):
        """one tick, enqueued on torch's current stream; noise [nu,B] is added to u_0 in the plant"""
        st = C.c_void_p(torch.cuda.current_stream(self.tdev).cuda_stream)
        p, B = self.paths, self.B
        _lib.check(self.lib.nmpc_path_nearest_device(self.ctl.solver.device, B, C.c_void_p(p.segments.data_ptr()),
                                                     C.c_void_p(p.offsets.data_ptr()), p.n_paths, C.c_void_p(self.path_id.data_ptr()),
                                                     C.c_void_p(self.pose.data_ptr()), self.back, self.ahead,
                                                     C.c_void_p(self.u.data_ptr()), st), "nmpc_path_nearest_device")
        self.disc.get_next_n_poses(p, self.path_id, self.u, out=self.refs)
        if self.spec.name == "tric":
            self.ctl.set_steering_wheel_angle(self.steer)
        self.out = self.ctl.run(self.pose, self.vel, self.refs, out=self.out)
        _lib.check(self.lib.nmpc_plant_step_device(self.ctl.solver._h, B, self.ctl.dt, None if noise is None else C.c_void_p(noise.data_ptr()),
                                                   C.c_void_p(self.x.data_ptr()), C.c_void_p(self.pose.data_ptr()),
                                                   C.c_void_p(self.vel.data_ptr()), C.c_void_p(self.steer.data_ptr()), st),
                   "nmpc_plant_step_device")
        return self.out

    def run_engine(self, ticks: int, noise: torch.Tensor | None = None, sqp_max_iter: int = 1, sqp_tol: float = 0.0,
                   shift: bool = False):
        """`ticks` closed-loop ticks through ONE call of the C ABI (`nmpc_rollout_device`): the library enqueues nearest
        point -> reference poses -> controller tick (RTI, or SQP to convergence) -> plant step -> optional warm-start
        shift for every tick on torch's current stream; no Python (and no host synchronisation) between ticks.
        noise [ticks,nu,B] or None.  Returns dict(pose [ticks+1,3,B], cmd [ticks,3,B], failed [ticks] int32)."""
        f64 = dict(dtype=torch.float64, device=self.tdev)
        B = self.B
        traj = torch.empty(ticks + 1, 3, B, **f64); cmds = torch.empty(ticks, 3, B, **f64)
        failed = torch.zeros(ticks, dtype=torch.int32, device=self.tdev)
        if noise is not None:
            assert noise.is_cuda and noise.is_contiguous() and noise.shape == (ticks, self.spec.nu, B)
        o = _lib.RolloutOpts(dt=self.ctl.dt, back=self.back, ahead=self.ahead, sqp_tol=float(sqp_tol),
                             is_holonomic=int(self.disc.is_holonomic), sqp_max_iter=int(sqp_max_iter), shift=int(bool(shift)))
        p = self.paths
        st = C.c_void_p(torch.cuda.current_stream(self.tdev).cuda_stream)
        vp = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        _lib.check(self.lib.nmpc_rollout_device(self.ctl.solver._h, B, int(ticks), C.byref(o), vp(p.segments), vp(p.offsets), p.n_paths,
                                                vp(self.path_id), vp(self.u), vp(self.x), vp(self.pose), vp(self.vel),
                                                vp(self.steer) if self.spec.name == "tric" else None, vp(noise), vp(traj), vp(cmds),
                                                vp(failed), st), "nmpc_rollout_device")
        return dict(pose=traj, cmd=cmds, failed=failed)

    def run(self, ticks: int, noise: torch.Tensor | None = None):
        """`ticks` closed-loop ticks; noise [ticks,nu,B] or None.  Returns dict(pose [ticks+1,3,B], cmd [ticks,3,B],
        failed [ticks] = robots with a non-zero solver status per tick), CUDA tensors; asynchronous."""
        f64 = dict(dtype=torch.float64, device=self.tdev)
        traj = torch.empty(ticks + 1, 3, self.B, **f64); cmds = torch.empty(ticks, 3, self.B, **f64)
        failed = torch.zeros(ticks, dtype=torch.int64, device=self.tdev)
        traj[0].copy_(self.pose)
        for t in range(ticks):
            out = self.step(None if noise is None else noise[t])
            cmds[t].copy_(out["cmd"]); traj[t + 1].copy_(self.pose)
            failed[t] = (out["status"] != 0).sum()
        return dict(pose=traj, cmd=cmds, failed=failed)
