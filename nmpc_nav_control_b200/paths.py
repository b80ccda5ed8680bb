"""Batched path discretisation: `PathDiscretizer::getNextNPoses`
(src/nmpc_nav_control/PathDiscretizer.cpp:14-63) for B robots on the device, writing the N+1 reference poses in the
structure-of-arrays layout `nmpc_ctrl_tick_device` reads.

The reference walks a list of `parametric_trajectories_common::TPath` curves, a PRIVATE dependency that is not part
of the reference tree; here a path is a list of segments from the families SURVEY.md 8(f2) names, sixteen doubles
each (`nmpc_path_segment` in include/nmpc_b200.h):

    [kind, vel, th0, th1, cx[0..5], cy[0..5]]
    kind 0: x(u) = sum cx[i] u^i, y(u) = sum cy[i] u^i     (line, cubic Bezier, any polynomial up to degree 5)
    kind 1: x(u) = cx[0] + cx[1] cos(cx[2] + cx[3] u), y(u) = cy[0] + cx[1] sin(cx[2] + cx[3] u)     (arc)
    vel: signed speed on the segment; th0, th1: holonomic heading at u = 0, 1

Everything goes through the C ABI (`nmpc_path_discretize_device`); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib

SEG = 16


def line(p0, p1, vel, th0=0.0, th1=0.0):
    r = np.zeros(SEG)
    r[1], r[2], r[3] = vel, th0, th1
    r[4], r[5] = p0[0], p1[0] - p0[0]
    r[10], r[11] = p0[1], p1[1] - p0[1]
    return r


def bezier3(p0, p1, p2, p3, vel, th0=0.0, th1=0.0):
    """cubic Bezier with control points p0..p3, in the power basis"""
    r = np.zeros(SEG)
    r[1], r[2], r[3] = vel, th0, th1
    for a, o in ((0, 4), (1, 10)):
        b0, b1, b2, b3 = p0[a], p1[a], p2[a], p3[a]
        r[o], r[o + 1], r[o + 2], r[o + 3] = b0, 3.0 * (b1 - b0), 3.0 * (b2 - 2.0 * b1 + b0), b3 - 3.0 * b2 + 3.0 * b1 - b0
    return r


def arc(center, radius, a0, a1, vel, th0=0.0, th1=0.0):
    """circular arc from polar angle a0 to a1 about `center`"""
    r = np.zeros(SEG)
    r[0], r[1], r[2], r[3] = 1.0, vel, th0, th1
    r[4], r[5], r[6], r[7] = center[0], radius, a0, a1 - a0
    r[10] = center[1]
    return r


class PathSet:
    """several paths packed for the device: segments [n_seg_total, 16], offsets [n_paths + 1]"""

    def __init__(self, paths, device: int = 0):
        self.tdev = torch.device("cuda", device)
        segs, off = [], [0]
        for p in paths:
            p = np.asarray(p, dtype=np.float64).reshape(-1, SEG)
            if len(p) < 1:
                raise ValueError("a path needs at least one segment")
            segs.append(p); off.append(off[-1] + len(p))
        self.n_paths = len(paths)
        self.host_segments = np.concatenate(segs, axis=0)
        self.host_offsets = np.array(off, dtype=np.int32)
        self.segments = torch.from_numpy(self.host_segments).to(self.tdev)
        self.offsets = torch.from_numpy(self.host_offsets).to(self.tdev)


class BatchedPathDiscretizer:
    """PathDiscretizer(sample_period, num_poses, is_holonomic) (PathDiscretizer.cpp:5-12) for a batch of robots"""

    def __init__(self, sample_period: float, num_poses: int, is_holonomic: bool = False, device: int = 0):
        self.sample_period, self.num_poses, self.is_holonomic = float(sample_period), int(num_poses), bool(is_holonomic)
        self.lib = _lib.load()
        self.device = device
        self.tdev = torch.device("cuda", device)

    def get_next_n_poses(self, paths: PathSet, path_id: torch.Tensor, nearest_sample_u: torch.Tensor,
                         out: torch.Tensor | None = None, stream: torch.cuda.Stream | None = None) -> torch.Tensor:
        """path_id [B] int32, nearest_sample_u [B] float64 (segment index + parameter, as active_path_u_ in
        NMPCNavControlROS.cpp:668); returns the poses [num_poses, 3, B] (x, y, theta); asynchronous"""
        B = path_id.shape[0]
        assert path_id.is_cuda and path_id.dtype == torch.int32 and nearest_sample_u.is_cuda and nearest_sample_u.dtype == torch.float64
        assert nearest_sample_u.shape == (B,)
        if out is None:
            out = torch.empty(self.num_poses, 3, B, dtype=torch.float64, device=self.tdev)
        assert out.shape == (self.num_poses, 3, B) and out.is_contiguous()
        st = stream if stream is not None else torch.cuda.current_stream(self.tdev)
        _lib.check(self.lib.nmpc_path_discretize_device(
            self.device, B, C.c_void_p(paths.segments.data_ptr()), C.c_void_p(paths.offsets.data_ptr()), paths.n_paths,
            C.c_void_p(path_id.data_ptr()), C.c_void_p(nearest_sample_u.data_ptr()), self.sample_period, self.num_poses,
            int(self.is_holonomic), C.c_void_p(out.data_ptr()), C.c_void_p(st.cuda_stream)), "nmpc_path_discretize_device")
        return out
