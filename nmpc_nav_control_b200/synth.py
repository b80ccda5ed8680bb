"""Synthetic OCP instances (initial states + reference paths) for tests and benchmarks.

The reference has no input fixtures; this follows SURVEY.md Appendix D: a
counter-based RNG so that instance *i* of a model is identical for every batch size, shard
and device, initial states inside the velocity envelope, and reference paths shaped like
what `PathDiscretizer::getNextNPoses` (src/nmpc_nav_control/PathDiscretizer.cpp:14-63) plus
the wrapper's padding / angle unwrapping (src/nmpc_nav_control/NMPCNavControlDiff.cpp:104-118,
NMPCNavControl.cpp:25-31) deliver: N+1 poses spaced |v|*dt along a curve, padded with the
last pose, theta unwrapped in a chain that starts at the robot heading.

Everything is plain torch integer / fp64 tensor arithmetic, so it runs on CPU (tests) and on
the GPU (bench: inputs are created where they are consumed).
"""
from __future__ import annotations

import math

import torch

from .problem import ModelSpec

SEED = 20261018
_M64 = (1 << 64) - 1


def _s64(v: int) -> int:
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


_GAMMA = _s64(0x9E3779B97F4A7C15)
_C1 = _s64(0xBF58476D1CE4E5B9)
_C2 = _s64(0x94D049BB133111EB)


def _lsr(z: torch.Tensor, s: int) -> torch.Tensor:
    return (z >> s) & ((1 << (64 - s)) - 1)


def _mix(z: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser on int64 tensors (two's-complement wrap-around arithmetic)."""
    z = (z ^ _lsr(z, 30)) * _C1
    z = (z ^ _lsr(z, 27)) * _C2
    return z ^ _lsr(z, 31)


def _mix_int(v: int) -> int:
    v &= _M64
    v = ((v ^ (v >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    v = ((v ^ (v >> 27)) * 0x94D049BB133111EB) & _M64
    return v ^ (v >> 31)


class CounterRng:
    """u(i, d) in [0,1): draw d (< 64) of instance i; stateless."""

    def __init__(self, seed: int, model_id: int, index: torch.Tensor):
        base = _mix_int((seed * 0x9E3779B97F4A7C15 + model_id + 1) & _M64)
        self.base = _s64(base)
        self.index = index.to(torch.int64)

    def uniform(self, draw: int, lo: float = 0.0, hi: float = 1.0) -> torch.Tensor:
        ctr = (self.index * 64 + (draw + 1)) * _GAMMA + self.base
        bits = _lsr(_mix(ctr), 11)
        u = bits.to(torch.float64) * (1.0 / 9007199254740992.0)
        return lo + (hi - lo) * u

    def normal_pair(self, draw: int, sigma: float):
        u1 = self.uniform(draw).clamp_min(1e-300)
        u2 = self.uniform(draw + 1)
        r = torch.sqrt(-2.0 * torch.log(u1)) * sigma
        return r * torch.cos(2.0 * math.pi * u2), r * torch.sin(2.0 * math.pi * u2)


def _norm_ang(a: torch.Tensor) -> torch.Tensor:
    """wrap to (-pi, pi] like include/nmpc_nav_control/utils.h:normAngRad"""
    return torch.atan2(torch.sin(a), torch.cos(a))


def _sinc(a: torch.Tensor) -> torch.Tensor:
    small = a.abs() < 1e-6
    safe = torch.where(small, torch.ones_like(a), a)
    return torch.where(small, 1.0 - a * a / 6.0, torch.sin(safe) / safe)


def make_instances(spec: ModelSpec, start: int, count: int, device="cpu", seed: int = SEED,
                   pose_only: bool = False, terminal_hack: bool = False) -> dict:
    """Instances [start, start+count) of model `spec`, instance-major (AoS) fp64 tensors.

    returns dict(x0=[B,nx], yref=[B,N+1,ny] (or [B,N+1,3] if pose_only),
                 We=[B,nx] or None (terminal_hack: the diff wrapper's W_e switch,
                 NMPCNavControlDiff.cpp:127-139, applied to the codegen Q))
    """
    n, dt, nx, nv = spec.n, spec.dt, spec.nx, spec.nv
    idx = torch.arange(start, start + count, device=device, dtype=torch.int64)
    rng = CounterRng(seed, spec.model_id, idx)
    f64 = dict(dtype=torch.float64, device=device)

    # ---- D.2 initial state --------------------------------------------------------------
    x0 = torch.zeros(count, nx, **f64)
    x0[:, 0] = rng.uniform(0, -1.0, 1.0)
    x0[:, 1] = rng.uniform(1, -1.0, 1.0)
    x0[:, 2] = rng.uniform(2, -math.pi, math.pi)
    if spec.name == "tric":
        deg = math.pi / 180.0
        x0[:, 3] = rng.uniform(3, -0.6, 0.6)
        x0[:, 4] = rng.uniform(4, -20.0 * deg, 20.0 * deg)
        x0[:, 5] = (x0[:, 3] + rng.uniform(7, -0.1, 0.1)).clamp(-0.9, 0.9)
        x0[:, 6] = (x0[:, 4] + rng.uniform(8, -5.0 * deg, 5.0 * deg)).clamp(-27.0 * deg, 27.0 * deg)
    else:
        vmax = spec.ubx[0]
        for c in range(nv):
            x0[:, 3 + c] = rng.uniform(3 + c, -0.6, 0.6) * vmax
            x0[:, 3 + nv + c] = (x0[:, 3 + c] + rng.uniform(7 + c, -0.1, 0.1)).clamp(-0.9 * vmax, 0.9 * vmax)

    # ---- D.3 reference path -------------------------------------------------------------
    goto = rng.uniform(11) >= 0.7
    nxn, nyn = rng.normal_pair(12, 0.05)
    xs = x0[:, 0] + nxn
    ys = x0[:, 1] + nyn
    ths = x0[:, 2] + rng.uniform(14, -0.3, 0.3)
    vp = rng.uniform(15, 0.2, 0.8)
    kap = rng.uniform(16, -1.5, 1.5)
    length = rng.uniform(17, 0.3, 3.0)
    omh = rng.uniform(18, -0.5, 0.5)
    gd = rng.uniform(19, 0.0, 2.0)
    gb = rng.uniform(20, -math.pi, math.pi)
    gth = rng.uniform(21, -math.pi, math.pi)

    i = torch.arange(n + 1, **f64)[None, :]                      # [1, N+1]
    s = torch.minimum(i * (vp[:, None] * dt), length[:, None])   # arc length, padded at the end
    half = 0.5 * kap[:, None] * s
    px = xs[:, None] + s * _sinc(half) * torch.cos(ths[:, None] + half)
    py = ys[:, None] + s * _sinc(half) * torch.sin(ths[:, None] + half)
    if spec.name == "omni4":
        t = s / vp[:, None]
        pth = ths[:, None] + omh[:, None] * t
    else:
        pth = ths[:, None] + kap[:, None] * s
    gx = (x0[:, 0] + gd * torch.cos(gb))[:, None].expand(-1, n + 1)
    gy = (x0[:, 1] + gd * torch.sin(gb))[:, None].expand(-1, n + 1)
    gt = gth[:, None].expand(-1, n + 1)
    px = torch.where(goto[:, None], gx, px)
    py = torch.where(goto[:, None], gy, py)
    pth = _norm_ang(torch.where(goto[:, None], gt, pth))

    # unwrap chain starting from the robot heading (NMPCNavControlDiff.cpp:104-112)
    prev = x0[:, 2].clone()
    cols = []
    for k in range(n + 1):
        cur = pth[:, k]
        delta = cur - prev
        cur = torch.where(delta > math.pi, cur - 2.0 * math.pi,
                          torch.where(delta < -math.pi, cur + 2.0 * math.pi, cur))
        cols.append(cur)
        prev = cur
    pth = torch.stack(cols, dim=1)

    width = 3 if pose_only else spec.ny
    yref = torch.zeros(count, n + 1, width, **f64)
    yref[:, :, 0] = px
    yref[:, :, 1] = py
    yref[:, :, 2] = pth

    We = None
    if terminal_hack:
        same = ((yref[:, n, 0] == yref[:, n - 1, 0]) & (yref[:, n, 1] == yref[:, n - 1, 1])
                & (yref[:, n, 2] == yref[:, n - 1, 2]))
        q = torch.tensor(spec.Q, **f64)
        We = q[None, :].repeat(count, 1)
        We[:, :3] = torch.where(same[:, None], 100.0 * q[None, :3], q[None, :3])
    return dict(x0=x0, yref=yref, We=We)
