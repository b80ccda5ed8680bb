#!/usr/bin/env python
"""Experiment: one batch solved as `parts` sub-batches, each with its own solver and CUDA stream, so that the
latency-bound tail of one sub-batch's lane-group phase overlaps the lockstep sweeps of another.
    python tools/exp_split.py diff 65536 1,2,3,4"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nmpc_nav_control_b200 import synth
from nmpc_nav_control_b200.problem import MODELS
from nmpc_nav_control_b200.solver import BatchedRtiSolver


def run(name, total, parts, steps=4, warm=2):
    spec = MODELS[name]
    dev = torch.device("cuda", 0)
    inst = synth.make_instances(spec, 0, total, device=dev, pose_only=True)
    x0 = inst["x0"].t().contiguous(); yref = inst["yref"].permute(1, 2, 0).contiguous()
    per = (total // parts + 31) // 32 * 32
    ctx = []
    for p in range(parts):
        a, b = p * per, min(total, (p + 1) * per)
        ctx.append(dict(x0=x0[:, a:b].contiguous(), yref=yref[:, :, a:b].contiguous(), s=BatchedRtiSolver(spec, b - a),
                        st=torch.cuda.Stream(dev), out=dict(status=torch.empty(b - a, dtype=torch.int32, device=dev),
                                                            qp_iter=torch.empty(b - a, dtype=torch.int32, device=dev))))
    torch.cuda.synchronize()

    def step():
        for c in ctx:
            c["s"].reset_async(c["st"]); c["s"].solve_device(c["x0"], c["yref"], out=c["out"], stream=c["st"])
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    it = sum(float(c["out"]["qp_iter"].double().sum()) for c in ctx) / total
    bad = sum(int((c["out"]["status"] != 0).sum()) for c in ctx)
    for c in ctx:
        c["s"].close()
    return dict(model=name, total=total, parts=parts, ms_per_step=round(ms, 2), solves_per_s=round(total / ms * 1e3), mean_qp_iter=it, bad=bad)


if __name__ == "__main__":
    name, total = sys.argv[1], int(sys.argv[2])
    for p in sys.argv[3].split(","):
        print(json.dumps(run(name, total, int(p))), flush=True)
