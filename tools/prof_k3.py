"""one cold RTI solve of B instances (default diff) — the command profiled under ncu"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nmpc_nav_control_b200 import synth
from nmpc_nav_control_b200.problem import MODELS
from nmpc_nav_control_b200.solver import BatchedRtiSolver

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
name = sys.argv[2] if len(sys.argv) > 2 else "diff"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
spec = MODELS[name]
inst = synth.make_instances(spec, 0, B, device="cuda", pose_only=True)
x0 = inst["x0"].t().contiguous(); yref = inst["yref"].permute(1, 2, 0).contiguous()
s = BatchedRtiSolver(spec, B)
if os.environ.get("PROF_ITER_MAX"):          # experiment: cap the interior-point iterations (bounds the straggler tail)
    s.set_opts(iter_max=int(os.environ["PROF_ITER_MAX"]))
for _ in range(reps):
    s.reset_async()
    out = s.solve_device(x0, yref)
    torch.cuda.synchronize()
    t = s.last_timing()
print(name, B, "status!=0:", int((out["status"] != 0).sum()), "mean iter", out["qp_iter"].double().mean().item(),
      "max iter", out["qp_iter"].max().item(), t, f"{B / (t['total_ms'] * 1e-3):.0f} solves/s")
