"""phase cycles of the block-per-instance kernel (a -DNMPC_SOLO_PROF build selected with NMPC_B200_LIB): one cold batch-1 solve"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nmpc_nav_control_b200 import synth, _lib
from nmpc_nav_control_b200.problem import MODELS
from nmpc_nav_control_b200.solver import BatchedRtiSolver
name = sys.argv[1] if len(sys.argv) > 1 else "diff"
spec = MODELS[name]
inst = synth.make_instances(spec, 0, 1, device="cuda", pose_only=True)
x0 = inst["x0"].t().contiguous(); yref = inst["yref"].permute(1, 2, 0).contiguous()
s = BatchedRtiSolver(spec, 1)
lib = _lib.load()
buf = (C.c_ulonglong * 24)()
for rep in range(3):
    s.reset_async(); out = s.solve_device(x0, yref); torch.cuda.synchronize()
    lib.nmpc_solo_prof(buf)
names = ["load", "adj_const", "adj_rec", "update", "reduce_B", "riccati", "gain", "fwd_pred", "step_pred", "reduce_F", "delta_rhs", "delta_bwd",
         "delta_ff", "delta_fwd", "step_delta", "reduce_Fd", "ric_A", "-", "-", "ric_B", "-", "-", "-", "-"]
tot = sum(buf)
print(name, "iters", int(out["qp_iter"][0]), "qp_ms", s.last_timing()["qp_ms"], "total cycles", tot)
for n, v in zip(names, buf):
    print(f"  {n:12s} {v:9d} cycles {v / tot:6.1%}")
