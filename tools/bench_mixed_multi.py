#!/usr/bin/env python
"""BASELINE config 5 across GPUs: the mixed omni4 / diff / tric batch of `--total` instances cut into contiguous shards,
one process per GPU (no collective on the solve path; NCCL only for the barrier and the max-over-ranks time).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 \\
        tools/bench_mixed_multi.py --total 1048576
Rank 0 prints one JSON line."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import torch.distributed as dist
import bench_models
from nmpc_nav_control_b200.shard import shard_range

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--total", type=int, default=1048576)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lo, hi = shard_range(a.total, rank, world)

    def sync():
        if world > 1:
            dist.barrier()
    r = bench_models.mixed(hi - lo, steps=a.steps, warm=2, device=local, start=lo // 3, sync=sync)
    t = torch.tensor([r["ms_per_step"]], dtype=torch.float64, device=f"cuda:{local}")
    bad = torch.tensor([float(sum(r["status_nonzero"].values()))], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(bad, op=dist.ReduceOp.SUM)
    if rank == 0:
        ms = float(t.item())
        print(json.dumps(dict(kind="mixed_multi_gpu", total=a.total, n_gpus=world, per_gpu=hi - lo, ms_per_step=ms,
                              solves_per_s=a.total / ms * 1e3, status_nonzero=int(bad.item()), scaling="strong",
                              timing="wall clock around barrier + synchronize, max over ranks", steps=a.steps)), flush=True)
    if world > 1:
        dist.destroy_process_group()
