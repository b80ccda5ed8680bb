"""N > 1 host logic on CPU: two gloo ranks build their shards of a batch from the counter-based
generator; the union must equal the single-process batch bit for bit, and the max-over-ranks
reduction used for timing must behave."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total, name, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nmpc_nav_control_b200 import synth
    from nmpc_nav_control_b200.problem import MODELS
    from nmpc_nav_control_b200.shard import shard_range
    a, b = shard_range(total, rank, world)
    inst = synth.make_instances(MODELS[name], a, b - a, pose_only=True)
    chk = torch.tensor([inst["x0"].sum().item(), inst["yref"].sum().item(), float(b - a)], dtype=torch.float64)
    dist.all_reduce(chk, op=dist.ReduceOp.SUM)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)      # stands in for a per-rank elapsed time
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        q.put((chk.numpy().copy(), float(t.item())))
    # every rank also returns its shard so the parent can compare bitwise
    torch.save(dict(x0=inst["x0"], yref=inst["yref"], a=a, b=b), os.path.join(os.environ["SHARD_TMP"], f"shard{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_shards_equal_single_process(tmp_path):
    sys.path.insert(0, ROOT)
    from nmpc_nav_control_b200 import synth
    from nmpc_nav_control_b200.problem import MODELS
    total, name, world = 1001, "diff", 2          # odd total: ragged shards
    os.environ["SHARD_TMP"] = str(tmp_path)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    chk, tmax = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    full = synth.make_instances(MODELS[name], 0, total, pose_only=True)
    s0 = torch.load(os.path.join(tmp_path, "shard0.pt")); s1 = torch.load(os.path.join(tmp_path, "shard1.pt"))
    assert (s0["a"], s0["b"], s1["a"], s1["b"]) == (0, 501, 501, 1001)
    assert torch.equal(torch.cat([s0["x0"], s1["x0"]]), full["x0"])
    assert torch.equal(torch.cat([s0["yref"], s1["yref"]]), full["yref"])
    assert chk[2] == total and tmax == 2.0
    assert abs(chk[0] - full["x0"].sum().item()) < 1e-9


def test_shard_range_and_mixed_split():
    from nmpc_nav_control_b200.shard import mixed_split, shard_range
    for total in (0, 1, 7, 65536, 1048576):
        for world in (1, 2, 4, 8):
            edges = [shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    m = mixed_split(1048576)          # 349,526 / 349,525 / 349,525
    assert m == {"omni4": 349526, "diff": 349525, "tric": 349525}


def _worker_paths(rank, world, port, total, tmp):
    """SURVEY 8(f2) under sharding: each rank discretises the paths of its contiguous block of robots (the device code
    compiled for the host) - robots are independent, the shards need no exchange"""
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import emul
    import pathcases
    from nmpc_nav_control_b200.shard import shard_range
    paths, pid, u0 = pathcases.cases(seed=17, n_paths=9, B=total)           # every rank derives the same fleet from the seed
    off = np.cumsum([0] + [len(p) for p in paths]).astype(np.int32)
    a, b = shard_range(total, rank, world)
    out = emul.emul_path_discretize(np.concatenate(paths), off, pid[a:b], u0[a:b], 0.025, 81)
    n = torch.tensor([float(b - a)], dtype=torch.float64)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    assert n.item() == total
    np.save(os.path.join(tmp, f"poses{rank}.npy"), out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_path_discretiser_shards_equal_single_process(tmp_path):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import emul
    import pathcases
    emul.build()                                       # build once here, not concurrently in the workers
    total, world = 77, 2
    ctx = mp.get_context("spawn")
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_paths, args=(r, world, port, total, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    paths, pid, u0 = pathcases.cases(seed=17, n_paths=9, B=total)
    off = np.cumsum([0] + [len(p) for p in paths]).astype(np.int32)
    full = emul.emul_path_discretize(np.concatenate(paths), off, pid, u0, 0.025, 81)
    got = np.concatenate([np.load(os.path.join(tmp_path, f"poses{r}.npy")) for r in range(world)], axis=2)
    assert np.array_equal(got, full)
