#!/usr/bin/env python
"""Secondary measurements (not the contract bench line): device-resident throughput of the three models at
several batch sizes, and the batch-1 latency through the host-buffer C ABI (what the ROS drop-in would see).
    python tools/bench_models.py [--latency-calls 1000]
Prints one JSON object per line."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nmpc_nav_control_b200 import synth
from nmpc_nav_control_b200.problem import MODELS
from nmpc_nav_control_b200.solver import BatchedRtiSolver


def throughput(name, B, steps=3, warm=2):
    spec = MODELS[name]
    dev = torch.device("cuda", 0)
    inst = synth.make_instances(spec, 0, B, device=dev, pose_only=True)
    x0 = inst["x0"].t().contiguous(); yref = inst["yref"].permute(1, 2, 0).contiguous()
    s = BatchedRtiSolver(spec, B)
    out = dict(status=torch.empty(B, dtype=torch.int32, device=dev), qp_iter=torch.empty(B, dtype=torch.int32, device=dev))
    for _ in range(warm):
        s.reset_async(); s.solve_device(x0, yref, out=out)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        s.reset_async(); s.solve_device(x0, yref, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    t = s.last_timing()
    r = dict(kind="throughput", model=name, batch=B, ms_per_step=ms, solves_per_s=B / ms * 1e3,
             mean_qp_iter=float(out["qp_iter"].double().mean()), max_qp_iter=int(out["qp_iter"].max()),
             status_nonzero=int((out["status"] != 0).sum()), kernel_ms=t)
    s.close()
    return r


def latency(name, calls):
    spec = MODELS[name]
    inst = synth.make_instances(spec, 0, 1, pose_only=True)
    x0 = inst["x0"].numpy().copy(); yref = inst["yref"].numpy().copy()
    s = BatchedRtiSolver(spec, 1)
    s.reset()
    out = None
    for _ in range(100):
        out = s.solve_host(x0, yref, out=out)
    ts = []
    for _ in range(calls):
        s.reset()                      # cold iterate: the worst case of a tick (a warm tick needs fewer QP iterations)
        t0 = time.perf_counter()
        out = s.solve_host(x0, yref, out=out)
        ts.append((time.perf_counter() - t0) * 1e6)
    ts = np.array(ts)
    r = dict(kind="latency_batch1", model=name, calls=calls, p50_us=float(np.percentile(ts, 50)), p95_us=float(np.percentile(ts, 95)),
             p99_us=float(np.percentile(ts, 99)), qp_iter=int(out["qp_iter"][0]), api="nmpc_rti_solve_host (cold iterate)")
    s.close()
    return r


def ctrl_tick(name, B, steps=3, warm=2):
    """the batched controller tick (glue kernels + RTI step) on device-resident inputs, cold iterate
    every step like the bench line, and the same through the host-buffer call"""
    from nmpc_nav_control_b200.controller import BatchedNavController
    spec = MODELS[name]
    dev = torch.device("cuda", 0)
    inst = synth.make_instances(spec, 0, B, device=dev, pose_only=True)
    yref = inst["yref"].permute(1, 2, 0).contiguous()                 # [N+1, 3, B]: the reference poses
    pose = inst["x0"].t().contiguous()[:3].contiguous()
    # the bench workload seen from the controller's side: measured twist = inverse kinematics of the actuator states
    # (exact for diff and tric, least-squares for omni4's four wheels), carried reference states = x0[3+nv:]
    a = inst["x0"].t().contiguous()[3:3 + spec.nv]
    vref0 = inst["x0"].t().contiguous()[3 + spec.nv:].contiguous()
    vel = torch.zeros(3, B, dtype=torch.float64, device=dev)
    steer = None
    if name == "diff":
        vel[0] = (a[1] + a[0]) / 2.0; vel[2] = (a[1] - a[0]) / spec.p[0]
    elif name == "omni4":
        vel[0] = (a[0] - a[1] + a[2] - a[3]) / 4.0; vel[1] = (-a[0] - a[1] + a[2] + a[3]) / 4.0
        vel[2] = (-a[0] - a[1] - a[2] - a[3]) / (2.0 * spec.p[0])
    else:
        vel[0] = a[0]; steer = a[1].contiguous()
    th = yref[:, 2, :]
    refs = yref.clone(); refs[:, 2, :] = torch.atan2(torch.sin(th), torch.cos(th))      # wrapped, as a path planner gives them
    c = BatchedNavController(spec, B, dt=spec.dt)
    if steer is not None:
        c.set_steering_wheel_angle(steer)
    carried = c.reference_states()[:, :B]
    out = None

    def step():
        nonlocal out
        c.solver.reset_async(); carried.copy_(vref0)
        out = c.run(pose, vel, refs, out=out)
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    t = c.solver.last_timing()
    hp, hv, hr = pose.t().contiguous().cpu().numpy(), vel.t().contiguous().cpu().numpy(), refs.permute(2, 0, 1).contiguous().cpu().numpy()
    ho = None
    hs = []
    for _ in range(warm + steps):
        c.solver.reset(); carried.copy_(vref0); torch.cuda.synchronize()
        t0 = time.perf_counter()
        ho = c.run_host(hp, hv, hr, out=ho)
        hs.append((time.perf_counter() - t0) * 1e3)
    r = dict(kind="controller_tick", model=name, batch=B, ms_per_tick=ms, ticks_per_s=B / ms * 1e3, solver_ms=t,
             glue_ms=ms - t["total_ms"], host_ms_per_tick=float(np.mean(hs[warm:])), host_ticks_per_s=B / float(np.mean(hs[warm:])) * 1e3,
             status_nonzero=int((out["status"] != 0).sum()), mean_qp_iter=float(out["qp_iter"].double().mean()),
             h2d_bytes_per_tick=int(hp.nbytes + hv.nbytes + hr.nbytes + ho["cmd"].nbytes), d2h_bytes_per_tick=int(ho["cmd"].nbytes + 8 * B))
    c.close()
    return r


def pathdisc(B, steps=20, warm=3):
    """N+1 reference poses for B robots on 512 seeded paths (lines, arcs, cubic Beziers)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import pathcases
    from nmpc_nav_control_b200 import paths as P
    paths, pid, _ = pathcases.cases(seed=9, n_paths=512, B=B)
    rng = np.random.default_rng(1)
    u0 = rng.uniform(0, 1, B) * np.array([len(paths[p]) for p in pid])
    ps = P.PathSet(paths)
    d = P.BatchedPathDiscretizer(0.025, 81, False)
    tp, tu = torch.from_numpy(pid).cuda(), torch.from_numpy(u0).cuda()
    out = d.get_next_n_poses(ps, tp, tu)
    for _ in range(warm):
        d.get_next_n_poses(ps, tp, tu, out=out)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        d.get_next_n_poses(ps, tp, tu, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return dict(kind="path_discretise", robots=B, poses=81, ms_per_call=ms, robots_per_s=B / ms * 1e3,
                out_bytes=int(out.numel() * 8), out_GBps=out.numel() * 8 / ms / 1e6)


def mixed(total, steps=2, warm=1, device=0, start=0, sync=None):
    """BASELINE config 5 on one GPU: `total` instances split in thirds over omni4 / diff / tric, one solver and one
    CUDA stream per model so that the three launch sequences overlap; device-resident inputs.  `start`: first instance
    index of this GPU's shard (multi-GPU run), `sync`: barrier called before and after the timed region"""
    dev = torch.device("cuda", device)
    third = total // 3
    sizes = {"omni4": total - 2 * third, "diff": third, "tric": third}
    ctx = {}
    for name, B in sizes.items():
        spec = MODELS[name]
        inst = synth.make_instances(spec, start, B, device=dev, pose_only=True)
        ctx[name] = dict(B=B, x0=inst["x0"].t().contiguous(), yref=inst["yref"].permute(1, 2, 0).contiguous(),
                         s=BatchedRtiSolver(spec, B, device=device), st=torch.cuda.Stream(dev),
                         out=dict(status=torch.empty(B, dtype=torch.int32, device=dev), qp_iter=torch.empty(B, dtype=torch.int32, device=dev)))
    torch.cuda.synchronize()

    def step():
        for c in ctx.values():
            c["s"].reset_async(c["st"])
            c["s"].solve_device(c["x0"], c["yref"], out=c["out"], stream=c["st"])
    for _ in range(warm):
        step()
    torch.cuda.synchronize(dev)
    if sync:
        sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize(dev)
    if sync:
        sync()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    r = dict(kind="mixed", total=total, sizes=sizes, ms_per_step=ms, solves_per_s=total / ms * 1e3,
             status_nonzero={n: int((c["out"]["status"] != 0).sum()) for n, c in ctx.items()},
             mean_qp_iter={n: float(c["out"]["qp_iter"].double().mean()) for n, c in ctx.items()})
    for c in ctx.values():
        c["s"].close()
    return r


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--latency-calls", type=int, default=1000)
    ap.add_argument("--batches", default="diff:65536,diff:131072,tric:65536,omni4:65536,omni4:262144")
    ap.add_argument("--mixed", type=int, default=0, help="BASELINE config 5: total instances of the mixed omni4/diff/tric batch")
    ap.add_argument("--ctrl", default="", help="SURVEY 8(f1): controller tick, e.g. diff:65536,tric:65536")
    ap.add_argument("--paths", type=int, default=0, help="SURVEY 8(f2): path discretisation for this many robots")
    a = ap.parse_args()
    if a.paths:
        print(json.dumps(pathdisc(a.paths)), flush=True)
        sys.exit(0)
    if a.ctrl:
        for item in a.ctrl.split(","):
            n, b = item.split(":")
            print(json.dumps(ctrl_tick(n, int(b))), flush=True)
        sys.exit(0)
    if a.mixed:
        print(json.dumps(mixed(a.mixed)), flush=True)
        sys.exit(0)
    for item in a.batches.split(","):
        if not item:
            continue
        n, b = item.split(":")
        print(json.dumps(throughput(n, int(b))), flush=True)
    for n in ("diff", "omni4", "tric") if a.latency_calls > 0 else ():
        print(json.dumps(latency(n, a.latency_calls)), flush=True)
