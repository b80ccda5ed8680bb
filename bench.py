#!/usr/bin/env python
"""Benchmark of the SQP-RTI hot path (BASELINE.json metric: batched OCP solves/sec, fp64).

A "step" = one SQP-RTI iteration of every instance of the batch from the reset (all-zero)
iterate: K1+K2 linearise, K3 interior-point QP (hybrid schedule: lockstep horizon sweeps while most
instances iterate, then the persistent lane-cooperative kernel for the rest), K4 step.  Workload at N=1: BASELINE config 2
(diff model, 65,536 random initial states / reference paths, SURVEY.md Appendix D inputs).
With N>1 every rank solves its own 65,536-instance shard (weak scaling, no collective on the
solve path; torch.distributed is only used for the barrier and the max-over-ranks time).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python bench.py --impl reference ...      # the CPU restatement (oracle) on the host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep stdout to the one JSON line

METRIC = "batched OCP solves/sec (fp64 SQP-RTI)"
UNIT = "solves/s"
WORKLOAD = "diff model batched SQP-RTI, 65,536 random initial states/reference paths per GPU (BASELINE config 2)"
MODEL = "diff"
BATCH = 65536
# algorithmic bytes and dense-equivalent flops per RTI solve
N_STAGES = 80


def bench_config(batch: int) -> dict:
    """the workload description, identical for both arms (`--impl ours` / `--impl reference`)"""
    return {"workload": WORKLOAD, "robot_model": MODEL, "batch_per_gpu": batch, "N": N_STAGES,
            "iterate": "reset (zero) before every step"}


def host_cores() -> int:
    """cores this process may run on (torch.distributed.run exports OMP_NUM_THREADS=1: the CPU arm sets its own count)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


ALG_BYTES = {"diff": 17512, "tric": 17512, "omni4": 29160}
F_LIN = {"diff": 3.73e5, "tric": 3.73e5, "omni4": 1.37e6}
F_ITER = {"diff": 2.69e5, "tric": 2.69e5, "omni4": 8.76e5}


def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.gpu = gpu
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [a.strip() for a in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(model: str, sample: int, nthreads: int = 0, repeats: int = 1):
    nthreads = nthreads or host_cores()
    """the oracle (kind = "port": acados-algorithm restatement, not acados) on the host cores,
    one solve per core via OpenMP"""
    from nmpc_nav_control_b200 import synth
    from nmpc_nav_control_b200.problem import MODELS
    from oracle import orc
    spec = MODELS[model]
    o = orc.Oracle(model, spec.codegen_defaults(), fast=True)
    inst = synth.make_instances(spec, 0, sample)
    x0 = inst["x0"].numpy().copy(); yref = inst["yref"].numpy().copy()
    best = None
    for _ in range(repeats):
        x = np.zeros((sample, spec.n + 1, spec.nx)); u = np.zeros((sample, spec.n, spec.nu))
        t0 = time.perf_counter()
        r = o.rti_batch(x0, yref, x, u, nthreads=nthreads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return dict(value=sample / best, cores=int(r["threads"]), seconds=best, mean_qp_iter=float(r["qp_iter"].mean()))


def run_reference(args):
    """--impl reference: the reference algorithm's CPU implementation (oracle port; real acados is
    not installable here) on all host cores of the box, bounded sample per step.
    Under torchrun rank 0 alone runs it (with every core of the box, whatever OMP_NUM_THREADS says)."""
    rank, world, _ = _dist_env()
    if rank != 0:
        return
    cores = host_cores()
    sample = int(min(args.batch, max(16384, 512 * cores)))
    cpu_baseline(MODEL, min(sample, 1024), nthreads=cores)          # warm-up / page-in
    for _ in range(max(0, args.warmup - 1)):
        cpu_baseline(MODEL, sample, nthreads=cores)
    t = []
    r = None
    for _ in range(args.steps):
        r = cpu_baseline(MODEL, sample, nthreads=cores)
        t.append(r["seconds"])
    sec = float(np.mean(t))
    val = sample / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args.batch),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": f"{sample} diff instances per step (same generator/seed as the GPU arm), OpenMP one solve per core, "
                                   f"{r['cores']} threads set explicitly"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "latency": cpu_latency(MODEL, 200),
    }
    print(json.dumps(line), flush=True)


def cpu_latency(model: str, calls: int) -> dict:
    """one RTI solve on one core through the oracle port (what `{m}_acados_solve` costs the reference per tick)"""
    from nmpc_nav_control_b200 import synth
    from nmpc_nav_control_b200.problem import MODELS
    from oracle import orc
    spec = MODELS[model]
    o = orc.Oracle(model, spec.codegen_defaults(), fast=True)
    inst = synth.make_instances(spec, 0, 1)
    x0 = inst["x0"].numpy().copy(); yref = inst["yref"].numpy().copy()
    ts = []
    for i in range(calls + 20):
        x = np.zeros((1, spec.n + 1, spec.nx)); u = np.zeros((1, spec.n, spec.nu))
        t0 = time.perf_counter()
        o.rti_batch(x0, yref, x, u, nthreads=1)
        ts.append((time.perf_counter() - t0) * 1e6)
    ts = np.array(ts[20:])
    return {"what": f"one cold-iterate {model} RTI solve, oracle port, 1 thread", "calls": calls,
            "p50_us": float(np.percentile(ts, 50)), "p95_us": float(np.percentile(ts, 95)), "p99_us": float(np.percentile(ts, 99))}


def gpu_latency(model: str, device: int, calls: int = 1000, warm: int = 100) -> dict:
    """host wall clock around a batch = 1 solve through the C ABI (H2D of x0 / yref, launches, D2H of
    u_0 / x_1 / status, sync) - what the ROS drop-in sees per tick; cold iterate (the worst case of a tick)"""
    from nmpc_nav_control_b200 import synth
    from nmpc_nav_control_b200.problem import MODELS
    from nmpc_nav_control_b200.solver import BatchedRtiSolver
    spec = MODELS[model]
    inst = synth.make_instances(spec, 0, 1, pose_only=True)
    x0 = inst["x0"].numpy().copy(); yref = inst["yref"].numpy().copy()
    s = BatchedRtiSolver(spec, 1, device=device)
    out = None
    for _ in range(warm):
        s.reset()
        out = s.solve_host(x0, yref, out=out)
    ts = []
    for _ in range(calls):
        s.reset()
        t0 = time.perf_counter()
        out = s.solve_host(x0, yref, out=out)
        ts.append((time.perf_counter() - t0) * 1e6)
    ts = np.array(ts)
    r = {"what": f"batch = 1 cold-iterate {model} RTI solve through nmpc_rti_solve_host (C ABI, host buffers)", "calls": calls,
         "warmup": warm, "p50_us": float(np.percentile(ts, 50)), "p95_us": float(np.percentile(ts, 95)),
         "p99_us": float(np.percentile(ts, 99)), "qp_iter": int(out["qp_iter"][0])}
    s.close()
    return r


def fp64_executed(qp_s: float, peak_tflops: float, batch: int):
    """executed fp64 flops of K3 (ncu smsp__sass_thread_inst_executed_op_d{add,mul,fma}_pred_on summed over the K3 launches
    of one step, fma x 2; profiles/k3_exec_flops.json, made by tools/ncu_flops.py) / K3 time of THIS run"""
    try:
        with open(os.path.join(ROOT, "profiles", "k3_exec_flops.json")) as f:
            j = json.load(f)
        if j.get("model") != MODEL or j.get("batch") != batch:
            return None
        fl = float(j["flops_per_step"])
        return {"kernel": "K3", "bound": "fp64", "achieved": fl / qp_s / 1e12, "peak": peak_tflops, "unit": "TFLOP/s",
                "frac": fl / qp_s / 1e12 / peak_tflops if peak_tflops > 0 else None,
                "note": "executed dadd + dmul + 2 dfma thread instructions (ncu counters of the committed capture, "
                        "profiles/k3_exec_flops.json) / K3 time of this run; the dense-equivalent figure is roofline_fp64"}
    except Exception:
        return None


def extra_configs(rank: int, world: int, local: int, barrier) -> dict:
    """BASELINE configs 3 and 5 under the same clock (short: 2 warm-ups + 3 steps each): omni4 262,144 on one GPU
    (rank 0 only) and the mixed omni4 / diff / tric 1,048,576-instance batch cut into contiguous shards over the ranks."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_models
    from nmpc_nav_control_b200.shard import shard_range
    out = {}
    if world == 1:
        r = bench_models.throughput("omni4", 262144, steps=3, warm=2)
        out["omni4_262144"] = {"value": r["solves_per_s"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "n_gpus": 1,
                               "mean_qp_iter": r["mean_qp_iter"], "status_nonzero": r["status_nonzero"], "steps": 3}
    total = 1048576
    lo, hi = shard_range(total, rank, world)
    r = bench_models.mixed(hi - lo, steps=3, warm=2, device=local, start=lo // 3, sync=barrier)
    t = torch.tensor([r["ms_per_step"]], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    out["mixed_1M"] = {"value": total / ms * 1e3, "unit": UNIT, "ms_per_step": ms, "n_gpus": world, "scaling": "strong",
                       "per_gpu": hi - lo, "steps": 3, "status_nonzero": int(sum(r["status_nonzero"].values())),
                       "timing": "wall clock around barrier + synchronize, max over ranks"}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from nmpc_nav_control_b200 import synth
    from nmpc_nav_control_b200.problem import MODELS
    from nmpc_nav_control_b200.solver import BatchedRtiSolver, dfma_peak_tflops

    rank, world, local = _dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    spec = MODELS[MODEL]
    B = args.batch
    dev = torch.device("cuda", local)

    # inputs of this rank's shard, created on the device (instance index = global index)
    inst = synth.make_instances(spec, rank * B, B, device=dev, pose_only=True)
    x0_soa = inst["x0"].t().contiguous()
    yref_soa = inst["yref"].permute(1, 2, 0).contiguous()
    x0_pin = inst["x0"].cpu().pin_memory()
    yref_pin = inst["yref"].cpu().pin_memory()
    solver = BatchedRtiSolver(spec, B, device=local)
    out = dict(status=torch.empty(B, dtype=torch.int32, device=dev), qp_iter=torch.empty(B, dtype=torch.int32, device=dev))
    hout = dict(u0=torch.empty(B, spec.nu, dtype=torch.float64).pin_memory().numpy(),
                x1=torch.empty(B, spec.nx, dtype=torch.float64).pin_memory().numpy(),
                status=torch.empty(B, dtype=torch.int32).pin_memory().numpy(),
                qp_iter=torch.empty(B, dtype=torch.int32).pin_memory().numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        solver.reset_async()
        solver.solve_device(x0_soa, yref_soa, out=out)

    def step_host():
        solver.reset()
        solver.solve_host(x0_pin.numpy(), yref_pin.numpy(), out=hout)

    def timed(fn, steps):
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t.item())

    def timed_events(fn, steps):
        """device time with CUDA events on the launching stream, max over ranks"""
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t.item())

    # the solver launches on torch's current stream, so torch events bracket it
    for _ in range(max(3, args.warmup)):
        step_device()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed_events(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)
    status_bad = int((out["status"] != 0).sum().item())
    mean_iter = float(out["qp_iter"].double().mean().item())
    hist = torch.bincount(out["qp_iter"].long().clamp(min=0)).cpu().tolist()
    qp_iter_hist = {str(i): int(c) for i, c in enumerate(hist) if c}
    launches = solver.last_launches() * args.steps

    # per-kernel times (second pass, one event read-out per step) for the roofline of K3
    kt = {"linearize_ms": 0.0, "qp_ms": 0.0, "step_ms": 0.0, "total_ms": 0.0}
    nk = min(args.steps, 10)
    for _ in range(nk):
        step_device()
        t = solver.last_timing()
        for k in kt:
            kt[k] += t[k] / nk

    # end-to-end through the host-buffer call: H2D of x0/yref and D2H of u0,x1,status inside
    for _ in range(2):
        step_host()
    e2e_steps = max(3, min(args.steps, 10))
    ms_e2e = timed(step_host, e2e_steps) / e2e_steps
    e2e_value = world * B / (ms_e2e * 1e-3)
    h2d = B * (spec.nx + (spec.n + 1) * 3) * 8
    d2h = B * (spec.nu + spec.nx) * 8 + B * 8

    solver.close()                       # frees the workspaces before the larger configurations
    extras = None
    if not args.no_extra:
        extras = extra_configs(rank, world, local, barrier)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    latency = None
    if not args.no_extra:
        latency = gpu_latency(MODEL, local)
        if not args.no_cpu:
            latency["cpu_port"] = cpu_latency(MODEL, 200)
        # the other two models of the package, the same measurement (300 calls each)
        latency["other_models"] = {}
        for m in ("tric", "omni4"):
            g = gpu_latency(m, local, calls=300, warm=30)
            rec = {"p50_us": g["p50_us"], "p99_us": g["p99_us"], "qp_iter": g["qp_iter"]}
            if not args.no_cpu:
                rec["cpu_port_p50_us"] = cpu_latency(m, 60)["p50_us"]
            latency["other_models"][m] = rec

    # measured DRAM traffic of K3 per solve (ncu dram__bytes over all K3 launches of one step, see
    # tools/ncu_launches.py --json and profiles/): what the streaming roofline is computed from
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "k3_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("model") == MODEL and tj.get("batch") == B:
            traffic = float(tj["dram_bytes_per_step"])
    except Exception:
        pass
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md 6.65 TB/s)"
    fp64_peak = dfma_peak_tflops(local)
    qp_s = kt["qp_ms"] * 1e-3
    alg_gbs = B * ALG_BYTES[MODEL] / qp_s / 1e9
    fp64_ach = B * mean_iter * F_ITER[MODEL] / qp_s / 1e12

    # CPU baseline (oracle port) on this box's cores, bounded sample
    cpu = None
    if not args.no_cpu:
        sample = BATCH                                  # the whole workload: ~10 s of CPU work for the three passes
        cpu_baseline(MODEL, 1024)
        r = cpu_baseline(MODEL, sample, repeats=3)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": f"{sample} of the same diff instances, cold iterate, OpenMP one solve per core, best of 3 "
                         f"({r['seconds']:.2f} s); acados-algorithm restatement (oracle), not acados"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": bench_config(B),
        "run": {"l2": "inputs (131 MB) and the K3 workspace (GBs of per-instance interior-point records streamed by every sweep) "
                      "exceed the 126 MB L2 many times over; no explicit flush",
                "k3_schedule": os.environ.get("NMPC_K3", "default"), "mean_qp_iter": mean_iter, "status_nonzero": status_bad,
                "qp_iter_hist": qp_iter_hist},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e, "api": "nmpc_rti_solve_host (C ABI, pinned host buffers)"},
        "gpu_launches": launches,
        "kernel_ms": kt,
        "roofline": {"kernel": "K3 interior point = the k_sweep<B_FIRST|FDF|B>, k_handover_* and k_ipm_coop launches of one step", "bound": "hbm",
                     "achieved": alg_gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": alg_gbs / hbm_peak, "traffic": traffic, "peak_source": hbm_src,
                     "note": "achieved = algorithmic bytes (17,512 B/solve, SURVEY 8d) x instances / K3 time per step "
                             "(CUDA events inside the library, on the launching stream); traffic = measured DRAM bytes of "
                             "the same launches (ncu launch list under profiles/): K3 streams the 85 KB interior-point state "
                             "of every instance through HBM each sweep, see roofline_stream and DESIGN.md 5"},
        "roofline_stream": None if traffic is None else {
            "kernel": "K3", "bound": "hbm", "achieved": traffic / qp_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": traffic / qp_s / 1e9 / hbm_peak,
            "note": "measured DRAM traffic of the K3 launches (ncu, profiles/k3_traffic.json) / K3 time of this run"},
        "roofline_fp64": {"kernel": "K3", "bound": "fp64", "achieved": fp64_ach, "peak": fp64_peak, "unit": "TFLOP/s",
                          "frac": fp64_ach / fp64_peak if fp64_peak > 0 else None,
                          "peak_source": "self-measured DFMA micro-benchmark (MEASURED_PEAKS.json has no fp64 figure)",
                          "note": "achieved = dense-equivalent flops (2.69e5 per IPM iteration, SURVEY 8d) x measured mean iterations"},
        "roofline_fp64_executed": fp64_executed(qp_s, fp64_peak, B),
        "cpu_baseline": cpu,
        "latency": latency,
        "configs": extras,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the latency block and the config 3 / 5 sub-records")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
