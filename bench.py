#!/usr/bin/env python
"""Benchmark of the SQP-RTI hot path (BASELINE.json metric: batched OCP solves/sec, fp64).

A "step" = one SQP-RTI iteration of every instance of the batch from the reset (all-zero)
iterate: K1+K2 linearise, K3 interior-point QP (hybrid schedule: lockstep horizon sweeps while most
instances iterate, then the persistent lane-group kernel for the rest), K4 step.  Workload at N=1: BASELINE config 2
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
    not installable here) with all host threads, bounded sample per step."""
    rank, world, _ = _dist_env()
    if rank != 0:
        return
    from oracle import orc
    cores = orc.max_threads()
    sample = int(min(BATCH, max(16384, 512 * cores)))
    cpu_baseline(MODEL, min(sample, 1024))          # warm-up / page-in
    for _ in range(max(0, args.warmup - 1)):
        cpu_baseline(MODEL, sample)
    t = []
    r = None
    for _ in range(args.steps):
        r = cpu_baseline(MODEL, sample)
        t.append(r["seconds"])
    sec = float(np.mean(t))
    val = sample / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_per_step": sample, "robot_model": MODEL},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": f"{sample} diff instances per step (same generator/seed as the GPU arm), OpenMP one solve per core"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

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
        from oracle import orc
        cores = orc.max_threads()
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
        "config": {"workload": WORKLOAD, "robot_model": MODEL, "batch_per_gpu": B, "N": spec.n, "iterate": "reset (zero) before every step",
                   "l2": "inputs (131 MB) and the K3 workspace (5.6 GB tile state streamed by every sweep, 3 GB of group records) "
                         "exceed the 126 MB L2 many times over; no explicit flush",
                   "k3_schedule": os.environ.get("NMPC_K3", "hybrid"),
                   "mean_qp_iter": mean_iter, "status_nonzero": status_bad},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e, "api": "nmpc_rti_solve_host (C ABI, pinned host buffers)"},
        "gpu_launches": launches,
        "kernel_ms": kt,
        "roofline": {"kernel": "K3 interior point = the k_sweep<B_FIRST|FDF|B>, k_handover_* and k_ipm_group launches of one step", "bound": "hbm",
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
        "cpu_baseline": cpu,
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
