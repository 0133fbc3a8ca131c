#!/usr/bin/env python
"""bench.py — headline benchmark of the WVA optimization hot path on B200.

Metric (BASELINE.json): (model,variant,replica) evals/sec + solver wall-ms, next to the reference
algorithm on the box's host cores.  One "step" = one pass of the hot path over one batch of synthetic
input: System.Calculate (sizing of every (server, accelerator) candidate) -> Manager.Optimize
(per-server cheapest feasible candidate + by-type totals) -> the (server, accelerator, replica) grid
of QueueAnalyzer.Analyze evaluations.  An "evaluation" = one chain solve; a step's
evaluations = S*A*R grid evaluations + the sizing solves counted on the device.

Workload at 1 GPU = BASELINE.json configs[1]: 1k models x 16 variants x 128 replica levels, N = 128
(K = 1408 states), 3 service classes, unlimited.  With --gpus N the model set is sharded (weak scaling:
every rank owns 1k models) and the per-shard partials are all-reduced with NCCL.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "llm-d-workload-variant-autoscaler_b200"

S_PER_GPU, A, NB, R = 1000, 16, 128, 128
ALG_BYTES_PER_EVAL = 17.9   # 16 B written + amortised inputs per grid evaluation
METRIC = "(model,variant,replica) evals/sec"


def workload(rank: int, world: int):
    synth = importlib.import_module(PKG + ".synth")
    # stream 2 = BASELINE config 2; other ranks draw their own models from the same generator family
    return synth.queue_system(S_PER_GPU, A, NB, n_classes=3, stream=2 + 100 * rank, R=R)


def config_dict(world):
    return {"workload": f"BASELINE configs[1]: {S_PER_GPU} models x {A} variants x {R} replica levels per GPU, "
                        f"state-dependent M/M/1/K sizing + replica grid, N={NB}, K={11 * NB}, 3 service classes, "
                        "unlimited allocator",
            "models_per_gpu": S_PER_GPU, "variants": A, "replica_levels": R, "max_batch": NB,
            "chain_states": 11 * NB + 1, "parallelism": f"model-sharded x{world}",
            "l2": "flushed between steps (256 MiB write inside the timed region)",
            "evaluation": "one chain solve: S*A*R grid Analyze calls + sizing solves counted on device"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax = max(smax, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_leg(sample_servers: int):
    """The reference algorithm (oracle port: the reference is Go and cannot be built here) on the host
    cores, on a bounded sample of the same workload."""
    from tests import oracle_lib
    orc = oracle_lib.load()
    synth = importlib.import_module(PKG + ".synth")
    d = synth.queue_system(sample_servers, A, NB, n_classes=3, stream=2, R=R)
    cores = orc.num_threads()
    t0 = time.perf_counter()
    cand = orc.calculate(d, nthreads=cores)
    t1 = time.perf_counter()
    sol = orc.solve(d, cand)
    t2 = time.perf_counter()
    orc.analyze_grid(d, R, nthreads=cores, full=True)
    t3 = time.perf_counter()
    evals = sample_servers * A * R + cand["_solves"]
    return {"value": evals / (t3 - t0), "unit": "evals/s", "cores": cores, "kind": "port",
            "sample": f"{sample_servers} of {S_PER_GPU} models x {A} variants x {R} levels (N={NB}), "
                      f"oracle C++ restatement, OpenMP over servers",
            "calculate_ms": (t1 - t0) * 1e3, "solve_ms": (t2 - t1) * 1e3, "grid_ms": (t3 - t2) * 1e3,
            "evals": int(evals), "seconds": t3 - t0}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tests import oracle_lib
    oracle_lib.load()
    vals = []
    last = None
    for i in range(args.warmup + args.steps):
        last = cpu_reference_leg(args.ref_sample)
        if i >= args.warmup:
            vals.append(last)
    secs = sum(v["seconds"] for v in vals)
    evals = sum(v["evals"] for v in vals)
    value = evals / secs
    line = {"metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": secs / len(vals) * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args.gpus),
            "impl": "reference",
            "cpu_baseline": {"value": value, "unit": "evals/s", "cores": last["cores"], "kind": "port",
                             "sample": last["sample"]},
            "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "solver_wall_ms": {"calculate": last["calculate_ms"], "solve": last["solve_ms"], "grid": last["grid_ms"],
                               "note": "on the bounded sample"}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-sample", type=int, default=100, help="servers per step in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    pkg = importlib.import_module(PKG)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (there is no CPU fallback); --impl reference runs on the host")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    sysd = workload(rank, world)
    eng = pkg.Engine(device=local)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    partial = torch.zeros(2 * sysd["n_types"] + 2, dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step():
        """inputs already in HBM: sizing + allocator + grid; partials all-reduced across shards"""
        flush.zero_()
        eng.calculate()
        t = eng.timing()
        solves, states, calc_ms = t["chain_solves"], t["chain_states"], t["calculate_ms"]
        eng.solve()
        solve_ms = eng.timing()["solve_ms"]
        eng.grid_run(R, full=True)
        t = eng.timing()
        if world > 1:
            dist.all_reduce(partial)   # per-shard capacity / cost partials (refreshed by e2e_step's fetch)
        return dict(size_solves=solves, size_states=states, calc_ms=calc_ms, solve_ms=solve_ms,
                    grid_ms=t["grid_ms"], grid_solves=t["chain_solves"], grid_states=t["chain_states"])

    def e2e_step():
        """the call a user makes, host buffers in, host results out (H2D and D2H inside)"""
        sol = eng.optimize(sysd)
        eng.grid_run(R, full=False)
        fr = eng.grid_fetch_frontier()
        if world > 1:
            p = torch.from_numpy(np.concatenate([sol["type_count"].astype(np.float64), sol["type_cost"],
                                                 [float((sol["state"] == 1).sum()), float(fr.sum())]])).to(dev)
            dist.all_reduce(p)
            p = p.cpu()
        return sol, fr

    h2d = sum(np.asarray(v).nbytes for k, v in sysd.items() if isinstance(v, np.ndarray))
    S = sysd["n_servers"]
    d2h = S * (1 + 9 * 4) + sysd["n_types"] * 16 + S * A * 4

    eng.load_system(sysd)
    for _ in range(args.warmup):
        info = resident_step()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    launches0 = eng.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    infos = []
    for _ in range(args.steps):
        infos.append(resident_step())
    e1.record()
    barrier()
    launches = eng.launch_count() - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    clocks = sampler.stop() if sampler else None

    evals_local = sum(S * A * R + i["size_solves"] for i in infos)
    ev = torch.tensor([float(evals_local)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ev)
    value = float(ev.item()) / (total_ms * 1e-3)

    # end-to-end arm
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        sol, fr = e2e_step()
        size_solves = eng.timing()  # (grid counters overwrite; recomputed below)
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = float(ev.item()) / (float(ms2.item()) * 1e-3)

    if rank == 0:
        last = infos[-1]
        grid_ms = float(np.mean([i["grid_ms"] for i in infos]))
        calc_ms = float(np.mean([i["calc_ms"] for i in infos]))
        solve_ms = float(np.mean([i["solve_ms"] for i in infos]))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        dominant = "grid_kernel" if grid_ms >= calc_ms else "sizer_kernel"
        dom_ms = max(grid_ms, calc_ms)
        if dominant == "grid_kernel":
            alg_bytes = S * A * R * ALG_BYTES_PER_EVAL
        else:
            alg_bytes = S * A * (24 + 36.0 / A + 37)
        achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
        dfma, ddiv = eng.microbench_fp64()
        # FP64 work actually executed: 5 pipe ops per pass-1 state, 12 per pass-2 state
        fp64_ops = 8.5 * (last["grid_states"] if dominant == "grid_kernel" else last["size_states"])
        line = {
            "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(world),
            "solver_wall_ms": {"calculate": calc_ms, "solve": solve_ms, "grid": grid_ms,
                               "calculate_plus_solve": calc_ms + solve_ms},
            "evals_per_step": {"grid": S * A * R, "sizing_solves": int(last["size_solves"]),
                               "states_grid": int(last["grid_states"]), "states_sizing": int(last["size_states"])},
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": None,
                         "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                         "note": "the path is FP64-latency/throughput bound, not HBM bound (SURVEY §0.4): see fp64"},
            "fp64": {"kernel": dominant, "achieved_ops_per_s": fp64_ops / (dom_ms * 1e-3),
                     "peak_dfma_per_s": dfma, "peak_ddiv_per_s": ddiv,
                     "frac_of_dfma_peak": fp64_ops / (dom_ms * 1e-3) / dfma,
                     "note": "pipe ops = 8.5 x states visited (5 in pass 1, 12 in pass 2); peaks measured in this run"},
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": float(ms2.item()) / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "host": {"nproc": os.cpu_count()},
        }
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = {k: v for k, v in cpu_reference_leg(max(args.ref_sample, 400)).items()
                                    if k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
