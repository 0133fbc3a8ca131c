"""Limited-capacity Manager.Optimize over model shards on N GPUs (torchrun, NCCL): every rank sizes its shard, the
candidates are all-gathered on the device, every rank runs the greedy sweep on the merged set.  Checks the result
against the whole system solved on one GPU and prints one JSON line with the timings.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
       tools/run_sharded_limited.py [S=20000] [A=16] [N=128] [policy=None]"""
import importlib, json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("llm-d-workload-variant-autoscaler_b200")

S = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
A = int(sys.argv[2]) if len(sys.argv) > 2 else 16
N = int(sys.argv[3]) if len(sys.argv) > 3 else 128
policy = sys.argv[4] if len(sys.argv) > 4 else "None"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

d = pkg.synth.queue_system(S, A, N, stream=3, saturation_policy=policy)
with pkg.Engine(local) as e:
    e.load_system(d); e.calculate(); e.solve()
    lim = pkg.synth.limit_capacity(d, e.solution()["type_count"], 0.6)

    def whole():
        e.load_system(lim); e.calculate(); e.solve()
        return e.solution()

    steps = {}

    def sharded():
        return pkg.sharding.solve_sharded(e, lim, rank, world, device=dev, timings=steps)

    out = {}
    for name, fn in (("whole_1gpu", whole), ("sharded", sharded)):
        ts = []
        for it in range(4):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            sol = fn()
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        t = torch.tensor([min(ts[1:])], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name + "_wall_ms"] = float(t.item())
        out[name] = sol
    same = all(np.array_equal(np.asarray(out["whole_1gpu"][k]).view(np.uint8), np.asarray(out["sharded"][k]).view(np.uint8)) for k in out["whole_1gpu"])
    flag = torch.tensor([1 if same else 0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        st = np.asarray(out["sharded"]["state"])
        print(json.dumps({"servers": S, "accelerators": A, "N": N, "policy": policy, "n_gpus": world,
                          "whole_1gpu_wall_ms": out["whole_1gpu_wall_ms"], "sharded_wall_ms": out["sharded_wall_ms"],
                          "identical_on_all_ranks": bool(flag.item()), "allocated": int((st == 1).sum()), "unallocated": int((st == 0).sum()),
                          "gathered_bytes_per_rank": ((S + world - 1) // world) * A * 37,
                          "sharded_steps_ms_rank0": {k: round(v, 3) for k, v in steps.items()}}))
if world > 1:
    dist.destroy_process_group()
