"""Model-sharded limited / unlimited solve and V1 saturation over N ranks THROUGH THE C-ABI (NCCL inside the library).
Run under torchrun:  python -m torch.distributed.run --nproc-per-node N tools/run_multi_gpu.py [--servers S] [--check]
Prints one JSON line from rank 0: per-phase device times (max over ranks), exchange times, and — with --check — whether
every rank's solution equals a whole-system run on one GPU bit for bit."""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

pkg = importlib.import_module("llm-d-workload-variant-autoscaler_b200")

ap = argparse.ArgumentParser()
ap.add_argument("--servers", type=int, default=20000)
ap.add_argument("--acc", type=int, default=16)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--check", action="store_true")
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
eng = pkg.Engine(local)
idt = torch.zeros(128, dtype=torch.uint8, device=dev)
if rank == 0:
    idt.copy_(torch.frombuffer(bytearray(pkg.comm_unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
eng.comm_init(world, rank, idt.cpu().numpy().tobytes())

d = pkg.synth.queue_system(args.servers, args.acc, args.batch, stream=97)
eng.load_system(d); eng.calculate(); eng.set_optimizer(True); eng.solve()
un = eng.solution()
cap = np.maximum(1, np.floor(np.asarray(un["type_count"], np.float64) * 0.6)).astype(np.int32)
rows = []
for _ in range(args.reps):
    dist.barrier()
    eng.calculate(); c = eng.timing()["calculate_ms"]
    eng.set_optimizer(True); eng.solve(); t1 = eng.timing()
    eng.set_capacity(cap); eng.set_optimizer(False, False, "None"); eng.solve(); t2 = eng.timing()
    rows.append([c, t1["solve_ms"], t1["exchange_ms"], t2["solve_ms"], t2["exchange_ms"]])
lim = eng.solution()
t = torch.tensor(np.min(np.array(rows), axis=0), dtype=torch.float64, device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
ok = True
if args.check:
    with pkg.Engine(local) as e1:                      # no communicator: the whole system on this GPU
        e1.load_system(d); e1.calculate(); e1.set_optimizer(True); e1.solve()
        a = e1.solution()
        e1.set_capacity(cap); e1.set_optimizer(False, False, "None"); e1.solve()
        b = e1.solution()
    for k in ("state", "acc", "num_replicas", "batch_size", "cost", "value", "itl", "ttft", "rho", "max_arrv_rate"):
        ok &= np.array_equal(np.asarray(a[k]).view(np.uint8), np.asarray(un[k]).view(np.uint8))
        ok &= np.array_equal(np.asarray(b[k]).view(np.uint8), np.asarray(lim[k]).view(np.uint8))
    ok &= np.array_equal(a["type_count"], un["type_count"]) and np.array_equal(b["type_count"], lim["type_count"])
    ok &= bool(np.allclose(a["type_cost"], un["type_cost"], rtol=1e-12))
    # V1 saturation: each rank its own block, partials all-reduced
    sb = pkg.synth.saturation_batch(2000, 32, stream=4 + rank)
    eng.saturation_upload(sb); eng.saturation_run(False)
    r = eng.saturation_fetch(fields=("partials", "partials_all"))
    p = torch.tensor(r["partials"].astype(np.float64), device=dev)
    dist.all_reduce(p)
    ok &= np.array_equal(p.cpu().numpy().astype(np.int64), r["partials_all"])
flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"world": world, "servers": args.servers, "acc": args.acc, "N": args.batch,
                      "calculate_ms": t[0].item(), "solve_unlimited_ms": t[1].item(), "exchange_unlimited_ms": t[2].item(),
                      "solve_limited_ms": t[3].item(), "exchange_limited_ms": t[4].item(),
                      "identical_to_one_gpu": bool(flag.item() == 1.0) if args.check else None}))
eng.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1.0 else 1)
