"""Host <-> device copy bandwidth with 1, 2, 4, 8 GPUs of one box copying at the same time (pinned memory, one process per GPU):
the ceiling of every end-to-end number in bench.py (the e2e legs copy their inputs in and their results out inside the timed
region).  Run:  python -m torch.distributed.run --nproc-per-node 8 tools/pcie_probe.py   -> one JSON line on rank 0."""
import json
import os

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
MB = 256
host_a, host_b = torch.empty(MB << 20, dtype=torch.uint8).pin_memory(), torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
dev_a, dev_b = torch.empty(MB << 20, dtype=torch.uint8, device="cuda"), torch.empty(MB << 20, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(kind, reps=8):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream())
    s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        if kind in ("d2h", "both"):
            with torch.cuda.stream(s1):
                host_a.copy_(dev_a, non_blocking=True)
        if kind in ("h2d", "both"):
            with torch.cuda.stream(s2):
                dev_b.copy_(host_b, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    return reps * MB * (1 << 20) / (e0.elapsed_time(e1) * 1e-3) / 1e9  # GB/s per direction


out = {}
for active in (1, 2, 4, 8):
    if active > world:
        break
    for kind in ("d2h", "h2d", "both"):
        if rank < active:
            run(kind, 2)
            v = run(kind)
        else:
            if world > 1:
                dist.barrier()
                dist.barrier()
            v = 0.0
        t = torch.tensor([v], device="cuda")
        if world > 1:
            g = [torch.zeros(1, device="cuda") for _ in range(world)]
            dist.all_gather(g, t)
            vals = [float(x) for x in g][:active]
        else:
            vals = [v]
        out[f"{active}gpu_{kind}"] = {"per_gpu_GBps": [round(x, 1) for x in vals], "aggregate_GBps": round(sum(vals), 1)}
if rank == 0:
    print(json.dumps({"probe": "pinned host<->device copies, 256 MB each, per direction", "world": world, "results": out}))
if world > 1:
    dist.destroy_process_group()
