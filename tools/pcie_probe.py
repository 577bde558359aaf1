#!/usr/bin/env python
"""Host<->device copy bandwidth with N ranks copying at the same time (run under torchrun):
what the end-to-end path can reach at most on this box.  Every rank moves the bench's per-step
volumes (106 MB up, 315 MB down) between pinned host memory and its GPU, alone and together."""
import json
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl")
dev = torch.device("cuda", local)
up_h = torch.empty(106 << 20, dtype=torch.uint8).pin_memory()
dn_h = torch.empty(315 << 20, dtype=torch.uint8).pin_memory()
up_d = torch.empty_like(up_h, device=dev)
dn_d = torch.empty_like(dn_h, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


def up():
    with torch.cuda.stream(s1):
        up_d.copy_(up_h, non_blocking=True)


def down():
    with torch.cuda.stream(s2):
        dn_h.copy_(dn_d, non_blocking=True)


def both():
    up(); down()


res = torch.tensor([up_h.numel() / timed(up) / 1e9, dn_h.numel() / timed(down) / 1e9,
                    (up_h.numel() + dn_h.numel()) / timed(both) / 1e9, timed(both) * 1e3], device=dev, dtype=torch.float64)
if world > 1:
    allr = [torch.empty_like(res) for _ in range(world)]
    dist.all_gather(allr, res)
else:
    allr = [res]
if rank == 0:
    print(json.dumps({"n": world, "cpus": os.cpu_count(),
                      "per_rank_gbs [h2d, d2h, both, both_ms]": [[round(float(v), 2) for v in r.tolist()] for r in allr]}))
if world > 1:
    dist.destroy_process_group()
