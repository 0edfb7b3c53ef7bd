"""PCIe probe: pinned H2D / D2H bandwidth on this box (context for bench.py's e2e number).
    python tools/h2d_probe.py                                   # one GPU
    torchrun --nproc-per-node N tools/h2d_probe.py [--bind]     # N GPUs copying at the same time; --bind pins each rank
                                                                # to its GPU's local CPUs first (lash_bind_thread_to_device)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
bound = 0
if "--bind" in sys.argv:
    from lash_b200.capi import lib
    bound = lib().lash_bind_thread_to_device(local)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out = {"rank": rank, "world": world, "bound_cpus": bound, "cpus_allowed": len(os.sched_getaffinity(0))}
for mb in (64, 1250):
    n = mb * 1000 * 1000
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, (src, dst) in {"h2d": (h, d), "d2h": (d, h)}.items():
        best = 0.0
        for _ in range(5):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dst.copy_(src, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            best = max(best, n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        out[f"{name}_{mb}MB_GBps"] = round(best, 2)
print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
