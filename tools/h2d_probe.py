"""PCIe probe: pinned H2D / D2H bandwidth on this box (context for bench.py's e2e number)."""
import json
import torch

out = {}
for mb in (64, 1250):
    n = mb * 1000 * 1000
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, (src, dst) in {"h2d": (h, d), "d2h": (d, h)}.items():
        best = 0.0
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dst.copy_(src, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            best = max(best, n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        out[f"{name}_{mb}MB_GBps"] = round(best, 2)
print(json.dumps(out))
