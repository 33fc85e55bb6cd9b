"""Measure the pinned D2H / H2D copy ceiling of this box (the binding roof of the read stage)."""
import json, torch
res = {}
for mb in (64, 256, 1024):
    n = mb << 20
    d = torch.empty(n, dtype=torch.uint8, device="cuda"); h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    for name, (dst, src) in {"d2h": (h, d), "h2d": (d, h)}.items():
        for _ in range(2): dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): dst.copy_(src, non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        res[f"{name}_{mb}MiB_GBps"] = 5 * n / (e0.elapsed_time(e1) / 1e3) / 1e9
print(json.dumps(res))
