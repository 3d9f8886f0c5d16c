#!/usr/bin/env python
"""Is host<->device copy bandwidth on this box stable?  (explains run-to-run jitter of the e2e leg)"""
import time, torch
n = 700 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
for rep in range(3):
    ts = []
    for k in range(12):
        t0 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print("H2D 700 MiB ms:", " ".join(f"{x:.1f}" for x in ts), flush=True)
    ts = []
    for k in range(12):
        t0 = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print("D2H 700 MiB ms:", " ".join(f"{x:.1f}" for x in ts), flush=True)
    # many small round trips (launch + sync latency)
    x = torch.zeros(1, device="cuda")
    t0 = time.perf_counter()
    for k in range(2000):
        x.add_(1); torch.cuda.synchronize()
    print(f"launch+sync round trip: {(time.perf_counter() - t0) / 2000 * 1e6:.1f} us", flush=True)
