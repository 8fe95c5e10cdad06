#!/usr/bin/env python
"""Where the end-to-end step goes: pure H2D of the compact arrays (the floor), engine step with rows left on the device, full step."""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_chr19 as bench
from metheor_b200 import engine, batch as B
b, _ = bench.make_workload(0)
hostc = B.to_compact(b)
pinned = {}
for k, v in list(hostc.items()):
    if isinstance(v, np.ndarray) and v.size:
        pinned[k] = torch.from_numpy(v.view({np.dtype("uint16"): np.int16}.get(v.dtype, v.dtype))).pin_memory()
        hostc[k] = pinned[k]
dev = {k: torch.empty_like(v, device="cuda") for k, v in pinned.items()}
nbytes = sum(v.numel() * v.element_size() for v in pinned.values())
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
def h2d():
    for k in pinned: dev[k].copy_(pinned[k], non_blocking=True)
ms = t(h2d)
out = {"compact_bytes": nbytes, "pure_h2d_ms": ms, "pure_h2d_GBps": nbytes / ms / 1e6}
for flags, name in ((engine.FLAG_KEEP_ON_DEVICE, "step_rows_on_device_ms"), (0, "step_full_ms")):
    ctx = engine.Context(engine.default_params(("pdr", "lpmd"), flags=flags), [bench.CONTIG_LEN])
    def step():
        ctx.reset(); ctx.submit_compact(hostc); ctx.finish()
    out[name] = t(step)
    ctx.close()
ctx = engine.Context(engine.default_params(("pdr", "lpmd"), flags=engine.FLAG_KEEP_ON_DEVICE | engine.FLAG_PROFILE), [bench.CONTIG_LEN])
for _ in range(3):
    ctx.reset(); ctx.submit_compact(hostc); ctx.finish()
out["kernels_ms"] = {k: round(v["ms"], 4) for k, v in ctx.stats()["kernels"].items()}
print(json.dumps(out))
