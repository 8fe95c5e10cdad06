#!/usr/bin/env python
"""Runs mhl, pm, fdrp, qfdrp once each (after one warm-up pass) on the bench workload, for an ncu capture of the gather kernels."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_chr19 as bench
from metheor_b200 import engine
cov = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
length = int(sys.argv[2]) if len(sys.argv) > 2 else bench.CONTIG_LEN
b, _ = bench.make_workload(0, cov, length)
view = {np.dtype("uint32"): np.int32, np.dtype("uint16"): np.int16, np.dtype("uint64"): np.int64}
devb = dict(b)
for k in ("start", "end", "meta", "cpg_off", "cpg_pos", "cpg_rel", "meth"):
    devb[k] = torch.from_numpy(b[k].view(view.get(b[k].dtype, b[k].dtype))).cuda()
for m in [(x,) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else ("mhl", "pm", "fdrp", "qfdrp"))]:
    ctx = engine.Context(engine.default_params(m, flags=engine.FLAG_KEEP_ON_DEVICE), [length])
    for _ in range(2):
        ctx.reset(); ctx.submit(devb); ctx.finish()
    ctx.close()
print("done")
