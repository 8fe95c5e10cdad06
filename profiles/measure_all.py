#!/usr/bin/env python
"""Per-measure kernel times on the bench workload (synthetic 30x chr19-sized contig), inputs resident in HBM.
   python profiles/measure_all.py [coverage] -> JSON lines on stdout"""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_chr19 as bench
from metheor_b200 import engine

cov = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
b, sites = bench.make_workload(0, cov, bench.CONTIG_LEN)
dev = torch.device("cuda", 0)
view = {np.dtype("uint32"): np.int32, np.dtype("uint16"): np.int16, np.dtype("uint64"): np.int64}
devb = dict(b)
for k in ("start", "end", "meta", "cpg_off", "cpg_pos", "cpg_rel", "meth"):
    devb[k] = torch.from_numpy(b[k].view(view.get(b[k].dtype, b[k].dtype))).to(dev)
R, I = b["n_reads"], b["n_cpg"]
for measures in (("pdr",), ("lpmd",), ("mhl",), ("pm",), ("me",), ("pm", "me"), ("fdrp",), ("qfdrp",), ("pdr", "lpmd", "mhl", "pm", "me", "fdrp", "qfdrp")):
    ctx = engine.Context(engine.default_params(measures, flags=engine.FLAG_KEEP_ON_DEVICE | engine.FLAG_PROFILE), [bench.CONTIG_LEN])
    acc, n = {}, 3
    for it in range(2 + n):
        ctx.reset(); ctx.submit(devb); res = ctx.finish()
        if it >= 2:
            for k, v in ctx.stats()["kernels"].items():
                acc[k] = acc.get(k, 0.0) + v["ms"] / n
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        ctx.reset(); ctx.submit(devb); res = ctx.finish()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    rows = {m: (res[m]["n"] if "n" in res[m] else None) for m in res}
    print(json.dumps({"measures": measures, "coverage": cov, "reads": R, "calls": I, "sites": ctx.stats()["n_sites"], "ms_per_pass": round(ms, 3),
                      "reads_per_sec": R / ms * 1e3, "rows": rows, "kernels_ms": {k: round(v, 4) for k, v in acc.items()}}), flush=True)
    ctx.close()
