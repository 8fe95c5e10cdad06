#!/usr/bin/env python
"""Whole-genome-scale resident run: the bench workload (synthetic 30x chr19-sized contig) submitted as K contigs of one
reference (K = 24: 1.4 Gb, 281 M reads, 748 M CpG calls, 25 M sites), all held in HBM as ONE region.
   python profiles/scale_wg.py [K] -> JSON lines"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_chr19 as bench
from metheor_b200 import engine

K = int(sys.argv[1]) if len(sys.argv) > 1 else 24
b, _ = bench.make_workload(0)
view = {np.dtype("uint32"): np.int32, np.dtype("uint16"): np.int16, np.dtype("uint64"): np.int64}
devb = dict(b)
for k in ("start", "end", "meta", "cpg_off", "cpg_pos", "cpg_rel", "meth"):
    devb[k] = torch.from_numpy(b[k].view(view.get(b[k].dtype, b[k].dtype))).cuda()
R, I = b["n_reads"] * K, b["n_cpg"] * K
for measures in (("pdr", "lpmd"), ("pm", "me"), ("mhl",), ("fdrp",), ("pdr", "lpmd", "mhl", "pm", "me", "fdrp", "qfdrp")):
    ctx = engine.Context(engine.default_params(measures, flags=engine.FLAG_KEEP_ON_DEVICE), [bench.CONTIG_LEN] * K)
    best = None
    for it in range(3):
        ctx.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for tid in range(K):
            ctx.submit(dict(devb, tid=tid))
        res = ctx.finish()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    st = ctx.stats()
    free, total = torch.cuda.mem_get_info()
    print(json.dumps({"measures": measures, "contigs": K, "reads": R, "calls": I, "sites": st["n_sites"], "regions": st["n_regions"],
                      "seconds": round(best, 4), "reads_per_sec": R / best, "rows": {m: res[m].get("n") for m in res},
                      "hbm_used_GB": round((total - free) / 1e9, 1)}), flush=True)
    ctx.close()
