#!/usr/bin/env python
"""Run measure sets over the whole-genome bench workload (or some of its contigs) with device-resident inputs — the command that
ncu wraps (profiles/r2_ncu.sh).  Writes a sidecar JSON with the number of engine kernels each set launched, in order, so that
an ncu launch list of the same process can be cut into per-set segments (profiles/r2_traffic.py).
   python profiles/wg_pass.py --sets pm+me,all7 [--contigs 0,20] [--scale 1.0] [--coverage 30] [--passes 1] [--sidecar out.json]
A set named chr19:pdr+lpmd runs on the chr19-sized contig of BASELINE.json configs[1] instead."""
import argparse, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from metheor_b200 import engine, synth_gpu as G

ap = argparse.ArgumentParser()
ap.add_argument("--sets", default="all7")
ap.add_argument("--contigs", default="")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--coverage", type=float, default=30.0)
ap.add_argument("--passes", type=int, default=1)
ap.add_argument("--warm", type=int, default=0)
ap.add_argument("--sidecar", default="")
ap.add_argument("--profile", action="store_true", help="per-kernel CUDA-event times (MTH_FLAG_PROFILE) of the last pass")
a = ap.parse_args()
dev = torch.device("cuda", 0)
contigs = G.genome(a.scale)
ref_len = [l for _, l in contigs]
tids = [int(x) for x in a.contigs.split(",")] if a.contigs else list(range(len(contigs)))
wg = [G.make_contig(dev, bench.SEED, t, ref_len[t], a.coverage) for t in tids]
torch.cuda.synchronize()
R, I = sum(b["n_reads"] for b in wg), sum(b["n_cpg"] for b in wg)
side = {"reads": R, "calls": I, "contigs": tids, "scale": a.scale, "coverage": a.coverage, "sets": []}
c19 = None
for name in a.sets.split(","):
    where, _, ms = name.rpartition(":")
    measures = bench.ALL7 if ms == "all7" else tuple(ms.split("+"))
    if where == "chr19":
        if c19 is None:
            import bench_chr19 as X
            b, _ = X.make_workload(0, 30.0, X.CONTIG_LEN)
            view = {np.dtype("uint32"): np.int32, np.dtype("uint16"): np.int16, np.dtype("uint64"): np.int64}
            c19 = dict(b)
            for k in ("start", "end", "meta", "cpg_off", "cpg_pos", "cpg_rel", "meth"):
                c19[k] = torch.from_numpy(b[k].view(view.get(b[k].dtype, b[k].dtype))).to(dev)
        batches, rl = [c19], [X.CONTIG_LEN]
    else:
        batches, rl = wg, ref_len
    ctx = engine.Context(engine.default_params(measures, flags=engine.FLAG_KEEP_ON_DEVICE | (engine.FLAG_PROFILE if a.profile else 0)), rl, device=0)
    launches = 0
    t0 = time.perf_counter()
    for it in range(a.warm + a.passes):
        ctx.reset()
        for b in batches:
            ctx.submit(b)
        res = ctx.finish()
        launches += ctx.stats()["kernel_launches"]
    st = ctx.stats()
    side["sets"].append({"name": name, "measures": list(measures), "passes": a.warm + a.passes, "engine_kernel_launches": int(launches),
                         "reads": int(st["n_reads"]), "calls": int(st["n_cpg"]), "sites": int(st["n_sites"]),
                         "rows": {m: int(res[m]["n"]) for m in res if "n" in res[m]}, "pair_ops": int(st["fdrp_pair_ops"]),
                         "seconds": time.perf_counter() - t0, "fallback_sites": [int(st["fallback_sites_mhl"]), int(st["fallback_sites_fdrp"])],
                         "kernels_ms": {k: round(v["ms"], 3) for k, v in st["kernels"].items()} if a.profile else None})
    ctx.close()
print(json.dumps(side))
if a.sidecar:
    json.dump(side, open(a.sidecar, "w"), indent=1)
