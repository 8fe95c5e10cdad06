import json, os, sys
import numpy as np, torch
sys.path.insert(0, '.')
import bench_chr19 as bench
from metheor_b200 import engine
b, _ = bench.make_workload(0, 30.0, 20_000_000)
view = {np.dtype("uint32"): np.int32, np.dtype("uint16"): np.int16, np.dtype("uint64"): np.int64}
devb = dict(b)
for k in ("start", "end", "meta", "cpg_off", "cpg_pos", "cpg_rel", "meth"):
    devb[k] = torch.from_numpy(b[k].view(view.get(b[k].dtype, b[k].dtype))).cuda()
for name, ov in (("default", {}), ("no_evaluate(min_depth=1e6)", dict(min_depth=1000000)), ("min_overlap=1000 (pairs skipped early)", dict(min_overlap=1000)), ("max_depth=8", dict(max_depth=8, min_depth=1))):
    for m in ("fdrp", "qfdrp", "mhl"):
        if m == "mhl" and name != "default" and "no_eval" not in name: continue
        prm = {m: ov} if m != "mhl" else {m: ({k: v for k, v in ov.items() if k == "min_depth"})}
        ctx = engine.Context(engine.default_params((m,), flags=engine.FLAG_KEEP_ON_DEVICE | engine.FLAG_PROFILE, **prm), [20_000_000])
        for _ in range(3):
            ctx.reset(); ctx.submit(devb); r = ctx.finish()
        k = ctx.stats()["kernels"]
        print(name, m, {x: round(v["ms"], 3) for x, v in k.items() if x in ("k_fdrp", "k_qfdrp", "k_mhl")}, r[m]["n"])
        ctx.close()
