#!/usr/bin/env python
"""A/B of library builds on the bench workload (kernel-tuning experiments, see profiles/build_variant.sh).
   python profiles/ingest_ab.py LIB [LIB ...]    LIB = path of a libmetheor_b200 build ('default' = the shipped one)
   python profiles/ingest_ab.py --park           only generate the workload; `--child LIB` then runs one build
The parent generates the synthetic 30x chr19-sized contig once and parks it in /dev/shm; one child per library loads it,
runs pdr+lpmd / pdr / lpmd / pm+me with the inputs resident in HBM and prints one JSON line: per-kernel ms (the
engine's own CUDA-event timing, MTH_FLAG_PROFILE), ms per pass, and a digest of every result so that builds can be
compared for equality as well as for speed."""
import hashlib, json, os, subprocess, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PARK = "/dev/shm/mth_ab_workload.npz"
KEYS = ("start", "end", "meta", "cpg_off", "cpg_pos", "cpg_rel", "meth")


def child(lib):
    import torch
    import bench_chr19 as bench
    from metheor_b200 import engine
    z = np.load(PARK)
    b = {k: z[k] for k in KEYS}
    b.update(tid=0, n_reads=int(z["n_reads"]), n_cpg=int(z["n_cpg"]))
    dev = torch.device("cuda", 0)
    view = {np.dtype("uint32"): np.int32, np.dtype("uint16"): np.int16, np.dtype("uint64"): np.int64}
    devb = dict(b)
    for k in KEYS:
        devb[k] = torch.from_numpy(b[k].view(view.get(b[k].dtype, b[k].dtype))).to(dev)
    out = {"lib": lib, "runs": []}
    for measures in (("pdr", "lpmd"), ("pdr",), ("lpmd",), ("pm", "me")):
        # digest of the rows (copied back once)
        ctx = engine.Context(engine.default_params(measures), [bench.CONTIG_LEN])
        ctx.submit(devb)
        res = ctx.finish()
        h = hashlib.sha256()
        for m in sorted(res):
            for k in sorted(res[m]):
                v = res[m][k]
                h.update(f"{m}.{k}".encode())
                h.update(np.ascontiguousarray(v).tobytes() if isinstance(v, np.ndarray) else repr(v).encode())
        ctx.close()
        ctx = engine.Context(engine.default_params(measures, flags=engine.FLAG_KEEP_ON_DEVICE | engine.FLAG_PROFILE), [bench.CONTIG_LEN])
        acc, n = {}, 5
        for it in range(3 + n):
            ctx.reset(); ctx.submit(devb); ctx.finish()
            if it >= 3:
                for k, v in ctx.stats()["kernels"].items():
                    acc[k] = acc.get(k, 0.0) + v["ms"] / n
        ctx.close()
        ctx = engine.Context(engine.default_params(measures, flags=engine.FLAG_KEEP_ON_DEVICE), [bench.CONTIG_LEN])
        for _ in range(3):
            ctx.reset(); ctx.submit(devb); ctx.finish()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10):
            ctx.reset(); ctx.submit(devb); ctx.finish()
        e1.record(); torch.cuda.synchronize()
        ctx.close()
        out["runs"].append({"measures": measures, "ms_per_pass": round(e0.elapsed_time(e1) / 10, 4), "digest": h.hexdigest()[:16],
                            "kernels_ms": {k: round(v, 4) for k, v in acc.items()}})
    print(json.dumps(out), flush=True)


def main():
    if sys.argv[1] == "--child":
        return child(sys.argv[2])
    import bench_chr19 as bench
    made = not os.path.exists(PARK)
    if made:
        b, _ = bench.make_workload(0, bench.COVERAGE, bench.CONTIG_LEN)
        np.savez(PARK, n_reads=b["n_reads"], n_cpg=b["n_cpg"], **{k: b[k] for k in KEYS})
    if sys.argv[1] == "--park":  # leave the workload in /dev/shm for children started by hand (ncu)
        return
    for lib in sys.argv[1:]:
        env = dict(os.environ)
        if lib != "default":
            env["METHEOR_B200_LIB"] = os.path.abspath(lib)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", lib], env=env, check=False)
    if made:
        os.remove(PARK)


if __name__ == "__main__":
    main()
