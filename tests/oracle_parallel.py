"""The CPU oracle on every host core.  TEST INFRASTRUCTURE ONLY (tests/ and bench.py's checker / cpu_baseline legs).

The reference is single-threaded; the only way to run it on N cores is N independent processes.  For ONE contig that is
still exact when every process gets a position bin plus a halo of reads and keeps only the rows it owns — the same
argument as the engine's multi-GPU sharding (metheor_b200/shard.py): every contributor and every flush trigger of a site
p starts in [p - Lmax + 1, p + 1].  Used to check the engine's rows on a whole contig of the bench workload in seconds
(FDRP / qFDRP alone need minutes on one core)."""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

_NB = None   # the contig's reads (numpy batch), inherited by the workers through fork
_JOB = None


def _worker(k):
    from metheor_b200 import batch as B
    from oracle_lib import Oracle
    import time
    measure, prm, cuts, halo, seed = _JOB
    lo, hi = cuts[k], cuts[k + 1]
    start = _NB["start"]
    if measure == "lpmd":  # per-read measure: disjoint slices, counters add up
        a, e = int(np.searchsorted(start, lo, "left")), int(np.searchsorted(start, hi, "left"))
    else:
        a, e = int(np.searchsorted(start, lo - halo, "left")), int(np.searchsorted(start, hi, "right"))
    if e <= a:
        return None, 0.0
    sub = B.slice_range(_NB, a, e)
    o = Oracle.from_soa(**B.to_oracle_soa([sub]))
    t0 = time.perf_counter()
    if measure == "pdr":
        r = o.pdr(**prm)
        key = "pos"
    elif measure == "mhl":
        r = o.mhl(**prm)
        key = "pos"
    elif measure in ("fdrp", "qfdrp"):
        r = o.fdrp(seed=seed, quantitative=(measure == "qfdrp"), **prm)
        key = "pos"
    elif measure in ("pm", "me", "quartets"):
        r = o.quartets(**prm)
        key = "p1"
    elif measure == "lpmd":
        r = o.lpmd(**prm)
        dt = time.perf_counter() - t0
        o.close()
        return r, dt
    else:
        raise ValueError(measure)
    dt = time.perf_counter() - t0
    keep = (r[key] >= lo) & (r[key] < hi)
    out = {k2: v[keep] for k2, v in r.items()}
    o.close()
    return out, dt


def run(nb, measure, prm, n_proc=None, interval=None, seed=0):
    """Rows of `measure` for the reads of ONE contig `nb` (numpy batch), restricted to sites (quartets: first site) in
    `interval` = (lo, hi) (default: everything), computed by n_proc oracle processes.
    -> (rows dict | lpmd dict, {"wall_s", "cpu_s", "procs"})"""
    global _NB, _JOB
    import time
    n_proc = n_proc or os.cpu_count() or 1
    start = np.asarray(nb["start"], np.int64)
    span = int((np.asarray(nb["end"], np.int64) - start).max(initial=0)) + 1
    halo = span + 2
    lo, hi = interval if interval is not None else (int(start[0]) - 2 if len(start) else 0, int(start[-1]) + span + 2 if len(start) else 1)
    a, e = int(np.searchsorted(start, lo, "left")), int(np.searchsorted(start, hi, "left"))
    n_bins = max(1, min(n_proc * 4, (e - a) // 20000 or 1))  # several bins per process: islands make bins uneven
    cuts = [lo] + [int(start[a + (e - a) * k // n_bins]) for k in range(1, n_bins)] + [hi]
    cuts = sorted(set(cuts))
    _NB, _JOB = nb, (measure, dict(prm), cuts, halo, seed)
    t0 = time.perf_counter()
    if n_proc == 1:
        parts = [_worker(k) for k in range(len(cuts) - 1)]
    else:
        with mp.get_context("fork").Pool(n_proc) as pool:
            parts = pool.map(_worker, range(len(cuts) - 1), chunksize=1)
    wall = time.perf_counter() - t0
    _NB = _JOB = None
    cpu = sum(p[1] for p in parts)
    parts = [p[0] for p in parts if p[0] is not None]
    info = {"wall_s": wall, "cpu_s": cpu, "procs": n_proc, "bins": len(cuts) - 1}
    if measure == "lpmd":
        tot = {k: sum(int(p[k]) for p in parts) for k in ("n_read", "n_valid_read", "n_conc", "n_disc")}
        tot["lpmd"] = np.float32(tot["n_disc"]) / np.float32(tot["n_conc"] + tot["n_disc"]) if (tot["n_conc"] + tot["n_disc"]) else np.float32("nan")
        return tot, info
    if not parts:
        return None, info
    return {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}, info
