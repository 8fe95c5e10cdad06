#!/usr/bin/env python
"""bench.py — metheor_b200 on BASELINE.json's whole-genome configuration, per measure.

Workload (`config.workload`): BASELINE.json configs[2] — synthetic 30x WGBS over 24 contigs with the hg38 chromosome
lengths (3.09 Gb, ~28 M CpG sites, ~0.62 G 150-bp single-end reads on both strands, ~0.82 G CpG calls), generated on the
GPU from an integer hash (metheor_b200/synth_gpu.py; bit-identical on CPU and CUDA).  The headline measure set is that
config's `pm` + `me`; the line also carries one block per measure (pdr, lpmd, mhl, pm, me, fdrp, qfdrp) on the SAME
workload, the combined passes, and — as secondary legs — configs[1] (pdr + lpmd on a chr19-sized contig: the round-1
headline), configs[3] (fdrp + qfdrp at 60x), the BAM -> TSV path and `tag`.

One "step" = one full pass of the hot path over the whole read set: per contig mth_submit (ingest: validation, site
bitmap, LPMD) -> site dictionary -> measure kernels -> row emission, then mth_finish.
  value     : reads/s, SoA batches already resident in HBM (device pointers handed to mth_submit, rows left in HBM)
  e2e       : reads/s through the same C-ABI calls with HOST buffers: pinned SoA arrays in the layout north_star names
              (per-read int32 CpG positions + packed uint64 methylation words), H2D of every batch and D2H of the rows inside
              the timed region; nothing is pre-encoded on the host
  roofline  : the measure's dominant kernel: algorithmic bytes (SURVEY 8d / DESIGN.md §4) over its CUDA-event time, against
              MEASURED_PEAKS.json; `roofline_measure` = SURVEY 8d bytes of the measure over the sum of its kernels
  cpu_baseline : the CPU oracle (C++ restatement of metheor 0.1.9 — the Rust binary cannot be built here), one thread like
              the reference, on a stated bounded sample of the same workload
  parity_full_size : the rows the engine produced in the FULL pass, restricted to one whole contig (chr21; a prefix of it
              for fdrp / qfdrp), bit-identical to the oracle run on that contig's reads (contigs are independent in the
              reference: a tid change flushes every accumulator, readutil.rs:290-295), plus digests that tie every
              single-measure pass to the all-seven pass that was checked

N > 1 (torchrun): STRONG scaling — the same genome is cut into N position bins (+ halo reads, MTH_META_HALO), one bin per
rank, no data-path collective; the one exchange is the library's NCCL all-reduce of LPMD's counters (mth_allreduce), once
per pass.  Rank 0 then runs the whole genome alone (untimed) and checks that the owned rows of all ranks together have
the same digest.
`--impl reference`: the CPU oracle alone on the parity contig (rank 0 only; no GPU needed).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SEED = 20260102
COVERAGE = 30.0
HEADLINE = ("pm", "me")
ALL7 = ("pdr", "lpmd", "mhl", "pm", "me", "fdrp", "qfdrp")
SINGLE = tuple((m,) for m in ALL7)
COMBOS = (("pm", "me"), ("pdr", "lpmd"), ("fdrp", "qfdrp"), ALL7)
PARITY_TID = 20                      # chr21: the smallest contig (46.7 Mb, ~9.3 M reads at 30x)
FDRP_PARITY_SPAN = 8_000_000         # fdrp / qfdrp rows are checked on the first 8 Mb of the parity contig (oracle: O(d^2 * 403) per site)
HALO = 1024                          # reads starting up to this far before a bin also go to its rank (>= longest reference span + 2)
DTYPE = "u32/u64 integer + f32 finalisation"


def workload_name(cov, scale):
    s = (f"synthetic {cov:g}x WGBS whole genome: 24 contigs with the hg38 chromosome lengths (3.09 Gb), ~28 M CpG sites, "
         f"150-bp SE reads on both strands (BASELINE.json configs[2]), seed {SEED}")
    return s if scale == 1.0 else s + f", contig lengths x {scale:g}"


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


_REAL_STDOUT = None


def ensure_built():
    need = [os.path.join(ROOT, "metheor_b200", "csrc", "libmetheor_b200.so"), os.path.join(ROOT, "metheor_b200", "host", "libmetheor_host.so"),
            os.path.join(ROOT, "oracle", "_build", "liboracle.so")]
    if not all(os.path.exists(p) for p in need):
        if int(os.environ.get("LOCAL_RANK", "0")) == 0:
            import __graft_entry__
            __graft_entry__.build()
        else:
            t0 = time.time()
            while not all(os.path.exists(p) for p in need) and time.time() - t0 < 600:
                time.sleep(1.0)
            time.sleep(2.0)


# ---------------------------------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------------------------------
# A region (one contig part on one GPU) costs the light passes ~0.12 ms next to their cost per base (~25 launches of small kernels,
# measured at N = 8: DESIGN.md 7) — as much as ~30 Mb of genome in the pm + me pass at 30x.  The position bins of the resident
# workload are balanced on bases + REGION_COST_BP x contigs (shard.plan_bins region_cost), scaled with 30 / coverage; the heavier
# 60x / 100x legs (a region boundary is worth ~2 Mb there) are balanced on length alone.
REGION_COST_BP = 30_000_000


def plan_bins_by_length(ref_len, world, region_cost=0):
    from metheor_b200 import shard
    return shard.plan_bins(ref_len, world, region_cost=region_cost)


def gen_shard(torch, dev, contigs, coverage, intervals, world):
    """The reads a rank needs for its intervals [(tid, lo, hi)]: start in [lo - HALO, hi]; halo copies carry MTH_META_HALO.
    -> (list of torch batches, owned reads, owned calls)"""
    from metheor_b200 import synth_gpu as G
    out, owned_r, owned_i = [], 0, 0
    for tid, lo, hi in intervals:
        length = contigs[tid][1]
        whole = world == 1 or (lo == 0 and hi >= length)
        b = G.make_contig(dev, SEED, tid, length, coverage, start_range=None if whole else (lo - HALO, hi + 1))
        if not whole:
            st = b["start"]
            halo = (st < lo) | (st >= hi)
            b["meta"] = torch.where(halo, b["meta"] | (1 << 9), b["meta"])
            cnt = (b["cpg_off"][1:] - b["cpg_off"][:-1]).to(torch.int64)
            owned_r += int((~halo).sum()); owned_i += int(cnt[~halo].sum())
        else:
            owned_r += b["n_reads"]; owned_i += b["n_cpg"]
        out.append(b)
    return out, owned_r, owned_i


def batch_bytes(b, with_rel):
    return 24 * b["n_reads"] + 4 + (6 if with_rel else 4) * b["n_cpg"]


# ---------------------------------------------------------------------------------------------------------------------
# device-side digests of result rows (order-independent 64-bit sums: shards add up to the whole)
# ---------------------------------------------------------------------------------------------------------------------
class _DevArr:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _dt(torch, dev, ptr, n, typestr):
    if not ptr or n == 0:
        return torch.zeros(0, dtype=torch.int64, device=dev)
    return torch.as_tensor(_DevArr(ptr, n, typestr), device=dev).to(torch.int64)


_K = [-7046029254386353131, -4658895280553007687, -7723592293110705685, 0x2545F4914F6CDD1D, 0x27D4EB2F165667C5, 0x165667B19E3779F9,
      -3750763034362895579]


def _mix(torch, cols):
    x = torch.zeros_like(cols[0])
    for k, c in zip(_K, cols):
        x = (x ^ c) * k
        x = x ^ ((x >> 29) & ((1 << 35) - 1))
    return x


def rows_digest(torch, dev, ctx, measures, owned=None):
    """{measure: [n_rows, digest]} from the device-resident rows of the last finish(); owned = [(tid, lo, hi)] keeps only the
    rows whose (first) position lies in one of the intervals (position-bin sharding)."""
    r = ctx.results_device()
    out = {}

    def own_mask(tid, pos):
        if owned is None:
            return None
        m = torch.zeros_like(tid, dtype=torch.bool)
        for t, lo, hi in owned:
            m |= (tid == t) & (pos >= (lo if lo > 0 else -1)) & (pos < hi)
        return m

    for name in ("pdr", "mhl", "fdrp", "qfdrp"):
        if name not in measures:
            continue
        s = getattr(r, name)
        n = int(s.n)
        cols = [_dt(torch, dev, s.tid, n, "<i4"), _dt(torch, dev, s.pos, n, "<i4"), _dt(torch, dev, s.value, n, "<i4")]
        if name == "pdr":
            cols += [_dt(torch, dev, s.n_conc, n, "<i4"), _dt(torch, dev, s.n_disc, n, "<i4")]
        x = _mix(torch, cols)
        m = own_mask(cols[0], cols[1])
        if m is not None:
            x = x[m]
        out[name] = [int(x.numel()), int(x.sum())]
    for name in ("pm", "me"):
        if name not in measures:
            continue
        s = getattr(r, name)
        n = int(s.n)
        cols = [_dt(torch, dev, getattr(s, k), n, "<i4") for k in ("tid", "p1", "p2", "p3", "p4", "value")]
        x = _mix(torch, cols)
        m = own_mask(cols[0], cols[1])
        if m is not None:
            x = x[m]
        out[name] = [int(x.numel()), int(x.sum())]
    return out


# ---------------------------------------------------------------------------------------------------------------------
# algorithmic bytes and kernel attribution
# ---------------------------------------------------------------------------------------------------------------------
def survey_bytes(m, R, I, C, Q):
    return {"pdr": 16 * R + 4 * I + 12 * C, "lpmd": 16 * R + 6 * I + 16, "mhl": 24 * R + 4 * I + 8 * C, "pm": 16 * R + 4 * I + 20 * Q,
            "me": 16 * R + 4 * I + 20 * Q, "fdrp": 24 * R + 4 * I + 8 * C, "qfdrp": 24 * R + 4 * I + 8 * C}[m]


def kernel_bytes(k, measures, R, I, C, Q):
    """Compulsory bytes of ONE kernel family over a pass (what it must read and write at least once), never more than the
    SURVEY 8d figure of the measure it serves."""
    lp = "lpmd" in measures
    if k == "k_ingest":
        return 16 * R + (6 if lp else 4) * I + (32 if lp else 0)
    if k in ("k_pdr_scatter", "k_pdr_tile"):
        return 5 * I + 8 * C                      # cpg_pos + flag byte per call, two u32 counters per site
    if k == "k_pdr_gather":
        return 16 * R + 4 * I + 8 * C
    if k == "k_mhl":
        return 24 * R + 4 * I + 8 * C
    if k in ("k_fdrp", "k_qfdrp"):
        return 24 * R + 4 * I + 8 * C
    if k == "k_fdrp_qfdrp":
        return 24 * R + 4 * I + 16 * C
    if k in ("k_pm_scatter", "k_me_scatter"):
        return 5 * I + 64 * C                     # cpg_pos + flag byte per call, one 16-bin u32 histogram per site
    if k in ("k_pm_count", "k_me_count", "k_pm_emit", "k_me_emit", "k_pm_me_emit", "k_pm_hist", "k_me_hist"):
        return 64 * C + 24 * Q
    return None


MEASURE_KERNELS = {"pdr": ("k_pdr",), "lpmd": (), "mhl": ("k_mhl",), "pm": ("k_pm",), "me": ("k_me",), "fdrp": ("k_fdrp",), "qfdrp": ("k_qfdrp",)}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


# ---------------------------------------------------------------------------------------------------------------------
# CPU oracle legs
# ---------------------------------------------------------------------------------------------------------------------
ORACLE_PRM = dict(pdr=dict(min_depth=10, min_cpgs=4, min_qual=10), mhl=dict(min_depth=10, min_cpgs=4, min_qual=10),
                  pm=dict(min_depth=10, min_qual=10), me=dict(min_depth=10, min_qual=10),
                  fdrp=dict(min_qual=10, min_depth=10, max_depth=40, min_overlap=35),
                  qfdrp=dict(min_qual=10, min_depth=10, max_depth=40, min_overlap=35), lpmd=dict(min_distance=2, max_distance=16, min_qual=10))
CPU_SAMPLE = dict(pdr=None, lpmd=None, mhl=None, pm=None, me=None, fdrp=250_000, qfdrp=150_000)  # reads; None = the whole parity contig


def oracle_single(nb, m, n_reads=None, repeat=1):
    """One thread, like the reference: -> (reads/s, seconds, reads) for measure m on the first n_reads reads of nb."""
    from metheor_b200 import batch as B
    from oracle_lib import Oracle
    n = nb["n_reads"] if n_reads is None else min(n_reads, nb["n_reads"])
    sub = nb if n == nb["n_reads"] else B.slice_range(nb, 0, n)
    o = Oracle.from_soa(**B.to_oracle_soa([sub]))
    best = None
    for _ in range(repeat):
        t0 = time.perf_counter()
        if m == "pdr": o.pdr(**ORACLE_PRM["pdr"])
        elif m == "lpmd": o.lpmd(**ORACLE_PRM["lpmd"])
        elif m == "mhl": o.mhl(**ORACLE_PRM["mhl"])
        elif m in ("pm", "me", "pm+me"): o.quartets(**ORACLE_PRM["pm"])
        elif m in ("fdrp", "qfdrp"): o.fdrp(quantitative=(m == "qfdrp"), **ORACLE_PRM[m])
        else: raise ValueError(m)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    o.close()
    return n / best, best, n


def f32bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


def parity_contig(rows, nb, tid, n_proc):
    """Engine rows of the full pass (host arrays) restricted to contig `tid` against the oracle on that contig's reads."""
    import oracle_parallel as OP
    out = {}
    L = int(nb["start"][-1]) + 4096 if nb["n_reads"] else 1
    for m in ALL7:
        if m not in rows:
            continue
        t0 = time.perf_counter()
        try:
            if m == "lpmd":
                continue  # global scalar: checked by additivity below
            span = (0, min(FDRP_PARITY_SPAN, L)) if m in ("fdrp", "qfdrp") else None
            key = "p1" if m in ("pm", "me") else "pos"
            g = rows[m]
            sel = np.asarray(g["tid"]) == tid
            if span is not None:
                sel &= np.asarray(g[key]) < span[1]
            want, info = OP.run(nb, "quartets" if m in ("pm", "me") else m, ORACLE_PRM[m], n_proc=n_proc, interval=span)
            ok, n_w = False, 0
            if want is not None:
                n_w = len(want[key])
                vkey = {"pdr": "pdr", "pm": "pm", "me": "me"}.get(m, "value")
                cols = [key] + (["p2", "p3", "p4"] if key == "p1" else []) + (["n_conc", "n_disc"] if m == "pdr" else [])
                ok = int(sel.sum()) == n_w and all(np.array_equal(np.asarray(g[c])[sel], want[c]) for c in cols) and \
                    np.array_equal(f32bits(np.asarray(g["value"])[sel]), f32bits(want[vkey]))
            out[m] = {"contig": f"tid {tid}" + (f", sites < {span[1]}" if span else ", whole contig"), "rows": int(sel.sum()), "oracle_rows": int(n_w),
                      "rows_bit_identical_to_oracle": bool(ok), "oracle_wall_s": round(info["wall_s"], 2), "oracle_cpu_s": round(info["cpu_s"], 2),
                      "oracle_procs": info["procs"], "seconds": round(time.perf_counter() - t0, 2)}
        except Exception as e:  # a checker problem must never cost the bench line
            out[m] = {"error": repr(e)}
    return out


def run_reference(args):
    """`--impl reference`: the CPU oracle (port of metheor 0.1.9, one thread like the reference) on the headline measure set,
    each step = the whole parity contig of the same workload (regenerated on the CPU, bit-identical to the GPU's)."""
    import torch
    from metheor_b200 import synth_gpu as G
    contigs = G.genome(args.scale)
    t_all = time.perf_counter()
    nb = G.to_numpy_batch(G.make_contig("cpu", SEED, PARITY_TID, contigs[PARITY_TID][1], args.coverage))
    for _ in range(min(args.warmup, 2)):
        oracle_single(nb, "pm+me", 1_000_000)
    rps, times = [], []
    for _ in range(args.steps):
        r, dt, n = oracle_single(nb, "pm+me")
        rps.append(r); times.append(dt)
    v = float(np.mean(rps))
    # informational: the same oracle on ALL host cores (position bins + halo of the contig, tests/oracle_parallel.py) — what a
    # host could do at best with this algorithm; the reference itself is single-threaded, so `value` stays the one-thread figure
    all_cores = None
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_parallel as OP
        best = None
        for _ in range(2):
            _, info = OP.run(nb, "quartets", ORACLE_PRM["pm"], n_proc=os.cpu_count())
            if best is None or info["wall_s"] < best["wall_s"]:
                best = info
        all_cores = {"cores": int(best["procs"]), "unit": "reads/s",
                     # upper bound: only the oracle's own compute seconds (summed over the bins), spread perfectly over the cores
                     "compute_only_upper_bound": nb["n_reads"] * best["procs"] / max(best["cpu_s"], 1e-9),
                     # lower bound: wall clock of the fork pool incl. slicing the contig per bin and returning the rows
                     "wall_incl_process_pool_and_marshalling": nb["n_reads"] / best["wall_s"],
                     "oracle_compute_seconds_sum": best["cpu_s"], "wall_seconds": best["wall_s"],
                     "how": "the same oracle pass split into position bins (+ halo) of the contig, one process per core; not a mode the reference has"}
    except Exception as e:  # informational only
        all_cores = {"error": repr(e)}
    emit({"impl": "reference", "metric": "reads_per_sec", "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
          "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
          "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
          "config": {"workload": workload_name(args.coverage, args.scale), "measures": list(HEADLINE)},
          "cpu_baseline": {"value": v, "unit": "reads/s", "cores": 1, "kind": "port",
                           "sample": f"each step = all {nb['n_reads']} reads of contig tid {PARITY_TID} (chr21) of the workload, pm + me "
                                     f"(one quartet pass, pm.rs:85-128 / me.rs:90-132); C++ restatement of metheor 0.1.9, "
                                     f"single-threaded like the reference, not the Rust binary"},
          "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
          "all_cores": all_cores, "host": {"nproc": os.cpu_count()}, "wall_s": time.perf_counter() - t_all})


# ---------------------------------------------------------------------------------------------------------------------
def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # stray prints of libraries go to stderr; emit() writes the result to the real stdout
    ensure_built()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--coverage", type=float, default=COVERAGE)
    ap.add_argument("--scale", type=float, default=1.0, help="shrink every contig (quick runs); 1.0 = the BASELINE workload")
    ap.add_argument("--region-cost", type=float, default=REGION_COST_BP, help="N > 1: bases one contig boundary is worth when the bins are balanced (at 30x; 0 = length alone)")
    ap.add_argument("--measure-steps", type=int, default=5, help="timed passes of every non-headline measure set")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip every CPU oracle leg (cpu_baseline, parity)")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary legs (chr19 configs[1], 60x, BAM, tag)")
    ap.add_argument("--chr19-coverage", type=float, default=30.0)
    ap.add_argument("--wg60", type=float, default=60.0, help="coverage of the configs[3] leg (fdrp + qfdrp); 0 disables")
    ap.add_argument("--wg100", type=float, default=100.0, help="coverage of the configs[4] leg (all seven measures); 0 disables")
    ap.add_argument("--tag-reads", type=int, default=2_000_000)
    ap.add_argument("--bam-reads", type=int, default=2_000_000)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from metheor_b200 import engine
    from metheor_b200 import synth_gpu as G
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream()
    t_begin = time.perf_counter()
    import bench_chr19 as X

    contigs = G.genome(args.scale)
    ref_len = [l for _, l in contigs]
    region_cost = int(args.region_cost * args.scale * 30.0 / max(args.coverage, 1e-9)) if world > 1 else 0
    intervals = plan_bins_by_length(ref_len, world, region_cost)[rank]
    intervals_len = plan_bins_by_length(ref_len, world)[rank]  # the 60x / 100x legs: length alone
    t0 = time.perf_counter()
    wg, owned_R, owned_I = gen_shard(torch, dev, contigs, args.coverage, intervals, world)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    R_loc, I_loc = sum(b["n_reads"] for b in wg), sum(b["n_cpg"] for b in wg)
    tot = torch.tensor([owned_R, owned_I], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(tot)
    R, I = int(tot[0]), int(tot[1])  # the whole genome's reads / calls (halo copies not counted)

    def make_ctx(measures, flags, comm=False):
        c = engine.Context(engine.default_params(measures, flags=flags), ref_len, device=local_rank)
        c.set_stream(stream.cuda_stream)
        if comm and world > 1:  # the library's own communicator: the id travels through torch.distributed
            idt = torch.zeros(128, dtype=torch.uint8, device=dev)
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(engine.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            c.comm_init_rank(world, rank, bytes(idt.cpu().numpy().tobytes()))
        return c

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = torch.tensor([e0.elapsed_time(e1), wall * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            per = [torch.zeros_like(ms) for _ in range(world)]
            dist.all_gather(per, ms)
            timed.per_rank_ms = [float(x[0]) / steps for x in per]  # device time per pass of every rank: the max is what counts
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms[0]), float(ms[1])

    _prepared = {}

    def prep(batches):
        """ctypes structs of the batches, marshalled once (outside every timed region): the timed loops call the C ABI directly"""
        key = id(batches)
        if key not in _prepared:
            _prepared[key] = [engine.Context.prepare(b) for b in batches]  # raw pointers only: cleared whenever a workload is freed
        return _prepared[key]

    def resident_pass(ctx, batches, lp):
        ctx.reset()
        for mb in prep(batches):
            ctx.submit(mb)
        res = ctx.finish()
        if world > 1 and lp:
            res["lpmd_all_ranks"] = ctx.allreduce()  # the path's one exchange: NCCL sum of 4 int64 inside the library
        return res

    def measure_block(measures, steps, batches=wg, n_reads=R, n_calls=I, profile=True, owned_iv=None):
        """Resident timing + per-kernel profile of one measure set -> dict"""
        lp = "lpmd" in measures
        ctx = make_ctx(measures, engine.FLAG_KEEP_ON_DEVICE, comm=lp)
        res = {}
        def step():
            res.update(resident_pass(ctx, batches, lp))
        for _ in range(args.warmup):
            step()
        ms_dev, ms_wall = timed(step, steps)
        st = ctx.stats()
        dig = rows_digest(torch, dev, ctx, measures, owned=None if world == 1 else (owned_iv if owned_iv is not None else intervals))
        ms_step = ms_dev / steps
        per_rank = None
        if world > 1:
            info = torch.tensor([st["n_regions"], st["kernel_launches"], sum(b["n_reads"] for b in batches)], device=dev, dtype=torch.int64)
            allinfo = [torch.zeros_like(info) for _ in range(world)]
            dist.all_gather(allinfo, info)
            per_rank = [{"ms_per_step": round(m, 4), "regions": int(x[0]), "launches": int(x[1]), "reads_incl_halo": int(x[2])}
                        for m, x in zip(timed.per_rank_ms, allinfo)]
        blk = {"measures": list(measures), "steps": steps, "ms_per_step": ms_step, "wall_ms_per_step": ms_wall / steps, "per_rank": per_rank,
               "value": n_reads / (ms_step * 1e-3), "unit": "reads/s", "launches_per_step": int(st["kernel_launches"]),
               "regions_per_step": int(st["n_regions"]), "digest": dig}
        counts = torch.tensor([st["n_sites"]] + [dig[m][0] if m in dig else 0 for m in ALL7] + [st["fdrp_pair_ops"]], device=dev, dtype=torch.int64)
        if world > 1:  # sites / rows / pair-ops of the whole genome (halo sites are counted by more than one rank: an upper bound)
            dist.all_reduce(counts)
            dt = torch.tensor([v[1] for v in dig.values()], device=dev, dtype=torch.int64)
            dist.all_reduce(dt)
            blk["digest"] = {m: [int(counts[1 + ALL7.index(m)]), int(x)] for m, x in zip(dig, dt)}
        C = int(counts[0])
        rows = {m: int(counts[1 + ALL7.index(m)]) for m in measures if m != "lpmd"}
        blk.update(cpg_sites=C, rows=rows, cpgs_per_sec=C / (ms_step * 1e-3))
        if "lpmd" in res:
            l = res.get("lpmd_all_ranks") or res["lpmd"]
            blk["lpmd"] = {k: (float(v) if k == "lpmd" else int(v)) for k, v in l.items() if k != "pairs"}
        if st["fdrp_pair_ops"]:
            blk["pair_ops"] = int(counts[-1])
            blk["pair_ops_per_sec"] = int(counts[-1]) / (ms_step * 1e-3)
        ctx.close()
        if profile:
            pctx = make_ctx(measures, engine.FLAG_KEEP_ON_DEVICE | engine.FLAG_PROFILE)
            acc, PS = {}, 2
            for it in range(1 + PS):
                resident_pass(pctx, batches, False)
                if it >= 1:
                    for k, v in pctx.stats()["kernels"].items():
                        a = acc.setdefault(k, [0, 0.0]); a[0] += v["launches"]; a[1] += v["ms"]
            pst = pctx.stats()
            pctx.close()
            blk["kernels"] = {k: {"launches_per_step": a[0] / PS, "ms_per_step": a[1] / PS} for k, a in acc.items() if not k.startswith("~")}
            det = {k[1:]: round(a[1] / PS, 3) for k, a in acc.items() if k.startswith("~")}  # tile kernel vs per-site fall-back, inside k_mhl / k_fdrp*
            if det:
                blk["kernel_detail_ms"] = det
            if pst["fallback_sites_mhl"] or pst["fallback_sites_fdrp"]:
                blk["fallback_sites"] = {"mhl": int(pst["fallback_sites_mhl"]), "fdrp": int(pst["fallback_sites_fdrp"])}
        return blk

    def add_roofline(blk, traffic):
        """roofline (dominant kernel) + roofline_measure (SURVEY bytes / all kernels of the pass) for a measure block (this rank's share)."""
        peak, peak_src = peaks()
        measures, kern = blk["measures"], blk.get("kernels") or {}
        Rl, Il = R_loc, I_loc
        C = blk["cpg_sites"] // world if world > 1 else blk["cpg_sites"]
        Q = max([v for m, v in blk["rows"].items() if m in ("pm", "me")] + [0]) // world
        cand = {k: v for k, v in kern.items() if kernel_bytes(k, measures, Rl, Il, C, Q)}
        if not cand:
            return
        hot = max(cand, key=lambda k: cand[k]["ms_per_step"])
        nl = max(1.0, cand[hot]["launches_per_step"])
        ab = kernel_bytes(hot, measures, Rl, Il, C, Q)
        ach = ab / (cand[hot]["ms_per_step"] * 1e-3) / 1e9
        tr = (traffic.get("wg", {}).get("+".join(measures), {}) or {}).get(hot)
        blk["roofline"] = {"bound": "hbm", "kernel": hot, "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": ach / peak,
                           "traffic": (tr["dram_bytes"] / tr["launches"]) if tr else None, "launches_per_step": nl,
                           "algorithmic_bytes_per_launch": ab / nl, "ms_per_launch": cand[hot]["ms_per_step"] / nl,
                           "algorithmic_bytes_per_step": ab, "ms_per_step": cand[hot]["ms_per_step"]}
        if tr:
            blk["roofline"]["traffic_capture"] = tr
        sb = sum(survey_bytes(m, Rl, Il, C, Q) for m in measures)
        kms = sum(v["ms_per_step"] for v in kern.values())
        blk["roofline_measure"] = {"algorithmic_bytes": sb, "kernel_ms_sum": kms, "kernel_ms_over_step": kms / blk["ms_per_step"],
                                   "achieved": sb / (kms * 1e-3) / 1e9, "unit": "GB/s", "frac": sb / (kms * 1e-3) / 1e9 / peak,
                                   "formula": "SURVEY.md 8d: " + " + ".join(measures)}

    traffic = load_traffic()
    sampler = X.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---------------- headline: pm + me on the whole genome, resident ----------------
    head = measure_block(HEADLINE, args.steps)
    add_roofline(head, traffic)

    # ---------------- one block per measure, and the combined passes ----------------
    blocks, combos = {}, {}
    for ms in SINGLE:
        blocks[ms[0]] = measure_block(ms, args.measure_steps)
        add_roofline(blocks[ms[0]], traffic)
    for ms in COMBOS[1:]:
        combos["+".join(ms)] = measure_block(ms, args.measure_steps)
        add_roofline(combos["+".join(ms)], traffic)

    # ---------------- end to end through host buffers (headline measure set) ----------------
    e2e, eres_all = None, None
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    need = sum(batch_bytes(b, False) for b in wg)
    e_batches = wg
    if need * 2.5 > avail:  # bounded by host memory: use the contigs that fit, say so
        keep, acc = [], 0
        for b in wg:
            if (acc + batch_bytes(b, False)) * 2.5 > avail:
                break
            keep.append(b); acc += batch_bytes(b, False)
        e_batches = keep
    if e_batches:
        host = []
        for b in e_batches:
            hb = {k: v for k, v in b.items() if not isinstance(v, torch.Tensor)}
            for k in ("start", "end", "meta", "cpg_off", "cpg_pos", "meth", "meth_off"):
                if b.get(k) is not None:
                    hb[k] = torch.empty(b[k].shape, dtype=b[k].dtype, pin_memory=True)
                    hb[k].copy_(b[k])
            hb["cpg_rel"] = None
            host.append(hb)
        ectx = make_ctx(HEADLINE, 0)
        eres = {}

        host_mb = [engine.Context.prepare(hb) for hb in host]

        def e_step():
            ectx.reset()
            for mb in host_mb:
                ectx.submit(mb)
            eres.update(ectx.finish(copy=False))
        for _ in range(2):
            e_step()
        e_steps = max(3, min(args.steps, 5))
        _, e_wall = timed(e_step, e_steps)
        est = ectx.stats()
        Re = sum(b["n_reads"] for b in e_batches) if len(e_batches) < len(wg) or world > 1 else R
        if world > 1:
            Re = R
        e2e = {"value": Re / (e_wall / e_steps * 1e-3), "unit": "reads/s", "ms_per_step": e_wall / e_steps,
               "h2d_bytes_per_step": int(est["h2d_bytes"]), "d2h_bytes_per_step": int(est["d2h_bytes"]), "steps": e_steps,
               "measures": list(HEADLINE), "contigs": len(e_batches), "aggregate_h2d_GBps": int(est["h2d_bytes"]) * world / (e_wall / e_steps * 1e-3) / 1e9,
               "limiter": "PCIe: every pass moves the whole SoA host -> device (53-54 GB/s on one GPU; GPUs that share a PCIe switch uplink share "
                          "that bandwidth, so the aggregate stops near 180-210 GB/s on an 8-GPU box)",
               "wire_format": "SoA (mth_submit): pinned host arrays in the layout north_star names, one batch per contig; no host-side "
                              "encoding inside or outside the timed region; rows come back into the context's pinned buffers",
               "rows": {m: int(eres[m]["n"]) for m in HEADLINE}}
        if world == 1 and len(e_batches) == len(wg):
            assert all(eres[m]["n"] == head["rows"][m] for m in HEADLINE), "e2e rows differ from the resident pass"
        ectx.close()
        del host, eres
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- parity at full size + CPU baselines (rank 0, N = 1) ----------------
    parity, cpu_blocks, cpu_head = None, {}, None
    if world == 1 and not args.no_cpu_baseline:
        n_proc = os.cpu_count() or 1
        actx = make_ctx(ALL7, 0)
        ares = resident_pass(actx, wg, False)  # the whole genome, all seven measures, rows to the host
        adig = rows_digest(torch, dev, actx, ALL7)
        nb = G.to_numpy_batch(wg[PARITY_TID])
        parity = parity_contig(ares, nb, PARITY_TID, n_proc)
        # LPMD is one scalar over the file: additivity over contigs (each contig alone through the engine) + the oracle on the parity contig
        lsum = {k: 0 for k in ("n_read", "n_valid_read", "n_conc", "n_disc")}
        lctx = make_ctx(("lpmd",), engine.FLAG_KEEP_ON_DEVICE)
        l_par = None
        for b in wg:
            lctx.reset(); lctx.submit(b)
            l = lctx.finish()["lpmd"]
            if b["tid"] == PARITY_TID:
                l_par = l
            for k in lsum:
                lsum[k] += int(l[k])
        lctx.close()
        import oracle_parallel as OP
        lw, linfo = OP.run(nb, "lpmd", ORACLE_PRM["lpmd"], n_proc=n_proc)
        la = ares["lpmd"]
        parity["lpmd"] = {"contig": f"tid {PARITY_TID}, whole contig", "counters_identical_to_oracle": bool(all(int(l_par[k]) == int(lw[k]) for k in lsum)),
                          "value_bit_identical_to_oracle": bool(f32bits(l_par["lpmd"]) == f32bits(lw["lpmd"])),
                          "whole_genome_counters_equal_sum_over_contigs": bool(all(int(la[k]) == lsum[k] for k in lsum)),
                          "oracle_wall_s": round(linfo["wall_s"], 2)}
        # every timed pass produced the rows that were checked: digests of the single-measure / combined passes == all-seven pass
        same = {}
        for name, blk in list(blocks.items()) + list(combos.items()) + [("pm+me(headline)", head)]:
            same[name] = bool(all(blk["digest"][m] == adig[m] for m in blk["digest"]))
        parity["digests_equal_all_seven_pass"] = same
        parity["properties"] = {
            "rows_sorted_by_tid_pos": bool(all((np.diff(np.asarray(ares[m]["tid"], np.int64) * (1 << 32) + np.asarray(ares[m]["pos"], np.int64)) > 0).all()
                                               for m in ("pdr", "mhl", "fdrp", "qfdrp") if ares[m]["n"] > 1)),
            "pm_me_same_quartets": bool(ares["pm"]["n"] == ares["me"]["n"] and all(np.array_equal(ares["pm"][k], ares["me"][k]) for k in ("tid", "p1", "p2", "p3", "p4"))),
            "qfdrp_le_fdrp_same_sites": bool(ares["fdrp"]["n"] == ares["qfdrp"]["n"] and np.array_equal(ares["fdrp"]["pos"], ares["qfdrp"]["pos"])
                                             and bool((np.asarray(ares["qfdrp"]["value"]) <= np.asarray(ares["fdrp"]["value"]) + 1e-6).all())),
            "values_in_unit_interval": bool(all(float(np.nanmin(ares[m]["value"])) >= -1e-7 and float(np.nanmax(ares[m]["value"])) <= 1.0 + 1e-6
                                                for m in ("pdr", "mhl", "fdrp", "qfdrp", "pm", "me") if ares[m]["n"]))}
        for m in blocks:
            blocks[m]["parity_full_size"] = dict(parity.get(m, {}), digest_equals_checked_pass=same.get(m))
        actx.close()
        del ares
        # CPU baselines: one thread, bounded samples of the parity contig
        for m in ALL7:
            try:
                rps, dt, n = oracle_single(nb, m, CPU_SAMPLE[m])
                cpu_blocks[m] = {"value": rps, "unit": "reads/s", "cores": 1, "kind": "port", "seconds": dt,
                                 "sample": f"first {n} reads of contig tid {PARITY_TID} (chr21) of the workload, reference defaults; C++ restatement of "
                                           f"metheor 0.1.9, single-threaded like the reference; host has {os.cpu_count()} cores"}
                blocks[m]["cpu_baseline"] = cpu_blocks[m]
            except Exception as e:
                blocks[m]["cpu_baseline"] = {"error": repr(e)}
        try:
            rps, dt, n = oracle_single(nb, "pm+me")
            cpu_head = {"value": rps, "unit": "reads/s", "cores": 1, "kind": "port", "seconds": dt,
                        "sample": f"all {n} reads of contig tid {PARITY_TID} (chr21) of the workload, pm + me (one quartet pass); C++ restatement of "
                                  f"metheor 0.1.9, single-threaded like the reference; host has {os.cpu_count()} cores"}
        except Exception as e:
            cpu_head = {"error": repr(e)}
        del nb

    # ---------------- N > 1: the shards together == the whole genome on one GPU ----------------
    sharded_check = None
    _prepared.clear()
    if world > 1:
        del wg
        torch.cuda.empty_cache()
        if rank == 0:
            full, _, _ = gen_shard(torch, dev, contigs, args.coverage, [(t, 0, l) for t, l in enumerate(ref_len)], 1)
            octx = engine.Context(engine.default_params(ALL7, flags=engine.FLAG_KEEP_ON_DEVICE), ref_len, device=local_rank)
            octx.reset()
            for b in full:
                octx.submit(b)
            ol = octx.finish()["lpmd"]
            od = rows_digest(torch, dev, octx, ALL7)
            a7 = combos["+".join(ALL7)]
            sharded_check = {"digests_equal_single_gpu": {m: bool(a7["digest"][m] == od[m]) for m in od},
                             "lpmd_counters_equal_single_gpu": bool(all(int(a7["lpmd"][k]) == int(ol[k]) for k in ("n_read", "n_valid_read", "n_conc", "n_disc"))),
                             "what": "owned rows of all ranks (row count + order-independent 64-bit digest per measure, summed over the ranks) and the "
                                     "all-reduced LPMD counters against the same genome processed by rank 0 alone"}
            octx.close()
            del full
        barrier()
    else:
        del wg
    torch.cuda.empty_cache()

    # ---------------- the larger BASELINE configs, sharded over the ranks like the headline ----------------
    extra = {}

    def coverage_leg(key, cov, measures, what, steps=2):
        """One measure set at another coverage of the same genome (every rank generates its own position bin)."""
        nonlocal R_loc, I_loc
        try:
            t0 = time.perf_counter()
            wc, r_own, i_own = gen_shard(torch, dev, contigs, cov, intervals_len, world)
            torch.cuda.synchronize()
            gsec = time.perf_counter() - t0
            tt = torch.tensor([r_own, i_own], device=dev, dtype=torch.int64)
            if world > 1:
                dist.all_reduce(tt)
            R_save, I_save = R_loc, I_loc
            R_loc, I_loc = sum(b["n_reads"] for b in wc), sum(b["n_cpg"] for b in wc)
            blk = measure_block(measures, steps, batches=wc, n_reads=int(tt[0]), n_calls=int(tt[1]), owned_iv=intervals_len)
            add_roofline(blk, {})
            R_loc, I_loc = R_save, I_save
            blk.update(workload=f"{what}: {' + '.join(measures)}, {workload_name(cov, args.scale)}", reads=int(tt[0]), cpg_calls=int(tt[1]),
                       n_gpus=world, generate_seconds=gsec)
            _prepared.clear()
            del wc
            torch.cuda.empty_cache()
            return blk
        except Exception as e:  # a secondary leg must never cost the bench line
            _prepared.clear()
            torch.cuda.empty_cache()
            return {"error": repr(e)}

    if not args.no_extra:
        if args.wg60 > 0:
            extra["config3_wg60x_fdrp_qfdrp"] = coverage_leg("config3", args.wg60, ("fdrp", "qfdrp"), f"BASELINE.json configs[3] on {world} GPU(s)")
        if args.wg100 > 0:
            extra["config4_wg100x_all_seven"] = coverage_leg("config4", args.wg100, ALL7, f"BASELINE.json configs[4] on {world} GPU(s)")
    if world == 1 and not args.no_extra:
        try:
            c19, b19 = X.chr19_leg(args, torch, dev, stream, max(3, min(args.steps, 10)))
            extra["config1_chr19_pdr_lpmd"] = c19
        except Exception as e:
            extra["config1_chr19_pdr_lpmd"] = {"error": repr(e)}
            b19 = None
        torch.cuda.empty_cache()
        if args.bam_reads > 0 and not args.no_cpu_baseline and b19 is not None:
            try:
                extra["bam_end_to_end"] = X.bam_leg(b19, args.bam_reads, X.CONTIG_LEN)
            except Exception as e:
                extra["bam_end_to_end"] = {"error": repr(e)}
        if args.tag_reads > 0 and not args.no_cpu_baseline:
            try:
                extra["tag"] = X.tag_leg(args.tag_reads, X.CONTIG_LEN)
            except Exception as e:
                extra["tag"] = {"error": repr(e)}

    if rank == 0:
        ms_step = head["ms_per_step"]
        line = {"metric": "reads_per_sec", "value": head["value"], "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
                "dtype": DTYPE, "data": "synthetic",
                "config": {"workload": workload_name(args.coverage, args.scale), "measures": list(HEADLINE), "reads": R, "cpg_calls": I,
                           "cpg_sites": head["cpg_sites"], "rows": head["rows"], "seed": SEED, "contigs": len(contigs),
                           "l2": "inputs (%.1f GB per step) larger than L2" % ((16 * R + 4 * I + 8 * R) / 1e9),
                           "parallelism": (f"one genome cut into {world} position bins (+{HALO}-bp halo), one bin per GPU, balanced on bases + {region_cost} x contigs "
                                           f"(a contig part is a region with a fixed cost; the 60x / 100x legs: on length alone); NCCL all-reduce of LPMD's 4 "
                                           f"counters inside the library (mth_allreduce), once per pass") if world > 1 else "single GPU",
                           "generate_seconds": gen_s},
                "cpgs_per_sec": head["cpgs_per_sec"], "wall_ms_per_step": head["wall_ms_per_step"],
                "e2e": e2e, "gpu_launches": int(head["launches_per_step"] * args.steps), "launches_per_step": head["launches_per_step"],
                "roofline": head.get("roofline"), "roofline_measure": head.get("roofline_measure"), "kernels": head.get("kernels"),
                "cpu_baseline": cpu_head, "parity_full_size": parity, "measures": blocks, "combined": combos, "sharded_check": sharded_check,
                "clocks": clocks, "bench_wall_s": None}
        line.update(extra)
        line["bench_wall_s"] = time.perf_counter() - t_begin
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
