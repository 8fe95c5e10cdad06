#!/usr/bin/env python
"""bench.py — BASELINE.json configs[1]: `pdr` + `lpmd` over synthetic 30x WGBS of a chr19-sized contig.

One "step" = one full pass of the hot path over the whole read set (11.7 M reads, ~31 M CpG calls): ingest
(validation + site marking + LPMD), site dictionary, PDR counters, row emission.

  value : reads/s with the SoA batch already resident in HBM (device pointers handed to mth_submit, rows left in HBM)
  e2e   : reads/s through the same C-ABI calls with HOST (pinned) buffers: H2D of the batch and D2H of the rows inside
          the timed region
  roofline : the slowest kernel of the step, algorithmic bytes / its CUDA-event time, vs MEASURED_PEAKS.json
  cpu_baseline : the CPU oracle (C++ restatement of metheor 0.1.9, NOT the Rust binary) on a bounded sample, 1 core

N > 1 (torchrun): every rank processes its own chr19-sized contig (weak scaling, genomic sharding needs no data-path
collective); LPMD's four int64 counters are all-reduced over NCCL each step.
`--impl reference` times the CPU oracle alone (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20260101
CONTIG_LEN = 58_617_616
COVERAGE = 30.0
WORKLOAD = "pdr+lpmd, synthetic 30x WGBS, chr19-sized contig (58.6 Mb, ~1.1 M CpG sites, 150-bp SE reads, both strands)"
MEASURES = ("pdr", "lpmd")
CPU_SAMPLE_READS = 12_000_000  # the whole chr19-sized read set: a full pass takes the oracle only a few seconds


def make_workload(rank, coverage=COVERAGE, length=CONTIG_LEN):
    from metheor_b200 import synth
    b, sites = synth.chr19_like(seed=SEED + 1000 * rank, coverage=coverage, length=length)
    return b, sites


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def oracle_time(b, n_sample, steps=1):
    """CPU oracle (pdr + lpmd, reference defaults) on the first n_sample reads. -> (reads/s, seconds per pass)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from metheor_b200 import batch as B
    from oracle_lib import Oracle
    sub = B.slice_reads(b, 0, min(n_sample, b["n_reads"]))
    o = Oracle.from_soa(**B.to_oracle_soa([sub]))
    best = None
    for _ in range(steps):
        t0 = time.perf_counter()
        _ORACLE_LAST["pdr"] = o.pdr(10, 4, 10)
        _ORACLE_LAST["lpmd"] = o.lpmd(2, 16, 10)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    _ORACLE_LAST["reads"] = sub["n_reads"]
    return sub["n_reads"] / best, best, sub["n_reads"]


_ORACLE_LAST = {}  # results of the last oracle_time pass: the checker of `parity_full_size`


def parity_full_size(eng, R):
    """Engine rows of the end-to-end leg (whole bench workload, through the C ABI with host buffers) against the oracle pass
    that was timed as cpu_baseline: bit-exact rows and LPMD counters at BASELINE.json's full size."""
    w, wl = _ORACLE_LAST.get("pdr"), _ORACLE_LAST.get("lpmd")
    if w is None or _ORACLE_LAST.get("reads") != R:
        return None
    g, gl = eng["pdr"], eng["lpmd"]
    f32 = lambda a: np.asarray(a, np.float32).view(np.uint32)
    n = int(g["n"])
    rows_ok = (n == len(w["pos"]) and all(np.array_equal(np.asarray(g[k])[:n], w[k]) for k in ("tid", "pos", "n_conc", "n_disc"))
               and np.array_equal(f32(g["value"])[:n], f32(w["pdr"])))
    lpmd_ok = all(int(gl[k]) == int(wl[k]) for k in ("n_read", "n_valid_read", "n_conc", "n_disc")) and \
        (f32(gl["lpmd"]) == f32(wl["lpmd"]) or (np.isnan(gl["lpmd"]) and np.isnan(wl["lpmd"])))
    return {"reads": int(R), "pdr_rows": n, "pdr_rows_bit_identical_to_oracle": bool(rows_ok), "lpmd_identical_to_oracle": bool(lpmd_ok),
            "checked": "rows and counters of the e2e leg vs the oracle pass timed as cpu_baseline (tid, pos, n_conc, n_disc, f32 bit patterns)"}


_ORACLE_PARTS = None
_ORACLE_CACHE = {}


def _oracle_slice(args):
    """worker of oracle_all_cores: one oracle pass (pdr + lpmd) over one slice of the reads; returns its seconds"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle
    o = _ORACLE_CACHE.get(args)
    if o is None:  # warm-up call: build this slice's read set once (inherited through fork: nothing is pickled)
        o = _ORACLE_CACHE[args] = Oracle.from_soa(**_ORACLE_PARTS[args])
    t0 = time.perf_counter()
    o.pdr(10, 4, 10)
    o.lpmd(2, 16, 10)
    return time.perf_counter() - t0


def oracle_all_cores(b, n_proc):
    """What the CPU could do at best with every host core: the reads cut into n_proc position slices, one oracle process
    each (the reference itself is single-threaded and could only be run per contig this way; slices of ONE contig are
    not bit-identical at their edges, so this is an optimistic throughput bound, not a parity run)."""
    import multiprocessing as mp
    from metheor_b200 import batch as B
    R = b["n_reads"]
    cuts = [R * k // n_proc for k in range(n_proc + 1)]
    global _ORACLE_PARTS
    _ORACLE_PARTS = [B.to_oracle_soa([B.slice_reads(b, lo, hi)]) for lo, hi in zip(cuts[:-1], cuts[1:])]
    with mp.get_context("fork").Pool(n_proc) as pool:
        for _ in range(2):  # warm-up: every worker ends up holding every slice's decoded reads
            pool.map(_oracle_slice, list(range(n_proc)) * 4, chunksize=1)
        t0 = time.perf_counter()
        secs = pool.map(_oracle_slice, range(n_proc), chunksize=1)
        wall = time.perf_counter() - t0
    _ORACLE_PARTS = None
    return {"value": R / wall, "unit": "reads/s", "cores": n_proc, "kind": "port", "seconds": wall, "slowest_slice_seconds": max(secs),
            "sample": f"all {R} reads in {n_proc} position slices, one single-threaded oracle process per slice, timed from "
                      f"dispatch to the last slice done (decoded reads already in each worker's memory); optimistic bound: the "
                      f"reference has no such mode and slice edges are not bit-identical"}


def run_reference(args, rank):
    if rank != 0:
        return
    b, _ = make_workload(0)
    n = CPU_SAMPLE_READS
    times = []
    for _ in range(args.warmup):
        oracle_time(b, n // 8)
    t_all0 = time.perf_counter()
    rps = []
    for _ in range(args.steps):
        r, dt, nn = oracle_time(b, n)
        rps.append(r); times.append(dt)
    v = float(np.mean(rps))
    line = {"impl": "reference", "metric": "reads_per_sec", "value": v, "unit": "reads/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32/u64 integer + f32 finalisation", "data": "synthetic",
            "config": {"workload": WORKLOAD, "measures": list(MEASURES)},
            "cpu_baseline": {"value": v, "unit": "reads/s", "cores": 1, "kind": "port",
                             "sample": f"first {n} reads of the workload; C++ restatement of metheor 0.1.9 (single-threaded "
                                       f"like the reference), not the Rust binary"},
            "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "host": {"nproc": os.cpu_count()}, "wall_s": time.perf_counter() - t_all0}
    emit(line)


def bam_leg(b, n_reads, length):
    """BAM -> TSV through the shipped host (BGZF inflate + record decode on all cores, compact batches, GPU engine, TSV
    writer) next to the oracle's CLI (single thread, its own BAM reader) on the same file; outputs must be identical."""
    import tempfile
    from metheor_b200 import batch as B
    from metheor_b200 import host, synth_bam
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    oracle_lib.build()
    sub = B.slice_reads(b, 0, min(n_reads, b["n_reads"]))
    out = {}
    with tempfile.TemporaryDirectory() as d:
        bam = os.path.join(d, "synthetic.bam")
        info = synth_bam.write_bam(bam, [("chr19", length)], [sub], threads=os.cpu_count() or 8)
        out.update(records=info["records"], bam_bytes=info["bytes_compressed"], uncompressed_bytes=info["bytes_uncompressed"],
                   host_threads=os.cpu_count())
        for m in MEASURES:
            tsv, st = os.path.join(d, f"{m}.tsv"), os.path.join(d, f"{m}.json")
            best = None
            for _ in range(3):
                t0 = time.perf_counter()
                host.run(m, bam, tsv, stats_json=st)
                dt = time.perf_counter() - t0
                if best is None or dt < best[0]:
                    best = (dt, json.load(open(st)))
            t0 = time.perf_counter()
            r = subprocess.run([oracle_lib.CLI_PATH, m, "-i", bam, "-o", tsv + ".oracle"], capture_output=True, text=True)
            dt_o = time.perf_counter() - t0
            same = r.returncode == 0 and open(tsv, "rb").read() == open(tsv + ".oracle", "rb").read()
            out[m] = {"reads_per_sec": info["records"] / best[0], "seconds": best[0],
                      "uncompressed_MB_per_sec": info["bytes_uncompressed"] / best[0] / 1e6, "stage_seconds": best[1]["seconds"],
                      "cpu_oracle_cli_seconds": dt_o, "cpu_oracle_cli_reads_per_sec": info["records"] / dt_o,
                      "tsv_identical_to_oracle": bool(same)}
    return out


def tag_leg(n_reads, length, read_len=150, seed=7):
    """`tag` (XM synthesis): synthetic plain-`150M` reads over a random chr19-sized genome through the C ABI
    (mth_tag: host arrays in, XM strings in pinned host memory out), with the kernel's own device time and the Python oracle
    (single thread, bounded sample) beside it; the sample's tags must be identical."""
    from metheor_b200 import tag
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tag_oracle
    rng = np.random.default_rng(seed)
    genome = rng.choice(np.frombuffer(b"ACGT", np.uint8), length, p=[0.29, 0.21, 0.21, 0.29])
    genome[rng.integers(0, length, length // 500)] = ord("N")
    pos = np.sort(rng.integers(0, length - read_len, n_reads)).astype(np.int32)
    tid = np.zeros(n_reads, np.int32)
    rc = (rng.random(n_reads) < 0.5).astype(np.uint8)
    l_seq = np.full(n_reads, read_len, np.int32)
    cigar_off = np.arange(n_reads + 1, dtype=np.uint32)
    cigar = np.full(n_reads, read_len << 4, np.uint32)
    nb = (read_len + 1) // 2
    seq_off = (np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(nb))
    codes = np.array([1, 2, 4, 8, 8, 8, 15], np.uint8)  # A C G T T T N: bisulfite-like
    seq4 = (codes[rng.integers(0, 7, n_reads * nb)] << 4) | codes[rng.integers(0, 7, n_reads * nb)]
    out = {"reads": n_reads, "read_len": read_len, "genome_bases": length}
    with tag.Genome([length]) as g:
        t0 = time.perf_counter()
        g.set_contig(0, genome)
        out["genome_upload_seconds"] = time.perf_counter() - t0
        best, kms = None, None
        for _ in range(4):
            t0 = time.perf_counter()
            off, ln, xm, status = g.tag_arrays(tid, pos, rc, l_seq, cigar_off, cigar, seq_off, seq4, raw=True)
            dt = time.perf_counter() - t0
            if best is None or dt < best:
                best, kms = dt, g.last_kernel_ms()
        assert not status.any()
        bytes_per_read = nb + (read_len + 4) + read_len + 4 + 4 + 1 + 4 + 4 + 8 + 8 + 8 + 4 + 1  # SEQ, genome, tag, scalars / offsets
        out.update(reads_per_sec=n_reads / best, seconds=best, kernel_ms=kms, kernel_reads_per_sec=n_reads / (kms * 1e-3),
                   kernel_algorithmic_GBps=bytes_per_read * n_reads / (kms * 1e-3) / 1e9, algorithmic_bytes_per_read=bytes_per_read,
                   h2d_bytes=int(tid.nbytes + pos.nbytes + rc.nbytes + l_seq.nbytes + cigar_off.nbytes + cigar.nbytes + seq_off.nbytes + seq4.nbytes
                                 + 16 * (n_reads + 1)), d2h_bytes=int(xm.nbytes + 5 * n_reads))
        # CPU oracle on a bounded sample, tags compared
        m = min(n_reads, 20000)
        text = bytes(genome).decode()
        nt = "=ACMGRSVTWYHKDBN"
        t0 = time.perf_counter()
        same = True
        for i in range(m):
            sq = seq4[i * nb:(i + 1) * nb]
            s = "".join(nt[c >> 4] + nt[c & 15] for c in sq)[:read_len]
            want = tag_oracle.xm_string(16 if rc[i] else 0, int(pos[i]), [(read_len, "M")], s, text, length, False)
            got = bytes(xm[int(off[i]):int(off[i]) + int(ln[i])]).decode()
            same = same and (got == want)
        dt = time.perf_counter() - t0
        out.update(cpu_oracle_reads_per_sec=m / dt, cpu_oracle_sample=f"first {m} reads, single-threaded Python restatement of tag.rs",
                   tags_identical_to_oracle=bool(same))
    return out


class _DevI64:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def ensure_built():
    """The libraries are built in-tree by __graft_entry__.build() and travel with the snapshot; a bare checkout builds them here."""
    need = [os.path.join(ROOT, "metheor_b200", "csrc", "libmetheor_b200.so"), os.path.join(ROOT, "metheor_b200", "host", "libmetheor_host.so"),
            os.path.join(ROOT, "oracle", "_build", "liboracle.so")]
    if not all(os.path.exists(p) for p in need):
        if int(os.environ.get("LOCAL_RANK", "0")) == 0:
            import __graft_entry__
            __graft_entry__.build()
        else:  # under torchrun only one rank builds
            t0 = time.time()
            while not all(os.path.exists(p) for p in need) and time.time() - t0 < 600:
                time.sleep(1.0)
            time.sleep(2.0)


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line on stdout.  Everything else any library prints (NCCL's version banner, ...) was routed to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # stray prints of libraries go to stderr; emit() writes the result to the real stdout
    ensure_built()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--coverage", type=float, default=COVERAGE)
    ap.add_argument("--length", type=int, default=CONTIG_LEN)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tag-reads", type=int, default=2_000_000, help="reads of the extra `tag` (XM synthesis) leg; 0 disables it")
    ap.add_argument("--bam-reads", type=int, default=2_000_000,
                    help="also time the BAM -> TSV path (C++ host + engine) on a BAM of the first N reads; 0 disables")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from metheor_b200 import engine
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)

    b, sites = make_workload(rank, args.coverage, args.length)
    R, I = b["n_reads"], b["n_cpg"]
    view = {np.dtype("uint32"): np.int32, np.dtype("uint16"): np.int16, np.dtype("uint64"): np.int64}
    host, devb = dict(b), dict(b)
    for k in ("start", "end", "meta", "cpg_off", "cpg_pos", "cpg_rel", "meth"):
        t = torch.from_numpy(b[k].view(view.get(b[k].dtype, b[k].dtype)))
        host[k] = t.pin_memory()
        devb[k] = t.to(dev)
    stream = torch.cuda.current_stream()

    def make_ctx(flags):
        c = engine.Context(engine.default_params(MEASURES, flags=flags), [args.length], device=local_rank)
        c.set_stream(stream.cuda_stream)
        return c

    # The path's one exchange: LPMD's four int64 counters (NCCL sum).  It runs on a side stream behind an event, on a copy
    # of the counters, so a rank can start its next pass while the (tiny, latency-bound) collective is in flight; the
    # timed region ends only after the last collective has completed (the compute stream waits for it before e1).
    ar_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    ar_buf = [torch.zeros(4, dtype=torch.int64, device=dev) for _ in range(2)] if world > 1 else None
    ar_state = {"k": 0, "last": None}

    def allreduce_lpmd(ctx):
        if world > 1:
            t = torch.as_tensor(_DevI64(ctx.lpmd_counters_device_ptr(), 4), device=dev)
            buf = ar_buf[ar_state["k"] & 1]
            ar_state["k"] += 1
            if ar_state["last"] is not None:
                stream.wait_event(ar_state["last"])  # the buffer's previous collective (two passes ago at the latest) is done
            buf.copy_(t)
            ready = torch.cuda.Event()
            ready.record(stream)
            with torch.cuda.stream(ar_stream):
                ar_stream.wait_event(ready)
                dist.all_reduce(buf)
                done = torch.cuda.Event()
                done.record(ar_stream)
            ar_state["last"] = done
            ar_state["result"] = buf

    def allreduce_join():
        if world > 1 and ar_state["last"] is not None:
            stream.wait_event(ar_state["last"])

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
        allreduce_join()
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = torch.tensor([e0.elapsed_time(e1), wall * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms[0]), float(ms[1])

    # ---------------- resident (value) ----------------
    ctx = make_ctx(engine.FLAG_KEEP_ON_DEVICE)
    rows = {}

    def step_resident():
        ctx.reset()
        ctx.submit(devb)
        rows.update(ctx.finish())
        allreduce_lpmd(ctx)

    for _ in range(args.warmup):
        step_resident()
    launches0 = ctx.stats()["kernel_launches"]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, ms_wall = timed(step_resident, args.steps)
    st = ctx.stats()
    n_sites = st["n_sites"]
    n_rows = rows["pdr"]["n"]
    ms_step = ms_dev / args.steps
    total_reads = R * world
    value = total_reads / (ms_step * 1e-3)

    # ---------------- per-kernel times (profile flag: CUDA events around every kernel) ----------------
    pctx = make_ctx(engine.FLAG_KEEP_ON_DEVICE | engine.FLAG_PROFILE)
    PSTEPS = 5
    for _ in range(3):
        pctx.reset(); pctx.submit(devb); pctx.finish()
    # stats are reset by reset(): collect from the LAST step only, repeated PSTEPS times for an average
    acc = {}
    for _ in range(PSTEPS):
        pctx.reset(); pctx.submit(devb); pctx.finish()
        for k, v in pctx.stats()["kernels"].items():
            a = acc.setdefault(k, [0, 0.0])
            a[0] += v["launches"]; a[1] += v["ms"]
    kern = {k: {"launches_per_step": a[0] / PSTEPS, "ms_per_step": a[1] / PSTEPS} for k, a in acc.items()}
    pctx.close()
    C = n_sites
    alg = {  # algorithmic bytes per launch, DESIGN.md §5
        "k_ingest": 16 * R + 6 * I + 32,          # LPMD: meta+cpg_off+meth per read, cpg_pos+cpg_rel per CpG call, 4 counters
        "k_pdr_scatter": 16 * R + 4 * I + 8 * C,  # PDR: per-read fields, cpg_pos per call, 2 u32 counters per site
        "k_pdr_gather": 16 * R + 4 * I + 8 * C,
        "k_sites_count": C * 4, "k_sites_emit": C * 4, "pdr_rows_count": C * 12, "k_pdr_emit": C * 8 + n_rows * 20,
    }
    hot = max((k for k in kern if k in ("k_ingest", "k_pdr_scatter", "k_pdr_gather")), key=lambda k: kern[k]["ms_per_step"])
    peak, peak_src = peaks()
    ach = alg[hot] / (kern[hot]["ms_per_step"] * 1e-3) / 1e9
    step_alg = (16 * R + 4 * I + 12 * C) + (16 * R + 6 * I + 16)  # PDR + LPMD
    kern_ms_total = sum(v["ms_per_step"] for v in kern.values())
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(hot, {}).get("dram_bytes")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": hot, "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic, "algorithmic_bytes_per_launch": alg[hot],
                "ms_per_launch": kern[hot]["ms_per_step"]}
    roofline_step = {"algorithmic_bytes": step_alg, "kernel_ms_sum": kern_ms_total,
                     "achieved": step_alg / (kern_ms_total * 1e-3) / 1e9, "unit": "GB/s",
                     "frac": step_alg / (kern_ms_total * 1e-3) / 1e9 / peak}

    # ---------------- end to end through host buffers ----------------
    # The shipped host (metheor_b200/host) hands batches over in the compact wire format (mth_submit_compact) with the
    # dense block encodings: ~7 B per read + 1.125 B per call cross PCIe and the device expands them.  `e2e` times exactly
    # that call sequence with pinned host arrays; `e2e_compact` is the plain compact format (9 B + 2.125 B) and `e2e_soa`
    # mth_submit with the full SoA layout (24 B + 6 B).
    from metheor_b200 import batch as B

    def pin(d):
        d = dict(d)
        for k, v in list(d.items()):
            if isinstance(v, np.ndarray):
                t = torch.from_numpy(v.view({np.dtype("uint16"): np.int16, np.dtype("uint32"): np.int32}.get(v.dtype, v.dtype)))
                d[k] = t.pin_memory() if t.numel() else t
        return d

    hostc = pin(B.to_compact(b))
    # dense block encodings (MTH_CENC_START16 | MTH_CENC_DELTA8), handed over like the streaming host does: a few
    # batches of <= 2 M reads, so that the copy of one overlaps the expansion + ingest of the previous one
    CH = 1 << 21
    hostd = [pin(B.to_compact(B.slice_reads(b, lo, min(R, lo + CH)), dense=True)) for lo in range(0, R, CH)]
    ectx = make_ctx(0)
    eres = {}

    def make_step(submit, payload):
        def step():
            ectx.reset()
            submit(payload)
            eres.update(ectx.finish(copy=False))  # rows are read where the C ABI leaves them: the context's pinned host buffers
            allreduce_lpmd(ectx)
        return step

    e_steps = max(3, min(args.steps, 10))
    e2e = {}
    def submit_dense(chunks):
        for hc in chunks:
            ectx.submit_compact(hc)

    for name, st_fn in (("soa", make_step(ectx.submit, host)), ("compact", make_step(ectx.submit_compact, hostc)),
                        ("dense", make_step(submit_dense, hostd))):
        for _ in range(args.warmup):
            st_fn()
        _, e_wall = timed(st_fn, e_steps)
        est = ectx.stats()
        assert eres["pdr"]["n"] == n_rows
        assert (eres["lpmd"]["n_conc"], eres["lpmd"]["n_disc"]) == (rows["lpmd"]["n_conc"], rows["lpmd"]["n_disc"]) or world > 1
        e2e[name] = {"value": total_reads / (e_wall / e_steps * 1e-3), "unit": "reads/s", "ms_per_step": e_wall / e_steps,
                     "h2d_bytes_per_step": int(est["h2d_bytes"]), "d2h_bytes_per_step": int(est["d2h_bytes"]), "steps": e_steps}

    # clocks / throttle reasons were sampled from the start of the resident timed region to the end of the e2e one
    clocks = sampler.stop() if rank == 0 else None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rps, dt, nn = oracle_time(b, CPU_SAMPLE_READS, steps=3)
        cpu = {"value": rps, "unit": "reads/s", "cores": 1, "kind": "port", "seconds": dt,
               "sample": f"first {nn} reads of the workload (pdr+lpmd, defaults); C++ restatement of metheor 0.1.9, "
                         f"single-threaded like the reference; host has {os.cpu_count()} cores"}

    parity_full = None
    if cpu is not None:
        try:
            parity_full = parity_full_size(eres, R)
        except Exception as e:  # a checker problem must never cost the bench line
            parity_full = {"error": repr(e)}

    cpu_all = None
    if cpu is not None and (os.cpu_count() or 1) > 1:
        try:
            cpu_all = oracle_all_cores(b, os.cpu_count())
        except Exception as e:
            cpu_all = {"error": repr(e)}

    bam = None
    if rank == 0 and world == 1 and args.bam_reads > 0 and not args.no_cpu_baseline:
        try:
            bam = bam_leg(b, args.bam_reads, args.length)
        except Exception as e:  # the extra leg must never cost the bench line
            bam = {"error": repr(e)}

    tag_res = None
    if rank == 0 and world == 1 and args.tag_reads > 0 and not args.no_cpu_baseline:
        try:
            tag_res = tag_leg(args.tag_reads, args.length)
        except Exception as e:
            tag_res = {"error": repr(e)}

    lpmd_all = None
    if world > 1 and ar_state.get("result") is not None:
        torch.cuda.synchronize()
        tot = ar_state["result"].cpu().numpy()  # n_read, n_valid_read, n_conc, n_disc summed over the ranks
        lpmd_all = {"n_read": int(tot[0]), "n_conc": int(tot[2]), "n_disc": int(tot[3]),
                    "lpmd": float(np.float32(tot[3]) / np.float32(tot[2] + tot[3]))}

    if rank == 0:
        line = {"metric": "reads_per_sec", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u32/u64 integer + f32 finalisation", "data": "synthetic",
                "config": {"workload": WORKLOAD, "measures": list(MEASURES), "reads_per_gpu": R, "cpg_calls_per_gpu": I,
                           "cpg_sites_per_gpu": int(C), "pdr_rows_per_gpu": int(n_rows), "seed": SEED,
                           "l2": "inputs (%.0f MB per step) larger than L2" % ((16 * R + 6 * I + 8 * R) / 1e6),
                           "parallelism": f"genomic sharding x{world}, NCCL all-reduce of 4 LPMD counters"},
                "cpgs_per_sec": C * world / (ms_step * 1e-3), "wall_ms_per_step": ms_wall / args.steps,
                "e2e": dict(e2e["dense"], wire_format="compact + dense block encodings (mth_submit_compact, enc = START16 | DELTA8), "
                                                      f"{len(hostd)} batches of <= {CH} reads"),
                "e2e_compact": dict(e2e["compact"], wire_format="compact (mth_submit_compact, enc = 0), one batch"),
                "e2e_soa": dict(e2e["soa"], wire_format="SoA (mth_submit), one batch"),
                "gpu_launches": int((st["kernel_launches"] - 0) * args.steps),
                "launches_per_step": int(st["kernel_launches"]),
                "roofline": roofline, "roofline_step": roofline_step, "kernels": kern, "cpu_baseline": cpu, "parity_full_size": parity_full, "cpu_baseline_all_cores": cpu_all, "bam_end_to_end": bam, "tag": tag_res,
                "clocks": clocks, "pdr_path": {1: "scatter", 2: "gather", 3: "scatter+gather(hazard sites)"}.get(st["pdr_path"]), "lpmd": float(rows["lpmd"]["lpmd"]),
                "lpmd_all_ranks": lpmd_all}
        emit(line)
    ctx.close(); ectx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
