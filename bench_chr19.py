"""bench_chr19.py — secondary legs of bench.py: BASELINE.json configs[1] (`pdr` + `lpmd` over synthetic 30x WGBS of a
chr19-sized contig; the round-1 headline, kept for continuity), the BAM -> TSV leg and the `tag` leg.

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
        os.sync()          # the file was just written: let the write-back finish before anything is timed
        time.sleep(1.0)
        out.update(records=info["records"], bam_bytes=info["bytes_compressed"], uncompressed_bytes=info["bytes_uncompressed"],
                   host_threads=os.cpu_count())
        for m in MEASURES:
            tsv, st = os.path.join(d, f"{m}.tsv"), os.path.join(d, f"{m}.json")
            best = None
            for _ in range(7):  # (the host side of these boxes is shared: single runs vary a lot)
                t0 = time.perf_counter()
                host.run(m, bam, tsv, stats_json=st)
                dt = time.perf_counter() - t0
                if best is None or dt < best[0]:
                    best = (dt, json.load(open(st)))
            t0 = time.perf_counter()
            r = subprocess.run([oracle_lib.CLI_PATH, m, "-i", bam, "-o", tsv + ".oracle"], capture_output=True, text=True)
            dt_o = time.perf_counter() - t0
            same = r.returncode == 0 and open(tsv, "rb").read() == open(tsv + ".oracle", "rb").read()
            # the same file through the host-side decoder (all cores: inflate + record decode on the CPU) for comparison
            best_h = None
            for _ in range(3):
                t0 = time.perf_counter()
                host.run(m, bam, tsv + ".hostdec", stats_json=st, decode_host=1)
                dt = time.perf_counter() - t0
                if best_h is None or dt < best_h[0]:
                    best_h = (dt, json.load(open(st)))
            same_h = open(tsv, "rb").read() == open(tsv + ".hostdec", "rb").read()
            out[m] = {"reads_per_sec": info["records"] / best[0], "seconds": best[0], "decode": best[1].get("decode"),
                      "device_decode": best[1].get("device_decode"),
                      "host_decode_path": {"reads_per_sec": info["records"] / best_h[0], "seconds": best_h[0], "stage_seconds": best_h[1]["seconds"],
                                           "tsv_identical": bool(same_h)},
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


def chr19_leg(args, torch, dev, stream, steps):
    """BASELINE.json configs[1] on one GPU: resident value, per-kernel roofline, end-to-end legs, CPU oracle, full-size parity."""
    from metheor_b200 import batch as B
    from metheor_b200 import engine
    b, sites = make_workload(0, args.chr19_coverage, CONTIG_LEN)
    R, I = b["n_reads"], b["n_cpg"]
    view = {np.dtype("uint32"): np.int32, np.dtype("uint16"): np.int16, np.dtype("uint64"): np.int64}
    host, devb = dict(b), dict(b)
    for k in ("start", "end", "meta", "cpg_off", "cpg_pos", "cpg_rel", "meth"):
        t = torch.from_numpy(b[k].view(view.get(b[k].dtype, b[k].dtype)))
        host[k] = t.pin_memory()
        devb[k] = t.to(dev)

    def make_ctx(flags):
        c = engine.Context(engine.default_params(MEASURES, flags=flags), [CONTIG_LEN], device=dev.index or 0)
        c.set_stream(stream.cuda_stream)
        return c

    def timed(step, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(n):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3

    ctx = make_ctx(engine.FLAG_KEEP_ON_DEVICE)
    rows = {}

    def step_resident():
        ctx.reset(); ctx.submit(devb); rows.update(ctx.finish())

    for _ in range(max(3, args.warmup)):
        step_resident()
    ms_dev, ms_wall = timed(step_resident, steps)
    st = ctx.stats()
    C, n_rows, ms_step = st["n_sites"], rows["pdr"]["n"], ms_dev / steps

    pctx = make_ctx(engine.FLAG_KEEP_ON_DEVICE | engine.FLAG_PROFILE)
    for _ in range(3):
        pctx.reset(); pctx.submit(devb); pctx.finish()
    acc, PSTEPS = {}, 5
    for _ in range(PSTEPS):
        pctx.reset(); pctx.submit(devb); pctx.finish()
        for k, v in pctx.stats()["kernels"].items():
            a = acc.setdefault(k, [0, 0.0]); a[0] += v["launches"]; a[1] += v["ms"]
    kern = {k: {"launches_per_step": a[0] / PSTEPS, "ms_per_step": a[1] / PSTEPS} for k, a in acc.items()}
    pctx.close()
    alg = {"k_ingest": 16 * R + 6 * I + 32,   # SURVEY 8d LPMD: meta + cpg_off + meth per read, cpg_pos + cpg_rel per call
           "k_pdr_scatter": 5 * I + 8 * C,    # the kernel's own compulsory bytes: cpg_pos + flag byte per call, 2 u32 counters per site
           "k_pdr_gather": 16 * R + 4 * I + 8 * C}
    hot = max((k for k in kern if k in alg), key=lambda k: kern[k]["ms_per_step"])
    peak, peak_src = peaks()
    ach = alg[hot] / (kern[hot]["ms_per_step"] * 1e-3) / 1e9
    step_alg = (16 * R + 4 * I + 12 * C) + (16 * R + 6 * I + 16)
    kms = sum(v["ms_per_step"] for v in kern.values())
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("chr19", {}).get(hot, {}).get("dram_bytes")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": hot, "achieved": ach, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "algorithmic_bytes_per_launch": alg[hot], "ms_per_launch": kern[hot]["ms_per_step"]}
    roofline_step = {"algorithmic_bytes": step_alg, "kernel_ms_sum": kms, "kernel_ms_over_step": kms / ms_step,
                     "achieved": step_alg / (ms_step * 1e-3) / 1e9, "unit": "GB/s", "frac": step_alg / (ms_step * 1e-3) / 1e9 / peak}

    # ---- end to end: host (pinned) buffers in, rows in the context's pinned buffers out ----
    def pin(d):
        d = dict(d)
        for k, v in list(d.items()):
            if isinstance(v, np.ndarray):
                t = torch.from_numpy(v.view({np.dtype("uint16"): np.int16, np.dtype("uint32"): np.int32}.get(v.dtype, v.dtype)))
                d[k] = t.pin_memory() if t.numel() else t
        return d

    CH = 1 << 21
    hostd = [pin(B.to_compact(B.slice_reads(b, lo, min(R, lo + CH)), dense=True)) for lo in range(0, R, CH)]
    ectx = make_ctx(0)
    eres = {}

    def make_step(submit, payload):
        def step():
            ectx.reset(); submit(payload); eres.update(ectx.finish(copy=False))
        return step

    def submit_dense(chunks):
        for hc in chunks:
            ectx.submit_compact(hc)

    e_steps = max(3, min(steps, 10))
    e2e = {}
    for name, st_fn in (("soa", make_step(ectx.submit, host)), ("dense", make_step(submit_dense, hostd))):
        for _ in range(3):
            st_fn()
        _, e_wall = timed(st_fn, e_steps)
        est = ectx.stats()
        assert eres["pdr"]["n"] == n_rows
        e2e[name] = {"value": R / (e_wall / e_steps * 1e-3), "unit": "reads/s", "ms_per_step": e_wall / e_steps,
                     "h2d_bytes_per_step": int(est["h2d_bytes"]), "d2h_bytes_per_step": int(est["d2h_bytes"]), "steps": e_steps}
    cpu = parity_full = cpu_all = None
    if not args.no_cpu_baseline:
        rps, dt, nn = oracle_time(b, CPU_SAMPLE_READS, steps=2)
        cpu = {"value": rps, "unit": "reads/s", "cores": 1, "kind": "port", "seconds": dt,
               "sample": f"all {nn} reads of the workload (pdr+lpmd, defaults); C++ restatement of metheor 0.1.9, single-threaded like the "
                         f"reference; host has {os.cpu_count()} cores"}
        try:
            parity_full = parity_full_size(eres, R)
        except Exception as e:
            parity_full = {"error": repr(e)}
        if (os.cpu_count() or 1) > 1:
            try:
                cpu_all = oracle_all_cores(b, os.cpu_count())
            except Exception as e:
                cpu_all = {"error": repr(e)}
    out = {"workload": WORKLOAD, "measures": list(MEASURES), "reads": R, "cpg_calls": I, "cpg_sites": int(C), "pdr_rows": int(n_rows),
           "steps": steps, "value": R / (ms_step * 1e-3), "unit": "reads/s", "ms_per_step": ms_step, "wall_ms_per_step": ms_wall / steps,
           "cpgs_per_sec": C / (ms_step * 1e-3), "launches_per_step": int(st["kernel_launches"]),
           "e2e": dict(e2e["soa"], wire_format="SoA (mth_submit: the layout BASELINE.json's north_star names), one batch; nothing "
                                               "happens to the reads on the host inside or outside the timed region"),
           "e2e_dense_preencoded": dict(e2e["dense"], wire_format="compact + dense block encodings (mth_submit_compact), batches of <= "
                                        f"{CH} reads; the host-side ENCODE is NOT in the timed region (the C++ host pays ~0.04 s per M "
                                        "reads for it on 16 threads: bam_end_to_end.stage_seconds.assemble)"),
           "roofline": roofline, "roofline_step": roofline_step, "kernels": kern, "cpu_baseline": cpu, "parity_full_size": parity_full,
           "cpu_baseline_all_cores": cpu_all, "lpmd": float(rows["lpmd"]["lpmd"]),
           "pdr_path": {1: "scatter", 2: "gather", 3: "scatter+gather(hazard sites)"}.get(st["pdr_path"])}
    ctx.close(); ectx.close()
    return out, b
