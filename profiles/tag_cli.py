"""`metheor tag` end to end on a synthetic BAM (plain 150M reads without XM) and a random chr19-sized FASTA: wall time and
the host's per-stage seconds (--stats).  Usage: python profiles/tag_cli.py [n_reads]  -> one JSON line."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metheor_b200 import batch as B  # noqa: E402
from metheor_b200 import host, synth, synth_bam  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    length = 58_617_616
    b, _ = synth.chr19_like(coverage=max(1.0, n * 150 / length * 1.2))
    sub = B.slice_reads(b, 0, min(n, b["n_reads"]))
    rng = np.random.default_rng(3)
    genome = rng.choice(np.frombuffer(b"ACGT", np.uint8), length, p=[0.29, 0.21, 0.21, 0.29])
    with tempfile.TemporaryDirectory() as d:
        fa, bam, out, st = (os.path.join(d, x) for x in ("g.fa", "in.bam", "out.sam", "st.json"))
        with open(fa, "wb") as f:
            f.write(b">chr19\n")
            full = length // 60
            body = np.empty((full, 61), np.uint8)
            body[:, :60] = genome[:full * 60].reshape(full, 60)
            body[:, 60] = 10
            f.write(body.tobytes() + genome[full * 60:].tobytes() + b"\n")
        info = synth_bam.write_bam(bam, [("chr19", length)], [sub], threads=os.cpu_count() or 8, with_xm=False)
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            host.tag(bam, out, fa, stats_json=st)
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, json.load(open(st)))
        lines = sum(1 for ln in open(out) if not ln.startswith("@"))
        print(json.dumps({"records": info["records"], "bam_bytes": info["bytes_compressed"], "sam_bytes": os.path.getsize(out),
                          "output_records": lines, "seconds": best[0], "reads_per_sec": info["records"] / best[0],
                          "stage_seconds": best[1]["seconds"], "host_threads": os.cpu_count()}))


if __name__ == "__main__":
    main()
