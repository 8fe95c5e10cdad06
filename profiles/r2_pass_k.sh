#!/bin/bash
O=gpurun_out/${1:-r2k}; mkdir -p $O
python - > $O/out.txt 2>&1 <<'PY'
import json, sys, time, tempfile, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import bench_chr19 as X
from metheor_b200 import batch as B, synth_bam, host
b, _ = X.make_workload(0, 30.0, X.CONTIG_LEN)
d = tempfile.mkdtemp()
bam = os.path.join(d, "s.bam")
sub = B.slice_reads(b, 0, 2_000_000)
synth_bam.write_bam(bam, [("chr19", X.CONTIG_LEN)], [sub], threads=16)
for m in ("pm", "pdr", "pm", "me", "qfdrp", "pm"):
    for rep in range(4):
        st = os.path.join(d, "st.json")
        t0 = time.perf_counter()
        host.run(m, bam, os.path.join(d, "o.tsv"), stats_json=st)
        dt = time.perf_counter() - t0
        s = json.load(open(st))
        print(m, rep, "%.3f s" % dt, s["seconds"], s["device_decode"], {k: round(v["ms"], 2) for k, v in s["gpu"][0]["kernels"].items()})
PY
tail -c 6000 $O/out.txt
