#!/bin/bash
O=gpurun_out/${1:-r2l}; mkdir -p $O
METHEOR_DEBUG_TIMING=1 python - > $O/out.txt 2>&1 <<'PY'
import json, sys, time, tempfile, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import bench_chr19 as X
from metheor_b200 import batch as B, synth_bam, host
b, _ = X.make_workload(0, 30.0, X.CONTIG_LEN)
d = tempfile.mkdtemp()
bam = os.path.join(d, "s.bam")
sub = B.slice_reads(b, 0, 2_000_000)
synth_bam.write_bam(bam, [("chr19", X.CONTIG_LEN)], [sub], threads=16)
for m, kw in (("pdr", {}), ("pdr", dict(decode_host=1)), ("lpmd", {}), ("lpmd", {}), ("lpmd", dict(decode_host=1)), ("lpmd", {}), ("mhl", {}), ("lpmd", {})):
    st = os.path.join(d, "st.json")
    t0 = time.perf_counter()
    host.run(m, bam, os.path.join(d, "o.tsv"), stats_json=st, **kw)
    dt = time.perf_counter() - t0
    s = json.load(open(st))
    print(m, kw, "%.3f s" % dt, s["seconds"]["stream"], {k: s["device_decode"][k] for k in ("create_s", "reserve_s", "stage_s", "window_s", "submit_s")}, flush=True)
PY
cat $O/out.txt | cut -c1-400
