#!/bin/bash
# ncu --set full of the GPU inflate kernel on the bench BAM (2 M reads, 14 064 BGZF members)
O=gpurun_out/${1:-R2d}; mkdir -p $O
cat > /tmp/inflate_once.py <<'PY'
import sys, os, tempfile
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import bench_chr19 as X
from metheor_b200 import bamdec, batch as B, synth_bam
b, _ = X.make_workload(0, 30.0, X.CONTIG_LEN)
d = tempfile.mkdtemp()
bam = os.path.join(d, "s.bam")
synth_bam.write_bam(bam, [("chr19", X.CONTIG_LEN)], [B.slice_reads(b, 0, 2_000_000)], threads=16)
data = open(bam, "rb").read()
mem = bamdec.bgzf_members(data)
out, status, ms = bamdec.inflate_members(data, mem)
print(len(mem), len(out), ms, int(status.any()))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bgzf_inflate -c 1 -o $O/prof_inflate -f python /tmp/inflate_once.py > $O/inflate.log 2>&1
ncu -i $O/prof_inflate.ncu-rep --page raw --csv > $O/prof_inflate_raw.csv 2>/dev/null
ncu -i $O/prof_inflate.ncu-rep --page source --csv --print-source cuda,sass > $O/src_k_bgzf_inflate.csv 2>/dev/null
gzip -9 $O/src_k_bgzf_inflate.csv; rm -f $O/prof_inflate.ncu-rep
tail -3 $O/inflate.log; ls -la $O
