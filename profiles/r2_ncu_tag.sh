#!/bin/bash
# ncu --set full of k_tag on the bench's tag leg (2 M plain 150M reads over a chr19-sized genome)
O=gpurun_out/${1:-R2e}; mkdir -p $O
cat > /tmp/tag_once.py <<'PY'
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import bench_chr19 as X
import json
r = X.tag_leg(2_000_000, X.CONTIG_LEN)
print(json.dumps({k: r[k] for k in ("kernel_ms", "kernel_algorithmic_GBps", "tags_identical_to_oracle")}))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tag -c 1 -o $O/prof_tag -f python /tmp/tag_once.py > $O/tag.log 2>&1
ncu -i $O/prof_tag.ncu-rep --page raw --csv > $O/prof_tag_raw.csv 2>/dev/null
ncu -i $O/prof_tag.ncu-rep --page source --csv --print-source cuda,sass > $O/src_k_tag.csv 2>/dev/null
gzip -9 -f $O/src_k_tag.csv; rm -f $O/prof_tag.ncu-rep
tail -2 $O/tag.log; ls -la $O
