#!/bin/bash
O=gpurun_out/${1:-r2i}; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_bamdec.py tests/test_cli_gpu.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 900 python - > $O/bam_leg.json 2> $O/bam_leg.err <<'PY'
import json, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import bench_chr19 as X
b, _ = X.make_workload(0, 30.0, X.CONTIG_LEN)
r = X.bam_leg(b, 2_000_000, X.CONTIG_LEN)
print(json.dumps(r))
for m in ("pdr", "lpmd"):
    print(m, r[m]["reads_per_sec"], r[m]["device_decode"], "host path", r[m]["host_decode_path"]["reads_per_sec"], r[m]["tsv_identical_to_oracle"], file=sys.stderr)
PY
tail -c 1500 $O/bam_leg.err
