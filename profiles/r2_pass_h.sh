#!/bin/bash
O=gpurun_out/${1:-r2h}; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -15 $O/pytest.log
timeout 900 python - > $O/bam_leg.json 2> $O/bam_leg.err <<'PY'
import json, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import bench_chr19 as X
b, _ = X.make_workload(0, 30.0, X.CONTIG_LEN)
print(json.dumps(X.bam_leg(b, 2_000_000, X.CONTIG_LEN)))
PY
tail -c 3000 $O/bam_leg.json; tail -c 1500 $O/bam_leg.err
