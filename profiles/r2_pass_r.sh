#!/bin/bash
# round 2, pass r: k_tag with 8 lanes per read (8 bases per lane and step)
O=gpurun_out/${1:-r2r}; mkdir -p $O
timeout 900 python -m pytest tests/test_tag.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -8 $O/pytest.log
timeout 600 python - > $O/tag_leg.json 2> $O/tag_leg.err <<'PY'
import json, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import bench_chr19 as X
r = X.tag_leg(2_000_000, X.CONTIG_LEN)
print(json.dumps(r))
PY
tail -3 $O/tag_leg.err; cat $O/tag_leg.json
