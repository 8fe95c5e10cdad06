#!/bin/bash
# round 2, pass u: FDRP tile instances of 16 / 8 sites for 60x / 100x
O=gpurun_out/${1:-r2u}; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "deep_piles or default_flags or dense_islands or fixture or position_bin" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
for cov in 60 100; do
  python profiles/wg_pass.py --contigs 18,19,20 --coverage $cov --sets fdrp+qfdrp --warm 1 --profile 2>/dev/null | tail -1 > $O/new_$cov.json
  METHEOR_FDRP_TILE=sparse32 python profiles/wg_pass.py --contigs 18,19,20 --coverage $cov --sets fdrp+qfdrp --warm 1 --profile 2>/dev/null | tail -1 > $O/old_$cov.json
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/*_*.json")):
    d=json.load(open(f))
    for s in d["sets"]:
        print(f.split("/")[-1], s["name"], s["reads"], s["fallback_sites"], s["rows"], {k:v for k,v in s["kernels_ms"].items() if "fdrp" in k})
PY
