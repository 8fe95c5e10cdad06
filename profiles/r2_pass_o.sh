#!/bin/bash
O=gpurun_out/${1:-r2o}; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
python profiles/wg_pass.py --contigs 0,1 --sets fdrp,qfdrp,fdrp+qfdrp --warm 1 --profile 2>/dev/null | tail -1 > $O/fdrp.json
python - <<PY
import json
d=json.load(open("$O/fdrp.json"))
for s in d["sets"]: print(s["name"], s["fallback_sites"], {k:v for k,v in s["kernels_ms"].items() if k.startswith("~") or k.startswith("k_fdrp") or k.startswith("k_qfdrp")})
PY
