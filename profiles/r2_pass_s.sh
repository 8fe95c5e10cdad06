#!/bin/bash
# round 2, pass s: MHL tile kernel — per-site fallback for long reads, 4 CTAs per SM at whole-genome density
O=gpurun_out/${1:-r2s}; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -4 $O/pytest.log
python profiles/wg_pass.py --sets mhl,all7 --warm 1 --profile 2>/dev/null | tail -1 > $O/mhl.json
python - <<PY
import json
d=json.load(open("$O/mhl.json"))
for s in d["sets"]: print(s["name"], round(s["seconds"],4), s.get("fallback_sites"), {k:v for k,v in s["kernels_ms"].items() if "mhl" in k})
PY
