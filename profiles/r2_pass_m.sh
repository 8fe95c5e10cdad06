#!/bin/bash
O=gpurun_out/${1:-r2m}; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.log
python profiles/wg_pass.py --contigs 0,1 --sets mhl,fdrp --warm 1 --profile 2>/dev/null | tail -1 > $O/mhl.json
python - <<PY
import json
d=json.load(open("$O/mhl.json"))
for s in d["sets"]: print(s["name"], s["fallback_sites"], {k:v for k,v in s["kernels_ms"].items() if k.startswith("~") or k in ("k_mhl","k_fdrp")})
PY
