#!/bin/bash
# round 2, pass t: mixed-site list for the PM/ME gather passes; MHL staging unrolled
O=gpurun_out/${1:-r2t}; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -4 $O/pytest.log
python profiles/wg_pass.py --sets pm+me,pm,mhl --warm 2 --passes 5 2>/dev/null | tail -1 > $O/timing.json
python profiles/wg_pass.py --sets pm+me,mhl --warm 1 --profile 2>/dev/null | tail -1 > $O/kernels.json
python - <<PY
import json
d=json.load(open("$O/timing.json"))
for s in d["sets"]: print(s["name"], "wall per pass ms", round(1e3*s["seconds"]/s["passes"],3))
d=json.load(open("$O/kernels.json"))
for s in d["sets"]: print(s["name"], s.get("fallback_sites"), s["kernels_ms"])
PY
