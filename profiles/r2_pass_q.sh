#!/bin/bash
# round 2, pass q: pipelined regions (deferred phase A / phase B)
O=gpurun_out/${1:-r2q}; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -15 $O/pytest.log
for mode in pipe serial; do
  if [ $mode = serial ]; then export METHEOR_NO_PIPELINE=1; else unset METHEOR_NO_PIPELINE; fi
  timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-extra > $O/bench_$mode.json 2> $O/bench_$mode.err; echo "rc=$?" >> $O/bench_$mode.err
  tail -c 300 $O/bench_$mode.err
  python - <<PY
import json
d=json.loads(open("$O/bench_$mode.json").readline())
print("$mode headline", d["ms_per_step"], d["value"])
for k,v in list(d["measures"].items())+list(d["combined"].items()): print("  ", k, v["ms_per_step"])
PY
done
