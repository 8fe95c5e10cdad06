OUT=gpurun_out/${1:-r01n}; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -12 $OUT/pytest_gpu.log
python profiles/measure_all.py > $OUT/measure_all.jsonl 2> $OUT/err.log; tail -2 $OUT/err.log
python - <<PY
import json
for l in open("$OUT/measure_all.jsonl"):
    d=json.loads(l); print(d['measures'], d['ms_per_pass'], 'ms', d['rows']); print('   ', {k:v for k,v in d['kernels_ms'].items() if v>0.03})
PY
