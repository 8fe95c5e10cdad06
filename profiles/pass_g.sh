OUT=gpurun_out/${1:-r02g}; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -12 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 10 --no-cpu-baseline > $OUT/bench.json 2>$OUT/bench.err; tail -3 $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print(round(d["ms_per_step"],4), {k:round(v["ms_per_step"],4) for k,v in d["kernels"].items()})
for k in ("e2e","e2e_compact","e2e_soa"): print(k, round(d[k]["ms_per_step"],3), round(d[k]["value"]/1e9,3), d[k]["h2d_bytes_per_step"])
PY
