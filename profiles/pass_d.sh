OUT=gpurun_out/${1:-r01g}; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -15 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 10 > $OUT/bench.json 2>$OUT/bench.err; tail -3 $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print(round(d["ms_per_step"],4), {k:round(v["ms_per_step"],4) for k,v in d["kernels"].items()})
print("e2e", d["e2e"]); print("e2e_soa", d["e2e_soa"]); print("roofline", d["roofline"]); print("cpu", d["cpu_baseline"])
PY
