OUT=gpurun_out/r01c; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -5 $OUT/pytest_gpu.log
for v in "" $(ls metheor_b200/csrc/variant_*.so); do
  n=$(basename "$v" .so); n=${n:-default}
  METHEOR_B200_LIB=$v timeout 300 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench_$n.json 2>>$OUT/bench.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_$n.json"))
print("$n", round(d["ms_per_step"],4), {k:round(v["ms_per_step"],4) for k,v in d["kernels"].items()}, "e2e", round(d["e2e"]["ms_per_step"],3))
PY
done
