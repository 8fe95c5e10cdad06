OUT=gpurun_out/r01b; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -5 $OUT/pytest_gpu.log
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $OUT/sanitizer.log 2>&1; tail -4 $OUT/sanitizer.log
for v in "" metheor_b200/csrc/variant_minb6.so metheor_b200/csrc/variant_minb8.so; do
  METHEOR_B200_LIB=$v timeout 300 python bench.py --steps 5 --no-cpu-baseline > $OUT/bench_$(basename "$v" .so).json 2>>$OUT/bench.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_$(basename "$v" .so).json"))
print("$v", d["ms_per_step"], {k:round(v["ms_per_step"],4) for k,v in d["kernels"].items()})
PY
done
