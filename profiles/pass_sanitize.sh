OUT=gpurun_out/${1:-r02s}; mkdir -p $OUT
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -3 $OUT/memcheck_smoke.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "compact or pairs or dense or long_spans or batches or more_than_64" > $OUT/memcheck_tests.log 2>&1; echo "memcheck tests rc=$?"; tail -4 $OUT/memcheck_tests.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"; tail -3 $OUT/racecheck_smoke.log
