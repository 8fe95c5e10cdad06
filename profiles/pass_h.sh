#!/bin/bash
# Sanitizer runs for the rewritten k_ingest + whole-genome-scale resident run.   gpurun --timeout 400 -- 'bash profiles/pass_h.sh r05s'
OUT=gpurun_out/${1:-r05s}; mkdir -p $OUT
timeout 120 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -2 $OUT/memcheck_smoke.log
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py tests/test_zz_lpmd_windows.py -m gpu -x -q -k "dense or more_than_64 or batches or windows" > $OUT/memcheck_tests.log 2>&1; echo "memcheck tests rc=$?"; tail -3 $OUT/memcheck_tests.log
timeout 120 compute-sanitizer --tool racecheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"; tail -2 $OUT/racecheck_smoke.log
timeout 200 python profiles/scale_wg.py > $OUT/scale_wg.jsonl 2> $OUT/scale_wg.err; echo "scale rc=$?"; cut -c1-260 $OUT/scale_wg.jsonl
