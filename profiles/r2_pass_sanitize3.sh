#!/bin/bash
# compute-sanitizer over the 16- / 8-site FDRP tile instances (deep piles: reservoir sampling in the tile kernel)
OUT=gpurun_out/${1:-r2san3}; mkdir -p $OUT
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "deep_piles" > $OUT/memcheck_fdrp.log 2>&1; echo "memcheck rc=$?"; tail -3 $OUT/memcheck_fdrp.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "deep_piles" > $OUT/racecheck_fdrp.log 2>&1; echo "racecheck rc=$?"; tail -3 $OUT/racecheck_fdrp.log
