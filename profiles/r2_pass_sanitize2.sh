#!/bin/bash
# round 2 (late changes): compute-sanitizer over the grouped k_tag path (word loads / 8-byte stores), the register-cached
# mixed-site quartets + side stream, the 1024-thread block-sum scan, the per-site island fallback of k_mhl_site
OUT=gpurun_out/${1:-r2san2}; mkdir -p $OUT
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_tag.py -m gpu -x -q -k "random_reads or golden" > $OUT/memcheck_tag.log 2>&1; echo "memcheck tag rc=$?"; tail -4 $OUT/memcheck_tag.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mixed_quartet or dense_islands or long_spans or default_flags or fixture" > $OUT/memcheck_parity.log 2>&1; echo "memcheck parity rc=$?"; tail -4 $OUT/memcheck_parity.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_tag.py -m gpu -x -q -k "random_reads and 1-False" > $OUT/racecheck_tag.log 2>&1; echo "racecheck tag rc=$?"; tail -4 $OUT/racecheck_tag.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mixed_quartet or dense_islands" > $OUT/racecheck_parity.log 2>&1; echo "racecheck parity rc=$?"; tail -4 $OUT/racecheck_parity.log
