#!/bin/bash
# round 2: compute-sanitizer over the kernels that are new this round (GPU inflate / BAM decode, CpG-set filter, PM/ME row
# histograms, fused FDRP + qFDRP, sparse tile instances) and the smoke run
OUT=gpurun_out/${1:-r2san}; mkdir -p $OUT
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -3 $OUT/memcheck_smoke.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_bamdec.py tests/test_gpu_parity.py -m gpu -x -q -k "bamdec or inflate or device_decode or cpg_set or regions_with_host or dense_islands or default_flags" > $OUT/memcheck_tests.log 2>&1; echo "memcheck tests rc=$?"; tail -4 $OUT/memcheck_tests.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"; tail -3 $OUT/racecheck_smoke.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_bamdec.py -m gpu -x -q -k "synthetic_wgbs or inflate" > $OUT/racecheck_bamdec.log 2>&1; echo "racecheck bamdec rc=$?"; tail -4 $OUT/racecheck_bamdec.log
