#!/bin/bash
# round 2, pass c: GPU tests + whole-genome bench after the PM/ME row-histogram redesign and the sparse-density tile instances
set -x
O=gpurun_out/${1:-r2c}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 1200 python bench.py --steps 10 --no-extra > $O/bench_full.json 2> $O/bench_full.err; echo "rc=$?" >> $O/bench_full.err
tail -c 600 $O/pytest.log; tail -c 1500 $O/bench_full.err
