#!/bin/bash
# round 2, pass a: GPU tests + the new whole-genome bench at a small scale, then at full size
set -x
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/gpu.txt; nproc >> $O/gpu.txt; free -g >> $O/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python bench.py --scale 0.02 --steps 3 --measure-steps 2 --wg60 0 > $O/bench_small.json 2> $O/bench_small.err; echo "rc=$?" >> $O/bench_small.err
timeout 1500 python bench.py --steps 10 > $O/bench_full.json 2> $O/bench_full.err; echo "rc=$?" >> $O/bench_full.err
tail -c 600 $O/pytest.log; tail -c 1500 $O/bench_small.err; tail -c 1500 $O/bench_full.err
