#!/bin/bash
# round 2: the driver's command lines (bench.py for both arms) + the GPU test-suite
O=gpurun_out/${1:-r2full}; mkdir -p $O
true
true
timeout 1800 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
tail -c 1200 $O/bench.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "rc=$?" >> $O/bench_reference.err
tail -c 300 $O/bench_reference.err; head -c 400 $O/bench_reference.json
