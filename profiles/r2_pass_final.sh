#!/bin/bash
# round 2, final state: the whole GPU test-suite, the driver's two bench command lines, smoke(), ncu of k_tag
O=gpurun_out/${1:-r2final}; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -4 $O/pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -3 $O/smoke.log
timeout 1800 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
tail -c 600 $O/bench.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "rc=$?" >> $O/bench_reference.err
tail -c 300 $O/bench_reference.err; head -c 400 $O/bench_reference.json
bash profiles/r2_ncu_tag.sh ${1:-r2final}
