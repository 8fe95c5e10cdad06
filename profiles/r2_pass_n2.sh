#!/bin/bash
# round 2: multi-GPU checks on N GPUs of one box: CLI --gpus N (bins / contigs) vs the oracle, bench.py strong scaling under torchrun
N=${2:-2}
O=gpurun_out/${1:-r2n2}; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 600 python -m pytest tests/test_cli_gpu.py -m gpu -x -q -k "multi_gpu" > $O/pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "rc=$?" >> $O/bench_n$N.err
tail -3 $O/pytest_multi.log; tail -c 1200 $O/bench_n$N.err; head -c 600 $O/bench_n$N.json
