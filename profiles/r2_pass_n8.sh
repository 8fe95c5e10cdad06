#!/bin/bash
# strong scaling of the whole-genome bench on N GPUs of one box (torchrun), final round-2 state
N=${2:-8}
O=gpurun_out/${1:-r2n8b}; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "rc=$?" >> $O/bench_n$N.err
tail -c 600 $O/bench_n$N.err; head -c 300 $O/bench_n$N.json
