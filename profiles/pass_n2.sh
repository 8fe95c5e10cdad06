#!/bin/bash
# 2-GPU weak-scaling bench only (torchrun + NCCL).   gpurun --gpus 2 --timeout 150 -- 'bash profiles/pass_n2.sh r05m'
OUT=gpurun_out/${1:-r05m}; mkdir -p $OUT
timeout 140 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --tag-reads 0 --bam-reads 0 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "rc=$?"; tail -2 $OUT/bench_n2.err
python -c "
import json; d=json.load(open('$OUT/bench_n2.json')); print(d['n_gpus'], d['ms_per_step'], d['value']/1e9, 'e2e', d['e2e']['ms_per_step'], d['e2e']['value']/1e9, d['lpmd'], d.get('lpmd_all_ranks'))"
