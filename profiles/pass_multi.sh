OUT=gpurun_out/${1:-r01u}; mkdir -p $OUT
N=${2:-2}
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "rc=$?"; tail -3 $OUT/bench_n$N.err
python -c "
import json; d=json.load(open('$OUT/bench_n$N.json')); print(d['n_gpus'], d['ms_per_step'], d['value']/1e9, 'e2e', d['e2e']['ms_per_step'], d['e2e']['value']/1e9, d['lpmd'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $OUT/ref_n$N.json 2>> $OUT/bench_n$N.err; echo "ref rc=$?"; head -c 300 $OUT/ref_n$N.json
# the CLI sharding contigs over both GPUs
python - <<PY
import sys, os, subprocess
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from metheor_b200 import synth, synth_bam, host
import oracle_lib
oracle_lib.build()
refs=[("chrA", 400000), ("chrB", 250000), ("chrC", 300000)]
bs=[]
for t,(n,l) in enumerate(refs):
    s=synth.make_sites(700+t,l); bs.append(synth.make_reads(710+t,s,l,25.0,tid=t))
synth_bam.write_bam('/tmp/m.bam', refs, bs)
for m in ("pdr","lpmd","mhl","pm","fdrp"):
    r=host.cli(m,"-i","/tmp/m.bam","-o",f"/tmp/m_{m}.tsv","--gpus",$N)
    o=subprocess.run([oracle_lib.CLI_PATH,m,"-i","/tmp/m.bam","-o",f"/tmp/o_{m}.tsv"])
    print(m, r.returncode, r.stderr[:200], open(f"/tmp/m_{m}.tsv").read()==open(f"/tmp/o_{m}.tsv").read())
PY
