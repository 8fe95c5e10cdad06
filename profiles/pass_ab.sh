#!/bin/bash
# k_ingest A/B pass (run under gpurun): parity tests on the shipped build, then every variant_*.so against it, then one
# full ncu capture of the shipped k_ingest.   gpurun --timeout 900 -- 'bash profiles/pass_ab.sh ab1'
TAG=${1:-ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 300 python profiles/ingest_ab.py --park
timeout 900 python profiles/ingest_ab.py default metheor_b200/csrc/variant_*.so > $OUT/ab.jsonl 2> $OUT/ab.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_ingest' -s 4 -c 1 -o $OUT/prof_ingest \
    python profiles/ingest_ab.py --child default > $OUT/ncu.log 2>&1
rm -f /dev/shm/mth_ab_workload.npz
python -c "
import json,sys
for l in open('$OUT/ab.jsonl'):
    d=json.loads(l)
    print(d['lib'])
    for r in d['runs']:
        print('   ', '+'.join(r['measures']), r['ms_per_pass'], r['digest'], 'k_ingest', r['kernels_ms'].get('k_ingest'))
"
tail -3 $OUT/ab.err
