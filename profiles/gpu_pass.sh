#!/bin/bash
# One GPU pass (run under gpurun): parity tests, bench line, ncu launch list, ncu full captures of the hot kernels.
# Usage: gpurun --timeout 1800 -- 'bash profiles/gpu_pass.sh r01a'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt; lscpu | grep 'Model name' >> $OUT/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json | head -c 3000
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_ingest|k_pdr_scatter|k_sites_emit|k_pdr_emit' -s 12 -c 4 \
    -o $OUT/prof_pdr_lpmd python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls -la $OUT
timeout 600 python profiles/measure_all.py > $OUT/measure_all.jsonl 2>> $OUT/bench.err
# `tag` (XM synthesis): CLI profile, launch time and one full capture of k_tag
timeout 300 python profiles/tag_cli.py 1000000 2>> $OUT/bench.err | tail -1 > $OUT/tag_cli.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_tag|k_upper' -c 8 --csv --log-file $OUT/tag_launches.csv \
    python -c "import bench; bench.tag_leg(2000000, 58617616)" > $OUT/tag_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_tag' -s 1 -c 1 -o $OUT/prof_tag \
    python -c "import bench; bench.tag_leg(2000000, 58617616)" >> $OUT/tag_under_ncu.log 2>&1
