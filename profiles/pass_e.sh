python bench.py --steps 10 --no-cpu-baseline > gpurun_out/e.json 2>gpurun_out/e.err; tail -2 gpurun_out/e.err
python -c "
import json; d=json.load(open('gpurun_out/e.json')); print(d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'soa', d['e2e_soa']['ms_per_step'])"
