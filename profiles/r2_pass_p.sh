#!/bin/bash
# round 2, pass p: register-cached mixed-site quartets + side stream, 1024-thread scan_sums, funnel-shift bit reader, one-wave windows
O=gpurun_out/${1:-r2p}; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -4 $O/pytest.log
timeout 900 python - > $O/bam_leg.json 2> $O/bam_leg.err <<'PY'
import json, sys, time, tempfile, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import bench_chr19 as X
from metheor_b200 import bamdec, batch as B, synth_bam
b, _ = X.make_workload(0, 30.0, X.CONTIG_LEN)
with tempfile.TemporaryDirectory() as d:
    bam = os.path.join(d, "s.bam")
    sub = B.slice_reads(b, 0, 2_000_000)
    synth_bam.write_bam(bam, [("chr19", X.CONTIG_LEN)], [sub], threads=16)
    data = open(bam, "rb").read()
    mem = bamdec.bgzf_members(data)
    for rep in range(3):
        out, status, ms = bamdec.inflate_members(data, mem)
        print("inflate kernel: %d members, %d -> %d bytes, %.2f ms, %.1f GB/s out, bad=%d" % (len(mem), len(data), len(out), ms, len(out) / ms / 1e6, int(status.any())), file=sys.stderr)
r = X.bam_leg(b, 2_000_000, X.CONTIG_LEN)
print(json.dumps(r))
for m in ("pdr", "lpmd"):
    print(m, r[m]["reads_per_sec"], r[m]["device_decode"], r[m]["stage_seconds"], "host path", r[m]["host_decode_path"]["reads_per_sec"], r[m]["tsv_identical_to_oracle"], file=sys.stderr)
PY
tail -c 2500 $O/bam_leg.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-extra > $O/bench_quick.json 2> $O/bench_quick.err; echo "rc=$?" >> $O/bench_quick.err
tail -c 300 $O/bench_quick.err
python - <<PY
import json
d=json.loads(open("$O/bench_quick.json").readline())
print("headline", d["ms_per_step"], d["value"])
for k,v in list(d["measures"].items())+list(d["combined"].items()): print(k, v["ms_per_step"])
print(d["kernels"])
PY
