#!/bin/bash
# round 2: ncu evidence for every kernel family on the whole-genome bench workload (1 GPU).
#  (1) launch list + DRAM bytes of every engine kernel over full passes of every measure set  -> traffic.json, per-kernel shares
#  (2) --set full capture (source-level) of one all-seven pass + the single-measure FDRP / qFDRP / PM passes on ONE contig (chr1)
set -x
O=gpurun_out/${1:-R2b}; mkdir -p $O
SETS="pm+me,pdr,lpmd,mhl,pm,me,fdrp,qfdrp,pdr+lpmd,fdrp+qfdrp,all7,chr19:pdr+lpmd"
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:'k_' --csv --log-file $O/wg_metrics.csv \
    python profiles/wg_pass.py --sets $SETS --sidecar $O/wg_sidecar.json > $O/wg_pass.log 2>&1
echo "metrics rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_' -o $O/prof_chr1 -f \
    python profiles/wg_pass.py --contigs 0 --sets all7,fdrp,qfdrp --sidecar $O/chr1_sidecar.json > $O/chr1_pass.log 2>&1
echo "full rc=$?"
ls -la $O
ncu -i $O/prof_chr1.ncu-rep --page raw --csv > $O/prof_chr1_raw.csv 2>/dev/null
SZ=$(stat -c %s $O/prof_chr1.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -gt 45000000 ]; then
  for k in k_ingest k_fdrp_tile k_mhl_site k_quartet_scatter k_quartet_hist k_quartet_canon_emit k_pdr_scatter k_mhl k_fdrp k_quartet; do
    ncu -i $O/prof_chr1.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:"$k\b" > $O/src_$k.csv 2>/dev/null
  done
  gzip -9 $O/src_*.csv
  rm -f $O/prof_chr1.ncu-rep
fi
gzip -9 $O/wg_metrics.csv
ls -la $O
