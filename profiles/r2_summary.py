#!/usr/bin/env python
"""`ncu --set full` raw page (csv) of profiles/r2_ncu.sh pass (2) -> one markdown table row per distinct kernel (first launch of each):
duration, DRAM bytes and throughput, SM throughput, issue-active, achieved occupancy, registers, executed warp instructions.
   python profiles/r2_summary.py gpurun_out/R2c/prof_chr1_raw.csv gpurun_out/R2c/chr1_sidecar.json"""
import csv, json, re, sys
rows = list(csv.reader(open(sys.argv[1])))
side = json.load(open(sys.argv[2])) if len(sys.argv) > 2 else None
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
def kname(full):
    n = full[5:] if full.startswith("void ") else full
    n = n.replace("mth::", "").replace("<unnamed>::", "")
    m = re.match(r"([A-Za-z0-9_]+)(<[^>]*>)?", n)
    return (m.group(1) + (m.group(2) or "")) if m else n
K = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue-active %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
     ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp instr"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
if side:
    s0 = side["sets"][0]
    print(f"workload: contig(s) {side['contigs']} of the bench genome, {side['reads']} reads, {side['calls']} calls; first set: {s0['name']} ({s0['sites']} sites)\n")
print("| kernel | " + " | ".join(n for _, n in K) + " |")
print("|---|" + "---|" * len(K))
seen = set()
for r in rows[2:]:
    name = kname(r[col["Kernel Name"]])
    if name in seen or name.startswith("k_scan"):
        continue
    seen.add(name)
    cells = []
    for k, _ in K:
        if k not in col:
            cells.append("-"); continue
        v, u = r[col[k]], units[col[k]]
        try:
            f = float(v.replace(",", ""))
            v = f"{f:.3g}" if abs(f) < 1000 else f"{f:,.0f}"
        except ValueError:
            pass
        cells.append(f"{v} {u}".strip())
    print(f"| {name} | " + " | ".join(cells) + " |")
