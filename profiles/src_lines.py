#!/usr/bin/env python
"""Per-source-line shares of executed warp instructions and stall samples from an `ncu --page source --csv --print-source cuda,sass`
dump (one kernel).   python profiles/src_lines.py gpurun_out/R2b/src_k_fdrp_tile.csv [top] [file-filter]"""
import csv, sys, collections
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
flt = sys.argv[3] if len(sys.argv) > 3 else ""
inst, smp, text = collections.Counter(), collections.Counter(), {}
cur, fname, hdr = None, "", None
first_kernel_done = False
for r in csv.reader(open(path, errors="replace")):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; iE = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[0]:
        cur = (fname, int(r[0])); text[cur] = r[1].strip(); continue
    if cur:
        f = lambda x: int(x) if x.isdigit() else 0
        inst[cur] += f(r[iE]); smp[cur] += f(r[iS])
ti, ts = sum(inst.values()) or 1, sum(smp.values()) or 1
print(f"{path}: {ti} warp instructions, {ts} stall samples (all launches in the dump)")
for k, c in inst.most_common(top):
    if flt and flt not in k[0]:
        continue
    print(f"{100*c/ti:5.1f}% instr {100*smp[k]/ts:5.1f}% stall  {k[0]}:{k[1]:<4d} {text[k][:120]}")
