#!/usr/bin/env python
"""Per-CUDA-source-line instruction and stall-sample shares of one kernel in an .ncu-rep (needs -lineinfo + --import-source).
   python profiles/lines.py gpurun_out/r01e/prof.ncu-rep k_ingest [top]"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern],
                     capture_output=True, text=True).stdout
inst, smp, text = collections.Counter(), collections.Counter(), {}
cur, fname, hdr, seen_kernel = None, "", None, 0
for r in csv.reader(out.splitlines()):
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
print(f"{kern}: {ti} warp instructions, {ts} stall samples")
for k, c in inst.most_common(top):
    print(f"{100*c/ti:5.1f}% instr {100*smp[k]/ts:5.1f}% stall  {k[0]}:{k[1]:<4d} {text[k][:110]}")
