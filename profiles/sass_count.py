#!/usr/bin/env python
"""Static SASS size of a k_ingest instance by phase (segments between BAR.SYNCs), for instruction-count work without a GPU.
   python profiles/sass_count.py [object file] [mangled-name substring]"""
import subprocess, sys
obj = sys.argv[1] if len(sys.argv) > 1 else "metheor_b200/csrc/build/k_ingest.o"
want = sys.argv[2] if len(sys.argv) > 2 else "k_ingestILb0ELb1"
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
seg, cur, on = [], 0, False
for l in out.splitlines():
    if "Function :" in l:
        if on:
            seg.append(cur)
        on = want in l
        cur = 0
        if on:
            print(l.strip())
        continue
    if not on or "/*" not in l or l.strip().startswith("/* 0x"):
        continue
    body = l.split("*/", 1)[1] if "*/" in l else ""
    ins = body.strip().split(";")[0].strip()
    if not ins:
        continue
    cur += 1
    if "BAR.SYNC" in ins:
        seg.append(cur)
        cur = 0
if on:
    seg.append(cur)
print("instructions per segment:", seg, "total", sum(seg))
