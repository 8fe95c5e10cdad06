#!/usr/bin/env python
"""Turn one GPU pass (gpurun_out/<tag>/, written by profiles/gpu_pass.sh) into the tracked summary profiles/<tag>_*.
   python profiles/summarize.py r01a          (needs `ncu` on PATH for the .ncu-rep -> csv step; no GPU needed)"""
import csv, re
import glob
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic"]


def kname(full_name):
    """'void mth::k_ingest<(bool)0, (bool)1>(mth::IngestArgs)' -> 'k_ingest' (instances of a template are one kernel here)."""
    n = full_name[5:] if full_name.startswith("void ") else full_name
    for pre in ("mth::", "<unnamed>::"):
        n = n.replace(pre, "")
    return re.split(r"[<(]", n)[0].strip()


def launches(path):
    """-> OrderedDict kernel -> [n, total_ns] from an `ncu --metrics gpu__time_duration.sum --csv` log."""
    agg = OrderedDict()
    with open(path) as f:
        rows = [r for r in csv.reader(l for l in f if l.startswith('"'))]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    for r in rows[1:]:
        name = kname(r[ik])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", ""))
    return agg


def full(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in raw.splitlines() if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": kname(r[hdr.index("Kernel Name")])}
        for k in KEYS:
            if k in hdr:
                d[k] = f"{r[hdr.index(k)]} {units[hdr.index(k)]}".strip()
        out.append(d)
    return out


def top_stalls(rep, kernel, n=12):
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in src.splitlines() if l.startswith('"')))
    if len(rows) < 3:
        return []
    hdr = rows[1]
    iS, isrc = hdr.index("# Samples"), hdr.index("Source")
    stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    body = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break  # next launch of the same kernel
        if len(r) == len(hdr) and r[0] != "Address":
            body.append(r)
    tot = sum(int(r[iS] or 0) for r in body) or 1
    out = []
    for r in sorted(body, key=lambda r: -int(r[iS] or 0))[:n]:
        st = sorted(((int(r[i] or 0), hdr[i]) for i in stalls), reverse=True)[0]
        out.append((100.0 * int(r[iS] or 0) / tot, r[isrc].strip(), st[1]))
    return out


def main():
    tag = sys.argv[1]
    d = os.path.join(ROOT, "gpurun_out", tag)
    md = [f"# GPU pass {tag}", ""]
    for name in ("gpu.txt", "nproc.txt", "pytest_gpu.log"):
        p = os.path.join(d, name)
        if os.path.exists(p):
            md += [f"## {name}", "```", open(p).read().strip()[-1500:], "```", ""]
    for name in sorted(glob.glob(os.path.join(d, "bench*.json"))):
        txt = open(name).read().strip()
        if txt:
            md += [f"## {os.path.basename(name)} (not under a profiler)", "```json", txt, "```", ""]
            with open(os.path.join(ROOT, "profiles", f"{tag}_{os.path.basename(name)}"), "w") as f:
                f.write(txt + "\n")
    for lc in sorted(glob.glob(os.path.join(d, "launches*.csv"))):
        agg = launches(lc)
        tot = sum(a[1] for a in agg.values()) or 1
        md += [f"## ncu launch list: {os.path.basename(lc)} (`--metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)", "",
               "| kernel | launches | total us | share | avg us |", "|---|---|---|---|---|"]
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            md.append(f"| {k} | {a[0]} | {a[1] / 1e3:.1f} | {100 * a[1] / tot:.1f}% | {a[1] / a[0] / 1e3:.2f} |")
        md.append("")
    for rep in sorted(glob.glob(os.path.join(d, "*.ncu-rep"))):
        md += [f"## ncu --set full: {os.path.basename(rep)}", ""]
        seen = set()
        for k in full(rep):
            md.append(f"### {k['kernel']}")
            md += [f"- {m}: {v}" for m, v in k.items() if m != "kernel"]
            if k["kernel"] not in seen:
                seen.add(k["kernel"])
                st = top_stalls(rep, k["kernel"])
                if st:
                    md += ["", "top stall sites (share of warp-stall samples, SASS, dominant reason):", "```"]
                    md += [f"{p:5.1f}%  {s[:100]:100s} {r}" for p, s, r in st] + ["```"]
            md.append("")
    # per-launch DRAM traffic of the captured kernels: bench.py quotes it as roofline.traffic
    traffic = {}
    for rep in sorted(glob.glob(os.path.join(d, "*.ncu-rep"))):
        for k in full(rep):
            def num(key):
                v, u = (k.get(key, "0 byte").split() + ["byte"])[:2]
                return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            traffic.setdefault(k["kernel"], {"dram_bytes": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
                                             "dram_read_bytes": num("dram__bytes_read.sum"), "dram_write_bytes": num("dram__bytes_write.sum"),
                                             "source": f"profiles/{tag}_summary.md (ncu --set full, one launch)"})
    if traffic:
        with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
            json.dump(traffic, f, indent=1)
    out = os.path.join(ROOT, "profiles", f"{tag}_summary.md")
    with open(out, "w") as f:
        f.write("\n".join(md) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
