#!/usr/bin/env python
"""ncu metrics CSV (gpu__time_duration, dram bytes per launch) + the sidecar of profiles/wg_pass.py -> per measure set and kernel
family: launches, summed duration, summed DRAM bytes.  Writes profiles/traffic.json (read by bench.py for roofline.traffic) and
prints a markdown table.
   python profiles/r2_traffic.py gpurun_out/R2b/wg_metrics.csv gpurun_out/R2b/wg_sidecar.json [profiles/traffic.json]"""
import csv, json, re, sys
from collections import OrderedDict

FAMILY = [("k_ingest", "k_ingest"), ("k_sites_count", "k_sites_count"), ("k_sites_emit", "k_sites_emit"), ("k_pdr_scatter", "k_pdr_scatter"),
          ("k_pdr_tile", "k_pdr_tile"), ("k_pdr_hazard", "k_pdr_gather"), ("k_pdr_gather", "k_pdr_gather"), ("k_pdr_rowcnt", "pdr_rows_count"),
          ("k_pdr_emit", "k_pdr_emit"), ("k_mhl_site", "k_mhl"), ("k_mhl", "k_mhl"), ("k_fdrp_tile", "k_fdrp*"), ("k_fdrp", "k_fdrp*"),
          ("k_quartet_scatter", "k_pm_scatter"), ("k_quartet_canon", "k_pm_count/emit"), ("k_quartet", "k_pm_count/emit"),
          ("k_site_emit", "k_site_emit"), ("k_scan", "rows_count(scan)"), ("k_expand", "k_expand"), ("k_lpmd_pairs", "k_lpmd_pairs")]


def kname(full):
    n = full[5:] if full.startswith("void ") else full
    n = n.replace("mth::", "").replace("<unnamed>::", "")
    return re.split(r"[<(]", n)[0].strip()


def main():
    rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
    side = json.load(open(sys.argv[2]))
    hdr = rows[0]
    iid, ik, im, iv = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    launches = OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(int(r[iid]), {"kernel": kname(r[ik])})
        d[r[im]] = float(r[iv].replace(",", ""))
    # k_fdrp_tables runs once per device (pair-index / division tables) and is not one of the engine's counted per-region launches
    seq = [l for l in launches.values() if l["kernel"] != "k_fdrp_tables"]
    total = sum(s["engine_kernel_launches"] for s in side["sets"])
    if total != len(seq):
        print(f"WARNING: sidecar counts {total} engine launches, ncu listed {len(seq)}", file=sys.stderr)
    out, at = {}, 0
    for s in side["sets"]:
        seg = seq[at:at + s["engine_kernel_launches"]]
        at += s["engine_kernel_launches"]
        agg = OrderedDict()
        for l in seg:
            a = agg.setdefault(l["kernel"], {"launches": 0, "time_ms": 0.0, "dram_bytes": 0.0})
            a["launches"] += 1
            a["time_ms"] += l.get("gpu__time_duration.sum", 0.0) / 1e6
            a["dram_bytes"] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
        # bench.py's kernel families (ProfScope names) -> ncu kernels
        fam = OrderedDict()
        for k, v in agg.items():
            name = k
            if k in ("k_mhl_site", "k_mhl"): name = "k_mhl"
            elif k in ("k_fdrp_tile", "k_fdrp"):
                m = s["measures"]
                name = "k_fdrp_qfdrp" if ("fdrp" in m and "qfdrp" in m) else ("k_qfdrp" if "qfdrp" in m else "k_fdrp")
            elif k == "k_quartet_scatter": name = "k_me_scatter" if s["measures"] == ["me"] else "k_pm_scatter"
            elif k in ("k_pdr_hazard", "k_pdr_gather"): name = "k_pdr_gather"
            f = fam.setdefault(name, {"launches": 0, "time_ms": 0.0, "dram_bytes": 0.0, "ncu_kernels": []})
            f["launches"] += v["launches"]; f["time_ms"] += v["time_ms"]; f["dram_bytes"] += v["dram_bytes"]; f["ncu_kernels"].append(k)
        for f in fam.values():
            f["passes"] = s["passes"]
            f["workload"] = f"{s['name']}: {s['reads']} reads, {s['calls']} calls, {s['sites']} sites, {s['passes']} pass(es)"
        where = "chr19" if s["name"].startswith("chr19:") else "wg"
        key = "+".join(s["measures"])
        if where == "chr19":
            out.setdefault("chr19", {}).update(fam)
        else:
            out.setdefault("wg", {})[key] = fam
        print(f"\n### {s['name']}  ({s['reads']} reads, {s['calls']} calls, {s['sites']} sites)\n")
        print("| kernel (ncu) | launches | time ms (cold, serialised) | DRAM GB | GB/s |")
        print("|---|---|---|---|---|")
        for k, v in agg.items():
            print(f"| {k} | {v['launches']} | {v['time_ms']:.3f} | {v['dram_bytes']/1e9:.3f} | {v['dram_bytes']/1e6/max(v['time_ms'],1e-9):.0f} |")
    if len(sys.argv) > 3:
        json.dump(out, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
