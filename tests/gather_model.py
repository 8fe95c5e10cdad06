"""Pure-Python model of the engine's per-site GATHER formulation, used to check the algorithm itself
against the streaming oracle before/independently of the CUDA code: for a CpG site p, walk the reads whose start lies
in [p - Lmax + 1, p + 1] in file order, split them into segments at every flush trigger, and report the last segment
whose depth reaches min_depth.  Small inputs only (Python loops)."""
import numpy as np


def _reads(soa):
    off = soa["cpg_off"]
    out = []
    for i in range(len(soa["start"])):
        a, b = int(off[i]), int(off[i + 1])
        out.append(dict(start=int(soa["start"][i]), end=int(soa["end"][i]), mapq=int(soa["mapq"][i]),
                        pos=[int(x) for x in soa["cpg_pos"][a:b]], meth=[int(x) for x in soa["cpg_meth"][a:b]]))
    return out


def site_segments(soa, contrib_ok, trigger_ok, slack):
    """-> {p: [segment, ...]} where a segment is the list of contributing read indices (file order)."""
    reads = _reads(soa)
    starts = np.asarray(soa["start"])
    lmax = max((r["end"] - r["start"] + 1 for r in reads), default=0)
    sites = sorted({p for r in reads for p in r["pos"]})
    res = {}
    for p in sites:
        lo = int(np.searchsorted(starts, p - lmax + 1, "left"))
        hi = int(np.searchsorted(starts, p + 1, "right"))
        segs, cur = [], []
        for j in range(lo, hi):
            r = reads[j]
            if not r["pos"]:
                continue
            if p in r["pos"]:
                if contrib_ok(r):
                    cur.append(j)
            elif trigger_ok(r) and r["pos"][0] > p + slack:
                if cur:
                    segs.append(cur)
                    cur = []
        if cur:
            segs.append(cur)
        res[p] = segs
    return res, reads


def pdr(soa, min_depth, min_cpgs, min_qual):
    ok = lambda r: len(r["pos"]) >= min_cpgs and r["mapq"] >= min_qual and len(r["pos"]) > 0
    segs, reads = site_segments(soa, ok, ok, 150)
    rows = []
    for p, ss in segs.items():
        best = None
        for s in ss:
            if len(s) >= min_depth:
                d = sum(1 for j in s if len(set(reads[j]["meth"])) > 1)
                best = (p, len(s) - d, d)
        if best:
            rows.append(best)
    return rows


def site_sets_strict(soa, min_depth, contrib_ok, trigger_ok):
    segs, reads = site_segments(soa, contrib_ok, trigger_ok, 0)
    out = {}
    for p, ss in segs.items():
        best = None
        for s in ss:
            if len(s) >= min_depth:
                best = s
        if best:
            out[p] = best
    return out, reads
