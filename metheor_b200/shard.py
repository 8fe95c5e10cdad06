"""Genomic sharding of the hot path across GPUs — host-side planning only, no arithmetic.

Two partitionings, both without a data-path collective:
  * contigs  : whole contigs per rank, longest first onto the least loaded rank (what `metheor --gpus N` does in C++).
    A tid change flushes every accumulator in the reference (readutil.rs:290-295), so contigs are independent.
  * bins     : contiguous position ranges of the linearised genome, balanced by length or by read count.  A rank
    receives its own reads (start in [lo, hi)) plus a HALO: reads starting up to `halo` bases before lo (they can call
    sites >= lo, or be flush triggers for them) and reads starting at hi or hi+... that start <= hi (a reverse-strand
    read starting at p+1 calls site p).  Sites (and quartets / pair rows, by first position) are OWNED by the bin that
    contains them; every contributor and every flush trigger of an owned site lies inside the rank's read range, so the
    reference's segment semantics stay exact.  Halo reads carry META_HALO so that LPMD's global counters count each
    read exactly once (by its owner); rows of non-owned sites are dropped with `owned_rows`.
The only exchange is the all-reduce (sum) of LPMD's four int64 counters.
"""
import numpy as np

from .batch import select_reads

META_HALO = 1 << 9   # include/metheor_b200.h: MTH_META_HALO
MAX_REF_SPAN = 65024  # engine limit on end - start + 1: a safe halo for any accepted input


def plan_contigs(ref_len, world):
    """-> list (per rank) of tid lists."""
    order = sorted(range(len(ref_len)), key=lambda t: -ref_len[t])
    load = [0] * world
    out = [[] for _ in range(world)]
    for t in order:
        g = min(range(world), key=lambda r: load[r])
        out[g].append(t)
        load[g] += ref_len[t]
    return [sorted(x) for x in out]


def plan_bins(ref_len, world, weights=None, region_cost=0):
    """Split the linearised genome into `world` contiguous ranges of equal length, or of equal read count when
    `weights` is a per-contig list of sorted read starts.  -> list (per rank) of (tid, lo, hi) intervals, hi exclusive.
    region_cost (bases, length mode): every contig a rank touches is a region of its own on the GPU and a region has a fixed
    cost (launches of ~25 small kernels, two host waits) next to the cost per base; the ranges are balanced on
    bases + region_cost x contigs by giving every contig `region_cost` bases of padding in front (a cut that falls into the
    padding moves to the contig's first base).  0 = balance on length alone."""
    n_ref = len(ref_len)
    cuts = [(0, 0)]  # rank r covers [cuts[r], cuts[r+1]) in (tid, pos) order
    if weights is None:
        P = int(region_cost)
        base = [P]  # first base of contig t in the padded coordinate
        for l in ref_len:
            base.append(base[-1] + int(l) + P)
        total = base[-1] - P
        for r in range(1, world):  # integer arithmetic: host/run.cpp plan_bins computes the same cuts
            x = total * r // world
            tid = min(int(np.searchsorted(np.asarray(base, np.int64) - P, x, "right")) - 1, n_ref - 1)
            cuts.append((tid, max(0, x - base[tid])))
    else:
        counts = np.array([len(w) for w in weights], np.int64)
        base = np.concatenate([[0], np.cumsum(counts)])
        for r in range(1, world):
            k = int(round(base[-1] * r / world))  # global index of the first read of rank r
            tid = min(int(np.searchsorted(base, k, "right")) - 1, n_ref - 1)
            kk = k - int(base[tid])
            cuts.append((tid, int(weights[tid][kk]) if kk < counts[tid] else int(ref_len[tid])))
    cuts.append((n_ref - 1, int(ref_len[-1])))
    for r in range(1, len(cuts)):  # keep the cut list monotone
        if cuts[r] < cuts[r - 1]:
            cuts[r] = cuts[r - 1]
    out = []
    for r in range(world):
        (t0, p0), (t1, p1) = cuts[r], cuts[r + 1]
        iv = []
        for tid in range(t0, t1 + 1):
            lo = p0 if tid == t0 else 0
            hi = p1 if tid == t1 else int(ref_len[tid])
            if hi > lo:
                iv.append((tid, lo, hi))
        out.append(iv)
    return out


def select_shard(b, intervals, halo=MAX_REF_SPAN):
    """Reads of batch `b` (one contig) a rank needs for its intervals -> (sub-batch with META_HALO set on halo reads,
    number of owned reads).  None if the rank has nothing on this contig."""
    mine = [(lo, hi) for tid, lo, hi in intervals if tid == b["tid"]]
    if not mine:
        return None, 0
    start = np.asarray(b["start"], np.int64)
    need = np.zeros(b["n_reads"], bool)
    own = np.zeros(b["n_reads"], bool)
    for lo, hi in mine:
        need |= (start >= lo - halo) & (start <= hi)
        own |= (start >= lo) & (start < hi)
    sub = select_reads(b, need)
    sub["meta"] = np.where(own[need], sub["meta"], sub["meta"] | META_HALO).astype(np.uint32)
    return sub, int(own.sum())


def owned_rows(rows, intervals, pos_key="pos"):
    """Keep the rows whose (first) position lies in one of the rank's intervals."""
    tid, pos = np.asarray(rows["tid"]), np.asarray(rows[pos_key])
    keep = np.zeros(len(tid), bool)
    for t, lo, hi in intervals:
        keep |= (tid == t) & (pos >= (lo if lo > 0 else -1)) & (pos < hi)  # a reverse read at 0 calls site -1 (readutil.rs:338)
    out = {}
    for k, v in rows.items():
        out[k] = v[keep] if isinstance(v, np.ndarray) and len(v) == len(keep) else v
    if "n" in out:
        out["n"] = int(keep.sum())
    return out


def merge_rows(parts, keys=("tid", "pos")):
    """Concatenate per-rank row dicts and restore the global (tid, pos, ...) order."""
    cols = [k for k in parts[0] if isinstance(parts[0][k], np.ndarray)]
    cat = {k: np.concatenate([p[k] for p in parts]) for k in cols}
    order = np.lexsort(tuple(cat[k] for k in reversed(keys)))
    out = {k: v[order] for k, v in cat.items()}
    out["n"] = len(order)
    return out
