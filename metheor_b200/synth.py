"""Deterministic synthetic WGBS read generator producing the engine's SoA batch layout directly.

Not part of the hot path: it only feeds tests and bench.py with Bismark-like reads (sorted starts, both strands,
first-order-Markov methylation along each read, ~1 % no-calls, ~8 % low-mapq reads) of a requested coverage over a
synthetic contig whose CpG density mimics hg38 chr19 (1.89 CpG / 100 bp, CpG-island clustering).
"""
import numpy as np

MAPQ_HIGH = 42


def make_sites(seed, length, mean_gap=53.0, island_site_frac=0.10, island_seq_frac=0.01):
    """Sorted unique CpG positions over [0, length): a two-state (island / open sea) gap model."""
    rng = np.random.default_rng(seed)
    n_target = int(length / mean_gap)
    n_isl = int(n_target * island_site_frac)
    n_sea = n_target - n_isl
    # islands: 1 % of the sequence in segments of ~1 kb
    n_seg = max(1, int(length * island_seq_frac / 1000))
    seg_start = np.sort(rng.integers(0, max(1, length - 1000), n_seg))
    isl = seg_start[rng.integers(0, n_seg, n_isl)] + rng.integers(0, 1000, n_isl)
    sea = rng.integers(0, length - 1, n_sea)
    pos = np.unique(np.concatenate([isl, sea]).astype(np.int64))
    # a CpG occupies 2 bp: drop sites adjacent to the previous one
    keep = np.ones(len(pos), bool)
    keep[1:] = np.diff(pos) >= 2
    pos = pos[keep]
    return pos[pos < length - 1].astype(np.int32)


def make_reads(seed, sites, length, coverage, read_len=150, tid=0, nocall=0.01, lowq=0.08, rev=0.5, stay=0.85,
               del_frac=0.0, del_max=30, n_reads=None, drop_empty=False):
    """-> dict(tid, n_reads, n_cpg, start, end, meta, cpg_off, cpg_pos, cpg_rel, meth) numpy arrays (one contig).

    del_frac > 0 gives that fraction of reads one deletion of 1..del_max bp (reference span > read_len), which is
    what makes PDR's 150-bp flush slack matter."""
    rng = np.random.default_rng(seed)
    sites = np.asarray(sites, np.int64)
    R = int(n_reads if n_reads is not None else coverage * length / read_len)
    start = np.sort(rng.integers(0, max(1, length - read_len - del_max - 1), R)).astype(np.int64)
    is_rev = rng.random(R) < rev
    dlen = np.where(rng.random(R) < del_frac, rng.integers(1, del_max + 1, R), 0).astype(np.int64)
    dq = rng.integers(1, read_len - 1, R).astype(np.int64)  # query offset where the deletion sits
    span = read_len + dlen
    end = start + span - 1
    mapq = np.where(rng.random(R) < lowq, rng.integers(0, 10, R), MAPQ_HIGH).astype(np.uint32)

    # candidate sites: forward reads call site x at abspos x, reverse reads at abspos x+1 (pos = abspos-1 = x)
    shift = is_rev.astype(np.int64)
    lo = np.searchsorted(sites, start - shift, "left")
    hi = np.searchsorted(sites, end - shift, "right")
    cnt = (hi - lo).astype(np.int64)
    tot = int(cnt.sum())
    ridx = np.repeat(np.arange(R, dtype=np.int64), cnt)
    first = np.cumsum(cnt) - cnt
    sidx = lo[ridx] + (np.arange(tot, dtype=np.int64) - first[ridx])
    x = sites[sidx]
    abspos = x + shift[ridx]
    off_in_ref = abspos - start[ridx]
    in_del = (dlen[ridx] > 0) & (off_in_ref >= dq[ridx]) & (off_in_ref < dq[ridx] + dlen[ridx])
    called = (~in_del) & (rng.random(tot) >= nocall)
    rel = off_in_ref - np.where(off_in_ref >= dq[ridx] + dlen[ridx], dlen[ridx], 0)

    # methylation: per-site beta, first-order Markov along the read
    srng = np.random.default_rng(seed ^ 0x5EED)
    mix = srng.random(len(sites)) < 0.6
    beta = np.where(mix, srng.beta(8, 1.5, len(sites)), srng.beta(1.2, 6, len(sites)))

    ridx, x, rel, sidx = ridx[called], x[called], rel[called], sidx[called]
    n = len(ridx)
    cnt2 = np.bincount(ridx, minlength=R).astype(np.int64)
    off = np.zeros(R + 1, np.int64)
    np.cumsum(cnt2, out=off[1:])
    fresh_draw = rng.random(n) < beta[sidx]
    is_first = np.zeros(n, bool)
    is_first[off[:-1][cnt2 > 0]] = True
    fresh = is_first | (rng.random(n) >= stay)
    src = np.maximum.accumulate(np.where(fresh, np.arange(n), 0))
    state = fresh_draw[src]

    # pack methylation bits: bit k of the read's word(s) = k-th CpG of the read
    k_in_read = np.arange(n, dtype=np.int64) - off[ridx]
    words_per_read = np.maximum(1, (cnt2 + 63) // 64)
    moff = np.zeros(R + 1, np.int64)
    np.cumsum(words_per_read, out=moff[1:])
    meth = np.zeros(int(moff[-1]), np.uint64)
    np.bitwise_or.at(meth, moff[ridx] + k_in_read // 64, state.astype(np.uint64) << (k_in_read % 64).astype(np.uint64))

    b = dict(tid=int(tid), n_reads=R, n_cpg=n, start=start.astype(np.int32), end=end.astype(np.int32),
             meta=(mapq | ((~is_rev).astype(np.uint32) << 8)).astype(np.uint32), cpg_off=off.astype(np.uint32),
             cpg_pos=x.astype(np.int32), cpg_rel=rel.astype(np.uint16), meth=meth,
             meth_off=None if int(words_per_read.max(initial=1)) == 1 else moff.astype(np.uint32))
    if drop_empty:
        from .batch import select_reads
        b = select_reads(b, cnt2 > 0)
    return b


def chr19_like(seed=20260101, coverage=30.0, length=58_617_616, **kw):
    """BASELINE.json configs[1]: synthetic 30x WGBS over a chr19-sized contig (58.6 Mb, ~1.1 M CpG sites)."""
    sites = make_sites(seed, length)
    return make_reads(seed + 1, sites, length, coverage, **kw), sites
