"""Whole-genome synthetic WGBS generator (BASELINE.json configs[2..4]) written with torch tensor ops so
that a 30x / 60x / 100x human-sized read set (0.6 - 2 G reads) is produced directly in HBM in seconds.

Not part of the hot path: it only feeds bench.py and the tests.  Same read model as `synth.make_reads` (sorted uniform
starts, both strands, first-order Markov methylation along a read, 1 % no-calls, 8 % low-mapq reads) over 24 contigs with
the hg38 chromosome lengths and ~28 M CpG sites (10 % of them in CpG islands covering 1 % of the sequence).

Every random decision is an INTEGER hash of (seed, stream, index) — no floating point, no device RNG — so the same call
on device="cpu" and device="cuda" yields bit-identical reads: the CPU oracle legs of bench.py (`--impl reference`, which
must not need the GPU) regenerate exactly the contig the engine saw.
"""
import numpy as np
import torch

HG38 = (("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555), ("chr5", 181538259),
        ("chr6", 170805979), ("chr7", 159345973), ("chr8", 145138636), ("chr9", 138394717), ("chr10", 133797422),
        ("chr11", 135086622), ("chr12", 133275309), ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189),
        ("chr16", 90338345), ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616), ("chr20", 64444167),
        ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415))
GENOME_LEN = sum(l for _, l in HG38)          # 3 088 269 832
WG_SITES = 28_200_000
MEAN_GAP = GENOME_LEN / WG_SITES               # ~109.5 bp
MAPQ_HIGH = 42
READ_LEN = 150

_M64 = (1 << 64) - 1


def _s64(x):
    """python int -> the int64 with the same low 64 bits"""
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


_C1, _C2, _C3 = _s64(0x9E3779B97F4A7C15), _s64(0xBF58476D1CE4E5B9), _s64(0x94D049BB133111EB)


def _lsr(x, k):
    """logical shift right of an int64 tensor (torch's >> is arithmetic)"""
    return (x >> k) & ((1 << (64 - k)) - 1)


def hash63(seed, stream, idx):
    """splitmix64 finaliser of (seed, stream, idx) -> int64 tensor with 63 uniform bits (non-negative).  int64
    multiplication wraps identically on CPU and CUDA."""
    key = _s64((seed * 0x9E3779B97F4A7C15) ^ ((stream + 1) * 0xD1B54A32D192ED03))
    x = idx * _C1 + key
    x = (x ^ _lsr(x, 30)) * _C2
    x = (x ^ _lsr(x, 27)) * _C3
    x = x ^ _lsr(x, 31)
    return _lsr(x, 1)


def make_sites_np(seed, length, mean_gap=MEAN_GAP, island_site_frac=0.10, island_seq_frac=0.01):
    """Sorted CpG positions of one contig (numpy, host) and their per-site methylation level as a 24-bit integer
    threshold: a two-state (island / open sea) gap model; beta ~ 0.6 Beta(8, 1.5) + 0.4 Beta(1.2, 6)."""
    rng = np.random.default_rng(seed)
    n_target = int(length / mean_gap)
    n_isl = int(n_target * island_site_frac)
    n_seg = max(1, int(length * island_seq_frac / 1000))
    seg_start = np.sort(rng.integers(0, max(1, length - 1000), n_seg))
    isl = seg_start[rng.integers(0, n_seg, n_isl)] + rng.integers(0, 1000, n_isl)
    sea = rng.integers(0, length - 1, n_target - n_isl)
    pos = np.unique(np.concatenate([isl, sea]).astype(np.int64))
    keep = np.ones(len(pos), bool)
    keep[1:] = np.diff(pos) >= 2  # a CpG occupies 2 bp
    pos = pos[keep]
    pos = pos[pos < length - 1]
    mix = rng.random(len(pos)) < 0.6
    beta = np.where(mix, rng.beta(8, 1.5, len(pos)), rng.beta(1.2, 6, len(pos)))
    return pos.astype(np.int32), np.minimum((beta * (1 << 24)).astype(np.int64), (1 << 24) - 1).astype(np.int32)


def make_contig(device, seed, tid, length, coverage, sites=None, beta_q=None, read_len=READ_LEN, nocall=0.01, lowq=0.08,
                stay=0.85, mean_gap=MEAN_GAP, start_range=None):
    """One contig of reads as torch tensors on `device` in the layout of mth_batch (start/end/meta i32, cpg_off i32,
    cpg_pos i32, cpg_rel i16, meth i64; unsigned fields carried in the signed type of the same width).
    start_range=(lo, hi): only the reads with lo <= start < hi (read indices stay those of the whole contig, so a slice
    is bit-identical to the same reads of the full contig).  -> dict"""
    if sites is None:
        sites, beta_q = make_sites_np(seed * 1000 + tid, length, mean_gap)
    dev = torch.device(device)
    sites_t = torch.as_tensor(np.asarray(sites, np.int64), device=dev)
    beta_t = torch.as_tensor(np.asarray(beta_q, np.int64), device=dev)
    R = int(coverage * length / read_len)
    s0 = seed * 1000 + tid
    idx = torch.arange(R, dtype=torch.int64, device=dev)
    start = torch.sort(hash63(s0, 0, idx) % max(1, length - read_len - 1)).values
    if start_range is not None:
        a = int(torch.searchsorted(start, torch.tensor([start_range[0]], device=dev))[0])
        b = int(torch.searchsorted(start, torch.tensor([start_range[1]], device=dev))[0])
        start, idx = start[a:b].contiguous(), idx[a:b].contiguous()
        R = b - a
    h = hash63(s0, 1, idx)                 # bits 0: strand, 8..: low-quality draw, 40..: low mapq value
    is_rev = h & 1
    low = (_lsr(h, 8) & 0xFFFFFF) < int(lowq * (1 << 24))
    mapq = torch.where(low, _lsr(h, 40) % 10, torch.full_like(h, MAPQ_HIGH))
    end = start + (read_len - 1)
    # candidate sites: forward reads call site x at abspos x, reverse reads at abspos x + 1 (pos = abspos - 1 = x)
    lo = torch.searchsorted(sites_t, start - is_rev, right=False)
    hi = torch.searchsorted(sites_t, end - is_rev, right=True)
    cnt = hi - lo
    tot = int(cnt.sum())
    ridx = torch.repeat_interleave(idx - idx[0] if R else idx, cnt, output_size=tot)
    first = torch.cumsum(cnt, 0) - cnt
    k_in = torch.arange(tot, dtype=torch.int64, device=dev) - first[ridx]
    sidx = lo[ridx] + k_in
    gread = ridx + (idx[0] if R else 0)    # read index within the whole contig: the hash key of a call is (read, k)
    hc = hash63(s0, 2, gread * 256 + k_in)
    called = (hc & 0xFFFFFF) >= int(nocall * (1 << 24))
    fresh_draw = (_lsr(hc, 24) & 0xFFFFFF) < beta_t[sidx]
    restart = (_lsr(hc, 48) & 0x7FFF) >= int(stay * (1 << 15))
    del hc, gread
    # keep the called sites only; methylation state: first-order Markov along the read's retained calls
    ridx, sidx, k_in = ridx[called], sidx[called], None
    fresh_draw, restart = fresh_draw[called], restart[called]
    n = int(ridx.numel())
    cnt2 = torch.bincount(ridx, minlength=R) if n else torch.zeros(R, dtype=torch.int64, device=dev)
    off = torch.zeros(R + 1, dtype=torch.int64, device=dev)
    torch.cumsum(cnt2, 0, out=off[1:])
    ar = torch.arange(n, dtype=torch.int64, device=dev)
    k2 = ar - off[ridx]
    fresh = (k2 == 0) | restart
    src = torch.cummax(torch.where(fresh, ar, torch.zeros_like(ar)), 0).values if n else ar
    state = fresh_draw[src].to(torch.int64)
    x = sites_t[sidx]
    rel = x + is_rev[ridx] - start[ridx]   # query index of a `150M` read
    words = torch.clamp((cnt2 + 63) // 64, min=1)
    multi = bool(int(words.max()) > 1) if R else False
    moff = torch.zeros(R + 1, dtype=torch.int64, device=dev)
    torch.cumsum(words, 0, out=moff[1:])
    meth = torch.zeros(int(moff[-1]), dtype=torch.int64, device=dev)
    if n:
        meth.index_add_(0, moff[ridx] + (k2 >> 6), state << (k2 & 63))  # distinct bits: sum == or (bit 63 wraps correctly)
    meta = (mapq | ((1 - is_rev) << 8)).to(torch.int32)
    return dict(tid=int(tid), n_reads=R, n_cpg=n, start=start.to(torch.int32), end=end.to(torch.int32), meta=meta,
                cpg_off=off.to(torch.int32), cpg_pos=x.to(torch.int32), cpg_rel=rel.to(torch.int16), meth=meth,
                meth_off=moff.to(torch.int32) if multi else None, n_meth_words=int(moff[-1]), n_sites_model=int(len(sites)))


def to_numpy_batch(b):
    """torch batch (any device) -> the numpy dict the oracle helpers and tests use (unsigned dtypes restored)."""
    u = dict(meta=np.uint32, cpg_off=np.uint32, cpg_rel=np.uint16, meth=np.uint64, meth_off=np.uint32)
    out = {}
    for k, v in b.items():
        if isinstance(v, torch.Tensor):
            a = v.detach().cpu().numpy()
            out[k] = a.view(u[k]) if k in u else a
        else:
            out[k] = v
    return out


def genome(scale=1.0):
    """[(name, length)] of the 24 hg38 chromosomes, optionally shrunk (tests)."""
    return [(n, max(2000, int(l * scale))) for n, l in HG38]
