"""CPU: the bench's test infrastructure — the torch whole-genome generator (bit-identical slices, valid batches) and the all-core
oracle (position bins + halo == one single-threaded pass, the same argument the multi-GPU sharding rests on)."""
import numpy as np

import oracle_parallel as OP
import parity
from metheor_b200 import batch as B
from metheor_b200 import bamdec, synth_gpu as G
from oracle_lib import Oracle


def test_generator_is_deterministic_valid_and_sliceable():
    a = G.to_numpy_batch(G.make_contig("cpu", 20260102, 20, 1_500_000, 30.0))
    b = G.to_numpy_batch(G.make_contig("cpu", 20260102, 20, 1_500_000, 30.0))
    for k in ("start", "end", "meta", "cpg_off", "cpg_pos", "cpg_rel", "meth"):
        assert np.array_equal(a[k], b[k]), k
    st = a["start"].astype(np.int64)
    off = a["cpg_off"].astype(np.int64)
    cnt = np.diff(off)
    ridx = np.repeat(np.arange(len(cnt)), cnt)
    pos = a["cpg_pos"].astype(np.int64)
    assert (np.diff(st) >= 0).all() and (a["end"] - a["start"] == 149).all()
    assert (pos >= st[ridx] - 1).all() and (pos <= a["end"][ridx]).all() and cnt.max() <= 64
    rev = 1 - ((a["meta"] >> 8) & 1)
    assert (a["cpg_rel"] == pos + rev[ridx] - st[ridx]).all()
    assert 0.45 < ((a["meta"] >> 8) & 1).mean() < 0.55 and 0.06 < ((a["meta"] & 0xFF) < 10).mean() < 0.10
    # a start-range slice is bit-identical to the same reads of the whole contig (what the multi-GPU bench relies on)
    s = G.to_numpy_batch(G.make_contig("cpu", 20260102, 20, 1_500_000, 30.0, start_range=(400_000, 900_000)))
    lo, hi = np.searchsorted(st, 400_000), np.searchsorted(st, 900_000)
    full = B.slice_range(a, int(lo), int(hi))
    for k in ("start", "end", "meta", "cpg_off", "cpg_pos", "cpg_rel", "meth"):
        assert np.array_equal(full[k], s[k]), k
    # another seed / contig gives other reads
    c = G.to_numpy_batch(G.make_contig("cpu", 20260102, 21, 1_500_000, 30.0))
    assert not np.array_equal(a["start"][:1000], c["start"][:1000])
    assert len(G.genome()) == 24 and sum(l for _, l in G.genome()) == G.GENOME_LEN


def test_all_core_oracle_equals_one_pass():
    nb = G.to_numpy_batch(G.make_contig("cpu", 5, 3, 600_000, 25.0, mean_gap=40.0, nocall=0.03))
    o = Oracle.from_soa(**B.to_oracle_soa([nb]))
    f32 = lambda x: np.asarray(x, np.float32).view(np.uint32)
    for m in ("pdr", "mhl", "fdrp", "qfdrp", "pm", "lpmd"):
        prm = dict(parity.DEFAULTS[m])
        if m not in ("lpmd", "pm"):
            prm["min_depth"] = 5
        want = {"pdr": lambda: o.pdr(**prm), "mhl": lambda: o.mhl(**prm), "fdrp": lambda: o.fdrp(**prm), "qfdrp": lambda: o.fdrp(quantitative=True, **prm),
                "pm": lambda: o.quartets(**prm), "lpmd": lambda: o.lpmd(**prm)}[m]()
        got, info = OP.run(nb, "quartets" if m == "pm" else m, prm, n_proc=3)
        if m == "lpmd":
            assert all(int(got[k]) == int(want[k]) for k in ("n_read", "n_valid_read", "n_conc", "n_disc")) and f32(got["lpmd"]) == f32(want["lpmd"])
            continue
        assert info["bins"] > 1
        for k, v in want.items():
            a, b = (f32(got[k]), f32(v)) if v.dtype == np.float32 else (got[k], v)
            assert np.array_equal(a, b), (m, k)
        # a sub-interval: rows of the interval only
        sub, _ = OP.run(nb, "quartets" if m == "pm" else m, prm, n_proc=2, interval=(100_000, 300_000))
        key = "p1" if m == "pm" else "pos"
        sel = (want[key] >= 100_000) & (want[key] < 300_000)
        assert np.array_equal(sub[key], want[key][sel])


def test_bgzf_member_walk():
    import struct, zlib
    import bamio
    data = bytes(np.random.default_rng(1).integers(0, 40, 200_000, dtype=np.uint8))
    raw = bamio.bgzf_compress(data, block=30_000)
    mem = bamdec.bgzf_members(raw)
    assert len(mem) == 8 and mem[-1][2] == 0 and sum(m[2] for m in mem) == len(data) and mem[-1][3] == len(raw)
    assert b"".join(zlib.decompress(raw[o:o + s], -15) for o, s, _, _, _ in mem) == data
    assert all(zlib.crc32(zlib.decompress(raw[o:o + s], -15)) == c for o, s, _, _, c in mem)
