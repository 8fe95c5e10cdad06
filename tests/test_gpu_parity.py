"""GPU parity: the CUDA engine (through the C ABI) against the CPU oracle, bit-exact, on the reference's fixtures
and on seeded synthetic reads that exercise what the fixtures do not (both strands, deletions, no-calls, low mapq,
segment hazards, several batches / contigs / regions, deep piles)."""
import numpy as np
import pytest

import parity
from metheor_b200 import batch as B
from metheor_b200 import engine, synth

pytestmark = pytest.mark.gpu

ALL = ("pdr", "lpmd", "mhl", "pm", "me", "fdrp", "qfdrp")
CHR1 = [248956422]


def test_reference_fixture_pins(golden, pins):
    """The values the reference's unit tests assert (tests/golden/reference_pins.json), now from the GPU."""
    f32 = np.float32
    for c in pins["pdr"]["cases"]:
        b = parity.records_to_batch(golden[c["input"]]["reads"])
        res, _ = engine.run_batches([b], CHR1, ["pdr"], pdr=dict(min_depth=0, min_cpgs=c["min_cpgs"], min_qual=10))
        assert res["pdr"]["n"] == c["n_rows"]
        if c["n_rows"]:
            assert (res["pdr"]["value"] == f32(c["pdr"])).all() and (res["pdr"]["n_conc"] == c["n_conc"]).all() \
                and (res["pdr"]["n_disc"] == c["n_disc"]).all()
    for c in pins["lpmd"]["cases"]:
        b = parity.records_to_batch(golden[c["input"]]["reads"])
        v = engine.run_batches([b], CHR1, ["lpmd"])[0]["lpmd"]["lpmd"]
        assert np.isnan(v) if c["lpmd"] == "NaN" else v == f32(c["lpmd"])
    for c in pins["mhl"]["cases"]:
        b = parity.records_to_batch(golden[c["input"]]["reads"])
        r = engine.run_batches([b], CHR1, ["mhl"], mhl=dict(min_depth=0, min_cpgs=0, min_qual=10))[0]["mhl"]
        assert r["n"] == c["n_rows"] and (r["value"] == f32(c.get("mhl", 0))).all()
    for name in ("pm", "me"):
        for c in pins[name]["cases"]:
            b = parity.records_to_batch(golden[c["input"]]["reads"])
            r = engine.run_batches([b], CHR1, [name], **{name: dict(min_depth=0, min_qual=10)})[0][name]
            assert r["n"] == c["n_quartets"] and (r["value"] == f32(c.get(name, 0))).all(), (name, c, r["value"])
    for name in ("fdrp", "qfdrp"):
        a = pins[name]["args"]
        for c in pins[name]["cases"]:
            b = parity.records_to_batch(golden[c["input"]]["reads"])
            r = engine.run_batches([b], CHR1, [name], **{name: dict(min_qual=c["min_qual"], min_depth=a["min_depth"],
                                                                    max_depth=a["max_depth"], min_overlap=a["min_overlap"])})[0][name]
            assert list(r["pos"]) == c["positions"]
            if c["positions"]:
                if name + "_f32_of" in c:
                    x, y = c[name + "_f32_of"].split("/")
                    want = f32(x) / f32(y)
                else:
                    want = f32(c[name])
                if c["exact"]:
                    assert (r["value"] == want).all(), (name, c, r["value"])
                else:
                    assert (np.abs(r["value"] - want) < c["tol"]).all()


@pytest.mark.parametrize("name", ["test1", "test2", "test3", "test4", "test5", "test6"])
def test_fixtures_default_flags_vs_oracle(golden, name):
    b = parity.records_to_batch(golden[name]["reads"])
    parity.check_all([b], CHR1, ALL)
    parity.check_all([b], CHR1, ALL, pdr=dict(min_depth=1, min_cpgs=1), mhl=dict(min_depth=1, min_cpgs=1),
                     pm=dict(min_depth=1), me=dict(min_depth=1), fdrp=dict(min_depth=1, min_overlap=1, max_depth=100),
                     qfdrp=dict(min_depth=1, min_overlap=1, max_depth=100))


def test_chr19_real_reads(golden):
    """1000 real Bismark reads (both strands, 24-29 bp) of the reference's tag fixture."""
    fx = golden["chr19_1000"]
    b = parity.records_to_batch(fx["reads"])
    ref_len = [l for _, l in fx["refs"]]
    parity.check_all([b], ref_len, ALL, pdr=dict(min_depth=2, min_cpgs=1), mhl=dict(min_depth=2, min_cpgs=1),
                     pm=dict(min_depth=1), me=dict(min_depth=1), fdrp=dict(min_depth=2, min_overlap=5),
                     qfdrp=dict(min_depth=2, min_overlap=5))


def _synth(seed, length=200_000, cov=30.0, **kw):
    sites = synth.make_sites(seed, length)
    return synth.make_reads(seed + 1, sites, length, cov, **kw)


@pytest.mark.parametrize("seed", [11, 12])
def test_synthetic_default_flags(seed):
    b = _synth(seed)
    res, st = parity.check_all([b], [200_000], ALL)
    assert st["pdr_path"] == 1  # 150-bp reads: the scatter path is provably exact
    assert res["pdr"]["n"] > 1000 and res["mhl"]["n"] > 100 and res["pm"]["n"] > 100 and res["fdrp"]["n"] > 1000


def test_synthetic_forced_gather_equals_scatter():
    b = _synth(13)
    parity.check_all([b], [200_000], ("pdr",), flags=engine.FLAG_FORCE_GATHER)


def test_synthetic_long_spans_segment_hazards():
    """Deletions stretch reads to > 150 reference bases: PDR must take the gather path and replay flush segments."""
    b = _synth(14, read_len=140, del_frac=0.5, del_max=60, nocall=0.05, lowq=0.15)
    res, st = parity.check_all([b], [200_000], ALL, pdr=dict(min_depth=3, min_cpgs=2), mhl=dict(min_depth=3, min_cpgs=2),
                               fdrp=dict(min_depth=3), qfdrp=dict(min_depth=3), pm=dict(min_depth=3), me=dict(min_depth=3))
    assert st["pdr_path"] == 3 and st["max_ref_span"] > 150  # scatter + segment-exact gather on the hazard sites


def test_synthetic_dense_islands_many_cpgs_per_read():
    sites = synth.make_sites(15, 60_000, mean_gap=6.0)
    b = synth.make_reads(16, sites, 60_000, 20.0)
    n = np.diff(b["cpg_off"].astype(np.int64)).max()
    assert n > 16
    parity.check_all([b], [60_000], ALL, pdr=dict(min_depth=5), mhl=dict(min_depth=5))


def test_more_than_64_cpgs_per_read_uses_meth_off():
    sites = np.arange(10, 20_000, 2, dtype=np.int32)  # a CpG every 2 bp: 75 calls per 150-bp read
    b = synth.make_reads(17, sites, 20_000, 12.0, nocall=0.02)
    assert b["meth_off"] is not None and np.diff(b["cpg_off"].astype(np.int64)).max() > 64
    parity.check_all([b], [20_000], ALL, pdr=dict(min_depth=4), mhl=dict(min_depth=4), fdrp=dict(min_depth=4),
                     qfdrp=dict(min_depth=4), pm=dict(min_depth=4), me=dict(min_depth=4))


def test_batches_contigs_and_zero_cpg_reads():
    """Several batches per contig and several contigs in one region: identical to one oracle pass over all reads."""
    lens = [150_000, 80_000, 120_000]
    batches = []
    for tid, L in enumerate(lens):
        b = _synth(20 + tid, length=L, cov=25.0, tid=tid)
        b["tid"] = tid
        cuts = [0, b["n_reads"] // 3, b["n_reads"] // 3 + 1, (2 * b["n_reads"]) // 3, b["n_reads"]]
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            batches.append(B.slice_reads(b, lo, hi))
    res, st = parity.check_all(batches, lens, ALL)
    assert st["n_regions"] == 1 and set(np.unique(res["pdr"]["tid"])) == {0, 1, 2}


def test_region_split_when_contigs_do_not_fit():
    lens = [1_400_000_000, 1_300_000_000, 900_000]
    batches = []
    for tid, off in ((0, 1_399_000_000), (1, 5_000), (2, 100)):
        b = _synth(30 + tid, length=100_000, cov=15.0)
        b["tid"] = tid
        for k in ("start", "end", "cpg_pos"):
            b[k] = (b[k].astype(np.int64) + off).astype(np.int32)
        batches.append(b)
    res, st = parity.check_all(batches, lens, ("pdr", "lpmd", "mhl", "pm"))
    assert st["n_regions"] == 2


def test_deep_piles_reservoir_sampling_matches_seeded_oracle():
    b = _synth(40, length=30_000, cov=90.0)
    parity.check_all([b], [30_000], ("fdrp", "qfdrp"), seed=1234)
    parity.check_all([b], [30_000], ("fdrp", "qfdrp"), seed=99, fdrp=dict(max_depth=4096), qfdrp=dict(max_depth=4096))


def test_device_resident_batch_is_borrowed_zero_copy():
    torch = pytest.importorskip("torch")
    b = _synth(50)
    d = dict(b)
    for k in ("start", "end", "meta", "cpg_off", "cpg_pos", "cpg_rel", "meth"):
        a = b[k]
        view = {np.dtype("uint32"): np.int32, np.dtype("uint16"): np.int16, np.dtype("uint64"): np.int64}.get(a.dtype, a.dtype)
        d[k] = torch.from_numpy(a.view(view)).cuda()
    want, _ = engine.run_batches([b], [200_000], ALL)
    got, st = engine.run_batches([d], [200_000], ALL)
    assert st["h2d_bytes"] == 0
    for m in want:
        for k in want[m]:
            assert np.array_equal(np.asarray(want[m][k]), np.asarray(got[m][k]), equal_nan=True), (m, k)


def test_error_paths():
    b = _synth(60, length=50_000, cov=5.0)
    bad = dict(b)
    bad["start"] = b["start"][::-1].copy()
    with pytest.raises(engine.EngineError) as e:
        engine.run_batches([bad], [50_000], ("pdr",))
    assert e.value.code in (engine._lib.ERR_UNSORTED, engine._lib.ERR_INVALID)
    b2 = dict(b); b2["tid"] = 1
    b1 = dict(b); b1["tid"] = 0
    with pytest.raises(engine.EngineError) as e:
        engine.run_batches([b2, b1], [50_000, 50_000], ("pdr",))
    assert e.value.code == engine._lib.ERR_UNSORTED
    with pytest.raises(engine.EngineError):
        engine.run_batches([dict(b, cpg_rel=None)], [50_000], ("lpmd",))  # LPMD needs the query indices
    empty = B.slice_reads(b, 0, 0)
    res, _ = engine.run_batches([empty], [50_000], ALL)
    assert res["pdr"]["n"] == 0 and np.isnan(res["lpmd"]["lpmd"])


def test_lpmd_pairs_table(golden):
    """lpmd --pairs (lpmd.rs:89-122): SURVEY Appendix B expects six rows `chr1 a b 0.5 8 8` on test1."""
    b = parity.records_to_batch(golden["test1"]["reads"])
    res, _ = parity.check_all([b], CHR1, ("lpmd",), lpmd=dict(want_pairs=1))
    pr = res["lpmd"]["pairs"]
    assert list(zip(pr["pos1"], pr["pos2"])) == [(0, 2), (0, 4), (0, 6), (2, 4), (2, 6), (4, 6)]
    assert (pr["lpmd"] == np.float32(0.5)).all() and (pr["n_conc"] == 8).all() and (pr["n_disc"] == 8).all()
    for seed, kw in ((81, {}), (82, dict(read_len=140, del_frac=0.5, del_max=60, nocall=0.05, lowq=0.15))):
        s = _synth(seed, **kw)
        res, _ = parity.check_all([s], [200_000], ("lpmd", "pdr"), lpmd=dict(want_pairs=1, min_distance=1, max_distance=40))
        assert res["lpmd"]["pairs"]["n"] > 1000
    lens = [150_000, 80_000]
    batches = []
    for tid, L in enumerate(lens):
        x = _synth(90 + tid, length=L, cov=20.0, tid=tid)
        batches += [B.slice_reads(x, 0, x["n_reads"] // 2), B.slice_reads(x, x["n_reads"] // 2, x["n_reads"])]
    parity.check_all(batches, lens, ("lpmd",), lpmd=dict(want_pairs=1))


def test_position_bin_sharding_with_halo_equals_one_pass():
    """two virtual ranks (run one after the other on this GPU) each take a position bin + halo; owned
    rows concatenated and LPMD counters summed equal a single engine pass and the oracle."""
    from metheor_b200 import shard
    lens = [180_000, 70_000]
    data = []
    for tid, L in enumerate(lens):
        data.append(_synth(100 + tid, length=L, cov=25.0, tid=tid, read_len=140, del_frac=0.3, del_max=50, nocall=0.03, lowq=0.1))
    ov = dict(pdr=dict(min_depth=5, min_cpgs=2), mhl=dict(min_depth=5, min_cpgs=2), fdrp=dict(min_depth=5, min_overlap=20),
              qfdrp=dict(min_depth=5, min_overlap=20), pm=dict(min_depth=5), me=dict(min_depth=5), lpmd=dict(want_pairs=1))
    whole, _ = parity.check_all(data, lens, ALL, **ov)
    for world in (2, 3):
        plan = shard.plan_bins(lens, world, weights=[b["start"] for b in data])
        parts, lp = [], np.zeros(4, np.int64)
        for r in range(world):
            mine = [x for x in (shard.select_shard(b, plan[r], halo=400)[0] for b in data) if x is not None and x["n_reads"]]
            res, _ = engine.run_batches(mine, lens, ALL, **{m: dict(parity.DEFAULTS[m], **ov.get(m, {})) for m in ALL})
            lp += np.array([res["lpmd"][k] for k in ("n_read", "n_valid_read", "n_conc", "n_disc")], np.int64)
            part = {m: shard.owned_rows(res[m], plan[r]) for m in ("pdr", "mhl", "fdrp", "qfdrp")}
            part.update({m: shard.owned_rows(res[m], plan[r], pos_key="p1") for m in ("pm", "me")})
            part["pairs"] = shard.owned_rows(res["lpmd"]["pairs"], plan[r], pos_key="pos1")
            parts.append(part)
        assert tuple(lp) == tuple(whole["lpmd"][k] for k in ("n_read", "n_valid_read", "n_conc", "n_disc"))
        for m in ("pdr", "mhl", "fdrp", "qfdrp", "pm", "me", "pairs"):
            keys = dict(pm=("tid", "p1", "p2", "p3", "p4"), me=("tid", "p1", "p2", "p3", "p4"), pairs=("tid", "pos1", "pos2")).get(m, ("tid", "pos"))
            merged = shard.merge_rows([p[m] for p in parts], keys=keys)
            want = whole["lpmd"]["pairs"] if m == "pairs" else whole[m]
            assert merged["n"] == want["n"] and merged["n"] > 100, (m, world)
            for k, v in want.items():
                if isinstance(v, np.ndarray):
                    a, b2 = merged[k], v
                    if a.dtype == np.float32:
                        a, b2 = a.view(np.uint32), b2.view(np.uint32)
                    assert np.array_equal(a, b2), (m, k, world)


def test_compact_wire_format_equals_soa(golden):
    """mth_submit_compact (9 B/read + 2.125 B/call, expanded on the device) gives the same rows as mth_submit."""
    ov = dict(pdr=dict(min_depth=3, min_cpgs=2), mhl=dict(min_depth=3, min_cpgs=2), fdrp=dict(min_depth=3), qfdrp=dict(min_depth=3),
              pm=dict(min_depth=3), me=dict(min_depth=3), lpmd=dict(want_pairs=1))
    b = parity.records_to_batch(golden["test4"]["reads"])
    parity.check_all([b], CHR1, ALL, compact=True)
    fx = golden["chr19_1000"]
    parity.check_all([parity.records_to_batch(fx["reads"])], [l for _, l in fx["refs"]], ALL, compact=True, **ov)
    # deletions: query indices are not implied by the positions -> rel_exc
    s = _synth(110, read_len=140, del_frac=0.5, del_max=60, nocall=0.05, lowq=0.15)
    c = B.to_compact(s)
    assert c["n_rel"] > 0 and (c["flags"] & 4).any() and not (c["flags"] & 4).all()
    parity.check_all([s], [200_000], ALL, compact=True, **ov)
    # several batches and contigs, formats alternating
    lens = [150_000, 80_000, 120_000]
    batches = []
    for tid, L in enumerate(lens):
        x = _synth(120 + tid, length=L, cov=25.0, tid=tid)
        cuts = [0, x["n_reads"] // 3, (2 * x["n_reads"]) // 3, x["n_reads"]]
        batches += [B.slice_reads(x, lo, hi) for lo, hi in zip(cuts[:-1], cuts[1:])]
    parity.check_all(batches, lens, ALL, compact=True)
    parity.check_all(batches, lens, ALL, compact="mix")
    # a read denser than the compact format allows falls back to the SoA batch; both kinds in one region
    dense = synth.make_reads(17, np.arange(10, 20_000, 2, dtype=np.int32), 20_000, 12.0, nocall=0.02)
    sparse = _synth(130, length=150_000, tid=1)
    parity.check_all([dense, sparse], [20_000, 150_000], ALL, compact=True, pdr=dict(min_depth=4), mhl=dict(min_depth=4))
    with pytest.raises(ValueError):
        B.to_compact(dense)


def test_dense_block_encodings_equal_soa(golden):
    """MTH_CENC_START16 | MTH_CENC_DELTA8: 16-bit start offsets per block of 256 reads and 8-bit chained call deltas, with
    per-block fall-back to the wide types (coverage gaps, long reads)."""
    ov = dict(pdr=dict(min_depth=3, min_cpgs=2), mhl=dict(min_depth=3, min_cpgs=2), fdrp=dict(min_depth=3), qfdrp=dict(min_depth=3),
              pm=dict(min_depth=3), me=dict(min_depth=3), lpmd=dict(want_pairs=1))
    parity.check_all([parity.records_to_batch(golden["test4"]["reads"])], CHR1, ALL, compact="dense")
    s = _synth(140, read_len=140, del_frac=0.5, del_max=60, nocall=0.05, lowq=0.15)
    d = B.to_compact(s, dense=True)
    assert d["enc"] == 3 and d["n_delta8"] > 0 and "start" not in d
    parity.check_all([s], [200_000], ALL, compact="dense", **ov)
    # a coverage gap wider than 65535 inside a block (32-bit start block) and a read with a 300-bp gap between calls (16-bit block)
    a = _synth(141, length=400_000, cov=6.0)
    keep = (a["start"] < 120_000) | (a["start"] > 260_000)
    a = B.select_reads(a, keep)
    longr = dict(tid=0, pos=395_000, flag=0, mapq=42, cigar="10M300N20M", xm="Z........." + "." * 10 + "Z........z")
    lb = parity.records_to_batch([longr])
    a2 = B.select_reads(a, a["start"] < 394_000)
    d = B.to_compact(a2, dense=True)
    assert d["n_start_exc"] == 256 and (d["blk_start"] < 0).sum() == 1
    dl = B.to_compact(lb, dense=True)
    assert dl["n_delta16"] == 3 and dl["n_delta8"] == 0
    parity.check_all([a2, lb], [400_000], ALL, compact="dense", **ov)
    # several batches / contigs, all three wire formats interleaved
    lens = [150_000, 80_000, 120_000]
    batches = []
    for tid, L in enumerate(lens):
        x = _synth(150 + tid, length=L, cov=25.0, tid=tid)
        cuts = [0, x["n_reads"] // 4, x["n_reads"] // 2, (3 * x["n_reads"]) // 4, x["n_reads"]]
        batches += [B.slice_reads(x, lo, hi) for lo, hi in zip(cuts[:-1], cuts[1:])]
    parity.check_all(batches, lens, ALL, compact="dense")
    parity.check_all(batches, lens, ALL, compact="mix")


def test_compact_error_paths_and_reserve():
    b = _synth(160, length=60_000, cov=8.0)
    c = B.to_compact(b, dense=True)
    ctx = engine.Context(engine.default_params(ALL), [60_000])
    try:
        assert ctx._L.mth_reserve(ctx._h, 10_000_000, 40_000_000) == 0  # capacity hint: harmless, clamped to free memory
        bad = dict(c, enc=8)
        with pytest.raises(engine.EngineError) as e:
            ctx.submit_compact(bad)
        assert e.value.code == engine._lib.ERR_INVALID
        ctx.reset()
        bad = dict(c, n_delta8=c["n_delta8"] - 1)  # n_delta8 + n_delta16 != n_cpg
        with pytest.raises(engine.EngineError):
            ctx.submit_compact(bad)
        ctx.reset()
        bad = dict(c)
        bad["n_cpg8"] = c["n_cpg8"].copy()
        bad["n_cpg8"][5] = 200  # more than 64 calls in a compact read: reported, never silently truncated
        with pytest.raises(engine.EngineError) as e:
            ctx.submit_compact(bad)
            ctx.finish()
        assert e.value.code in (engine._lib.ERR_UNSUPPORTED, engine._lib.ERR_INVALID)
        ctx.reset()
        plain = B.to_compact(b)
        unsorted_ = dict(plain, start=plain["start"][::-1].copy())
        with pytest.raises(engine.EngineError) as e:
            ctx.submit_compact(unsorted_)
            ctx.finish()
        assert e.value.code in (engine._lib.ERR_UNSORTED, engine._lib.ERR_INVALID)
        ctx.reset()
        ctx.submit_compact(c)  # the context is still usable after the failures
        res = ctx.finish()
        assert res["pdr"]["n"] >= 0 and res["lpmd"]["n_read"] == b["n_reads"]
    finally:
        ctx.close()


def test_pdr_flush_segment_is_replayed_on_hazard_sites_only():
    """A constructed PDR flush (pdr.rs:160-177): read T (a 200-base skip, first CpG 212 bases behind its start) flushes the
    CpGs of read A before read B contributes to the same CpG again -> the reference reports B's segment only."""
    reads = [dict(tid=0, pos=100, flag=0, mapq=42, cigar="20M", xm=".....Z....Z........."),   # A: CpGs 105 (Z), 110 (Z)
             dict(tid=0, pos=101, flag=0, mapq=42, cigar="10M200N20M", xm="." * 10 + "..Z......z.........."),  # T: first CpG 313
             dict(tid=0, pos=104, flag=0, mapq=42, cigar="20M", xm=".z....Z............."),    # B: CpGs 105 (z), 110 (Z) -> discordant
             dict(tid=0, pos=300, flag=0, mapq=42, cigar="20M", xm=".............Z......")]    # C: CpG 313
    b = parity.records_to_batch(reads)
    res, st = parity.check_all([b], [10_000], ("pdr",), pdr=dict(min_depth=1, min_cpgs=1, min_qual=10))
    assert st["pdr_path"] == 3
    r = res["pdr"]
    got = {int(p): (int(c), int(d)) for p, c, d in zip(r["pos"], r["n_conc"], r["n_disc"])}
    assert got[105] == (0, 1) and got[110] == (0, 1), got   # only B's segment survives (A's was flushed, then overwritten)
    assert got[313] == (1, 1)                               # T (discordant) + C


def test_cpg_set_filter_on_the_device():
    """--cpg-set as a device bitmap (mth_set_cpg_set; readutil.rs:87-95, 347-374): the host ships unfiltered calls, the engine
    drops the ones outside the set before anything else — SoA, compact and dense batches, several batches and contigs, reads
    that lose all their calls, reads with more than 64 calls."""
    lens = [150_000, 90_000]
    rng = np.random.default_rng(77)
    batches, st, sp = [], [], []
    for tid, L in enumerate(lens):
        sites = synth.make_sites(500 + tid, L)
        b = synth.make_reads(510 + tid, sites, L, 25.0, tid=tid, del_frac=0.1)
        keep = sites[rng.random(len(sites)) < 0.55]
        st.append(np.full(len(keep), tid, np.int32)); sp.append(keep.astype(np.int32))
        cuts = [0, b["n_reads"] // 2, b["n_reads"]]
        batches += [B.slice_reads(b, lo, hi) for lo, hi in zip(cuts[:-1], cuts[1:])]
    cs = (np.concatenate(st), np.concatenate(sp))
    kw = dict(pdr=dict(min_depth=4, min_cpgs=2), mhl=dict(min_depth=4, min_cpgs=2), fdrp=dict(min_depth=4), qfdrp=dict(min_depth=4),
              pm=dict(min_depth=3), me=dict(min_depth=3), lpmd=dict(want_pairs=1))
    for compact in (False, True, "dense", "mix"):
        res, stt = parity.check_all(batches, lens, ALL, cpg_set=cs, compact=compact, **{k: dict(v) for k, v in kw.items()})
        assert res["pdr"]["n"] > 100 and res["pm"]["n"] > 10 and stt["n_cpg"] < sum(b["n_cpg"] for b in batches)
    # every call filtered out; an empty set; entries in random order with duplicates and positions nobody calls
    parity.check_all(batches, lens, ALL, cpg_set=(np.zeros(1, np.int32), np.array([7], np.int32)))
    parity.check_all(batches, lens, ("pdr", "lpmd", "pm"), cpg_set=(np.zeros(0, np.int32), np.zeros(0, np.int32)))
    perm = rng.permutation(len(cs[0]))
    parity.check_all(batches, lens, ("pdr", "mhl", "lpmd"), cpg_set=(np.concatenate([cs[0][perm], cs[0][:50]]), np.concatenate([cs[1][perm], cs[1][:50]])),
                     pdr=dict(min_depth=4, min_cpgs=2), mhl=dict(min_depth=4, min_cpgs=2))
    # more than 64 calls per read (meth_off): methylation words squeezed across word boundaries
    sites = np.arange(10, 20_000, 2, dtype=np.int32)
    b = synth.make_reads(517, sites, 20_000, 10.0, nocall=0.02)
    keep = sites[rng.random(len(sites)) < 0.9]
    parity.check_all([b], [20_000], ALL, cpg_set=(np.zeros(len(keep), np.int32), keep), pdr=dict(min_depth=4), mhl=dict(min_depth=4),
                     fdrp=dict(min_depth=4), qfdrp=dict(min_depth=4), pm=dict(min_depth=4), me=dict(min_depth=4))


def test_regions_with_host_soa_batches_and_mixed_quartets():
    """Several regions fed through mth_submit with HOST SoA batches (the copy stream writes the next region's reads at arena
    offset 0 while the previous region's PM/ME and --pairs emit kernels may still be reading it: the copy stream must wait)."""
    lens = [1_400_000_000, 1_300_000_000, 1_200_000_000, 900_000]
    batches = []
    for tid, off in ((0, 1_399_000_000), (1, 5_000), (2, 700_000_000), (3, 100)):
        sites = synth.make_sites(600 + tid, 200_000, mean_gap=12.0)
        b = synth.make_reads(610 + tid, sites, 200_000, 20.0, tid=tid, nocall=0.05, del_frac=0.05)
        for k in ("start", "end", "cpg_pos"):
            b[k] = (b[k].astype(np.int64) + off).astype(np.int32)
        cuts = [0, b["n_reads"] // 2, b["n_reads"]]
        batches += [B.slice_reads(b, lo, hi) for lo, hi in zip(cuts[:-1], cuts[1:])]
    for _ in range(3):
        res, st = parity.check_all(batches, lens, ("pm", "me", "lpmd", "pdr"), pm=dict(min_depth=3), me=dict(min_depth=3),
                                   pdr=dict(min_depth=4), lpmd=dict(want_pairs=1))
        assert st["n_regions"] >= 3 and res["pm"]["n"] > 1000


@pytest.mark.gpu
@pytest.mark.parametrize("coverage", [40.0, 400.0])
def test_mixed_quartet_sites_shallow_and_deep_windows(coverage):
    """PM / ME sites with several quartet keys (no-calls and deletions): at 40x the window of a site fits the lanes' register
    cache of k_quartet (<= 128 reads), at 400x it does not and the keys are enumerated by passes over the window."""
    sites = synth.make_sites(910, 60_000, mean_gap=9.0)
    b = synth.make_reads(911, sites, 60_000, coverage, tid=0, nocall=0.08, del_frac=0.08)
    res, _ = parity.check_all([b], [60_000], ("pm", "me"), pm=dict(min_depth=2), me=dict(min_depth=2))
    assert res["pm"]["n"] > 1000 and res["me"]["n"] == res["pm"]["n"]


@pytest.mark.gpu
def test_chain_of_large_device_contigs_is_processed_behind_the_next_ingest(monkeypatch):
    """Four device-resident contigs of > 2^20 reads each: every one is a region of its own, read in place, and its site
    dictionary / measure kernels / row emission are queued behind the NEXT contigs' ingest passes (engine.cu mth_submit).
    Rows and LPMD counters must equal (a) the same regions processed strictly one after the other (METHEOR_NO_PIPELINE) and
    (b) the same reads fed as host batches, which are checked against the oracle."""
    torch = pytest.importorskip("torch")
    from metheor_b200 import synth_gpu as G
    lens = [22_000_000, 21_000_000, 23_000_000, 400_000, 20_500_000, 21_500_000]
    dev_batches = [G.make_contig("cuda:0", 77, tid, lens[tid], 8.0 if tid != 3 else 6.0, nocall=0.04) for tid in range(6)]
    # contig 3 is a small one: it ends the chain (0, 1, 2 are parked in turn) and is packed with 4 and 5 into one arena region
    assert sum(b["n_reads"] >= (1 << 20) for b in dev_batches) == 5
    host_batches = [G.to_numpy_batch(b) for b in dev_batches]
    want, _ = parity.check_all(host_batches, lens, ALL, pm=dict(min_depth=3), me=dict(min_depth=3))
    over = dict(pm=dict(min_depth=3), me=dict(min_depth=3))
    got, st = engine.run_batches(dev_batches, lens, ALL, **over)
    monkeypatch.setenv("METHEOR_NO_PIPELINE", "1")
    serial, st2 = engine.run_batches(dev_batches, lens, ALL, **over)
    assert st["h2d_bytes"] == 0 and st2["h2d_bytes"] == 0 and st["n_regions"] == st2["n_regions"] >= 4
    for other in (serial, want):
        for m in got:
            for k in got[m]:
                if isinstance(got[m][k], dict):
                    continue
                assert np.array_equal(np.asarray(got[m][k]), np.asarray(other[m][k]), equal_nan=True), (m, k)
