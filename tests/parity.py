"""Shared helpers of the GPU parity tests: run the engine through the C ABI and the oracle on the same reads."""
import numpy as np

from metheor_b200 import batch as B
from metheor_b200 import engine
from oracle_lib import Oracle


def bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


def assert_site_rows(got, want, what, value_key="value"):
    assert got["n"] == len(want["pos"]), f"{what}: {got['n']} rows vs oracle {len(want['pos'])}"
    assert np.array_equal(got["tid"], want["tid"]), f"{what}: tid"
    assert np.array_equal(got["pos"], want["pos"]), f"{what}: pos"
    gb, wb = bits(got["value"]), bits(want[value_key])
    if not np.array_equal(gb, wb):
        bad = np.flatnonzero(gb != wb)
        raise AssertionError(f"{what}: {len(bad)} of {len(gb)} values differ bitwise; first at row {bad[0]} pos "
                             f"{want['pos'][bad[0]]}: got {got['value'][bad[0]]!r} want {want[value_key][bad[0]]!r}")


def assert_quartet_rows(got, want, what, value_key):
    assert got["n"] == len(want["p1"]), f"{what}: {got['n']} rows vs oracle {len(want['p1'])}"
    for k in ("tid", "p1", "p2", "p3", "p4"):
        assert np.array_equal(got[k], want[k]), f"{what}: {k}"
    gb, wb = bits(got["value"]), bits(want[value_key])
    if not np.array_equal(gb, wb):
        bad = np.flatnonzero(gb != wb)
        raise AssertionError(f"{what}: {len(bad)} of {len(gb)} values differ; first row {bad[0]} got {got['value'][bad[0]]!r} "
                             f"want {want[value_key][bad[0]]!r}")
    if "counts" in got:
        assert np.array_equal(got["counts"], want["counts"]), f"{what}: counts"


DEFAULTS = dict(pdr=dict(min_depth=10, min_cpgs=4, min_qual=10), mhl=dict(min_depth=10, min_cpgs=4, min_qual=10),
                pm=dict(min_depth=10, min_qual=10), me=dict(min_depth=10, min_qual=10),
                fdrp=dict(min_qual=10, min_depth=10, max_depth=40, min_overlap=35),
                qfdrp=dict(min_qual=10, min_depth=10, max_depth=40, min_overlap=35),
                lpmd=dict(min_distance=2, max_distance=16, min_qual=10))


def check_all(batches, ref_len, measures, seed=0, flags=0, compact=False, cpg_set=None, **overrides):
    """Engine vs oracle, bit-exact, for every requested measure.  Returns (engine results, stats).
    cpg_set = (tid array, pos array): the engine filters on the device (mth_set_cpg_set), the oracle with filter_isin."""
    prm = {m: dict(DEFAULTS[m], **overrides.get(m, {})) for m in measures}
    res, stats = engine.run_batches(batches, ref_len, measures, flags=flags, seed=seed, compact=compact, cpg_set=cpg_set,
                                    **{k: dict(v) for k, v in prm.items()})
    orc = Oracle.from_soa(**B.to_oracle_soa(batches))
    if cpg_set is not None:
        orc.set_cpg_set(*cpg_set)
    if "pdr" in measures:
        w = orc.pdr(**prm["pdr"])
        assert_site_rows(res["pdr"], w, "pdr", "pdr")
        assert np.array_equal(res["pdr"]["n_conc"], w["n_conc"]) and np.array_equal(res["pdr"]["n_disc"], w["n_disc"]), "pdr counts"
    if "mhl" in measures:
        assert_site_rows(res["mhl"], orc.mhl(**prm["mhl"]), "mhl")
    for m in ("fdrp", "qfdrp"):
        if m in measures:
            assert_site_rows(res[m], orc.fdrp(seed=seed, quantitative=(m == "qfdrp"), **prm[m]), m)
    for m in ("pm", "me"):
        if m in measures:
            assert_quartet_rows(res[m], orc.quartets(**prm[m]), m, m)
    if "lpmd" in measures:
        want_pairs = bool(prm["lpmd"].pop("want_pairs", 0))
        w = orc.lpmd(pairs=want_pairs, **prm["lpmd"])
        g = res["lpmd"]
        if want_pairs:
            gp, wp = g["pairs"], w["pairs"]
            assert gp["n"] == len(wp["tid"]), f"lpmd pairs: {gp['n']} rows vs oracle {len(wp['tid'])}"
            for k in ("tid", "pos1", "pos2", "n_conc", "n_disc"):
                assert np.array_equal(gp[k], wp[k]), f"lpmd pairs: {k}"
            assert np.array_equal(bits(gp["lpmd"]), bits(wp["lpmd"])), "lpmd pairs: value"
        assert (g["n_read"], g["n_valid_read"], g["n_conc"], g["n_disc"]) == (w["n_read"], w["n_valid_read"], w["n_conc"], w["n_disc"]), (g, w)
        assert bits(g["lpmd"]) == bits(w["lpmd"]) or (np.isnan(g["lpmd"]) and np.isnan(w["lpmd"]))
    return res, stats


def records_to_batch(reads, tid=0):
    """Plain-Python decode of fixture records (any CIGAR, both strands) into one SoA batch — mirrors
    readutil.rs:24-53,323-345 for the tests (the product decoder is the C++ host)."""
    start, end, meta, off, pos, rel, meth = [], [], [], [0], [], [], []
    import bamio
    for r in reads:
        ref, qi, positions = r["pos"], 0, []
        for n, op in bamio.parse_cigar(r["cigar"]):
            if op in (0, 7, 8):
                positions += list(range(ref, ref + n)); ref += n
            elif op in (1, 4):
                positions += [None] * n
            elif op in (2, 3):
                ref += n
        al = [p for p in positions if p is not None]
        fwd = r["flag"] in (0, 99, 147)
        start.append(al[0] if al else -1); end.append(al[-1] if al else -1)
        meta.append(r["mapq"] | (int(fwd) << 8))
        for i, (p, ch) in enumerate(zip(positions, r["xm"])):
            if ch in "zZ" and p is not None:
                pos.append(p if fwd else p - 1); rel.append(i); meth.append(ch == "Z")
        off.append(len(pos))
    m, moff = B.pack_meth(off, np.asarray(meth, np.uint8))
    return dict(tid=tid, n_reads=len(reads), n_cpg=len(pos), start=np.asarray(start, np.int32), end=np.asarray(end, np.int32),
                meta=np.asarray(meta, np.uint32), cpg_off=np.asarray(off, np.uint32), cpg_pos=np.asarray(pos, np.int32),
                cpg_rel=np.asarray(rel, np.uint16), meth=m, meth_off=moff)
