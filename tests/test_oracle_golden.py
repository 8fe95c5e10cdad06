"""Pins the CPU oracle against every value the reference's own unit tests hold for the hot path
(values + file:line in tests/golden/reference_pins.json)."""
import math
import os
import subprocess

import numpy as np
import pytest

import bamio
import oracle_lib
from oracle_lib import Oracle


def f32(x):
    return np.float32(x)


def test_pdr_pins(fixture_bams, pins):
    a = pins["pdr"]["args"]
    for c in pins["pdr"]["cases"]:
        r = Oracle.open(fixture_bams[c["input"]]).pdr(a["min_depth"], c["min_cpgs"], a["min_qual"])
        assert len(r["pos"]) == c["n_rows"], c
        if c["n_rows"]:
            assert (r["pdr"] == f32(c["pdr"])).all() and (r["n_conc"] == c["n_conc"]).all() and (r["n_disc"] == c["n_disc"]).all(), c


def test_readutil_discordant_reads(fixture_bams, pins):
    # readutil.rs:420-440: 16 reads, 14 discordant  ==> every CpG of test1 sees n_disc 14 / n_conc 2 with no filters
    r = Oracle.open(fixture_bams["test1"]).pdr(0, 0, 0)
    assert (r["n_disc"] == pins["readutil"]["n_discordant_read"]).all()
    assert (r["n_conc"] + r["n_disc"] == pins["readutil"]["n_read"]).all()


def test_lpmd_pins(fixture_bams, pins):
    a = pins["lpmd"]["args"]
    for c in pins["lpmd"]["cases"]:
        v = Oracle.open(fixture_bams[c["input"]]).lpmd(a["min_distance"], a["max_distance"], a["min_qual"])["lpmd"]
        if c["lpmd"] == "NaN":
            assert math.isnan(v)
        else:
            assert v == f32(c["lpmd"]), c


def test_mhl_pins(fixture_bams, pins):
    a = pins["mhl"]["args"]
    for c in pins["mhl"]["cases"]:
        r = Oracle.open(fixture_bams[c["input"]]).mhl(a["min_depth"], a["min_cpgs"], a["min_qual"])
        assert len(r["pos"]) == c["n_rows"], c
        if c["n_rows"]:
            assert (r["value"] == f32(c["mhl"])).all(), (c, r["value"])


def test_pm_me_pins(fixture_bams, pins):
    for c in pins["pm"]["cases"]:
        r = Oracle.open(fixture_bams[c["input"]]).quartets(0, pins["pm"]["args"]["min_qual"])
        assert len(r["pm"]) == c["n_quartets"]
        if c["n_quartets"]:
            assert (r["pm"] == f32(c["pm"])).all(), (c, r["pm"])
    for c in pins["me"]["cases"]:
        r = Oracle.open(fixture_bams[c["input"]]).quartets(0, pins["me"]["args"]["min_qual"])
        assert len(r["me"]) == c["n_quartets"]
        if c["n_quartets"]:
            assert (r["me"] == f32(c["me"])).all(), (c, r["me"])
            if "depth" in c:
                assert (r["counts"].sum(axis=1) == c["depth"]).all()


def _val(c, key):
    if key + "_f32_of" in c:
        a, b = c[key + "_f32_of"].split("/")
        return f32(a) / f32(b)
    return f32(c[key])


def test_fdrp_qfdrp_pins(fixture_bams, pins):
    for name, quant in (("fdrp", False), ("qfdrp", True)):
        a = pins[name]["args"]
        for c in pins[name]["cases"]:
            r = Oracle.open(fixture_bams[c["input"]]).fdrp(min_qual=c["min_qual"], min_depth=a["min_depth"],
                                                           max_depth=a["max_depth"], min_overlap=a["min_overlap"],
                                                           quantitative=quant)
            assert list(r["pos"]) == c["positions"], c
            if not c["positions"]:
                continue
            want = _val(c, name)
            if c["exact"]:
                assert (r["value"] == want).all(), (c, r["value"])
            else:
                assert (np.abs(r["value"] - want) < c["tol"]).all(), (c, r["value"])


def test_qfdrp_pair_primitives(golden, pins):
    # qfdrp.rs:291-305 hamming(0,k); :332-355 shared CpGs(0,1) == 4; :309-330 pile size 16.
    # A two-read pile {0,k} with min_overlap 1 gives qfdrp = hamming/shared (denominator 1 pair).
    reads = golden["test1"]["reads"]
    for k, ham in enumerate(pins["qfdrp"]["hamming_0_k"]["values"], start=1):
        sub = [reads[0], reads[k]]
        soa = _soa_from_records(sub)
        v = Oracle.from_soa(**soa).qfdrp(min_qual=0, min_depth=2, max_depth=40, min_overlap=1)["value"]
        assert (v == f32(ham) / f32(pins["qfdrp"]["shared_cpgs_0_1"]["value"])).all()
    # pile size 16: FDRP denominator is C(16,2)=120 -> with all pairs discordant except identical ones
    full = Oracle.from_soa(**_soa_from_records(reads)).fdrp(min_qual=0, min_depth=16, max_depth=40, min_overlap=1)
    assert len(full["pos"]) == 4
    assert len(Oracle.from_soa(**_soa_from_records(reads)).fdrp(min_qual=0, min_depth=17, max_depth=40, min_overlap=1)["pos"]) == 0


def _soa_from_records(reads):
    """Plain-Python decode of pure-M forward records (enough for the fixtures)."""
    tid, start, end, mapq, off, pos, rel, meth = [], [], [], [], [0], [], [], []
    for r in reads:
        n = int(r["cigar"][:-1]); assert r["cigar"].endswith("M") and r["flag"] == 0
        tid.append(r["tid"]); start.append(r["pos"]); end.append(r["pos"] + n - 1); mapq.append(r["mapq"])
        for i, ch in enumerate(r["xm"]):
            if ch in "zZ":
                pos.append(r["pos"] + i); rel.append(i); meth.append(ch == "Z")
        off.append(len(pos))
    return dict(tid=tid, start=start, end=end, mapq=mapq, cpg_off=off, cpg_pos=pos, cpg_rel=rel, cpg_meth=meth)


def test_soa_entry_equals_bam_entry(fixture_bams, golden):
    for name in ("test1", "test4", "test6"):
        a = Oracle.open(fixture_bams[name])
        b = Oracle.from_soa(**_soa_from_records(golden[name]["reads"]))
        for fn in (lambda o: o.pdr(0, 0, 10), lambda o: o.mhl(0, 0, 10), lambda o: o.fdrp(min_depth=1, min_overlap=1),
                   lambda o: o.quartets(0, 10)):
            ra, rb = fn(a), fn(b)
            for k in ra:
                assert np.array_equal(ra[k], rb[k], equal_nan=True), (name, k)


def test_cli_default_rows(fixture_bams, tmp_path):
    """default-flag TSV rows on the fixtures, through the oracle CLI."""
    oracle_lib.build()

    def run(measure, name, *extra):
        out = tmp_path / f"{measure}_{name}.tsv"
        subprocess.check_call([oracle_lib.CLI_PATH, measure, "-i", fixture_bams[name], "-o", str(out), *extra])
        return out.read_text()

    assert run("pdr", "test1") == "".join(f"chr1\t{p}\t{p + 2}\t0.875\t2\t14\n" for p in (0, 2, 4, 6))
    assert run("pdr", "test2") == "".join(f"chr1\t{p}\t{p + 2}\t0\t16\t0\n" for p in (0, 2, 4, 6))
    assert run("pdr", "test3") == ""
    assert run("pdr", "test6") == ""
    assert run("mhl", "test4") == "".join(f"chr1\t{p}\t{p + 2}\t0.1625\n" for p in (0, 2, 4, 6, 13, 15, 17, 19))
    assert run("pm", "test1") == "chr1\t0\t2\t4\t6\t0.9375\n"
    assert run("me", "test1") == "chr1\t0\t2\t4\t6\t1\n"
    assert run("me", "test2") == "chr1\t0\t2\t4\t6\t0.25\n"
    assert run("lpmd", "test5").splitlines()[1].endswith("\tNaN")
    assert run("lpmd", "test1").splitlines()[0] == "name\tlpmd"
    assert run("fdrp", "test6") == "chr1\t2\t4\t0\nchr1\t13\t15\t0\n"
    assert run("fdrp", "test1", "-d", "1", "-D", "100", "-l", "1", "-q", "10") == "".join(
        f"chr1\t{p}\t{p + 2}\t1\n" for p in (0, 2, 4, 6))
    assert run("qfdrp", "test1", "-d", "1", "-D", "100", "-l", "1", "-q", "10") == "".join(
        f"chr1\t{p}\t{p + 2}\t0.53333336\n" for p in (0, 2, 4, 6))
    pairs = tmp_path / "pairs.tsv"
    run("lpmd", "test1", "-p", str(pairs))
    assert pairs.read_text() == "chrom\tcpg1\tcpg2\tlpmd\tn_concordant\tn_discordant\n" + "".join(
        f"chr1\t{a}\t{b}\t0.5\t8\t8\n" for a, b in ((0, 2), (0, 4), (0, 6), (2, 4), (2, 6), (4, 6)))


def test_rust_display_f32():
    for v, s in ((0.875, "0.875"), (1.0, "1"), (0.0, "0"), (-0.0, "-0"), (8 / 15, "0.53333336"), (0.1625, "0.1625"),
                 (1e-10, "0.0000000001"), (1 / 3, "0.33333334"), (float("nan"), "NaN"), (float("inf"), "inf")):
        assert oracle_lib.fmt_f32(v) == s


@pytest.mark.skipif(not os.path.exists("/root/reference/tests/test1.bam"), reason="reference tree not mounted")
def test_oracle_reads_original_reference_files(golden):
    """Only where /root/reference exists: the oracle's own BGZF/BAM/SAM parser on the untouched reference inputs
    decodes exactly the records committed in fixtures.json (and the rebuilt BAMs are therefore faithful)."""
    for k in range(1, 7):
        o = Oracle.open(f"/root/reference/tests/test{k}.bam")
        want = Oracle.from_soa(**_soa_from_records(golden[f"test{k}"]["reads"])).export_reads()
        got = o.export_reads()
        for key in ("start", "end", "mapq", "cpg_off", "cpg_pos", "cpg_rel", "cpg_meth"):
            assert np.array_equal(got[key], want[key]), (k, key)
    o = Oracle.open("/root/reference/tests/test.chr19.XM.sam")
    assert lib_n(o) == 1000 and o.all_xm_ok()


def lib_n(o):
    return oracle_lib.lib().orc_n_reads(o.h)
