"""CPU tests of the C++ host (no GPU): BGZF/BAM/SAM decode -> BismarkRead-level SoA against the oracle's decoder, the
--cpg-set filter, Rust-compatible f32 formatting, the C API surface and the clap-compatible command line."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import bamio
import recgen
from metheor_b200 import host
from oracle_lib import Oracle, fmt_f32

REFS = [("chr1", 200_000), ("chr2", 90_000), ("chrM", 16_569)]


def _same_decode(path, cpg_set=None):
    got = host.decode_file(path, cpg_set=cpg_set, threads=3)
    o = Oracle.open(path)
    want = o.export_reads()
    assert got["refs"] == o.refs()
    assert got["n_reads"] == len(want["tid"])
    for k in ("start", "end", "mapq"):
        assert np.array_equal(got[k], want[k]), k
    has = want["tid"] >= 0  # the oracle's BismarkRead only knows its contig through its CpGs (readutil.rs:15-21)
    if cpg_set is None:
        assert np.array_equal(got["tid"][has], want["tid"][has])
    if cpg_set is None:
        assert np.array_equal(got["cpg_off"], want["cpg_off"])
        assert np.array_equal(got["cpg_pos"], want["cpg_pos"])
        assert np.array_equal(got["cpg_rel"].astype(np.int32), want["cpg_rel"])
        assert np.array_equal(got["cpg_meth"], want["cpg_meth"])
    return got, want


def test_exports_and_header():
    L = host.lib()
    for s in host.EXPORTS:
        assert hasattr(L, s), s
    hdr = open(os.path.join(os.path.dirname(host.HERE), "include", "metheor_host.h")).read()
    declared = set(re.findall(r"\b(mthh_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(host.EXPORTS), declared ^ set(host.EXPORTS)


@pytest.mark.parametrize("name", ["test1", "test2", "test3", "test4", "test5", "test6"])
def test_reference_fixture_bams(fixture_bams, name):
    got, _ = _same_decode(fixture_bams[name])
    assert got["n_reads"] in (16, 32)


def test_chr19_sam_and_bam(golden, tmp_path):
    fx = golden["chr19_1000"]
    refs = [tuple(r) for r in fx["refs"]]
    sam, bam = str(tmp_path / "c.sam"), str(tmp_path / "c.bam")
    bamio.write_sam(sam, refs, fx["reads"])
    bamio.write_bam(bam, refs, fx["reads"], with_seq=True)
    a, _ = _same_decode(sam)
    b, _ = _same_decode(bam)
    assert a["n_reads"] == 1000 and np.array_equal(a["cpg_pos"], b["cpg_pos"]) and a["n_cpg"] > 1000


@pytest.mark.parametrize("seed,block", [(1, 0xFF00), (2, 700), (3, 61)])
def test_random_cigars_strands_and_small_bgzf_blocks(tmp_path, seed, block):
    """Records straddle BGZF members (tiny blocks) and windows; indels, clips, ref-skips, odd flags, short XM."""
    reads = recgen.random_records(seed, REFS, 3000)
    p = str(tmp_path / "r.bam")
    bamio.write_bam(p, REFS, reads, with_seq=(seed == 1), block=block)
    got, _ = _same_decode(p)
    assert got["n_cpg"] > 5000 and (np.diff(got["cpg_off"]) == 0).any()
    s = str(tmp_path / "r.sam")
    bamio.write_sam(s, REFS, reads)
    _same_decode(s)


def test_cpg_set_filter(tmp_path):
    reads = recgen.random_records(5, REFS, 2000)
    p = str(tmp_path / "r.bam")
    bamio.write_bam(p, REFS, reads)
    full = host.decode_file(p)
    rng = np.random.default_rng(0)
    ridx = np.repeat(np.arange(full["n_reads"]), np.diff(full["cpg_off"]))
    keep = rng.random(full["n_cpg"]) < 0.5
    pairs = sorted(set(zip(full["tid"][ridx][keep].tolist(), full["cpg_pos"][keep].tolist())))
    bed = str(tmp_path / "set.bed")
    with open(bed, "w") as f:
        for t, x in pairs:
            f.write(f"{REFS[t][0]}\t{x}\t{x + 2}\n")
    got = host.decode_file(p, cpg_set=bed)
    inset = np.array([(t, x) in set(pairs) for t, x in zip(full["tid"][ridx].tolist(), full["cpg_pos"].tolist())])
    assert np.array_equal(got["cpg_pos"], full["cpg_pos"][inset])
    assert np.array_equal(got["cpg_rel"], full["cpg_rel"][inset])
    assert np.array_equal(np.diff(got["cpg_off"]), np.bincount(ridx[inset], minlength=full["n_reads"]))
    # the oracle applies the same set per measure: spot-check through PDR-independent decode counts
    o = Oracle.open(p)
    o.set_cpg_set_file(bed)
    with pytest.raises(host.HostError) as e:
        host.decode_file(p, cpg_set=str(tmp_path / "missing.bed"))
    assert e.value.status == 101 and "Could not read target CpG file" in e.value.msg
    with open(bed, "a") as f:
        f.write("chrUn\t5\t7\n")
    with pytest.raises(host.HostError):
        host.decode_file(p, cpg_set=bed)  # unknown contig: header.tid(chrom).unwrap() panics (bamutil.rs:24)


def test_input_errors(tmp_path):
    with pytest.raises(host.HostError) as e:
        host.decode_file(str(tmp_path / "no_such.bam"))
    assert e.value.status == 101 and "Error opening BAM file" in e.value.msg and "file not found" in e.value.msg \
        and "no_such.bam" in e.value.msg
    junk = tmp_path / "Cargo.toml"
    junk.write_text("[package]\nname = \"metheor\"\n")
    with pytest.raises(host.HostError) as e:
        host.decode_file(str(junk))
    assert e.value.status == 101 and "Error opening BAM file" in e.value.msg
    reads = recgen.random_records(7, REFS, 300)
    reads[100]["xm"] = None
    p = str(tmp_path / "noxm.bam")
    bamio.write_bam(p, REFS, reads)
    with pytest.raises(host.HostError) as e:
        host.decode_file(p)
    assert e.value.status == 101 and "Error reading XM tag in BAM record" in e.value.msg
    raw = open(p, "rb").read()
    t = tmp_path / "trunc.bam"
    t.write_bytes(raw[: len(raw) // 2])
    with pytest.raises(host.HostError):
        host.decode_file(str(t))


def test_f32_formatting_matches_rust_display():
    vals = [0.875, 1.0, 0.0, -0.0, 0.53333336, 1e-10, 0.1625, 1 / 3, 0.9375, 0.25, float("nan"), float("inf"), 2.5e-7, 123456.7]
    want = ["0.875", "1", "0", "-0", "0.53333336", "0.0000000001", "0.1625", "0.33333334", "0.9375", "0.25", "NaN", "inf",
            "0.00000025", "123456.7"]
    assert [host.format_f32(v) for v in vals] == want
    rng = np.random.default_rng(1)
    for v in rng.random(2000).astype(np.float32):
        s = host.format_f32(float(v))
        assert s == fmt_f32(float(v)) and np.float32(s) == v and "e" not in s


def test_cli_usage_matches_clap_contract(tmp_path):
    """tests/cli_error_handling.rs of the reference, against our binary (none of these reach the GPU)."""
    r = host.cli("--help")
    assert r.returncode == 0 and all(w in r.stdout for w in ("Usage:", "Commands:", "pdr", "fdrp", "tag"))
    r = host.cli("--version")
    assert r.returncode == 0 and "metheor" in r.stdout
    r = host.cli()
    assert r.returncode != 0 and "Usage:" in r.stderr
    r = host.cli("invalid_command")
    assert r.returncode != 0 and "error:" in r.stderr and "subcommand" in r.stderr
    for sub in ("pdr", "lpmd", "mhl", "pm", "me", "fdrp", "qfdrp"):
        r = host.cli(sub, "--output", "x.tsv")
        assert r.returncode == 2 and "required" in r.stderr
        r = host.cli(sub, "--input", "tests/test1.bam")
        assert r.returncode == 2 and "required" in r.stderr
    r = host.cli("pdr", "--input", "a.bam", "--output", "x.tsv", "--min-depth", "-5")
    assert r.returncode != 0 and ("invalid" in r.stderr or "error" in r.stderr)
    r = host.cli("pdr", "--input", "a.bam", "--output", "x.tsv", "--min-qual", "300")
    assert r.returncode != 0
    r = host.cli("tag", "--input", "a.bam", "--output", "x.sam")
    assert r.returncode != 0 and "required" in r.stderr
    r = host.cli("pdr", "--help")
    assert r.returncode == 0 and all(w in r.stdout for w in ("PDR", "--input", "--output"))
    r = host.cli("lpmd", "--help")
    assert r.returncode == 0 and all(w in r.stdout for w in ("LPMD", "--min-distance", "--max-distance"))
    r = host.cli("pdr", "-i", str(tmp_path / "no_such.bam"), "-o", str(tmp_path / "o.tsv"))
    assert r.returncode == 101 and "file not found" in r.stderr and "no_such.bam" in r.stderr
    junk = tmp_path / "Cargo.toml"
    junk.write_text("[package]\n")
    r = host.cli("pdr", "-i", str(junk), "-o", str(tmp_path / "o.tsv"))
    assert r.returncode == 101 and "Error opening BAM file" in r.stderr


def test_multi_window_stream_matches_generator(tmp_path):
    """~18 MB of records = several 8 MiB windows of the decode-only API: records straddle window and BGZF member
    boundaries, the producer thread inflates and walks ahead of the decoder."""
    from metheor_b200 import synth, synth_bam
    L = 200_000
    sites = synth.make_sites(91, L)
    b = synth.make_reads(92, sites, L, 30.0)
    p = str(tmp_path / "big.bam")
    info = synth_bam.write_bam(p, [("chr19", L)], [b], threads=4)
    assert info["bytes_uncompressed"] > 17_000_000
    for threads in (1, 5):
        d = host.decode_file(p, threads=threads)
        assert d["n_reads"] == b["n_reads"]
        for k in ("start", "end", "cpg_pos", "cpg_rel"):
            assert np.array_equal(d[k], b[k]), k
        assert np.array_equal(d["mapq"], (b["meta"] & 0xFF).astype(np.uint8))
