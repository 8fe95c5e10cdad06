"""GPU: the `metheor` binary (C++ host + CUDA engine) against the oracle's CLI — TSV files byte for byte — on the
reference's fixture BAMs, on random records (indels, clips, both strands, three contigs) and on synthetic WGBS reads."""
import filecmp
import json
import os
import subprocess

import numpy as np
import pytest

import bamio
import oracle_lib
import recgen
from metheor_b200 import host, synth

pytestmark = pytest.mark.gpu

MEASURES = ("pdr", "lpmd", "mhl", "pm", "me", "fdrp", "qfdrp")
REFS = [("chr1", 200_000), ("chr2", 90_000), ("chrM", 16_569)]


def _oracle_cli(*args):
    oracle_lib.build()
    return subprocess.run([oracle_lib.CLI_PATH, *map(str, args)], capture_output=True, text=True)


_FLAG_FIELD = {"-d": "min_depth", "-p": "min_cpgs", "-q": "min_qual", "-D": "max_depth", "-l": "min_overlap", "-m": "min_distance",
               "-M": "max_distance", "-c": "cpg_set", "--seed": "seed"}


def _both(tmp_path, measure, bam, *flags, expect_rows=None, binary=False):
    """Engine vs oracle CLI on the same file.  binary=True spawns `metheor`; otherwise the same code runs through the
    library entry point mthh_run (what the binary calls), which saves a process + CUDA start-up per case."""
    a, b = str(tmp_path / f"{measure}.engine.tsv"), str(tmp_path / f"{measure}.oracle.tsv")
    if binary:
        r = host.cli(measure, "-i", bam, "-o", a, *flags)
        assert r.returncode == 0, r.stderr
    else:
        kw = {_FLAG_FIELD[flags[i]]: (flags[i + 1] if flags[i] == "-c" else int(flags[i + 1])) for i in range(0, len(flags), 2)}
        host.run(measure, bam, a, **kw)
    o = _oracle_cli(measure, "-i", bam, "-o", b, *flags)
    assert o.returncode == 0, o.stderr
    ta, tb = open(a).read(), open(b).read()
    assert ta == tb, f"{measure} {flags}: TSV differs\n--- engine\n{ta[:400]}\n--- oracle\n{tb[:400]}"
    if expect_rows is not None:
        assert ta.count("\n") == expect_rows, (measure, ta.count("\n"))
    return ta


@pytest.mark.parametrize("name", ["test1", "test2", "test3", "test4", "test5", "test6"])
def test_fixture_bams_default_flags(fixture_bams, tmp_path, name):
    for m in MEASURES:
        _both(tmp_path, m, fixture_bams[name], binary=(name == "test1"))


def test_fixture_expected_rows_from_survey_appendix_b(fixture_bams, tmp_path):
    """SURVEY.md Appendix B (derived from the reference's unit-test pins)."""
    t = _both(tmp_path, "pdr", fixture_bams["test1"], expect_rows=4)
    assert t.splitlines()[0] == "chr1\t0\t2\t0.875\t2\t14"
    assert _both(tmp_path, "pm", fixture_bams["test1"]) == "chr1\t0\t2\t4\t6\t0.9375\n"
    assert _both(tmp_path, "me", fixture_bams["test1"]) == "chr1\t0\t2\t4\t6\t1\n"
    assert _both(tmp_path, "mhl", fixture_bams["test4"], expect_rows=8).splitlines()[4] == "chr1\t13\t15\t0.1625"
    assert _both(tmp_path, "lpmd", fixture_bams["test5"]).splitlines()[1].endswith("\tNaN")
    # tests/output_validation.rs:304-405 flags
    t = _both(tmp_path, "qfdrp", fixture_bams["test1"], "-d", 1, "-D", 100, "-l", 1, "-q", 10, expect_rows=4)
    assert t.splitlines()[0] == "chr1\t0\t2\t0.53333336"
    t = _both(tmp_path, "fdrp", fixture_bams["test1"], "-d", 1, "-D", 100, "-l", 1, "-q", 10, expect_rows=4)
    assert t.splitlines()[0] == "chr1\t0\t2\t1"
    assert _both(tmp_path, "pdr", fixture_bams["test3"]) == ""  # the output file exists and is empty (pdr.rs:95-101)


def test_random_records_three_contigs(tmp_path):
    reads = recgen.random_records(21, REFS, 6000, max_len=80)
    bam = str(tmp_path / "r.bam")
    bamio.write_bam(bam, REFS, reads, block=3000)
    for m, flags in (("pdr", ("-d", 2, "-p", 2)), ("mhl", ("-d", 2, "-p", 2)), ("pm", ("-d", 1)), ("me", ("-d", 2)),
                     ("fdrp", ("-d", 2, "-l", 5)), ("qfdrp", ("-d", 2, "-l", 5, "-D", 8, "--seed", 5)), ("lpmd", ("-m", 1, "-M", 30))):
        t = _both(tmp_path, m, bam, *flags)
        assert t.count("\n") > (1 if m == "lpmd" else 20)
    pa, pb = str(tmp_path / "pairs.engine.tsv"), str(tmp_path / "pairs.oracle.tsv")
    assert host.cli("lpmd", "-i", bam, "-o", str(tmp_path / "l.tsv"), "-p", pa, "-m", 1, "-M", 30).returncode == 0
    assert _oracle_cli("lpmd", "-i", bam, "-o", str(tmp_path / "l2.tsv"), "-p", pb, "-m", 1, "-M", 30).returncode == 0
    assert open(pa).read() == open(pb).read() and open(pa).read().count("\n") > 100
    sam = str(tmp_path / "r.sam")
    bamio.write_sam(sam, REFS, reads)
    _both(tmp_path, "pdr", sam, "-d", 2, "-p", 2, binary=True)


def test_synthetic_wgbs_default_flags_and_cpg_set(tmp_path):
    length = 120_000
    sites = synth.make_sites(71, length)
    b = synth.make_reads(72, sites, length, 30.0)
    refs = [("chr19", length)]
    bam = str(tmp_path / "s.bam")
    bamio.write_bam(bam, refs, recgen.batch_to_records(b))
    for m in MEASURES:
        t = _both(tmp_path, m, bam)
        assert t.count("\n") > (1 if m == "lpmd" else 100), m
    bed = str(tmp_path / "set.bed")
    with open(bed, "w") as f:
        for x in sites[::2]:
            f.write(f"chr19\t{x}\t{x + 2}\n")
    for m in MEASURES:
        _both(tmp_path, m, bam, "-c", bed, *(() if m == "lpmd" else ("-d", 5)))
    stats = str(tmp_path / "stats.json")
    r = host.cli("pdr", "-i", bam, "-o", str(tmp_path / "x.tsv"), "--stats", stats, "--threads", 4)
    assert r.returncode == 0, r.stderr
    st = json.load(open(stats))
    assert st["records"] == b["n_reads"] and st["gpu"][0]["kernel_launches"] > 0 and st["seconds"]["total"] > 0


def test_missing_xm_aborts_like_the_reference(tmp_path):
    reads = recgen.random_records(31, REFS, 300)
    reads[50]["xm"] = None
    reads[50]["mapq"] = 0
    bam = str(tmp_path / "noxm.bam")
    bamio.write_bam(bam, REFS, reads)
    r = host.cli("pdr", "-i", bam, "-o", str(tmp_path / "o.tsv"))
    assert r.returncode == 101 and "Error reading XM tag in BAM record" in r.stderr
    # LPMD tests mapq first (lpmd.rs:177): the low-mapq record without XM is skipped
    r = host.cli("lpmd", "-i", bam, "-o", str(tmp_path / "o.tsv"))
    assert r.returncode == 0, r.stderr
    r = host.cli("pdr", "-i", bam.replace("noxm", "nodir/none"), "-o", str(tmp_path / "o.tsv"))
    assert r.returncode == 101
    ok = str(tmp_path / "ok.bam")
    bamio.write_bam(ok, REFS, recgen.random_records(32, REFS, 300))
    r = host.cli("pdr", "-i", ok, "-o", "/nonexistent_directory/readonly_output.tsv")
    assert r.returncode != 0


def test_edge_inputs(tmp_path):
    """Header-only BAM, unmapped reads behind the mapped ones, a CpG set that filters everything, reads at contig ends."""
    empty = str(tmp_path / "empty.bam")
    bamio.write_bam(empty, REFS, [])
    for m in MEASURES:
        t = _both(tmp_path, m, empty)
        assert t == ("name\tlpmd\n" + empty + "\tNaN\n" if m == "lpmd" else "")
    reads = recgen.random_records(41, REFS, 1500, p_simple=0.8)
    reads += [dict(tid=-1, pos=-1, flag=4, mapq=0, cigar="", xm="....z...Z.") for _ in range(20)]
    # reads touching both ends of a contig (forward and reverse strand)
    edge = [dict(tid=2, pos=0, flag=f, mapq=42, cigar="20M", xm="Z..z....Z.....z....Z") for f in (0, 16)] + \
           [dict(tid=2, pos=REFS[2][1] - 20, flag=f, mapq=42, cigar="20M", xm="Z..z....Z.....z....Z") for f in (0, 16)]
    mapped = [r for r in reads if r["tid"] >= 0]
    mapped = sorted(mapped + edge, key=lambda r: (r["tid"], r["pos"]))
    bam = str(tmp_path / "u.bam")
    bamio.write_bam(bam, REFS, mapped + [r for r in reads if r["tid"] < 0])
    for m in MEASURES:
        _both(tmp_path, m, bam, *(() if m == "lpmd" else ("-d", 1)))
    bed = str(tmp_path / "none.bed")
    open(bed, "w").write("chr1\t199999\t200001\n")
    for m in ("pdr", "lpmd", "pm"):
        _both(tmp_path, m, bam, "-c", bed)


def test_reads_with_more_than_64_calls_take_the_soa_batch(tmp_path):
    """A window that holds a read with > 64 CpG calls cannot use the compact wire format: the host falls back to the SoA
    batch with meth_off for it (DESIGN.md 'Limits': up to 256 calls per read)."""
    rng = np.random.default_rng(5)
    reads = []
    for i in range(400):
        n = 120
        xm = "".join(rng.choice(list("zZ"), n)) if i % 3 == 0 else "".join(rng.choice(list("zZ.."), n))
        reads.append(dict(tid=0, pos=int(100 + i // 4), flag=int(rng.choice([0, 16])), mapq=42, cigar=f"{n}M", xm=xm))
    bam = str(tmp_path / "dense.bam")
    bamio.write_bam(bam, REFS, reads)
    for m in MEASURES:
        t = _both(tmp_path, m, bam, *(() if m == "lpmd" else ("-d", 5)))
        assert t.count("\n") > (1 if m == "lpmd" else 50), m


def test_multi_gpu_position_bins_and_contigs(tmp_path):
    """`metheor --gpus N` (one process, N engine contexts): the genome cut into position bins with halo reads (default) or into
    whole contigs; LPMD's counters joined by the library's NCCL all-reduce.  TSVs must equal the oracle CLI byte for byte."""
    from metheor_b200 import _lib
    n_dev = _lib.lib().mth_device_count()
    if n_dev < 2:
        pytest.skip("needs at least 2 GPUs")
    n = min(n_dev, 4)
    refs = [("chrA", 150_000), ("chrB", 60_000), ("chrC", 90_000)]
    recs = []
    for tid, (_, length) in enumerate(refs):
        sites = synth.make_sites(90 + tid, length)
        b = synth.make_reads(190 + tid, sites, length, 25.0, tid=tid, del_frac=0.05)
        recs += recgen.batch_to_records(b)
    bam = str(tmp_path / "mg.bam")
    bamio.write_bam(bam, refs, recs, block=5000)
    for shard in ("bins", "contigs"):
        for m in MEASURES:
            a, b_ = str(tmp_path / f"{m}.{shard}.tsv"), str(tmp_path / f"{m}.oracle.tsv")
            r = host.cli(m, "-i", bam, "-o", a, "--gpus", n, "--shard", shard)
            assert r.returncode == 0, r.stderr
            if not os.path.exists(b_):
                o = _oracle_cli(m, "-i", bam, "-o", b_)
                assert o.returncode == 0, o.stderr
            assert open(a).read() == open(b_).read(), (m, shard)
            assert open(a).read().count("\n") > (1 if m == "lpmd" else 100)
    pa, pb = str(tmp_path / "pairs.mg.tsv"), str(tmp_path / "pairs.oracle.tsv")
    assert host.cli("lpmd", "-i", bam, "-o", str(tmp_path / "l.tsv"), "-p", pa, "--gpus", n).returncode == 0
    assert _oracle_cli("lpmd", "-i", bam, "-o", str(tmp_path / "l2.tsv"), "-p", pb).returncode == 0
    assert open(pa).read() == open(pb).read()


def test_output_formats_and_region(tmp_path):
    """Engine extensions behind the path: --format tsv.gz / bedgraph[.gz] and --region through the .bai's linear
    index.  A region run must give exactly the rows of the whole-file run whose site lies inside the region."""
    import gzip
    refs = [("chrA", 300_000), ("chrB", 120_000), ("chrC", 50_000)]
    recs = []
    for tid, (_, length) in enumerate(refs):
        if tid == 2:
            continue  # a contig without reads
        sites = synth.make_sites(700 + tid, length)
        recs += recgen.batch_to_records(synth.make_reads(710 + tid, sites, length, 20.0, tid=tid, del_frac=0.05))
    bam = str(tmp_path / "r.bam")
    block = 20_000
    rec_off = bamio.write_bam(bam, refs, recs, block=block)
    bamio.write_bai(bam, refs, recs, rec_off, block=block)
    whole = {}
    for m in ("pdr", "mhl", "fdrp", "pm"):
        whole[m] = str(tmp_path / f"{m}.tsv")
        assert host.cli(m, "-i", bam, "-o", whole[m]).returncode == 0
    # formats
    gz, bg, bgz = str(tmp_path / "p.tsv.gz"), str(tmp_path / "p.bedgraph"), str(tmp_path / "p.bedgraph.gz")
    assert host.cli("pdr", "-i", bam, "-o", gz, "--format", "tsv.gz").returncode == 0
    assert host.cli("pdr", "-i", bam, "-o", bg, "--format", "bedgraph").returncode == 0
    assert host.cli("mhl", "-i", bam, "-o", bgz, "--format", "bedgraph.gz").returncode == 0
    assert gzip.open(gz, "rb").read() == open(whole["pdr"], "rb").read()
    assert open(gz, "rb").read()[-28:] == bamio.BGZF_EOF and len(bamio.bgzf_decompress(open(gz, "rb").read())) == os.path.getsize(whole["pdr"])
    assert open(bg).read() == "".join("\t".join(l.split("\t")[:4]) + "\n" for l in open(whole["pdr"]).read().splitlines())
    assert gzip.open(bgz, "rt").read() == open(whole["mhl"]).read()
    r = host.cli("pm", "-i", bam, "-o", bg, "--format", "bedgraph")
    assert r.returncode == 2 and "per-CpG measure" in r.stderr
    # regions: inside a contig, across BGZF blocks, at the contig ends, a whole contig, an empty contig, 1-based inclusive ends
    for reg, (tid, lo, hi) in (("chrA:100001-150000", (0, 100_000, 150_000)), ("chrA:1-20000", (0, 0, 20_000)), ("chrB", (1, 0, 120_000)),
                               ("chrA:250,000-300,000", (0, 249_999, 300_000)), ("chrC", (2, 0, 50_000)), ("chrB:60000", (1, 59_999, 120_000))):
        for m in ("pdr", "mhl", "fdrp", "pm"):
            out = str(tmp_path / "reg.tsv")
            r = host.cli(m, "-i", bam, "-o", out, "--region", reg)
            assert r.returncode == 0, r.stderr
            name = refs[tid][0]
            want = [l for l in open(whole[m]).read().splitlines() if l.split("\t")[0] == name and lo <= int(l.split("\t")[1]) < hi]
            assert open(out).read().splitlines() == want, (reg, m)
        if tid != 2:
            assert len(want) > 10
    # LPMD over a region = LPMD over the reads that start inside it
    sub = [r for r in recs if r["tid"] == 0 and 100_000 <= r["pos"] < 150_000]
    bam2 = str(tmp_path / "sub.bam")
    bamio.write_bam(bam2, refs, sub)
    a, b_ = str(tmp_path / "l1.tsv"), str(tmp_path / "l2.tsv")
    assert host.cli("lpmd", "-i", bam, "-o", a, "--region", "chrA:100001-150000").returncode == 0
    assert _oracle_cli("lpmd", "-i", bam2, "-o", b_).returncode == 0
    assert open(a).read().splitlines()[1].split("\t")[1] == open(b_).read().splitlines()[1].split("\t")[1]
    assert host.cli("pdr", "-i", bam, "-o", a, "--region", "chrZ:1-5").returncode == 101
