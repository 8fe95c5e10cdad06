"""GPU: the device-side BGZF inflate and BAM record decode (csrc/inflate.cuh, csrc/bamdec.cu) against zlib and against the
C++ host decoder (which the CPU suite checks against the oracle's decoder, tests/test_host_decode.py)."""
import os
import zlib

import numpy as np
import pytest

import bamio
import recgen
from metheor_b200 import bamdec, host, synth

pytestmark = pytest.mark.gpu

REFS = [("chr1", 200_000), ("chr2", 90_000), ("chrM", 16_569)]


def _raw(data, level, strategy=zlib.Z_DEFAULT_STRATEGY):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    return c.compress(data) + c.flush()


def test_inflate_kernel_matches_zlib():
    rng = np.random.default_rng(5)
    text = (b"ACGTTGCANNNN" * 50 + bytes(rng.integers(33, 74, 700, dtype=np.uint8))) * 60
    cases = [b"", b"a", b"abc" * 5, bytes(rng.integers(0, 256, 65000, dtype=np.uint8)),  # incompressible: stored blocks
             text[:65000], text[:300], bytes(60000), bytes(rng.integers(0, 4, 64000, dtype=np.uint8)),
             bytes(rng.choice(np.frombuffer(b"zZ.hHxX", np.uint8), 65280))]
    blob, members, want = bytearray(), [], []
    for data in cases:
        for level, strat in ((1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED),
                             (0, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)):
            raw = _raw(data, level, strat)
            members.append((len(blob), len(raw), len(data), 0, zlib.crc32(data)))
            blob += raw
            want.append(data)
            blob += b"\0" * int(rng.integers(0, 5))  # unaligned payload starts
    got, status, ms = bamdec.inflate_members(bytes(blob), members)
    assert not status.any(), status
    assert got == b"".join(want)
    # corrupt streams are reported, never crash: flip bytes / truncate
    bad = bytearray(blob)
    for k in range(40, len(bad), 97):
        bad[k] ^= 0x5A
    _, status, _ = bamdec.inflate_members(bytes(bad), members)
    assert status.any()
    # a wrong CRC-32 in the trailer is reported for exactly that member (status 9), the others stay clean
    wrong = list(members)
    wrong[5] = wrong[5][:4] + (wrong[5][4] ^ 0x1,)
    got2, status, _ = bamdec.inflate_members(bytes(blob), wrong)
    assert status[5] == 9 and not np.delete(status, 5).any() and got2 == got


def _decode_all(path, window_members, **kw):
    data = open(path, "rb").read()
    members = bamdec.bgzf_members(data)
    # the header: inflate from the start with zlib until it is complete (what the host does)
    refs, hdr_len = bamio.read_header_len(path)
    dec = bamdec.Decoder([l for _, l in refs], **kw)
    out, tot = [], dict(n_records=0, n_dropped=0, n_dropped_mapq_ok=0)
    try:
        k = 0
        first = True
        while k < len(members):
            w = members[k:k + window_members]
            k += len(w)
            res, batches = dec.window(data, w, skip=hdr_len if first else 0, last=(k >= len(members)))
            first = False
            assert res.bad_record < 0
            out += batches
            for f in tot:
                tot[f] += getattr(res, f)
    finally:
        dec.close()
    return out, tot


def _compare_with_host(path, batches, tot, min_qual=0):
    d = host.decode_file(path)
    keep = np.diff(d["cpg_off"]) > 0
    assert sum(b["n_reads"] for b in batches) == int(keep.sum())
    assert tot["n_records"] == d["n_reads"] and tot["n_dropped"] == int((~keep).sum())
    assert tot["n_dropped_mapq_ok"] == int(((~keep) & (d["mapq"] >= min_qual)).sum())
    cat = lambda k: np.concatenate([b[k] for b in batches]) if batches else np.zeros(0)
    assert np.array_equal(np.concatenate([np.full(b["n_reads"], b["tid"], np.int32) for b in batches]), d["tid"][keep])
    assert np.array_equal(cat("start"), d["start"][keep]) and np.array_equal(cat("end"), d["end"][keep])
    assert np.array_equal(cat("meta") & 0x1FF, d["meta"][keep] & 0x1FF)
    assert np.array_equal(np.concatenate([np.diff(b["cpg_off"].astype(np.int64)) for b in batches]), np.diff(d["cpg_off"])[keep])
    assert all(b["cpg_off"][0] == 0 and b["cpg_off"][-1] == b["n_cpg"] for b in batches)
    assert np.array_equal(cat("cpg_pos"), d["cpg_pos"]) and np.array_equal(cat("cpg_rel"), d["cpg_rel"])
    from metheor_b200 import batch as B
    assert np.array_equal(np.concatenate([B.unpack_meth(b) for b in batches]), d["cpg_meth"])


@pytest.mark.parametrize("block,window", [(300, 1), (3000, 2), (3000, 1000), (60000, 3)])
def test_device_decode_equals_host_decoder(tmp_path, block, window):
    """Random records (indels, clips, skips, both strands, odd flags, three contigs + unmapped tail), small BGZF blocks so that
    records straddle members and windows: device SoA == host SoA."""
    reads = recgen.random_records(101, REFS, 4000, max_len=90)
    reads += [dict(tid=-1, pos=-1, flag=4, mapq=0, cigar="", xm="....z...Z.") for _ in range(30)]
    path = str(tmp_path / "r.bam")
    bamio.write_bam(path, REFS, reads, block=block)
    batches, tot = _decode_all(path, window)
    assert len({b["tid"] for b in batches}) == 3
    _compare_with_host(path, batches, tot)


def test_device_decode_synthetic_wgbs_and_missing_xm(tmp_path):
    length = 150_000
    sites = synth.make_sites(171, length)
    b = synth.make_reads(172, sites, length, 25.0)
    path = str(tmp_path / "s.bam")
    bamio.write_bam(path, [("chr19", length)], recgen.batch_to_records(b))
    batches, tot = _decode_all(path, 4)
    _compare_with_host(path, batches, tot)
    # a record without XM: reported with its index (the host then aborts like readutil.rs:45-51)
    reads = recgen.random_records(33, REFS, 500)
    reads[123]["xm"] = None
    path2 = str(tmp_path / "noxm.bam")
    bamio.write_bam(path2, REFS, reads)
    data = open(path2, "rb").read()
    refs, hdr_len = bamio.read_header_len(path2)
    dec = bamdec.Decoder([l for _, l in refs])
    res, _ = dec.window(data, bamdec.bgzf_members(data), skip=hdr_len, last=True)
    assert res.bad_record == 123 and not res.bad_is_corrupt
    dec.close()
    # LPMD order: a low-mapq record without XM is only counted (lpmd.rs:176-181)
    reads[123]["mapq"] = 0
    bamio.write_bam(path2, REFS, reads)
    data = open(path2, "rb").read()
    dec = bamdec.Decoder([l for _, l in refs], lpmd_order=True, min_qual=10)
    res, _ = dec.window(data, bamdec.bgzf_members(data), skip=hdr_len, last=True)
    assert res.bad_record < 0
    dec.close()


def test_records_longer_than_a_chain_chunk(tmp_path):
    """Records far larger than the 32 KiB chunks of the speculative boundary search (CIGARs with 12 000 operations: ~48 KB each):
    chunks without any record start, chains that jump over chunks, tiny windows that end inside such a record."""
    rng = np.random.default_rng(9)
    reads = recgen.random_records(55, REFS, 600, max_len=60)
    big = []
    for k in range(6):
        ops = "".join(f"{int(rng.integers(1, 3))}M1I" for _ in range(6000))
        qlen = sum(int(x) for x in ops.replace("M", " ").replace("I", " ").split())
        xm = "".join(rng.choice(list("zZ" + "." * 40 + "h" * 20), min(qlen, 900)))  # ~30 calls: within the device decoder's 64
        big.append(dict(tid=0, pos=int(1000 + 20_000 * k), flag=0, mapq=40, cigar=ops, xm=xm))
    allr = sorted([r for r in reads if r["tid"] >= 0] + big, key=lambda r: (r["tid"], r["pos"]))
    path = str(tmp_path / "big.bam")
    bamio.write_bam(path, REFS, allr, block=40_000)
    for window in (1, 2, 1000):
        batches, tot = _decode_all(path, window)
        _compare_with_host(path, batches, tot)
