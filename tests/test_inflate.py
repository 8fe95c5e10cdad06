"""The host's own raw-DEFLATE decoder (metheor_b200/host/inflate_fast.cpp) against zlib: every block type, compression
level and strategy, literal- and match-heavy data, exact-size / truncation / corruption rejects, and the BGZF reader's
zlib fall-back counter staying at zero on well-formed files.  CPU only."""
import os
import zlib

import numpy as np
import pytest

from metheor_b200 import host, synth, synth_bam
from metheor_b200 import batch as B


def _raw(data, level, strategy=zlib.Z_DEFAULT_STRATEGY, mem=8):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, mem, strategy)
    return c.compress(data) + c.flush()


def _payloads():
    rng = np.random.default_rng(11)
    out = [b"", b"a", b"abc" * 5, bytes(65280), os.urandom(20000), (b"ACGTTTGACCA" * 50 + os.urandom(30)) * 40]
    for n in (1, 7, 8, 9, 257, 258, 259, 300, 4096, 65280, 65535):
        out.append(bytes(rng.integers(0, 4, n, dtype=np.uint8) + 65))                      # 2-bit alphabet: literal pairs
        out.append(bytes(np.minimum(rng.geometric(0.08, n), 255).astype(np.uint8)))        # skewed: long codewords, subtables
        out.append(bytes(np.minimum(rng.geometric(0.5, n), 255).astype(np.uint8)))
        out.append(bytes(rng.integers(0, 256, n, dtype=np.uint8) & rng.integers(0, 256, n, dtype=np.uint8)))
    return out


def test_every_block_type_level_and_strategy_matches_zlib():
    n = 0
    for d in _payloads():
        for level in (0, 1, 4, 6, 9):
            for strat in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
                r = _raw(d, level, strat)
                assert host.inflate_raw(r, len(d)) == d, (len(d), level, strat)
                # the decoder must produce EXACTLY the announced size from EXACTLY this input
                assert host.inflate_raw(r, len(d) + 1) is None
                if d:
                    assert host.inflate_raw(r, len(d) - 1) is None
                if len(r) > 2:
                    assert host.inflate_raw(r[:-1], len(d)) is None
                n += 1
    assert n > 1000


def test_crc32_matches_zlib_at_every_length_class():
    rng = np.random.default_rng(2)
    for n in list(range(0, 200)) + [255, 256, 1023, 4096, 65279, 65280, 65535, 100003]:
        d = bytes(rng.integers(0, 256, n, dtype=np.uint8))
        assert host.crc32(d) == (zlib.crc32(d) & 0xFFFFFFFF), n


def test_multi_block_streams_and_sync_flushes():
    rng = np.random.default_rng(5)
    parts = [bytes(rng.integers(65, 70, 3000, dtype=np.uint8)), os.urandom(500), b"x" * 4000, b""]
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    r = b""
    for i, p in enumerate(parts):  # Z_SYNC_FLUSH / Z_FULL_FLUSH end a block and insert an empty stored block
        r += c.compress(p) + c.flush(zlib.Z_SYNC_FLUSH if i % 2 else zlib.Z_FULL_FLUSH)
    r += c.flush()
    want = b"".join(parts)
    assert host.inflate_raw(r, len(want)) == want


def test_corrupted_streams_never_decode_to_wrong_size_or_crash():
    rng = np.random.default_rng(9)
    d = bytes(np.minimum(rng.geometric(0.2, 30000), 255).astype(np.uint8))
    r = bytearray(_raw(d, 6))
    for _ in range(400):
        m = bytearray(r)
        for _ in range(int(rng.integers(1, 4))):
            m[int(rng.integers(0, len(m)))] ^= 1 << int(rng.integers(0, 8))
        got = host.inflate_raw(bytes(m), len(d))
        assert got is None or len(got) == len(d)  # a flipped literal may still be a well-formed stream; the CRC catches it
    for bad in (b"\x07", b"\x06", b"\x01\x05\x00\x00\x00", b"\x01\x05\x00\xfa\xffabc"):  # reserved type, bad / short stored
        assert host.inflate_raw(bad, 5) is None


def test_bgzf_reader_uses_the_fast_decoder_and_matches_zlib_path(tmp_path):
    b, _ = synth.chr19_like(coverage=0.2, length=2_000_000, seed=77)
    sub = B.slice_reads(b, 0, min(20000, b["n_reads"]))
    p = str(tmp_path / "s.bam")
    synth_bam.write_bam(p, [("chr19", 2_000_000)], [sub], threads=4)
    before = host.zlib_fallbacks()
    a = host.decode_file(p)
    assert host.zlib_fallbacks() == before, "well-formed BGZF members must not fall back to zlib"
    assert a["n_reads"] == sub["n_reads"] and np.array_equal(a["start"], sub["start"])
    # flip one payload bit of one member: the CRC check (or the decoder) must reject it, zlib gets the last word, error
    raw = bytearray(open(p, "rb").read())
    raw[len(raw) // 2] ^= 0x10
    q = str(tmp_path / "bad.bam")
    open(q, "wb").write(bytes(raw))
    with pytest.raises(host.HostError):
        host.decode_file(q)
    assert host.zlib_fallbacks() > before
