"""Vectorised writer of a realistic Bismark-like BAM for the synthetic reads of synth.py (plain `<len>M` alignments):
~450 bytes per record (24-byte name, 4-bit SEQ, QUAL, XM:Z, XR:Z, XG:Z, NM:i), BGZF members of <= 65280 bytes compressed
with zlib level 1 on a thread pool.  Only used to time the end-to-end BAM path (bench.py --bam-reads) and in tests."""
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .batch import unpack_meth


def _bgzf_block(data, level=1):
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    comp = c.compress(data) + c.flush()
    bsize = len(comp) + 25
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + comp +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def write_bam(path, refs, batches, read_len=150, seed=0, threads=16, block=0xFF00, with_xm=True):
    """batches: SoA dicts (one contig each, ascending tid) whose reads are plain `read_len`M alignments.
    with_xm=False leaves the XM:Z field out (input of `metheor tag`)."""
    rng = np.random.default_rng(seed)
    ht = ("@HD\tVN:1.0\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in refs)).encode()
    head = bytearray(b"BAM\x01" + struct.pack("<i", len(ht)) + ht + struct.pack("<i", len(refs)))
    for n, l in refs:
        nb = n.encode() + b"\0"
        head += struct.pack("<i", len(nb)) + nb + struct.pack("<i", l)
    name_len = 24
    n_seq = (read_len + 1) // 2
    o_name = 36
    o_cigar = o_name + name_len
    o_seq = o_cigar + 4
    o_qual = o_seq + n_seq
    o_xm = o_qual + read_len
    o_tail = o_xm + (3 + read_len + 1 if with_xm else 0)
    tail = b"XRZCT\0XGZCT\0NMC\x03"
    rec_len = o_tail + len(tail)
    chunks = [bytes(head)]
    n_total = 0
    for b in batches:
        R = b["n_reads"]
        rec = np.zeros((R, rec_len), np.uint8)
        core = np.zeros(R, dtype=[("bs", "<i4"), ("tid", "<i4"), ("pos", "<i4"), ("lrn", "u1"), ("mapq", "u1"), ("bin", "<u2"),
                                  ("ncig", "<u2"), ("flag", "<u2"), ("lseq", "<i4"), ("ntid", "<i4"), ("npos", "<i4"), ("tlen", "<i4")])
        core["bs"] = rec_len - 4
        core["tid"] = b["tid"]
        core["pos"] = b["start"]
        core["lrn"] = name_len
        core["mapq"] = b["meta"] & 0xFF
        core["bin"] = 4680
        core["ncig"] = 1
        core["flag"] = np.where((b["meta"] >> 8) & 1, 0, 16)
        core["lseq"] = read_len
        core["ntid"] = -1
        core["npos"] = -1
        rec[:, :36] = core.view(np.uint8).reshape(R, 36)
        name = np.frombuffer(b"SRR0000000.", np.uint8)
        rec[:, o_name:o_name + len(name)] = name
        idx = np.arange(n_total, n_total + R, dtype=np.int64)
        for d in range(12):  # 12 decimal digits of the read index
            rec[:, o_name + len(name) + 11 - d] = 48 + (idx // 10 ** d) % 10
        rec[:, o_cigar:o_cigar + 4] = np.frombuffer(struct.pack("<I", read_len << 4), np.uint8)
        rec[:, o_seq:o_seq + n_seq] = rng.choice(np.array([0x11, 0x12, 0x14, 0x18, 0x21, 0x28, 0x41, 0x48, 0x81, 0x88], np.uint8), (R, n_seq))
        rec[:, o_qual:o_qual + read_len] = rng.integers(28, 41, (R, read_len), dtype=np.uint8)
        if not with_xm:
            rec[:, o_tail:] = np.frombuffer(tail, np.uint8)
            chunks.append(rec.tobytes())
            n_total += R
            continue
        rec[:, o_xm:o_xm + 3] = np.frombuffer(b"XMZ", np.uint8)
        xm = rec[:, o_xm + 3:o_xm + 3 + read_len]
        xm[:] = ord(".")
        other = rng.random((R, read_len)) < 0.18  # CHH / CHG context calls
        xm[other] = rng.choice(np.frombuffer(b"hhhhxxHX", np.uint8), int(other.sum()))
        off = np.asarray(b["cpg_off"], np.int64)
        ridx = np.repeat(np.arange(R, dtype=np.int64), np.diff(off))
        xm[ridx, np.asarray(b["cpg_rel"], np.int64)] = np.where(unpack_meth(b) > 0, ord("Z"), ord("z"))
        non_cpg = np.ones((R, read_len), bool)
        non_cpg[ridx, np.asarray(b["cpg_rel"], np.int64)] = False
        zz = non_cpg & ((xm == ord("z")) | (xm == ord("Z")))
        xm[zz] = ord(".")
        rec[:, o_tail:] = np.frombuffer(tail, np.uint8)
        chunks.append(rec.tobytes())
        n_total += R
    raw = b"".join(chunks)
    pieces = [raw[i:i + block] for i in range(0, len(raw), block)]
    with ThreadPoolExecutor(threads) as ex:
        comp = list(ex.map(_bgzf_block, pieces))
    with open(path, "wb") as f:
        for c in comp:
            f.write(c)
        f.write(_bgzf_block(b""))
    return dict(records=n_total, bytes_uncompressed=len(raw), bytes_compressed=sum(map(len, comp)) + 28, record_bytes=rec_len)
