"""Helpers around the SoA read batch (the layout of `mth_batch` in include/metheor_b200.h), numpy side."""
import numpy as np

FIELDS = ("start", "end", "meta", "cpg_off", "cpg_pos", "cpg_rel", "meth", "meth_off")


def meth_word_offsets(b):
    if b.get("meth_off") is not None:
        return np.asarray(b["meth_off"], np.int64)
    return np.arange(b["n_reads"] + 1, dtype=np.int64)


def unpack_meth(b):
    """-> uint8[n_cpg] methylation flag per CpG incidence."""
    off = np.asarray(b["cpg_off"], np.int64)
    cnt = np.diff(off)
    ridx = np.repeat(np.arange(b["n_reads"], dtype=np.int64), cnt)
    k = np.arange(int(off[-1]), dtype=np.int64) - off[ridx]
    moff = meth_word_offsets(b)
    w = np.asarray(b["meth"], np.uint64)[moff[ridx] + k // 64]
    return ((w >> (k % 64).astype(np.uint64)) & np.uint64(1)).astype(np.uint8)


def pack_meth(cpg_off, flags):
    off = np.asarray(cpg_off, np.int64)
    cnt = np.diff(off)
    R = len(cnt)
    words = np.maximum(1, (cnt + 63) // 64)
    moff = np.zeros(R + 1, np.int64)
    np.cumsum(words, out=moff[1:])
    ridx = np.repeat(np.arange(R, dtype=np.int64), cnt)
    k = np.arange(int(off[-1]), dtype=np.int64) - off[ridx]
    meth = np.zeros(int(moff[-1]), np.uint64)
    np.bitwise_or.at(meth, moff[ridx] + k // 64, np.asarray(flags, np.uint64) << (k % 64).astype(np.uint64))
    return meth, (None if int(words.max(initial=1)) == 1 else moff.astype(np.uint32))


def select_reads(b, mask):
    """Sub-batch keeping reads where mask is True (file order preserved)."""
    mask = np.asarray(mask, bool)
    off = np.asarray(b["cpg_off"], np.int64)
    cnt = np.diff(off)
    keep_inc = np.repeat(mask, cnt)
    flags = unpack_meth(b)[keep_inc]
    new_off = np.zeros(int(mask.sum()) + 1, np.int64)
    np.cumsum(cnt[mask], out=new_off[1:])
    meth, moff = pack_meth(new_off, flags)
    return dict(tid=b["tid"], n_reads=int(mask.sum()), n_cpg=int(new_off[-1]), start=b["start"][mask], end=b["end"][mask],
                meta=b["meta"][mask], cpg_off=new_off.astype(np.uint32), cpg_pos=b["cpg_pos"][keep_inc],
                cpg_rel=None if b.get("cpg_rel") is None else b["cpg_rel"][keep_inc], meth=meth, meth_off=moff)


def slice_reads(b, lo, hi):
    m = np.zeros(b["n_reads"], bool)
    m[lo:hi] = True
    return select_reads(b, m)


def slice_range(b, lo, hi):
    """Reads [lo, hi) of a batch without touching the rest of it (array slices; methylation words are per read)."""
    off = np.asarray(b["cpg_off"], np.int64)
    i0, i1 = int(off[lo]), int(off[hi])
    moff = b.get("meth_off")
    if moff is None:
        meth, mo = b["meth"][lo:hi], None
    else:
        m = np.asarray(moff, np.int64)
        meth, mo = b["meth"][int(m[lo]):int(m[hi])], (m[lo:hi + 1] - m[lo]).astype(np.uint32)
    return dict(tid=b["tid"], n_reads=hi - lo, n_cpg=i1 - i0, start=b["start"][lo:hi], end=b["end"][lo:hi], meta=b["meta"][lo:hi],
                cpg_off=(off[lo:hi + 1] - i0).astype(np.uint32), cpg_pos=b["cpg_pos"][i0:i1],
                cpg_rel=None if b.get("cpg_rel") is None else b["cpg_rel"][i0:i1], meth=meth, meth_off=mo)


def to_oracle_soa(batches):
    """Concatenate batches (in file order) into the argument dict of tests/oracle_lib.Oracle.from_soa."""
    tid, start, end, mapq, pos, rel, meth, offs = [], [], [], [], [], [], [], [np.zeros(1, np.int64)]
    base = 0
    for b in batches:
        tid.append(np.full(b["n_reads"], b["tid"], np.int32))
        start.append(b["start"]); end.append(b["end"]); mapq.append((b["meta"] & 0xFF).astype(np.uint8))
        pos.append(b["cpg_pos"]); meth.append(unpack_meth(b))
        rel.append(np.asarray(b["cpg_rel"], np.int32) if b.get("cpg_rel") is not None
                   else (np.arange(b["n_cpg"], dtype=np.int64) - np.repeat(np.asarray(b["cpg_off"], np.int64)[:-1],
                                                                            np.diff(np.asarray(b["cpg_off"], np.int64)))).astype(np.int32))
        offs.append(np.asarray(b["cpg_off"], np.int64)[1:] + base)
        base += b["n_cpg"]
    cat = np.concatenate
    return dict(tid=cat(tid), start=cat(start), end=cat(end), mapq=cat(mapq), cpg_off=cat(offs), cpg_pos=cat(pos),
                cpg_rel=cat(rel), cpg_meth=cat(meth))


CBLOCK = 256  # MTH_CBLOCK


def to_compact(b, with_rel=True, dense=False):
    """SoA batch -> the compact wire format of include/metheor_b200.h (mth_batch_compact).  Needs <= 64 calls per read.
    dense=True adds the block encodings MTH_CENC_START16 | MTH_CENC_DELTA8 (7 B per read + 1.125 B per call)."""
    off = np.asarray(b["cpg_off"], np.int64)
    cnt = np.diff(off)
    if cnt.max(initial=0) > 64:
        raise ValueError("compact wire format holds at most 64 CpG calls per read")
    R = b["n_reads"]
    start = np.asarray(b["start"], np.int64)
    span = np.asarray(b["end"], np.int64) - start
    if span.max(initial=0) > 65023 or span.min(initial=0) < 0:
        raise ValueError("span outside the compact format")
    meta = np.asarray(b["meta"], np.uint32)
    fwd = ((meta >> 8) & 1).astype(np.int64)
    ridx = np.repeat(np.arange(R, dtype=np.int64), cnt)
    pos = np.asarray(b["cpg_pos"], np.int64)
    delta = pos - (start[ridx] - 1)
    flags = (fwd | (((meta >> 9) & 1) << 1)).astype(np.uint8)
    rel_exc = np.zeros(0, np.uint16)
    if with_rel and b.get("cpg_rel") is not None:
        rel = np.asarray(b["cpg_rel"], np.int64)
        odd = rel != delta - fwd[ridx]
        explicit = np.zeros(R, bool)
        explicit[np.unique(ridx[odd])] = True
        flags = flags | (explicit.astype(np.uint8) << 2)
        rel_exc = rel[explicit[ridx]].astype(np.uint16)
    out = dict(tid=b["tid"], n_reads=R, n_cpg=int(off[-1]), n_rel=len(rel_exc), start=np.asarray(b["start"], np.int32),
               span=span.astype(np.uint16), mapq=(meta & 0xFF).astype(np.uint8), n_cpg8=cnt.astype(np.uint8), flags=flags,
               cpg_delta=delta.astype(np.uint16), meth_bits=np.packbits(unpack_meth(b), bitorder="little"), rel_exc=rel_exc, enc=0)
    if not dense or R == 0:
        return out
    nb = (R + CBLOCK - 1) // CBLOCK
    blk = np.arange(R, dtype=np.int64) // CBLOCK
    # starts: 16-bit offsets from the block's first (= smallest) start; blocks spanning more than 65535 keep 32-bit starts
    first = start[np.arange(nb, dtype=np.int64) * CBLOCK]
    last = start[np.minimum((np.arange(nb, dtype=np.int64) + 1) * CBLOCK, R) - 1]
    wide_s = (last - first) > 65535
    exc_rank = np.cumsum(wide_s) - 1
    blk_start = np.where(wide_s, -(1 + exc_rank), first).astype(np.int32)
    off16 = np.where(wide_s[blk], 0, start - first[blk]).astype(np.uint16)
    start_exc = np.zeros(int(wide_s.sum()) * CBLOCK, np.int32)
    for e, bidx in enumerate(np.flatnonzero(wide_s)):
        seg = start[bidx * CBLOCK:(bidx + 1) * CBLOCK]
        start_exc[e * CBLOCK:e * CBLOCK + len(seg)] = seg
    # calls: deltas from the previous call of the read (first call: from start - 1); a block with a delta > 255 stays 16-bit
    is_first = np.zeros(len(pos), bool)
    is_first[off[:-1][cnt > 0]] = True
    prev = np.where(is_first, start[ridx] - 1, np.concatenate([[0], pos[:-1]]))
    dch = pos - prev
    cblk = blk[ridx]
    blk_max = np.zeros(nb, np.int64)
    np.maximum.at(blk_max, cblk, dch)
    wide_c = blk_max > 255
    ncall_blk = np.bincount(cblk, minlength=nb)
    n8 = np.where(wide_c, 0, ncall_blk)
    n16 = np.where(wide_c, ncall_blk, 0)
    off8 = np.cumsum(n8) - n8
    off16c = np.cumsum(n16) - n16
    blk_call_off = np.where(wide_c, off16c | 0x80000000, off8).astype(np.uint32)
    out.update(enc=3, start_off16=off16, blk_start=blk_start, start_exc=start_exc, n_start_exc=len(start_exc),
               cpg_delta8=dch[~wide_c[cblk]].astype(np.uint8), cpg_delta=dch[wide_c[cblk]].astype(np.uint16), blk_call_off=blk_call_off)
    out["n_delta8"], out["n_delta16"] = len(out["cpg_delta8"]), len(out["cpg_delta"])
    del out["start"]
    return out
