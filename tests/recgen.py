"""Seeded generator of alignment records (tid/pos/flag/mapq/CIGAR/XM) that exercises what the reference's fixtures
do not: both strands and odd flags, indels, clips, ref-skips, =/X ops, no-calls, XM shorter than the read."""
import numpy as np


def random_records(seed, refs, n, max_len=60, p_simple=0.6, xm_missing=False):
    rng = np.random.default_rng(seed)
    reads = []
    pos_by_tid = {}
    for tid, (_, ln) in enumerate(refs):
        k = n // len(refs)
        pos_by_tid[tid] = np.sort(rng.integers(0, max(1, ln - 6000), k))  # N ops can stretch a read by a few kb
    flags = [0, 16, 99, 147, 83, 163, 1024, 256, 0, 16]
    for tid in sorted(pos_by_tid):
        for pos in pos_by_tid[tid]:
            qlen = int(rng.integers(8, max_len))
            if rng.random() < p_simple:
                ops = [(qlen, "M")]
            else:
                ops, left = [], qlen
                if rng.random() < 0.3:
                    ops.append((int(rng.integers(1, 4)), "H"))
                if rng.random() < 0.4:
                    s = int(rng.integers(1, 5)); ops.append((s, "S")); left -= s
                while left > 0:
                    m = int(min(left, rng.integers(1, 20)))
                    ops.append((m, str(rng.choice(["M", "M", "M", "=", "X"])))); left -= m
                    if left <= 0:
                        break
                    c = rng.random()
                    if c < 0.3:
                        i = int(min(left, rng.integers(1, 4))); ops.append((i, "I")); left -= i
                    elif c < 0.6:
                        ops.append((int(rng.integers(1, 30)), "D"))
                    elif c < 0.7:
                        ops.append((int(rng.integers(50, 400)), "N"))
                    elif c < 0.75:
                        ops.append((1, "P"))
                if ops[-1][1] in "DNP":
                    ops.append((1, "M")); qlen += 1
                if rng.random() < 0.3:
                    s = int(rng.integers(1, 4)); ops.append((s, "S")); qlen += s
                qlen = sum(l for l, o in ops if o in "MIS=X")
            xm_len = qlen if rng.random() < 0.95 else max(1, qlen - int(rng.integers(1, 5)))
            xm = "".join(rng.choice(list("....zZzZhHxXuU"), xm_len))
            reads.append(dict(tid=tid, pos=int(pos), flag=int(rng.choice(flags)), mapq=int(rng.choice([0, 5, 10, 30, 42])),
                              cigar="".join(f"{l}{o}" for l, o in ops), xm=None if xm_missing and rng.random() < 0.01 else xm))
    return reads


def batch_to_records(b, read_len=150):
    """SoA batch of metheor_b200.synth (no deletions) -> alignment records with `{read_len}M` CIGARs and XM strings."""
    off = np.asarray(b["cpg_off"], np.int64)
    from metheor_b200.batch import unpack_meth
    meth = unpack_meth(b)
    recs = []
    for r in range(b["n_reads"]):
        xm = ["."] * read_len
        for k in range(off[r], off[r + 1]):
            xm[int(b["cpg_rel"][k])] = "Z" if meth[k] else "z"
        fwd = (int(b["meta"][r]) >> 8) & 1
        recs.append(dict(tid=int(b["tid"]), pos=int(b["start"][r]), flag=0 if fwd else 16, mapq=int(b["meta"][r]) & 0xFF,
                         cigar=f"{read_len}M", xm="".join(xm)))
    return recs
