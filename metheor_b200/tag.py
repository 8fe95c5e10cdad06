"""ctypes mirror of the engine's `tag` ABI (include/metheor_b200.h: mth_genome_*, mth_tag) — test / tooling plumbing.

The product host is metheor_b200/host/tag.cpp (`metheor tag`); this module lets the tests drive the same C entry points
with numpy arrays.  Mirrors determine_xm_tag_string's inputs (reference src/tag.rs:130-136): a record's position, CIGAR,
SEQ and strand decision, and the genome."""
import ctypes as C

import numpy as np

from . import _lib
from .engine import EngineError

NT16 = "=ACMGRSVTWYHKDBN"
_CODE = np.full(256, 15, np.uint8)
for _i, _c in enumerate(NT16):
    _CODE[ord(_c)] = _i
    _CODE[ord(_c.lower())] = _i
for _c, _v in zip("0123", (1, 2, 4, 8)):  # htslib seq_nt16_table
    _CODE[ord(_c)] = _v
CIGAR_OPS = "MIDNSHP=X"


def pack_seq(text):
    """SAM SEQ text -> BAM 4-bit bytes (high nibble first), l_seq."""
    if text == "*":
        return np.zeros(0, np.uint8), 0
    codes = _CODE[np.frombuffer(text.encode(), np.uint8)]
    n = len(codes)
    if n & 1:
        codes = np.append(codes, np.uint8(0))
    return ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8), n


def need_reverse_complement(flag, paired):
    rev, first, last = bool(flag & 16), bool(flag & 64), bool(flag & 128)
    if paired:
        return not ((not rev and first) or (rev and last))
    return rev


class Genome:
    def __init__(self, ref_len, device=0):
        self._L = _lib.lib()
        self._g = C.c_void_p()
        arr = (C.c_int64 * max(len(ref_len), 1))(*ref_len)
        rc = self._L.mth_genome_create(C.byref(self._g), device, len(ref_len), arr)
        if rc != 0:
            raise EngineError(rc, self._L.mth_genome_last_error(None).decode())

    def _check(self, rc):
        if rc != 0:
            raise EngineError(rc, self._L.mth_genome_last_error(self._g).decode())

    def set_contig(self, tid, seq):
        """seq: bytes / str / uint8 array, any letter case."""
        if isinstance(seq, str):
            seq = seq.encode()
        a = np.frombuffer(seq, np.uint8) if isinstance(seq, (bytes, bytearray)) else np.ascontiguousarray(seq, np.uint8)
        self._check(self._L.mth_genome_set_contig(self._g, tid, a.ctypes.data if a.size else None, a.size))

    def tag(self, reads, paired=False):
        """reads: iterable of dicts(tid, pos (0-based), flag, cigar [(len, op_char)], seq (SAM text)) -> (xm list, status array)."""
        reads = list(reads)
        n = len(reads)
        tid = np.array([r["tid"] for r in reads], np.int32)
        pos = np.array([r["pos"] for r in reads], np.int32)
        rcf = np.array([need_reverse_complement(r["flag"], paired) for r in reads], np.uint8)
        cig_off = np.zeros(n + 1, np.uint32)
        seq_off = np.zeros(n + 1, np.uint64)
        l_seq = np.zeros(n, np.int32)
        cig, seqs = [], []
        for i, r in enumerate(reads):
            cig += [(ln << 4) | CIGAR_OPS.index(op) for ln, op in r["cigar"]]
            cig_off[i + 1] = len(cig)
            s4, ls = pack_seq(r["seq"])
            seqs.append(s4)
            l_seq[i] = ls
            seq_off[i + 1] = seq_off[i] + np.uint64(len(s4))
        cigar = np.array(cig, np.uint32)
        seq4 = np.concatenate(seqs) if seqs else np.zeros(0, np.uint8)
        return self.tag_arrays(tid, pos, rcf, l_seq, cig_off, cigar, seq_off, seq4)

    def tag_arrays(self, tid, pos, rc, l_seq, cigar_off, cigar, seq_off, seq4, raw=False):
        """raw=True: (xm_off, xm_len, xm bytes, status) as numpy views of the engine's pinned result buffers."""
        keep = [np.ascontiguousarray(x, d) for x, d in ((tid, np.int32), (pos, np.int32), (rc, np.uint8), (l_seq, np.int32),
                (cigar_off, np.uint32), (cigar, np.uint32), (seq_off, np.uint64), (seq4, np.uint8))]
        b = _lib.TagBatch()
        b.n_reads = len(keep[0])
        for name, a in zip(("tid", "pos", "rc", "l_seq", "cigar_off", "cigar", "seq_off", "seq4"), keep):
            setattr(b, name, a.ctypes.data if a.size else None)
        res = _lib.TagResult()
        self._check(self._L.mth_tag(self._g, C.byref(b), C.byref(res)))
        n = res.n_reads
        if n == 0:
            return [], np.zeros(0, np.uint8)
        off = np.ctypeslib.as_array(C.cast(res.xm_off, C.POINTER(C.c_uint64)), (n + 1,))
        ln = np.ctypeslib.as_array(C.cast(res.xm_len, C.POINTER(C.c_uint32)), (n,))
        status = np.ctypeslib.as_array(C.cast(res.status, C.POINTER(C.c_uint8)), (n,)).copy()
        total = int(off[n])
        if raw:
            return off, ln, np.ctypeslib.as_array(C.cast(res.xm, C.POINTER(C.c_uint8)), (max(total, 1),))[:total], status
        xm = bytes(np.ctypeslib.as_array(C.cast(res.xm, C.POINTER(C.c_uint8)), (max(total, 1),))[:total])
        out = [xm[int(off[i]):int(off[i]) + int(ln[i])].decode() for i in range(n)]
        assert int(res.n_failed) == int((status != 0).sum())
        return out, status

    def last_kernel_ms(self):
        return float(self._L.mth_genome_last_kernel_ms(self._g))

    def close(self):
        if self._g:
            self._L.mth_genome_destroy(self._g)
            self._g = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
