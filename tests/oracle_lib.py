"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
CLI_PATH = os.path.join(ORACLE_DIR, "_build", "metheor_oracle")


def build():
    src = os.path.join(ORACLE_DIR, "metheor_oracle.cpp")
    if (not os.path.exists(LIB_PATH) or not os.path.exists(CLI_PATH)
            or os.path.getmtime(LIB_PATH) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp, i64, i32, u32, u64 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint32, C.c_uint64
        L.orc_open.restype = vp; L.orc_open.argtypes = [C.c_char_p]
        L.orc_from_soa.restype = vp; L.orc_from_soa.argtypes = [i64] + [vp] * 8
        L.orc_close.argtypes = [vp]
        L.orc_error.restype = C.c_char_p; L.orc_error.argtypes = [vp]
        L.orc_n_reads.restype = i64; L.orc_n_reads.argtypes = [vp]
        L.orc_n_cpgs.restype = i64; L.orc_n_cpgs.argtypes = [vp]
        L.orc_n_ref.restype = C.c_int; L.orc_n_ref.argtypes = [vp]
        L.orc_ref_name.restype = C.c_char_p; L.orc_ref_name.argtypes = [vp, C.c_int]
        L.orc_ref_len.restype = i64; L.orc_ref_len.argtypes = [vp, C.c_int]
        L.orc_all_xm_ok.restype = C.c_int; L.orc_all_xm_ok.argtypes = [vp]
        L.orc_export_reads.argtypes = [vp] * 9
        L.orc_set_cpg_set_file.restype = C.c_int; L.orc_set_cpg_set_file.argtypes = [vp, C.c_char_p]
        L.orc_set_cpg_set.argtypes = [vp, i64, vp, vp]
        L.orc_clear_cpg_set.argtypes = [vp]
        L.orc_pdr.restype = i64; L.orc_pdr.argtypes = [vp, u32, u32, u32]
        L.orc_pdr_rows.argtypes = [vp] * 6
        L.orc_mhl.restype = i64; L.orc_mhl.argtypes = [vp, u32, u32, u32]
        L.orc_fdrp.restype = i64; L.orc_fdrp.argtypes = [vp, C.c_int, u32, u32, u32, i32, u64]
        L.orc_fdrp_oob.restype = C.c_int; L.orc_fdrp_oob.argtypes = [vp]
        L.orc_site_rows.argtypes = [vp] * 4
        L.orc_quartets.restype = i64; L.orc_quartets.argtypes = [vp, u32, u32]
        L.orc_quartet_rows.argtypes = [vp] * 9
        L.orc_lpmd.restype = i64; L.orc_lpmd.argtypes = [vp, i32, i32, u32, C.c_int, vp, vp]
        L.orc_lpmd_pair_rows.argtypes = [vp] * 7
        L.orc_fmt_f32.restype = C.c_int; L.orc_fmt_f32.argtypes = [C.c_float, C.c_char_p, C.c_int]
        L.orc_reservoir_draw.restype = u32; L.orc_reservoir_draw.argtypes = [u64, i32, i32, u32]
        L.orc_compute_pm.restype = C.c_float; L.orc_compute_pm.argtypes = [vp]
        L.orc_compute_me.restype = C.c_float; L.orc_compute_me.argtypes = [vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def fmt_f32(v):
    buf = C.create_string_buffer(128)
    n = lib().orc_fmt_f32(C.c_float(v), buf, 128)
    return buf.raw[:n].decode()


class Oracle:
    """One decoded read set + the seven measures of the reference restated on the CPU."""

    def __init__(self, handle, keep=()):
        self.h = handle
        self._keep = keep
        if not handle:
            raise RuntimeError("oracle: null handle")

    @classmethod
    def open(cls, path):
        o = cls(lib().orc_open(path.encode()))
        err = lib().orc_error(o.h).decode()
        if err:
            raise IOError(err)
        return o

    @classmethod
    def from_soa(cls, tid, start, end, mapq, cpg_off, cpg_pos, cpg_rel, cpg_meth):
        arrs = [np.ascontiguousarray(tid, np.int32), np.ascontiguousarray(start, np.int32),
                np.ascontiguousarray(end, np.int32), np.ascontiguousarray(mapq, np.uint8),
                np.ascontiguousarray(cpg_off, np.int64), np.ascontiguousarray(cpg_pos, np.int32),
                None if cpg_rel is None else np.ascontiguousarray(cpg_rel, np.int32),
                np.ascontiguousarray(cpg_meth, np.uint8)]
        h = lib().orc_from_soa(len(arrs[0]), *[_p(a) for a in arrs])
        return cls(h)

    def close(self):
        if self.h:
            lib().orc_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ----- decoded reads --------------------------------------------------------------------
    def refs(self):
        L = lib()
        return [(L.orc_ref_name(self.h, i).decode(), L.orc_ref_len(self.h, i)) for i in range(L.orc_n_ref(self.h))]

    def all_xm_ok(self):
        return bool(lib().orc_all_xm_ok(self.h))

    def export_reads(self):
        L = lib()
        n, k = L.orc_n_reads(self.h), L.orc_n_cpgs(self.h)
        d = dict(tid=np.zeros(n, np.int32), start=np.zeros(n, np.int32), end=np.zeros(n, np.int32),
                 mapq=np.zeros(n, np.uint8), cpg_off=np.zeros(n + 1, np.int64), cpg_pos=np.zeros(k, np.int32),
                 cpg_rel=np.zeros(k, np.int32), cpg_meth=np.zeros(k, np.uint8))
        L.orc_export_reads(self.h, *[_p(d[x]) for x in ("tid", "start", "end", "mapq", "cpg_off", "cpg_pos", "cpg_rel", "cpg_meth")])
        return d

    def set_cpg_set(self, tid, pos):
        t, p = np.ascontiguousarray(tid, np.int32), np.ascontiguousarray(pos, np.int32)
        lib().orc_set_cpg_set(self.h, len(t), _p(t), _p(p))

    def set_cpg_set_file(self, path):
        if lib().orc_set_cpg_set_file(self.h, path.encode()) != 0:
            raise IOError(lib().orc_error(self.h).decode())

    def clear_cpg_set(self):
        lib().orc_clear_cpg_set(self.h)

    # ----- measures -------------------------------------------------------------------------
    def pdr(self, min_depth=10, min_cpgs=4, min_qual=10):
        L = lib()
        n = L.orc_pdr(self.h, min_depth, min_cpgs, min_qual)
        r = dict(tid=np.zeros(n, np.int32), pos=np.zeros(n, np.int32), pdr=np.zeros(n, np.float32),
                 n_conc=np.zeros(n, np.uint32), n_disc=np.zeros(n, np.uint32))
        L.orc_pdr_rows(self.h, _p(r["tid"]), _p(r["pos"]), _p(r["pdr"]), _p(r["n_conc"]), _p(r["n_disc"]))
        return r

    def _site(self, n):
        r = dict(tid=np.zeros(n, np.int32), pos=np.zeros(n, np.int32), value=np.zeros(n, np.float32))
        lib().orc_site_rows(self.h, _p(r["tid"]), _p(r["pos"]), _p(r["value"]))
        return r

    def mhl(self, min_depth=10, min_cpgs=4, min_qual=10):
        return self._site(lib().orc_mhl(self.h, min_depth, min_cpgs, min_qual))

    def fdrp(self, min_qual=10, min_depth=10, max_depth=40, min_overlap=35, seed=0, quantitative=False):
        return self._site(lib().orc_fdrp(self.h, int(quantitative), min_qual, min_depth, max_depth, min_overlap, seed))

    def qfdrp(self, **kw):
        return self.fdrp(quantitative=True, **kw)

    def fdrp_oob(self):
        return bool(lib().orc_fdrp_oob(self.h))

    def quartets(self, min_depth=10, min_qual=10):
        L = lib()
        n = L.orc_quartets(self.h, min_depth, min_qual)
        r = dict(tid=np.zeros(n, np.int32), p1=np.zeros(n, np.int32), p2=np.zeros(n, np.int32), p3=np.zeros(n, np.int32),
                 p4=np.zeros(n, np.int32), pm=np.zeros(n, np.float32), me=np.zeros(n, np.float32),
                 counts=np.zeros((n, 16), np.uint32))
        L.orc_quartet_rows(self.h, *[_p(r[x]) for x in ("tid", "p1", "p2", "p3", "p4", "pm", "me", "counts")])
        return r

    def lpmd(self, min_distance=2, max_distance=16, min_qual=10, pairs=False):
        L = lib()
        o4 = np.zeros(4, np.int32)
        v = C.c_float(0)
        n = L.orc_lpmd(self.h, min_distance, max_distance, min_qual, int(pairs), _p(o4), C.byref(v))
        res = dict(n_read=int(o4[0]), n_valid_read=int(o4[1]), n_conc=int(o4[2]), n_disc=int(o4[3]),
                   lpmd=np.float32(v.value))
        if pairs:
            r = dict(tid=np.zeros(n, np.int32), pos1=np.zeros(n, np.int32), pos2=np.zeros(n, np.int32),
                     lpmd=np.zeros(n, np.float32), n_conc=np.zeros(n, np.int32), n_disc=np.zeros(n, np.int32))
            L.orc_lpmd_pair_rows(self.h, *[_p(r[x]) for x in ("tid", "pos1", "pos2", "lpmd", "n_conc", "n_disc")])
            res["pairs"] = r
        return res
