"""Python plumbing over the device-side BGZF / BAM decoder of libmetheor_b200.so (mth_bamdec_*, mth_bgzf_inflate): BGZF member
walk on the host (what the C++ host does in host/run_gpu.cpp), windows of compressed bytes in, device-resident SoA batches out.
Tests and profiles only; nothing here computes."""
import ctypes as C
import struct

import numpy as np

from . import _lib
from ._lib import BamdecResult, BgzfMember


def bgzf_members(data, offset=0):
    """Walk the BGZF member headers of `data` (bytes) from `offset` -> list of (payload offset, payload size, isize, member end, crc)."""
    out = []
    n = len(data)
    o = offset
    while o + 18 <= n:
        if data[o] != 31 or data[o + 1] != 139:
            raise ValueError(f"not a gzip member at {o}")
        xlen = struct.unpack_from("<H", data, o + 10)[0]
        x, bsize = o + 12, None
        while x + 4 <= o + 12 + xlen:
            si1, si2, slen = data[x], data[x + 1], struct.unpack_from("<H", data, x + 2)[0]
            if si1 == 66 and si2 == 67 and slen == 2:
                bsize = struct.unpack_from("<H", data, x + 4)[0]
            x += 4 + slen
        if bsize is None:
            raise ValueError(f"gzip member without the BGZF extra field at {o}")
        total = bsize + 1
        pay = o + 12 + xlen
        crc, isize = struct.unpack_from("<II", data, o + total - 8)
        out.append((pay, total - (12 + xlen) - 8, isize, o + total, crc))
        o += total
    return out


def inflate_members(data, members, device=0):
    """GPU inflate of the given members of `data` -> (bytes, per-member status, kernel ms)."""
    L = _lib.lib()
    n = len(members)
    arr = (BgzfMember * max(n, 1))()
    tot = 0
    for i, mb in enumerate(members):
        off, size, isize = mb[0], mb[1], mb[2]
        arr[i].offset, arr[i].size, arr[i].isize = off, size, isize
        if len(mb) > 4:
            arr[i].crc, arr[i].flags = mb[4], 1
        tot += isize
    out = np.zeros(max(tot, 1), np.uint8)
    status = np.zeros(max(n, 1), np.int32)
    ms = C.c_double(0)
    buf = np.frombuffer(data, np.uint8)
    rc = L.mth_bgzf_inflate(device, buf.ctypes.data, len(data), arr, n, out.ctypes.data, out.nbytes, status.ctypes.data, C.byref(ms))
    if rc != 0:
        raise RuntimeError(f"mth_bgzf_inflate: {rc} {L.mth_bamdec_last_error(None).decode()}")
    return out[:tot].tobytes(), status[:n], ms.value


class Decoder:
    """mth_bamdec: windows of compressed BGZF members -> device SoA batches (fetched to numpy here for the tests)."""

    def __init__(self, ref_len, device=0, lpmd_order=False, min_qual=0):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        arr = (C.c_int64 * max(len(ref_len), 1))(*[int(x) for x in ref_len])
        rc = self._L.mth_bamdec_create(C.byref(self._h), device, len(ref_len), arr, int(lpmd_order), int(min_qual))
        if rc != 0:
            raise RuntimeError(f"mth_bamdec_create: {rc} {self._L.mth_bamdec_last_error(None).decode()}")

    def close(self):
        if self._h:
            self._L.mth_bamdec_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def window(self, data, members, skip=0, last=False, fetch=True):
        """-> (result struct, list of numpy batches).  `members` as from bgzf_members (absolute offsets into data)."""
        n = len(members)
        arr = (BgzfMember * max(n, 1))()
        for i, mb in enumerate(members):
            arr[i].offset, arr[i].size, arr[i].isize = mb[0], mb[1], mb[2]
            if len(mb) > 4:
                arr[i].crc, arr[i].flags = mb[4], 1
        buf = np.frombuffer(data, np.uint8)
        res = BamdecResult()
        rc = self._L.mth_bamdec_window(self._h, buf.ctypes.data, len(data), arr, n, int(skip), int(last), C.byref(res))
        if rc != 0:
            raise RuntimeError(f"mth_bamdec_window: {rc} {self._L.mth_bamdec_last_error(self._h).decode()}")
        batches = []
        if fetch:
            import torch
            for k in range(res.n_runs):
                b = res.runs[k]
                R, I = int(b.n_reads), int(b.n_cpg)

                def dev(ptr, n_el, dt, ts):
                    if n_el == 0:
                        return np.zeros(0, dt)
                    t = torch.as_tensor(_Arr(ptr, n_el, ts), device="cuda")
                    return t.cpu().numpy().view(dt).copy()
                batches.append(dict(tid=int(b.tid), n_reads=R, n_cpg=I, start=dev(b.start, R, np.int32, "<i4"), end=dev(b.end, R, np.int32, "<i4"),
                                    meta=dev(b.meta, R, np.uint32, "<i4"), cpg_off=dev(b.cpg_off, R + 1, np.uint32, "<i4"),
                                    cpg_pos=dev(b.cpg_pos, I, np.int32, "<i4"), cpg_rel=dev(b.cpg_rel, I, np.uint16, "<i2"),
                                    meth=dev(b.meth, R, np.uint64, "<i8"), meth_off=None))
        return res, batches


class _Arr:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}
