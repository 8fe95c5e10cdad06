"""Python host layer over the C ABI: a thin Context object plus numpy/torch plumbing.  All computation happens in
libmetheor_b200.so on the GPU; this file only marshals pointers."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (FLAG_FORCE_GATHER, FLAG_KEEP_ON_DEVICE, FLAG_PROFILE, FLAG_QUARTET_COUNTS, MEASURE_BITS, Batch,
                   BatchCompact, LpmdResult, Params, Results, Stats)


class EngineError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"metheor_b200 error {code}: {msg}")
        self.code = code


_NP = dict(start=np.int32, end=np.int32, meta=np.uint32, cpg_off=np.uint32, cpg_pos=np.int32, cpg_rel=np.uint16,
           meth=np.uint64, meth_off=np.uint32)


def default_params(measures=(), flags=0, seed=0, **overrides):
    """measures: iterable of names; overrides like pdr=dict(min_depth=1) set per-subcommand thresholds (lib.rs flags)."""
    p = Params()
    _lib.lib().mth_params_default(C.byref(p))
    for m in measures:
        p.measures |= MEASURE_BITS[m]
    p.flags = flags
    p.seed = seed
    for k, d in overrides.items():
        sub = getattr(p, k)
        for f, v in d.items():
            if not hasattr(sub, f):
                raise AttributeError(f"{k}.{f}")
            setattr(sub, f, v)
    return p


def _as_ptr(x):
    """numpy array / torch tensor / None -> (address or None, keepalive, is_device)."""
    if x is None:
        return None, None, False
    if isinstance(x, np.ndarray):
        return x.ctypes.data, x, False
    if hasattr(x, "data_ptr"):  # torch tensor
        return x.data_ptr(), x, bool(x.is_cuda)
    raise TypeError(type(x))


class Context:
    """One engine context on one GPU (mth_ctx).  submit(batch)... then finish() -> dict of numpy row arrays."""

    def __init__(self, params, ref_len, device=0):
        L = _lib.lib()
        self._L = L
        self.params = params
        self._h = C.c_void_p()
        arr = (C.c_int64 * len(ref_len))(*[int(x) for x in ref_len])
        rc = L.mth_ctx_create(C.byref(self._h), device, C.byref(params), len(ref_len), arr)
        if rc != 0:
            raise EngineError(rc, L.mth_last_error(None).decode())
        self._keep = []
        self._copy_rows = True

    def _check(self, rc):
        if rc != 0:
            raise EngineError(rc, self._L.mth_last_error(self._h).decode())

    def close(self):
        if self._h:
            self._L.mth_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        self._check(self._L.mth_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def submit(self, b):
        """b: dict with tid, n_reads, n_cpg and the SoA arrays (numpy host arrays or torch CUDA tensors), or a struct made by
        prepare() (the ctypes marshalling done once: a caller that submits the same arrays many times, like bench.py, should not
        pay ~100 us of Python per call inside its timed loop)."""
        if isinstance(b, Batch):
            self._check(self._L.mth_submit(self._h, C.byref(b)))
            return
        mb, keep = self.prepare(b, _with_keep=True)
        self._keep.extend(keep)
        self._check(self._L.mth_submit(self._h, C.byref(mb)))

    @staticmethod
    def prepare(b, _with_keep=False):
        """dict batch -> mth_batch struct (the caller keeps the arrays alive)."""
        keep_list = []
        mb = Batch()
        mb.tid, mb.n_reads, mb.n_cpg = int(b["tid"]), int(b["n_reads"]), int(b["n_cpg"])
        dev = None
        for f in ("start", "end", "meta", "cpg_off", "cpg_pos", "cpg_rel", "meth", "meth_off"):
            x = b.get(f)
            if isinstance(x, np.ndarray):
                x = np.ascontiguousarray(x, _NP[f])
            p, keep, is_dev = _as_ptr(x)
            if p is not None:
                dev = is_dev if dev is None else dev
                if dev != is_dev:
                    raise ValueError("batch mixes host and device arrays")
                keep_list.append(keep)
            setattr(mb, f, p)
        mb.mem_kind = 1 if dev else 0
        mo = b.get("meth_off")
        if mo is None:
            mb.n_meth_words = mb.n_reads
        else:
            mb.n_meth_words = int(b["n_meth_words"]) if "n_meth_words" in b else int(len(b["meth"]))
        return (mb, keep_list) if _with_keep else mb

    def submit_compact(self, b):
        """b: dict in the compact wire format (batch.to_compact): numpy host arrays or torch CUDA tensors."""
        mb = BatchCompact()
        mb.tid, mb.n_reads, mb.n_cpg, mb.n_rel = int(b["tid"]), int(b["n_reads"]), int(b["n_cpg"]), int(b["n_rel"])
        dt = dict(start=np.int32, span=np.uint16, mapq=np.uint8, n_cpg8=np.uint8, flags=np.uint8, cpg_delta=np.uint16,
                  meth_bits=np.uint8, rel_exc=np.uint16, start_off16=np.uint16, blk_start=np.int32, start_exc=np.int32,
                  cpg_delta8=np.uint8, blk_call_off=np.uint32)
        mb.enc = int(b.get("enc", 0))
        mb.n_start_exc, mb.n_delta8, mb.n_delta16 = int(b.get("n_start_exc", 0)), int(b.get("n_delta8", 0)), int(b.get("n_delta16", 0))
        dev = None
        for f, t in dt.items():
            x = b.get(f)
            if isinstance(x, np.ndarray):
                x = np.ascontiguousarray(x, t)
            p, keep, is_dev = _as_ptr(x)
            if p is not None and (not isinstance(x, np.ndarray) or x.size):
                dev = is_dev if dev is None else dev
                if dev != is_dev:
                    raise ValueError("batch mixes host and device arrays")
                self._keep.append(keep)
            setattr(mb, f, p)
        mb.mem_kind = 1 if dev else 0
        self._check(self._L.mth_submit_compact(self._h, C.byref(mb)))

    def set_cpg_set(self, tid, pos):
        """`--cpg-set` on the device: (tid, pos) pairs of the BED file; batches are then submitted unfiltered."""
        t, p = np.ascontiguousarray(tid, np.int32), np.ascontiguousarray(pos, np.int32)
        self._check(self._L.mth_set_cpg_set(self._h, len(t), t.ctypes.data, p.ctypes.data))

    def clear_cpg_set(self):
        self._check(self._L.mth_clear_cpg_set(self._h))

    def add_skipped_reads(self, n_reads, n_mapq_ok):
        self._check(self._L.mth_add_skipped_reads(self._h, n_reads, n_mapq_ok))

    def _np(self, ptr, n, dtype, cols=None):
        if not ptr or n == 0:
            return np.zeros((0,) if cols is None else (0, cols), dtype)
        count = n * (cols or 1)
        a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(count,))
        if self._copy_rows:
            a = a.copy()
        return a if cols is None else a.reshape(n, cols)

    def finish(self, copy=True):
        """-> dict measure -> rows.  copy=True: numpy copies owned by the caller; copy=False: zero-copy views of the
        context's pinned result buffers (the C ABI's own contract: valid until the next finish / reset / close).
        With FLAG_KEEP_ON_DEVICE only row counts are returned."""
        self._copy_rows = bool(copy)
        r = Results()
        self._check(self._L.mth_finish(self._h, C.byref(r)))
        self._keep.clear()
        out = {}
        M = self.params.measures
        for name in ("pdr", "mhl", "fdrp", "qfdrp"):
            if not (M & MEASURE_BITS[name]):
                continue
            s = getattr(r, name)
            d = dict(n=int(s.n), tid=self._np(s.tid, s.n, np.int32), pos=self._np(s.pos, s.n, np.int32),
                     value=self._np(s.value, s.n, np.float32))
            if name == "pdr":
                d["n_conc"] = self._np(s.n_conc, s.n, np.uint32)
                d["n_disc"] = self._np(s.n_disc, s.n, np.uint32)
            out[name] = d
        for name in ("pm", "me"):
            if not (M & MEASURE_BITS[name]):
                continue
            s = getattr(r, name)
            d = dict(n=int(s.n), tid=self._np(s.tid, s.n, np.int32), value=self._np(s.value, s.n, np.float32))
            for k in ("p1", "p2", "p3", "p4"):
                d[k] = self._np(getattr(s, k), s.n, np.int32)
            if s.counts:
                d["counts"] = self._np(s.counts, s.n, np.uint32, 16)
            out[name] = d
        if M & MEASURE_BITS["lpmd"]:
            out["lpmd"] = dict(n_read=r.lpmd.n_read, n_valid_read=r.lpmd.n_valid_read, n_conc=r.lpmd.n_conc,
                               n_disc=r.lpmd.n_disc, lpmd=np.float32(r.lpmd.lpmd))
            if self.params.lpmd.want_pairs:
                s = r.lpmd_pairs
                out["lpmd"]["pairs"] = dict(n=int(s.n), tid=self._np(s.tid, s.n, np.int32), pos1=self._np(s.pos1, s.n, np.int32),
                                            pos2=self._np(s.pos2, s.n, np.int32), lpmd=self._np(s.lpmd, s.n, np.float32),
                                            n_conc=self._np(s.n_conc, s.n, np.int32), n_disc=self._np(s.n_disc, s.n, np.int32))
        return out

    def results_device(self):
        r = Results()
        self._check(self._L.mth_results_device(self._h, C.byref(r)))
        return r

    def lpmd_counters_device_ptr(self):
        p = C.c_void_p()
        self._check(self._L.mth_lpmd_counters_device(self._h, C.byref(p)))
        return p.value

    def lpmd_refresh(self):
        r = LpmdResult()
        self._check(self._L.mth_lpmd_refresh(self._h, C.byref(r)))
        return dict(n_read=r.n_read, n_valid_read=r.n_valid_read, n_conc=r.n_conc, n_disc=r.n_disc, lpmd=np.float32(r.lpmd))

    # ---- multi-GPU: one context per rank, joined by NCCL inside the library (mth_comm_*, mth_allreduce) ----
    def comm_init_rank(self, n_ranks, rank, unique_id):
        """unique_id: the 128 bytes rank 0 got from `comm_unique_id()` (shipped to the ranks by the caller)."""
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        self._check(self._L.mth_comm_init_rank(self._h, n_ranks, rank, C.cast(buf, C.c_void_p)))

    def allreduce(self):
        """Sum LPMD's counters over the ranks (after finish()); -> the global LPMD result."""
        self._check(self._L.mth_allreduce(self._h))
        return self.lpmd_refresh()

    def comm_n_ranks(self):
        return int(self._L.mth_comm_n_ranks(self._h))

    def reset(self):
        self._check(self._L.mth_reset(self._h))
        self._keep.clear()

    def sync(self):
        self._check(self._L.mth_sync(self._h))

    def stats(self):
        s = Stats()
        self._check(self._L.mth_get_stats(self._h, C.byref(s)))
        d = {k: getattr(s, k) for k in ("n_reads", "n_cpg", "n_sites", "n_regions", "kernel_launches", "h2d_bytes",
                                         "d2h_bytes", "fdrp_pair_ops", "fallback_sites_mhl", "fallback_sites_fdrp", "max_ref_span", "pdr_path")}
        d["kernels"] = {s.kernel[i].name.decode(): dict(launches=s.kernel[i].launches, ms=s.kernel[i].ms)
                        for i in range(s.n_kernel_stats)}
        return d


def comm_unique_id():
    buf = (C.c_char * 128)()
    rc = _lib.lib().mth_comm_unique_id(C.cast(buf, C.c_void_p))
    if rc != 0:
        raise EngineError(rc, _lib.lib().mth_last_error(None).decode())
    return bytes(buf)


def comm_init_all(contexts):
    """Single process, several contexts (one per GPU): ncclCommInitAll."""
    arr = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
    rc = _lib.lib().mth_comm_init_all(arr, len(contexts))
    if rc != 0:
        raise EngineError(rc, _lib.lib().mth_last_error(contexts[0]._h).decode())


def allreduce_group(contexts):
    arr = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
    rc = _lib.lib().mth_allreduce_group(arr, len(contexts))
    if rc != 0:
        raise EngineError(rc, _lib.lib().mth_last_error(contexts[0]._h).decode())
    return [c.lpmd_refresh() for c in contexts]


def run_batches(batches, ref_len, measures, device=0, flags=0, seed=0, compact=False, cpg_set=None, **overrides):
    """Convenience: one context, submit every batch, finish.  -> (results dict, stats dict)
    compact: send the batches in the compact wire format (True), or alternate between the two formats ("mix")."""
    from .batch import to_compact
    ctx = Context(default_params(measures, flags=flags, seed=seed, **overrides), ref_len, device)
    try:
        if cpg_set is not None:
            ctx.set_cpg_set(*cpg_set)
        for k, b in enumerate(batches):
            if compact and (compact != "mix" or k % 2 == 0) and b["n_reads"] and \
                    np.diff(np.asarray(b["cpg_off"], np.int64)).max(initial=0) <= 64:
                ctx.submit_compact(to_compact(b, dense=(compact == "dense" or (compact == "mix" and k % 4 == 0))))
            else:
                ctx.submit(b)
        res = ctx.finish()
        return res, ctx.stats()
    finally:
        ctx.close()
