"""ctypes binding of the C++ host (include/metheor_host.h): libmetheor_host.so and the `metheor` binary.
The host decodes BAM/SAM on the CPU cores and drives the GPU engine; nothing here computes a measure."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_DIR = os.path.join(HERE, "host")
LIB_PATH = os.path.join(HOST_DIR, "libmetheor_host.so")
BIN_PATH = os.path.join(HERE, "bin", "metheor")
MEASURES = dict(pdr=0, lpmd=1, mhl=2, pm=3, me=4, fdrp=5, qfdrp=6)


class Options(C.Structure):
    _fields_ = [("measure", C.c_int32), ("input", C.c_char_p), ("output", C.c_char_p), ("cpg_set", C.c_char_p),
                ("pairs", C.c_char_p), ("min_depth", C.c_uint32), ("min_cpgs", C.c_uint32), ("min_qual", C.c_uint32),
                ("max_depth", C.c_uint32), ("min_overlap", C.c_int32), ("min_distance", C.c_int32),
                ("max_distance", C.c_int32), ("device", C.c_int32), ("n_gpus", C.c_int32), ("shard_contigs", C.c_int32), ("decode_host", C.c_int32), ("out_format", C.c_int32), ("region", C.c_char_p), ("threads", C.c_int32),
                ("seed", C.c_uint64), ("stats_json", C.c_char_p)]


class Decoded(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("n_cpg", C.c_int64), ("n_ref", C.c_int32), ("ref_name", C.POINTER(C.c_char_p)),
                ("ref_len", C.POINTER(C.c_int64)), ("tid", C.c_void_p), ("start", C.c_void_p), ("end", C.c_void_p),
                ("meta", C.c_void_p), ("cpg_off", C.c_void_p), ("cpg_pos", C.c_void_p), ("cpg_rel", C.c_void_p),
                ("cpg_meth", C.c_void_p)]


EXPORTS = ["mthh_options_default", "mthh_run", "mthh_main", "mthh_decode_file", "mthh_decoded_free", "mthh_format_f32", "mthh_inflate_raw",
           "mthh_zlib_fallbacks", "mthh_tag", "mthh_crc32", "mthh_plan_shards"]


class Interval(C.Structure):
    _fields_ = [("rank", C.c_int32), ("tid", C.c_int32), ("lo", C.c_int64), ("hi", C.c_int64)]


class HostError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"metheor host error (status {status}): {msg}")
        self.status, self.msg = status, msg


def build():
    subprocess.check_call(["make", "-C", HOST_DIR, "-j8"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C {HOST_DIR}`")
        L = C.CDLL(LIB_PATH)
        L.mthh_options_default.argtypes = [C.POINTER(Options), C.c_int32]; L.mthh_options_default.restype = None
        L.mthh_run.argtypes = [C.POINTER(Options), C.c_char_p, C.c_size_t]; L.mthh_run.restype = C.c_int
        L.mthh_decode_file.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.POINTER(C.POINTER(Decoded)), C.c_char_p, C.c_size_t]
        L.mthh_decode_file.restype = C.c_int
        L.mthh_decoded_free.argtypes = [C.POINTER(Decoded)]; L.mthh_decoded_free.restype = None
        L.mthh_format_f32.argtypes = [C.c_float, C.c_char_p, C.c_int]; L.mthh_format_f32.restype = C.c_int
        L.mthh_inflate_raw.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]; L.mthh_inflate_raw.restype = C.c_int
        L.mthh_zlib_fallbacks.argtypes = []; L.mthh_zlib_fallbacks.restype = C.c_int64
        L.mthh_crc32.argtypes = [C.c_char_p, C.c_size_t]; L.mthh_crc32.restype = C.c_uint32
        L.mthh_tag.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, C.c_char_p, C.c_char_p, C.c_size_t]
        L.mthh_tag.restype = C.c_int
        _lib = L
    return _lib


def _arr(ptr, n, dtype):
    if not ptr or n == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=(n,)).copy()


def decode_file(path, cpg_set=None, threads=0):
    """BAM/SAM -> dict of numpy arrays, one entry per record (what BismarkRead::new yields); no GPU involved."""
    L = lib()
    out = C.POINTER(Decoded)()
    err = C.create_string_buffer(4096)
    rc = L.mthh_decode_file(path.encode(), cpg_set.encode() if cpg_set else None, threads, C.byref(out), err, 4096)
    if rc != 0:
        raise HostError(rc, err.value.decode())
    d = out.contents
    try:
        R, I = d.n_reads, d.n_cpg
        meta = _arr(d.meta, R, np.uint32)
        return dict(refs=[(d.ref_name[i].decode(), int(d.ref_len[i])) for i in range(d.n_ref)], n_reads=R, n_cpg=I,
                    tid=_arr(d.tid, R, np.int32), start=_arr(d.start, R, np.int32), end=_arr(d.end, R, np.int32),
                    meta=meta, mapq=(meta & 0xFF).astype(np.uint8), cpg_off=_arr(d.cpg_off, R + 1, np.int64),
                    cpg_pos=_arr(d.cpg_pos, I, np.int32), cpg_rel=_arr(d.cpg_rel, I, np.uint16),
                    cpg_meth=_arr(d.cpg_meth, I, np.uint8))
    finally:
        L.mthh_decoded_free(out)


def run(measure, input, output, **kw):
    """One subcommand through the library entry point (the `metheor` binary calls the same function)."""
    L = lib()
    o = Options()
    L.mthh_options_default(C.byref(o), MEASURES[measure])
    o.input, o.output = input.encode(), output.encode()
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v.encode() if isinstance(v, str) else v)
    err = C.create_string_buffer(4096)
    rc = L.mthh_run(C.byref(o), err, 4096)
    if rc != 0:
        raise HostError(rc, err.value.decode())


def plan_shards(ref_len, world, by_contig=False):
    """The multi-GPU plan of `metheor --gpus N` -> list (per rank) of (tid, lo, hi); same result as shard.plan_bins / plan_contigs."""
    L = lib()
    L.mthh_plan_shards.argtypes = [C.c_int32, C.POINTER(C.c_int64), C.c_int32, C.c_int32, C.POINTER(Interval), C.c_int32]
    L.mthh_plan_shards.restype = C.c_int32
    arr = (C.c_int64 * len(ref_len))(*[int(x) for x in ref_len])
    cap = len(ref_len) + 2 * world + 4
    out = (Interval * cap)()
    n = L.mthh_plan_shards(len(ref_len), arr, world, int(by_contig), out, cap)
    res = [[] for _ in range(world)]
    for k in range(n):
        res[out[k].rank].append((out[k].tid, out[k].lo, out[k].hi))
    return res


def format_f32(v):
    buf = C.create_string_buffer(128)
    n = lib().mthh_format_f32(C.c_float(v), buf, 128)
    return buf.raw[:n].decode()


def tag(input, output, genome, device=0, threads=0, stats_json=None):
    """`metheor tag -i input -o output -g genome` in-process (reference src/tag.rs:386-443); raises HostError."""
    err = C.create_string_buffer(4096)
    rc = lib().mthh_tag(str(input).encode(), str(output).encode(), str(genome).encode(), device, threads,
                        stats_json.encode() if stats_json else None, err, 4096)
    if rc != 0:
        raise HostError(rc, err.value.decode())


def inflate_raw(data, out_len):
    """The host's own DEFLATE decoder on one raw stream -> bytes, or None if it rejects the stream."""
    out = C.create_string_buffer(max(out_len, 1))
    ok = lib().mthh_inflate_raw(data, len(data), out, out_len)
    return out.raw[:out_len] if ok else None


def crc32(data):
    return int(lib().mthh_crc32(data, len(data)))


def zlib_fallbacks():
    return int(lib().mthh_zlib_fallbacks())


def cli(*args, **kw):
    """Run the `metheor` binary; -> CompletedProcess (text mode)."""
    return subprocess.run([BIN_PATH, *map(str, args)], capture_output=True, text=True, **kw)
