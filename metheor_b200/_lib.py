"""ctypes binding of libmetheor_b200.so (include/metheor_b200.h).  The library is the product: if it is missing or
cannot be loaded this module raises — there is no Python/CPU fallback."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# METHEOR_B200_LIB selects another build of the same library (kernel-variant experiments); never a fallback
LIB_PATH = os.environ.get("METHEOR_B200_LIB") or os.path.join(CSRC, "libmetheor_b200.so")

MTH_PDR, MTH_LPMD, MTH_MHL, MTH_PM, MTH_ME, MTH_FDRP, MTH_QFDRP = (1 << i for i in range(7))
MTH_ALL = 0x7F
MEASURE_BITS = dict(pdr=MTH_PDR, lpmd=MTH_LPMD, mhl=MTH_MHL, pm=MTH_PM, me=MTH_ME, fdrp=MTH_FDRP, qfdrp=MTH_QFDRP)
FLAG_KEEP_ON_DEVICE, FLAG_PROFILE, FLAG_QUARTET_COUNTS, FLAG_FORCE_GATHER = 1, 2, 4, 8
MTH_OK, ERR_INVALID, ERR_CUDA, ERR_UNSORTED, ERR_UNSUPPORTED, ERR_STATE = 0, -1, -2, -3, -4, -5
ABI_VERSION = 2

u32, i32, i64, u64, vp = C.c_uint32, C.c_int32, C.c_int64, C.c_uint64, C.c_void_p


class PdrParams(C.Structure):
    _fields_ = [("min_depth", u32), ("min_cpgs", u32), ("min_qual", u32)]


class LpmdParams(C.Structure):
    _fields_ = [("min_distance", i32), ("max_distance", i32), ("min_qual", u32), ("want_pairs", u32)]


class MhlParams(C.Structure):
    _fields_ = [("min_depth", u32), ("min_cpgs", u32), ("min_qual", u32)]


class QuartetParams(C.Structure):
    _fields_ = [("min_depth", u32), ("min_qual", u32)]


class FdrpParams(C.Structure):
    _fields_ = [("min_qual", u32), ("min_depth", u32), ("max_depth", u32), ("min_overlap", i32)]


class Params(C.Structure):
    _fields_ = [("abi_version", u32), ("measures", u32), ("flags", u32), ("reserved", u32), ("pdr", PdrParams),
                ("lpmd", LpmdParams), ("mhl", MhlParams), ("pm", QuartetParams), ("me", QuartetParams),
                ("fdrp", FdrpParams), ("qfdrp", FdrpParams), ("seed", u64)]


class Batch(C.Structure):
    _fields_ = [("tid", i32), ("mem_kind", i32), ("n_reads", i64), ("n_cpg", i64), ("n_meth_words", i64),
                ("start", vp), ("end", vp), ("meta", vp), ("cpg_off", vp), ("cpg_pos", vp), ("cpg_rel", vp),
                ("meth", vp), ("meth_off", vp)]


class BatchCompact(C.Structure):
    _fields_ = [("tid", i32), ("mem_kind", i32), ("n_reads", i64), ("n_cpg", i64), ("n_rel", i64), ("start", vp), ("span", vp),
                ("mapq", vp), ("n_cpg8", vp), ("flags", vp), ("cpg_delta", vp), ("meth_bits", vp), ("rel_exc", vp),
                ("enc", u32), ("reserved", u32), ("start_off16", vp), ("blk_start", vp), ("start_exc", vp), ("n_start_exc", i64),
                ("cpg_delta8", vp), ("blk_call_off", vp), ("n_delta8", i64), ("n_delta16", i64)]


class SiteRows(C.Structure):
    _fields_ = [("n", i64), ("tid", vp), ("pos", vp), ("value", vp), ("n_conc", vp), ("n_disc", vp)]


class QuartetRows(C.Structure):
    _fields_ = [("n", i64), ("tid", vp), ("p1", vp), ("p2", vp), ("p3", vp), ("p4", vp), ("value", vp), ("counts", vp)]


class LpmdResult(C.Structure):
    _fields_ = [("n_read", i64), ("n_valid_read", i64), ("n_conc", i64), ("n_disc", i64), ("lpmd", C.c_float)]


class PairRows(C.Structure):
    _fields_ = [("n", i64), ("tid", vp), ("pos1", vp), ("pos2", vp), ("lpmd", vp), ("n_conc", vp), ("n_disc", vp)]


class Results(C.Structure):
    _fields_ = [("pdr", SiteRows), ("mhl", SiteRows), ("fdrp", SiteRows), ("qfdrp", SiteRows), ("pm", QuartetRows),
                ("me", QuartetRows), ("lpmd", LpmdResult), ("lpmd_pairs", PairRows)]


class KernelStat(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("launches", i64), ("ms", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("n_reads", i64), ("n_cpg", i64), ("n_sites", i64), ("n_regions", i64), ("kernel_launches", i64),
                ("h2d_bytes", i64), ("d2h_bytes", i64), ("fdrp_pair_ops", i64), ("fallback_sites_mhl", i64), ("fallback_sites_fdrp", i64), ("max_ref_span", i32), ("pdr_path", i32),
                ("n_kernel_stats", i32), ("kernel", KernelStat * 48)]


class TagBatch(C.Structure):
    _fields_ = [("n_reads", i64), ("tid", vp), ("pos", vp), ("rc", vp), ("l_seq", vp), ("cigar_off", vp), ("cigar", vp),
                ("seq_off", vp), ("seq4", vp)]


class TagResult(C.Structure):
    _fields_ = [("n_reads", i64), ("n_failed", i64), ("xm_off", vp), ("xm_len", vp), ("xm", vp), ("status", vp)]


class BgzfMember(C.Structure):
    _fields_ = [("offset", u64), ("size", u32), ("isize", u32), ("crc", u32), ("flags", u32)]


class BamdecResult(C.Structure):
    _fields_ = [("n_records", i64), ("n_dropped", i64), ("n_dropped_mapq_ok", i64), ("n_runs", i32), ("max_cpgs", i32), ("runs", C.POINTER(Batch)),
                ("max_span", i64), ("bad_record", i64), ("bad_is_corrupt", i32), ("reserved", i32), ("uncompressed_bytes", u64),
                ("ms_inflate", C.c_double), ("ms_boundaries", C.c_double), ("ms_decode", C.c_double), ("chain_repairs", i64)]


EXPORTS = ["mth_params_default", "mth_ctx_create", "mth_ctx_destroy", "mth_set_stream", "mth_submit", "mth_submit_compact", "mth_reserve",
           "mth_add_skipped_reads", "mth_finish", "mth_results_device", "mth_lpmd_counters_device", "mth_lpmd_refresh",
           "mth_reset", "mth_sync", "mth_sync_copies", "mth_get_stats", "mth_last_error", "mth_host_alloc", "mth_host_free",
           "mth_device_count", "mth_version", "mth_reservoir_draw", "mth_genome_create", "mth_genome_set_contig", "mth_tag",
           "mth_genome_destroy", "mth_genome_last_error", "mth_genome_last_kernel_ms", "mth_comm_unique_id", "mth_comm_init_rank",
           "mth_comm_init_all", "mth_allreduce", "mth_allreduce_group", "mth_comm_destroy", "mth_comm_n_ranks", "mth_set_cpg_set", "mth_clear_cpg_set", "mth_bamdec_create",
           "mth_bamdec_window", "mth_bamdec_stage", "mth_bamdec_destroy", "mth_bamdec_last_error", "mth_bgzf_inflate"]


def build(force=False):
    """Compile the CUDA library in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    if force:
        subprocess.check_call(["make", "-C", CSRC, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", CSRC, "-j8"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C {CSRC}` (python -c 'import __graft_entry__ as g; "
                           "g.build()'); metheor_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    P = C.POINTER
    L.mth_params_default.argtypes = [P(Params)]; L.mth_params_default.restype = None
    L.mth_ctx_create.argtypes = [P(vp), C.c_int, P(Params), i32, P(i64)]; L.mth_ctx_create.restype = C.c_int
    L.mth_ctx_destroy.argtypes = [vp]; L.mth_ctx_destroy.restype = C.c_int
    L.mth_set_stream.argtypes = [vp, vp]; L.mth_set_stream.restype = C.c_int
    L.mth_submit.argtypes = [vp, P(Batch)]; L.mth_submit.restype = C.c_int
    L.mth_reserve.argtypes = [vp, i64, i64]; L.mth_reserve.restype = C.c_int
    L.mth_submit_compact.argtypes = [vp, P(BatchCompact)]; L.mth_submit_compact.restype = C.c_int
    L.mth_add_skipped_reads.argtypes = [vp, i64, i64]; L.mth_add_skipped_reads.restype = C.c_int
    L.mth_finish.argtypes = [vp, P(Results)]; L.mth_finish.restype = C.c_int
    L.mth_results_device.argtypes = [vp, P(Results)]; L.mth_results_device.restype = C.c_int
    L.mth_lpmd_counters_device.argtypes = [vp, P(vp)]; L.mth_lpmd_counters_device.restype = C.c_int
    L.mth_lpmd_refresh.argtypes = [vp, P(LpmdResult)]; L.mth_lpmd_refresh.restype = C.c_int
    L.mth_reset.argtypes = [vp]; L.mth_reset.restype = C.c_int
    L.mth_sync.argtypes = [vp]; L.mth_sync.restype = C.c_int
    L.mth_sync_copies.argtypes = [vp]; L.mth_sync_copies.restype = C.c_int
    L.mth_get_stats.argtypes = [vp, P(Stats)]; L.mth_get_stats.restype = C.c_int
    L.mth_last_error.argtypes = [vp]; L.mth_last_error.restype = C.c_char_p
    L.mth_host_alloc.argtypes = [C.c_size_t]; L.mth_host_alloc.restype = vp
    L.mth_host_free.argtypes = [vp]; L.mth_host_free.restype = None
    L.mth_device_count.argtypes = []; L.mth_device_count.restype = C.c_int
    L.mth_version.argtypes = []; L.mth_version.restype = C.c_char_p
    L.mth_reservoir_draw.argtypes = [u64, i32, i32, u32]; L.mth_reservoir_draw.restype = u32
    L.mth_set_cpg_set.argtypes = [vp, i64, vp, vp]; L.mth_set_cpg_set.restype = C.c_int
    L.mth_clear_cpg_set.argtypes = [vp]; L.mth_clear_cpg_set.restype = C.c_int
    L.mth_bamdec_create.argtypes = [P(vp), C.c_int, i32, P(i64), u32, u32]; L.mth_bamdec_create.restype = C.c_int
    L.mth_bamdec_window.argtypes = [vp, vp, C.c_size_t, P(BgzfMember), i64, u64, C.c_int, P(BamdecResult)]; L.mth_bamdec_window.restype = C.c_int
    L.mth_bamdec_stage.argtypes = [vp, C.c_int, vp, C.c_size_t]; L.mth_bamdec_stage.restype = C.c_int
    L.mth_bamdec_destroy.argtypes = [vp]; L.mth_bamdec_destroy.restype = C.c_int
    L.mth_bamdec_last_error.argtypes = [vp]; L.mth_bamdec_last_error.restype = C.c_char_p
    L.mth_bgzf_inflate.argtypes = [C.c_int, vp, C.c_size_t, P(BgzfMember), i64, vp, C.c_size_t, vp, P(C.c_double)]; L.mth_bgzf_inflate.restype = C.c_int
    L.mth_comm_unique_id.argtypes = [vp]; L.mth_comm_unique_id.restype = C.c_int
    L.mth_comm_init_rank.argtypes = [vp, C.c_int, C.c_int, vp]; L.mth_comm_init_rank.restype = C.c_int
    L.mth_comm_init_all.argtypes = [P(vp), C.c_int]; L.mth_comm_init_all.restype = C.c_int
    L.mth_allreduce.argtypes = [vp]; L.mth_allreduce.restype = C.c_int
    L.mth_allreduce_group.argtypes = [P(vp), C.c_int]; L.mth_allreduce_group.restype = C.c_int
    L.mth_comm_destroy.argtypes = [vp]; L.mth_comm_destroy.restype = C.c_int
    L.mth_comm_n_ranks.argtypes = [vp]; L.mth_comm_n_ranks.restype = C.c_int
    L.mth_genome_create.argtypes = [P(vp), C.c_int, i32, P(i64)]; L.mth_genome_create.restype = C.c_int
    L.mth_genome_set_contig.argtypes = [vp, i32, vp, i64]; L.mth_genome_set_contig.restype = C.c_int
    L.mth_tag.argtypes = [vp, P(TagBatch), P(TagResult)]; L.mth_tag.restype = C.c_int
    L.mth_genome_destroy.argtypes = [vp]; L.mth_genome_destroy.restype = C.c_int
    L.mth_genome_last_error.argtypes = [vp]; L.mth_genome_last_error.restype = C.c_char_p
    L.mth_genome_last_kernel_ms.argtypes = [vp]; L.mth_genome_last_kernel_ms.restype = C.c_double
    _lib = L
    return L
