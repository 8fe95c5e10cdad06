/* metheor_b200.h — C ABI of the B200-native methylation-heterogeneity engine (libmetheor_b200.so).
 *
 * The reference (dohlee/metheor v0.1.9) has no FFI of its own: its seam for this path is the per-measure
 *   compute_helper(input, <thresholds>, cpg_set) -> map      src/pdr.rs:119  src/lpmd.rs:154  src/mhl.rs:135
 *                                                            src/pm.rs:85    src/me.rs:90      src/fdrp.rs:176
 *                                                            src/qfdrp.rs:188
 * fed by BismarkRead accessors (src/readutil.rs:55-95).  This header is the boundary a host (the C++ host in
 * metheor_b200/host, or a Rust host binding these symbols) calls instead of the body of those
 * loops: the host decodes BAM records + XM tags into structure-of-arrays batches (what BismarkRead::new,
 * src/readutil.rs:24-53,323-345, produces, AFTER the optional --cpg-set filter of src/readutil.rs:87-95) and the
 * engine does every per-read / per-CpG / per-quartet / per-read-pair computation on the GPU.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function returns MTH_OK (0) or a negative
 * mth_status and never aborts; mth_last_error() gives the message.  One caller thread per context; one context
 * per GPU.  The library refuses to run without a CUDA device (no CPU fallback).
 */
#ifndef METHEOR_B200_H
#define METHEOR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTH_ABI_VERSION 2

typedef enum {
    MTH_OK = 0,
    MTH_ERR_INVALID = -1,      /* bad argument / malformed batch */
    MTH_ERR_CUDA = -2,         /* CUDA runtime error (no device, out of memory, launch failure) */
    MTH_ERR_UNSORTED = -3,     /* reads not sorted by (tid, start): the engine requires coordinate-sorted input */
    MTH_ERR_UNSUPPORTED = -4,  /* input outside the engine's documented limits (see DESIGN.md "Limits") */
    MTH_ERR_STATE = -5         /* call sequence error (e.g. submit after finish without reset) */
} mth_status;

/* measure bitmask */
#define MTH_PDR   (1u << 0)
#define MTH_LPMD  (1u << 1)
#define MTH_MHL   (1u << 2)
#define MTH_PM    (1u << 3)
#define MTH_ME    (1u << 4)
#define MTH_FDRP  (1u << 5)
#define MTH_QFDRP (1u << 6)
#define MTH_ALL   0x7Fu

/* context flags */
#define MTH_FLAG_KEEP_ON_DEVICE (1u << 0) /* finish() leaves rows in HBM only (mth_results_device); host pointers NULL */
#define MTH_FLAG_PROFILE        (1u << 1) /* bracket every kernel with CUDA events; mth_get_stats reports per-kernel ms */
#define MTH_FLAG_QUARTET_COUNTS (1u << 2) /* also return the 16 pattern counts per PM/ME row */
#define MTH_FLAG_FORCE_GATHER   (1u << 3) /* PDR: always use the segment-exact gather kernel (testing) */

/* thresholds: one block per reference subcommand, same names/defaults as src/lib.rs:24-231 */
typedef struct { uint32_t min_depth, min_cpgs, min_qual; } mth_pdr_params;                     /* lib.rs:28-52  (10,4,10) */
typedef struct { int32_t min_distance, max_distance; uint32_t min_qual, want_pairs; } mth_lpmd_params; /* lib.rs:198-231 (2,16,10) */
typedef struct { uint32_t min_depth, min_cpgs, min_qual; } mth_mhl_params;                     /* lib.rs:166-192 (10,4,10) */
typedef struct { uint32_t min_depth, min_qual; } mth_quartet_params;                           /* lib.rs:55-98  (10,10) */
typedef struct { uint32_t min_qual, min_depth, max_depth; int32_t min_overlap; } mth_fdrp_params; /* lib.rs:101-163 (10,10,40,35) */

typedef struct {
    uint32_t abi_version; /* MTH_ABI_VERSION */
    uint32_t measures;    /* MTH_* bitmask */
    uint32_t flags;       /* MTH_FLAG_* */
    uint32_t reserved;
    mth_pdr_params pdr;
    mth_lpmd_params lpmd;
    mth_mhl_params mhl;
    mth_quartet_params pm, me;
    mth_fdrp_params fdrp, qfdrp;
    uint64_t seed;        /* reservoir sampling once a pile exceeds max_depth (reference: unseeded, fdrp.rs:90) */
} mth_params;

void mth_params_default(mth_params* p); /* reference defaults, measures = 0 */

/* One batch of decoded reads of ONE contig, in file order (start ascending).  Replaces the per-record
 * BismarkRead of src/readutil.rs:15-21.  Arrays are caller-owned and must stay valid until the next
 * mth_finish()/mth_sync()/mth_sync_copies() returns (copies are asynchronous).  mem_kind: 0 = host memory (pinned preferred,
 * see mth_host_alloc), 1 = device memory on the context's GPU that stays valid until mth_finish (the engine may read it in
 * place), 2 = device memory that is only valid until mth_sync_copies (always copied).
 *   start/end : first / last aligned reference position of the read (readutil.rs:25-33)
 *   meta      : bits 0-7 mapq (pdr.rs:150), bit 8 = forward strand (informational), bit 9 = MTH_META_HALO
 *   cpg_off   : n_reads+1 prefix offsets into cpg_pos / cpg_rel, cpg_off[0] == 0
 *   cpg_pos   : strand-adjusted CpG positions (readutil.rs:332-339), strictly increasing within a read
 *   cpg_rel   : query index of each CpG (readutil.rs:335); only read when MTH_LPMD is set, may be NULL otherwise
 *   meth      : packed methylation calls, bit k of a read's word(s) = its k-th CpG is 'Z' (readutil.rs:258)
 *   meth_off  : n_reads+1 word offsets into meth, or NULL meaning exactly one word per read (needs <= 64 CpGs/read)
 */
/* Position-bin sharding: a rank also receives the reads that start shortly before its bin (halo).
 * They contribute to the rank's own sites like any read, but LPMD's global counters (lpmd.rs:176-191) must count each
 * read once, on the rank that owns it: halo copies carry this bit and are skipped by the LPMD counters only. */
#define MTH_META_HALO (1u << 9)

typedef struct {
    int32_t tid;
    int32_t mem_kind;
    int64_t n_reads;
    int64_t n_cpg;
    int64_t n_meth_words; /* length of meth[] in 64-bit words (= n_reads when meth_off is NULL) */
    const int32_t* start;
    const int32_t* end;
    const uint32_t* meta;
    const uint32_t* cpg_off;
    const int32_t* cpg_pos;
    const uint16_t* cpg_rel;
    const uint64_t* meth;
    const uint32_t* meth_off;
} mth_batch;

/* The same reads in a compact wire format (~1/3 of the bytes of mth_batch: 9 B per read + 2.125 B per CpG call), for
 * hosts whose link to the GPU is the bottleneck (PCIe: the SoA of a 30x chr19 is 468 MB, 172 MB in this form).  The
 * engine expands it on the device into the layout above (k_expand) and proceeds identically.  Restrictions: at most 64
 * CpG calls per read (use mth_batch for a batch that contains a denser read); end - start <= 65023.
 *   start      : as in mth_batch
 *   span       : end - start
 *   mapq       : pdr.rs:150
 *   n_cpg      : CpG calls of the read (replaces cpg_off; prefix-summed on the device)
 *   flags      : bit 0 forward strand, bit 1 halo copy (MTH_META_HALO), bit 2 query indices given explicitly in rel_exc
 *   cpg_delta  : per call, position - (start - 1)                       (a call lies in [start-1, end])
 *   meth_bits  : per call, 1 bit: call x of the batch is bit (x & 7) of byte (x >> 3)
 *   rel_exc    : only read when MTH_LPMD is set: the query indices (readutil.rs:335) of the calls of the reads that
 *                have flag bit 2, in order.  All other reads are plain `<len>M` alignments, for which the query index
 *                is implied: position - start (+1 on the reverse strand, readutil.rs:332-339).
 */
#define MTH_CFLAG_FORWARD 1u
#define MTH_CFLAG_HALO 2u
#define MTH_CFLAG_REL_EXPLICIT 4u
typedef struct {
    int32_t tid;
    int32_t mem_kind;
    int64_t n_reads;
    int64_t n_cpg;
    int64_t n_rel;            /* entries of rel_exc */
    const int32_t* start;
    const uint16_t* span;
    const uint8_t* mapq;
    const uint8_t* n_cpg8;
    const uint8_t* flags;
    const uint16_t* cpg_delta;
    const uint8_t* meth_bits; /* (n_cpg + 7) / 8 bytes */
    const uint16_t* rel_exc;
    /* ---- optional denser encodings (enc != 0): 7 B per read + 1.125 B per call on sorted short-read data ----------
     * Reads are grouped in blocks of MTH_CBLOCK = 256 (block b = reads [256 b, 256 b + 256) of this batch); a block that
     * does not fit the narrow type falls back to the wide one, block by block, so nothing needs an escape code.
     * MTH_CENC_START16: `start` above is not read.  start = blk_start[b] + start_off16[r]; a block whose starts span
     *   more than 65535 has blk_start[b] = -(1 + e) and its 256 absolute starts are start_exc[256 e .. 256 e + 255].
     * MTH_CENC_DELTA8: call positions are deltas from the PREVIOUS call of the read (first call: from start - 1).
     *   blk_call_off[b] = offset of the block's first call in cpg_delta8; with bit 31 set the block holds a delta > 255
     *   and its calls are 16-bit deltas in `cpg_delta` at offset (blk_call_off[b] & 0x7FFFFFFF) instead.  Without this
     *   bit of `enc`, `cpg_delta` keeps its meaning above (one 16-bit offset from start - 1 per call). */
    uint32_t enc;
    uint32_t reserved;
    const uint16_t* start_off16;
    const int32_t* blk_start;
    const int32_t* start_exc;
    int64_t n_start_exc;      /* entries of start_exc (256 per exception block) */
    const uint8_t* cpg_delta8;
    const uint32_t* blk_call_off;
    int64_t n_delta8;         /* entries of cpg_delta8 */
    int64_t n_delta16;        /* entries of cpg_delta when MTH_CENC_DELTA8 is set */
} mth_batch_compact;
#define MTH_CENC_START16 1u
#define MTH_CENC_DELTA8 2u
#define MTH_CBLOCK 256

/* Result rows.  Arrays are owned by the context and valid until the next mth_finish / mth_reset / mth_ctx_destroy. */
typedef struct {          /* pdr.rs:102-116, mhl.rs:122-131, fdrp.rs:169-172, qfdrp.rs:181-184: sorted by (tid,pos) */
    int64_t n;
    const int32_t* tid;
    const int32_t* pos;
    const float* value;
    const uint32_t* n_conc; /* PDR only, else NULL */
    const uint32_t* n_disc; /* PDR only, else NULL */
} mth_site_rows;

typedef struct {          /* pm.rs:53-60, me.rs:57-65; sorted by (tid,p1,p2,p3,p4) (reference order is unspecified) */
    int64_t n;
    const int32_t* tid;
    const int32_t* p1;
    const int32_t* p2;
    const int32_t* p3;
    const int32_t* p4;
    const float* value;
    const uint32_t* counts; /* n*16 when MTH_FLAG_QUARTET_COUNTS, else NULL */
} mth_quartet_rows;

typedef struct {          /* lpmd.rs:7-15,51-55 (reference counters are i32; these are int64) */
    int64_t n_read, n_valid_read, n_conc, n_disc;
    float lpmd;
} mth_lpmd_result;

typedef struct {          /* lpmd.rs:89-122 (--pairs), sorted by (tid,pos1,pos2) */
    int64_t n;
    const int32_t* tid;
    const int32_t* pos1;
    const int32_t* pos2;
    const float* lpmd;
    const int32_t* n_conc;
    const int32_t* n_disc;
} mth_pair_rows;

typedef struct {
    mth_site_rows pdr, mhl, fdrp, qfdrp;
    mth_quartet_rows pm, me;
    mth_lpmd_result lpmd;
    mth_pair_rows lpmd_pairs;
} mth_results;

#define MTH_MAX_KERNEL_STATS 48
typedef struct {
    int64_t n_reads, n_cpg, n_sites, n_regions;
    int64_t kernel_launches;      /* engine kernels launched since create/reset */
    int64_t h2d_bytes, d2h_bytes; /* bytes copied since create/reset */
    int64_t fdrp_pair_ops;        /* read pairs compared by the FDRP / qFDRP kernels (sum over closed segments of n(n-1)/2) */
    int64_t fallback_sites_mhl, fallback_sites_fdrp; /* MTH_FLAG_PROFILE: sites the tile kernels handed to the per-site kernels */
    int32_t max_ref_span;         /* longest end-start+1 seen */
    int32_t pdr_path;             /* 0 none, 1 scatter (no flush possible), 2 gather everywhere (MTH_FLAG_FORCE_GATHER), 3 scatter + gather on hazard sites */
    int32_t n_kernel_stats;
    struct { char name[32]; int64_t launches; double ms; } kernel[MTH_MAX_KERNEL_STATS]; /* MTH_FLAG_PROFILE */
} mth_stats;

typedef struct mth_ctx mth_ctx;

/* n_ref / ref_len: the BAM reference list (bamutil.rs:13-25), needed to lay contigs out in the device coordinate. */
int mth_ctx_create(mth_ctx** out, int device, const mth_params* params, int32_t n_ref, const int64_t* ref_len);
int mth_ctx_destroy(mth_ctx* ctx);
/* Use `cuda_stream` (a cudaStream_t) for all kernels instead of the context's own stream; NULL restores it. */
int mth_set_stream(mth_ctx* ctx, void* cuda_stream);
/* Asynchronous: enqueues the host->device copy on the copy stream and the ingest kernels behind it. */
int mth_submit(mth_ctx* ctx, const mth_batch* batch);
/* Capacity hint: pre-allocates the arena for about n_reads reads / n_cpg CpG calls so that a streaming host does not
 * pay for re-allocations while batches arrive.  Never required; clamped to the free device memory. */
int mth_reserve(mth_ctx* ctx, int64_t n_reads, int64_t n_cpg);
/* Same as mth_submit for the compact wire format (host or device memory). */
int mth_submit_compact(mth_ctx* ctx, const mth_batch_compact* batch);
/* `--cpg-set` on the device (replaces get_target_cpgs + filter_isin, readutil.rs:347-374 and :87-95): the host hands over the BED
 * file's (tid, pos) pairs once, before the first batch, and ships its batches UNFILTERED; the engine drops every call that is not
 * in the set before anything else is computed from a read — exactly where the reference filters.  Entries may come in any order
 * and may repeat.  mth_clear_cpg_set returns to unfiltered operation.  Both are only legal before the first submit / after
 * mth_reset (MTH_ERR_STATE otherwise). */
int mth_set_cpg_set(mth_ctx* ctx, int64_t n, const int32_t* tid, const int32_t* pos);
int mth_clear_cpg_set(mth_ctx* ctx);
/* Host reads that carried no CpG call were dropped before submit: only LPMD's n_read counts them (lpmd.rs:176). */
int mth_add_skipped_reads(mth_ctx* ctx, int64_t n_reads, int64_t n_reads_mapq_ok);
/* Closes the input, runs the measure kernels, brings the rows back (unless KEEP_ON_DEVICE) and synchronises. */
int mth_finish(mth_ctx* ctx, mth_results* out);
/* Device-resident view of the last results (same struct, device pointers). */
int mth_results_device(mth_ctx* ctx, mth_results* out);
/* int64[4] on the device: n_read, n_valid_read, n_conc, n_disc — what mth_allreduce sums over the ranks. */
int mth_lpmd_counters_device(mth_ctx* ctx, void** dev_ptr);
/* Recompute the LPMD scalar of the last results from (all-reduced) device counters. */
int mth_lpmd_refresh(mth_ctx* ctx, mth_lpmd_result* out);
int mth_reset(mth_ctx* ctx);        /* forget all input and results, keep allocations */
int mth_sync(mth_ctx* ctx);         /* wait for everything enqueued so far */
/* Wait only for the host->device copies of the batches submitted so far: after it returns their host arrays may be
 * reused (a streaming host keeps a small ring of pinned batches), while the ingest kernels keep running. */
int mth_sync_copies(mth_ctx* ctx);
int mth_get_stats(mth_ctx* ctx, mth_stats* out);
const char* mth_last_error(mth_ctx* ctx); /* ctx may be NULL: last create error */

/* ---- multi-GPU (north_star: "shard by contig/position bin ... a single NCCL all-reduce at the end") ---------
 * The reference has nothing to bind here (single process, single thread); a multi-GPU host shards the genome by contig or
 * position bin (reads of a bin + a halo of MTH_META_HALO copies, see mth_batch), runs one context per GPU and joins the
 * contexts with ONE collective after mth_finish: the sum of LPMD's four int64 counters (lpmd.rs:176-191 accumulates them
 * over the whole file).  NCCL is loaded at run time (dlopen of libnccl.so.2: the copy already in the process, e.g.
 * PyTorch's, else the system one); a host that never calls these functions does not need NCCL installed.
 *   multi-process (one rank per GPU): rank 0 calls mth_comm_unique_id and ships the 128 bytes to the other ranks by any
 *     means (MPI, torch.distributed, a file); every rank calls mth_comm_init_rank, then mth_allreduce after mth_finish.
 *   single process, N contexts: mth_comm_init_all once, then mth_allreduce_group after every context has finished.
 * After the all-reduce every context reports the global LPMD (mth_lpmd_refresh, mth_lpmd_counters_device). */
#define MTH_COMM_ID_BYTES 128
int mth_comm_unique_id(void* id128);                                   /* ncclGetUniqueId */
int mth_comm_init_rank(mth_ctx* ctx, int n_ranks, int rank, const void* id128); /* ncclCommInitRank on the context's device */
int mth_comm_init_all(mth_ctx** ctxs, int n);                          /* ncclCommInitAll over the contexts' devices */
int mth_allreduce(mth_ctx* ctx);                                       /* ncclAllReduce(sum, int64) on the compute stream */
int mth_allreduce_group(mth_ctx** ctxs, int n);                        /* the same for N contexts of one process (group call) */
int mth_comm_destroy(mth_ctx* ctx);                                    /* also done by mth_ctx_destroy */
int mth_comm_n_ranks(mth_ctx* ctx);                                    /* 0 without a communicator */

void* mth_host_alloc(size_t bytes); /* pinned host memory (cudaHostAlloc) or NULL */
void mth_host_free(void* p);
int mth_device_count(void);
const char* mth_version(void);

/* ---------------------------------------------------------------------------------------------------------------------
 * `tag`: Bismark XM strings synthesised from read sequence + reference genome (replaces determine_xm_tag_string,
 * src/tag.rs:130-384, called per record from tag::run, src/tag.rs:408-424).  The genome lives in HBM (one upper-cased
 * byte per base, every contig of the header), a batch ships CIGAR + 4-bit SEQ as a BAM record stores them, one warp per
 * read builds the aligned read/reference columns and classifies every reference C by its 3-base context.
 * ------------------------------------------------------------------------------------------------------------------- */
typedef struct mth_genome mth_genome;

/* per-read status of mth_tag: where the reference panics, the read gets a non-zero status and an empty tag */
enum {
    MTH_TAG_OK = 0,
    MTH_TAG_NO_COMPLEMENT = 1, /* reverse-strand read with a base outside the complement map, e.g. '=' (tag.rs:23) */
    MTH_TAG_NO_CONTEXT = 2,    /* look-ahead over deletions found no second base (tag.rs:301) */
    MTH_TAG_BAD_CONTIG = 3,    /* tid outside the genome, or contig not loaded (tag.rs:152) */
    MTH_TAG_PAST_END = 4,      /* alignment ends past the contig (tag.rs:167-172) */
    MTH_TAG_SHORT_SEQ = 5      /* SEQ shorter than the CIGAR's M + I bases */
};

typedef struct mth_tag_batch {
    int64_t n_reads;
    const int32_t* tid;        /* [n_reads] */
    const int32_t* pos;        /* [n_reads] 0-based leftmost position (Record::reference_start) */
    const uint8_t* rc;         /* [n_reads] 1: work on the reverse complement (tag.rs:141-144; need_reverse_complement for pairs) */
    const int32_t* l_seq;      /* [n_reads] number of bases in SEQ */
    const uint32_t* cigar_off; /* [n_reads + 1] offsets into cigar[] */
    const uint32_t* cigar;     /* BAM encoding: len << 4 | op, op index into "MIDNSHP=X" */
    const uint64_t* seq_off;   /* [n_reads + 1] byte offsets into seq4[] */
    const uint8_t* seq4;       /* BAM encoding: two bases per byte, high nibble first, codes of "=ACMGRSVTWYHKDBN" */
} mth_tag_batch;

typedef struct mth_tag_result {
    int64_t n_reads;
    int64_t n_failed;          /* reads with status != MTH_TAG_OK */
    const uint64_t* xm_off;    /* [n_reads + 1] read r's tag is xm[xm_off[r] .. xm_off[r] + xm_len[r]) */
    const uint32_t* xm_len;    /* [n_reads] (can be shorter than the M + I bases: tag.rs:300-331 has no final else) */
    const uint8_t* xm;         /* tag characters, not NUL-terminated */
    const uint8_t* status;     /* [n_reads] MTH_TAG_* */
} mth_tag_result;

/* ref_len: the header's @SQ LN values (tag.rs:58-77).  Contigs are loaded one by one, any letter case (upper-cased on the
 * device like tag.rs:162).  `len` may differ from ref_len[tid] only by being longer (extra bases are ignored). */
int mth_genome_create(mth_genome** out, int device, int32_t n_ref, const int64_t* ref_len);
int mth_genome_set_contig(mth_genome* g, int32_t tid, const uint8_t* seq, int64_t len);
/* Tags one batch; result arrays are pinned host memory owned by the genome until the next mth_tag / destroy. */
int mth_tag(mth_genome* g, const mth_tag_batch* batch, mth_tag_result* out);
/* Device time of the last mth_tag's kernel (CUDA events around the launch), milliseconds; < 0 if none ran yet. */
double mth_genome_last_kernel_ms(mth_genome* g);
int mth_genome_destroy(mth_genome* g);
const char* mth_genome_last_error(mth_genome* g); /* g may be NULL: last create error */

/* ---------------------------------------------------------------------------------------------------------------------
 * BGZF inflate + BAM record decode on the device.  Replaces, for BAM input on one GPU, the per-record host
 * work of the reference: htslib's bgzf inflate + bam_read1 behind bam::Reader::records() (bamutil.rs:4-11) and BismarkRead::new
 * + get_cpgs (readutil.rs:24-53, 323-345).  The host walks the BGZF member headers (18 bytes per <= 64 KiB member) and hands the
 * COMPRESSED bytes over window by window; the decoder returns device-resident mth_batch runs (mem_kind 2, one per contig
 * present in the window) for mth_submit.  mem_kind 2 = device memory that is only valid until the decoder's next window: the
 * engine copies it (never borrows it); call mth_sync_copies before the next mth_bamdec_window.
 * ------------------------------------------------------------------------------------------------------------------- */
typedef struct mth_bamdec mth_bamdec;
typedef struct {
    uint64_t offset;   /* first byte of the member's raw DEFLATE payload within the compressed window */
    uint32_t size;     /* payload bytes (member size - header - 8) */
    uint32_t isize;    /* uncompressed size (the member's ISIZE field) */
    uint32_t crc;      /* CRC-32 of the uncompressed bytes (the member's CRC32 field) */
    uint32_t flags;    /* bit 0: verify `crc` on the device after inflating (what htslib does for every block) */
} mth_bgzf_member;
typedef struct {
    int64_t n_records;                    /* records that END in this window (a record cut by the window's end moves to the next) */
    int64_t n_dropped, n_dropped_mapq_ok; /* records not shipped: no retained CpG call, or (lpmd_order) mapq < min_qual; LPMD still counts them */
    int32_t n_runs;                       /* contig runs among the kept reads */
    int32_t max_cpgs;
    const mth_batch* runs;                /* n_runs batches in DEVICE memory (mem_kind 2), valid until the next window / destroy */
    int64_t max_span;
    int64_t bad_record;                   /* -1, else the first record (index within the window) without a usable XM:Z tag (readutil.rs:45-51) */
    int32_t bad_is_corrupt;               /* that record does not parse at all */
    int32_t reserved;
    uint64_t uncompressed_bytes;
    double ms_inflate, ms_boundaries, ms_decode; /* device time per stage, cumulative since create */
    int64_t chain_repairs;                /* windows whose speculative record chain needed the sequential repair */
} mth_bamdec_result;
/* lpmd_order / min_qual: lpmd.rs:176-181 tests mapq BEFORE it builds the read (low-mapq records are only counted and may lack XM). */
int mth_bamdec_create(mth_bamdec** out, int device, int32_t n_ref, const int64_t* ref_len, uint32_t lpmd_order, uint32_t min_qual);
/* Uploads the compressed bytes of an upcoming window into staging slot 0 or 1 (blocking; any host memory, e.g. the memory-mapped
 * file).  May be called from a second thread while mth_bamdec_window works on the other slot. */
int mth_bamdec_stage(mth_bamdec* d, int slot, const uint8_t* comp, size_t comp_bytes);
/* comp: host memory holding the window's compressed bytes, or NULL: then comp_bytes is the staging slot (0 / 1) filled by
 * mth_bamdec_stage.  skip: bytes at the start of the FIRST window's output that are the BAM header.
 * last != 0: no window follows — a trailing partial record is an error. */
int mth_bamdec_window(mth_bamdec* d, const uint8_t* comp, size_t comp_bytes, const mth_bgzf_member* members, int64_t n_members,
                      uint64_t skip, int last, mth_bamdec_result* out);
int mth_bamdec_destroy(mth_bamdec* d);
const char* mth_bamdec_last_error(mth_bamdec* d); /* d may be NULL: last create error */
/* The inflate kernel alone (tests, profiles): members of `comp` -> out_host, concatenated; per-member status (0 = ok). */
int mth_bgzf_inflate(int device, const uint8_t* comp, size_t comp_bytes, const mth_bgzf_member* members, int64_t n, uint8_t* out_host,
                     size_t out_bytes, int32_t* status_host, double* kernel_ms);

/* Seeded replacement draw used for reservoir sampling: returns j in 1..=total (documented in DESIGN.md). */
uint32_t mth_reservoir_draw(uint64_t seed, int32_t tid, int32_t pos, uint32_t total);

#ifdef __cplusplus
}
#endif
#endif /* METHEOR_B200_H */
