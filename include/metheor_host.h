/* metheor_host.h — C API of the host half of the drop-in (libmetheor_host.so + the `metheor` binary).
 *
 * The reference is one Rust crate: CLI (src/main.rs:15-122, src/lib.rs:24-231) -> per-measure compute()
 * (src/pdr.rs:82, lpmd.rs:125, mhl.rs:101, pm.rs:63, me.rs:68, fdrp.rs:148, qfdrp.rs:160) -> BAM decode
 * (src/bamutil.rs:4-25 over rust-htslib, src/readutil.rs:24-53,323-374) -> TSV rows.  Rust is not available in
 * this image, so the host is C++17 over zlib; it decodes BGZF/BAM (or SAM text) on all host cores into the pinned
 * structure-of-arrays batches of include/metheor_b200.h, streams them to the GPU engine (libmetheor_b200.so) and
 * writes the reference's TSV formats.  A Rust host would call the same engine ABI.
 *
 * Plain C types only.  Functions return 0 on success; on failure they return the process exit status the
 * reference would have produced (101 = Rust panic, 2 = clap usage error) and fill `err`.
 */
#ifndef METHEOR_HOST_H
#define METHEOR_HOST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    MTHH_PDR = 0, MTHH_LPMD, MTHH_MHL, MTHH_PM, MTHH_ME, MTHH_FDRP, MTHH_QFDRP
} mthh_measure;

/* One reference subcommand invocation: same fields and defaults as src/lib.rs:24-231. */
typedef struct {
    int32_t measure;        /* mthh_measure */
    const char* input;      /* -i/--input  : BAM or SAM (bamutil.rs:4-11 auto-detects) */
    const char* output;     /* -o/--output */
    const char* cpg_set;    /* -c/--cpg-set BED or NULL (readutil.rs:347-374) */
    const char* pairs;      /* lpmd -p/--pairs or NULL (lpmd.rs:89-122) */
    uint32_t min_depth;     /* -d */
    uint32_t min_cpgs;      /* -p (pdr, mhl) */
    uint32_t min_qual;      /* -q */
    uint32_t max_depth;     /* -D (fdrp, qfdrp) */
    int32_t min_overlap;    /* -l (fdrp, qfdrp) */
    int32_t min_distance;   /* -m (lpmd) */
    int32_t max_distance;   /* -M (lpmd) */
    /* engine options (not in the reference; long flags only, no collision with -i -o -d -p -q -c -D -l -m -M -g) */
    int32_t device;         /* --device N : first GPU to use (default 0) */
    int32_t n_gpus;         /* --gpus N   : the genome is sharded over N GPUs of this node (default 1): position bins + halo */
    int32_t shard_contigs;  /* --shard contigs : shard whole contigs instead of position bins (default 0 = bins) */
    int32_t decode_host;    /* --decode host   : inflate + decode BAM records on the host cores instead of the GPU (default 0: GPU for
                               BAM input on one GPU; SAM text and --gpus > 1 always decode on the host) */
    int32_t out_format;     /* --format tsv | tsv.gz | bedgraph | bedgraph.gz : bit 0 = BGZF-compressed output, bit 1 = bedGraph
                               (chrom start end value; per-CpG measures only).  Default 0: the reference's plain TSV */
    const char* region;     /* --region chr[:beg-end] (1-based, inclusive like samtools) or NULL: only the rows of sites inside the
                               region are computed and written — identical to the same rows of a whole-file run; with a .bai next
                               to the BAM the decoder seeks to the region through the linear index */
    int32_t threads;        /* --threads N: decode threads, 0 = all cores */
    uint64_t seed;          /* --seed     : reservoir sampling seed once a pile exceeds max_depth */
    const char* stats_json; /* --stats F  : write reads/s, per-stage seconds and kernel stats as JSON, or NULL */
} mthh_options;

void mthh_options_default(mthh_options* o, int32_t measure);

/* Runs one subcommand end to end on the GPU engine.  No CPU fallback: fails if no CUDA device is present. */
int mthh_run(const mthh_options* o, char* err, size_t errcap);

/* argv-level entry used by the `metheor` binary (clap-compatible parsing, help and error texts). */
int mthh_main(int argc, char** argv);

/* ---- decode only (no GPU needed): what BismarkRead::new produces for every record, as SoA ---------------------
 * Used by the CPU test-suite to check the host decoder against the oracle, and by tools that feed mth_submit
 * themselves.  Reads without any retained CpG are kept here (n_cpg == 0) so that counts line up with the file. */
typedef struct {
    int64_t n_reads, n_cpg;
    int32_t n_ref;
    const char* const* ref_name;
    const int64_t* ref_len;
    const int32_t* tid;      /* per read */
    const int32_t* start;
    const int32_t* end;
    const uint32_t* meta;    /* mapq | fwd << 8 */
    const int64_t* cpg_off;  /* n_reads + 1 */
    const int32_t* cpg_pos;
    const uint16_t* cpg_rel;
    const uint8_t* cpg_meth; /* 0/1 per CpG call */
} mthh_decoded;

int mthh_decode_file(const char* path, const char* cpg_set, int32_t threads, mthh_decoded** out, char* err, size_t errcap);
void mthh_decoded_free(mthh_decoded* d);

/* `metheor tag` (reference src/tag.rs:386-443): XM tags synthesised on the GPU from SEQ + CIGAR + the FASTA `genome`,
 * appended as the last aux field, written as SAM text.  threads 0 = all cores; stats_json may be NULL.
 * 0 on success, otherwise the exit status (101 where the reference panics) with the message in `err`. */
int mthh_tag(const char* input, const char* output, const char* genome, int32_t device, int32_t threads,
             const char* stats_json, char* err, size_t errcap);

/* The multi-GPU sharding plan `metheor --gpus N` uses: the linearised genome cut into `world` position bins of
 * equal length (by_contig = 0) or whole contigs, longest first onto the least loaded rank (by_contig = 1).  Writes up to `cap`
 * intervals (rank, tid, lo, hi — hi exclusive) ordered by rank; returns the number of intervals. */
typedef struct { int32_t rank, tid; int64_t lo, hi; } mthh_interval;
int mthh_plan_shards(int32_t n_ref, const int64_t* ref_len, int32_t world, int32_t by_contig, mthh_interval* out, int32_t cap);

/* Rust `{}` formatting of an f32 (shortest round-trip digits, positional, "NaN", "inf", "-0"); returns length. */
int mthh_format_f32(float v, char* buf, int cap);

/* The host's own raw-DEFLATE decoder (host/inflate_fast.cpp), exposed for its tests: decodes in[0,in_len) into exactly
 * out_len bytes; 1 on success, 0 if it rejects the stream (the BGZF reader then falls back to zlib).
 * mthh_zlib_fallbacks: how many BGZF members this process handed to zlib after the fast decoder rejected them. */
int mthh_inflate_raw(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len);
int64_t mthh_zlib_fallbacks(void);
uint32_t mthh_crc32(const uint8_t* p, size_t n); /* the BGZF reader's CRC-32 (PCLMULQDQ folding), == zlib crc32(0, p, n) */

#ifdef __cplusplus
}
#endif
#endif /* METHEOR_HOST_H */
