// kernels.h — host-callable launchers of the sm_100a kernels (definitions in k_*.cu).
#pragma once
#include "common.cuh"
#include "../../include/metheor_b200.h"

namespace mth {

// Every launcher returns the number of kernels it launched (for mth_stats.kernel_launches).

// ---- ingest (k_ingest.cu) -------------------------------------------------------------------
int launch_add_u32(uint32_t* a, int64_t n, uint32_t add, cudaStream_t s);
int launch_iota_u32(uint32_t* a, int64_t n, uint32_t base, cudaStream_t s);
int launch_add_i32(int32_t* a, int64_t n, int32_t add, cudaStream_t s);
struct IngestArgs {
    ReadsView rv;             // region-wide view (already contains the batch)
    int64_t r0, n;            // reads [r0, r0+n) are the new batch
    int64_t i0;               // first CpG call of the batch
    const uint16_t* cpg_rel;  // device, REGION-wide like rv.cpg_pos (same indices), or nullptr; 16-byte aligned base
    unsigned long long* bitmap;  // bit (p+1) set for every CpG position p seen
    int32_t lin_lo, lin_hi;   // valid position range of this contig in device coordinates: [lin_lo-1, lin_hi)
    int do_lpmd;
    mth_lpmd_params lpmd;
    int do_pdr;               // also classify reads for PDR (CF_PDR_C / CF_PDR_D in call_flags)
    mth_pdr_params pdr;
    int do_pm, do_me;         // mapq filter of pm.rs:111 / me.rs:115 -> CF_PM_OK / CF_ME_OK
    uint32_t pm_min_qual, me_min_qual;
    uint8_t* call_flags;      // region-wide, one byte per CpG call (common.cuh CF_*)
    RegionScalars* sc;
};
int launch_ingest(const IngestArgs& a, cudaStream_t s);

// ---- compact wire format -> SoA arena (k_expand.cu) ------------------------------------------
struct ExpandArgs {
    // compact batch (device copies)
    int64_t n;
    const int32_t* start; const uint16_t* span; const uint8_t* mapq; const uint8_t* n_cpg8; const uint8_t* flags;
    const uint16_t* cpg_delta; const uint8_t* meth_bits; const uint16_t* rel_exc;
    uint32_t bit_base;        // bit of meth_bits[0] that belongs to the first call (pieces of a batch start mid-byte)
    int64_t n_calls, n_delta8, n_delta16, n_rel, n_start_exc;  // declared sizes: a batch whose per-read counts disagree is reported, never read out of bounds
    uint32_t enc;             // MTH_CENC_* dense encodings (whole batches only)
    const uint16_t* start_off16; const int32_t* blk_start; const int32_t* start_exc;
    const uint8_t* cpg_delta8; const uint32_t* blk_call_off;
    // destination: region arena, reads from r0, calls from i0, meth words from w0
    int64_t r0, i0, w0;
    int32_t lin_off;
    int32_t* start_out; int32_t* end_out; uint32_t* meta_out; uint32_t* off_out; int32_t* pos_out; uint16_t* rel_out;  // rel_out: nullptr unless LPMD
    uint64_t* meth_out; uint32_t* moff_out;  // moff_out: nullptr unless the region already uses meth_off
    uint32_t* block_calls; uint32_t* block_rel;  // scratch, one entry per 256 reads
    uint32_t* err;
};
int launch_expand(const ExpandArgs& a, unsigned long long* total_scratch, cudaStream_t s);
int launch_scan_sums(uint32_t* sums, int64_t n, unsigned long long* total, cudaStream_t s);  // single block, in place

// ---- --cpg-set on the device (k_cpgset.cu) ------------------------------------------------------
// bits (lin_off + p + 1) of `bitmap` for the set positions p of one contig
int launch_cpgset_mark(const int32_t* pos, int64_t n, int32_t lin_off, int32_t len, unsigned long long* bitmap, cudaStream_t s);
// Drops the calls of reads [r0, r0 + n) (calls from i0 on) that are not in the set: methylation words squeezed in place, kept
// positions / query indices written to pos_tmp / rel_tmp (batch-relative), cpg_off rewritten, *total = kept calls (device).
int launch_cpgset_filter(int64_t r0, int64_t n, int64_t i0, uint32_t* off, const int32_t* pos, const uint16_t* rel, uint64_t* meth,
                         const uint32_t* meth_off, const unsigned long long* set_bitmap, int64_t set_words, uint32_t* kept, uint32_t* scan_scratch,
                         unsigned long long* total, int32_t* pos_tmp, uint16_t* rel_tmp, cudaStream_t s);

// ---- site dictionary + scans (k_sites.cu) ---------------------------------------------------
// phase 1: popcount per 1024-word block -> block_sums, scanned in place; total -> sc->n_sites
int launch_sites_count(const unsigned long long* bitmap, int64_t n_words, uint32_t* block_sums, RegionScalars* sc,
                       cudaStream_t s);
// phase 2: word_prefix[w] = #sites before word w; site_pos[rank] = position
int launch_sites_emit(const unsigned long long* bitmap, int64_t n_words, const uint32_t* block_sums,
                      uint32_t* word_prefix, int32_t* site_pos, cudaStream_t s);
// in-place exclusive scan of a[0..n) (u32); total written to *total (device, u64). scratch >= ceil(n/2048)+1 u32
int launch_exclusive_scan_u32(uint32_t* a, int64_t n, uint32_t* scratch, unsigned long long* total, cudaStream_t s);

// ---- PDR (k_pdr.cu) -------------------------------------------------------------------------
int launch_pdr_scatter(const int32_t* cpg_pos, const uint8_t* call_flags, int64_t n_calls, const unsigned long long* bitmap,
                       int64_t n_words, const uint32_t* word_prefix, const RegionScalars* sc, uint32_t* cnt2, cudaStream_t s);
int launch_pdr_gather(const ReadsView& rv, const int32_t* site_pos, int64_t C, const RegionScalars* sc,
                      uint32_t* cnt2, mth_pdr_params prm, const uint8_t* only, cudaStream_t s);  // only != nullptr: flagged sites
// hazard[rank] = 1 for the sites a PDR flush can touch (k_pdr.cu): the scatter counts are exact everywhere else
int launch_pdr_hazard(const ReadsView& rv, const unsigned long long* bitmap, const uint32_t* word_prefix, uint8_t* hazard,
                      cudaStream_t s);
// rowcnt[s] = 1 iff site s yields a row
int launch_pdr_rowcnt(const uint32_t* cnt2, int64_t C, uint32_t min_depth, uint32_t* rowcnt, cudaStream_t s);
struct SiteRowsDev { int32_t* tid; int32_t* pos; float* value; uint32_t* n_conc; uint32_t* n_disc; };
int launch_pdr_emit(const uint32_t* cnt2, const uint32_t* rowoff, const int32_t* site_pos, int64_t C, uint32_t min_depth,
                    ContigTable ct, SiteRowsDev rows, int64_t row_base, cudaStream_t s);

// ---- MHL (k_mhl.cu) -------------------------------------------------------------------------
// value[s], rowcnt[s] in {0,1}
int launch_mhl(const ReadsView& rv, const int32_t* site_pos, int64_t C, const RegionScalars* sc, mth_mhl_params prm,
               float* value, uint32_t* rowcnt, const uint8_t* only, cudaStream_t s);  // only != nullptr: just the flagged sites
// thread-per-site form (k_mhl_site.cu): takes every site whose tile fits shared memory, flags the rest in fallback[] (C bytes,
// zeroed by the caller) for launch_mhl(..., only = fallback)
int launch_mhl_site(const ReadsView& rv, const int32_t* site_pos, int64_t C, const RegionScalars* sc, mth_mhl_params prm,
                    float* value, uint32_t* rowcnt, uint8_t* fallback, cudaStream_t s);
// generic: rows for sites with rowcnt (pre-scan value kept in `flag`), value from dense array
int launch_site_emit(const float* value, const uint32_t* rowoff, uint64_t n_rows_region, const int32_t* site_pos,
                     int64_t C, ContigTable ct, SiteRowsDev rows, int64_t row_base, cudaStream_t s);
int gather_grid(int64_t C);  // grid size for the warp-per-site kernels
int launch_count_flags(const uint8_t* flags, int64_t n, unsigned long long* total, cudaStream_t s);  // *total += #non-zero bytes

// ---- PM / ME (k_quartet.cu) -----------------------------------------------------------------
struct QuartetRowsDev { int32_t* tid; int32_t* p1; int32_t* p2; int32_t* p3; int32_t* p4; float* value; uint32_t* counts; };
// Streaming per-call pass (k_quartet.cu): every call that starts a quartet inside its read is an observation (site, slot,
// pattern) when the quartet is CANONICAL or skips exactly one dictionary site (4 slots per site); it is counted in
// qcnt[site][slot] and appended to the observation list.  Anything else sets mixed[site] and the site is left to the gather
// kernels below.  ok_bit = CF_PM_OK or CF_ME_OK.
int launch_quartet_scatter(const int32_t* cpg_pos, const uint8_t* call_flags, int64_t n_calls, const unsigned long long* bitmap,
                           int64_t n_words, const uint32_t* word_prefix, const RegionScalars* sc, uint32_t ok_bit, uint32_t* qcnt,
                           uint8_t* mixed, uint32_t* obs_site, uint8_t* obs_vp, unsigned long long* obs_n, cudaStream_t s);
// rows per canonical site = slots with count >= min_depth; mixed sites get 0 (overwritten by the gather count)
int launch_quartet_canon_count(const uint32_t* qcnt, const uint8_t* mixed, int64_t C, uint32_t min_depth, uint32_t* rowcnt, cudaStream_t s);
// observation list -> 16-bin histogram per output row (hrows[row][16], zeroed by the caller; row = rowoff[site] + slot rank)
int launch_quartet_hist(const uint32_t* obs_site, const uint8_t* obs_vp, const unsigned long long* obs_n, int64_t n_obs_max,
                        const uint32_t* qcnt, const uint8_t* mixed, const uint32_t* rowoff, uint32_t min_depth, uint32_t* hrows,
                        cudaStream_t s);
// rows of the canonical sites from the row histograms; kind 0 = PM into rows_a, 1 = ME into rows_a, 2 = PM into rows_a AND ME
// into rows_b (identical thresholds: identical rows)
int launch_quartet_canon_emit(const uint32_t* qcnt, const uint8_t* mixed, const uint32_t* hrows, const int32_t* site_pos, int64_t C,
                              uint32_t min_depth, int kind, const uint32_t* rowoff, const float* me_lut, int me_lut_max, ContigTable ct,
                              QuartetRowsDev rows_a, int64_t base_a, QuartetRowsDev rows_b, int64_t base_b, cudaStream_t s);
// gather pass 1: rowcnt[s] = number of quartets starting at site s whose depth >= min_depth.  `mixed` != nullptr restricts
// both gather passes to the sites it flags (the others keep what the canonical kernels wrote).
// the sites flagged in mixed[] as a list (any order) + its length, for the two gather passes below
int launch_quartet_mixed_list(const uint8_t* mixed, int64_t C, uint32_t* list, unsigned long long* n_list, cudaStream_t s);
int launch_quartet_count(const ReadsView& rv, const int32_t* site_pos, int64_t C, const RegionScalars* sc,
                         mth_quartet_params prm, const uint8_t* mixed, const uint32_t* mixed_list, const unsigned long long* mixed_n,
                         uint32_t* rowcnt, cudaStream_t s);
// pass 2: write the rows (sorted by key within a site) at rowoff[s]; kind 0 = PM, 1 = ME, 2 = PM into rows AND ME into rows_b
int launch_quartet_emit(const ReadsView& rv, const int32_t* site_pos, int64_t C, const RegionScalars* sc,
                        mth_quartet_params prm, int kind, const uint8_t* mixed, const uint32_t* mixed_list, const unsigned long long* mixed_n,
                        const uint32_t* rowoff, const float* me_lut,
                        int me_lut_max, ContigTable ct, QuartetRowsDev rows, int64_t row_base, QuartetRowsDev rows_b, int64_t row_base_b,
                        cudaStream_t s);

// ---- LPMD --pairs (k_pairs.cu) ----------------------------------------------------------------
struct PairRowsDev { int32_t* tid; int32_t* pos1; int32_t* pos2; float* lpmd; int32_t* n_conc; int32_t* n_disc; };
int launch_lpmd_pairs_count(const ReadsView& rv, const uint16_t* rel, const int32_t* site_pos, int64_t C,
                            const RegionScalars* sc, mth_lpmd_params prm, uint32_t* rowcnt, cudaStream_t s);
int launch_lpmd_pairs_emit(const ReadsView& rv, const uint16_t* rel, const int32_t* site_pos, int64_t C,
                           const RegionScalars* sc, mth_lpmd_params prm, const uint32_t* rowoff, ContigTable ct,
                           PairRowsDev rows, int64_t row_base, cudaStream_t s);

// ---- FDRP / qFDRP (k_fdrp.cu) ---------------------------------------------------------------
int launch_fdrp(const ReadsView& rv, const int32_t* site_pos, int64_t C, RegionScalars* sc, mth_fdrp_params prm,
                int quantitative, uint64_t seed, ContigTable ct, void* scratch, size_t scratch_bytes, float* value,
                uint32_t* rowcnt, float* value_q, uint32_t* rowcnt_q, const uint8_t* only, cudaStream_t s);  // only != nullptr: just the flagged sites
// quantitative: 0 = FDRP, 1 = qFDRP, 2 = BOTH from one pile + one pair loop (fdrp.rs:124-145 and qfdrp.rs:137-157 share the
// pile, the overlap test and the Hamming distance): FDRP -> value / rowcnt, qFDRP -> value_q / rowcnt_q
// tile form (k_fdrp_tile.cu): takes every site whose tile fits shared memory, flags the rest in fallback[] (C bytes, zeroed by
// the caller) for launch_fdrp(..., only = fallback)
int launch_fdrp_tile(const ReadsView& rv, const int32_t* site_pos, int64_t C, const unsigned long long* bitmap, int64_t n_words,
                     const uint32_t* word_prefix, RegionScalars* sc, mth_fdrp_params prm, int quantitative, uint64_t seed,
                     ContigTable ct, float* value, uint32_t* rowcnt, float* value_q, uint32_t* rowcnt_q, uint8_t* fallback,
                     cudaStream_t s);
size_t fdrp_scratch_bytes(mth_fdrp_params prm, int quantitative);

}  // namespace mth
