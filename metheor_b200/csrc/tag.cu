// tag.cu — XM synthesis (`metheor tag`): the genome resident in HBM, one warp per read.
//
// Replaces determine_xm_tag_string (reference src/tag.rs:130-384).  The reference materialises two aligned strings per
// read (read bases / reference bases, '-' for gaps) with two context columns on either side, reverse-complements both for
// reverse-strand reads, walks them once and reverses the tag back.  Here:
//   phase A  the warp walks the CIGAR once and writes the body columns (M, I, D only — tag.rs:186-233 ignores every other
//            op) to an L2-resident scratch line of the read; the four context bases come straight from the genome
//            (N outside the contig, tag.rs:165-173);
//   phase B  lane i classifies target column i (index arithmetic stands in for the reverse complement), the warp
//            compacts the emitted characters in order with a ballot (a column can emit nothing: deletion columns and
//            contexts outside CG/CHG/CHH/unknown, tag.rs:300-331), and reverse-strand tags are reversed in place.
// HBM traffic per read: l_seq/2 B of SEQ + ~l_seq B of genome in, ~l_seq B of tag out; the columns stay in L2.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/metheor_b200.h"

namespace {

constexpr uint32_t FULL = 0xffffffffu;
constexpr int TAG_BLOCK = 256;
#ifndef TAG_MINB
#define TAG_MINB 4  // resident CTAs per SM the register budget is capped for (5 and 6 measured: spills, no gain)
#endif

struct TagArgs {
    int64_t n_reads;
    const int32_t* tid;
    const int32_t* pos;
    const uint8_t* rc;
    const int32_t* l_seq;
    const uint32_t* cigar_off;
    const uint32_t* cigar;
    const uint64_t* seq_off;
    const uint8_t* seq4;
    const uint64_t* col_off;   // [n+1] body columns (M + I + D) per read
    const uint64_t* xm_off;    // [n+1] tag capacity (M + I) per read
    const uint8_t* genome;     // upper-cased bases, contigs back to back
    const int64_t* contig_off; // [n_ref]
    const int64_t* contig_len; // [n_ref] header LN
    const int64_t* loaded_len; // [n_ref] bases actually loaded (-1: contig missing)
    int32_t n_ref;
    uint8_t* col_read;
    uint8_t* col_ref;
    uint8_t* xm;
    uint32_t* xm_len;
    uint8_t* status;
};

// tag.rs:78-99; 0 = not in the map (the reference panics on the HashMap index)
__device__ __forceinline__ uint8_t complement(uint8_t c) {
    switch (c) {
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; case 'N': return 'N';
        case 'M': return 'K'; case 'R': return 'Y'; case 'W': return 'W'; case 'S': return 'S'; case 'Y': return 'R';
        case 'K': return 'M'; case 'V': return 'B'; case 'H': return 'D'; case 'D': return 'H'; case 'B': return 'V';
        case '-': return '-';
        default: return 0;
    }
}

__device__ __forceinline__ bool is_h(uint8_t c) { return c == 'A' || c == 'T' || c == 'C'; }
__device__ __forceinline__ bool is_unknown(uint8_t c) { return c == '-' || c == 'N'; }

// Context of a reference C -> tag character for read base rd, or 0 when the reference emits nothing.  n_ctx: how many
// context characters exist after the C (2 normally, 1 when the look-ahead ran into the end of the columns).
__device__ __forceinline__ uint8_t classify(uint8_t c1, uint8_t c2, int n_ctx, uint8_t rd) {
    uint8_t lower;
    if (c1 == 'G') lower = 'z';                                               // tag.rs:301, 339
    else if (n_ctx == 2 && is_h(c1) && c2 == 'G') lower = 'x';                // CHG, tag.rs:32-34
    else if (n_ctx == 2 && is_h(c1) && is_h(c2)) lower = 'h';                 // CHH, tag.rs:36-41
    else if (is_unknown(c1) || (n_ctx == 2 && is_unknown(c2))) lower = 'u';   // tag.rs:43-52
    else return 0;
    return rd == 'C' ? (uint8_t)(lower - 32) : rd == 'T' ? lower : (uint8_t)'.';
}

// One read, any number of columns: the columns go through the global scratch line of the read (long reads, huge deletions).
__device__ __noinline__ void tag_read_generic(const TagArgs& a, const int64_t r, const int lane) {
    do {
        const int32_t tid = a.tid[r];
        const int64_t start = a.pos[r];
        const bool rc = a.rc[r] != 0;
        const int32_t l_seq = a.l_seq[r];
        const uint32_t c_lo = a.cigar_off[r], c_hi = a.cigar_off[r + 1];
        const uint64_t col0 = a.col_off[r];
        const int64_t B = (int64_t)(a.col_off[r + 1] - col0);
        uint8_t* const xm = a.xm + a.xm_off[r];
        uint8_t st = MTH_TAG_OK;
        if (tid < 0 || tid >= a.n_ref || a.loaded_len[tid] < 0 || start < 0) st = MTH_TAG_BAD_CONTIG;
        if (st) {
            if (lane == 0) { a.status[r] = st; a.xm_len[r] = 0; }
            continue;
        }
        const int64_t chrom = a.contig_len[tid], loaded = a.loaded_len[tid];
        const uint8_t* const g = a.genome + a.contig_off[tid];
        auto ref_at = [&](int64_t p) -> uint8_t { return (p < 0 || p >= chrom || p >= loaded) ? (uint8_t)'N' : g[p]; };
        const uint8_t* const sq = a.seq4 + a.seq_off[r];
        auto base_at = [&](int64_t i) -> uint8_t {
            if (i >= l_seq) return 'N';
            const uint32_t code = (sq[i >> 1] >> ((~i & 1) << 2)) & 15u;
            return (uint8_t)"=ACMGRSVTWYHKDBN"[code];
        };
        uint8_t* const cr = a.col_read + col0;
        uint8_t* const cf = a.col_ref + col0;

        // ---- phase A: body columns ----
        int64_t c0 = 0, ur = 0, uf = 0, ref_span = 0;
        bool unmappable = false;
        for (uint32_t k = c_lo; k < c_hi; k++) {
            const uint32_t v = a.cigar[k];
            const int64_t len = v >> 4;
            const uint32_t op = v & 15u;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref_span += len;  // bam_endpos: M D N = X
            if (op > 2) continue;  // tag.rs:232 `_ => {}`
            if (c0 + len > B) break;  // offsets inconsistent with the CIGAR (rejected on the host; never write outside)
            for (int64_t j = lane; j < len; j += 32) {
                const uint8_t rd = op == 2 ? (uint8_t)'-' : base_at(ur + j);
                const uint8_t rf = op == 1 ? (uint8_t)'-' : ref_at(start + uf + j);
                cr[c0 + j] = rd;
                cf[c0 + j] = rf;
                if (rc && (!complement(rd) || !complement(rf))) unmappable = true;
            }
            c0 += len;
            if (op != 2) ur += len;
            if (op != 1) uf += len;
        }
        const int64_t end = start + (ref_span ? ref_span : 1);
        const uint8_t p0 = ref_at(start - 2), p1 = ref_at(start - 1), s0 = ref_at(end), s1 = ref_at(end + 1);
        if (rc && (!complement(p0) || !complement(p1))) unmappable = true;
        if (end > chrom || (end + 2 < chrom ? end + 2 : chrom) > loaded) st = MTH_TAG_PAST_END;
        else if (ur > l_seq) st = MTH_TAG_SHORT_SEQ;
        else if (__any_sync(FULL, unmappable)) st = MTH_TAG_NO_COMPLEMENT;
        __syncwarp();
        if (st) {
            if (lane == 0) { a.status[r] = st; a.xm_len[r] = 0; }
            continue;
        }

        // ---- phase B: classify target column i, compact in order ----
        const int64_t L = B + 2;
        auto R = [&](int64_t i) -> uint8_t { return i >= B ? (uint8_t)'-' : rc ? complement(cr[B - 1 - i]) : cr[i]; };
        auto F = [&](int64_t i) -> uint8_t {
            if (i < B) return rc ? complement(cf[B - 1 - i]) : cf[i];
            if (rc) return complement(i == B ? p1 : p0);
            return i == B ? s0 : s1;
        };
        uint32_t n_out = 0;
        bool no_context = false;
        for (int64_t base = 0; base < B; base += 32) {
            const int64_t i = base + lane;
            uint8_t ch = 0;
            if (i < B) {
                const uint8_t rd = R(i);
                if (rd == '-') {
                    ch = 0;
                } else if (rd == 'N') {
                    ch = '.';
                } else if (F(i) == 'C') {
                    if ((R(i + 1) == '-' || R(i + 2) == '-') && i != L - 3 && i != L - 4) {  // tag.rs:268-299
                        uint8_t ctx[2] = {0, 0};
                        int found = 0;
                        for (int64_t k = 1; found != 2 && i + k <= L - 1; k++)
                            if (R(i + k) != '-') ctx[found++] = F(i + k);
                        if (found == 0) no_context = true;
                        else ch = classify(ctx[0], ctx[1], found, rd);
                    } else {
                        ch = classify(F(i + 1), F(i + 2), 2, rd);
                    }
                } else {
                    ch = '.';
                }
            }
            const uint32_t m = __ballot_sync(FULL, ch != 0);
            if (ch) xm[n_out + __popc(m & ((1u << lane) - 1u))] = ch;
            n_out += __popc(m);
        }
        if (__any_sync(FULL, no_context)) {
            if (lane == 0) { a.status[r] = MTH_TAG_NO_CONTEXT; a.xm_len[r] = 0; }
            continue;
        }
        __syncwarp();
        if (rc) {  // tag.rs:380-383
            for (uint32_t k = lane; k < n_out / 2; k += 32) {
                const uint8_t x = xm[k], y = xm[n_out - 1 - k];
                xm[k] = y;
                xm[n_out - 1 - k] = x;
            }
        }
        if (lane == 0) { a.status[r] = MTH_TAG_OK; a.xm_len[r] = n_out; }
    } while (false);
}

constexpr int TAG_CAP = 512;     // target columns (body + 2) a warp keeps in shared memory
constexpr int TAG_MAXREF = 512;  // contig tables of up to this many contigs are copied to shared memory

// Per-read scalars, loaded together at the top of the read so that the independent global loads overlap.  (Loading the
// NEXT read's header one iteration ahead was measured: no gain, the kernel is issue-bound, and it cost spills.)
struct ReadHdr {
    uint64_t col_lo, col_hi, seq_off, xm_off;
    int32_t tid, pos, l_seq;
    uint32_t c_lo, c_hi, first_op;
    bool rc;
};
__device__ __forceinline__ ReadHdr load_hdr(const TagArgs& a, int64_t r) {
    ReadHdr h;
    h.col_lo = a.col_off[r]; h.col_hi = a.col_off[r + 1];
    h.seq_off = a.seq_off[r]; h.xm_off = a.xm_off[r];
    h.tid = a.tid[r]; h.pos = a.pos[r]; h.l_seq = a.l_seq[r];
    h.c_lo = a.cigar_off[r]; h.c_hi = a.cigar_off[r + 1];
    h.first_op = h.c_hi > h.c_lo ? a.cigar[h.c_lo] : 0u;
    h.rc = a.rc[r] != 0;
    return h;
}

// Tables and per-warp lines in shared memory, handed to the per-read routines.
struct TagShared {
    const uint8_t *comp, *nt, *ctx, *cls, *rdc, *clsr;
    const int64_t *t_off, *t_len, *t_loaded;
    uint8_t *tr, *tf, *to;  // this warp's target columns (read / reference) and character line
};

// One read by a whole warp, any CIGAR.  Reads of up to TAG_CAP - 2 body columns stay in shared memory: phase A writes
// the TARGET columns (already reverse-complemented for reverse-strand reads, tag.rs:244-257), phase B classifies them
// into a shared character line and counts, phase C writes the tag (reversed for reverse-strand reads, tag.rs:380-383).
__device__ __noinline__ void tag_read_warp(const TagArgs& a, const TagShared& T, const int64_t r, const int lane) {
    uint8_t* const tr = T.tr;
    uint8_t* const tf = T.tf;
    uint8_t* const to = T.to;
    {
    const ReadHdr h = load_hdr(a, r);
    const int64_t B64 = (int64_t)(h.col_hi - h.col_lo);
    if (B64 + 2 > TAG_CAP) {
        tag_read_generic(a, r, lane);
        __syncwarp();
        return;
    }
    const int B = (int)B64;
    const int32_t tid = h.tid;
    const int64_t start = h.pos;
    const bool rc = h.rc;
    const int32_t l_seq = h.l_seq;
    const uint32_t c_lo = h.c_lo, c_hi = h.c_hi;
    if (tid < 0 || tid >= a.n_ref || T.t_loaded[tid] < 0 || start < 0) {
        if (lane == 0) { a.status[r] = MTH_TAG_BAD_CONTIG; a.xm_len[r] = 0; }
        return;
    }
    const int64_t chrom = T.t_len[tid], loaded = T.t_loaded[tid];
    const int64_t lim = chrom < loaded ? chrom : loaded;
    const uint8_t* const g = a.genome + T.t_off[tid];
    const uint8_t* const sq = a.seq4 + h.seq_off;

    // ---- phase A: target columns ----
    int c0 = 0;
    int64_t ur = 0, uf = 0, ref_span = 0;
    bool unmappable = false;
    for (uint32_t k = c_lo; k < c_hi; k++) {
        const uint32_t v = k == c_lo ? h.first_op : a.cigar[k];
        const int len = (int)(v >> 4);
        const uint32_t op = v & 15u;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref_span += v >> 4;
        if (op > 2) continue;
        if (c0 + len > B) break;
        for (int j = lane; j < len; j += 32) {
            uint8_t rd = '-', rf = '-';
            if (op != 2) {
                const int64_t i = ur + j;
                rd = i < l_seq ? T.nt[(sq[i >> 1] >> ((~i & 1) << 2)) & 15u] : (uint8_t)'N';
            }
            if (op != 1) {
                const int64_t p = start + uf + j;
                rf = p < lim ? g[p] : (uint8_t)'N';
            }
            int t = c0 + j;
            if (rc) {
                rd = T.comp[rd];
                rf = T.comp[rf];
                unmappable |= !rd || !rf;
                t = B - 1 - t;
            }
            tr[t] = rd;
            tf[t] = rf;
        }
        c0 += len;
        if (op != 2) ur += len;
        if (op != 1) uf += len;
    }
    const int64_t end = start + (ref_span ? ref_span : 1);
    if (lane < 2) {  // the two context columns behind the body
        int64_t p = rc ? start - 1 - lane : end + lane;
        uint8_t c = (p >= 0 && p < lim) ? g[p] : (uint8_t)'N';
        if (rc) {
            c = T.comp[c];
            unmappable |= !c;
        }
        tr[B + lane] = '-';
        tf[B + lane] = c;
    }
    uint8_t st = MTH_TAG_OK;
    if (end > chrom || (end + 2 < chrom ? end + 2 : chrom) > loaded) st = MTH_TAG_PAST_END;
    else if (ur > l_seq) st = MTH_TAG_SHORT_SEQ;
    else if (__any_sync(FULL, unmappable)) st = MTH_TAG_NO_COMPLEMENT;
    __syncwarp();
    if (st) {
        if (lane == 0) { a.status[r] = st; a.xm_len[r] = 0; }
        __syncwarp();
        return;
    }

    // ---- phase B: classify ----
    const int L = B + 2;
    int n_out = 0;
    bool no_context = false;
    for (int base = 0; base < B; base += 32) {
        const int i = base + lane;
        uint8_t ch = 0;
        if (i < B) {
            const uint8_t rd = tr[i];
            if (rd == '-') {
                ch = 0;
            } else if (rd == 'N') {
                ch = '.';
            } else if (tf[i] == 'C') {
                if ((tr[i + 1] == '-' || tr[i + 2] == '-') && i != L - 3 && i != L - 4) {  // tag.rs:268-299
                    uint8_t ctx[2] = {0, 0};
                    int found = 0;
                    for (int k = 1; found != 2 && i + k <= L - 1; k++)
                        if (tr[i + k] != '-') ctx[found++] = tf[i + k];
                    if (found == 0) no_context = true;
                    else ch = classify(ctx[0], ctx[1], found, rd);
                } else {
                    ch = T.ctx[T.cls[tf[i + 1]] * 12 + T.cls[tf[i + 2]] * 3 + T.rdc[rd]];
                }
            } else {
                ch = '.';
            }
        }
        const uint32_t m = __ballot_sync(FULL, ch != 0);
        if (ch) to[n_out + __popc(m & ((1u << lane) - 1u))] = ch;
        n_out += __popc(m);
    }
    if (__any_sync(FULL, no_context)) {
        if (lane == 0) { a.status[r] = MTH_TAG_NO_CONTEXT; a.xm_len[r] = 0; }
        __syncwarp();
        return;
    }
    __syncwarp();
    // ---- phase C: the tag, in read orientation ----
    uint8_t* const xm = a.xm + h.xm_off;
    for (int k = lane; k < n_out; k += 32) xm[k] = to[rc ? n_out - 1 - k : k];
    if (lane == 0) { a.status[r] = MTH_TAG_OK; a.xm_len[r] = (uint32_t)n_out; }
    __syncwarp();
    }
}

// The common read — CIGAR `<l_seq>M`, at least two bases away from both contig ends — has no gap columns: column k is read
// base k over reference base start + k, its context is the next two reference bases (forward) or the complements of the two
// BEFORE it (reverse strand: reverse-complementing the columns, classifying and reversing the tag back puts the character of
// column B-1-k at position k again), and every column emits one character.  Four lanes share a read (8 reads per warp); a
// lane takes 8 consecutive bases per step: one 32-bit word of packed SEQ, 12 reference bytes from three funnel-shifted words,
// one 8-byte store.  A read that turns out to need the general machinery — a context that emits nothing, a base without a
// complement on the reverse strand (tag.rs:78-99 panics; status NO_COMPLEMENT) — is redone by tag_read_warp.
constexpr int TAG_GROUP = 4;               // lanes per read: 4 x 8 bases = one 32-byte sector of tag per step; 19 steps-of-8 of a 150-base read fill 5 x 4 slots
constexpr int TAG_RPW = 32 / TAG_GROUP;    // reads per warp and step
// nibbles of a 32-bit word of packed SEQ (little endian, base j = high nibble of byte j / 2 for even j) that hold bases 0 .. nb - 1
__constant__ uint32_t c_valid[9] = {0x00000000u, 0x000000F0u, 0x000000FFu, 0x0000F0FFu, 0x0000FFFFu,
                                    0x00F0FFFFu, 0x00FFFFFFu, 0xF0FFFFFFu, 0xFFFFFFFFu};

__global__ void __launch_bounds__(TAG_BLOCK, TAG_MINB) k_tag(const __grid_constant__ TagArgs a) {
    __shared__ uint8_t s_comp[256];
    __shared__ uint8_t s_read[TAG_BLOCK / 32][TAG_CAP + 4];
    __shared__ uint8_t s_ref[TAG_BLOCK / 32][TAG_CAP + 4];
    __shared__ uint8_t s_out[TAG_BLOCK / 32][TAG_CAP];
    __shared__ int64_t s_tab[3][TAG_MAXREF];
    // classify() as three byte look-ups: context class of each of the two bases behind the C (0 'G', 1 A/T/C, 2 '-'/'N',
    // 3 anything else), class of the read base (0 'C', 1 'T', 2 other), and the 4 x 4 x 3 table of tag characters.
    // s_clsr: the context class of the COMPLEMENT of a reference base (reverse strand), 4 = the base has no complement.
    __shared__ uint8_t s_cls[256], s_rdc[256], s_clsr[256], s_ctx[64], s_nt[16];
    for (int c = threadIdx.x; c < 256; c += blockDim.x) {
        const uint8_t cc = complement((uint8_t)c);
        s_comp[c] = cc;
        s_cls[c] = c == 'G' ? 0 : is_h((uint8_t)c) ? 1 : is_unknown((uint8_t)c) ? 2 : 3;
        s_clsr[c] = !cc ? 4 : cc == 'G' ? 0 : is_h(cc) ? 1 : is_unknown(cc) ? 2 : 3;
        s_rdc[c] = c == 'C' ? 0 : c == 'T' ? 1 : 2;
    }
    if (threadIdx.x < 48) {
        static const char probe[4] = {'G', 'A', 'N', 'R'};  // one representative per context class
        const int k1 = threadIdx.x / 12, k2 = (threadIdx.x / 3) & 3, kr = threadIdx.x % 3;
        s_ctx[threadIdx.x] = classify((uint8_t)probe[k1], (uint8_t)probe[k2], 2, kr == 0 ? (uint8_t)'C' : kr == 1 ? (uint8_t)'T' : (uint8_t)'A');
    } else if (threadIdx.x < 64) {
        s_ctx[threadIdx.x] = 0;
    }
    if (threadIdx.x < 16) s_nt[threadIdx.x] = (uint8_t)"=ACMGRSVTWYHKDBN"[threadIdx.x];
    const bool tab_smem = a.n_ref <= TAG_MAXREF;
    if (tab_smem)
        for (int c = threadIdx.x; c < a.n_ref; c += blockDim.x) {
            s_tab[0][c] = a.contig_off[c];
            s_tab[1][c] = a.contig_len[c];
            s_tab[2][c] = a.loaded_len[c];
        }
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    TagShared T;
    T.comp = s_comp; T.nt = s_nt; T.ctx = s_ctx; T.cls = s_cls; T.rdc = s_rdc; T.clsr = s_clsr;
    T.t_off = tab_smem ? s_tab[0] : a.contig_off;
    T.t_len = tab_smem ? s_tab[1] : a.contig_len;
    T.t_loaded = tab_smem ? s_tab[2] : a.loaded_len;
    T.tr = s_read[wib]; T.tf = s_ref[wib]; T.to = s_out[wib];
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int sub = lane / TAG_GROUP, gl = lane % TAG_GROUP;
    for (int64_t r0 = warp * TAG_RPW; r0 < a.n_reads; r0 += n_warps * TAG_RPW) {
        const int64_t r = r0 + sub;
        const bool have = r < a.n_reads;
        bool fast = false, bad = false;
        if (have) {
            const uint32_t c_lo = a.cigar_off[r], c_hi = a.cigar_off[r + 1];
            const int32_t tid = a.tid[r], l_seq = a.l_seq[r];
            const int64_t start = a.pos[r];
            const bool rc = a.rc[r] != 0;
            if (c_hi == c_lo + 1 && l_seq > 0 && tid >= 0 && tid < a.n_ref && start >= 2) {
                const uint32_t op = a.cigar[c_lo];
                const int64_t chrom = T.t_len[tid], loaded = T.t_loaded[tid];
                const int64_t lim = chrom < loaded ? chrom : loaded;  // loaded < 0: contig missing, never fast
                fast = op == ((uint32_t)l_seq << 4) && start + l_seq + 2 <= lim;
                if (fast) {
                    const uint8_t* const sq = a.seq4 + a.seq_off[r];
                    const uint8_t* const g = a.genome + T.t_off[tid] + start;
                    uint8_t* const xm = a.xm + a.xm_off[r];  // 8-byte aligned, capacity a multiple of 8 (mth_tag)
                    const int U = (l_seq + 7) >> 3;
                    const uint8_t* const cls2 = rc ? s_clsr : s_cls;
                    const uint32_t c_char = rc ? 'G' : 'C';
                    // class of the read base per 4-bit code, 2 bits each: forward C(2)->0 T(8)->1; reverse G(4)->0 A(1)->1 '='(0)->3
                    const uint32_t kr_lut = rc ? 0xAAAAA8A7u : 0xAAA9AA8Au;
                    for (int u = gl; u < U; u += TAG_GROUP) {
                        const int k0 = u << 3;
                        const int nb = l_seq - k0 < 8 ? l_seq - k0 : 8;
                        // 8 read bases: 4 bytes of packed SEQ at any alignment
                        const uintptr_t sa = (uintptr_t)(sq + 4 * u);
                        const uint32_t* const sw = reinterpret_cast<const uint32_t*>(sa & ~(uintptr_t)3);
                        const uint32_t seqw = __funnelshift_r(__ldg(sw), __ldg(sw + 1), (uint32_t)(sa & 3u) * 8u);
                        // reference bytes k0 - 2 .. k0 + 9
                        const uintptr_t ga = (uintptr_t)(g + k0 - 2);
                        const uint32_t* const gw = reinterpret_cast<const uint32_t*>(ga & ~(uintptr_t)3);
                        const uint32_t gs = (uint32_t)(ga & 3u) * 8u;
                        const uint32_t x0 = __ldg(gw), x1 = __ldg(gw + 1), x2 = __ldg(gw + 2), x3 = __ldg(gw + 3);
                        const uint32_t y[3] = {__funnelshift_r(x0, x1, gs), __funnelshift_r(x1, x2, gs), __funnelshift_r(x2, x3, gs)};
                        // Everything below is the same instruction stream for both strands (the four reads of a warp differ in
                        // strand: a branch on it would run both sides): the strand selects a table half, a constant and offsets.
                        // ... and no branch inside: per base one table look-up and selects.  What needs the general path is
                        // only DETECTED here, generously (a false alarm just sends the read through tag_read_warp).
                        uint32_t cl[12];
                        uint32_t any_cl = 0u;  // class 4 = no complement (reverse-strand table only)
#pragma unroll
                        for (int j = 0; j < 12; j++) {
                            cl[j] = cls2[(y[j >> 2] >> (8 * (j & 3))) & 0xffu];
                            any_cl |= cl[j];
                        }
                        uint32_t out[2] = {0u, 0u};
                        bool odd = (any_cl & 4u) != 0u;
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            // twice the 4-bit code of base j (high nibble of byte j / 2 for even j): the bit offset into kr_lut
                            const int sh = 8 * (j >> 1) + ((j & 1) ? 0 : 4) - 1;
                            const uint32_t rdn2 = (sh >= 0 ? seqw >> sh : seqw << 1) & 30u;
                            const uint32_t f = (y[(j + 2) >> 2] >> (8 * ((j + 2) & 3))) & 0xffu;
                            // context classes: the next two reference bases, or (reverse strand) the complements of the two before
                            const uint32_t k1 = rc ? cl[j + 1] : cl[j + 3], k2 = rc ? cl[j] : cl[j + 4];
                            const uint32_t kr = (kr_lut >> rdn2) & 3u;  // 0: the read shows the C, 1: the T, 2: anything else, 3: '=' on the reverse strand
                            const uint32_t t = s_ctx[k1 * 12u + k2 * 3u + kr];  // 64 entries: classes 4 / 3 index zeros or neighbours, and are flagged
                            const bool is_c = f == c_char && rdn2 != 30u;
                            odd |= is_c && t == 0u;  // a context that emits nothing
                            out[j >> 2] |= (is_c ? t : (uint32_t)'.') << (8 * (j & 3));
                        }
                        if (rc) {  // '=' (code 0) has no complement: any zero nibble among the bases that exist (borrows can only raise false alarms)
                            const uint32_t sv = seqw | ~c_valid[nb];
                            odd |= ((sv - 0x11111111u) & ~sv & 0x88888888u) != 0u;
                        }
                        bad |= odd;
                        *reinterpret_cast<uint2*>(xm + k0) = make_uint2(out[0], out[1]);
                    }
                }
            }
        }
        const uint32_t badm = __ballot_sync(FULL, bad);
        const bool group_bad = ((badm >> (sub * TAG_GROUP)) & ((1u << TAG_GROUP) - 1u)) != 0u;
        if (have && fast && !group_bad && gl == 0) { a.status[r] = MTH_TAG_OK; a.xm_len[r] = (uint32_t)a.l_seq[r]; }
        uint32_t slow = __ballot_sync(FULL, have && gl == 0 && (!fast || group_bad));
        while (slow) {  // the reads that need the general machinery, one after the other by the whole warp
            const int l = __ffs(slow) - 1;
            slow &= slow - 1;
            tag_read_warp(a, T, r0 + l / TAG_GROUP, lane);
            __syncwarp();
        }
    }
}

// to_uppercase of tag.rs:162 (ASCII letters), 16 bases per thread step
__global__ void k_upper(uint8_t* p, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint8_t c = p[i];
        if (c >= 'a' && c <= 'z') p[i] = (uint8_t)(c - 32);
    }
}

std::string g_genome_create_err;

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    bool pinned_host;
    explicit DevBuf(bool host = false) : pinned_host(host) {}
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        release();
        size_t want = n + n / 4 + 64;
        cudaError_t e = pinned_host ? cudaHostAlloc((void**)&p, want * sizeof(T), cudaHostAllocDefault) : cudaMalloc((void**)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want; else p = nullptr;
        return e;
    }
    void release() {
        if (p) { if (pinned_host) cudaFreeHost(p); else cudaFree(p); }
        p = nullptr;
        cap = 0;
    }
};

}  // namespace

struct mth_genome {
    int device = 0;
    cudaStream_t stream = nullptr;
    int32_t n_ref = 0;
    std::vector<int64_t> ref_len, contig_off, loaded;
    uint8_t* d_genome = nullptr;
    int64_t total = 0;
    int64_t *d_contig_off = nullptr, *d_contig_len = nullptr, *d_loaded = nullptr;
    bool tables_dirty = true;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double last_kernel_ms = -1.0;
    std::string err;
    int sm_count = 148;
    // per-batch device buffers
    DevBuf<int32_t> d_tid, d_pos, d_lseq;
    DevBuf<uint8_t> d_rc, d_seq4, d_colr, d_colf, d_xm, d_status;
    DevBuf<uint32_t> d_cigar_off, d_cigar, d_xm_len;
    DevBuf<uint64_t> d_seq_off, d_col_off, d_xm_off;
    // pinned results
    DevBuf<uint64_t> h_xm_off{true}, h_col_off{true};
    DevBuf<uint32_t> h_xm_len{true};
    DevBuf<uint8_t> h_xm{true}, h_status{true};
};

#define G_CUDA(call)                                                                        \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            g->err = std::string(#call) + ": " + cudaGetErrorString(e_);                    \
            return MTH_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)

extern "C" {

const char* mth_genome_last_error(mth_genome* g) { return g ? g->err.c_str() : g_genome_create_err.c_str(); }

int mth_genome_create(mth_genome** out, int device, int32_t n_ref, const int64_t* ref_len) {
    if (!out || n_ref < 0 || (n_ref && !ref_len)) { g_genome_create_err = "mth_genome_create: bad arguments"; return MTH_ERR_INVALID; }
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        g_genome_create_err = "no CUDA device: the metheor_b200 engine has no CPU fallback";
        return MTH_ERR_CUDA;
    }
    if (device < 0 || device >= n_dev) { g_genome_create_err = "mth_genome_create: no such device"; return MTH_ERR_INVALID; }
    mth_genome* g = new mth_genome();
    g->device = device;
    g->n_ref = n_ref;
    int64_t off = 0;
    for (int32_t t = 0; t < n_ref; t++) {
        if (ref_len[t] < 0) { delete g; g_genome_create_err = "mth_genome_create: negative contig length"; return MTH_ERR_INVALID; }
        g->ref_len.push_back(ref_len[t]);
        g->contig_off.push_back(off);
        g->loaded.push_back(-1);
        off += ref_len[t];
    }
    g->total = off;
    auto fail = [&](const char* what, cudaError_t e) {
        g_genome_create_err = std::string(what) + ": " + cudaGetErrorString(e);
        mth_genome_destroy(g);
        return MTH_ERR_CUDA;
    };
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", e);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) g->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    if ((e = cudaMalloc((void**)&g->d_genome, (size_t)off + 64)) != cudaSuccess) return fail("cudaMalloc(genome)", e);
    const size_t tb = sizeof(int64_t) * (size_t)(n_ref ? n_ref : 1);
    if ((e = cudaMalloc((void**)&g->d_contig_off, tb)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc((void**)&g->d_contig_len, tb)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc((void**)&g->d_loaded, tb)) != cudaSuccess) return fail("cudaMalloc", e);
    *out = g;
    return MTH_OK;
}

int mth_genome_set_contig(mth_genome* g, int32_t tid, const uint8_t* seq, int64_t len) {
    if (!g) return MTH_ERR_INVALID;
    if (tid < 0 || tid >= g->n_ref || len < 0 || (len && !seq)) { g->err = "mth_genome_set_contig: bad arguments"; return MTH_ERR_INVALID; }
    G_CUDA(cudaSetDevice(g->device));
    const int64_t n = len < g->ref_len[(size_t)tid] ? len : g->ref_len[(size_t)tid];
    uint8_t* dst = g->d_genome + g->contig_off[(size_t)tid];
    if (n) {
        G_CUDA(cudaMemcpyAsync(dst, seq, (size_t)n, cudaMemcpyHostToDevice, g->stream));
        int64_t blocks = (n + 256 * 16 - 1) / (256 * 16);
        const int64_t cap = (int64_t)g->sm_count * 16;
        if (blocks > cap) blocks = cap;
        k_upper<<<(unsigned)blocks, 256, 0, g->stream>>>(dst, n);
        G_CUDA(cudaGetLastError());
        G_CUDA(cudaStreamSynchronize(g->stream));  // `seq` is the caller's again
    }
    g->loaded[(size_t)tid] = n;
    g->tables_dirty = true;
    return MTH_OK;
}

int mth_tag(mth_genome* g, const mth_tag_batch* b, mth_tag_result* out) {
    if (!g) return MTH_ERR_INVALID;
    if (!b || !out || b->n_reads < 0) { g->err = "mth_tag: bad arguments"; return MTH_ERR_INVALID; }
    const int64_t n = b->n_reads;
    memset(out, 0, sizeof(*out));
    if (n == 0) return MTH_OK;
    if (!b->tid || !b->pos || !b->rc || !b->l_seq || !b->cigar_off || !b->seq_off) { g->err = "mth_tag: null array"; return MTH_ERR_INVALID; }
    if (n > (int64_t)1 << 31) { g->err = "mth_tag: more than 2^31 reads in one batch"; return MTH_ERR_UNSUPPORTED; }
    G_CUDA(cudaSetDevice(g->device));
    const uint32_t n_cigar = b->cigar_off[n];
    const uint64_t n_seq = b->seq_off[n];
    if ((n_cigar && !b->cigar) || (n_seq && !b->seq4)) { g->err = "mth_tag: null cigar / seq4"; return MTH_ERR_INVALID; }
    // offsets of the column scratch (M + I + D) and of the tag buffer (M + I), from the CIGARs
    G_CUDA(g->h_col_off.ensure((size_t)n + 1));
    G_CUDA(g->h_xm_off.ensure((size_t)n + 1));
    uint64_t cols = 0, xms = 0;
    for (int64_t r = 0; r < n; r++) {
        g->h_col_off.p[r] = cols;
        g->h_xm_off.p[r] = xms;
        const uint32_t lo = b->cigar_off[r], hi = b->cigar_off[r + 1];
        if (hi < lo || hi > n_cigar || b->seq_off[r + 1] < b->seq_off[r] || b->l_seq[r] < 0 ||
            (uint64_t)(b->seq_off[r + 1] - b->seq_off[r]) < ((uint64_t)b->l_seq[r] + 1) / 2) {
            g->err = "mth_tag: inconsistent offsets at read " + std::to_string(r);
            return MTH_ERR_INVALID;
        }
        for (uint32_t k = lo; k < hi; k++) {
            const uint32_t v = b->cigar[k], op = v & 15u;
            if (op <= 2) cols += v >> 4;
            if (op <= 1) xms += v >> 4;
        }
        xms = (xms + 7) & ~(uint64_t)7;  // every tag starts 8-byte aligned and owns whole 8-byte words (k_tag stores 8 characters at a time)
    }
    g->h_col_off.p[n] = cols;
    g->h_xm_off.p[n] = xms;

    if (g->tables_dirty) {
        const size_t tb = sizeof(int64_t) * (size_t)g->n_ref;
        if (tb) {
            G_CUDA(cudaMemcpyAsync(g->d_contig_off, g->contig_off.data(), tb, cudaMemcpyHostToDevice, g->stream));
            G_CUDA(cudaMemcpyAsync(g->d_contig_len, g->ref_len.data(), tb, cudaMemcpyHostToDevice, g->stream));
            G_CUDA(cudaMemcpyAsync(g->d_loaded, g->loaded.data(), tb, cudaMemcpyHostToDevice, g->stream));
            G_CUDA(cudaStreamSynchronize(g->stream));
        }
        g->tables_dirty = false;
    }
    const size_t N = (size_t)n;
    G_CUDA(g->d_tid.ensure(N)); G_CUDA(g->d_pos.ensure(N)); G_CUDA(g->d_lseq.ensure(N)); G_CUDA(g->d_rc.ensure(N));
    G_CUDA(g->d_cigar_off.ensure(N + 1)); G_CUDA(g->d_cigar.ensure(n_cigar ? n_cigar : 1));
    G_CUDA(g->d_seq_off.ensure(N + 1)); G_CUDA(g->d_seq4.ensure(n_seq + 16));  // k_tag reads whole words: up to 7 bytes past a read's SEQ
    G_CUDA(g->d_col_off.ensure(N + 1)); G_CUDA(g->d_xm_off.ensure(N + 1));
    G_CUDA(g->d_colr.ensure(cols ? cols : 1)); G_CUDA(g->d_colf.ensure(cols ? cols : 1));
    G_CUDA(g->d_xm.ensure(xms ? xms : 1)); G_CUDA(g->d_xm_len.ensure(N)); G_CUDA(g->d_status.ensure(N));
    G_CUDA(g->h_xm.ensure(xms ? xms : 1)); G_CUDA(g->h_xm_len.ensure(N)); G_CUDA(g->h_status.ensure(N));
    cudaStream_t s = g->stream;
    const auto H2D = cudaMemcpyHostToDevice;
    G_CUDA(cudaMemcpyAsync(g->d_tid.p, b->tid, N * 4, H2D, s));
    G_CUDA(cudaMemcpyAsync(g->d_pos.p, b->pos, N * 4, H2D, s));
    G_CUDA(cudaMemcpyAsync(g->d_lseq.p, b->l_seq, N * 4, H2D, s));
    G_CUDA(cudaMemcpyAsync(g->d_rc.p, b->rc, N, H2D, s));
    G_CUDA(cudaMemcpyAsync(g->d_cigar_off.p, b->cigar_off, (N + 1) * 4, H2D, s));
    if (n_cigar) G_CUDA(cudaMemcpyAsync(g->d_cigar.p, b->cigar, (size_t)n_cigar * 4, H2D, s));
    G_CUDA(cudaMemcpyAsync(g->d_seq_off.p, b->seq_off, (N + 1) * 8, H2D, s));
    if (n_seq) G_CUDA(cudaMemcpyAsync(g->d_seq4.p, b->seq4, (size_t)n_seq, H2D, s));
    G_CUDA(cudaMemcpyAsync(g->d_col_off.p, g->h_col_off.p, (N + 1) * 8, H2D, s));
    G_CUDA(cudaMemcpyAsync(g->d_xm_off.p, g->h_xm_off.p, (N + 1) * 8, H2D, s));

    TagArgs a;
    a.n_reads = n;
    a.tid = g->d_tid.p; a.pos = g->d_pos.p; a.rc = g->d_rc.p; a.l_seq = g->d_lseq.p;
    a.cigar_off = g->d_cigar_off.p; a.cigar = g->d_cigar.p; a.seq_off = g->d_seq_off.p; a.seq4 = g->d_seq4.p;
    a.col_off = g->d_col_off.p; a.xm_off = g->d_xm_off.p;
    a.genome = g->d_genome; a.contig_off = g->d_contig_off; a.contig_len = g->d_contig_len; a.loaded_len = g->d_loaded;
    a.n_ref = g->n_ref;
    a.col_read = g->d_colr.p; a.col_ref = g->d_colf.p; a.xm = g->d_xm.p; a.xm_len = g->d_xm_len.p; a.status = g->d_status.p;
    const int64_t reads_per_block = (TAG_BLOCK / 32) * TAG_RPW;
    int64_t blocks = (n + reads_per_block - 1) / reads_per_block;
    const int64_t cap = (int64_t)g->sm_count * TAG_MINB;  // resident CTAs of 256 threads per SM (launch bounds)
    if (blocks > cap) blocks = cap;
    if (!g->ev0) { G_CUDA(cudaEventCreate(&g->ev0)); G_CUDA(cudaEventCreate(&g->ev1)); }
    G_CUDA(cudaEventRecord(g->ev0, s));
    k_tag<<<(unsigned)blocks, TAG_BLOCK, 0, s>>>(a);
    G_CUDA(cudaGetLastError());
    G_CUDA(cudaEventRecord(g->ev1, s));
    const auto D2H = cudaMemcpyDeviceToHost;
    if (xms) G_CUDA(cudaMemcpyAsync(g->h_xm.p, g->d_xm.p, (size_t)xms, D2H, s));
    G_CUDA(cudaMemcpyAsync(g->h_xm_len.p, g->d_xm_len.p, N * 4, D2H, s));
    G_CUDA(cudaMemcpyAsync(g->h_status.p, g->d_status.p, N, D2H, s));
    G_CUDA(cudaStreamSynchronize(s));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g->ev0, g->ev1) == cudaSuccess) g->last_kernel_ms = ms;
    int64_t failed = 0;
    for (int64_t r = 0; r < n; r++) failed += g->h_status.p[r] != MTH_TAG_OK;
    out->n_reads = n;
    out->n_failed = failed;
    out->xm_off = g->h_xm_off.p;
    out->xm_len = g->h_xm_len.p;
    out->xm = g->h_xm.p;
    out->status = g->h_status.p;
    return MTH_OK;
}

double mth_genome_last_kernel_ms(mth_genome* g) { return g ? g->last_kernel_ms : -1.0; }

int mth_genome_destroy(mth_genome* g) {
    if (!g) return MTH_OK;
    cudaSetDevice(g->device);
    if (g->ev0) cudaEventDestroy(g->ev0);
    if (g->ev1) cudaEventDestroy(g->ev1);
    if (g->stream) { cudaStreamSynchronize(g->stream); cudaStreamDestroy(g->stream); }
    if (g->d_genome) cudaFree(g->d_genome);
    if (g->d_contig_off) cudaFree(g->d_contig_off);
    if (g->d_contig_len) cudaFree(g->d_contig_len);
    if (g->d_loaded) cudaFree(g->d_loaded);
    g->d_tid.release(); g->d_pos.release(); g->d_lseq.release(); g->d_rc.release(); g->d_seq4.release();
    g->d_colr.release(); g->d_colf.release(); g->d_xm.release(); g->d_status.release(); g->d_cigar_off.release();
    g->d_cigar.release(); g->d_xm_len.release(); g->d_seq_off.release(); g->d_col_off.release(); g->d_xm_off.release();
    g->h_xm_off.release(); g->h_col_off.release(); g->h_xm_len.release(); g->h_xm.release(); g->h_status.release();
    delete g;
    return MTH_OK;
}

}  // extern "C"
