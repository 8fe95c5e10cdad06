// k_cpgset.cu — `--cpg-set` on the device (readutil.rs:87-95 filter_isin, :347-374 get_target_cpgs).
//
// The reference reads the BED file into a HashSet<(tid, pos)> and drops every CpG call that is not in it BEFORE anything else
// is computed from the read, so quartets, stretches, pairs and "first CpG" are formed over the retained calls.  Here the set is
// a bitmap over the region's linear coordinate (bit p + 1 like the site bitmap, 1 bit per base), filled from the sorted
// (tid, pos) list the host handed over (mth_set_cpg_set); a freshly copied / expanded batch is filtered in the arena before
// k_ingest sees it:
//   k_cpgset_count   thread per read: test its calls against the bitmap, squeeze the methylation bits of the kept calls
//                    together (in place: a read owns its words), write the kept count
//   (exclusive scan of the counts, k_sites.cu)
//   k_cpgset_scatter thread per read: kept positions / query indices -> a scratch copy at the new offsets, new cpg_off
// and the host copies the scratch back over the batch's slice of the call arrays (it needs the kept total to know where the
// next batch goes: one stream synchronisation per batch, only when a CpG set is active).
#include "kernels.h"

namespace mth {

__global__ void k_cpgset_mark(const int32_t* __restrict__ pos, int64_t n, int32_t lin_off, int32_t len, unsigned long long* __restrict__ bitmap) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t p = pos[i];
    if (p < -1 || p >= len) return;  // outside the contig: can never match a call
    const uint32_t bit = (uint32_t)(lin_off + p + 1);
    atomicOr(&bitmap[bit >> 6], 1ull << (bit & 63));
}

// positions are validated by k_ingest AFTER the filter: a corrupt one must not index outside the bitmap here
__device__ __forceinline__ bool in_set(const unsigned long long* __restrict__ bm, int64_t n_words, int32_t p) {
    const uint32_t bit = (uint32_t)(p + 1);
    if (p < -1 || (int64_t)(bit >> 6) >= n_words) return false;
    return (bm[bit >> 6] >> (bit & 63)) & 1ull;
}

struct CpgsetArgs {
    int64_t r0, n, i0;                 // reads [r0, r0 + n) of the arena, their calls start at i0
    const uint32_t* off;               // arena cpg_off (old offsets, absolute)
    const int32_t* pos;
    const uint16_t* rel;               // nullptr unless LPMD
    uint64_t* meth;
    const uint32_t* meth_off;          // nullptr: one word per read
    const unsigned long long* set_bitmap;
    int64_t set_words;
    uint32_t* kept;                    // [n] out: kept calls per read (scanned in place afterwards)
    int32_t* pos_tmp;
    uint16_t* rel_tmp;
    uint32_t* off_out;                 // arena cpg_off (written by the scatter pass)
    const unsigned long long* total;   // device: kept calls of the batch (after the scan)
};

__global__ void __launch_bounds__(256) k_cpgset_count(CpgsetArgs a) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n) return;
    const int64_t j = a.r0 + r;
    const uint32_t o0 = a.off[j], n = a.off[j + 1] - o0;
    const size_t w0 = a.meth_off ? a.meth_off[j] : (size_t)j;
    if (n == 0) { a.kept[r] = 0; return; }  // (a read without calls may own no methylation word at all)
    uint32_t k_out = 0;
    uint64_t cur_in = 0, cur_out = 0;
    for (uint32_t k = 0; k < n && k < (uint32_t)MAX_CPGS_PER_READ; k++) {
        if ((k & 63u) == 0) cur_in = a.meth[w0 + (k >> 6)];
        if (in_set(a.set_bitmap, a.set_words, a.pos[o0 + k])) {
            cur_out |= ((cur_in >> (k & 63u)) & 1ull) << (k_out & 63u);
            k_out++;
            if ((k_out & 63u) == 0) {  // a full output word: its input word has been consumed entirely (k_out <= k + 1)
                a.meth[w0 + (k_out >> 6) - 1] = cur_out;
                cur_out = 0;
            }
        }
    }
    if ((k_out & 63u) != 0 || k_out == 0) a.meth[w0 + (k_out >> 6)] = cur_out;
    a.kept[r] = k_out;
}

__global__ void __launch_bounds__(256) k_cpgset_scatter(CpgsetArgs a) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n) return;
    const int64_t j = a.r0 + r;
    const uint32_t o0 = a.off[j], n = a.off[j + 1] - o0;
    uint32_t at = a.kept[r];  // exclusive scan: first kept call of this read within the batch
    // the old offsets of read j + 1 are still needed by the thread of read j + 1: new offsets go to a second array
    for (uint32_t k = 0; k < n && k < (uint32_t)MAX_CPGS_PER_READ; k++) {
        const int32_t p = a.pos[o0 + k];
        if (in_set(a.set_bitmap, a.set_words, p)) {
            a.pos_tmp[at] = p;
            if (a.rel) a.rel_tmp[at] = a.rel[o0 + k];
            at++;
        }
    }
}

// cpg_off[r0 + r] = i0 + exclusive[r]  (separate pass: the scatter pass still reads the old offsets)
__global__ void __launch_bounds__(256) k_cpgset_offsets(CpgsetArgs a) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > a.n) return;
    a.off_out[a.r0 + r] = (uint32_t)(a.i0 + (r < a.n ? a.kept[r] : (uint32_t)*a.total));
}

int launch_cpgset_mark(const int32_t* pos, int64_t n, int32_t lin_off, int32_t len, unsigned long long* bitmap, cudaStream_t s) {
    if (n <= 0) return 0;
    k_cpgset_mark<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(pos, n, lin_off, len, bitmap);
    return 1;
}

int launch_cpgset_filter(int64_t r0, int64_t n, int64_t i0, uint32_t* off, const int32_t* pos, const uint16_t* rel, uint64_t* meth,
                         const uint32_t* meth_off, const unsigned long long* set_bitmap, int64_t set_words, uint32_t* kept, uint32_t* scan_scratch,
                         unsigned long long* total, int32_t* pos_tmp, uint16_t* rel_tmp, cudaStream_t s) {
    if (n <= 0) return 0;
    CpgsetArgs a{r0, n, i0, off, pos, rel, meth, meth_off, set_bitmap, set_words, kept, pos_tmp, rel_tmp, off, total};
    const unsigned g = (unsigned)((n + 255) / 256), g1 = (unsigned)((n + 1 + 255) / 256);
    int k = 0;
    k_cpgset_count<<<g, 256, 0, s>>>(a);
    k++;
    k += launch_exclusive_scan_u32(kept, n, scan_scratch, total, s);
    k_cpgset_scatter<<<g, 256, 0, s>>>(a);
    k_cpgset_offsets<<<g1, 256, 0, s>>>(a);
    return k + 2;
}

}  // namespace mth
