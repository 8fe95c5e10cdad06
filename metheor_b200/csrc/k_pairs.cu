// k_pairs.cu — the optional `lpmd --pairs` table (lpmd.rs:70-87 accumulate, :89-122 print): for every CpG pair
// (a, b) seen in one read with min_distance <= query distance <= max_distance, the number of reads in which the two
// calls agree / disagree, and lpmd = n_d as f32 / (n_c as f32 + n_d as f32) (lpmd.rs:111).
//
// LPMD never flushes, so there are no segments: the warp that owns site a visits every read calling a (they all start
// inside a's window, gather.cuh) and enumerates a's partner sites in ascending order — one pass to find the next
// partner, one pass to count it — which yields the rows already sorted by (tid, pos1, pos2) (lpmd.rs:94 sorts the keys).
// Two kernels from one template: COUNT (rows per anchor site) and EMIT.
#include "gather.cuh"
#include "kernels.h"

namespace mth {

struct PairArgs {
    ReadsView rv;
    const uint16_t* rel;  // region-wide query indices, parallel to rv.cpg_pos
    const int32_t* site_pos;
    int64_t C;
    const RegionScalars* sc;
    mth_lpmd_params prm;
    uint32_t* rowcnt;        // COUNT: out
    const uint32_t* rowoff;  // EMIT: in
    ContigTable ct;
    PairRowsDev rows;
    int64_t row_base;
};

template <bool EMIT>
__global__ void __launch_bounds__(GATHER_BLOCK) k_lpmd_pairs(PairArgs a) {
    const ReadsView& rv = a.rv;
    const int lane = lane_id();
    const int32_t dmin = a.prm.min_distance, dmax = a.prm.max_distance;
    const uint32_t min_qual = a.prm.min_qual;
    for_each_site(rv, a.site_pos, a.C, a.sc->lmax, [&](int64_t s, int32_t p, int64_t lo, int32_t target) {
        uint32_t n_rows = 0;
        const int64_t out0 = EMIT ? (a.row_base + a.rowoff[s]) : 0;
        int32_t last = p;
        for (;;) {
            // ---- next partner position > last over all reads calling p ----
            int32_t best = INT32_MAX;
            scan_window(rv, lo, p, target, [&](const LaneRead& lr) {
                int32_t cand = INT32_MAX;
                if (lr.idx >= 0 && lr.mapq >= min_qual) {  // lpmd.rs:177
                    const int32_t* cp = rv.cpg_pos + lr.o0;
                    const uint16_t* rl = a.rel + lr.o0;
                    const int32_t r0 = rl[lr.idx];
                    for (uint32_t k = (uint32_t)lr.idx + 1; k < lr.n; k++) {
                        int32_t d = (int32_t)rl[k] - r0;
                        if (d > dmax) break;       // readutil.rs:184
                        if (d < dmin) continue;    // readutil.rs:196
                        if (cp[k] > last) { cand = cp[k]; break; }
                    }
                }
                best = min(best, __reduce_min_sync(FULL, cand));
            });
            if (best == INT32_MAX) break;
            // ---- count the pair (p, best) ----
            uint32_t nc = 0, nd = 0;
            scan_window(rv, lo, p, target, [&](const LaneRead& lr) {
                bool hit = false, conc = false;
                if (lr.idx >= 0 && lr.mapq >= min_qual && (uint32_t)lr.idx + 1 < lr.n) {
                    const int32_t* cp = rv.cpg_pos + lr.o0;
                    int k = find_pos(cp + lr.idx + 1, lr.n - (uint32_t)lr.idx - 1, best);
                    if (k >= 0) {
                        k += lr.idx + 1;
                        const uint16_t* rl = a.rel + lr.o0;
                        int32_t d = (int32_t)rl[k] - (int32_t)rl[lr.idx];
                        if (d >= dmin && d <= dmax) {
                            hit = true;
                            conc = meth_bit(rv, lr.j, (uint32_t)lr.idx) == meth_bit(rv, lr.j, (uint32_t)k);  // readutil.rs:200
                        }
                    }
                }
                uint32_t hm = __ballot_sync(FULL, hit), cm = __ballot_sync(FULL, hit && conc);
                nc += __popc(cm);
                nd += __popc(hm) - __popc(cm);
            });
            if (EMIT && lane == 0) {
                int64_t r = out0 + n_rows;
                int32_t tid, pos;
                delinearize(a.ct, p, &tid, &pos);
                a.rows.tid[r] = tid;
                a.rows.pos1[r] = pos;
                a.rows.pos2[r] = best - (p - pos);
                a.rows.lpmd[r] = __fdiv_rn((float)nd, __fadd_rn((float)nc, (float)nd));  // lpmd.rs:111
                a.rows.n_conc[r] = (int32_t)nc;
                a.rows.n_disc[r] = (int32_t)nd;
            }
            n_rows++;
            last = best;
        }
        if (!EMIT && lane == 0) a.rowcnt[s] = n_rows;
    });
}

int launch_lpmd_pairs_count(const ReadsView& rv, const uint16_t* rel, const int32_t* site_pos, int64_t C,
                            const RegionScalars* sc, mth_lpmd_params prm, uint32_t* rowcnt, cudaStream_t s) {
    if (C <= 0) return 0;
    PairArgs a;
    memset(&a, 0, sizeof(a));
    a.rv = rv; a.rel = rel; a.site_pos = site_pos; a.C = C; a.sc = sc; a.prm = prm; a.rowcnt = rowcnt;
    k_lpmd_pairs<false><<<gather_grid(C), GATHER_BLOCK, 0, s>>>(a);
    return 1;
}

int launch_lpmd_pairs_emit(const ReadsView& rv, const uint16_t* rel, const int32_t* site_pos, int64_t C,
                           const RegionScalars* sc, mth_lpmd_params prm, const uint32_t* rowoff, ContigTable ct,
                           PairRowsDev rows, int64_t row_base, cudaStream_t s) {
    if (C <= 0) return 0;
    PairArgs a;
    memset(&a, 0, sizeof(a));
    a.rv = rv; a.rel = rel; a.site_pos = site_pos; a.C = C; a.sc = sc; a.prm = prm; a.rowoff = rowoff; a.ct = ct;
    a.rows = rows; a.row_base = row_base;
    k_lpmd_pairs<true><<<gather_grid(C), GATHER_BLOCK, 0, s>>>(a);
    return 1;
}

}  // namespace mth
