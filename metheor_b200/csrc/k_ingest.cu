// k_ingest.cu — per-batch kernels: offset/coordinate fix-ups, validation, CpG-site marking and LPMD.
//
// k_ingest is one pass over a freshly copied batch (16 B/read + 4 B/CpG [+2 B/CpG for LPMD] of HBM reads):
//   * validates what every later kernel relies on (sorted starts, monotone offsets, CpG positions strictly
//     increasing and inside [start-1, end]) and records max(end-start+1) for the gather windows;
//   * marks each CpG position in the region's site bitmap (load-test then atomicOr: bits are only ever set,
//     so a stale 0 only costs a redundant atomic);
//   * LPMD (lpmd.rs:175-199 + readutil.rs:166-224): all in-read CpG pairs with min <= d(query index) <= max,
//     concordant iff equal methylation; warp-reduced into four 64-bit device counters.
#include "kernels.h"

namespace mth {

__global__ void k_add_u32(uint32_t* a, int64_t n, uint32_t add) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += add;
}
__global__ void k_iota_u32(uint32_t* a, int64_t n, uint32_t base) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = base + (uint32_t)i;
}
__global__ void k_add_i32(int32_t* a, int64_t n, int32_t add) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += add;
}

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

int launch_add_u32(uint32_t* a, int64_t n, uint32_t add, cudaStream_t s) {
    if (n <= 0) return 0;
    k_add_u32<<<grid_for(n, 256), 256, 0, s>>>(a, n, add);
    return 1;
}
int launch_iota_u32(uint32_t* a, int64_t n, uint32_t base, cudaStream_t s) {
    if (n <= 0) return 0;
    k_iota_u32<<<grid_for(n, 256), 256, 0, s>>>(a, n, base);
    return 1;
}
int launch_add_i32(int32_t* a, int64_t n, int32_t add, cudaStream_t s) {
    if (n <= 0) return 0;
    k_add_i32<<<grid_for(n, 256), 256, 0, s>>>(a, n, add);
    return 1;
}

__global__ void __launch_bounds__(256) k_ingest(IngestArgs a) {
    const ReadsView& rv = a.rv;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool in = t < a.n;
    int64_t j = a.r0 + (in ? t : 0);
    uint32_t err = 0;
    int32_t span = 0;
    uint32_t lp_valid = 0, lp_c = 0, lp_d = 0;
    if (in) {
        int32_t s = rv.start[j], e = rv.end[j];
        if (j > 0 && s < rv.start[j - 1]) err |= ERRBIT_UNSORTED;
        span = e - s + 1;
        if (span < 1 || span > MAX_REF_SPAN) err |= ERRBIT_SPAN;
        if (s < a.lin_lo || e >= a.lin_hi) err |= ERRBIT_POS_RANGE;
        uint32_t o0 = rv.cpg_off[j], o1 = rv.cpg_off[j + 1];
        if (o1 < o0 || (int64_t)o1 > rv.I) { err |= ERRBIT_BAD_OFFSETS; o1 = o0; }
        uint32_t n = o1 - o0;
        uint32_t cap = rv.meth_off ? (uint32_t)MAX_CPGS_PER_READ : 64u;
        if (n > cap) { err |= ERRBIT_TOO_MANY_CPGS; n = 0; }
        if (rv.meth_off && n > 0) {
            uint32_t m0 = rv.meth_off[j], m1 = rv.meth_off[j + 1];
            if (m1 < m0 || (m1 - m0) * 64u < n) { err |= ERRBIT_BAD_OFFSETS; n = 0; }
        }
        int32_t prev = s - 2;
        for (uint32_t k = 0; k < n; k++) {
            int32_t p = rv.cpg_pos[o0 + k];
            if (p <= prev) err |= ERRBIT_CPG_ORDER;
            if (p < s - 1 || p > e) { err |= ERRBIT_POS_RANGE; prev = p; continue; }
            prev = p;
            uint32_t bit = (uint32_t)(p + 1);
            unsigned long long* w = a.bitmap + (bit >> 6);
            unsigned long long m = 1ull << (bit & 63);
            if (!(*w & m)) atomicOr(w, m);
        }
        if (a.do_lpmd && !(err & (ERRBIT_BAD_OFFSETS | ERRBIT_TOO_MANY_CPGS))) {
            uint32_t mapq = rv.meta[j] & 0xFFu;
            if (mapq >= a.lpmd.min_qual) {   // lpmd.rs:177
                lp_valid = 1;
                const uint16_t* rel = a.cpg_rel + ((int64_t)o0 - a.i0);
                uint64_t w0 = n ? meth_word(rv, j, 0) : 0;
                for (uint32_t k = 1; k < n; k++) {
                    int32_t rk = rel[k];
                    uint32_t mk = k < 64 ? (uint32_t)((w0 >> k) & 1ull) : meth_bit(rv, j, k);
                    for (int32_t q = (int32_t)k - 1; q >= 0; q--) {
                        int32_t d = rk - (int32_t)rel[q];
                        if (d > a.lpmd.max_distance) break;      // readutil.rs:184 (anchors popped from the front)
                        if (d < a.lpmd.min_distance) continue;   // readutil.rs:196
                        uint32_t mq = q < 64 ? (uint32_t)((w0 >> q) & 1ull) : meth_bit(rv, j, (uint32_t)q);
                        if (mq == mk) lp_c++; else lp_d++;       // readutil.rs:200-214
                    }
                }
            }
        }
    }
    // warp-level reductions, then a handful of atomics per warp
    int32_t wmax = __reduce_max_sync(FULL, span);
    uint32_t werr = __reduce_or_sync(FULL, err);
    if (lane_id() == 0) {
        atomicMax(&a.sc->lmax, wmax);
        if (werr) atomicOr(&a.sc->err, werr);
    }
    if (a.do_lpmd) {
        uint32_t nread = __popc(__ballot_sync(FULL, in));
        uint32_t nvalid = __reduce_add_sync(FULL, lp_valid);
        uint32_t c = __reduce_add_sync(FULL, lp_c);
        uint32_t d = __reduce_add_sync(FULL, lp_d);
        if (lane_id() == 0) {
            atomicAdd(&a.sc->lpmd[0], (unsigned long long)nread);
            if (nvalid) atomicAdd(&a.sc->lpmd[1], (unsigned long long)nvalid);
            if (c) atomicAdd(&a.sc->lpmd[2], (unsigned long long)c);
            if (d) atomicAdd(&a.sc->lpmd[3], (unsigned long long)d);
        }
    }
}

int launch_ingest(const IngestArgs& a, cudaStream_t s) {
    if (a.n <= 0) return 0;
    k_ingest<<<grid_for(a.n, 256), 256, 0, s>>>(a);
    return 1;
}

}  // namespace mth
