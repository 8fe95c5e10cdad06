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

// ---- TMA (cp.async.bulk) + mbarrier helpers: 1-D bulk copies global -> shared, completion on an mbarrier ----
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

// One CTA = one tile of ING_TILE consecutive reads.  The tile's CpG calls are one contiguous slice of cpg_pos /
// cpg_rel: thread 0 arms an mbarrier and issues two TMA bulk copies of the slice into shared memory while every
// thread loads its read's fixed-size fields (coalesced).  Then
//   phase 1 (one thread per CpG call, strided): mark the site bitmap — independent loads/atomics, no serial chain;
//   phase 2 (one thread per read, calls from shared memory): validation + LPMD pair loop at shared-memory latency.
// Tiles with more than ING_CAP calls (dense CpG islands) read the slice from global memory instead.
constexpr int ING_TILE = 256;
constexpr int ING_CAP = 3072;
#ifndef ING_MINB
#define ING_MINB 8
#endif

__global__ void __launch_bounds__(ING_TILE, ING_MINB) k_ingest(IngestArgs a) {
    __shared__ __align__(16) int32_t s_pos[ING_CAP + 8];
    __shared__ __align__(16) uint16_t s_rel[ING_CAP + 8];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_g0, s_lo, s_hi, s_staged;
    __shared__ uint32_t s_red[ING_TILE / 32][6];

    const ReadsView& rv = a.rv;
    const int tid = threadIdx.x;
    const int64_t tile0 = a.r0 + (int64_t)blockIdx.x * ING_TILE;
    const int64_t tile1 = min(a.r0 + a.n, tile0 + ING_TILE);

    if (tid == 0) {
        uint32_t lo = rv.cpg_off[tile0], hi = rv.cpg_off[tile1];
        if (hi < lo || (int64_t)hi > rv.I) hi = lo;  // reported as ERRBIT_BAD_OFFSETS by the per-read check below
        // 16-byte aligned bulk range [g0, g1) (8 calls = 32 B of cpg_pos = 16 B of cpg_rel), never past the arrays
        uint32_t g0 = lo & ~7u;
        uint32_t g1 = min((hi + 7u) & ~7u, (uint32_t)(rv.I & ~7ll));
        uint32_t staged = 0;
        if (hi > lo && hi - g0 <= (uint32_t)ING_CAP && g1 > g0) {
            staged = g1 - g0;
            mbar_init(&s_bar, 1);
            mbar_expect_tx(&s_bar, staged * (a.do_lpmd ? 6u : 4u));
            bulk_g2s(s_pos, rv.cpg_pos + g0, staged * 4u, &s_bar);
            if (a.do_lpmd) bulk_g2s(s_rel, a.cpg_rel + g0, staged * 2u, &s_bar);
        }
        s_g0 = g0; s_lo = lo; s_hi = hi; s_staged = staged;
    }

    // per-read fixed-size fields: coalesced, independent of the bulk copies in flight
    const int64_t j = tile0 + tid;
    const bool in = j < tile1;
    int32_t s = 0, e = 0;
    uint32_t o0 = 0, o1 = 0, meta = 0;
    if (in) {
        s = rv.start[j]; e = rv.end[j]; meta = rv.meta[j];
        o0 = rv.cpg_off[j]; o1 = rv.cpg_off[j + 1];
    }
    __syncthreads();
    const uint32_t g0 = s_g0, lo = s_lo, hi = s_hi, staged = s_staged;
    const bool fits = hi - g0 <= (uint32_t)ING_CAP;  // slice addressable in shared memory
    if (fits && hi > lo) {
        // the (at most 7) calls beyond the last 16-byte boundary of the arrays are fetched with plain loads
        for (uint32_t x = g0 + staged + tid; x < hi; x += ING_TILE) {
            s_pos[x - g0] = rv.cpg_pos[x];
            if (a.do_lpmd) s_rel[x - g0] = a.cpg_rel[x];
        }
        if (staged) mbar_wait(&s_bar, 0);
        __syncthreads();
    }
    // accessors indexed by the GLOBAL call index
    auto pos_at = [&](uint32_t x) -> int32_t { return fits ? s_pos[x - g0] : rv.cpg_pos[x]; };
    auto rel_at = [&](uint32_t x) -> int32_t { return fits ? (int32_t)s_rel[x - g0] : (int32_t)a.cpg_rel[x]; };

    // ---- phase 1: site bitmap, one thread per CpG call ----
    for (uint32_t x = lo + tid; x < hi; x += ING_TILE) {
        int32_t p = pos_at(x);
        if (p < a.lin_lo - 1 || p >= a.lin_hi) continue;  // flagged per read below; never touch memory outside the contig
        uint32_t bit = (uint32_t)(p + 1);
        unsigned long long* w = a.bitmap + (bit >> 6);
        unsigned long long m = 1ull << (bit & 63);
        if (!(__ldg((const unsigned long long*)w) & m)) atomicOr(w, m);
    }

    // ---- phase 2: per read ----
    uint32_t err = 0;
    int32_t span = 0;
    uint32_t lp_read = 0, lp_valid = 0, lp_c = 0, lp_d = 0;
    if (in) {
        if (j > 0 && s < rv.start[j - 1]) err |= ERRBIT_UNSORTED;
        span = e - s + 1;
        if (span < 1 || span > MAX_REF_SPAN) err |= ERRBIT_SPAN;
        if (s < a.lin_lo || e >= a.lin_hi) err |= ERRBIT_POS_RANGE;
        if (o1 < o0 || o0 < lo || o1 > hi) { err |= ERRBIT_BAD_OFFSETS; o1 = o0; }
        uint32_t n = o1 - o0;
        uint32_t cap = rv.meth_off ? (uint32_t)MAX_CPGS_PER_READ : 64u;
        if (n > cap) { err |= ERRBIT_TOO_MANY_CPGS; n = 0; }
        if (rv.meth_off && n > 0) {
            uint32_t m0 = rv.meth_off[j], m1 = rv.meth_off[j + 1];
            if (m1 < m0 || (m1 - m0) * 64u < n) { err |= ERRBIT_BAD_OFFSETS; n = 0; }
        }
        int32_t prev = s - 2;
        for (uint32_t k = 0; k < n; k++) {
            int32_t p = pos_at(o0 + k);
            if (p <= prev) err |= ERRBIT_CPG_ORDER;
            if (p < s - 1 || p > e) err |= ERRBIT_POS_RANGE;
            prev = p;
        }
        if (a.do_lpmd && !(meta & META_HALO)) lp_read = 1;  // lpmd.rs:176 (a halo copy is counted by its owner rank)
        if (lp_read && !(err & (ERRBIT_BAD_OFFSETS | ERRBIT_TOO_MANY_CPGS))) {
            if ((meta & 0xFFu) >= a.lpmd.min_qual) {   // lpmd.rs:177
                lp_valid = 1;
                uint64_t w0 = n > 1 ? meth_word(rv, j, 0) : 0;
                for (uint32_t k = 1; k < n; k++) {
                    int32_t rk = rel_at(o0 + k);
                    uint32_t mk = k < 64 ? (uint32_t)((w0 >> k) & 1ull) : meth_bit(rv, j, k);
                    for (int32_t q = (int32_t)k - 1; q >= 0; q--) {
                        int32_t d = rk - rel_at(o0 + (uint32_t)q);
                        if (d > a.lpmd.max_distance) break;      // readutil.rs:184 (anchors popped from the front)
                        if (d < a.lpmd.min_distance) continue;   // readutil.rs:196
                        uint32_t mq = q < 64 ? (uint32_t)((w0 >> q) & 1ull) : meth_bit(rv, j, (uint32_t)q);
                        if (mq == mk) lp_c++; else lp_d++;       // readutil.rs:200-214
                    }
                }
            }
        }
    }
    // ---- block reduction, then a handful of atomics per CTA ----
    const int lane = lane_id(), warp = tid >> 5;
    int32_t wmax = __reduce_max_sync(FULL, span);
    uint32_t werr = __reduce_or_sync(FULL, err);
    uint32_t wv = __reduce_add_sync(FULL, lp_valid), wc = __reduce_add_sync(FULL, lp_c), wd = __reduce_add_sync(FULL, lp_d);
    uint32_t wr = __reduce_add_sync(FULL, lp_read);
    if (lane == 0) {
        s_red[warp][0] = (uint32_t)wmax; s_red[warp][1] = werr; s_red[warp][2] = wv; s_red[warp][3] = wc; s_red[warp][4] = wd; s_red[warp][5] = wr;
    }
    __syncthreads();
    if (tid == 0) {
        int32_t bmax = 0;
        uint32_t berr = 0, bv = 0, bc = 0, bd = 0, br = 0;
#pragma unroll
        for (int w = 0; w < ING_TILE / 32; w++) {
            bmax = max(bmax, (int32_t)s_red[w][0]); berr |= s_red[w][1]; bv += s_red[w][2]; bc += s_red[w][3]; bd += s_red[w][4]; br += s_red[w][5];
        }
        if (bmax > 0) atomicMax(&a.sc->lmax, bmax);
        if (berr) atomicOr(&a.sc->err, berr);
        if (a.do_lpmd) {
            if (br) atomicAdd(&a.sc->lpmd[0], (unsigned long long)br);
            if (bv) atomicAdd(&a.sc->lpmd[1], (unsigned long long)bv);
            if (bc) atomicAdd(&a.sc->lpmd[2], (unsigned long long)bc);
            if (bd) atomicAdd(&a.sc->lpmd[3], (unsigned long long)bd);
        }
    }
}

int launch_ingest(const IngestArgs& a, cudaStream_t s) {
    if (a.n <= 0) return 0;
    k_ingest<<<grid_for(a.n, ING_TILE), ING_TILE, 0, s>>>(a);
    return 1;
}

}  // namespace mth
