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
#include "tma.cuh"

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

// One CTA = one tile of ING_TILE = 256 x ING_RPT consecutive reads (ING_RPT reads per thread, interleaved so that every
// global load is coalesced).  The tile's CpG calls are one contiguous slice of cpg_pos / cpg_rel:
//   prologue : thread 0 reads the two slice bounds, arms an mbarrier and issues the TMA bulk copies of the slice into
//              shared memory (tma.cuh); meanwhile every thread loads the fixed-size fields of its reads — offsets and
//              starts go to shared memory, the rest stays in registers;
//   per read : validation (sortedness, offsets, span -> lmax, first/last call inside [start-1, end]) and the read-level
//              decisions every later per-call step needs, written as one FLAG BYTE per call into shared memory:
//              methylated, first call of its read, read counts for LPMD (lpmd.rs:177), PDR state of the read
//              (filters pdr.rs:147-155 + concordance readutil.rs:134-145);
//   per call : order check against the previous call, OR of the site into a 16 384-position window of the site bitmap
//              held in shared memory, the LPMD pairs this call closes (walk back over the read's earlier calls while
//              the query distance is <= max_distance, readutil.rs:166-224), and the flag byte goes to global memory
//              (call_flags) so that k_pdr_scatter is a pure per-call kernel;
//   epilogue : non-zero window words merge into the global bitmap with fire-and-forget atomics; block reduction of the
//              scalars, <= 6 global atomics per CTA.
// After the prologue no thread waits on global memory again.  Tiles with more than ING_CAP calls (dense CpG islands)
// are processed in several passes over runs of reads whose calls fit the slice.
#ifndef ING_RPT
#define ING_RPT 2
#endif
#ifndef ING_MINB
#define ING_MINB 6
#endif
constexpr int ING_THREADS = 256;
constexpr int ING_TILE = ING_THREADS * ING_RPT;
constexpr int ING_CAP = 1024 * ING_RPT;   // calls staged per pass (avg 2.7 per read on WGBS); denser tiles take several passes
constexpr int ING_WIN_WORDS = 512;  // 32-bit words of the shared bitmap window = 16 384 positions

struct ReadVerdict {  // what the per-read step decides
    uint32_t err, n, flags;
    bool bad;
};

// Validation of one read + the read-level part of its calls' flag byte.  first / last = its first / last call position.
__device__ __forceinline__ ReadVerdict judge_read(const IngestArgs& a, int64_t j, int32_t s, int32_t e, int32_t prev_start,
                                                  uint32_t meta, uint32_t o0, uint32_t o1, uint32_t lo, uint32_t hi) {
    const ReadsView& rv = a.rv;
    ReadVerdict v;
    v.err = 0; v.bad = false; v.flags = 0;
    if (s < prev_start) v.err |= ERRBIT_UNSORTED;
    int32_t span = e - s + 1;
    if (span < 1 || span > MAX_REF_SPAN) v.err |= ERRBIT_SPAN;
    if (s < a.lin_lo || e >= a.lin_hi) v.err |= ERRBIT_POS_RANGE;
    if (o1 < o0 || o0 < lo || o1 > hi) { v.err |= ERRBIT_BAD_OFFSETS; o1 = o0; v.bad = true; }
    v.n = o1 - o0;
    uint32_t cap = rv.meth_off ? (uint32_t)MAX_CPGS_PER_READ : 64u;
    if (v.n > cap) { v.err |= ERRBIT_TOO_MANY_CPGS; v.n = 0; v.bad = true; }
    if (rv.meth_off && v.n > 0) {
        uint32_t m0 = rv.meth_off[j], m1 = rv.meth_off[j + 1];
        if (m1 < m0 || (m1 - m0) * 64u < v.n) { v.err |= ERRBIT_BAD_OFFSETS; v.n = 0; v.bad = true; }
    }
    const uint32_t mapq = meta & 0xFFu;
    if (a.do_lpmd && !(meta & META_HALO) && !v.bad && mapq >= a.lpmd.min_qual) v.flags |= CF_LPMD;  // lpmd.rs:177
    if (a.do_pm && mapq >= a.pm_min_qual) v.flags |= CF_PM_OK;                                       // pm.rs:111
    if (a.do_me && mapq >= a.me_min_qual) v.flags |= CF_ME_OK;                                       // me.rs:115
    return v;
}

// What thread 0 plans for one pass over (part of) a tile: reads [ra, rb) of the tile and their calls [lo, hi).
struct TilePlan {
    int ra, rb;            // tile-relative read range; rb == 0: not planned yet
    uint32_t g0, lo, hi;   // staged slice starts at call g0 (16-byte aligned), calls [lo, hi) belong to the reads
    uint32_t staged;       // calls covered by the TMA bulk copies (0: none issued)
    uint32_t parity;       // mbarrier phase to wait for
};

// Called by ONE thread: arms the mbarrier and issues the bulk copies of the calls [lo, hi) (see tma.cuh for the
// alignment rules).  `uses` counts the copies issued on this barrier so far (phase parity).
template <int CAP>
__device__ __forceinline__ TilePlan plan_pass(int ra, int rb, uint32_t lo, uint32_t hi, int64_t n_calls_total, const int32_t* pos_src,
                                              const uint16_t* rel_src, int32_t* s_pos, uint16_t* s_rel, uint64_t* bar, uint32_t* uses) {
    TilePlan pl;
    pl.ra = ra; pl.rb = rb;
    if (hi < lo || (int64_t)hi > n_calls_total) hi = lo;  // reported as ERRBIT_BAD_OFFSETS by the per-read check
    pl.g0 = lo & ~7u;
    if (hi - pl.g0 > (uint32_t)CAP) hi = pl.g0 + CAP;     // only reachable with corrupt offsets (a read has <= 256 calls)
    pl.lo = lo; pl.hi = hi;
    pl.staged = 0;
    pl.parity = *uses & 1u;
    const uint32_t g1 = min((hi + 7u) & ~7u, (uint32_t)(n_calls_total & ~7ll));
    if (hi > lo && g1 > pl.g0) {
        pl.staged = g1 - pl.g0;
        mbar_expect_tx(bar, pl.staged * (rel_src ? 6u : 4u));
        bulk_g2s(s_pos, pos_src + pl.g0, pl.staged * 4u, bar);
        if (rel_src) bulk_g2s(s_rel, rel_src + pl.g0, pl.staged * 2u, bar);
        *uses += 1;
    }
    return pl;
}

__global__ void __launch_bounds__(ING_THREADS, ING_MINB) k_ingest(IngestArgs a) {
    __shared__ __align__(16) int32_t s_pos[ING_CAP + 8];
    __shared__ __align__(16) uint16_t s_rel[ING_CAP + 8];
    __shared__ uint8_t s_flags[ING_CAP + 8];
    __shared__ uint32_t s_off[ING_TILE + 1];
    __shared__ int32_t s_start[ING_TILE + 1];  // s_start[0] = start of the read before the tile
    __shared__ uint32_t s_bm[ING_WIN_WORDS];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ TilePlan s_plan;
    __shared__ uint32_t s_red[ING_THREADS / 32][6];

    const ReadsView& rv = a.rv;
    const int tid = threadIdx.x;
    const int64_t tile0 = a.r0 + (int64_t)blockIdx.x * ING_TILE;
    const int64_t tile1 = min(a.r0 + a.n, tile0 + ING_TILE);
    const int nr = (int)(tile1 - tile0);
    const uint16_t* rel_src = a.do_lpmd ? a.cpg_rel : nullptr;
    uint32_t tma_uses = 0;  // thread 0 only

    if (tid == 0) {
        mbar_init(&s_bar, 1);
        const uint32_t lo = rv.cpg_off[tile0], hi = rv.cpg_off[tile1];
        s_off[nr] = hi;
        s_start[0] = tile0 > 0 ? rv.start[tile0 - 1] : INT32_MIN;
        if (hi < lo || hi - (lo & ~7u) <= (uint32_t)ING_CAP)  // the common case: the whole tile in one pass, copies start now
            s_plan = plan_pass<ING_CAP>(0, nr, lo, hi, rv.I, rv.cpg_pos, rel_src, s_pos, s_rel, &s_bar, &tma_uses);
        else
            s_plan.rb = 0;  // dense tile (CpG island): passes are planned from the offsets once they are in shared memory
    }
    for (int k = tid; k < ING_WIN_WORDS; k += ING_THREADS) s_bm[k] = 0;

    // fixed-size fields of this thread's reads: coalesced, all issued before anything waits
    int32_t rs[ING_RPT], re[ING_RPT];
    uint32_t rmeta[ING_RPT];
    uint64_t rmw[ING_RPT];
#pragma unroll
    for (int i = 0; i < ING_RPT; i++) {
        const int idx = tid + i * ING_THREADS;
        rs[i] = 0; re[i] = 0; rmeta[i] = 0; rmw[i] = 0;
        if (idx < nr) {
            const int64_t j = tile0 + idx;
            rs[i] = rv.start[j]; re[i] = rv.end[j]; rmeta[i] = rv.meta[j];
            s_off[idx] = rv.cpg_off[j];
            if (!rv.meth_off) rmw[i] = rv.meth[j];
        }
    }
#pragma unroll
    for (int i = 0; i < ING_RPT; i++) {
        const int idx = tid + i * ING_THREADS;
        if (idx < nr) s_start[idx + 1] = rs[i];
    }
    __syncthreads();

    // thread 0: next pass = the longest run of reads from `ra` whose calls fit the shared-memory slice
    auto plan_next = [&](int ra) {
        const uint32_t lo = s_off[ra], g0 = lo & ~7u;
        int a_ = ra + 1, b_ = nr;  // largest rb in [ra+1, nr] with s_off[rb] - g0 <= CAP (a read never exceeds CAP)
        while (a_ < b_) {
            int m = (a_ + b_ + 1) >> 1;
            if (s_off[m] >= lo && s_off[m] - g0 <= (uint32_t)ING_CAP) a_ = m; else b_ = m - 1;
        }
        s_plan = plan_pass<ING_CAP>(ra, a_, lo, s_off[a_], rv.I, rv.cpg_pos, rel_src, s_pos, s_rel, &s_bar, &tma_uses);
    };

    uint32_t err = 0;
    int32_t span_max = 0;
    uint32_t lp0 = 0, lp1 = 0, lp2 = 0, lp3 = 0;  // n_read, n_valid_read, n_conc, n_disc
    uint32_t* bitmap32 = (uint32_t*)a.bitmap;             // little-endian view of the 64-bit words
    const uint32_t wbase = ((uint32_t)s_start[1]) >> 5;   // calls lie at or after start[tile0] - 1, i.e. bit >= start[tile0]
    const int32_t dmin = a.lpmd.min_distance, dmax = a.lpmd.max_distance;

    if (s_plan.rb == 0) {  // uniform: first pass of a dense tile
        __syncthreads();
        if (tid == 0) plan_next(0);
        __syncthreads();
    }
    for (;;) {
        const TilePlan pl = s_plan;
        const uint32_t g0 = pl.g0, lo = pl.lo, hi = pl.hi;
        for (uint32_t x = g0 + pl.staged + tid; x < hi; x += ING_THREADS) {  // tail beyond the last 16-byte boundary of the arrays
            s_pos[x - g0] = rv.cpg_pos[x];
            if (a.do_lpmd) s_rel[x - g0] = a.cpg_rel[x];
        }
        if (pl.staged) mbar_wait(&s_bar, pl.parity);
        __syncthreads();
        // ---- per read ----
#pragma unroll
        for (int i = 0; i < ING_RPT; i++) {
            const int idx = tid + i * ING_THREADS;
            if (idx < pl.ra || idx >= pl.rb) continue;
            const int64_t j = tile0 + idx;
            const uint32_t o0 = s_off[idx];
            ReadVerdict v = judge_read(a, j, rs[i], re[i], s_start[idx], rmeta[i], o0, s_off[idx + 1], lo, hi);
            err |= v.err;
            span_max = max(span_max, re[i] - rs[i] + 1);
            if (a.do_lpmd && !(rmeta[i] & META_HALO)) lp0++;  // lpmd.rs:176 (a halo copy is counted by its owner rank)
            if (v.flags & CF_LPMD) lp1++;
            const uint32_t n = v.n;
            if (n == 0) continue;
            const uint32_t y0 = o0 - g0;
            // calls are strictly increasing (checked per call below), so the two ends bound them all
            if (s_pos[y0] < rs[i] - 1 || s_pos[y0 + n - 1] > re[i]) err |= ERRBIT_POS_RANGE;
            uint32_t base = v.flags;
            if (a.do_pdr && n >= a.pdr.min_cpgs && (rmeta[i] & 0xFFu) >= a.pdr.min_qual) {  // pdr.rs:147-155
                bool disc;
                if (!rv.meth_off) {
                    const uint64_t m = low_mask64(n), x = rmw[i] & m;
                    disc = x != 0 && x != m;  // readutil.rs:134-145
                } else {
                    disc = read_discordant(rv, j, n);
                }
                base |= disc ? CF_PDR_D : CF_PDR_C;
            }
            if (!rv.meth_off) {
                uint64_t w = rmw[i];
                s_flags[y0] = (uint8_t)(base | CF_FIRST | (uint32_t)(w & 1ull));
                for (uint32_t k = 1; k < n; k++) {
                    w >>= 1;
                    s_flags[y0 + k] = (uint8_t)(base | (uint32_t)(w & 1ull));
                }
            } else {
                for (uint32_t k = 0; k < n; k++) s_flags[y0 + k] = (uint8_t)(base | meth_bit(rv, j, k) | (k == 0 ? CF_FIRST : 0u));
            }
        }
        __syncthreads();
        // ---- per CpG call ----
        const uint32_t ylo = lo - g0;
        for (uint32_t x = lo + tid; x < hi; x += ING_THREADS) {
            const uint32_t y = x - g0;
            const int32_t p = s_pos[y];
            const uint32_t f = s_flags[y];
            a.call_flags[x] = (uint8_t)f;
            if (!(f & CF_FIRST) && y > ylo) {
                if (p <= s_pos[y - 1]) err |= ERRBIT_CPG_ORDER;
                if (a.do_lpmd) {
                    const int32_t rk = s_rel[y];
                    if (rk <= (int32_t)s_rel[y - 1]) err |= ERRBIT_CPG_ORDER;  // query indices increase along a read
                    if (f & CF_LPMD) {
                        // readutil.rs:166-224: anchors are the read's earlier calls; those further than max_distance have
                        // been popped (:184), those closer than min_distance are skipped (:196)
                        for (uint32_t z = y - 1;; z--) {
                            const int32_t d = rk - (int32_t)s_rel[z];
                            if (d > dmax) break;
                            const uint32_t fz = s_flags[z];
                            if (d >= dmin) {
                                if (((fz ^ f) & CF_METH) == 0) lp2++; else lp3++;  // readutil.rs:200-214
                            }
                            if ((fz & CF_FIRST) || z == ylo) break;
                        }
                    }
                }
            }
            if (p < a.lin_lo - 1 || p >= a.lin_hi) { err |= ERRBIT_POS_RANGE; continue; }  // never touch memory outside the contig
            const uint32_t bit = (uint32_t)(p + 1);
            const uint32_t w = (bit >> 5) - wbase;
            if (w < (uint32_t)ING_WIN_WORDS) atomicOr(&s_bm[w], 1u << (bit & 31));
            else atomicOr(&bitmap32[bit >> 5], 1u << (bit & 31));
        }
        if (pl.rb >= nr) break;
        __syncthreads();  // everyone is done with this pass's slice
        if (tid == 0) plan_next(pl.rb);
        __syncthreads();
    }
    __syncthreads();
    for (int k = tid; k < ING_WIN_WORDS; k += ING_THREADS) {
        uint32_t v = s_bm[k];
        if (v) atomicOr(&bitmap32[wbase + k], v);
    }

    // ---- block reduction, then a handful of atomics per CTA ----
    const int lane = lane_id(), warp = tid >> 5;
    int32_t wmax = __reduce_max_sync(FULL, span_max);
    uint32_t werr = __reduce_or_sync(FULL, err);
    uint32_t wr = __reduce_add_sync(FULL, lp0), wv = __reduce_add_sync(FULL, lp1);
    uint32_t wc = __reduce_add_sync(FULL, lp2), wd = __reduce_add_sync(FULL, lp3);
    if (lane == 0) {
        s_red[warp][0] = (uint32_t)max(wmax, 0); s_red[warp][1] = werr; s_red[warp][2] = wv; s_red[warp][3] = wc; s_red[warp][4] = wd;
        s_red[warp][5] = wr;
    }
    __syncthreads();
    if (tid == 0) {
        int32_t bmax = 0;
        uint32_t berr = 0, bv = 0, bc = 0, bd = 0, br = 0;
#pragma unroll
        for (int w = 0; w < ING_THREADS / 32; w++) {
            bmax = max(bmax, (int32_t)s_red[w][0]); berr |= s_red[w][1]; bv += s_red[w][2]; bc += s_red[w][3]; bd += s_red[w][4];
            br += s_red[w][5];
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
    k_ingest<<<grid_for(a.n, ING_TILE), ING_THREADS, 0, s>>>(a);
    return 1;
}

}  // namespace mth
