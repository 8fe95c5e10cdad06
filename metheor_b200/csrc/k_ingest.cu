// k_ingest.cu — per-batch kernels: offset/coordinate fix-ups, validation, CpG-site marking and LPMD.
//
// k_ingest is one pass over a freshly copied batch (24 B/read + 4 B/CpG [+2 B/CpG for LPMD] of HBM reads):
//   * validates what every later kernel relies on (sorted starts, monotone offsets, CpG positions strictly
//     increasing and inside [start-1, end]) and records max(end-start+1) for the gather windows;
//   * marks each CpG position in the region's site bitmap (shared-memory window per tile, merged with atomicOr:
//     bits are only ever set);
//   * writes one flag byte per call (call_flags) carrying the read-level verdicts the per-call kernels need;
//   * LPMD (lpmd.rs:175-199 + readutil.rs:166-224): all in-read CpG pairs with min <= d(query index) <= max,
//     concordant iff equal methylation; block-reduced into four 64-bit device counters.
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
//   prologue : every thread issues the loads of the fixed-size fields of its reads; while they are in flight thread 0
//              reads the two slice bounds, arms an mbarrier and issues the TMA bulk copies of the slice into shared
//              memory (tma.cuh); offsets and starts then go to shared memory, the rest stays in registers;
//   per read : validation (sortedness, offsets, span -> lmax, first/last call inside [start-1, end]) and the read-level
//              decisions every later per-call step needs, written as one FLAG BYTE per call into shared memory:
//              methylated, first call of its read, read counts for LPMD (lpmd.rs:177), PDR state of the read
//              (filters pdr.rs:147-155 + concordance readutil.rs:134-145);
//   per call : every thread takes groups of 4 consecutive calls: one 128-bit load of the positions, one 32-bit load of
//              the flag bytes, one 64-bit load of the query indices, and the 4 flag bytes go to global memory
//              (call_flags) as one 32-bit store so that k_pdr_scatter is a pure per-call kernel.  Order check against
//              the previous call, OR of the site into a 16 384-position window of the site bitmap held in shared
//              memory.  LPMD pairs (readutil.rs:166-224): the two nearest anchors of every call are tested branch-free
//              from registers (packed key 4 * query index + methylation bit: one subtraction gives distance window
//              and concordance), the rare third and later anchors in one flat walk shared by the warp.  The first and
//              last group of a slice and calls outside the contig take a plain one-call-at-a-time path;
//   epilogue : non-zero window words merge into the global bitmap with fire-and-forget atomics; block reduction of the
//              scalars, <= 6 global atomics per CTA.
// After the prologue no thread waits on global memory again.  Tiles with more than ING_CAP calls (dense CpG islands)
// are processed in several passes over runs of reads whose calls fit the slice.  The kernel is instantiated per
// (meth_off present, LPMD requested) so that neither costs anything when absent.
// Tuning (profiles/r05_ingest_ab.md): 1024-read tiles; the LPMD instance needs 62 registers to stay spill-free, so it
// runs 4 CTAs per SM, the others 5.
#ifndef ING_RPT
#define ING_RPT 4
#endif
#ifndef ING_MINB_LPMD
#define ING_MINB_LPMD 4
#endif
#ifndef ING_MINB
#define ING_MINB 5
#endif
constexpr int ING_CPT = 4;  // calls per thread and sweep of the per-call phase
constexpr int ING_THREADS = 256;
constexpr int ING_TILE = ING_THREADS * ING_RPT;
constexpr int ING_CAP = 1024 * ING_RPT;   // calls staged per pass (avg 2.7 per read on WGBS); denser tiles take several passes
constexpr int ING_WIN_WORDS = 512;  // 32-bit words of the shared bitmap window = 16 384 positions

struct ReadVerdict {  // what the per-read step decides
    uint32_t err, n, flags;
    bool bad;
};

// Validation of one read + the read-level part of its calls' flag byte.  first / last = its first / last call position.
// WIDE: the region keeps its methylation words behind meth_off (some read has more than 64 calls).
template <bool WIDE, bool LPMD>
__device__ __forceinline__ ReadVerdict judge_read(const IngestArgs& a, int64_t j, int32_t s, int32_t e, int32_t prev_start,
                                                  uint32_t meta, uint32_t o0, uint32_t o1, uint32_t lo, uint32_t hi) {
    const ReadsView& rv = a.rv;
    ReadVerdict v;
    v.err = 0; v.bad = false; v.flags = 0;
    if (s < prev_start) v.err |= ERRBIT_UNSORTED;
    if ((uint32_t)(e - s) >= (uint32_t)MAX_REF_SPAN) v.err |= ERRBIT_SPAN;  // span = e - s + 1 must lie in [1, MAX_REF_SPAN]
    if (s < a.lin_lo || e >= a.lin_hi) v.err |= ERRBIT_POS_RANGE;
    if (o1 < o0 || o0 < lo || o1 > hi) { v.err |= ERRBIT_BAD_OFFSETS; o1 = o0; v.bad = true; }
    v.n = o1 - o0;
    if (v.n > (WIDE ? (uint32_t)MAX_CPGS_PER_READ : 64u)) { v.err |= ERRBIT_TOO_MANY_CPGS; v.n = 0; v.bad = true; }
    if (WIDE && v.n > 0) {
        uint32_t m0 = rv.meth_off[j], m1 = rv.meth_off[j + 1];
        if (m1 < m0 || (m1 - m0) * 64u < v.n) { v.err |= ERRBIT_BAD_OFFSETS; v.n = 0; v.bad = true; }
    }
    const uint32_t mapq = meta & 0xFFu;
    if (LPMD && !(meta & META_HALO) && !v.bad && mapq >= a.lpmd.min_qual) v.flags |= CF_LPMD;  // lpmd.rs:177
    if (a.do_pm && mapq >= a.pm_min_qual) v.flags |= CF_PM_OK;                                  // pm.rs:111
    if (a.do_me && mapq >= a.me_min_qual) v.flags |= CF_ME_OK;                                  // me.rs:115
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

template <bool WIDE, bool LPMD>
__global__ void __launch_bounds__(ING_THREADS, LPMD ? ING_MINB_LPMD : ING_MINB) k_ingest(IngestArgs a) {
    __shared__ __align__(16) int32_t s_pos[ING_CAP + 8];
    __shared__ __align__(16) uint16_t s_rel[LPMD ? ING_CAP + 8 : 8];
    __shared__ __align__(16) uint8_t s_flags[ING_CAP + 8];
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
    const uint16_t* rel_src = LPMD ? a.cpg_rel : nullptr;
    uint32_t tma_uses = 0;  // thread 0 only

    // fixed-size fields of this thread's reads: coalesced, all issued before anything waits (thread 0 plans the slice
    // copies while they are in flight)
    int32_t rs[ING_RPT], re[ING_RPT];
    uint32_t rmeta[ING_RPT], roff[ING_RPT];
    uint64_t rmw[ING_RPT];
#pragma unroll
    for (int i = 0; i < ING_RPT; i++) {
        const int idx = tid + i * ING_THREADS;
        rs[i] = 0; re[i] = 0; rmeta[i] = 0; rmw[i] = 0; roff[i] = 0;
        if (idx < nr) {
            const int64_t j = tile0 + idx;
            roff[i] = rv.cpg_off[j];
            rs[i] = rv.start[j]; re[i] = rv.end[j]; rmeta[i] = rv.meta[j];
            if (!WIDE) rmw[i] = rv.meth[j];
        }
    }
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
#pragma unroll
    for (int i = 0; i < ING_RPT; i++) {
        const int idx = tid + i * ING_THREADS;
        if (idx < nr) { s_off[idx] = roff[i]; s_start[idx + 1] = rs[i]; }
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
    const int32_t p_lo = a.lin_lo - 1, p_hi = a.lin_hi;   // a call position is valid in [p_lo, p_hi)
    // the distance window in units of K (see the per-call phase); query indices are < 65 536, so clamping loses nothing
    const int32_t dmin_c = max(min(dmin, 70000), -70000), dmax_c = max(min(dmax, 70000), -70000);
    const int32_t k4lo = dmax_c >= dmin_c ? 4 * dmin_c - 1 : INT32_MAX, k4hi = dmax_c >= dmin_c ? 4 * dmax_c + 1 : INT32_MIN;  // empty window: nothing matches
    const uint32_t k4span = dmax_c >= dmin_c ? (uint32_t)(k4hi - k4lo) : 0u;

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
            if (LPMD) s_rel[x - g0] = a.cpg_rel[x];
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
            ReadVerdict v = judge_read<WIDE, LPMD>(a, j, rs[i], re[i], s_start[idx], rmeta[i], o0, s_off[idx + 1], lo, hi);
            err |= v.err;
            span_max = max(span_max, re[i] - rs[i] + 1);
            if (LPMD && !(rmeta[i] & META_HALO)) lp0++;  // lpmd.rs:176 (a halo copy is counted by its owner rank)
            if (v.flags & CF_LPMD) lp1++;
            const uint32_t n = v.n;
            if (n == 0) continue;
            const uint32_t y0 = o0 - g0;
            // calls are strictly increasing (checked per call below), so the two ends bound them all
            if (s_pos[y0] < rs[i] - 1 || s_pos[y0 + n - 1] > re[i]) err |= ERRBIT_POS_RANGE;
            uint32_t base = v.flags;
            if (a.do_pdr && n >= a.pdr.min_cpgs && (rmeta[i] & 0xFFu) >= a.pdr.min_qual) {  // pdr.rs:147-155
                bool disc;
                if (!WIDE) {
                    const uint64_t m = low_mask64(n), x = rmw[i] & m;
                    disc = x != 0 && x != m;  // readutil.rs:134-145
                } else {
                    disc = read_discordant(rv, j, n);
                }
                base |= disc ? CF_PDR_D : CF_PDR_C;
            }
            if (!WIDE) {
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
        // ---- per CpG call: every thread takes groups of ING_CPT consecutive calls of the staged slice (vector loads from
        //      shared memory, one vector store of the flag bytes); the slice starts at g0, a multiple of 8 calls ----
        const uint32_t ylo = lo - g0, yhi = hi - g0;
        const uint32_t p_span = (uint32_t)(p_hi - p_lo);
        for (uint32_t y0 = (uint32_t)tid * ING_CPT; y0 < yhi; y0 += ING_THREADS * ING_CPT) {
            if (y0 + ING_CPT <= ylo) continue;  // the calls before `lo` belong to the previous tile / pass
            int32_t p[ING_CPT];
            uint32_t f4;       // flag byte of call y0 + k in bits [8k, 8k + 8)
            uint64_t rel4 = 0; // query index of call y0 + k in bits [16k, 16k + 16)
            const int4 v = *reinterpret_cast<const int4*>(&s_pos[y0]);
            p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
            f4 = *reinterpret_cast<const uint32_t*>(&s_flags[y0]);
            if (LPMD) rel4 = *reinterpret_cast<const uint64_t*>(&s_rel[y0]);
            bool fast = y0 >= ylo && y0 + ING_CPT <= yhi;  // all calls of the group belong to this pass ...
#pragma unroll
            for (int k = 0; k < ING_CPT; k++) fast = fast && (uint32_t)(p[k] - p_lo) < p_span;  // ... and lie inside the contig
            uint8_t* fdst = a.call_flags + (size_t)g0 + y0;  // g0 + y0 is a multiple of ING_CPT
            uint32_t need = 0;  // calls of this group that close LPMD pairs
            if (fast) {
                *reinterpret_cast<uint32_t*>(fdst) = f4;
                const bool head = y0 == ylo;  // call y0 opens the slice: nothing before it to compare with
                int32_t prev = head ? INT32_MIN : s_pos[y0 - 1];
                // LPMD (readutil.rs:166-224): the anchors of a call are the earlier calls of its read at query distance
                // dmin..dmax.  The two nearest anchors of every call are counted right here from registers (97 % of the
                // calls of WGBS reads have no third one), further ones in the walk below.  K = 4 * query index +
                // methylation bit of the calls y0 - 2 .. y0 + ING_CPT - 1: for an anchor pair, K_call - K_anchor =
                // 4 d + (-1 | 0 | +1), so d in [dmin, dmax] <=> the difference lies in [4 dmin - 1, 4 dmax + 1], and the
                // difference is odd exactly when the two calls disagree (readutil.rs:200-214).  fol = "an earlier call
                // of the same read precedes it".
                int32_t K[ING_CPT + 2];
                bool fol[ING_CPT + 2];
                uint32_t na = 0, nd = 0;  // anchor pairs of this group, discordant ones
                if (LPMD) {
                    uint32_t r2 = 0, f2 = CF_FIRST | (CF_FIRST << 8);
                    if (!head) {  // then y0 >= ING_CPT >= 2
                        r2 = *reinterpret_cast<const uint32_t*>(&s_rel[y0 - 2]);
                        f2 = *reinterpret_cast<const uint16_t*>(&s_flags[y0 - 2]);
                    }
                    K[0] = (int32_t)(((r2 & 0xFFFFu) << 2) | (f2 & CF_METH));
                    K[1] = (int32_t)(((r2 >> 16) << 2) | ((f2 >> 8) & CF_METH));
                    fol[0] = !(f2 & CF_FIRST) && y0 > ylo + 2;
                    fol[1] = !(f2 & (CF_FIRST << 8)) && y0 > ylo + 1;
                }
                bool disorder = false;
#pragma unroll
                for (int k = 0; k < ING_CPT; k++) {
                    const uint32_t f = f4 >> (8 * k);
                    const bool follows = !(f & CF_FIRST) && !(k == 0 && head);  // an earlier call of the same read precedes it
                    disorder |= follows && p[k] <= prev;
                    prev = p[k];
                    if (LPMD) {
                        const int i = k + 2;
                        K[i] = (int32_t)((((uint32_t)(rel4 >> (16 * k)) & 0xFFFFu) << 2) | (f & CF_METH));
                        fol[i] = follows;
                        const int32_t e1 = K[i] - K[i - 1], e2 = K[i] - K[i - 2];
                        disorder |= follows && e1 <= 1;  // query indices increase along a read (d1 <= 0)
                        const bool on = follows && (f & CF_LPMD);
                        const bool a1 = on && (uint32_t)(e1 - k4lo) <= k4span;                // nearest anchor
                        const bool a2 = on && fol[i - 1] && (uint32_t)(e2 - k4lo) <= k4span;  // second nearest
                        na += (uint32_t)a1 + (uint32_t)a2;
                        nd += (a1 ? (uint32_t)e1 & 1u : 0u) + (a2 ? (uint32_t)e2 & 1u : 0u);
                        // a third anchor can only exist when the chain of predecessors goes on and the second one is in reach
                        if (on && fol[i - 1] && fol[i - 2] && e2 <= k4hi) need |= 1u << k;
                    }
                    const uint32_t bit = (uint32_t)(p[k] + 1);
                    const uint32_t w = (bit >> 5) - wbase;
                    if (w < (uint32_t)ING_WIN_WORDS) atomicOr(&s_bm[w], 1u << (bit & 31));
                    else atomicOr(&bitmap32[bit >> 5], 1u << (bit & 31));
                }
                if (disorder) err |= ERRBIT_CPG_ORDER;
                if (LPMD) { lp2 += na - nd; lp3 += nd; }
            } else {
                // the first / last group of the slice, or a call outside the contig: one call at a time
#pragma unroll 1
                for (int k = 0; k < ING_CPT; k++) {
                    const uint32_t y = y0 + k;
                    if (y < ylo || y >= yhi) continue;
                    const int32_t pk = s_pos[y];
                    const uint32_t f = s_flags[y];
                    fdst[k] = (uint8_t)f;
                    if (!(f & CF_FIRST) && y > ylo) {
                        if (pk <= s_pos[y - 1]) err |= ERRBIT_CPG_ORDER;
                        if (LPMD) {
                            if (s_rel[y] <= s_rel[y - 1]) err |= ERRBIT_CPG_ORDER;
                            if (f & CF_LPMD) need |= 1u << k;
                        }
                    }
                    if ((uint32_t)(pk - p_lo) >= p_span) { err |= ERRBIT_POS_RANGE; continue; }  // never touch memory outside the contig
                    const uint32_t bit = (uint32_t)(pk + 1);
                    const uint32_t w = (bit >> 5) - wbase;
                    if (w < (uint32_t)ING_WIN_WORDS) atomicOr(&s_bm[w], 1u << (bit & 31));
                    else atomicOr(&bitmap32[bit >> 5], 1u << (bit & 31));
                }
            }
            if (LPMD) {
                // readutil.rs:166-224: anchors are the read's earlier calls; those further than max_distance have been
                // popped (:184), those closer than min_distance are skipped (:196).  One flat loop over (call, anchor)
                // steps so that the lanes of a warp walk their different calls side by side.  A fast group has counted
                // the two nearest anchors already and resumes at the third.
                const uint32_t zback = fast ? 3u : 1u;
                uint32_t z = 0, fk = 0;
                int32_t rk = 0;
                bool walking = false;
                for (;;) {
                    if (!walking) {
                        if (!need) break;
                        const int k = __ffs(need) - 1;
                        need &= need - 1;
                        rk = (int32_t)((rel4 >> (16 * k)) & 0xFFFFu);
                        fk = f4 >> (8 * k);
                        z = y0 + k - zback;
                        walking = true;
                    }
                    const int32_t d = rk - (int32_t)s_rel[z];
                    const uint32_t fz = s_flags[z];
                    if (d <= dmax && d >= dmin) {
                        if (((fz ^ fk) & CF_METH) == 0) lp2++; else lp3++;  // readutil.rs:200-214
                    }
                    if (d > dmax || (fz & CF_FIRST) || z == ylo) walking = false; else z--;
                }
            }
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
    uint32_t wr = 0, wv = 0, wc = 0, wd = 0;
    if (LPMD) {
        wr = __reduce_add_sync(FULL, lp0); wv = __reduce_add_sync(FULL, lp1);
        wc = __reduce_add_sync(FULL, lp2); wd = __reduce_add_sync(FULL, lp3);
    }
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
        if (LPMD) {
            if (br) atomicAdd(&a.sc->lpmd[0], (unsigned long long)br);
            if (bv) atomicAdd(&a.sc->lpmd[1], (unsigned long long)bv);
            if (bc) atomicAdd(&a.sc->lpmd[2], (unsigned long long)bc);
            if (bd) atomicAdd(&a.sc->lpmd[3], (unsigned long long)bd);
        }
    }
}

int launch_ingest(const IngestArgs& a, cudaStream_t s) {
    if (a.n <= 0) return 0;
    const unsigned grid = grid_for(a.n, ING_TILE);
    const bool wide = a.rv.meth_off != nullptr, lpmd = a.do_lpmd != 0;
    if (wide) {
        if (lpmd) k_ingest<true, true><<<grid, ING_THREADS, 0, s>>>(a);
        else k_ingest<true, false><<<grid, ING_THREADS, 0, s>>>(a);
    } else {
        if (lpmd) k_ingest<false, true><<<grid, ING_THREADS, 0, s>>>(a);
        else k_ingest<false, false><<<grid, ING_THREADS, 0, s>>>(a);
    }
    return 1;
}

}  // namespace mth
