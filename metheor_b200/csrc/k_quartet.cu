// k_quartet.cu — PM (epipolymorphism, pm.rs:85-128 + :42-51) and ME (methylation entropy, me.rs:90-132 + :42-55):
// 16-pattern histograms of 4 read-consecutive CpGs (readutil.rs:97-132).
//
// A quartet is keyed by its four positions; the warp that owns site p handles every quartet whose FIRST CpG is p.
// All reads calling p lie in p's window (gather.cuh); a read whose k-th CpG is p contributes key (pos[k+1],
// pos[k+2], pos[k+3]) and pattern 8*m[k]+4*m[k+1]+2*m[k+2]+m[k+3] when k+3 < n.  Reads that miss a call give a
// different key for the same p, so a site can own several quartets: the common single-key case is resolved in one
// pass over the window (count for the first key seen while checking that no other key occurs); otherwise keys are
// enumerated in ascending order, one counting pass each — exact for any number of keys and it yields the rows
// already sorted by (p1,p2,p3,p4) (the reference prints HashMap order, pm.rs:76).
// No flushing exists for PM/ME (whole-input maps), so there are no segments here.
// Two kernels from one template: COUNT (rows per site, for the output offsets) and EMIT (rows).
#include "gather.cuh"
#include "kernels.h"

namespace mth {

struct QKey {
    int32_t a, b, c;
};
__device__ __forceinline__ bool key_less(const QKey& x, const QKey& y) {
    if (x.a != y.a) return x.a < y.a;
    if (x.b != y.b) return x.b < y.b;
    return x.c < y.c;
}
__device__ __forceinline__ bool key_eq(const QKey& x, const QKey& y) { return x.a == y.a && x.b == y.b && x.c == y.c; }

// warp-wide lexicographic minimum of the keys of the lanes in `mask` (mask != 0)
__device__ __forceinline__ QKey warp_min_key(uint32_t mask, QKey k) {
    const int32_t BIG = INT32_MAX;
    bool in = (mask >> lane_id()) & 1u;
    int32_t a = in ? k.a : BIG;
    int32_t ma = __reduce_min_sync(FULL, a);
    bool ina = in && k.a == ma;
    int32_t b = ina ? k.b : BIG;
    int32_t mb = __reduce_min_sync(FULL, b);
    bool inb = ina && k.b == mb;
    int32_t cc = inb ? k.c : BIG;
    int32_t mc = __reduce_min_sync(FULL, cc);
    return QKey{ma, mb, mc};
}

// glibc / ARM optimized-routines log2f (sysdeps/ieee754/flt-32/e_log2f.c, LOG2F_TABLE_BITS = 4, N = 16), evaluated
// in double without contraction.  Only used for quartet depths beyond the host-libm table (see engine.cu
// build_me_lut); PARITY UNPINNED there (table constants restated from the published algorithm).
__device__ const double LOG2F_INVC[16] = {0x1.661ec79f8f3bep+0, 0x1.571ed4aaf883dp+0, 0x1.49539f0f010bp+0,  0x1.3c995b0b80385p+0,
                                          0x1.30d190c8864a5p+0, 0x1.25e227b0b8eap+0,  0x1.1bb4a4a1a343fp+0, 0x1.12358f08ae5bap+0,
                                          0x1.0953f419900a7p+0, 0x1p+0,               0x1.e608cfd9a47acp-1, 0x1.ca4b31f026aap-1,
                                          0x1.b2036576afce6p-1, 0x1.9c2d163a1aa2dp-1, 0x1.886e6037841edp-1, 0x1.767dcf5534862p-1};
__device__ const double LOG2F_LOGC[16] = {-0x1.efec65b963019p-2, -0x1.b0b6832d4fca4p-2, -0x1.7418b0a1fb77bp-2, -0x1.39de91a6dcf7bp-2,
                                          -0x1.01d9bf3f2b631p-2, -0x1.97c1d1b3b7afp-3,  -0x1.2f9e393af3c9fp-3, -0x1.960cbbf788d5cp-4,
                                          -0x1.a6f9db6475fcep-5, 0x0p+0,                0x1.338ca9f24f53dp-4,  0x1.476a9543891bap-3,
                                          0x1.e840b4ac4e4d2p-3,  0x1.40645f0c6651cp-2,  0x1.88e9c2c1b9ff8p-2,  0x1.ce0a44eb17bccp-2};
__device__ __forceinline__ float dev_log2f(float x) {
    const double* invc = LOG2F_INVC;
    const double* logc = LOG2F_LOGC;
    const double A0 = -0x1.712b6f70a7e4dp-2, A1 = 0x1.ecabf496832ep-2, A2 = -0x1.715479ffae3dep-1, A3 = 0x1.715475f35c8b8p0;
    uint32_t ix = __float_as_uint(x);
    if (ix == 0x3f800000u) return 0.f;
    uint32_t tmp = ix - 0x3f330000u;
    int i = (tmp >> 19) & 15;
    uint32_t top = tmp & 0xff800000u;
    uint32_t iz = ix - top;
    int k = (int32_t)tmp >> 23;
    double z = (double)__uint_as_float(iz);
    double r = __dadd_rn(__dmul_rn(z, invc[i]), -1.0);
    double y0 = __dadd_rn(logc[i], (double)k);
    double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(A1, r), A2);
    y = __dadd_rn(__dmul_rn(A0, r2), y);
    double pp = __dadd_rn(__dmul_rn(A3, r), y0);
    y = __dadd_rn(__dmul_rn(y, r2), pp);
    return (float)y;
}

__device__ __forceinline__ float pm_value(const uint32_t* cnt, uint32_t total) {  // pm.rs:42-51
    float pm = 1.0f;
    float ft = (float)total;
    for (int k = 0; k < 16; k++) {
        float q = __fdiv_rn((float)cnt[k], ft);
        pm = __fsub_rn(pm, __fmul_rn(q, q));
    }
    return pm;
}
__device__ __forceinline__ float me_value(const uint32_t* cnt, uint32_t total, const float* __restrict__ lut, int lut_max) {
    float me = 0.0f;  // me.rs:42-55
    float ft = (float)total;
    for (int k = 0; k < 16; k++) {
        uint32_t ck = cnt[k];
        if (ck == 0) continue;
        float t;
        if ((int)total <= lut_max) {
            t = lut[(size_t)total * (total + 1) / 2 + ck];
        } else {
            float p = __fdiv_rn((float)ck, ft);
            t = __fmul_rn(p, dev_log2f(p));
        }
        me = __fadd_rn(me, t);
    }
    return __fmul_rn(me, -0.25f);
}

struct QuartetArgs {
    ReadsView rv;
    const int32_t* site_pos;
    int64_t C;
    const RegionScalars* sc;
    mth_quartet_params prm;
    int kind;                 // 0 PM, 1 ME, 2 PM into rows and ME into rows_b (EMIT only)
    const uint8_t* mixed;     // nullptr: every site; else only the sites flagged by k_quartet_scatter
    const uint32_t* mixed_list;           // the flagged sites as a list (k_quartet_mixed_list), or nullptr: walk the flags
    const unsigned long long* mixed_n;    // its length (device)
    uint32_t* rowcnt;         // COUNT: out
    const uint32_t* rowoff;   // EMIT: in
    const float* me_lut;
    int me_lut_max;
    ContigTable ct;
    QuartetRowsDev rows;
    int64_t row_base;
    QuartetRowsDev rows_b;    // kind 2: the ME rows
    int64_t row_base_b;
};

constexpr int QCACHE = 4;  // chunks of 32 reads of a site's window whose quartets stay in registers between the passes

template <bool EMIT>
__global__ void __launch_bounds__(GATHER_BLOCK) k_quartet(QuartetArgs a) {
    __shared__ uint32_t s_cnt[GATHER_BLOCK / 32][16];
    uint32_t* cnt = s_cnt[threadIdx.x >> 5];
    const ReadsView& rv = a.rv;
    const int lane = lane_id();
    const uint32_t min_qual = a.prm.min_qual, min_depth = a.prm.min_depth;

    auto site_body = [&](int64_t s, int32_t p, int64_t lo, int32_t target) {
        uint32_t n_rows = 0;
        const int64_t out0 = EMIT ? (a.row_base + a.rowoff[s]) : 0;
        const int64_t out0b = EMIT ? (a.row_base_b + a.rowoff[s]) : 0;

        auto lane_quartet = [&](const LaneRead& lr, QKey* key, uint32_t* pat) -> bool {
            // pm.rs:111 / me.rs:115 mapq filter; readutil.rs:101-105 needs 4 CpGs from the site onwards
            if (lr.idx < 0 || lr.mapq < min_qual || (uint32_t)lr.idx + 3 >= lr.n) return false;
            const int32_t* cp = rv.cpg_pos + lr.o0 + lr.idx;
            key->a = cp[1]; key->b = cp[2]; key->c = cp[3];
            uint32_t k = (uint32_t)lr.idx;
            *pat = (meth_bit(rv, lr.j, k) << 3) | (meth_bit(rv, lr.j, k + 1) << 2) | (meth_bit(rv, lr.j, k + 2) << 1) |
                   meth_bit(rv, lr.j, k + 3);
            return true;
        };
        auto finish_key = [&](const QKey& key) {
            __syncwarp();
            uint32_t c = lane < 16 ? cnt[lane] : 0u;
            uint32_t total = __reduce_add_sync(FULL, c);
            if (total > 0 && total >= min_depth) {  // pm.rs:77 / me.rs:82
                if (EMIT && lane == 0) {
                    int64_t r = out0 + n_rows;
                    int32_t tid, pos;
                    delinearize(a.ct, p, &tid, &pos);
                    int32_t off = p - pos;  // linear offset of the contig
                    a.rows.tid[r] = tid;
                    a.rows.p1[r] = pos;
                    a.rows.p2[r] = key.a - off;
                    a.rows.p3[r] = key.b - off;
                    a.rows.p4[r] = key.c - off;
                    a.rows.value[r] = a.kind == 1 ? me_value(cnt, total, a.me_lut, a.me_lut_max) : pm_value(cnt, total);
                    if (a.rows.counts)
                        for (int k = 0; k < 16; k++) a.rows.counts[(size_t)r * 16 + k] = cnt[k];
                    if (a.kind == 2) {
                        const int64_t r2 = out0b + n_rows;
                        a.rows_b.tid[r2] = tid;
                        a.rows_b.p1[r2] = pos;
                        a.rows_b.p2[r2] = key.a - off;
                        a.rows_b.p3[r2] = key.b - off;
                        a.rows_b.p4[r2] = key.c - off;
                        a.rows_b.value[r2] = me_value(cnt, total, a.me_lut, a.me_lut_max);
                        if (a.rows_b.counts)
                            for (int k = 0; k < 16; k++) a.rows_b.counts[(size_t)r2 * 16 + k] = cnt[k];
                    }
                }
                n_rows++;
            }
            __syncwarp();
            if (lane < 16) cnt[lane] = 0;
            __syncwarp();
        };
        auto count_chunk = [&](bool match, uint32_t pat) {
            uint32_t mm = __ballot_sync(FULL, match);
            if (match) {
                uint32_t peers = __match_any_sync(mm, pat);
                if (lane == __ffs(peers) - 1) cnt[pat] += __popc(peers);  // one leader per pattern: no conflicts
            }
            __syncwarp();
        };

        if (lane < 16) cnt[lane] = 0;
        __syncwarp();

        // ---- pass 0: optimistic single-key pass ----
        // The lanes keep what they saw (key, pattern, valid) for the first QCACHE chunks of the window in registers: a site with
        // several keys is then resolved without touching memory again.  (Each pass over the window is a chain of dependent
        // loads — start, offsets, positions, bits — of several microseconds; mixed sites are rare, so a launch lasts as long
        // as ONE site takes and that latency, not throughput, is what the passes cost.)
        bool have_guess = false, mixed = false;
        QKey guess{0, 0, 0};
        QKey ck[QCACHE];
        uint32_t cpv[QCACHE];  // pattern | 16 when the lane's read of that chunk starts a quartet at p
#pragma unroll
        for (int c = 0; c < QCACHE; c++) { ck[c] = QKey{0, 0, 0}; cpv[c] = 0u; }
        int n_chunks = 0;
        scan_window(rv, lo, p, target, [&](const LaneRead& lr) {
            QKey key{0, 0, 0};
            uint32_t pat = 0;
            bool valid = lane_quartet(lr, &key, &pat);
#pragma unroll
            for (int c = 0; c < QCACHE; c++)
                if (c == n_chunks) { ck[c] = key; cpv[c] = valid ? (pat | 16u) : 0u; }
            n_chunks++;
            uint32_t vm = __ballot_sync(FULL, valid);
            if (!vm) return;
            if (!have_guess) {
                int src = __ffs(vm) - 1;
                guess.a = __shfl_sync(FULL, key.a, src);
                guess.b = __shfl_sync(FULL, key.b, src);
                guess.c = __shfl_sync(FULL, key.c, src);
                have_guess = true;
            }
            bool match = valid && key_eq(key, guess);
            if (__ballot_sync(FULL, match) != vm) mixed = true;
            count_chunk(match, pat);
        });
        if (have_guess && !mixed) {
            finish_key(guess);
        } else if (have_guess) {
            // ---- general path: enumerate keys in ascending order ----
            if (lane < 16) cnt[lane] = 0;
            __syncwarp();
            QKey last{INT32_MIN, INT32_MIN, INT32_MIN};
            bool first_round = true;
            if (n_chunks <= QCACHE) {  // from the registers
                while (true) {
                    bool cand = false;
                    QKey mine{INT32_MAX, INT32_MAX, INT32_MAX};
#pragma unroll
                    for (int c = 0; c < QCACHE; c++) {
                        const bool v = (cpv[c] & 16u) && (first_round || key_less(last, ck[c]));
                        if (v && (!cand || key_less(ck[c], mine))) { mine = ck[c]; cand = true; }
                    }
                    const uint32_t cmask = __ballot_sync(FULL, cand);
                    if (!cmask) break;
                    const QKey best = warp_min_key(cmask, mine);
#pragma unroll
                    for (int c = 0; c < QCACHE; c++)
                        if (c < n_chunks) count_chunk((cpv[c] & 16u) && key_eq(ck[c], best), cpv[c] & 15u);
                    finish_key(best);
                    last = best;
                    first_round = false;
                }
            } else {  // a window of more than 32 * QCACHE reads: one pass over the window per step
                while (true) {
                    bool found = false;
                    QKey best{INT32_MAX, INT32_MAX, INT32_MAX};
                    scan_window(rv, lo, p, target, [&](const LaneRead& lr) {
                        QKey key{0, 0, 0};
                        uint32_t pat = 0;
                        bool valid = lane_quartet(lr, &key, &pat);
                        bool cand = valid && (first_round || key_less(last, key));
                        uint32_t cmask = __ballot_sync(FULL, cand);
                        if (!cmask) return;
                        QKey m = warp_min_key(cmask, key);
                        if (!found || key_less(m, best)) best = m;
                        found = true;
                    });
                    if (!found) break;
                    scan_window(rv, lo, p, target, [&](const LaneRead& lr) {
                        QKey key{0, 0, 0};
                        uint32_t pat = 0;
                        bool valid = lane_quartet(lr, &key, &pat);
                        count_chunk(valid && key_eq(key, best), pat);
                    });
                    finish_key(best);
                    last = best;
                    first_round = false;
                }
            }
        }
        if (!EMIT && lane == 0) a.rowcnt[s] = n_rows;
    };
    if (a.mixed_list) for_each_listed_site(rv, a.site_pos, a.sc->lmax, a.mixed_list, *a.mixed_n, site_body);
    else for_each_site(rv, a.site_pos, a.C, a.sc->lmax, site_body, a.mixed);
}

// ---- streaming path: canonical quartets --------------------------------------------------------------------
// A quartet observation is a call x with three more calls of the same read behind it (readutil.rs:97-132); its key is
// (pos[x], pos[x+1], pos[x+2], pos[x+3]).  When those are four CONSECUTIVE sites of the dictionary — rank(pos[x+3]) ==
// rank(pos[x]) + 3, true unless a read skipped a site (no-call, deletion) — the key is implied by the first site: slot 0 of
// that site.  A read that skipped exactly ONE of the next four sites (a no-call) gives one of three other keys —
// (r+1,r+2,r+4), (r+1,r+3,r+4), (r+2,r+3,r+4) — the slots 1..3; in that order the four keys are already sorted the way the
// rows must be.  Anything else (two skips, deletions) flags the site as mixed and leaves it to the gather kernel.
// Everything needed is in cpg_pos[] and call_flags[] (5 B per call): methylation bits, read boundaries (CF_FIRST) and the
// mapq verdict; no per-read data is touched.
//
// Most sites never start a quartet (at whole-genome density 87 % of them have no row), so there is no per-site
// 16-pattern histogram.  Pass 1 (k_quartet_scatter) counts the observations per (site, slot) — 16 B per site — and
// appends every observation (site, slot, pattern) to a compact list; the row counts and the scan follow from the counts;
// pass 2 (k_quartet_hist) walks the LIST (not the calls) and adds each observation of a slot that reached min_depth into
// the 16-bin histogram of its output ROW (64 B per row, rows << sites); the emit kernel turns row histograms into PM / ME.
constexpr int QV = 4;  // key variants (slots) per site
constexpr int QS_THREADS = 256;
constexpr int QS_CPT = 8;
constexpr int QS_TILE = QS_THREADS * QS_CPT;
constexpr int QS_WIN = 256;  // 64-bit bitmap words in the shared window = 16 384 positions

__global__ void __launch_bounds__(QS_THREADS) k_quartet_scatter(const int32_t* __restrict__ cpg_pos, const uint8_t* __restrict__ call_flags,
                                                                int64_t n_calls, const unsigned long long* __restrict__ bitmap,
                                                                int64_t n_words, const uint32_t* __restrict__ word_prefix,
                                                                const RegionScalars* __restrict__ sc, uint32_t ok_bit,
                                                                uint32_t* __restrict__ qcnt, uint8_t* __restrict__ mixed,
                                                                uint32_t* __restrict__ obs_site, uint8_t* __restrict__ obs_vp,
                                                                unsigned long long* __restrict__ obs_n) {
    __shared__ int32_t s_pos[QS_TILE + 4];
    __shared__ uint8_t s_fl[QS_TILE + 4];
    __shared__ unsigned long long s_bmw[QS_WIN];
    __shared__ uint32_t s_pref[QS_WIN];
    __shared__ uint32_t s_wsum[QS_THREADS / 32];
    __shared__ uint32_t s_nobs;
    __shared__ unsigned long long s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t x0 = (int64_t)blockIdx.x * QS_TILE;
    const int nt = (int)min((int64_t)QS_TILE, n_calls - x0);
    const int nh = (int)min((int64_t)QS_TILE + 3, n_calls - x0);  // with the 3-call halo
    if (tid == 0) s_nobs = 0;
    for (int y = tid; y < nh; y += QS_THREADS) {
        s_pos[y] = cpg_pos[x0 + y];
        s_fl[y] = call_flags[x0 + y];
    }
    const int32_t wlo = max(cpg_pos[x0] - sc->lmax + 1, 0);  // every call of the tile is at or after P0 - lmax (reads are sorted)
    const uint32_t w0 = (uint32_t)wlo >> 6;
    unsigned long long bw = 0;
    if ((int64_t)w0 + tid < n_words) bw = bitmap[w0 + tid];
    const uint32_t rank0 = word_prefix[w0];
    {
        uint32_t c = (uint32_t)__popcll(bw), inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_wsum[warp] = inc;
        s_bmw[tid] = bw;
        __syncthreads();
        uint32_t base = 0;
#pragma unroll
        for (int w = 0; w < QS_THREADS / 32; w++) base += (w < warp) ? s_wsum[w] : 0u;
        s_pref[tid] = base + inc - c;
    }
    __syncthreads();
    auto rank_of = [&](int32_t p) -> uint32_t {
        const uint32_t bit = (uint32_t)(p + 1);
        const uint32_t w = (bit >> 6) - w0;
        if (w < (uint32_t)QS_WIN) return rank0 + s_pref[w] + (uint32_t)__popcll(s_bmw[w] & ((1ull << (bit & 63)) - 1ull));
        const uint32_t gw = bit >> 6;
        return __ldg(word_prefix + gw) + (uint32_t)__popcll(__ldg(bitmap + gw) & ((1ull << (bit & 63)) - 1ull));
    };
    uint32_t o_site[QS_CPT], o_slot[QS_CPT];  // this thread's observations: site rank, (list index << 8) | slot << 4 | pattern
    int n_mine = 0;
#pragma unroll
    for (int it = 0; it < QS_CPT; it++) {
        const int y = it * QS_THREADS + tid;
        o_site[it] = 0xffffffffu;
        o_slot[it] = 0;
        if (y >= nt || y + 3 >= nh) continue;
        const uint32_t f0 = s_fl[y];
        if (!(f0 & ok_bit)) continue;
        const uint32_t f1 = s_fl[y + 1], f2 = s_fl[y + 2], f3 = s_fl[y + 3];
        if ((f1 | f2 | f3) & CF_FIRST) continue;  // fewer than four calls left in this read
        const uint32_t pat = ((f0 & CF_METH) << 3) | ((f1 & CF_METH) << 2) | ((f2 & CF_METH) << 1) | (f3 & CF_METH);
        const uint32_t r = rank_of(s_pos[y]);
        const uint32_t g3 = rank_of(s_pos[y + 3]) - r;
        int v = -1;  // slot: which of the dictionary's next sites the read called
        if (g3 == 3) {
            v = 0;                                             // (r+1, r+2, r+3)
        } else if (g3 == 4) {
            const uint32_t g1 = rank_of(s_pos[y + 1]) - r, g2 = rank_of(s_pos[y + 2]) - r;
            if (g1 == 1) v = g2 == 2 ? 1 : 2;                  // (r+1, r+2, r+4) / (r+1, r+3, r+4)
            else v = 3;                                        // (r+2, r+3, r+4)
        }
        if (v >= 0) {
            atomicAdd(&qcnt[(size_t)r * QV + v], 1u);
            o_site[it] = r;
            o_slot[it] = ((uint32_t)v << 4) | pat;
            n_mine++;
        } else {
            mixed[r] = 1;
        }
    }
    // append this CTA's observations to the list: one shared counter, ONE global atomic per CTA
    const uint32_t my0 = n_mine ? atomicAdd(&s_nobs, (uint32_t)n_mine) : 0u;
    __syncthreads();
    if (tid == 0) s_base = s_nobs ? atomicAdd(obs_n, (unsigned long long)s_nobs) : 0ull;
    __syncthreads();
    unsigned long long at = s_base + my0;
#pragma unroll
    for (int it = 0; it < QS_CPT; it++) {
        if (o_site[it] != 0xffffffffu) {
            obs_site[at] = o_site[it];
            obs_vp[at] = (uint8_t)o_slot[it];
            at++;
        }
    }
}

// rows per site from the slot counts: slots that reached min_depth (pm.rs:77 / me.rs:82), in slot order
__global__ void __launch_bounds__(256) k_quartet_canon_count(const uint32_t* __restrict__ qcnt, const uint8_t* __restrict__ mixed, int64_t C,
                                                             uint32_t min_depth, uint32_t* __restrict__ rowcnt) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= C) return;
    if (mixed[s]) return;  // rowcnt[s] of a mixed site belongs to the gather pass, which may be running beside this kernel
    const uint4 q = reinterpret_cast<const uint4*>(qcnt)[s];
    const uint32_t md = max(min_depth, 1u);
    rowcnt[s] = (q.x >= md) + (q.y >= md) + (q.z >= md) + (q.w >= md);
}

// pass 2: every listed observation whose (site, slot) reached min_depth goes into the histogram of its row
__global__ void __launch_bounds__(256) k_quartet_hist(const uint32_t* __restrict__ obs_site, const uint8_t* __restrict__ obs_vp,
                                                      const unsigned long long* __restrict__ obs_n, const uint32_t* __restrict__ qcnt,
                                                      const uint8_t* __restrict__ mixed, const uint32_t* __restrict__ rowoff,
                                                      uint32_t min_depth, uint32_t* __restrict__ hrows) {
    const unsigned long long n = *obs_n;
    const uint32_t md = max(min_depth, 1u);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t s = obs_site[i];
        if (mixed[s]) continue;
        const uint32_t vp = obs_vp[i], v = vp >> 4, pat = vp & 15u;
        const uint4 q = reinterpret_cast<const uint4*>(qcnt)[s];
        const uint32_t c[4] = {q.x, q.y, q.z, q.w};
        if (c[v] < md) continue;
        uint32_t before = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) before += (k < (int)v && c[k] >= md) ? 1u : 0u;
        atomicAdd(&hrows[((size_t)rowoff[s] + before) * 16 + pat], 1u);
    }
}

// rows of the canonical sites from the row histograms (thread per site); mixed sites are skipped (gather kernels).
// BOTH: PM into rows_a and ME into rows_b from one read of the histograms (same thresholds: same rows).
template <int KIND>  // 0 PM, 1 ME, 2 both
__global__ void __launch_bounds__(256) k_quartet_canon_emit(const uint32_t* __restrict__ qcnt, const uint8_t* __restrict__ mixed,
                                                            const uint32_t* __restrict__ hrows, const int32_t* __restrict__ site_pos,
                                                            int64_t C, uint32_t min_depth, const uint32_t* __restrict__ rowoff,
                                                            const float* __restrict__ me_lut, int me_lut_max, ContigTable ct,
                                                            QuartetRowsDev rows_a, int64_t base_a, QuartetRowsDev rows_b, int64_t base_b) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= C || mixed[s]) return;
    const uint4 q = reinterpret_cast<const uint4*>(qcnt)[s];
    const uint32_t c[4] = {q.x, q.y, q.z, q.w};
    const uint32_t md = max(min_depth, 1u);
    if (!(c[0] >= md || c[1] >= md || c[2] >= md || c[3] >= md)) return;
    int32_t tid, pos;
    delinearize(ct, site_pos[s], &tid, &pos);
    const int32_t off = site_pos[s] - pos;
    int64_t r = rowoff[s];
#pragma unroll
    for (int v = 0; v < QV; v++) {
        if (c[v] < md) continue;
        uint32_t cnt[16];
        const uint4* h = reinterpret_cast<const uint4*>(hrows + (size_t)r * 16);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint4 x = h[k];
            cnt[4 * k] = x.x; cnt[4 * k + 1] = x.y; cnt[4 * k + 2] = x.z; cnt[4 * k + 3] = x.w;
        }
        const uint32_t total = c[v];
        // the key of slot v: sites s+1..s+4 of the dictionary minus the one the reads skipped
        const int d2 = v == 3 ? 2 : 1, d3 = v >= 2 ? 3 : 2, d4 = v >= 1 ? 4 : 3;
        const int32_t p2 = site_pos[s + d2] - off, p3 = site_pos[s + d3] - off, p4 = site_pos[s + d4] - off;
        if (KIND == 0 || KIND == 2) {
            const int64_t o = base_a + r;
            rows_a.tid[o] = tid; rows_a.p1[o] = pos; rows_a.p2[o] = p2; rows_a.p3[o] = p3; rows_a.p4[o] = p4;
            rows_a.value[o] = pm_value(cnt, total);
            if (rows_a.counts)
                for (int k = 0; k < 16; k++) rows_a.counts[(size_t)o * 16 + k] = cnt[k];
        }
        if (KIND == 1 || KIND == 2) {
            QuartetRowsDev& rw = KIND == 1 ? rows_a : rows_b;
            const int64_t o = (KIND == 1 ? base_a : base_b) + r;
            rw.tid[o] = tid; rw.p1[o] = pos; rw.p2[o] = p2; rw.p3[o] = p3; rw.p4[o] = p4;
            rw.value[o] = me_value(cnt, total, me_lut, me_lut_max);
            if (rw.counts)
                for (int k = 0; k < 16; k++) rw.counts[(size_t)o * 16 + k] = cnt[k];
        }
        r++;
    }
}

int launch_quartet_scatter(const int32_t* cpg_pos, const uint8_t* call_flags, int64_t n_calls, const unsigned long long* bitmap,
                           int64_t n_words, const uint32_t* word_prefix, const RegionScalars* sc, uint32_t ok_bit, uint32_t* qcnt,
                           uint8_t* mixed, uint32_t* obs_site, uint8_t* obs_vp, unsigned long long* obs_n, cudaStream_t s) {
    if (n_calls <= 0) return 0;
    k_quartet_scatter<<<(unsigned)((n_calls + QS_TILE - 1) / QS_TILE), QS_THREADS, 0, s>>>(cpg_pos, call_flags, n_calls, bitmap, n_words,
                                                                                          word_prefix, sc, ok_bit, qcnt, mixed, obs_site,
                                                                                          obs_vp, obs_n);
    return 1;
}

int launch_quartet_canon_count(const uint32_t* qcnt, const uint8_t* mixed, int64_t C, uint32_t min_depth, uint32_t* rowcnt, cudaStream_t s) {
    if (C <= 0) return 0;
    k_quartet_canon_count<<<(unsigned)((C + 255) / 256), 256, 0, s>>>(qcnt, mixed, C, min_depth, rowcnt);
    return 1;
}

int launch_quartet_hist(const uint32_t* obs_site, const uint8_t* obs_vp, const unsigned long long* obs_n, int64_t n_obs_max,
                        const uint32_t* qcnt, const uint8_t* mixed, const uint32_t* rowoff, uint32_t min_depth, uint32_t* hrows,
                        cudaStream_t s) {
    if (n_obs_max <= 0) return 0;
    int64_t blocks = (n_obs_max + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    k_quartet_hist<<<(unsigned)blocks, 256, 0, s>>>(obs_site, obs_vp, obs_n, qcnt, mixed, rowoff, min_depth, hrows);
    return 1;
}

int launch_quartet_canon_emit(const uint32_t* qcnt, const uint8_t* mixed, const uint32_t* hrows, const int32_t* site_pos, int64_t C,
                              uint32_t min_depth, int kind, const uint32_t* rowoff, const float* me_lut, int me_lut_max, ContigTable ct,
                              QuartetRowsDev rows_a, int64_t base_a, QuartetRowsDev rows_b, int64_t base_b, cudaStream_t s) {
    if (C <= 0) return 0;
    const unsigned g = (unsigned)((C + 255) / 256);
    if (kind == 0)
        k_quartet_canon_emit<0><<<g, 256, 0, s>>>(qcnt, mixed, hrows, site_pos, C, min_depth, rowoff, me_lut, me_lut_max, ct, rows_a, base_a, rows_b, base_b);
    else if (kind == 1)
        k_quartet_canon_emit<1><<<g, 256, 0, s>>>(qcnt, mixed, hrows, site_pos, C, min_depth, rowoff, me_lut, me_lut_max, ct, rows_a, base_a, rows_b, base_b);
    else
        k_quartet_canon_emit<2><<<g, 256, 0, s>>>(qcnt, mixed, hrows, site_pos, C, min_depth, rowoff, me_lut, me_lut_max, ct, rows_a, base_a, rows_b, base_b);
    return 1;
}

// the flagged sites as a list: 16 flags per thread, the (few) hits appended with one atomic each; order does not matter
__global__ void __launch_bounds__(256) k_quartet_mixed_list(const uint8_t* __restrict__ mixed, int64_t C, uint32_t* __restrict__ list,
                                                            unsigned long long* __restrict__ n_list) {
    const int64_t s0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (s0 >= C) return;
    if (s0 + 16 <= C) {  // `mixed` has 64 bytes of slack and is 16-byte aligned (cudaMalloc)
        const uint4 v = *reinterpret_cast<const uint4*>(mixed + s0);
        if (!(v.x | v.y | v.z | v.w)) return;
    }
    for (int64_t s = s0; s < min(C, s0 + 16); s++)
        if (mixed[s]) list[atomicAdd(n_list, 1ull)] = (uint32_t)s;
}

int launch_quartet_mixed_list(const uint8_t* mixed, int64_t C, uint32_t* list, unsigned long long* n_list, cudaStream_t s) {
    if (C <= 0) return 0;
    k_quartet_mixed_list<<<(unsigned)((C + 4095) / 4096), 256, 0, s>>>(mixed, C, list, n_list);
    return 1;
}

// grid of the gather kernels over a list of flagged sites: they are few (one warp each, a few per SM)
static int listed_grid(int64_t C) { return min(gather_grid(C), 148 * 8); }

int launch_quartet_count(const ReadsView& rv, const int32_t* site_pos, int64_t C, const RegionScalars* sc,
                         mth_quartet_params prm, const uint8_t* mixed, const uint32_t* mixed_list, const unsigned long long* mixed_n,
                         uint32_t* rowcnt, cudaStream_t s) {
    if (C <= 0) return 0;
    QuartetArgs a;
    memset(&a, 0, sizeof(a));
    a.rv = rv; a.site_pos = site_pos; a.C = C; a.sc = sc; a.prm = prm; a.rowcnt = rowcnt; a.mixed = mixed;
    a.mixed_list = mixed_list; a.mixed_n = mixed_n;
    k_quartet<false><<<mixed_list ? listed_grid(C) : gather_grid(C), GATHER_BLOCK, 0, s>>>(a);
    return 1;
}

int launch_quartet_emit(const ReadsView& rv, const int32_t* site_pos, int64_t C, const RegionScalars* sc,
                        mth_quartet_params prm, int kind, const uint8_t* mixed, const uint32_t* mixed_list, const unsigned long long* mixed_n,
                        const uint32_t* rowoff, const float* me_lut,
                        int me_lut_max, ContigTable ct, QuartetRowsDev rows, int64_t row_base, QuartetRowsDev rows_b, int64_t row_base_b,
                        cudaStream_t s) {
    if (C <= 0) return 0;
    QuartetArgs a;
    memset(&a, 0, sizeof(a));
    a.rv = rv; a.site_pos = site_pos; a.C = C; a.sc = sc; a.prm = prm; a.kind = kind; a.rowoff = rowoff; a.mixed = mixed;
    a.mixed_list = mixed_list; a.mixed_n = mixed_n;
    a.me_lut = me_lut; a.me_lut_max = me_lut_max; a.ct = ct; a.rows = rows; a.row_base = row_base; a.rows_b = rows_b; a.row_base_b = row_base_b;
    k_quartet<true><<<mixed_list ? listed_grid(C) : gather_grid(C), GATHER_BLOCK, 0, s>>>(a);
    return 1;
}

}  // namespace mth
