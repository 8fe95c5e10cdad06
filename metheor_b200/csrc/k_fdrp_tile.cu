// k_fdrp_tile.cu — FDRP / qFDRP (fdrp.rs:12-145,176-246; qfdrp.rs:97-157,188-258): warp per site over TILE-STAGED reads.
//
// k_fdrp.cu rebuilds, for every site, the rank-space masks of every read of its pile (two passes over the reads' calls,
// a union bitmap, global scratch) although a read's verdict against another read does not depend on the site.  Here a
// CTA takes FT_SITES consecutive sites; the reads that can matter for them (one contiguous range of the sorted read
// arrays) are staged in shared memory ONCE: start, end, first call, eligibility, and their calls as 128-bit masks over
// the tile's site ranks (called / methylated, plus the one call that can lie outside [start, end]).  Each warp then
// takes sites of the tile: the window walk, the contributor test (one bit of the read's mask), the flush-trigger test
// and the segment replay all read shared memory, the pile is a list of 16-bit read indices, and the O(d^2) pair loop
// (32 pairs per step, lexicographic order, qFDRP's ordered f32 sum) ANDs two staged masks per pair.
//
// What does not fit — tiles with more than FT_RCAP reads, calls further than 32 sites from the tile (dense islands),
// max_depth > 64 — is flagged in `fallback` and done by k_fdrp afterwards (same results).  Two instances differ only in how
// many reads a tile may stage: CpG-dense contigs (chr19-like) need ~700 per 64 sites at 30x, whole-genome density ~1500.
#include <stdlib.h>
#include <string.h>

#include "gather.cuh"
#include "kernels.h"

namespace mth {

constexpr int FT_SITES = 64;       // sites per CTA (dense instance); the sparse instances take 32 or 64
constexpr int FT_THREADS = 256;
constexpr int FT_WARPS = FT_THREADS / 32;
constexpr int FT_RCAP = 1024;      // reads staged per tile (dense instance: chr19-like CpG density, ~700 reads per 64 sites at 30x)
constexpr int FT_RCAP_SPARSE = 2048;  // sparse instance: whole-genome density (a 64-site tile spans ~7 kb: ~1500 reads at 30x)
constexpr int FT_MARGIN = 32;      // site ranks representable before / after the tile
constexpr int FT_WIN = 256;        // 64-bit words of the site bitmap staged for rank lookups (16 384 positions)
constexpr int FT_MAXD = 64;        // pile slots per warp
constexpr int FT_MAX_READ_LEN = 201;  // fdrp.rs:10

// Pair t (lexicographic order over i < j, itertools combinations(2): fdrp.rs:128) of a pile of depth n, as i | j << 8, for every
// n <= FT_MAXD: the table of depth n starts at C(n, 3).  Replaces the per-lane index arithmetic of the pair loop (12 % of the
// kernel's instructions, profiles/R2c_lines_k_fdrp_tile.txt) by one cached load.  Filled once per device by k_fdrp_tables.
constexpr int FT_PAIR_TAB = (FT_MAXD + 1) * FT_MAXD * (FT_MAXD - 1) / 6;
__device__ uint16_t g_pair_tab[FT_PAIR_TAB];
// ham / shared for the small operands of qfdrp.rs:152, the very same f32 division done once
constexpr int FT_DIV_LUT = 32;
__device__ float g_div_lut[(FT_DIV_LUT + 1) * (FT_DIV_LUT + 1)];

__global__ void k_fdrp_tables() {
    const uint32_t n = blockIdx.x + 2;  // one CTA per depth 2..FT_MAXD
    uint16_t* tab = g_pair_tab + (size_t)n * (n - 1) * (n - 2) / 6;
    for (uint32_t i = threadIdx.x; i + 1 < n; i += blockDim.x) {
        const uint32_t t0 = i * (n - 1) - i * (i - 1) / 2;  // pairs before row i: sum_{r<i} (n - 1 - r)
        for (uint32_t j = i + 1; j < n; j++) tab[t0 + (j - i - 1)] = (uint16_t)(i | (j << 8));
    }
    if (blockIdx.x == 0)
        for (int k = threadIdx.x; k < (FT_DIV_LUT + 1) * (FT_DIV_LUT + 1); k += blockDim.x) {
            const int sh = k / (FT_DIV_LUT + 1), h = k % (FT_DIV_LUT + 1);
            g_div_lut[k] = sh ? __fdiv_rn((float)h, (float)sh) : 0.f;
        }
}

template <int RCAP>
struct FtSmem {
    unsigned long long cm[RCAP][2], mm[RCAP][2];
    int32_t start[RCAP], end[RCAP], first[RCAP];
    uint8_t flag[RCAP];       // 1: eligible (mapq >= min_qual, >= 1 call)
    uint8_t ub[RCAP];         // mask bit of the call outside [start, end] (reverse-strand call at start-1), 255 = none
    uint16_t pile[FT_WARPS][FT_MAXD];
    unsigned long long bmw[FT_WIN];
    uint32_t pref[FT_WIN];
    uint32_t wsum[FT_WARPS];
    long long ra;
    int nreads, bad;
};

template <int SITES, int RCAP>
__global__ void __launch_bounds__(FT_THREADS, RCAP <= 1024 ? 3 : 2) k_fdrp_tile(ReadsView rv, const int32_t* __restrict__ site_pos, int64_t C,
                                                             const unsigned long long* __restrict__ bitmap, int64_t n_words,
                                                             const uint32_t* __restrict__ word_prefix,
                                                             RegionScalars* __restrict__ scal, mth_fdrp_params prm, int quant,
                                                             uint64_t seed, ContigTable ct, float* __restrict__ value,
                                                             uint32_t* __restrict__ rowcnt, float* __restrict__ value_q,
                                                             uint32_t* __restrict__ rowcnt_q, uint8_t* __restrict__ fallback) {
    extern __shared__ __align__(16) unsigned char ft_raw[];
    FtSmem<RCAP>& sh = *reinterpret_cast<FtSmem<RCAP>*>(ft_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t lmax = scal->lmax;
    const uint32_t D = prm.max_depth;
    const int64_t n_tiles = (C + SITES - 1) / SITES;
    uint16_t* pile = sh.pile[warp];
    unsigned long long pair_ops = 0;  // per warp, flushed with one atomic at the end

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t s0 = tile * SITES;
        const int ns = (int)min((int64_t)SITES, C - s0);
        __syncthreads();  // previous tile fully consumed
        const int32_t p_first = site_pos[s0], p_last = site_pos[s0 + ns - 1];
        // window of the site bitmap for rank lookups: every call of a tile read lies in [p_first - lmax, p_last + lmax]
        const uint32_t w0 = (uint32_t)max(p_first - lmax, 0) >> 6;
        if (warp == 0) {
            const int64_t ra = warp_lower_bound(rv.start, rv.R, p_first - lmax + 1);
            const int64_t rb = warp_lower_bound(rv.start, rv.R, p_last + 2);
            if (lane == 0) {
                sh.ra = ra;
                sh.nreads = (int)min(rb - ra, (int64_t)RCAP + 1);
                sh.bad = (D > (uint32_t)FT_MAXD || ((uint32_t)(p_last + lmax + 1) >> 6) - w0 >= (uint32_t)FT_WIN) ? 1 : 0;
            }
        }
        {   // bitmap window + block exclusive scan of its popcounts (one word per thread)
            unsigned long long bw = 0;
            if ((int64_t)w0 + tid < n_words) bw = bitmap[w0 + tid];
            const uint32_t c = (uint32_t)__popcll(bw);
            uint32_t inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(FULL, inc, o);
                if (lane >= o) inc += t;
            }
            if (lane == 31) sh.wsum[warp] = inc;
            sh.bmw[tid] = bw;
            __syncthreads();
            uint32_t base = 0;
#pragma unroll
            for (int w = 0; w < FT_WARPS; w++) base += (w < warp) ? sh.wsum[w] : 0u;
            sh.pref[tid] = base + inc - c;
        }
        __syncthreads();
        const int64_t ra = sh.ra;
        const int nreads = sh.nreads;
        if (nreads > RCAP || sh.bad) {
            if (tid < ns) fallback[s0 + tid] = 1;
            continue;
        }
        // ---- stage the reads: scalars + call masks in tile rank space (bit b = site rank s0 - FT_MARGIN + b) ----
        const uint32_t rank0 = word_prefix[w0];
        const int64_t rank_base = s0 - FT_MARGIN;
        int overflow = 0;
        for (int r = tid; r < nreads; r += FT_THREADS) {
            const int64_t j = ra + r;
            const uint32_t o0 = rv.cpg_off[j], n = rv.cpg_off[j + 1] - o0;
            const int32_t st = rv.start[j], en = rv.end[j];
            sh.start[r] = st;
            sh.end[r] = en;
            sh.first[r] = n ? rv.cpg_pos[o0] : INT32_MIN;
            sh.flag[r] = ((rv.meta[j] & 0xFFu) >= prm.min_qual && n > 0) ? 1 : 0;  // fdrp.rs:205, :208
            unsigned long long c0 = 0, c1 = 0, m0 = 0, m1 = 0;
            uint32_t ub = 255;
            for (uint32_t k = 0; k < n; k++) {
                const int32_t x = rv.cpg_pos[o0 + k];
                const uint32_t bit = (uint32_t)(x + 1);
                const uint32_t w = (bit >> 6) - w0;
                const int64_t g = (int64_t)rank0 + sh.pref[w] + __popcll(sh.bmw[w] & ((1ull << (bit & 63)) - 1ull));
                const int64_t b = g - rank_base;
                if (b < 0 || b >= 128) { overflow = 1; continue; }
                const unsigned long long one = 1ull << (b & 63);
                const bool meth = meth_bit(rv, j, k) != 0;
                if (b < 64) { c0 |= one; if (meth) m0 |= one; } else { c1 |= one; if (meth) m1 |= one; }
                if (x < st || x > en) ub = (uint32_t)b;  // fdrp.rs:65-72: bit1/bit2 without bit0
            }
            sh.cm[r][0] = c0; sh.cm[r][1] = c1; sh.mm[r][0] = m0; sh.mm[r][1] = m1;
            sh.ub[r] = (uint8_t)ub;
        }
        if (__syncthreads_or(overflow)) {  // a read reaches beyond the representable ranks (dense CpG island)
            if (tid < ns) fallback[s0 + tid] = 1;
            continue;
        }

        // ---- one warp per site ----
        for (int ls = warp; ls < ns; ls += FT_WARPS) {
            const int32_t p = site_pos[s0 + ls];
            const uint32_t pbit = (uint32_t)(ls + FT_MARGIN);
            int lo = 0, hi = nreads;  // first read with start >= p - lmax + 1 (all lanes compute the same)
            while (lo < hi) { int m = (lo + hi) >> 1; if (sh.start[m] < p - lmax + 1) lo = m + 1; else hi = m; }
            uint32_t total = 0;
            float best = 0.f, best_q = 0.f;  // quant == 2 (both measures from one pair loop): best = FDRP, best_q = qFDRP
            bool have = false;
            const bool want_q = quant != 0, want_d = quant != 1;

            // All pairs (i < j) of the pile, 32 per step in lexicographic order (itertools combinations(2), fdrp.rs:128).
            // (A row-wise form — pile reads in registers, row i broadcast from shared memory — was measured slower: 143 ms
            // against 119 ms for FDRP on the whole-genome workload, twice as slow for qFDRP: 30 short rows instead of 12 full
            // steps; profiles/R2_summary.md.)
            auto evaluate = [&](uint32_t n) {
                const uint64_t P = (uint64_t)n * (n - 1) / 2;
                pair_ops += P;
                const uint16_t* __restrict__ tab = g_pair_tab + (size_t)n * (n - 1) * (n - 2) / 6;
                float acc = 0.f;
                uint32_t disc = 0;
                for (uint64_t t0 = 0; t0 < P; t0 += 32) {
                    float term = 0.f;
                    if (t0 + lane < P) {
                        const uint32_t e = __ldg(tab + t0 + lane);  // this lane's pair (i, j) of the step
                        const int ri = pile[e & 0xFFu], rj = pile[e >> 8];
                        const int32_t ov = min(sh.end[ri], sh.end[rj]) - max(sh.start[ri], sh.start[rj]) + 1;  // fdrp.rs:97-107
                        if (ov >= prm.min_overlap && ov > 0) {  // fdrp.rs:133-136 (without overlap: ham = 0, adds nothing)
                            const unsigned long long b0 = sh.cm[ri][0] & sh.cm[rj][0], b1 = sh.cm[ri][1] & sh.cm[rj][1];
                            unsigned long long v0 = b0, v1 = b1;  // both called AND both covered: drop the uncovered calls
                            const uint32_t ui = sh.ub[ri], uj = sh.ub[rj];
                            if ((ui & uj) != 255u) {  // rare: a reverse-strand call at start - 1 (fdrp.rs:65-72)
                                if (ui < 64u) v0 &= ~(1ull << ui); else if (ui < 128u) v1 &= ~(1ull << (ui - 64u));
                                if (uj < 64u) v0 &= ~(1ull << uj); else if (uj < 128u) v1 &= ~(1ull << (uj - 64u));
                            }
                            const uint32_t ham = (uint32_t)__popcll(v0 & (sh.mm[ri][0] ^ sh.mm[rj][0])) +
                                                 (uint32_t)__popcll(v1 & (sh.mm[ri][1] ^ sh.mm[rj][1]));  // fdrp.rs:109-122
                            if (want_q && ham) {  // qfdrp.rs:152
                                const uint32_t shr = (uint32_t)(__popcll(b0) + __popcll(b1));
                                term = shr <= (uint32_t)FT_DIV_LUT ? __ldg(&g_div_lut[shr * (FT_DIV_LUT + 1) + ham]) : __fdiv_rn((float)ham, (float)shr);
                            }
                            if (want_d) disc += ham ? 1u : 0u;  // fdrp.rs:138-140
                        }
                    }
                    if (want_q) {  // sequential f32 accumulation in pair order
                        uint32_t nz = __ballot_sync(FULL, term != 0.f);
                        if (__popc(nz) > 8) {
                            // dense step: fold all 32 lanes in order, branch-free (x + 0.0f == x exactly, acc >= 0)
#pragma unroll
                            for (int src = 0; src < 32; src++) acc = __fadd_rn(acc, __shfl_sync(FULL, term, src));
                        } else {
                            while (nz) {
                                const int src = __ffs(nz) - 1;
                                acc = __fadd_rn(acc, __shfl_sync(FULL, term, src));
                                nz &= nz - 1;
                            }
                        }
                    }
                }
                const float den = __fdiv_rn((float)((unsigned long long)n * (unsigned long long)(n - 1)), 2.0f);  // fdrp.rs:143
                if (want_d) {
                    const float v = __fdiv_rn((float)__reduce_add_sync(FULL, disc), den);
                    best = v;
                }
                if (want_q) {
                    const float v = __fdiv_rn(acc, den);
                    if (quant == 2) best_q = v; else best = v;
                }
            };
            auto close = [&]() {
                const uint32_t depth = min(total, D);
                if (depth > 0 && depth >= prm.min_depth) {  // fdrp.rs:215
                    __syncwarp();
                    evaluate(depth);
                    have = true;
                }
                total = 0;
                __syncwarp();
            };

            for (int c = 0;; c++) {
                const int r = lo + 32 * c + lane;
                const bool in = r < nreads && sh.start[r] <= p + 1;
                if (!__ballot_sync(FULL, in)) break;
                const bool elig = in && sh.flag[r];
                const bool calls = elig && ((sh.cm[r][pbit >> 6] >> (pbit & 63)) & 1ull);
                const bool trig = elig && !calls && sh.first[r] > p;  // fdrp.rs:213-222 (strict)
                uint32_t cm = __ballot_sync(FULL, calls), tm = __ballot_sync(FULL, trig);
                while (cm | tm) {  // replay contributors / triggers in file order
                    const int tpos = tm ? (__ffs(tm) - 1) : 32;
                    const uint32_t below = tpos >= 32 ? FULL : ((1u << tpos) - 1u);
                    const uint32_t cb = cm & below;
                    if (cb) {
                        // add_read, fdrp.rs:51-95: reads reaching beyond +-201 bp of the site are not added (no depth increment)
                        const bool mine = (cb >> lane) & 1u;
                        const bool acc = mine && !(sh.start[r] < p - FT_MAX_READ_LEN) && !(sh.end[r] > p + FT_MAX_READ_LEN);
                        const uint32_t am = __ballot_sync(FULL, acc);
                        const uint32_t t0 = total + __popc(am & ((1u << lane) - 1u));
                        int slot = -1;
                        if (acc) {
                            if (t0 < D) {
                                slot = (int)t0;
                            } else {  // reservoir: seeded draw instead of thread_rng
                                int32_t ctid, cpos;
                                delinearize(ct, p, &ctid, &cpos);
                                const uint32_t jd = reservoir_draw(seed, ctid, cpos, t0 + 1);
                                if (jd <= D) slot = (int)jd - 1;
                            }
                        }
                        const uint32_t wm = __ballot_sync(FULL, slot >= 0);
                        if (slot >= 0) {  // several replacements may hit one slot within a chunk: the latest read must win
                            const uint32_t same = __match_any_sync(wm, slot);
                            if (lane == 31 - __clz(same)) pile[slot] = (uint16_t)r;
                        }
                        total += __popc(am);
                        cm &= ~cb;
                        __syncwarp();
                    }
                    if (tpos < 32) {
                        close();
                        const int nc = cm ? (__ffs(cm) - 1) : 32;  // triggers up to the next contributor add nothing
                        tm &= ~(nc >= 32 ? FULL : ((1u << nc) - 1u));
                    }
                }
            }
            close();
            if (lane == 0) {
                value[s0 + ls] = best;
                rowcnt[s0 + ls] = have ? 1u : 0u;
                if (quant == 2) {
                    value_q[s0 + ls] = best_q;
                    rowcnt_q[s0 + ls] = have ? 1u : 0u;
                }
            }
        }
    }
    if (lane == 0 && pair_ops) atomicAdd(&scal->fdrp_pairs, pair_ops);
}

int launch_fdrp_tile(const ReadsView& rv, const int32_t* site_pos, int64_t C, const unsigned long long* bitmap, int64_t n_words,
                     const uint32_t* word_prefix, RegionScalars* sc, mth_fdrp_params prm, int quantitative, uint64_t seed,
                     ContigTable ct, float* value, uint32_t* rowcnt, float* value_q, uint32_t* rowcnt_q, uint8_t* fallback,
                     cudaStream_t s) {
    if (C <= 0) return 0;
    static bool tables_ready[64] = {false};
    {
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64 && !tables_ready[dev]) {  // the pair-index table and the division table, once per device
            k_fdrp_tables<<<FT_MAXD - 1, 64, 0, s>>>();
            cudaStreamSynchronize(s);  // once per device: other streams of this process may use the tables next
            tables_ready[dev] = true;
        }
    }
    static bool attr_set = false;
    static int force = 0;  // METHEOR_FDRP_TILE = dense | sparse32 | sparse64 | sparse16 | sparse8: kernel-variant experiments (profiles/)
    if (!attr_set) {
        cudaFuncSetAttribute(k_fdrp_tile<64, FT_RCAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FtSmem<FT_RCAP>));
        cudaFuncSetAttribute(k_fdrp_tile<32, FT_RCAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FtSmem<FT_RCAP>));
        cudaFuncSetAttribute(k_fdrp_tile<16, FT_RCAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FtSmem<FT_RCAP>));
        cudaFuncSetAttribute(k_fdrp_tile<8, FT_RCAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FtSmem<FT_RCAP>));
        cudaFuncSetAttribute(k_fdrp_tile<64, FT_RCAP_SPARSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FtSmem<FT_RCAP_SPARSE>));
        const char* e = getenv("METHEOR_FDRP_TILE");
        if (e) force = !strcmp(e, "dense") ? 1 : !strcmp(e, "sparse32") ? 2 : !strcmp(e, "sparse64") ? 3 : !strcmp(e, "sparse16") ? 4 : !strcmp(e, "sparse8") ? 5 : 0;
        attr_set = true;
    }
    // Reads a tile has to stage ~ sites x reads per site gap (coverage / CpG density): the instance with the most sites per
    // tile whose 1024-read capacity covers that with some room — 64 sites at chr19-like density and 30x, 32 at whole-genome
    // density and 30x, 16 at 60x, 8 (one site per warp) at 100x.  Without the small instances every tile of a 60x / 100x genome
    // overflowed and ALL sites went through the per-site kernel (245 of 290 ms, 340 of 471 ms of those passes).
    const double per_site = (double)rv.R / (double)C;
    int variant = force;
    if (!variant) {
        const double need = per_site * 1.3;
        variant = 64.0 * need <= (double)FT_RCAP ? 1 : 32.0 * need <= (double)FT_RCAP ? 2 : 16.0 * need <= (double)FT_RCAP ? 4 : 5;
    }
    const int sites = variant == 2 ? 32 : variant == 4 ? 16 : variant == 5 ? 8 : 64;
    int64_t tiles = (C + sites - 1) / sites;
    if (tiles > 148 * 48) tiles = 148 * 48;
#define MTH_FT_LAUNCH(S, R)                                                                                                              \
    k_fdrp_tile<S, R><<<(unsigned)tiles, FT_THREADS, sizeof(FtSmem<R>), s>>>(rv, site_pos, C, bitmap, n_words, word_prefix, sc, prm, quantitative, \
                                                                            seed, ct, value, rowcnt, value_q, rowcnt_q, fallback)
    if (variant == 1) MTH_FT_LAUNCH(64, FT_RCAP);
    else if (variant == 2) MTH_FT_LAUNCH(32, FT_RCAP);
    else if (variant == 4) MTH_FT_LAUNCH(16, FT_RCAP);
    else if (variant == 5) MTH_FT_LAUNCH(8, FT_RCAP);
    else MTH_FT_LAUNCH(64, FT_RCAP_SPARSE);
#undef MTH_FT_LAUNCH
    return 1;
}

}  // namespace mth
