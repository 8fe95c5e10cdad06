// k_mhl_site.cu — methylation haplotype load (mhl.rs:135-208, AssociatedReads mhl.rs:12-81), one THREAD per CpG site.
//
// The warp-per-site form (k_mhl.cu) is issue-bound on cross-lane bookkeeping: a site only has ~30 reads to look at.
// Here MS_SITES consecutive sites form a tile; the reads that can matter for them are one contiguous range of the sorted
// read arrays and are staged in shared memory once (start, first call, call count | mapq, methylation word, and the
// calls themselves); then every thread replays the reference's streaming loop for its own site, sequentially and in
// file order, entirely out of shared memory: a read contributes if the site is among its calls, a read whose first
// call lies behind the site closes the segment (mhl.rs:162-173), the per-read stretch histogram is the
// x_l = x_{l-1} & (x_{l-1} >> 1) popcount chain, and the per-site S[l] / N[n] accumulators are two small
// shared-memory arrays per thread.  32 sites advance per warp instruction instead of one.
//
// What does not fit — tiles with more than MS_RCAP reads or MS_CCAP calls or > 64 calls per read (whole tile), sites with a
// contributing read of more than MS_L calls (dense CpG islands; only those sites) or a segment deeper than the 16-bit
// accumulators — is flagged in `fallback` and done by the warp-per-site kernel afterwards.
#include <stdlib.h>
#include <string.h>

#include "gather.cuh"
#include "kernels.h"

namespace mth {

constexpr int MS_SITES = 128;    // sites (= threads) per CTA
constexpr int MS_RCAP = 2048;    // reads staged per tile (dense instance: chr19-like density, ~1400 reads per 128 sites at 30x)
constexpr int MS_CCAP = 6144;    // calls staged per tile
// sparse instance: whole-genome density (a 128-site tile spans ~14 kb: ~2900 reads, ~3800 calls at 30x, sd ~9 %).  54 KB of shared
// memory: FOUR resident CTAs per SM instead of the three a 4096 / 8192 instance allows — the kernel is latency-bound, warps in
// flight are what it lacks; the few tiles over capacity go to the per-site kernel.
constexpr int MS_RCAP_SPARSE = 3584;
constexpr int MS_CCAP_SPARSE = 5120;
constexpr int MS_L = 16;         // longest read (in calls) the per-thread accumulators hold
constexpr int MS_SPAN = 60000;   // positions a tile may span (calls are staged as 16-bit offsets)

// Shared-memory footprint decides how many sites are in flight per SM, so everything is packed: 10 bytes per read,
// 2 bytes per call, 16-bit accumulators (a segment deeper than 65 535 reads goes to the per-site kernel).
template <int RCAP, int CCAP>
struct MsSmem {
    // one 64-bit word per read, fetched with ONE shared-memory load in the per-site loop:
    //   bits 0-15 start (offset from the tile base) | 16-31 first call (offset; 0xFFFF: no call) | 32-47 index of its first call
    //   in pos[] | 48-55 mapq | 56-63 number of calls (<= MS_L)
    unsigned long long rec[RCAP];
    uint16_t mbits[RCAP];     // methylation bits of the read's calls (n <= MS_L = 16 calls in this kernel)
    uint16_t pos[CCAP];       // calls, as offsets from the tile base
    uint16_t S[MS_L + 1][MS_SITES];   // [l][thread]: bank-conflict-free
    uint16_t N[MS_L + 1][MS_SITES];
    long long ra;
    unsigned int c0;
    int nreads, ncalls, bad;
};

template <int RCAP, int CCAP>
__global__ void __launch_bounds__(MS_SITES) k_mhl_site(ReadsView rv, const int32_t* __restrict__ site_pos, int64_t C,
                                                       const RegionScalars* __restrict__ sc, mth_mhl_params prm,
                                                       float* __restrict__ value, uint32_t* __restrict__ rowcnt,
                                                       uint8_t* __restrict__ fallback) {
    extern __shared__ __align__(16) unsigned char ms_raw[];
    MsSmem<RCAP, CCAP>& sh = *reinterpret_cast<MsSmem<RCAP, CCAP>*>(ms_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t lmax = sc->lmax;
    const int64_t n_tiles = (C + MS_SITES - 1) / MS_SITES;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t s0 = tile * MS_SITES;
        const int ns = (int)min((int64_t)MS_SITES, C - s0);
        __syncthreads();  // previous tile fully consumed
        if (warp == 0) {
            const int64_t ra = warp_lower_bound(rv.start, rv.R, site_pos[s0] - lmax + 1);
            const int64_t rb = warp_lower_bound(rv.start, rv.R, site_pos[s0 + ns - 1] + 2);
            if (lane == 0) {
                const unsigned int c0 = rv.cpg_off[ra], c1 = rv.cpg_off[rb];
                sh.ra = ra;
                sh.c0 = c0;
                sh.nreads = (int)min(rb - ra, (int64_t)RCAP + 1);
                sh.ncalls = (int)min(c1 - c0, (unsigned int)CCAP + 1u);
                // multi-word reads or a tile too wide for 16-bit offsets: per-site kernel
                sh.bad = (rv.meth_off != nullptr || site_pos[s0 + ns - 1] - site_pos[s0] + 2 * lmax > MS_SPAN) ? 1 : 0;
            }
        }
        __syncthreads();
        const int64_t ra = sh.ra;
        const uint32_t c0 = sh.c0;
        const int nreads = sh.nreads, ncalls = sh.ncalls;
        if (nreads > RCAP || ncalls > CCAP || sh.bad) {
            if (tid < ns) fallback[s0 + tid] = 1;
            continue;
        }
        const int32_t base = site_pos[s0] - lmax - 1;  // every call of a tile read is >= base
        // (unrolling the two staging loops x4 / x8 to put more loads in flight per thread was measured: 21.7 ms against 20.7 ms
        // for the whole-genome pass — no gain)
        for (int r = tid; r < nreads; r += MS_SITES) {
            const int64_t j = ra + r;
            const uint32_t o0 = rv.cpg_off[j], n = rv.cpg_off[j + 1] - o0;
            sh.rec[r] = (unsigned long long)(uint16_t)(rv.start[j] - base) | 0xFFFF0000ull | ((unsigned long long)(uint16_t)(o0 - c0) << 32) |
                        ((unsigned long long)(rv.meta[j] & 0xFFu) << 48) | ((unsigned long long)min(n, 255u) << 56);
            sh.mbits[r] = (uint16_t)rv.meth[j];  // staged once (coalesced) instead of one global load per (site, read)
        }
        for (int y = tid; y < ncalls; y += MS_SITES) sh.pos[y] = (uint16_t)(rv.cpg_pos[c0 + y] - base);
        __syncthreads();
        for (int r = tid; r < nreads; r += MS_SITES) {
            const unsigned long long w = sh.rec[r];
            if (w >> 56) sh.rec[r] = (w & ~0xFFFF0000ull) | ((unsigned long long)sh.pos[(uint32_t)(w >> 32) & 0xFFFFu] << 16);
        }
        __syncthreads();
        if (tid >= ns) continue;

        // ---- one thread per site: mhl.rs:155-205 over the site's window, in file order ----
        const int32_t p = site_pos[s0 + tid];
        const uint32_t pq = (uint32_t)(p - base);  // the site in tile offsets
        bool deep = false;
        const uint32_t key_lo = pq - (uint32_t)lmax + 1u, key_hi = pq + 1u;  // reads with start in [p - lmax + 1, p + 1], as tile offsets
        int lo = 0, hi = nreads;  // first read with start >= p - lmax + 1
        while (lo < hi) { int m = (lo + hi) >> 1; if (((uint32_t)sh.rec[m] & 0xFFFFu) < key_lo) lo = m + 1; else hi = m; }
#pragma unroll
        for (int l = 0; l <= MS_L; l++) { sh.S[l][tid] = 0; sh.N[l][tid] = 0; }
        uint32_t depth = 0, maxn = 0;
        float best = 0.f;
        bool have = false;
        auto close = [&]() {
            if (depth == 0) return;
            if (depth >= prm.min_depth) {  // mhl.rs:165
                // D[l] = sum over reads of max(0, n - l + 1) = suffix sum of the suffix counts of N (mhl.rs:56-64)
                uint32_t ge = 0, dl = 0;
                uint32_t Dl[MS_L + 1];
#pragma unroll
                for (int l = MS_L; l >= 1; l--) {
                    ge += sh.N[l][tid];
                    dl += ge;
                    Dl[l] = dl;
                }
                float mhl = 0.f;
#pragma unroll
                for (int l = 1; l <= MS_L; l++) {  // ascending l, sequential f32 adds (canonical order)
                    const uint32_t cnt = sh.S[l][tid];
                    if (cnt) mhl = __fadd_rn(mhl, __fdiv_rn(__fmul_rn((float)l, (float)cnt), (float)Dl[l]));
                }
                // mhl.rs:46-48 sums 1..max_n in f32: every partial sum is an integer < 2^24, so the closed form is exact
                best = __fdiv_rn(mhl, (float)(maxn * (maxn + 1) / 2));  // mhl.rs:71
                have = true;
            }
#pragma unroll
            for (int l = 1; l <= MS_L; l++) { sh.S[l][tid] = 0; sh.N[l][tid] = 0; }
            depth = maxn = 0;
        };
        for (int r = lo; r < nreads; r++) {
            const unsigned long long w = sh.rec[r];
            if (((uint32_t)w & 0xFFFFu) > key_hi) break;
            const uint32_t m = (uint32_t)(w >> 48), n = m >> 8;
            if (n == 0) continue;
            if (((uint32_t)(w >> 16) & 0xFFFFu) > pq) { close(); continue; }  // mhl.rs:162-173: ANY read with >= 1 CpG flushes what lies before its first CpG
            const uint16_t* cp = sh.pos + ((uint32_t)(w >> 32) & 0xFFFFu);
            // A read with more calls than the accumulators hold (a CpG island) whose calls reach p: THIS site goes to the per-site
            // kernel — not the whole tile, whose other sites mostly lie outside the island.  (Decided from the read's last call,
            // without walking its calls: an island read would cost every site of its window a scan of up to 255 calls.)
            if (n > (uint32_t)MS_L) {
                if ((uint32_t)cp[n - 1] >= pq) { deep = true; break; }
                continue;
            }
            // does the read call p?  (its calls are sorted; n <= MS_L)
            bool calls = false;
            for (uint32_t k = 0; k < n; k++) {
                const uint32_t x = cp[k];
                if (x >= pq) { calls = x == pq; break; }
            }
            if (!calls) continue;
            if ((m & 0xFFu) < prm.min_qual || n < prm.min_cpgs) continue;  // mhl.rs:176, :181
            if (++depth >= 4000u) deep = true;  // 16 calls x 4000 reads still fit the 16-bit accumulators
            maxn = max(maxn, n);
            sh.N[n][tid]++;  // mhl.rs:75-80
            unsigned long long x = (unsigned long long)sh.mbits[r] & low_mask64(n);
            for (uint32_t l = 1; x; l++) {  // stretch_info[l] = popc(x_l), readutil.rs:147-164
                sh.S[l][tid] += (uint32_t)__popcll(x);
                x &= x >> 1;
            }
        }
        close();
        if (deep) { fallback[s0 + tid] = 1; continue; }  // deeper than the 16-bit accumulators allow: per-site kernel
        value[s0 + tid] = best;
        rowcnt[s0 + tid] = have ? 1u : 0u;
    }
}

int launch_mhl_site(const ReadsView& rv, const int32_t* site_pos, int64_t C, const RegionScalars* sc, mth_mhl_params prm,
                    float* value, uint32_t* rowcnt, uint8_t* fallback, cudaStream_t s) {
    if (C <= 0) return 0;
    using Dense = MsSmem<MS_RCAP, MS_CCAP>;
    using Sparse = MsSmem<MS_RCAP_SPARSE, MS_CCAP_SPARSE>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_mhl_site<MS_RCAP, MS_CCAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Dense));
        cudaFuncSetAttribute(k_mhl_site<MS_RCAP_SPARSE, MS_CCAP_SPARSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Sparse));
        attr_set = true;
    }
    int64_t tiles = (C + MS_SITES - 1) / MS_SITES;
    if (tiles > 148 * 64) tiles = 148 * 64;
    // reads a tile has to stage ~ MS_SITES x reads per site gap: pick the instance whose capacity covers it with some room
    const double est = (double)MS_SITES * (double)rv.R / (double)C;
    static int force = -1;  // METHEOR_MHL_TILE = dense | sparse: kernel-variant experiments (profiles/)
    if (force < 0) {
        const char* e = getenv("METHEOR_MHL_TILE");
        force = e ? (!strcmp(e, "dense") ? 1 : !strcmp(e, "sparse") ? 2 : 0) : 0;
    }
    if (force == 1 || (force == 0 && est * 1.3 <= (double)MS_RCAP))
        k_mhl_site<MS_RCAP, MS_CCAP><<<(unsigned)tiles, MS_SITES, sizeof(Dense), s>>>(rv, site_pos, C, sc, prm, value, rowcnt, fallback);
    else
        k_mhl_site<MS_RCAP_SPARSE, MS_CCAP_SPARSE><<<(unsigned)tiles, MS_SITES, sizeof(Sparse), s>>>(rv, site_pos, C, sc, prm, value, rowcnt, fallback);
    return 1;
}

}  // namespace mth
