// k_pdr.cu — PDR: per-CpG concordant/discordant read counts (pdr.rs:119-212, readutil.rs:134-145).
//
// Two kernels produce the same dense per-site counters cnt2[2*s] = n_concordant, cnt2[2*s+1] = n_discordant:
//   * k_pdr_scatter: one thread per read, int32 atomics into the site counters.  Exact whenever no read spans
//     more than 150 reference bases: a flush trigger must have its first CpG > p + 150 (pdr.rs:162) while
//     starting no later than a subsequent contributor of p (start <= p + 1), which needs a span >= 151.
//   * k_pdr_gather: the segment-exact gather (gather.cuh) for everything else.
// k_pdr_rowcnt / k_pdr_emit turn the counters into rows: pdr = d as f32 / (c as f32 + d as f32) (pdr.rs:47-49).
#include "gather.cuh"
#include "kernels.h"
#include "tma.cuh"

namespace mth {

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

__device__ __forceinline__ bool pdr_read_ok(const mth_pdr_params& prm, uint32_t mapq, uint32_t n) {
    // pdr.rs:147 (n < min_cpgs), :150 (mapq < min_qual), :155 (no CpGs)
    return n >= prm.min_cpgs && mapq >= prm.min_qual && n > 0;
}

// k_pdr_scatter is a pure per-CALL kernel: k_ingest already classified every read (filters pdr.rs:147-155 + concordance
// state readutil.rs:134-145) and replicated the verdict on each of its calls (call_flags, CF_PDR_C / CF_PDR_D), so the
// scatter reads 5 bytes per call (position + flag byte) and no per-read data at all.
// One CTA = PCS_TILE consecutive calls (vectorised int4 / uchar4 loads).  Reads are sorted, so the calls of a tile lie
// in [P0 - lmax, ...) where P0 is the tile's first call (a later read starts at or after the read owning P0):
//   * the 16 384-position window of the site bitmap starting there is loaded into shared memory and block-scanned, so
//     rank(p) = rank0 + prefix[w] + popc(word & below) needs shared memory only;
//   * counts go to shared-memory counters [local rank][state]; non-zero counters are flushed with one fire-and-forget
//     global atomic each (one per (tile, site, state) instead of one per call);
//   * calls outside the window / rank budget take the direct path (dictionary lookup + global atomic).
#ifndef PCS_CPT
#define PCS_CPT 8    // calls per thread, multiple of 4
#endif
#ifndef PCS_MINB
#define PCS_MINB 6
#endif
constexpr int PCS_THREADS = 256;
constexpr int PCS_TILE = PCS_THREADS * PCS_CPT;
constexpr int PCS_WIN = 256;                   // 64-bit bitmap words in the window = 16 384 positions
constexpr int PCS_RANKS = 64 * PCS_CPT;        // local site ranks with a shared-memory counter pair

__device__ __forceinline__ uint32_t site_rank(const unsigned long long* __restrict__ bitmap,
                                              const uint32_t* __restrict__ word_prefix, int32_t pos) {
    uint32_t bit = (uint32_t)(pos + 1);
    uint32_t w = bit >> 6;
    return __ldg(word_prefix + w) + (uint32_t)__popcll(__ldg(bitmap + w) & ((1ull << (bit & 63)) - 1ull));
}

__global__ void __launch_bounds__(PCS_THREADS, PCS_MINB) k_pdr_scatter(const int32_t* __restrict__ cpg_pos,
                                                                       const uint8_t* __restrict__ call_flags, int64_t n_calls,
                                                                       const unsigned long long* __restrict__ bitmap, int64_t n_words,
                                                                       const uint32_t* __restrict__ word_prefix,
                                                                       const RegionScalars* __restrict__ sc,
                                                                       uint32_t* __restrict__ cnt2) {
    __shared__ unsigned long long s_bmw[PCS_WIN];
    __shared__ uint32_t s_pref[PCS_WIN];
    __shared__ __align__(16) uint32_t s_cnt[2 * PCS_RANKS];
    __shared__ uint32_t s_wsum[PCS_THREADS / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t x0 = (int64_t)blockIdx.x * PCS_TILE;

    // this thread's calls: groups of 4 consecutive calls, groups interleaved across the CTA (coalesced 16 B / 4 B loads)
    int4 pv[PCS_CPT / 4];
    uchar4 fv[PCS_CPT / 4];
#pragma unroll
    for (int g = 0; g < PCS_CPT / 4; g++) {
        const int64_t x = x0 + (int64_t)(g * PCS_THREADS + tid) * 4;
        pv[g] = make_int4(0, 0, 0, 0);
        fv[g] = make_uchar4(0, 0, 0, 0);
        if (x + 4 <= n_calls) {
            pv[g] = *reinterpret_cast<const int4*>(cpg_pos + x);
            fv[g] = *reinterpret_cast<const uchar4*>(call_flags + x);
        } else if (x < n_calls) {
            int32_t tp[4] = {0, 0, 0, 0};
            uint8_t tf[4] = {0, 0, 0, 0};
            for (int k = 0; k < 4 && x + k < n_calls; k++) { tp[k] = cpg_pos[x + k]; tf[k] = call_flags[x + k]; }
            pv[g] = make_int4(tp[0], tp[1], tp[2], tp[3]);
            fv[g] = make_uchar4(tf[0], tf[1], tf[2], tf[3]);
        }
    }
    // window of the site bitmap: bit index of a call = position + 1 >= P0 - lmax + 1
    const int32_t wlo = max(cpg_pos[x0] - sc->lmax + 1, 0);
    const uint32_t w0 = (uint32_t)wlo >> 6;
    unsigned long long bw = 0;
    if ((int64_t)w0 + tid < n_words) bw = bitmap[w0 + tid];
    const uint32_t rank0 = word_prefix[w0];
    {
        uint4* z = reinterpret_cast<uint4*>(s_cnt);
        for (int k = tid; k < 2 * PCS_RANKS / 4; k += PCS_THREADS) z[k] = make_uint4(0, 0, 0, 0);
    }
    // local rank prefix of the window words (block exclusive scan of their popcounts)
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
        for (int w = 0; w < PCS_THREADS / 32; w++) base += (w < warp) ? s_wsum[w] : 0u;
        s_pref[tid] = base + inc - c;
    }
    __syncthreads();

    auto count = [&](int32_t p, uint32_t f) {
        const uint32_t cd = (f >> 3) & 3u;  // CF_PDR_C -> 1, CF_PDR_D -> 2
        if (!cd) return;
        const uint32_t bit = (uint32_t)(p + 1);
        const uint32_t w = (bit >> 6) - w0;
        if (w < (uint32_t)PCS_WIN) {
            const uint32_t r = s_pref[w] + (uint32_t)__popcll(s_bmw[w] & ((1ull << (bit & 63)) - 1ull));
            if (r < (uint32_t)PCS_RANKS) atomicAdd(&s_cnt[2 * r + (cd - 1)], 1u);
            else atomicAdd(&cnt2[2 * (size_t)(rank0 + r) + (cd - 1)], 1u);
        } else {
            atomicAdd(&cnt2[2 * (size_t)site_rank(bitmap, word_prefix, p) + (cd - 1)], 1u);
        }
    };
#pragma unroll
    for (int g = 0; g < PCS_CPT / 4; g++) {
        count(pv[g].x, fv[g].x);
        count(pv[g].y, fv[g].y);
        count(pv[g].z, fv[g].z);
        count(pv[g].w, fv[g].w);
    }
    __syncthreads();
    for (int k = tid; k < 2 * PCS_RANKS; k += PCS_THREADS) {
        uint32_t v = s_cnt[k];
        if (v) atomicAdd(&cnt2[2 * (size_t)rank0 + k], v);
    }
}

struct PdrGatherPolicy {
    static constexpr int SLACK = 150;  // pdr.rs:162 is_before(first, 150)
    const ReadsView& rv;
    mth_pdr_params prm;
    uint32_t* cnt2;
    uint32_t c, d, bc, bd;
    __device__ PdrGatherPolicy(const ReadsView& rv_, mth_pdr_params p, uint32_t* out) : rv(rv_), prm(p), cnt2(out) {}
    __device__ __forceinline__ bool contrib_ok(uint32_t mapq, uint32_t n) const { return pdr_read_ok(prm, mapq, n); }
    __device__ __forceinline__ bool trigger_ok(uint32_t mapq, uint32_t n) const { return pdr_read_ok(prm, mapq, n); }
    __device__ __forceinline__ void begin_site(int32_t) { c = d = bc = bd = 0; }
    __device__ __forceinline__ void add(uint32_t mask, const LaneRead& lr) {
        bool mine = (mask >> lane_id()) & 1u;
        bool disc = mine && read_discordant(rv, lr.j, lr.n);
        uint32_t dm = __ballot_sync(FULL, disc);
        d += __popc(dm);
        c += __popc(mask) - __popc(dm);
    }
    __device__ __forceinline__ void close() {
        if (c + d > 0 && c + d >= prm.min_depth) { bc = c; bd = d; }  // result.insert overwrites, pdr.rs:163-171
        c = d = 0;
    }
    __device__ __forceinline__ void end_site(int64_t s) {
        if (lane_id() == 0) { cnt2[2 * s] = bc; cnt2[2 * s + 1] = bd; }
    }
};

__global__ void __launch_bounds__(GATHER_BLOCK) k_pdr_gather(ReadsView rv, const int32_t* __restrict__ site_pos, int64_t C,
                                                             const RegionScalars* __restrict__ sc,
                                                             uint32_t* __restrict__ cnt2, mth_pdr_params prm,
                                                             const uint8_t* __restrict__ only) {
    PdrGatherPolicy pol(rv, prm, cnt2);
    gather_sites(rv, site_pos, C, sc->lmax, pol, only);
}

// Which sites can see a PDR flush at all?  A read j flushes the live CpG p (pdr.rs:160-177) iff p + 150 < first_cpg(j), and
// that only matters if a LATER read still contributes to p, i.e. starts at or before p + 1 — so start(j) <= p + 1 as well.
// Hence p lies in [start(j) - 1, first_cpg(j) - 151]: only reads whose first CpG call sits >= 150 bases behind their start
// (long reads / reads with deletions and a CpG-free head) create hazard sites, and only those few sites need the
// segment-exact gather kernel; everywhere else the scatter counts are provably what the reference reports.
__global__ void __launch_bounds__(256) k_pdr_hazard(ReadsView rv, const unsigned long long* __restrict__ bitmap,
                                                    const uint32_t* __restrict__ word_prefix, uint8_t* __restrict__ hazard) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= rv.R) return;
    const uint32_t o0 = rv.cpg_off[j];
    if (rv.cpg_off[j + 1] == o0) return;
    const int32_t s = rv.start[j], first = rv.cpg_pos[o0];
    if (first - s < 150) return;
    // ranks of the sites with position in [s - 1, first - 151]: site_rank(x) = number of sites below x
    const uint32_t r0 = site_rank(bitmap, word_prefix, s - 1), r1 = site_rank(bitmap, word_prefix, first - 150);
    for (uint32_t r = r0; r < r1; r++) hazard[r] = 1;
}

__global__ void k_pdr_rowcnt(const uint32_t* __restrict__ cnt2, int64_t C, uint32_t min_depth,
                             uint32_t* __restrict__ rowcnt) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= C) return;
    uint32_t t = cnt2[2 * s] + cnt2[2 * s + 1];
    rowcnt[s] = (t > 0 && t >= min_depth) ? 1u : 0u;
}

__global__ void k_pdr_emit(const uint32_t* __restrict__ cnt2, const uint32_t* __restrict__ rowoff,
                           const int32_t* __restrict__ site_pos, int64_t C, uint32_t min_depth, ContigTable ct,
                           SiteRowsDev rows, int64_t row_base) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= C) return;
    uint32_t c = cnt2[2 * s], d = cnt2[2 * s + 1];
    uint32_t t = c + d;
    if (!(t > 0 && t >= min_depth)) return;
    int64_t r = row_base + rowoff[s];
    int32_t tid, pos;
    delinearize(ct, site_pos[s], &tid, &pos);
    rows.tid[r] = tid;
    rows.pos[r] = pos;
    rows.value[r] = __fdiv_rn((float)d, __fadd_rn((float)c, (float)d));  // pdr.rs:47-49
    rows.n_conc[r] = c;
    rows.n_disc[r] = d;
}

int launch_pdr_scatter(const int32_t* cpg_pos, const uint8_t* call_flags, int64_t n_calls, const unsigned long long* bitmap,
                       int64_t n_words, const uint32_t* word_prefix, const RegionScalars* sc, uint32_t* cnt2, cudaStream_t s) {
    if (n_calls <= 0) return 0;
    k_pdr_scatter<<<grid_for(n_calls, PCS_TILE), PCS_THREADS, 0, s>>>(cpg_pos, call_flags, n_calls, bitmap, n_words, word_prefix, sc, cnt2);
    return 1;
}

int gather_grid(int64_t C) {
    int64_t warps = (C + SITES_PER_WARP - 1) / SITES_PER_WARP;
    int64_t blocks = (warps * 32 + GATHER_BLOCK - 1) / GATHER_BLOCK;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 64) blocks = 148 * 64;
    return (int)blocks;
}

int launch_pdr_gather(const ReadsView& rv, const int32_t* site_pos, int64_t C, const RegionScalars* sc,
                      uint32_t* cnt2, mth_pdr_params prm, const uint8_t* only, cudaStream_t s) {
    if (C <= 0) return 0;
    k_pdr_gather<<<gather_grid(C), GATHER_BLOCK, 0, s>>>(rv, site_pos, C, sc, cnt2, prm, only);
    return 1;
}

int launch_pdr_hazard(const ReadsView& rv, const unsigned long long* bitmap, const uint32_t* word_prefix, uint8_t* hazard,
                      cudaStream_t s) {
    if (rv.R <= 0) return 0;
    k_pdr_hazard<<<grid_for(rv.R, 256), 256, 0, s>>>(rv, bitmap, word_prefix, hazard);
    return 1;
}

int launch_pdr_rowcnt(const uint32_t* cnt2, int64_t C, uint32_t min_depth, uint32_t* rowcnt, cudaStream_t s) {
    if (C <= 0) return 0;
    k_pdr_rowcnt<<<grid_for(C, 256), 256, 0, s>>>(cnt2, C, min_depth, rowcnt);
    return 1;
}

int launch_pdr_emit(const uint32_t* cnt2, const uint32_t* rowoff, const int32_t* site_pos, int64_t C, uint32_t min_depth,
                    ContigTable ct, SiteRowsDev rows, int64_t row_base, cudaStream_t s) {
    if (C <= 0) return 0;
    k_pdr_emit<<<grid_for(C, 256), 256, 0, s>>>(cnt2, rowoff, site_pos, C, min_depth, ct, rows, row_base);
    return 1;
}

}  // namespace mth
