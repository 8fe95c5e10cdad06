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

namespace mth {

__device__ __forceinline__ bool pdr_read_ok(const mth_pdr_params& prm, uint32_t mapq, uint32_t n) {
    // pdr.rs:147 (n < min_cpgs), :150 (mapq < min_qual), :155 (no CpGs)
    return n >= prm.min_cpgs && mapq >= prm.min_qual && n > 0;
}

// One CTA = one tile of PDS_TILE consecutive reads (their CpG calls are one contiguous slice of cpg_pos).
//   phase 1, one thread per read : filters + concordance state -> code {0 skip, 1 concordant, 2 discordant} and the
//            owner table (call -> read of the tile) in shared memory;
//   phase 2, one thread per call : coalesced cpg_pos load, site rank from the dictionary, shared-memory atomics into a
//            window of PDS_WIN site ranks starting at the first site the tile can touch (reads are sorted, a 30x tile
//            of 256 reads covers ~1.3 kb = a few dozen sites), direct global atomics for ranks outside the window;
//   phase 3: non-zero window counters are flushed with one global atomic each.
// Global atomics drop from one per CpG call to one per (tile, site, state).
constexpr int PDS_TILE = 256;
constexpr int PDS_CAP = 4096;   // calls per tile with an owner entry; beyond that the per-read fallback runs
constexpr int PDS_WIN = 1024;   // site ranks aggregated in shared memory

__device__ __forceinline__ uint32_t site_rank(const unsigned long long* __restrict__ bitmap,
                                              const uint32_t* __restrict__ word_prefix, int32_t pos) {
    uint32_t bit = (uint32_t)(pos + 1);
    uint32_t w = bit >> 6;
    return __ldg(word_prefix + w) + (uint32_t)__popcll(__ldg(bitmap + w) & ((1ull << (bit & 63)) - 1ull));
}

__global__ void __launch_bounds__(PDS_TILE) k_pdr_scatter(ReadsView rv, const unsigned long long* __restrict__ bitmap,
                                                          const uint32_t* __restrict__ word_prefix,
                                                          uint32_t* __restrict__ cnt2, mth_pdr_params prm) {
    __shared__ uint32_t s_cnt[2 * PDS_WIN];
    __shared__ uint8_t s_owner[PDS_CAP];
    __shared__ uint8_t s_code[PDS_TILE];
    __shared__ uint32_t s_lo, s_hi, s_rank0;

    const int tid = threadIdx.x;
    const int64_t tile0 = (int64_t)blockIdx.x * PDS_TILE;
    const int64_t tile1 = min(rv.R, tile0 + PDS_TILE);
    if (tid == 0) {
        s_lo = rv.cpg_off[tile0];
        s_hi = rv.cpg_off[tile1];
        // every call of the tile lies at or after start[tile0] - 1 (reads sorted by start, calls in [start-1, end])
        int32_t pmin = rv.start[tile0] - 1;
        uint32_t bit = (uint32_t)(pmin + 1);
        uint32_t w = bit >> 6;
        s_rank0 = word_prefix[w] + (uint32_t)__popcll(bitmap[w] & ((1ull << (bit & 63)) - 1ull));
    }
    for (int k = tid; k < 2 * PDS_WIN; k += PDS_TILE) s_cnt[k] = 0;

    const int64_t j = tile0 + tid;
    uint32_t o0 = 0, n = 0, code = 0;
    if (j < tile1) {
        o0 = rv.cpg_off[j];
        n = rv.cpg_off[j + 1] - o0;
        uint32_t mapq = rv.meta[j] & 0xFFu;
        if (n > 0 && pdr_read_ok(prm, mapq, n)) code = read_discordant(rv, j, n) ? 2u : 1u;
    }
    s_code[tid] = (uint8_t)code;
    __syncthreads();
    const uint32_t lo = s_lo, hi = s_hi, rank0 = s_rank0;
    const bool tabled = hi - lo <= (uint32_t)PDS_CAP;
    if (tabled) {
        for (uint32_t k = 0; k < n; k++) s_owner[o0 - lo + k] = (uint8_t)tid;
        __syncthreads();
        for (uint32_t x = lo + tid; x < hi; x += PDS_TILE) {
            uint32_t cd = s_code[s_owner[x - lo]];
            if (!cd) continue;
            uint32_t r = site_rank(bitmap, word_prefix, rv.cpg_pos[x]);
            uint32_t rel = r - rank0;
            if (rel < (uint32_t)PDS_WIN) atomicAdd(&s_cnt[2 * rel + (cd - 1)], 1u);
            else atomicAdd(&cnt2[2 * (size_t)r + (cd - 1)], 1u);
        }
    } else if (code) {  // dense tile: one thread per read, straight to global memory
        for (uint32_t k = 0; k < n; k++) {
            uint32_t r = site_rank(bitmap, word_prefix, rv.cpg_pos[o0 + k]);
            atomicAdd(&cnt2[2 * (size_t)r + (code - 1)], 1u);
        }
    }
    __syncthreads();
    for (int k = tid; k < 2 * PDS_WIN; k += PDS_TILE) {
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
                                                             uint32_t* __restrict__ cnt2, mth_pdr_params prm) {
    PdrGatherPolicy pol(rv, prm, cnt2);
    gather_sites(rv, site_pos, C, sc->lmax, pol);
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

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

int launch_pdr_scatter(const ReadsView& rv, const unsigned long long* bitmap, const uint32_t* word_prefix,
                       uint32_t* cnt2, mth_pdr_params prm, cudaStream_t s) {
    if (rv.R <= 0) return 0;
    k_pdr_scatter<<<grid_for(rv.R, PDS_TILE), PDS_TILE, 0, s>>>(rv, bitmap, word_prefix, cnt2, prm);
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
                      uint32_t* cnt2, mth_pdr_params prm, cudaStream_t s) {
    if (C <= 0) return 0;
    k_pdr_gather<<<gather_grid(C), GATHER_BLOCK, 0, s>>>(rv, site_pos, C, sc, cnt2, prm);
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
