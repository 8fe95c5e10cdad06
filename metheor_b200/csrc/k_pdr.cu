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

__global__ void __launch_bounds__(256) k_pdr_scatter(ReadsView rv, const unsigned long long* __restrict__ bitmap,
                                                     const uint32_t* __restrict__ word_prefix,
                                                     uint32_t* __restrict__ cnt2, mth_pdr_params prm) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= rv.R) return;
    uint32_t o0 = rv.cpg_off[j], n = rv.cpg_off[j + 1] - o0;
    if (n == 0) return;
    uint32_t mapq = rv.meta[j] & 0xFFu;
    if (!pdr_read_ok(prm, mapq, n)) return;
    uint32_t disc = read_discordant(rv, j, n) ? 1u : 0u;
    for (uint32_t k = 0; k < n; k++) {
        uint32_t bit = (uint32_t)(rv.cpg_pos[o0 + k] + 1);
        uint32_t w = bit >> 6;
        uint32_t rank = word_prefix[w] + (uint32_t)__popcll(bitmap[w] & ((1ull << (bit & 63)) - 1ull));
        atomicAdd(&cnt2[2 * (size_t)rank + disc], 1u);
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
    k_pdr_scatter<<<grid_for(rv.R, 256), 256, 0, s>>>(rv, bitmap, word_prefix, cnt2, prm);
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
