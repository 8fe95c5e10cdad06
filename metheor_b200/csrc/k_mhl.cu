// k_mhl.cu — methylation haplotype load per CpG (mhl.rs:135-208, AssociatedReads mhl.rs:12-81,
// stretch histogram readutil.rs:147-164).
//
// Per read, with x_1 = its methylation bits (bit k = k-th CpG of the read) and x_l = x_{l-1} & (x_{l-1} >> 1),
// popc(x_l) is exactly stretch_info[l] (number of fully methylated windows of l consecutive CpGs of the read).
// Per site the warp keeps, for the live segment, S[l] = sum of stretch_info[l] and N[n] = number of reads with n
// CpGs in per-warp shared memory; closing a segment computes
//     mhl = ( sum_{l: S[l]>0} (l as f32 * S[l] as f32) / D[l] ) / sum_{l=1..max_n} l ,   D[l] = sum_reads max(0, n-l+1)
// with sequential f32 adds in ascending l (the reference iterates a HashMap, mhl.rs:50, i.e. unspecified order;
// ascending l is the canonical order shared with the oracle) and no FMA contraction.
#include "gather.cuh"
#include "kernels.h"

namespace mth {

constexpr int MHL_LCAP = MAX_CPGS_PER_READ;  // 256
constexpr int MHL_WARPS = GATHER_BLOCK / 32;

struct MhlPolicy {
    static constexpr int SLACK = 0;  // mhl.rs:164 strict cpg < first
    const ReadsView& rv;
    mth_mhl_params prm;
    float* value;
    uint32_t* rowcnt;
    uint32_t* S;  // per-warp shared: S[1..LCAP]
    uint32_t* N;  // per-warp shared: N[1..LCAP]
    uint32_t depth, maxn, maxl;
    float best;
    bool have;

    __device__ MhlPolicy(const ReadsView& rv_, mth_mhl_params p, float* v, uint32_t* rc, uint32_t* s_, uint32_t* n_)
        : rv(rv_), prm(p), value(v), rowcnt(rc), S(s_), N(n_) {}
    // mhl.rs:176 (mapq), :181 (min_cpgs)
    __device__ __forceinline__ bool contrib_ok(uint32_t mapq, uint32_t n) const { return mapq >= prm.min_qual && n >= prm.min_cpgs; }
    // mhl.rs:162: ANY read with >= 1 CpG flushes, before the filters
    __device__ __forceinline__ bool trigger_ok(uint32_t, uint32_t) const { return true; }
    __device__ __forceinline__ void begin_site(int32_t) { depth = maxn = maxl = 0; have = false; best = 0.f; }

    __device__ __forceinline__ void add(uint32_t mask, const LaneRead& lr) {
        const int lane = lane_id();
        bool mine = (mask >> lane) & 1u;
        uint32_t n = mine ? lr.n : 0u;
        depth += __popc(mask);
        const uint32_t nmax = __reduce_max_sync(FULL, n);
        maxn = max(maxn, nmax);
        if (mine) atomicAdd(&N[n], 1u);  // mhl.rs:75-80
        if (nmax <= 64) {
            // every contributor of this chunk fits one word: x_1 = methylation bits, x_l = x_{l-1} & (x_{l-1} >> 1)
            uint64_t x = mine ? (meth_word(rv, lr.j, 0) & low_mask64(n)) : 0ull;
            for (uint32_t l = 1; l <= 64; l++) {
                uint32_t tot = __reduce_add_sync(FULL, (uint32_t)__popcll(x));
                if (tot == 0) break;
                if (lane == 0) S[l] += tot;  // mhl.rs:36-41
                maxl = max(maxl, l);
                x &= x >> 1;
            }
        } else {
            uint64_t x[MAX_METH_WORDS];
            uint32_t nw = (n + 63) >> 6;
#pragma unroll
            for (int w = 0; w < MAX_METH_WORDS; w++) {
                x[w] = 0;
                if ((uint32_t)w < nw) x[w] = meth_word(rv, lr.j, w) & low_mask64(min(64u, n - 64u * w));
            }
            for (uint32_t l = 1; l <= (uint32_t)MHL_LCAP; l++) {
                uint32_t h = 0;
#pragma unroll
                for (int w = 0; w < MAX_METH_WORDS; w++) h += __popcll(x[w]);
                uint32_t tot = __reduce_add_sync(FULL, h);
                if (tot == 0) break;
                if (lane == 0) S[l] += tot;
                maxl = max(maxl, l);
#pragma unroll
                for (int w = 0; w < MAX_METH_WORDS; w++) {  // x &= x >> 1 across words
                    uint64_t hi = (w + 1 < MAX_METH_WORDS) ? x[w + 1] : 0ull;
                    x[w] &= (x[w] >> 1) | (hi << 63);
                }
            }
        }
        __syncwarp();
    }

    __device__ __forceinline__ void close() {
        if (depth == 0) return;
        __syncwarp();
        const int lane = lane_id();
        if (depth >= prm.min_depth) {  // mhl.rs:165
            float res;
            if (maxn <= 32) {
                // lane l-1 owns stretch length l: D[l] = sum_{n >= l} (n - l + 1) N[n] is a suffix sum of the suffix counts
                const uint32_t l = (uint32_t)lane + 1;
                uint32_t ge = l <= maxn ? N[l] : 0u;   // -> #reads with n >= l
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t t = __shfl_down_sync(FULL, ge, o);
                    if (lane + o < 32) ge += t;
                }
                uint32_t dl = ge;                      // -> D[l]
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t t = __shfl_down_sync(FULL, dl, o);
                    if (lane + o < 32) dl += t;
                }
                const uint32_t cnt = l <= maxl ? S[l] : 0u;
                // mhl.rs:56-68: (l as f32 * count as f32) / denom ; terms with count 0 do not exist in the reference's map
                const float term = cnt ? __fdiv_rn(__fmul_rn((float)l, (float)cnt), (float)dl) : 0.f;
                float mhl = 0.f;
                for (uint32_t k = 0; k < maxl; k++) {  // ascending l, sequential f32 adds (canonical order)
                    const float t = __shfl_sync(FULL, term, (int)k);
                    if (__shfl_sync(FULL, cnt, (int)k)) mhl = __fadd_rn(mhl, t);
                }
                // mhl.rs:46-48 sums 1..max_n in f32: every partial sum is an integer < 2^24, so the closed form is exact
                const float l_sum = (float)(maxn * (maxn + 1) / 2);
                res = __fdiv_rn(mhl, l_sum);           // mhl.rs:71
            } else {
                res = 0.f;
                if (lane == 0) {
                    uint32_t cnt_ge = 0, dl = 0;
                    for (uint32_t l = maxn; l >= 1; l--) {  // D[l] for l = maxn..1, stored over N[l]
                        cnt_ge += N[l];
                        dl += cnt_ge;
                        N[l] = dl;
                    }
                    float mhl = 0.f, l_sum = 0.f;
                    for (uint32_t l = 1; l <= maxn; l++) l_sum = __fadd_rn(l_sum, (float)l);  // mhl.rs:46-48
                    for (uint32_t l = 1; l <= maxl; l++) {                                     // mhl.rs:50-69
                        uint32_t cnt = S[l];
                        if (cnt == 0) continue;
                        mhl = __fadd_rn(mhl, __fdiv_rn(__fmul_rn((float)l, (float)cnt), (float)N[l]));
                    }
                    res = __fdiv_rn(mhl, l_sum);
                }
                res = __shfl_sync(FULL, res, 0);
            }
            best = res;
            have = true;
        }
        __syncwarp();
        for (uint32_t l = 1 + lane; l <= max(maxn, maxl); l += 32) { S[l] = 0; N[l] = 0; }
        __syncwarp();
        depth = maxn = maxl = 0;
    }

    __device__ __forceinline__ void end_site(int64_t s) {
        if (lane_id() == 0) {
            value[s] = best;
            rowcnt[s] = have ? 1u : 0u;
        }
    }
};

__global__ void __launch_bounds__(GATHER_BLOCK) k_mhl(ReadsView rv, const int32_t* __restrict__ site_pos, int64_t C,
                                                      const RegionScalars* __restrict__ sc, mth_mhl_params prm,
                                                      float* __restrict__ value, uint32_t* __restrict__ rowcnt,
                                                      const uint8_t* __restrict__ only) {
    __shared__ uint32_t sS[MHL_WARPS][MHL_LCAP + 1];
    __shared__ uint32_t sN[MHL_WARPS][MHL_LCAP + 1];
    int warp = threadIdx.x >> 5;
    for (int l = lane_id(); l <= MHL_LCAP; l += 32) { sS[warp][l] = 0; sN[warp][l] = 0; }
    __syncwarp();
    MhlPolicy pol(rv, prm, value, rowcnt, sS[warp], sN[warp]);
    gather_sites(rv, site_pos, C, sc->lmax, pol, only);
}

__global__ void k_site_emit(const float* __restrict__ value, const uint32_t* __restrict__ rowoff, int64_t C,
                            unsigned long long n_rows_region, const int32_t* __restrict__ site_pos, ContigTable ct,
                            SiteRowsDev rows, int64_t row_base) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= C) return;
    // rowoff is the exclusive scan of the 0/1 row counts: site s has a row iff the next offset differs
    uint32_t o = rowoff[s];
    uint32_t nxt = (s + 1 < C) ? rowoff[s + 1] : (uint32_t)n_rows_region;
    if (nxt == o) return;
    int64_t r = row_base + o;
    int32_t tid, pos;
    delinearize(ct, site_pos[s], &tid, &pos);
    rows.tid[r] = tid;
    rows.pos[r] = pos;
    rows.value[r] = value[s];
}

int launch_mhl(const ReadsView& rv, const int32_t* site_pos, int64_t C, const RegionScalars* sc, mth_mhl_params prm,
               float* value, uint32_t* rowcnt, const uint8_t* only, cudaStream_t s) {
    if (C <= 0) return 0;
    k_mhl<<<gather_grid(C), GATHER_BLOCK, 0, s>>>(rv, site_pos, C, sc, prm, value, rowcnt, only);
    return 1;
}

int launch_site_emit(const float* value, const uint32_t* rowoff, uint64_t n_rows_region, const int32_t* site_pos,
                     int64_t C, ContigTable ct, SiteRowsDev rows, int64_t row_base, cudaStream_t s) {
    if (C <= 0) return 0;
    unsigned long long n_rows = n_rows_region;
    k_site_emit<<<(unsigned)((C + 255) / 256), 256, 0, s>>>(value, rowoff, C, n_rows, site_pos, ct, rows, row_base);
    return 1;
}

}  // namespace mth
