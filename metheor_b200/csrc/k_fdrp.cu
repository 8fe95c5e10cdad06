// k_fdrp.cu — FDRP (fdrp.rs:12-145,176-246) and qFDRP (qfdrp.rs:97-157,188-258): per-CpG read pile (<= max_depth
// reads, reservoir-sampled beyond) and all-pairs comparison.
//
// The reference keeps, per pile read, a 403-byte window centred on the CpG (bit0 covered, bit1 CpG call, bit2
// methylated) and compares pairs byte by byte.  Here the warp that owns site p
//   1. collects the pile (read indices in file order; window rule fdrp.rs:58-63 drops reads reaching beyond
//      +-201 bp; once full, slot j-1 is replaced when the seeded draw j <= max_depth — fdrp.rs:81-94 with the
//      documented counter-based draw instead of thread_rng) while replaying flush segments (gather.cuh);
//   2. when a segment closes with depth >= min_depth, ranks the union of the pile's CpG positions (a 416-bit
//      position bitmap in shared memory + popcount prefix) and stores for every pile read three rank-space
//      bit masks: cpg (bit1), meth (bit2) and covered-cpg (bit0 & bit1);
//   3. evaluates the n(n-1)/2 pairs in lexicographic order 32 at a time:
//        overlap = max(0, min(e_i,e_j) - max(s_i,s_j) + 1)              (fdrp.rs:97-107; bit0 is start..=end)
//        ham     = popc(cpg_i & cpg_j & cov_i & cov_j & (meth_i ^ meth_j)) (fdrp.rs:109-122, qfdrp.rs:121-135)
//        shared  = popc(cpg_i & cpg_j)                                   (qfdrp.rs:109-119: bit0 NOT required)
//      FDRP adds 1 per pair with overlap >= min_overlap and ham > 0; qFDRP adds ham/shared — sequentially, in
//      pair order, in f32 (qfdrp.rs:152), which is replayed with warp shuffles so the rounding is identical;
//      both divide by (n*(n-1)) as f32 / 2.0.
#include "gather.cuh"
#include "kernels.h"

namespace mth {

constexpr int FDRP_MAX_READ_LEN = 201;          // fdrp.rs:10
constexpr int FDRP_WIN_LO = FDRP_MAX_READ_LEN + 1;  // positions p-202 .. p+213 -> 416 bits (a reverse-strand CpG may sit at start-1)
constexpr int FDRP_UWORDS = 13;                 // 13 x 32 bits
constexpr int FDRP_MAXW = 7;                    // <= 416 distinct positions -> <= 7 rank words
constexpr int FDRP_WARPS = GATHER_BLOCK / 32;

struct FdrpScratch {  // per-warp slices of one big allocation (global memory, L1/L2 resident)
    uint32_t* pile;          // [D] read indices
    int32_t* st;             // [D]
    int32_t* en;             // [D]
    unsigned long long* cm;  // [MAXW][D]
    unsigned long long* mm;  // [MAXW][D]
    unsigned long long* vm;  // [MAXW][D]
};

__host__ __device__ inline size_t fdrp_per_warp_bytes(uint32_t D) {
    size_t d = ((size_t)D + 3) & ~(size_t)3;
    return d * 4 * 3 + d * 8 * 3 * FDRP_MAXW;
}

struct FdrpPolicy {
    static constexpr int SLACK = 0;  // fdrp.rs:214 strict cpg < first
    const ReadsView& rv;
    mth_fdrp_params prm;
    int quant;  // 0 FDRP, 1 qFDRP, 2 both from one pair loop (value = FDRP, value_q = qFDRP)
    uint64_t seed;
    ContigTable ct;
    float* value;
    uint32_t* rowcnt;
    float* value_q;
    uint32_t* rowcnt_q;
    FdrpScratch sc;
    uint32_t D;        // max_depth
    uint32_t* U;       // shared: union bitmap [13]
    uint32_t* UP;      // shared: prefix [13]
    int32_t p;
    uint32_t total, depth;
    float best, best_q;
    bool have;
    unsigned long long pair_ops = 0;

    __device__ FdrpPolicy(const ReadsView& rv_, mth_fdrp_params prm_, int q, uint64_t seed_, ContigTable ct_, float* v,
                          uint32_t* rc, float* vq, uint32_t* rcq, FdrpScratch sc_, uint32_t* U_, uint32_t* UP_)
        : rv(rv_), prm(prm_), quant(q), seed(seed_), ct(ct_), value(v), rowcnt(rc), value_q(vq), rowcnt_q(rcq), sc(sc_),
          D(prm_.max_depth), U(U_), UP(UP_) {}
    // fdrp.rs:205 mapq, :208 no CpGs
    __device__ __forceinline__ bool contrib_ok(uint32_t mapq, uint32_t n) const { return mapq >= prm.min_qual && n > 0; }
    __device__ __forceinline__ bool trigger_ok(uint32_t mapq, uint32_t n) const { return mapq >= prm.min_qual && n > 0; }
    __device__ __forceinline__ void begin_site(int32_t p_) { p = p_; total = depth = 0; have = false; best = best_q = 0.f; }

    __device__ __forceinline__ void add(uint32_t mask, const LaneRead& lr) {
        const int lane = lane_id();
        bool mine = (mask >> lane) & 1u;
        int32_t e = mine ? rv.end[lr.j] : 0;
        // fdrp.rs:55-63: rel(start) < 0 or rel(end) > 402 -> the read is not added (no depth increment)
        bool acc = mine && !(lr.start < p - FDRP_MAX_READ_LEN) && !(e > p + FDRP_MAX_READ_LEN);
        uint32_t am = __ballot_sync(FULL, acc);
        if (!am) return;
        uint32_t rank = __popc(am & ((1u << lane) - 1u));
        uint32_t t0 = total + rank;  // number of reads offered before this one
        int slot = -1;
        if (acc) {
            if (t0 < D) {
                slot = (int)t0;  // fdrp.rs:81-85
            } else {               // fdrp.rs:87-94
                int32_t tid, pos;
                delinearize(ct, p, &tid, &pos);
                uint32_t j = reservoir_draw(seed, tid, pos, t0 + 1);
                if (j <= D) slot = (int)j - 1;
            }
        }
        // several replacements may hit one slot within a chunk: the latest read (highest lane) must win
        bool writer = slot >= 0;
        uint32_t wm = __ballot_sync(FULL, writer);
        if (writer) {
            uint32_t same = __match_any_sync(wm, slot);
            if (lane == 31 - __clz(same)) sc.pile[slot] = (uint32_t)lr.j;
        }
        total += __popc(am);
        depth = min(total, D);
        __syncwarp();
    }

    __device__ void evaluate() {
        const int lane = lane_id();
        const bool want_q = quant != 0, want_d = quant != 1;
        const uint32_t n = depth;
        const size_t Dp = ((size_t)D + 3) & ~(size_t)3;
        // ---- union of CpG positions of the pile, in window coordinates ----
        if (lane < FDRP_UWORDS) U[lane] = 0;
        __syncwarp();
        for (uint32_t e = lane; e < n; e += 32) {
            int64_t j = sc.pile[e];
            uint32_t o0 = rv.cpg_off[j], nc = rv.cpg_off[j + 1] - o0;
            sc.st[e] = rv.start[j];
            sc.en[e] = rv.end[j];
            for (uint32_t k = 0; k < nc; k++) {
                uint32_t bit = (uint32_t)(rv.cpg_pos[o0 + k] - (p - FDRP_WIN_LO));
                atomicOr(&U[bit >> 5], 1u << (bit & 31));
            }
        }
        __syncwarp();
        if (lane < FDRP_UWORDS) {
            uint32_t pre = 0;
            for (int w = 0; w < lane; w++) pre += __popc(U[w]);
            UP[lane] = pre;
        }
        __syncwarp();
        const uint32_t K = UP[FDRP_UWORDS - 1] + __popc(U[FDRP_UWORDS - 1]);
        const uint32_t NW = (K + 63) >> 6;
        // ---- rank-space masks per pile read ----
        for (uint32_t e = lane; e < n; e += 32) {
            int64_t j = sc.pile[e];
            uint32_t o0 = rv.cpg_off[j], nc = rv.cpg_off[j + 1] - o0;
            int32_t s = sc.st[e], en = sc.en[e];
            for (uint32_t w = 0; w < NW; w++) { sc.cm[w * Dp + e] = 0; sc.mm[w * Dp + e] = 0; sc.vm[w * Dp + e] = 0; }
            for (uint32_t k = 0; k < nc; k++) {
                int32_t x = rv.cpg_pos[o0 + k];
                uint32_t bit = (uint32_t)(x - (p - FDRP_WIN_LO));
                uint32_t r = UP[bit >> 5] + __popc(U[bit >> 5] & ((1u << (bit & 31)) - 1u));
                unsigned long long m = 1ull << (r & 63);
                size_t at = (size_t)(r >> 6) * Dp + e;
                sc.cm[at] |= m;                                   // bit1, fdrp.rs:72
                if (meth_bit(rv, j, k)) sc.mm[at] |= m;           // bit2, fdrp.rs:74-76
                if (x >= s && x <= en) sc.vm[at] |= m;            // bit0 & bit1 (fdrp.rs:65-67 marks start..=end)
            }
        }
        __syncwarp();
        // ---- all pairs (i<j) in lexicographic order ----
        const uint64_t P = (uint64_t)n * (n - 1) / 2;
        pair_ops += P;
        uint32_t i = 0, jj = 1 + lane;  // pair index t = lane maps to (0, 1+lane) before normalisation
        while (i < n && jj >= n) { jj = jj - n + i + 2; i++; }  // row i holds n-1-i pairs
        float acc = 0.f;
        uint32_t disc = 0;
        for (uint64_t t0 = 0; t0 < P; t0 += 32) {
            bool valid = (t0 + lane) < P;
            float term = 0.f;
            if (valid) {
                int32_t ov = min(sc.en[i], sc.en[jj]) - max(sc.st[i], sc.st[jj]) + 1;
                if (ov < 0) ov = 0;
                if (ov >= prm.min_overlap) {  // fdrp.rs:133-136
                    // word 0 outside the loop: a pile's CpG positions almost always fit 64 ranks (NW == 1)
                    unsigned long long both = sc.cm[i] & sc.cm[jj];
                    uint32_t shared = __popcll(both);
                    uint32_t ham = __popcll(both & sc.vm[i] & sc.vm[jj] & (sc.mm[i] ^ sc.mm[jj]));
                    for (uint32_t w = 1; w < NW; w++) {
                        both = sc.cm[w * Dp + i] & sc.cm[w * Dp + jj];
                        shared += __popcll(both);
                        ham += __popcll(both & sc.vm[w * Dp + i] & sc.vm[w * Dp + jj] & (sc.mm[w * Dp + i] ^ sc.mm[w * Dp + jj]));
                    }
                    if (want_q && ham) term = __fdiv_rn((float)ham, (float)shared);  // qfdrp.rs:152 (0 / shared adds nothing)
                    if (want_d) disc += (ham > 0) ? 1u : 0u;                         // fdrp.rs:138-140
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
            // advance this lane's pair by 32
            if (valid) {
                jj += 32;
                while (i < n && jj >= n) { jj = jj - n + i + 2; i++; }
            }
        }
        const float den = __fdiv_rn((float)((unsigned long long)n * (unsigned long long)(n - 1)), 2.0f);  // fdrp.rs:143
        if (want_d) best = __fdiv_rn((float)__reduce_add_sync(FULL, disc), den);  // += 1.0 per discordant pair: exact in f32 below 2^24
        if (want_q) {
            const float v = __fdiv_rn(acc, den);
            if (quant == 2) best_q = v; else best = v;
        }
    }

    __device__ __forceinline__ void close() {
        if (depth > 0 && depth >= prm.min_depth) {  // fdrp.rs:215
            evaluate();
            have = true;
        }
        total = depth = 0;
    }
    __device__ __forceinline__ void end_site(int64_t s) {
        if (lane_id() == 0) {
            value[s] = best;
            rowcnt[s] = have ? 1u : 0u;
            if (quant == 2) {
                value_q[s] = best_q;
                rowcnt_q[s] = have ? 1u : 0u;
            }
        }
    }
};

__global__ void __launch_bounds__(GATHER_BLOCK, 4) k_fdrp(ReadsView rv, const int32_t* __restrict__ site_pos, int64_t C,
                                                       RegionScalars* __restrict__ scal, mth_fdrp_params prm, int quant,
                                                       uint64_t seed, ContigTable ct, char* scratch, float* __restrict__ value,
                                                       uint32_t* __restrict__ rowcnt, float* __restrict__ value_q,
                                                       uint32_t* __restrict__ rowcnt_q, const uint8_t* __restrict__ only) {
    __shared__ uint32_t sU[FDRP_WARPS][FDRP_UWORDS + 1];
    __shared__ uint32_t sP[FDRP_WARPS][FDRP_UWORDS + 1];
    const int warp = threadIdx.x >> 5;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t D = prm.max_depth;
    const size_t Dp = ((size_t)D + 3) & ~(size_t)3;
    char* base = scratch + (size_t)warp_global * fdrp_per_warp_bytes(D);
    FdrpScratch fs;
    fs.pile = (uint32_t*)base;
    fs.st = (int32_t*)(base + Dp * 4);
    fs.en = (int32_t*)(base + Dp * 8);
    fs.cm = (unsigned long long*)(base + Dp * 12);
    fs.mm = fs.cm + Dp * FDRP_MAXW;
    fs.vm = fs.mm + Dp * FDRP_MAXW;
    FdrpPolicy pol(rv, prm, quant, seed, ct, value, rowcnt, value_q, rowcnt_q, fs, sU[warp], sP[warp]);
    gather_sites(rv, site_pos, C, scal->lmax, pol, only);
    if (lane_id() == 0 && pol.pair_ops) atomicAdd(&scal->fdrp_pairs, pol.pair_ops);
}

// grid is bounded so that the per-warp scratch stays below ~2 GB even for very large max_depth
static int fdrp_grid(int64_t C, uint32_t D) {
    int g = gather_grid(C);
    size_t per_block = fdrp_per_warp_bytes(D) * FDRP_WARPS;
    size_t cap = ((size_t)2 << 30) / per_block;
    if (cap < 148) cap = 148;
    if ((size_t)g > cap) g = (int)cap;
    return g;
}

size_t fdrp_scratch_bytes(mth_fdrp_params prm, int) {
    // sized for the largest grid fdrp_grid can return
    size_t per_block = fdrp_per_warp_bytes(prm.max_depth) * FDRP_WARPS;
    size_t cap = ((size_t)2 << 30) / per_block;
    if (cap < 148) cap = 148;
    size_t blocks = cap < (size_t)(148 * 64) ? cap : (size_t)(148 * 64);
    return blocks * per_block;
}

int launch_fdrp(const ReadsView& rv, const int32_t* site_pos, int64_t C, RegionScalars* sc, mth_fdrp_params prm,
                int quantitative, uint64_t seed, ContigTable ct, void* scratch, size_t scratch_bytes, float* value,
                uint32_t* rowcnt, float* value_q, uint32_t* rowcnt_q, const uint8_t* only, cudaStream_t s) {
    if (C <= 0) return 0;
    int g = fdrp_grid(C, prm.max_depth);
    if ((size_t)g * fdrp_per_warp_bytes(prm.max_depth) * FDRP_WARPS > scratch_bytes) return 0;
    k_fdrp<<<g, GATHER_BLOCK, 0, s>>>(rv, site_pos, C, sc, prm, quantitative, seed, ct, (char*)scratch, value, rowcnt, value_q, rowcnt_q, only);
    return 1;
}

}  // namespace mth
