// k_expand.cu — compact wire format (mth_batch_compact, include/metheor_b200.h) -> the region's SoA arena.
//
// Three launches per batch: per-block call totals of the 1-byte per-read call counts, a single-block scan of the block
// totals (k_sites.cu k_scan_sums, shared), and the expansion proper: every thread owns one read, recomputes its block's
// exclusive scan, and writes start / end / meta / cpg_off (already shifted to the region's linear coordinate and call
// base, so no fix-up kernels follow), then its calls: position = start - 1 + delta, query index (implied for plain
// `<len>M` alignments, from rel_exc otherwise) and the packed methylation word gathered from the bit stream.
#include "kernels.h"

namespace mth {

constexpr int EXP_BLOCK = 256;

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* s_warp, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < EXP_BLOCK / 32; w++) {
        uint32_t x = s_warp[w];
        base += (w < warp) ? x : 0u;
        tot += x;
    }
    *total = tot;
    __syncthreads();
    return base + inc - v;
}

// block_sums[2*b] = calls of block b, block_sums[2*b+1] = explicit query indices of block b
__global__ void __launch_bounds__(EXP_BLOCK) k_expand_count(ExpandArgs a) {
    __shared__ uint32_t s_warp[EXP_BLOCK / 32];
    const int64_t r = (int64_t)blockIdx.x * EXP_BLOCK + threadIdx.x;
    uint32_t n = 0, ne = 0;
    if (r < a.n) {
        n = a.n_cpg8[r];
        if (a.rel_out && (a.flags[r] & 4u)) ne = n;
    }
    uint32_t tot, tot_e;
    block_excl_scan(n, s_warp, &tot);
    block_excl_scan(ne, s_warp, &tot_e);
    if (threadIdx.x == 0) {
        a.block_calls[blockIdx.x] = tot;
        a.block_rel[blockIdx.x] = tot_e;
    }
}

__global__ void __launch_bounds__(EXP_BLOCK) k_expand(ExpandArgs a) {
    __shared__ uint32_t s_warp[EXP_BLOCK / 32];
    const int64_t r = (int64_t)blockIdx.x * EXP_BLOCK + threadIdx.x;
    const bool in = r < a.n;
    uint32_t n = 0, ne = 0, fl = 0;
    if (in) {
        n = a.n_cpg8[r];
        fl = a.flags[r];
        if (a.rel_out && (fl & 4u)) ne = n;
    }
    uint32_t tot;
    const uint32_t o_blk = block_excl_scan(n, s_warp, &tot);                 // call offset within this block of 256 reads
    const uint32_t o_local = o_blk + a.block_calls[blockIdx.x];              // ... within the batch (block_calls: scanned in place)
    const uint32_t e_local = block_excl_scan(ne, s_warp, &tot) + a.block_rel[blockIdx.x];
    if (!in) return;
    const int64_t j = a.r0 + r;
    int32_t s;
    if (a.enc & MTH_CENC_START16) {
        const int32_t bs = a.blk_start[blockIdx.x];
        const int64_t xi = bs >= 0 ? 0 : (int64_t)(-(bs + 1)) * EXP_BLOCK + threadIdx.x;
        if (xi >= a.n_start_exc && bs < 0) {
            atomicOr(a.err, ERRBIT_BAD_OFFSETS);
            s = 0;
        } else {
            s = bs >= 0 ? bs + (int32_t)a.start_off16[r] : a.start_exc[xi];
        }
    } else {
        s = a.start[r];
    }
    const int32_t s_lin = s + a.lin_off;
    a.start_out[j] = s_lin;
    a.end_out[j] = s_lin + (int32_t)a.span[r];
    a.meta_out[j] = (uint32_t)a.mapq[r] | ((fl & 1u) << 8) | ((fl & 2u) << 8);  // forward -> bit 8, halo -> bit 9
    a.off_out[j] = (uint32_t)(a.i0 + o_local);
    if (r == a.n - 1) a.off_out[j + 1] = (uint32_t)(a.i0 + o_local + n);
    if (n > 64) {  // not representable here: leave the calls out, k_ingest reports ERRBIT_TOO_MANY_CPGS via the offsets
        atomicOr(a.err, ERRBIT_TOO_MANY_CPGS);
        n = 64;
    }
    if ((int64_t)o_local + n > a.n_calls) {  // per-read counts exceed the declared number of calls: corrupt batch
        atomicOr(a.err, ERRBIT_BAD_OFFSETS);
        n = 0;
    }
    uint64_t mw = 0;
    const int32_t fwd = (int32_t)(fl & 1u);
    // where this read's call deltas live: 16-bit offsets from start - 1, or (dense) 8/16-bit deltas from the previous call
    const bool chained = (a.enc & MTH_CENC_DELTA8) != 0;
    const uint32_t bco = chained ? a.blk_call_off[blockIdx.x] : 0u;
    const bool wide = !chained || (bco & 0x80000000u);
    const size_t dbase = chained ? (size_t)(bco & 0x7FFFFFFFu) + o_blk : (size_t)o_local;
    if ((chained && (int64_t)(dbase + n) > (wide ? a.n_delta16 : a.n_delta8)) ||
        (a.rel_out && (fl & 4u) && (int64_t)e_local + n > a.n_rel)) {
        atomicOr(a.err, ERRBIT_BAD_OFFSETS);
        n = 0;
    }
    int32_t d_acc = 0;
    for (uint32_t k = 0; k < n; k++) {
        const uint32_t x = o_local + k;
        int32_t d = wide ? (int32_t)a.cpg_delta[dbase + k] : (int32_t)a.cpg_delta8[dbase + k];
        if (chained) { d_acc += d; d = d_acc; }
        a.pos_out[a.i0 + x] = s_lin - 1 + d;
        const uint32_t xb = x + a.bit_base;
        mw |= (uint64_t)((a.meth_bits[xb >> 3] >> (xb & 7)) & 1u) << k;
        if (a.rel_out) a.rel_out[a.i0 + x] = (fl & 4u) ? a.rel_exc[e_local + k] : (uint16_t)(d - fwd);
    }
    a.meth_out[a.w0 + r] = mw;
    if (a.moff_out) {
        a.moff_out[j] = (uint32_t)(a.w0 + r);
        if (r == a.n - 1) a.moff_out[j + 1] = (uint32_t)(a.w0 + r + 1);
    }
}

int launch_expand(const ExpandArgs& a, unsigned long long* total_scratch, cudaStream_t s) {
    if (a.n <= 0) return 0;
    const unsigned nb = (unsigned)((a.n + EXP_BLOCK - 1) / EXP_BLOCK);
    k_expand_count<<<nb, EXP_BLOCK, 0, s>>>(a);
    int k = 1;
    k += launch_scan_sums(a.block_calls, nb, total_scratch, s);
    k += launch_scan_sums(a.block_rel, nb, total_scratch + 1, s);
    k_expand<<<nb, EXP_BLOCK, 0, s>>>(a);
    return k + 1;
}

}  // namespace mth
