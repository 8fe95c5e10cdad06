// k_sites.cu — CpG-site dictionary (sorted unique positions of the region) from the site bitmap, plus the
// device-wide exclusive scan used to turn per-site row counts into output offsets.
//
// The dictionary replaces the reference's HashMap<CpGPosition, …> keys (pdr.rs:131, mhl.rs:147, fdrp.rs:193):
// rank(p) = word_prefix[(p+1)>>6] + popc(bitmap word below bit (p+1)&63); site_pos[rank] = p.
#include "kernels.h"

namespace mth {

constexpr int SCAN_BLOCK = 256;
constexpr int WORDS_PER_THREAD = 4;
constexpr int WORDS_PER_BLOCK = SCAN_BLOCK * WORDS_PER_THREAD;  // 1024 words = 65536 positions

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[SCAN_BLOCK / 32];
    __shared__ uint32_t block_total;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < SCAN_BLOCK / 32 ? warp_sums[lane] : 0;
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < SCAN_BLOCK / 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(FULL, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < SCAN_BLOCK / 32) warp_sums[lane] = winc - w;
        if (lane == SCAN_BLOCK / 32 - 1) block_total = winc;
    }
    __syncthreads();
    uint32_t res = inc - v + warp_sums[warp];
    *total = block_total;
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_sites_count(const unsigned long long* __restrict__ bitmap,
                                                            int64_t n_words, uint32_t* __restrict__ block_sums) {
    int64_t w0 = ((int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x) * WORDS_PER_THREAD;
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < WORDS_PER_THREAD; k++)
        if (w0 + k < n_words) c += __popcll(bitmap[w0 + k]);
    uint32_t total;
    block_exclusive_scan(c, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: in-place exclusive scan of sums[0..n), grand total -> *total.
// n is small (one entry per 65 536 positions / 2 048 sites) and this kernel sits on every region's critical path three or
// more times, so it is built for latency: 1024 threads, each owning a run of ceil(n / 1024) consecutive entries, ONE block-wide
// scan of the per-thread totals (two barriers) instead of a loop of 256-entry scans with a carry.
constexpr int SUMS_BLOCK = 1024;
__global__ void __launch_bounds__(SUMS_BLOCK) k_scan_sums(uint32_t* sums, int64_t n, unsigned long long* total) {
    __shared__ unsigned long long warp_tot[SUMS_BLOCK / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t per = (n + SUMS_BLOCK - 1) / SUMS_BLOCK;
    const int64_t i0 = (int64_t)threadIdx.x * per, i1 = min(n, i0 + per);
    unsigned long long mine = 0;
    for (int64_t i = i0; i < i1; i++) mine += sums[i];
    unsigned long long inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = warp_tot[lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(FULL, winc, o);
            if (lane >= o) winc += t;
        }
        warp_tot[lane] = winc - w;
        if (lane == 31) *total = winc;
    }
    __syncthreads();
    unsigned long long run = warp_tot[warp] + inc - mine;
    for (int64_t i = i0; i < i1; i++) {
        const uint32_t v = sums[i];
        sums[i] = (uint32_t)run;
        run += v;
    }
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_sites_emit(const unsigned long long* __restrict__ bitmap, int64_t n_words,
                                                           const uint32_t* __restrict__ block_offs,
                                                           uint32_t* __restrict__ word_prefix,
                                                           int32_t* __restrict__ site_pos) {
    int64_t w0 = ((int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x) * WORDS_PER_THREAD;
    unsigned long long x[WORDS_PER_THREAD];
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < WORDS_PER_THREAD; k++) {
        x[k] = (w0 + k < n_words) ? bitmap[w0 + k] : 0ull;
        c += __popcll(x[k]);
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(c, &total) + block_offs[blockIdx.x];
#pragma unroll
    for (int k = 0; k < WORDS_PER_THREAD; k++) {
        if (w0 + k < n_words) word_prefix[w0 + k] = ex;
        unsigned long long v = x[k];
        while (v) {
            int b = __ffsll((long long)v) - 1;
            site_pos[ex++] = (int32_t)((w0 + k) * 64 + b - 1);  // bit index = position + 1
            v &= v - 1;
        }
    }
}

int launch_scan_sums(uint32_t* sums, int64_t n, unsigned long long* total, cudaStream_t s) {
    k_scan_sums<<<1, SUMS_BLOCK, 0, s>>>(sums, n, total);
    return 1;
}

int launch_sites_count(const unsigned long long* bitmap, int64_t n_words, uint32_t* block_sums, RegionScalars* sc,
                       cudaStream_t s) {
    int64_t nb = (n_words + WORDS_PER_BLOCK - 1) / WORDS_PER_BLOCK;
    if (nb <= 0) nb = 1;
    k_sites_count<<<(unsigned)nb, SCAN_BLOCK, 0, s>>>(bitmap, n_words, block_sums);
    k_scan_sums<<<1, SUMS_BLOCK, 0, s>>>(block_sums, nb, &sc->n_sites);
    return 2;
}

int launch_sites_emit(const unsigned long long* bitmap, int64_t n_words, const uint32_t* block_sums,
                      uint32_t* word_prefix, int32_t* site_pos, cudaStream_t s) {
    int64_t nb = (n_words + WORDS_PER_BLOCK - 1) / WORDS_PER_BLOCK;
    if (nb <= 0) nb = 1;
    k_sites_emit<<<(unsigned)nb, SCAN_BLOCK, 0, s>>>(bitmap, n_words, block_sums, word_prefix, site_pos);
    return 1;
}

// ---- generic exclusive scan of u32 (row counts -> row offsets) ------------------------------
constexpr int ITEMS_PER_THREAD = 8;
constexpr int ITEMS_PER_BLOCK = SCAN_BLOCK * ITEMS_PER_THREAD;  // 2048

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_partial(const uint32_t* __restrict__ a, int64_t n,
                                                             uint32_t* __restrict__ sums) {
    int64_t i0 = ((int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x) * ITEMS_PER_THREAD;
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < ITEMS_PER_THREAD; k++)
        if (i0 + k < n) c += a[i0 + k];
    uint32_t total;
    block_exclusive_scan(c, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_final(uint32_t* __restrict__ a, int64_t n,
                                                           const uint32_t* __restrict__ offs) {
    int64_t i0 = ((int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x) * ITEMS_PER_THREAD;
    uint32_t v[ITEMS_PER_THREAD];
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < ITEMS_PER_THREAD; k++) {
        v[k] = (i0 + k < n) ? a[i0 + k] : 0;
        c += v[k];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(c, &total) + offs[blockIdx.x];
#pragma unroll
    for (int k = 0; k < ITEMS_PER_THREAD; k++) {
        if (i0 + k < n) a[i0 + k] = ex;
        ex += v[k];
    }
}

int launch_exclusive_scan_u32(uint32_t* a, int64_t n, uint32_t* scratch, unsigned long long* total, cudaStream_t s) {
    int64_t nb = (n + ITEMS_PER_BLOCK - 1) / ITEMS_PER_BLOCK;
    if (nb <= 0) nb = 1;
    k_scan_partial<<<(unsigned)nb, SCAN_BLOCK, 0, s>>>(a, n, scratch);
    k_scan_sums<<<1, SUMS_BLOCK, 0, s>>>(scratch, nb, total);
    k_scan_final<<<(unsigned)nb, SCAN_BLOCK, 0, s>>>(a, n, scratch);
    return 3;
}

__global__ void k_count_flags(const uint8_t* __restrict__ flags, int64_t n, unsigned long long* __restrict__ total) {
    unsigned int c = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) c += flags[i] ? 1u : 0u;
    c = __reduce_add_sync(FULL, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(total, (unsigned long long)c);
}
int launch_count_flags(const uint8_t* flags, int64_t n, unsigned long long* total, cudaStream_t s) {
    if (n <= 0) return 0;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_count_flags<<<(unsigned)blocks, 256, 0, s>>>(flags, n, total);
    return 1;
}

}  // namespace mth
