// tma.cuh — 1-D TMA bulk copies (cp.async.bulk, SASS UBLKCP) global -> shared with mbarrier completion, and the
// tile-slice staging shared by the streaming kernels (k_ingest, k_pdr_scatter).
#pragma once
#include "common.cuh"

namespace mth {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

// The CpG calls of a tile of consecutive reads are the contiguous slice [lo, hi) of cpg_pos / cpg_rel.  Bulk copies need
// 16-byte aligned addresses and sizes: the staged range is [g0, g0 + staged) with g0 = lo rounded down to 8 calls
// (32 B of cpg_pos, 16 B of cpg_rel) and the end rounded up to 8 calls but never past the last whole group of the
// arrays; the (at most 7) calls beyond that are fetched with plain loads by the caller (`tail`).
struct SliceStage {
    uint32_t g0, lo, hi, staged;
};

// Called by ONE thread.  Returns the staging plan and issues the copies (pos always, rel when rel_src != nullptr).
template <int CAP>
__device__ __forceinline__ SliceStage stage_slice(uint32_t lo, uint32_t hi, int64_t n_calls_total, const int32_t* pos_src,
                                                  const uint16_t* rel_src, int32_t* s_pos, uint16_t* s_rel, uint64_t* bar) {
    SliceStage st;
    if (hi < lo || (int64_t)hi > n_calls_total) hi = lo;  // reported as ERRBIT_BAD_OFFSETS by the per-read check
    st.lo = lo;
    st.hi = hi;
    st.g0 = lo & ~7u;
    st.staged = 0;
    uint32_t g1 = min((hi + 7u) & ~7u, (uint32_t)(n_calls_total & ~7ll));
    if (hi > lo && hi - st.g0 <= (uint32_t)CAP && g1 > st.g0) {
        st.staged = g1 - st.g0;
        mbar_init(bar, 1);
        mbar_expect_tx(bar, st.staged * (rel_src ? 6u : 4u));
        bulk_g2s(s_pos, pos_src + st.g0, st.staged * 4u, bar);
        if (rel_src) bulk_g2s(s_rel, rel_src + st.g0, st.staged * 2u, bar);
    }
    return st;
}

}  // namespace mth
