// common.cuh — shared device-side types and helpers of the metheor_b200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mth {

constexpr uint32_t FULL = 0xffffffffu;
constexpr int MAX_METH_WORDS = 4;          // <= 256 CpG calls per read (DESIGN.md "Limits")
constexpr int MAX_CPGS_PER_READ = 64 * MAX_METH_WORDS;
constexpr int CONTIG_GAP = 1 << 16;        // spacing between contigs in the device (linear) coordinate
constexpr int MAX_REF_SPAN = CONTIG_GAP - 512;
constexpr uint32_t META_HALO = 1u << 9;    // MTH_META_HALO
// per-call flag byte written by k_ingest (call_flags[]): read-level decisions replicated on every call of the read
constexpr uint32_t CF_METH = 1u, CF_FIRST = 2u, CF_LPMD = 4u, CF_PDR_C = 8u, CF_PDR_D = 16u, CF_PM_OK = 32u, CF_ME_OK = 64u;

// device error bits (ctx->d_err)
enum : uint32_t {
    ERRBIT_UNSORTED = 1u << 0,
    ERRBIT_BAD_OFFSETS = 1u << 1,
    ERRBIT_TOO_MANY_CPGS = 1u << 2,
    ERRBIT_CPG_ORDER = 1u << 3,
    ERRBIT_SPAN = 1u << 4,
    ERRBIT_POS_RANGE = 1u << 5,
    ERRBIT_PILE_OVERFLOW = 1u << 6,
};

// Structure-of-arrays view of all reads of the current region, in file order, device (linear) coordinates.
struct ReadsView {
    const int32_t* __restrict__ start;
    const int32_t* __restrict__ end;
    const uint32_t* __restrict__ meta;
    const uint32_t* __restrict__ cpg_off;   // R+1
    const int32_t* __restrict__ cpg_pos;    // I
    const uint64_t* __restrict__ meth;
    const uint32_t* __restrict__ meth_off;  // R+1 or nullptr (one word per read)
    int64_t R;
    int64_t I;
};

// Region-wide scalars living in device memory (written by the ingest kernels, read by everything after).
struct RegionScalars {
    int32_t lmax;          // max(end-start+1)
    int32_t last_start;    // start of the last read ingested so far (sortedness across batches)
    uint32_t err;          // ERRBIT_*
    uint32_t pad;
    unsigned long long n_sites;
    unsigned long long lpmd[4];  // n_read, n_valid_read, n_conc, n_disc
    unsigned long long fdrp_pairs;  // read pairs compared by the FDRP / qFDRP kernels (pair-ops of SURVEY 8d)
    unsigned long long fallback_sites[2];  // MTH_FLAG_PROFILE: sites the MHL / FDRP tile kernels handed to the per-site kernels
};

struct ContigTable {       // device copy: contigs of the current region, ascending lin_off
    int32_t n;
    const int32_t* lin_off;  // position p of contig k lives at lin_off[k] + p
    const int32_t* tid;
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint64_t meth_word(const ReadsView& rv, int64_t j, int w) {
    uint32_t base = rv.meth_off ? rv.meth_off[j] : (uint32_t)j;
    return rv.meth[(size_t)base + w];
}

// mask with the low n bits set (n in 0..64)
__device__ __forceinline__ uint64_t low_mask64(uint32_t n) { return n >= 64 ? ~0ull : ((1ull << n) - 1ull); }

// readutil.rs:134-145 — discordant iff the read's CpG states are not all equal
__device__ __forceinline__ bool read_discordant(const ReadsView& rv, int64_t j, uint32_t n) {
    uint32_t nw = (n + 63) >> 6;
    bool any1 = false, any0 = false;
    for (uint32_t w = 0; w < nw; w++) {
        uint32_t bits = min(64u, n - 64u * w);
        uint64_t m = low_mask64(bits);
        uint64_t x = meth_word(rv, j, (int)w) & m;
        any1 |= (x != 0);
        any0 |= (x != m);
    }
    return any1 && any0;
}

__device__ __forceinline__ uint32_t meth_bit(const ReadsView& rv, int64_t j, uint32_t k) {
    return (uint32_t)((meth_word(rv, j, (int)(k >> 6)) >> (k & 63)) & 1ull);
}

// index of p in the strictly increasing list a[0..n), or -1
__device__ __forceinline__ int find_pos(const int32_t* __restrict__ a, uint32_t n, int32_t p) {
    if (n <= 8) {
        for (uint32_t i = 0; i < n; i++) {
            int32_t v = a[i];
            if (v == p) return (int)i;
            if (v > p) return -1;
        }
        return -1;
    }
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (a[mid] < p) lo = mid + 1; else hi = mid;
    }
    return (lo < n && a[lo] == p) ? (int)lo : -1;
}

// first index x in [0,n] with a[x] >= key, cooperatively by one warp (a sorted ascending)
__device__ __forceinline__ int64_t warp_lower_bound(const int32_t* __restrict__ a, int64_t n, int32_t key) {
    const int lane = lane_id();
    int64_t lo = 0, hi = n;  // answer in [lo, hi]; a[x] < key for x < lo; hi == n or a[hi] >= key
    while (hi - lo > 32) {
        int64_t step = (hi - lo + 32) / 33;
        int64_t idx = lo + step * (lane + 1);
        bool ge = (idx >= hi) || (a[idx] >= key);
        uint32_t m = __ballot_sync(FULL, ge);
        int f = m ? (__ffs(m) - 1) : 32;
        int64_t nlo = (f == 0) ? lo : (lo + step * f + 1);
        int64_t nhi = (f == 32) ? hi : min(hi, lo + step * (f + 1));
        lo = nlo;
        hi = nhi;
    }
    int64_t idx = lo + lane;
    bool ge = (idx >= hi) || (a[idx] >= key);
    uint32_t m = __ballot_sync(FULL, ge);
    int f = m ? (__ffs(m) - 1) : 32;
    return min(hi, lo + f);
}

// (tid,pos) of a device-coordinate position
__device__ __forceinline__ void delinearize(const ContigTable& ct, int32_t lin, int32_t* tid, int32_t* pos) {
    int lo = 0, hi = ct.n;  // last k with lin_off[k] - 1 <= lin
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (ct.lin_off[mid] - 1 <= lin) lo = mid; else hi = mid;
    }
    *tid = ct.tid[lo];
    *pos = lin - ct.lin_off[lo];
}

// Seeded reservoir draw shared with the oracle (oracle/metheor_oracle.cpp reservoir_draw): j in 1..=total
__host__ __device__ __forceinline__ uint32_t reservoir_draw(uint64_t seed, int32_t tid, int32_t pos, uint32_t total) {
    uint64_t x = seed ^ ((uint64_t)(uint32_t)tid * 0x9E3779B97F4A7C15ULL) ^
                 ((uint64_t)(uint32_t)pos * 0xBF58476D1CE4E5B9ULL) ^ ((uint64_t)total * 0x94D049BB133111EBULL);
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return 1u + (uint32_t)(((x >> 32) * (uint64_t)total) >> 32);
}

}  // namespace mth
