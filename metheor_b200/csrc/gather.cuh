// gather.cuh — the per-site GATHER skeleton shared by PDR (exact path), MHL, FDRP/qFDRP and PM/ME.
//
// One warp owns a run of consecutive CpG sites.  For a site p every read that can matter — contributors (reads
// calling p) and flush triggers between them — starts in [p - Lmax + 1, p + 1] (reads are sorted by start; a CpG
// call lies in [start-1, end]), i.e. a contiguous window of the read arrays.  The warp walks that window 32 reads
// at a time in FILE ORDER; `__ballot_sync` turns "calls p" / "would flush p" into two bit masks and the
// reference's streaming semantics are replayed on the masks:
//   * a trigger read (first CpG > p + SLACK, eligible per measure) closes the live segment of p — the
//     `cpg2reads.retain(...)` flush of pdr.rs:160-177 / mhl.rs:163-172 / fdrp.rs:213-222;
//   * a closed segment with depth >= min_depth overwrites the site's result (`result.insert`), so the value
//     reported is that of the LAST qualifying segment;
//   * the segment still open at the end of the window is closed by the final flush (pdr.rs:199-210 …).
// Policies supply eligibility, SLACK and the per-segment accumulator.
#pragma once
#include "common.cuh"

namespace mth {

struct LaneRead {
    int64_t j;       // read index
    int32_t start;
    uint32_t o0;     // cpg_off[j]
    uint32_t n;      // number of CpG calls
    uint32_t mapq;
    int32_t first;   // first CpG position (INT32_MIN if none / read outside the window)
    int idx;         // index of the site within the read's CpG list, -1 if the read does not call it
    bool active;     // read lies in the window of the site
};

constexpr int SITES_PER_WARP = 16;
constexpr int GATHER_BLOCK = 256;

// Walk the read window of site p (starting at read index lo, a multiple-of-32-agnostic lower bound) and call
// f(lr) once per chunk of 32 reads with every lane participating.
template <class F>
__device__ __forceinline__ void scan_window(const ReadsView& rv, int64_t lo, int32_t p, int32_t target, F&& f) {
    const int lane = lane_id();
    for (int64_t base = lo; base < rv.R; base += 32) {
        LaneRead lr;
        lr.j = base + lane;
        bool in = lr.j < rv.R;
        lr.start = in ? rv.start[lr.j] : INT32_MAX;
        if (__shfl_sync(FULL, lr.start, 0) > p + 1) break;  // whole chunk starts after the site
        lr.active = in && lr.start <= p + 1 && lr.start >= target;
        lr.o0 = 0; lr.n = 0; lr.mapq = 0; lr.idx = -1; lr.first = INT32_MIN;
        if (lr.active) {
            lr.o0 = rv.cpg_off[lr.j];
            lr.n = rv.cpg_off[lr.j + 1] - lr.o0;
            if (lr.n) {
                lr.mapq = rv.meta[lr.j] & 0xFFu;
                const int32_t* cp = rv.cpg_pos + lr.o0;
                lr.first = cp[0];
                if (lr.first <= p && cp[lr.n - 1] >= p) lr.idx = find_pos(cp, lr.n, p);
            }
        }
        f(lr);
    }
}

// Iterate the sites owned by this warp; body(s, p, lo, target) is called with the window lower bound maintained.
// `only` != nullptr restricts the walk to the sites it flags.  Flagged sites cluster (CpG islands), so in that mode sites
// are dealt to the warps one by one, round robin, instead of in runs of SITES_PER_WARP: neighbours land on different warps.
template <class Body>
__device__ __forceinline__ void for_each_site(const ReadsView& rv, const int32_t* __restrict__ site_pos, int64_t C,
                                              int32_t lmax, Body&& body, const uint8_t* __restrict__ only = nullptr) {
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    if (only) {
        for (int64_t base = 0; base < C; base += n_warps * 32) {
            // 32 candidate sites per warp and step: lane l looks at site base + (l * n_warps + warp_global)
            const int64_t sl = base + (int64_t)lane_id() * n_warps + warp_global;
            uint32_t want = __ballot_sync(FULL, sl < C && only[sl] != 0);
            while (want) {
                const int l = __ffs(want) - 1;
                want &= want - 1;
                const int64_t s = base + (int64_t)l * n_warps + warp_global;
                const int32_t p = site_pos[s];
                const int32_t target = p - lmax + 1;
                const int64_t lo = warp_lower_bound(rv.start, rv.R, target);
                body(s, p, lo, target);
            }
        }
        return;
    }
    for (int64_t s0 = warp_global * SITES_PER_WARP; s0 < C; s0 += n_warps * SITES_PER_WARP) {
        int64_t s1 = min(C, s0 + SITES_PER_WARP);
        int64_t lo = warp_lower_bound(rv.start, rv.R, site_pos[s0] - lmax + 1);
        for (int64_t s = s0; s < s1; s++) {
            const int32_t p = site_pos[s];
            const int32_t target = p - lmax + 1;
            while (lo + 32 <= rv.R && rv.start[lo + 31] < target) lo += 32;  // warp-uniform
            body(s, p, lo, target);
        }
    }
}

// The sites of a compact list (built from per-site flags when the flagged sites are few: the walk over all the flags — and a grid
// sized for it — costs more than the sites themselves), dealt to the warps one by one.
template <class Body>
__device__ __forceinline__ void for_each_listed_site(const ReadsView& rv, const int32_t* __restrict__ site_pos, int32_t lmax,
                                                     const uint32_t* __restrict__ list, unsigned long long n_list, Body&& body) {
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (unsigned long long i = (unsigned long long)warp_global; i < n_list; i += (unsigned long long)n_warps) {
        const int64_t s = (int64_t)list[i];
        const int32_t p = site_pos[s];
        const int32_t target = p - lmax + 1;
        const int64_t lo = warp_lower_bound(rv.start, rv.R, target);
        body(s, p, lo, target);
    }
}

// Policy interface:
//   static constexpr int SLACK;
//   bool contrib_ok(mapq, n) / trigger_ok(mapq, n)
//   void begin_site(p) ; void add(mask, lr) ; void close() ; void end_site(site_index)
// `only` != nullptr: just the sites it flags (the ones a tile-form kernel handed back).
template <class Policy>
__device__ __forceinline__ void gather_sites(const ReadsView& rv, const int32_t* __restrict__ site_pos, int64_t C,
                                             int32_t lmax, Policy& pol, const uint8_t* __restrict__ only = nullptr) {
    for_each_site(rv, site_pos, C, lmax, [&](int64_t s, int32_t p, int64_t lo, int32_t target) {
        pol.begin_site(p);
        scan_window(rv, lo, p, target, [&](const LaneRead& lr) {
            bool contributes = lr.idx >= 0 && pol.contrib_ok(lr.mapq, lr.n);
            bool trig = lr.active && lr.n > 0 && lr.first > p + Policy::SLACK && pol.trigger_ok(lr.mapq, lr.n);
            uint32_t cm = __ballot_sync(FULL, contributes);
            uint32_t tm = __ballot_sync(FULL, trig);
            while (cm | tm) {  // replay contributors / triggers in file order
                int tpos = tm ? (__ffs(tm) - 1) : 32;
                uint32_t below = tpos >= 32 ? FULL : ((1u << tpos) - 1u);
                uint32_t cb = cm & below;
                if (cb) {
                    pol.add(cb, lr);
                    cm &= ~cb;
                }
                if (tpos < 32) {
                    pol.close();
                    int nc = cm ? (__ffs(cm) - 1) : 32;  // triggers up to the next contributor add nothing
                    uint32_t clr = nc >= 32 ? FULL : ((1u << nc) - 1u);
                    tm &= ~clr;
                }
            }
        });
        pol.close();
        pol.end_site(s);
    }, only);
}

}  // namespace mth
