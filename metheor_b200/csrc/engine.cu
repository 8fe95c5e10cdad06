// engine.cu — host side of libmetheor_b200.so: the C ABI of include/metheor_b200.h, device memory management,
// stream plumbing and the launch sequence.  No measure arithmetic happens on the host (the one exception is the
// table of p*log2f(p) values built once per context so that ME matches the host libm bit for bit, see k_quartet.cu).
//
// Data flow:
//   mth_submit : H2D copy of the SoA batch on the copy stream  ->  [offset / coordinate fix-ups]  ->  k_ingest
//                (validation, site bitmap, LPMD) on the compute stream, ordered by an event.
//   region close (contig set full, or mth_finish): bitmap -> site dictionary -> one kernel family per measure ->
//                row counts -> exclusive scan -> row emission into device row buffers.
//   mth_finish : D2H of the rows into pinned buffers owned by the context.
#include <dlfcn.h>
#include <emmintrin.h>
#include <math.h>
#include <nccl.h>  // types only: the library is loaded with dlopen at the first mth_comm_* call
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <vector>

#include "kernels.h"

using namespace mth;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};
struct HostBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct SiteRowsBuf {
    DevBuf tid, pos, value, nc, nd;
    HostBuf h_tid, h_pos, h_value, h_nc, h_nd;
    int64_t n = 0;
};
struct QuartetRowsBuf {
    DevBuf tid, p1, p2, p3, p4, value, counts;
    HostBuf h_tid, h_p1, h_p2, h_p3, h_p4, h_value, h_counts;
    int64_t n = 0;
};

struct PairRowsBuf {
    DevBuf tid, pos1, pos2, lpmd, nc, nd;
    HostBuf h_tid, h_pos1, h_pos2, h_lpmd, h_nc, h_nd;
    int64_t n = 0;
};

struct ProfSpan {
    const char* name;
    cudaEvent_t e0, e1;
    int launches;
};

constexpr int64_t COMPACT_PIECE = 1 << 21;  // reads per H2D piece of a compact host batch
constexpr int64_t BORROWED_REGION_MIN_READS = 1 << 20;  // see place_batch

enum { M_PDR = 0, M_MHL, M_FDRP, M_QFDRP, M_PM, M_ME, M_PAIRS, M_COUNT };

// What belongs to ONE region: its reads (view), its contigs and everything the ingest pass writes.  The context holds the
// region being filled; a region whose processing is deferred behind the next region's ingest pass (see mth_submit) is
// parked in a RegionJob, and swap_slot() puts it back into the context for the calls that work on it.
struct RegionSlot {
    bool region_active = false;
    int32_t cur_lin_off = 0;
    int64_t R = 0, I = 0, W = 0;
    bool borrowed = false;
    mth_batch bview;
    bool has_meth_off = false;
    std::vector<int32_t> reg_lin_off, reg_tid;
    DevBuf bitmap, scalars, a_flags;
    size_t bitmap_words_valid = 0;
    HostBuf h_scalars, h_totals;
};
// Host-side values the three steps of a region hand to each other: close (site count queued) -> phase A (measure kernels,
// row counts, scans) -> phase B (row emission).
struct RegionCarry {
    int64_t C = 0;
    int q_set[2] = {0, 1};
    unsigned long long tot[M_COUNT] = {0, 0, 0, 0, 0, 0, 0};
    size_t nct = 0;
    bool done = false;                            // nothing left to do (a region without sites ends in phase A)
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;   // deferred regions wait on events; nullptr = synchronise the stream
};
struct RegionJob {
    RegionSlot st;
    RegionCarry k;
};

}  // namespace

struct mth_ctx {
    int device = 0;
    mth_params prm;
    std::vector<int64_t> ref_len;
    cudaStream_t own_compute = nullptr, copy = nullptr, compute = nullptr;
    cudaStream_t aux = nullptr;  // side stream of process_region: the latency-bound mixed-site gather of PM / ME beside the streaming kernels
    cudaEvent_t ev_copy = nullptr, ev_compute = nullptr, ev_fork = nullptr, ev_join = nullptr;
    std::string err;
    int finished = 0;

    // Regions in flight behind the one being filled (borrowed device contigs only, see mth_submit): `pend` has its phase A
    // queued and waits for phase B; `next` is the parking place of the region that was just closed (and, between uses, the
    // owner of the spare buffer set).
    RegionJob pend, next;
    bool has_pend = false;
    bool pipeline = true;

    // region state
    bool region_active = false;
    int32_t last_tid = -1;
    int32_t cur_lin_off = 0;
    int64_t R = 0, I = 0, W = 0;
    bool borrowed = false;
    mth_batch bview;  // borrowed device pointers
    bool has_meth_off = false;
    std::vector<int32_t> reg_lin_off, reg_tid;

    // arena
    DevBuf a_start, a_end, a_meta, a_off, a_pos, a_rel, a_meth, a_moff, a_flags;
    DevBuf bitmap, word_prefix, block_sums, site_pos, scalars, totals, ct_lin, ct_tid;
    DevBuf cnt2, scan_scratch, fdrp_scratch, me_lut, lpmd_total;
    DevBuf gfallback;            // sites the thread-per-site gather kernels hand to the warp-per-site form
    // PM / ME: per-(site, slot) observation counts, mixed-site flags, the observation list (site / slot+pattern / length) and
    // the 16-bin histograms of the output rows (k_quartet.cu)
    DevBuf qcnt[2], qmixed[2], qobs_site[2], qobs_vp[2], qobs_n, qhrows[2], qmix_list[2], qmix_n;
    DevBuf stage[2][13], exp_blocks, exp_tot;  // compact wire format: two staging sets + scan scratch
    cudaEvent_t ev_stage_free[2] = {nullptr, nullptr};
    bool stage_busy[2] = {false, false};
    int stage_next = 0;
    DevBuf rowcnt[M_COUNT], value[M_COUNT];
    size_t bitmap_words_valid = 0;
    HostBuf h_scalars, h_totals;

    SiteRowsBuf rows_pdr, rows_mhl, rows_fdrp, rows_qfdrp;
    QuartetRowsBuf rows_pm, rows_me;
    PairRowsBuf rows_pairs;
    int64_t lpmd_total_host[4] = {0, 0, 0, 0};
    int me_lut_max = 0;

    // --cpg-set on the device (k_cpgset.cu): the set as per-contig sorted position lists, the region's set bitmap and scratch
    bool has_set = false;
    std::vector<int64_t> set_off;   // n_ref + 1 offsets into set_dev
    DevBuf set_dev, set_bitmap, set_kept, set_scan, set_total, set_pos_tmp, set_rel_tmp;
    HostBuf h_set_total, h_ct;
    size_t set_words_valid = 0;

    ncclComm_t comm = nullptr;  // multi-GPU: joins this context with its peers (mth_comm_init_*)
    int comm_ranks = 0;

    mth_stats stats;
    std::vector<ProfSpan> spans;
    std::vector<cudaEvent_t> ev_pool;
};

static std::string g_create_err;

#define CUDA_TRY(ctx, call)                                                                             \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(_e);                            \
            return MTH_ERR_CUDA;                                                                        \
        }                                                                                               \
    } while (0)

#define TRY(expr)                      \
    do {                               \
        int _rc = (expr);              \
        if (_rc != MTH_OK) return _rc; \
    } while (0)

static int fail(mth_ctx* c, int code, const std::string& msg) {
    c->err = msg;
    return code;
}

// ---- memory helpers ----------------------------------------------------------------------------
static int dev_reserve(mth_ctx* c, DevBuf& b, size_t bytes, size_t keep_bytes) {
    if (bytes <= b.cap) return MTH_OK;
    size_t ncap = bytes + bytes / 2 + 256;
    void* np = nullptr;
    // pending async work may still touch the old allocation
    CUDA_TRY(c, cudaStreamSynchronize(c->copy));
    CUDA_TRY(c, cudaStreamSynchronize(c->compute));
    cudaError_t e = cudaMalloc(&np, ncap);
    if (e != cudaSuccess) {
        ncap = bytes;
        CUDA_TRY(c, cudaMalloc(&np, ncap));
    }
    if (b.p && keep_bytes) CUDA_TRY(c, cudaMemcpy(np, b.p, keep_bytes, cudaMemcpyDeviceToDevice));
    if (b.p) CUDA_TRY(c, cudaFree(b.p));
    b.p = np;
    b.cap = ncap;
    return MTH_OK;
}
static int host_reserve(mth_ctx* c, HostBuf& b, size_t bytes) {
    if (bytes <= b.cap) return MTH_OK;
    if (b.p) CUDA_TRY(c, cudaFreeHost(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t ncap = bytes + bytes / 4 + 256;
    CUDA_TRY(c, cudaHostAlloc(&b.p, ncap, cudaHostAllocDefault));
    b.cap = ncap;
    return MTH_OK;
}
static void dev_free(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}
static void host_free(HostBuf& b) {
    if (b.p) cudaFreeHost(b.p);
    b.p = nullptr;
    b.cap = 0;
}

// ---- profiling -----------------------------------------------------------------------------------
static cudaEvent_t get_event(mth_ctx* c) {
    if (!c->ev_pool.empty()) {
        cudaEvent_t e = c->ev_pool.back();
        c->ev_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
struct ProfScope {
    mth_ctx* c;
    ProfSpan sp;
    bool on;
    ProfScope(mth_ctx* ctx, const char* name) : c(ctx) {
        on = (c->prm.flags & MTH_FLAG_PROFILE) != 0;
        sp.name = name;
        sp.launches = 0;
        if (on) {
            sp.e0 = get_event(c);
            sp.e1 = get_event(c);
            cudaEventRecord(sp.e0, c->compute);
        }
    }
    void add(int n) {
        sp.launches += n;
        c->stats.kernel_launches += n;
    }
    void add0(int n) { sp.launches += n; }  // nested detail scope ("~name"): the launch is counted by the enclosing scope
    ~ProfScope() {
        if (on) {
            cudaEventRecord(sp.e1, c->compute);
            c->spans.push_back(sp);
        }
    }
};
static void resolve_spans(mth_ctx* c) {
    for (auto& sp : c->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.e0, sp.e1) != cudaSuccess) ms = 0.f;
        int k = -1;
        for (int i = 0; i < c->stats.n_kernel_stats; i++)
            if (!strcmp(c->stats.kernel[i].name, sp.name)) k = i;
        if (k < 0 && c->stats.n_kernel_stats < MTH_MAX_KERNEL_STATS) {
            k = c->stats.n_kernel_stats++;
            strncpy(c->stats.kernel[k].name, sp.name, sizeof(c->stats.kernel[k].name) - 1);
            c->stats.kernel[k].launches = 0;
            c->stats.kernel[k].ms = 0;
        }
        if (k >= 0) {
            c->stats.kernel[k].launches += sp.launches;
            c->stats.kernel[k].ms += ms;
        }
        c->ev_pool.push_back(sp.e0);
        c->ev_pool.push_back(sp.e1);
    }
    c->spans.clear();
}

// ---- region management ---------------------------------------------------------------------------
static ReadsView make_view(mth_ctx* c) {
    ReadsView v;
    if (c->borrowed) {
        v.start = c->bview.start; v.end = c->bview.end; v.meta = c->bview.meta; v.cpg_off = c->bview.cpg_off;
        v.cpg_pos = c->bview.cpg_pos; v.meth = c->bview.meth; v.meth_off = c->bview.meth_off;
    } else {
        v.start = (const int32_t*)c->a_start.p; v.end = (const int32_t*)c->a_end.p; v.meta = (const uint32_t*)c->a_meta.p;
        v.cpg_off = (const uint32_t*)c->a_off.p; v.cpg_pos = (const int32_t*)c->a_pos.p;
        v.meth = (const uint64_t*)c->a_meth.p; v.meth_off = c->has_meth_off ? (const uint32_t*)c->a_moff.p : nullptr;
    }
    v.R = c->R;
    v.I = c->I;
    return v;
}

static int add_contig(mth_ctx* c, int32_t tid, int32_t lin_off) {
    c->reg_lin_off.push_back(lin_off);
    c->reg_tid.push_back(tid);
    c->cur_lin_off = lin_off;
    c->last_tid = tid;
    // bits (p+1) for p in [lin_off-1, lin_off+len): words covering [lin_off, lin_off+len+1]
    size_t w0 = (size_t)lin_off >> 6;
    size_t w1 = (((size_t)lin_off + (size_t)c->ref_len[tid] + 2) >> 6) + 1;
    TRY(dev_reserve(c, c->bitmap, w1 * 8, c->bitmap_words_valid * 8));
    CUDA_TRY(c, cudaMemsetAsync((char*)c->bitmap.p + w0 * 8, 0, (w1 - w0) * 8, c->compute));
    c->bitmap_words_valid = w1;
    if (c->has_set) {  // the contig's part of the CpG-set bitmap (same coordinate as the site bitmap)
        TRY(dev_reserve(c, c->set_bitmap, w1 * 8, c->set_words_valid * 8));
        CUDA_TRY(c, cudaMemsetAsync((char*)c->set_bitmap.p + w0 * 8, 0, (w1 - w0) * 8, c->compute));
        c->set_words_valid = w1;
        const int64_t a = c->set_off[(size_t)tid], b = c->set_off[(size_t)tid + 1];
        c->stats.kernel_launches += launch_cpgset_mark((const int32_t*)c->set_dev.p + a, b - a, lin_off, (int32_t)c->ref_len[tid],
                                                       (unsigned long long*)c->set_bitmap.p, c->compute);
    }
    return MTH_OK;
}

static int begin_region(mth_ctx* c, int32_t tid) {
    c->region_active = true;
    c->R = c->I = c->W = 0;
    c->borrowed = false;
    c->has_meth_off = false;
    c->reg_lin_off.clear();
    c->reg_tid.clear();
    c->bitmap_words_valid = 0;
    c->set_words_valid = 0;
    TRY(dev_reserve(c, c->scalars, sizeof(RegionScalars), 0));
    RegionScalars init;
    memset(&init, 0, sizeof(init));
    TRY(host_reserve(c, c->h_scalars, 2 * sizeof(RegionScalars)));
    memcpy((char*)c->h_scalars.p + sizeof(RegionScalars), &init, sizeof(init));
    CUDA_TRY(c, cudaMemcpyAsync(c->scalars.p, (char*)c->h_scalars.p + sizeof(RegionScalars), sizeof(init),
                                cudaMemcpyHostToDevice, c->compute));
    return add_contig(c, tid, 0);
}

static int ensure_arena(mth_ctx* c, int64_t R, int64_t I, int64_t W, bool need_rel, bool need_moff) {
    TRY(dev_reserve(c, c->a_start, (size_t)R * 4, (size_t)c->R * 4));
    TRY(dev_reserve(c, c->a_end, (size_t)R * 4, (size_t)c->R * 4));
    TRY(dev_reserve(c, c->a_meta, (size_t)R * 4, (size_t)c->R * 4));
    TRY(dev_reserve(c, c->a_off, (size_t)(R + 1) * 4, (size_t)(c->R + 1) * 4));
    TRY(dev_reserve(c, c->a_pos, (size_t)I * 4 + 4, (size_t)c->I * 4));
    TRY(dev_reserve(c, c->a_meth, (size_t)W * 8 + 8, (size_t)c->W * 8));
    if (need_rel) TRY(dev_reserve(c, c->a_rel, (size_t)I * 2 + 2, (size_t)c->I * 2));
    if (need_moff) TRY(dev_reserve(c, c->a_moff, (size_t)(R + 1) * 4, (size_t)(c->R + 1) * 4));
    return MTH_OK;
}

static int64_t batch_words(const mth_batch* b) { return b->meth_off ? b->n_meth_words : b->n_reads; }

// copy a borrowed single-batch region into the arena so that more batches can be appended
static int materialize(mth_ctx* c) {
    if (!c->borrowed) return MTH_OK;
    const mth_batch& b = c->bview;
    bool lp = (c->prm.measures & MTH_LPMD) != 0;
    c->borrowed = false;
    int64_t R = c->R, I = c->I, W = c->W;
    c->R = c->I = c->W = 0;  // nothing to preserve
    TRY(ensure_arena(c, R, I, W, lp, c->has_meth_off));
    c->R = R; c->I = I; c->W = W;
    cudaStream_t s = c->compute;
    CUDA_TRY(c, cudaMemcpyAsync(c->a_start.p, b.start, (size_t)R * 4, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(c, cudaMemcpyAsync(c->a_end.p, b.end, (size_t)R * 4, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(c, cudaMemcpyAsync(c->a_meta.p, b.meta, (size_t)R * 4, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(c, cudaMemcpyAsync(c->a_off.p, b.cpg_off, (size_t)(R + 1) * 4, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(c, cudaMemcpyAsync(c->a_pos.p, b.cpg_pos, (size_t)I * 4, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(c, cudaMemcpyAsync(c->a_meth.p, b.meth, (size_t)W * 8, cudaMemcpyDeviceToDevice, s));
    if (c->has_meth_off)
        CUDA_TRY(c, cudaMemcpyAsync(c->a_moff.p, b.meth_off, (size_t)(R + 1) * 4, cudaMemcpyDeviceToDevice, s));
    if (lp && I) CUDA_TRY(c, cudaMemcpyAsync(c->a_rel.p, b.cpg_rel, (size_t)I * 2, cudaMemcpyDeviceToDevice, s));
    return MTH_OK;
}

static int process_region(mth_ctx* c);
static void swap_slot(mth_ctx* c, RegionSlot& o);
static int region_close(mth_ctx* c, RegionCarry& k);
static int region_phase_a(mth_ctx* c, RegionCarry& k);
static int finish_pending(mth_ctx* c);

static int check_scalars_err(mth_ctx* c, uint32_t e) {
    if (!e) return MTH_OK;
    if (e & ERRBIT_UNSORTED) return fail(c, MTH_ERR_UNSORTED, "reads are not sorted by start position within a contig");
    if (e & ERRBIT_BAD_OFFSETS) return fail(c, MTH_ERR_INVALID, "cpg_off / meth_off are not monotone prefix offsets");
    if (e & ERRBIT_TOO_MANY_CPGS)
        return fail(c, MTH_ERR_UNSUPPORTED, "a read carries more CpG calls than supported (64 without meth_off, 256 with)");
    if (e & ERRBIT_CPG_ORDER) return fail(c, MTH_ERR_INVALID, "cpg_pos not strictly increasing within a read");
    if (e & ERRBIT_SPAN) return fail(c, MTH_ERR_UNSUPPORTED, "read reference span < 1 or > 65024 bp");
    if (e & ERRBIT_POS_RANGE)
        return fail(c, MTH_ERR_INVALID, "read or CpG position outside its contig / CpG outside [start-1, end]");
    if (e & ERRBIT_PILE_OVERFLOW) return fail(c, MTH_ERR_UNSUPPORTED, "FDRP pile scratch overflow");
    return fail(c, MTH_ERR_INVALID, "device-side validation failed");
}

// Host-side sums over one piece of a compact batch (SSE2 sum-of-absolute-differences: ~16 bytes per cycle, so that
// slicing a large host batch into pieces costs microseconds, not milliseconds):
//   *calls    = sum n8[r]
//   *explicit = sum n8[r] over reads whose flags carry MTH_CFLAG_REL_EXPLICIT (flags == nullptr: 0)
static void sum_call_counts(const uint8_t* n8, const uint8_t* flags, int64_t n, uint64_t* calls, uint64_t* explicit_rel) {
    const __m128i zero = _mm_setzero_si128(), bit = _mm_set1_epi8((char)MTH_CFLAG_REL_EXPLICIT);
    __m128i acc = zero, acc_e = zero;
    int64_t r = 0;
    for (; r + 16 <= n; r += 16) {
        const __m128i v = _mm_loadu_si128((const __m128i*)(n8 + r));
        acc = _mm_add_epi64(acc, _mm_sad_epu8(v, zero));
        if (flags) {
            const __m128i f = _mm_loadu_si128((const __m128i*)(flags + r));
            const __m128i m = _mm_cmpeq_epi8(_mm_and_si128(f, bit), bit);
            acc_e = _mm_add_epi64(acc_e, _mm_sad_epu8(_mm_and_si128(v, m), zero));
        }
    }
    uint64_t t[2], te[2];
    _mm_storeu_si128((__m128i*)t, acc);
    _mm_storeu_si128((__m128i*)te, acc_e);
    uint64_t c = t[0] + t[1], e = te[0] + te[1];
    for (; r < n; r++) {
        c += n8[r];
        if (flags && (flags[r] & MTH_CFLAG_REL_EXPLICIT)) e += n8[r];
    }
    *calls = c;
    *explicit_rel = e;
}

// Decides where a new batch goes: same region (possibly a new contig appended behind a gap) or a fresh region after the
// current one has been processed.
static int place_batch(mth_ctx* c, int32_t tid, int64_t n_reads, int64_t n_cpg, int64_t n_words) {
    if (c->region_active) {
        if (tid < c->last_tid) return fail(c, MTH_ERR_UNSORTED, "contigs must arrive in ascending tid order (coordinate-sorted input)");
        if (tid != c->last_tid) {
            int64_t noff = (((int64_t)c->cur_lin_off + c->ref_len[c->last_tid] + CONTIG_GAP) + 63) & ~63ll;
            bool fits = noff + c->ref_len[tid] + 64 < (int64_t)INT32_MAX && c->I + n_cpg < (int64_t)UINT32_MAX - 64 &&
                        c->R + n_reads < (int64_t)INT32_MAX - 64 && c->W + n_words < (int64_t)UINT32_MAX - 64;
            // A large device-resident contig that is being read in place (zero copy) closes its own region: packing the next
            // contig behind it would first copy the whole contig into the arena, which costs more than a region boundary.
            if (c->borrowed && c->R >= BORROWED_REGION_MIN_READS) fits = false;
            if (fits) {
                TRY(materialize(c));
                TRY(add_contig(c, tid, (int32_t)noff));
            } else {
                TRY(process_region(c));
            }
        } else if (c->I + n_cpg >= (int64_t)UINT32_MAX - 64 || c->R + n_reads >= (int64_t)INT32_MAX - 64) {
            return fail(c, MTH_ERR_UNSUPPORTED, "a single contig holds more than 2^32 CpG calls / 2^31 reads");
        }
    }
    if (!c->region_active) TRY(begin_region(c, tid));
    return MTH_OK;
}

// The ingest pass over reads [r0, r0+n) of the region (already in device memory, linear coordinates).
static int run_ingest(mth_ctx* c, int32_t tid, int64_t r0, int64_t n, int64_t i0, int64_t n_cpg, const uint16_t* rel_dev) {
    const bool lp = (c->prm.measures & MTH_LPMD) != 0;
    IngestArgs ia;
    ia.rv = make_view(c);
    ia.r0 = r0;
    ia.n = n;
    ia.i0 = i0;
    ia.cpg_rel = rel_dev;
    ia.bitmap = (unsigned long long*)c->bitmap.p;
    ia.lin_lo = c->cur_lin_off;
    ia.lin_hi = c->cur_lin_off + (int32_t)c->ref_len[tid];
    ia.do_lpmd = lp ? 1 : 0;
    ia.lpmd = c->prm.lpmd;
    ia.do_pdr = (c->prm.measures & MTH_PDR) ? 1 : 0;
    ia.pdr = c->prm.pdr;
    ia.do_pm = (c->prm.measures & MTH_PM) ? 1 : 0;
    ia.do_me = (c->prm.measures & MTH_ME) ? 1 : 0;
    ia.pm_min_qual = c->prm.pm.min_qual;
    ia.me_min_qual = c->prm.me.min_qual;
    TRY(dev_reserve(c, c->a_flags, (size_t)c->I + 64, (size_t)i0));
    ia.call_flags = (uint8_t*)c->a_flags.p;
    ia.sc = (RegionScalars*)c->scalars.p;
    {
        ProfScope ps(c, "k_ingest");
        ps.add(launch_ingest(ia, c->compute));
    }
    CUDA_TRY(c, cudaGetLastError());
    c->stats.n_reads += n;
    c->stats.n_cpg += n_cpg;
    return MTH_OK;
}

// --cpg-set: drop the calls of reads [r0, r0 + n) (calls [i0, i0 + n_cpg) of the arena) that are not in the set, compacting the
// batch's slice of the call arrays; *kept = calls left.  Synchronises the compute stream (the host must know where the next
// batch goes).
static int filter_batch(mth_ctx* c, int64_t r0, int64_t n, int64_t i0, int64_t n_cpg, int64_t* kept) {
    *kept = n_cpg;
    if (!c->has_set || n <= 0) return MTH_OK;
    const bool lp = (c->prm.measures & MTH_LPMD) != 0;
    TRY(dev_reserve(c, c->set_kept, (size_t)n * 4 + 64, 0));
    TRY(dev_reserve(c, c->set_scan, (size_t)((n + 2047) / 2048 + 2) * 4, 0));
    TRY(dev_reserve(c, c->set_total, 16, 0));
    TRY(dev_reserve(c, c->set_pos_tmp, (size_t)n_cpg * 4 + 64, 0));
    if (lp) TRY(dev_reserve(c, c->set_rel_tmp, (size_t)n_cpg * 2 + 64, 0));
    TRY(host_reserve(c, c->h_set_total, 16));
    cudaStream_t s = c->compute;
    {
        ProfScope ps(c, "k_cpgset_filter");
        ps.add(launch_cpgset_filter(r0, n, i0, (uint32_t*)c->a_off.p, (const int32_t*)c->a_pos.p, lp ? (const uint16_t*)c->a_rel.p : nullptr,
                                    (uint64_t*)c->a_meth.p, c->has_meth_off ? (const uint32_t*)c->a_moff.p : nullptr,
                                    (const unsigned long long*)c->set_bitmap.p, (int64_t)c->set_words_valid, (uint32_t*)c->set_kept.p,
                                    (uint32_t*)c->set_scan.p, (unsigned long long*)c->set_total.p, (int32_t*)c->set_pos_tmp.p,
                                    lp ? (uint16_t*)c->set_rel_tmp.p : nullptr, s));
    }
    CUDA_TRY(c, cudaMemcpyAsync(c->h_set_total.p, c->set_total.p, 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaStreamSynchronize(s));
    const int64_t k = (int64_t) * (unsigned long long*)c->h_set_total.p;
    if (k > n_cpg) return fail(c, MTH_ERR_INVALID, "cpg_off inconsistent with n_cpg");
    if (k) {
        CUDA_TRY(c, cudaMemcpyAsync((int32_t*)c->a_pos.p + i0, c->set_pos_tmp.p, (size_t)k * 4, cudaMemcpyDeviceToDevice, s));
        if (lp) CUDA_TRY(c, cudaMemcpyAsync((uint16_t*)c->a_rel.p + i0, c->set_rel_tmp.p, (size_t)k * 2, cudaMemcpyDeviceToDevice, s));
    }
    *kept = k;
    return MTH_OK;
}

// ---- C ABI -----------------------------------------------------------------------------------------
extern "C" {

int mth_set_cpg_set(mth_ctx* c, int64_t n, const int32_t* tid, const int32_t* pos) {
    if (!c || n < 0 || (n && (!tid || !pos))) return MTH_ERR_INVALID;
    if (c->region_active || c->stats.n_reads) return fail(c, MTH_ERR_STATE, "mth_set_cpg_set must be called before the first batch (after mth_reset)");
    CUDA_TRY(c, cudaSetDevice(c->device));
    const size_t n_ref = c->ref_len.size();
    std::vector<int64_t> cnt(n_ref + 1, 0);
    for (int64_t i = 0; i < n; i++) {
        if (tid[i] < 0 || (size_t)tid[i] >= n_ref) return fail(c, MTH_ERR_INVALID, "CpG set entry with a tid outside the reference list");
        cnt[(size_t)tid[i] + 1]++;
    }
    for (size_t t = 0; t < n_ref; t++) cnt[t + 1] += cnt[t];
    std::vector<int32_t> sorted((size_t)n);
    {
        std::vector<int64_t> at(cnt.begin(), cnt.end() - 1);
        for (int64_t i = 0; i < n; i++) sorted[(size_t)at[(size_t)tid[i]]++] = pos[i];
    }
    c->set_off = cnt;
    TRY(dev_reserve(c, c->set_dev, (size_t)n * 4 + 64, 0));
    if (n) CUDA_TRY(c, cudaMemcpy(c->set_dev.p, sorted.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    c->has_set = true;
    return MTH_OK;
}

int mth_clear_cpg_set(mth_ctx* c) {
    if (!c) return MTH_ERR_INVALID;
    if (c->region_active || c->stats.n_reads) return fail(c, MTH_ERR_STATE, "mth_clear_cpg_set must be called before the first batch (after mth_reset)");
    c->has_set = false;
    c->set_off.clear();
    return MTH_OK;
}

void mth_params_default(mth_params* p) {
    memset(p, 0, sizeof(*p));
    p->abi_version = MTH_ABI_VERSION;
    p->pdr = {10, 4, 10};
    p->lpmd = {2, 16, 10, 0};
    p->mhl = {10, 4, 10};
    p->pm = {10, 10};
    p->me = {10, 10};
    p->fdrp = {10, 10, 40, 35};
    p->qfdrp = {10, 10, 40, 35};
    p->seed = 0;
}

const char* mth_version(void) { return "metheor_b200 0.1.0 (sm_100a)"; }

int mth_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

void* mth_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void mth_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

uint32_t mth_reservoir_draw(uint64_t seed, int32_t tid, int32_t pos, uint32_t total) {
    return reservoir_draw(seed, tid, pos, total);
}

const char* mth_last_error(mth_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int mth_ctx_create(mth_ctx** out, int device, const mth_params* params, int32_t n_ref, const int64_t* ref_len) {
    if (!out || !params || n_ref < 0 || (n_ref > 0 && !ref_len)) { g_create_err = "null argument"; return MTH_ERR_INVALID; }
    *out = nullptr;
    if (params->abi_version != MTH_ABI_VERSION) { g_create_err = "mth_params.abi_version mismatch"; return MTH_ERR_INVALID; }
    if (params->measures & ~MTH_ALL) { g_create_err = "unknown measure bits"; return MTH_ERR_INVALID; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_err = std::string("no CUDA device available (the engine has no CPU fallback): ") + cudaGetErrorString(e);
        return MTH_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { g_create_err = "device index out of range"; return MTH_ERR_INVALID; }
    for (int32_t i = 0; i < n_ref; i++)
        if (ref_len[i] < 0 || ref_len[i] > (int64_t)INT32_MAX - 4 * CONTIG_GAP) {
            g_create_err = "contig length outside the supported range (< 2^31 - 2^18)";
            return MTH_ERR_UNSUPPORTED;
        }
    if ((params->measures & (MTH_FDRP)) && params->fdrp.max_depth == 0) { g_create_err = "fdrp.max_depth must be >= 1"; return MTH_ERR_INVALID; }
    if ((params->measures & (MTH_QFDRP)) && params->qfdrp.max_depth == 0) { g_create_err = "qfdrp.max_depth must be >= 1"; return MTH_ERR_INVALID; }
    mth_ctx* c = new mth_ctx();
    c->device = device;
    c->prm = *params;
    c->ref_len.assign(ref_len, ref_len + n_ref);
    memset(&c->stats, 0, sizeof(c->stats));
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->own_compute, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_compute, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_stage_free[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_stage_free[1], cudaEventDisableTiming) != cudaSuccess) {
        g_create_err = std::string("CUDA initialisation failed: ") + cudaGetErrorString(cudaGetLastError());
        delete c;
        return MTH_ERR_CUDA;
    }
    c->compute = c->own_compute;
    c->pipeline = getenv("METHEOR_NO_PIPELINE") == nullptr;  // A/B switch: regions strictly one after the other
    *out = c;
    return MTH_OK;
}

int mth_ctx_destroy(mth_ctx* c) {
    if (!c) return MTH_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    DevBuf* devs[] = {&c->a_start, &c->a_end, &c->a_meta, &c->a_off, &c->a_pos, &c->a_rel, &c->a_meth, &c->a_moff, &c->a_flags, &c->bitmap,
                      &c->word_prefix, &c->block_sums, &c->site_pos, &c->scalars, &c->totals, &c->ct_lin, &c->ct_tid, &c->cnt2,
                      &c->scan_scratch, &c->fdrp_scratch, &c->me_lut, &c->lpmd_total, &c->set_dev, &c->set_bitmap, &c->set_kept, &c->set_scan,
                      &c->set_total, &c->set_pos_tmp, &c->set_rel_tmp};
    for (DevBuf* b : devs) dev_free(*b);
    for (auto& set : c->stage)
        for (DevBuf& b : set) dev_free(b);
    dev_free(c->exp_blocks);
    dev_free(c->exp_tot);
    for (int q = 0; q < 2; q++) { dev_free(c->qcnt[q]); dev_free(c->qmixed[q]); dev_free(c->qobs_site[q]); dev_free(c->qobs_vp[q]); dev_free(c->qhrows[q]); dev_free(c->qmix_list[q]); }
    dev_free(c->qobs_n);
    dev_free(c->qmix_n);
    dev_free(c->gfallback);
    for (cudaEvent_t e : c->ev_stage_free)
        if (e) cudaEventDestroy(e);
    for (int m = 0; m < M_COUNT; m++) { dev_free(c->rowcnt[m]); dev_free(c->value[m]); }
    for (SiteRowsBuf* r : {&c->rows_pdr, &c->rows_mhl, &c->rows_fdrp, &c->rows_qfdrp}) {
        dev_free(r->tid); dev_free(r->pos); dev_free(r->value); dev_free(r->nc); dev_free(r->nd);
        host_free(r->h_tid); host_free(r->h_pos); host_free(r->h_value); host_free(r->h_nc); host_free(r->h_nd);
    }
    for (QuartetRowsBuf* r : {&c->rows_pm, &c->rows_me}) {
        dev_free(r->tid); dev_free(r->p1); dev_free(r->p2); dev_free(r->p3); dev_free(r->p4); dev_free(r->value); dev_free(r->counts);
        host_free(r->h_tid); host_free(r->h_p1); host_free(r->h_p2); host_free(r->h_p3); host_free(r->h_p4);
        host_free(r->h_value); host_free(r->h_counts);
    }
    {
        PairRowsBuf& r = c->rows_pairs;
        for (DevBuf* b : {&r.tid, &r.pos1, &r.pos2, &r.lpmd, &r.nc, &r.nd}) dev_free(*b);
        for (HostBuf* b : {&r.h_tid, &r.h_pos1, &r.h_pos2, &r.h_lpmd, &r.h_nc, &r.h_nd}) host_free(*b);
    }
    host_free(c->h_scalars);
    host_free(c->h_totals);
    host_free(c->h_set_total);
    host_free(c->h_ct);
    for (auto& sp : c->spans) { cudaEventDestroy(sp.e0); cudaEventDestroy(sp.e1); }
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    mth_comm_destroy(c);
    if (c->ev_copy) cudaEventDestroy(c->ev_copy);
    if (c->ev_compute) cudaEventDestroy(c->ev_compute);
    for (RegionJob* j : {&c->pend, &c->next}) {
        dev_free(j->st.bitmap); dev_free(j->st.scalars); dev_free(j->st.a_flags);
        host_free(j->st.h_scalars); host_free(j->st.h_totals);
        if (j->k.ev_a) cudaEventDestroy(j->k.ev_a);
        if (j->k.ev_b) cudaEventDestroy(j->k.ev_b);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->own_compute) cudaStreamDestroy(c->own_compute);
    if (c->copy) cudaStreamDestroy(c->copy);
    if (c->aux) cudaStreamDestroy(c->aux);
    delete c;
    return MTH_OK;
}

int mth_set_stream(mth_ctx* c, void* cuda_stream) {
    if (!c) return MTH_ERR_INVALID;
    CUDA_TRY(c, cudaStreamSynchronize(c->compute));
    c->compute = cuda_stream ? (cudaStream_t)cuda_stream : c->own_compute;
    return MTH_OK;
}

int mth_sync(mth_ctx* c) {
    if (!c) return MTH_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaStreamSynchronize(c->copy));
    CUDA_TRY(c, cudaStreamSynchronize(c->compute));
    return MTH_OK;
}

int mth_sync_copies(mth_ctx* c) {
    if (!c) return MTH_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaStreamSynchronize(c->copy));
    return MTH_OK;
}

int mth_reset(mth_ctx* c) {
    if (!c) return MTH_ERR_INVALID;
    TRY(mth_sync(c));
    c->region_active = false;
    c->finished = 0;
    c->R = c->I = c->W = 0;
    c->borrowed = false;
    c->last_tid = -1;
    c->has_pend = false;
    for (RegionJob* j : {&c->pend, &c->next}) {
        j->st.region_active = false;
        j->st.borrowed = false;
        j->st.R = j->st.I = j->st.W = 0;
    }
    c->rows_pdr.n = c->rows_mhl.n = c->rows_fdrp.n = c->rows_qfdrp.n = 0;
    c->rows_pm.n = c->rows_me.n = 0;
    c->rows_pairs.n = 0;
    memset(c->lpmd_total_host, 0, sizeof(c->lpmd_total_host));
    resolve_spans(c);
    memset(&c->stats, 0, sizeof(c->stats));
    c->err.clear();
    return MTH_OK;
}

int mth_add_skipped_reads(mth_ctx* c, int64_t n_reads, int64_t n_reads_mapq_ok) {
    if (!c || n_reads < 0 || n_reads_mapq_ok < 0) return MTH_ERR_INVALID;
    c->lpmd_total_host[0] += n_reads;
    c->lpmd_total_host[1] += n_reads_mapq_ok;
    return MTH_OK;
}

int mth_submit(mth_ctx* c, const mth_batch* b) {
    if (!c || !b) return MTH_ERR_INVALID;
    if (c->finished) return fail(c, MTH_ERR_STATE, "mth_submit after mth_finish (call mth_reset first)");
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (b->n_reads < 0 || b->n_cpg < 0) return fail(c, MTH_ERR_INVALID, "negative batch size");
    if (b->n_reads == 0) return MTH_OK;
    if (b->tid < 0 || (size_t)b->tid >= c->ref_len.size()) return fail(c, MTH_ERR_INVALID, "batch tid outside the reference list");
    if (!b->start || !b->end || !b->meta || !b->cpg_off || (b->n_cpg && !b->cpg_pos) || !b->meth)
        return fail(c, MTH_ERR_INVALID, "null array in batch");
    bool lp = (c->prm.measures & MTH_LPMD) != 0;
    if (lp && b->n_cpg && !b->cpg_rel) return fail(c, MTH_ERR_INVALID, "cpg_rel is required when MTH_LPMD is requested");
    if (b->mem_kind != 0 && b->mem_kind != 1 && b->mem_kind != 2) return fail(c, MTH_ERR_INVALID, "mem_kind must be 0 (host), 1 or 2 (device)");
    if (b->n_reads > (int64_t)INT32_MAX - 64 || b->n_cpg > (int64_t)UINT32_MAX - 64)
        return fail(c, MTH_ERR_UNSUPPORTED, "batch too large (reads < 2^31, CpG calls < 2^32)");
    int64_t bw = batch_words(b);
    if (bw < b->n_reads && b->meth_off == nullptr) return fail(c, MTH_ERR_INVALID, "n_meth_words inconsistent");

    // borrowed arrays are read in place; the TMA bulk copies of k_ingest need 16-byte aligned bases
    auto al16 = [](const void* p) { return ((uintptr_t)p & 15u) == 0; };
    const bool borrowable = b->mem_kind == 1 && al16(b->cpg_pos) && al16(b->cpg_rel) && !c->has_set;

    // A chain of large device-resident contigs (each is a region of its own, see place_batch): the region that ends here is
    // only CLOSED now (its site count is queued); its phase A is queued behind THIS batch's ingest pass and its phase B
    // behind the next one's, so that the host's two waits per region fall while the GPU has an ingest pass to run.
    bool deferred = false;
    if (c->pipeline && c->region_active && c->borrowed && c->R >= BORROWED_REGION_MIN_READS && b->tid > c->last_tid && borrowable) {
        RegionCarry& k = c->next.k;
        if (!k.ev_a && (cudaEventCreateWithFlags(&k.ev_a, cudaEventDisableTiming) != cudaSuccess ||
                        cudaEventCreateWithFlags(&k.ev_b, cudaEventDisableTiming) != cudaSuccess))
            return fail(c, MTH_ERR_CUDA, "cudaEventCreate failed");
        TRY(region_close(c, k));
        swap_slot(c, c->next.st);  // the closed region is parked; the context continues with the spare buffer set
        c->region_active = false;
        c->borrowed = false;
        c->R = c->I = c->W = 0;
        deferred = true;
    }

    TRY(place_batch(c, b->tid, b->n_reads, b->n_cpg, bw));

    const int32_t lin_off = c->cur_lin_off;
    const int64_t r0 = c->R, i0 = c->I, w0 = c->W;
    const bool can_borrow = borrowable && r0 == 0 && lin_off == 0;
    const uint16_t* rel_dev = nullptr;
    if (can_borrow) {
        c->borrowed = true;
        c->bview = *b;
        c->has_meth_off = b->meth_off != nullptr;
        rel_dev = b->cpg_rel;
    } else {
        TRY(materialize(c));
        bool need_moff = c->has_meth_off || b->meth_off != nullptr;
        TRY(ensure_arena(c, r0 + b->n_reads, i0 + b->n_cpg, w0 + bw, lp, need_moff));
        cudaStream_t cs = c->copy;
        // Batch k+1 only writes arena slots that no in-flight kernel of batch k reads: the shared terminal offset
        // a_off[r0] (== i0 after fix-up) is NOT rewritten — the copy skips the batch's own leading 0.
        if (b->mem_kind == 0 && (b->cpg_off[0] != 0 || (b->meth_off && b->meth_off[0] != 0)))
            return fail(c, MTH_ERR_INVALID, "cpg_off[0] / meth_off[0] must be 0");
        cudaMemcpyKind kind = b->mem_kind == 0 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
        size_t nR = (size_t)b->n_reads, nI = (size_t)b->n_cpg;
        size_t skip = r0 ? 1 : 0;
        CUDA_TRY(c, cudaMemcpyAsync((int32_t*)c->a_start.p + r0, b->start, nR * 4, kind, cs));
        CUDA_TRY(c, cudaMemcpyAsync((int32_t*)c->a_end.p + r0, b->end, nR * 4, kind, cs));
        CUDA_TRY(c, cudaMemcpyAsync((uint32_t*)c->a_meta.p + r0, b->meta, nR * 4, kind, cs));
        CUDA_TRY(c, cudaMemcpyAsync((uint32_t*)c->a_off.p + r0 + skip, b->cpg_off + skip, (nR + 1 - skip) * 4, kind, cs));
        if (nI) CUDA_TRY(c, cudaMemcpyAsync((int32_t*)c->a_pos.p + i0, b->cpg_pos, nI * 4, kind, cs));
        CUDA_TRY(c, cudaMemcpyAsync((uint64_t*)c->a_meth.p + w0, b->meth, (size_t)bw * 8, kind, cs));
        if (lp && nI) CUDA_TRY(c, cudaMemcpyAsync((uint16_t*)c->a_rel.p + i0, b->cpg_rel, nI * 2, kind, cs));
        if (b->meth_off)
            CUDA_TRY(c, cudaMemcpyAsync((uint32_t*)c->a_moff.p + r0 + skip, b->meth_off + skip, (nR + 1 - skip) * 4, kind, cs));
        if (b->mem_kind == 0)
            c->stats.h2d_bytes += (int64_t)(nR * 16 + 4 + nI * (lp ? 6 : 4) + (size_t)bw * 8 + (b->meth_off ? (nR + 1) * 4 : 0));
        CUDA_TRY(c, cudaEventRecord(c->ev_copy, cs));
        CUDA_TRY(c, cudaStreamWaitEvent(c->compute, c->ev_copy, 0));
        {
            ProfScope ps(c, "fixup");
            if (i0) ps.add(launch_add_u32((uint32_t*)c->a_off.p + r0 + 1, b->n_reads, (uint32_t)i0, c->compute));
            if (need_moff) {
                // reads ingested before the first batch that carried meth_off had one word each
                if (!c->has_meth_off && r0) ps.add(launch_iota_u32((uint32_t*)c->a_moff.p, r0 + 1, 0, c->compute));
                if (b->meth_off) {
                    if (w0) ps.add(launch_add_u32((uint32_t*)c->a_moff.p + r0 + skip, b->n_reads + 1 - (int64_t)skip, (uint32_t)w0, c->compute));
                } else {
                    ps.add(launch_iota_u32((uint32_t*)c->a_moff.p + r0 + skip, b->n_reads + 1 - (int64_t)skip, (uint32_t)(w0 + skip), c->compute));
                }
                c->has_meth_off = true;
            }
            if (lin_off) {
                ps.add(launch_add_i32((int32_t*)c->a_start.p + r0, b->n_reads, lin_off, c->compute));
                ps.add(launch_add_i32((int32_t*)c->a_end.p + r0, b->n_reads, lin_off, c->compute));
                ps.add(launch_add_i32((int32_t*)c->a_pos.p + i0, b->n_cpg, lin_off, c->compute));
            }
        }
        rel_dev = lp ? (const uint16_t*)c->a_rel.p : nullptr;
    }
    c->R = r0 + b->n_reads;
    c->I = i0 + b->n_cpg;
    c->W = w0 + bw;
    int64_t n_kept = b->n_cpg;
    if (c->has_set) {  // readutil.rs:87-95: the CpG-set filter comes before everything else
        TRY(filter_batch(c, r0, b->n_reads, i0, b->n_cpg, &n_kept));
        c->I = i0 + n_kept;
    }

    TRY(run_ingest(c, b->tid, r0, b->n_reads, i0, n_kept, rel_dev));
    if (deferred) {
        TRY(finish_pending(c));  // phase B of the region before the parked one
        swap_slot(c, c->next.st);
        const int rc = region_phase_a(c, c->next.k);
        swap_slot(c, c->next.st);
        TRY(rc);
        if (!c->next.k.done) {  // `next` becomes the region waiting for phase B; the finished one's buffers are the spare set
            std::swap(c->pend, c->next);
            c->has_pend = true;
        }
    }
    return MTH_OK;
}

int mth_reserve(mth_ctx* c, int64_t n_reads, int64_t n_cpg) {
    if (!c || n_reads < 0 || n_cpg < 0) return MTH_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(c, cudaMemGetInfo(&free_b, &total_b));
    const bool lp = (c->prm.measures & MTH_LPMD) != 0;
    const double per_read = 24.0, per_call = lp ? 7.0 : 5.0;
    double need = per_read * (double)n_reads + per_call * (double)n_cpg;
    const double budget = 0.5 * (double)free_b;
    if (need > budget) {  // keep the proportions, shrink to the budget
        const double f = budget / need;
        n_reads = (int64_t)((double)n_reads * f);
        n_cpg = (int64_t)((double)n_cpg * f);
    }
    n_reads = std::min<int64_t>(n_reads, (int64_t)INT32_MAX - 128);
    n_cpg = std::min<int64_t>(n_cpg, (int64_t)UINT32_MAX - 128);
    if (n_reads <= c->R && n_cpg <= c->I) return MTH_OK;
    const bool dbg = getenv("METHEOR_DEBUG_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    CUDA_TRY(c, cudaStreamSynchronize(c->copy));
    CUDA_TRY(c, cudaStreamSynchronize(c->compute));
    const double t1 = now();
    TRY(ensure_arena(c, std::max(n_reads, c->R), std::max(n_cpg, c->I), std::max(n_reads, c->W), lp, c->has_meth_off));
    TRY(dev_reserve(c, c->a_flags, (size_t)std::max(n_cpg, c->I) + 64, (size_t)c->I));
    if (dbg) fprintf(stderr, "mth_reserve: %lld reads %lld calls: wait for streams %.4f s, reallocation %.4f s\n", (long long)n_reads, (long long)n_cpg, t1 - t0, now() - t1);
    return MTH_OK;
}

int mth_submit_compact(mth_ctx* c, const mth_batch_compact* b) {
    if (!c || !b) return MTH_ERR_INVALID;
    if (c->finished) return fail(c, MTH_ERR_STATE, "mth_submit_compact after mth_finish (call mth_reset first)");
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (b->n_reads < 0 || b->n_cpg < 0 || b->n_rel < 0) return fail(c, MTH_ERR_INVALID, "negative batch size");
    if (b->n_reads == 0) return MTH_OK;
    if (b->tid < 0 || (size_t)b->tid >= c->ref_len.size()) return fail(c, MTH_ERR_INVALID, "batch tid outside the reference list");
    const bool enc_s16 = (b->enc & MTH_CENC_START16) != 0, enc_d8 = (b->enc & MTH_CENC_DELTA8) != 0;
    if (b->enc & ~(MTH_CENC_START16 | MTH_CENC_DELTA8)) return fail(c, MTH_ERR_INVALID, "unknown bits in mth_batch_compact.enc");
    if ((!enc_s16 && !b->start) || !b->span || !b->mapq || !b->n_cpg8 || !b->flags || (b->n_cpg && !b->meth_bits) ||
        (b->n_cpg && !enc_d8 && !b->cpg_delta))
        return fail(c, MTH_ERR_INVALID, "null array in compact batch");
    if (enc_s16 && (!b->start_off16 || !b->blk_start || (b->n_start_exc && !b->start_exc) || b->n_start_exc < 0))
        return fail(c, MTH_ERR_INVALID, "MTH_CENC_START16 needs start_off16, blk_start (and start_exc)");
    if (enc_d8 && (!b->blk_call_off || (b->n_delta8 && !b->cpg_delta8) || (b->n_delta16 && !b->cpg_delta) || b->n_delta8 < 0 ||
                   b->n_delta16 < 0 || b->n_delta8 + b->n_delta16 != b->n_cpg))
        return fail(c, MTH_ERR_INVALID, "MTH_CENC_DELTA8 needs blk_call_off, cpg_delta8 / cpg_delta with n_delta8 + n_delta16 == n_cpg");
    const bool lp = (c->prm.measures & MTH_LPMD) != 0;
    if (lp && b->n_rel && !b->rel_exc) return fail(c, MTH_ERR_INVALID, "rel_exc is required when n_rel > 0");
    if (b->mem_kind != 0 && b->mem_kind != 1) return fail(c, MTH_ERR_INVALID, "mem_kind must be 0 (host) or 1 (device)");
    if (b->n_reads > (int64_t)INT32_MAX - 64 || b->n_cpg > (int64_t)UINT32_MAX - 64)
        return fail(c, MTH_ERR_UNSUPPORTED, "batch too large (reads < 2^31, CpG calls < 2^32)");
    TRY(place_batch(c, b->tid, b->n_reads, b->n_cpg, b->n_reads));
    TRY(materialize(c));
    TRY(ensure_arena(c, c->R + b->n_reads, c->I + b->n_cpg, c->W + b->n_reads, lp, c->has_meth_off));
    TRY(dev_reserve(c, c->a_flags, (size_t)(c->I + b->n_cpg) + 64, (size_t)c->I));

    // Host batches go over in pieces of COMPACT_PIECE reads through two staging sets, so that the H2D copy of piece k+1
    // overlaps the expansion + ingest of piece k.  Device batches are expanded in place in one go.
    // (the dense encodings address their arrays by block of the whole batch: no slicing — hand them over in batches of a
    // few million reads, as the streaming host does anyway)
    const int64_t PIECE = (b->mem_kind == 0 && b->enc == 0) ? COMPACT_PIECE : b->n_reads;
    int64_t x0 = 0, e0 = 0;  // calls / explicit query indices before the piece
    for (int64_t ra = 0; ra < b->n_reads; ra += PIECE) {
        const int64_t rb = std::min(b->n_reads, ra + PIECE);
        int64_t nI = b->n_cpg, nE = lp ? b->n_rel : 0;
        if (PIECE < b->n_reads) {  // host memory: count this piece's calls
            uint64_t si = 0, se = 0;
            sum_call_counts(b->n_cpg8 + ra, (lp && b->n_rel) ? b->flags + ra : nullptr, rb - ra, &si, &se);
            nI = (int64_t)si;
            nE = (int64_t)se;
            if (x0 + nI > b->n_cpg || e0 + nE > b->n_rel) return fail(c, MTH_ERR_INVALID, "n_cpg8 / flags inconsistent with n_cpg / n_rel");
        }
        const int64_t r0 = c->R, i0 = c->I, w0 = c->W;
        const size_t nR = (size_t)(rb - ra);
        const int bit_base = (int)(x0 & 7);
        ExpandArgs ea;
        memset(&ea, 0, sizeof(ea));
        const size_t nblk = (nR + 255) / 256;
        const bool s16 = (b->enc & MTH_CENC_START16) != 0, d8 = (b->enc & MTH_CENC_DELTA8) != 0;
        // the arrays of this piece: {source, bytes}; order fixed, absent ones have 0 bytes
        const void* src[12] = {s16 ? nullptr : (const void*)(b->start + ra), b->span + ra, b->mapq + ra, b->n_cpg8 + ra, b->flags + ra,
                               d8 ? (const void*)b->cpg_delta : (const void*)(b->cpg_delta + x0), b->meth_bits + (x0 >> 3),
                               b->rel_exc ? b->rel_exc + e0 : nullptr, b->start_off16, b->blk_start, b->start_exc, b->cpg_delta8};
        const size_t sz[12] = {s16 ? 0 : nR * 4, nR * 2, nR, nR, nR, d8 ? (size_t)b->n_delta16 * 2 : (size_t)nI * 2,
                               ((size_t)nI + bit_base + 7) / 8, (size_t)nE * 2, s16 ? nR * 2 : 0, s16 ? nblk * 4 : 0,
                               s16 ? (size_t)b->n_start_exc * 4 : 0, d8 ? (size_t)b->n_delta8 : 0};
        const void* ptr[13];  // device addresses the expansion reads (12 arrays + blk_call_off)
        if (b->mem_kind == 0) {
            const int set = c->stage_next;
            c->stage_next ^= 1;
            DevBuf* st = c->stage[set];
            for (int k = 0; k < 12; k++) TRY(dev_reserve(c, st[k], sz[k] + 64, 0));
            TRY(dev_reserve(c, st[12], (d8 ? nblk * 4 : 0) + 64, 0));
            if (c->stage_busy[set]) CUDA_TRY(c, cudaStreamWaitEvent(c->copy, c->ev_stage_free[set], 0));
            for (int k = 0; k < 12; k++) {
                if (sz[k] && src[k]) CUDA_TRY(c, cudaMemcpyAsync(st[k].p, src[k], sz[k], cudaMemcpyHostToDevice, c->copy));
                c->stats.h2d_bytes += (int64_t)sz[k];
                ptr[k] = st[k].p;
            }
            if (d8) {
                CUDA_TRY(c, cudaMemcpyAsync(st[12].p, b->blk_call_off, nblk * 4, cudaMemcpyHostToDevice, c->copy));
                c->stats.h2d_bytes += (int64_t)nblk * 4;
            }
            ptr[12] = st[12].p;
            CUDA_TRY(c, cudaEventRecord(c->ev_copy, c->copy));
            CUDA_TRY(c, cudaStreamWaitEvent(c->compute, c->ev_copy, 0));
            ea.bit_base = (uint32_t)bit_base;
        } else {
            for (int k = 0; k < 12; k++) ptr[k] = src[k];
            ptr[0] = b->start; ptr[1] = b->span; ptr[2] = b->mapq; ptr[3] = b->n_cpg8; ptr[4] = b->flags; ptr[5] = b->cpg_delta;
            ptr[6] = b->meth_bits; ptr[7] = b->rel_exc;
            ptr[12] = b->blk_call_off;
        }
        ea.start = (const int32_t*)ptr[0]; ea.span = (const uint16_t*)ptr[1]; ea.mapq = (const uint8_t*)ptr[2];
        ea.n_cpg8 = (const uint8_t*)ptr[3]; ea.flags = (const uint8_t*)ptr[4]; ea.cpg_delta = (const uint16_t*)ptr[5];
        ea.meth_bits = (const uint8_t*)ptr[6]; ea.rel_exc = (const uint16_t*)ptr[7];
        ea.enc = b->enc;
        ea.n_calls = nI; ea.n_delta8 = b->n_delta8; ea.n_delta16 = b->n_delta16; ea.n_rel = nE; ea.n_start_exc = b->n_start_exc;
        ea.start_off16 = (const uint16_t*)ptr[8]; ea.blk_start = (const int32_t*)ptr[9]; ea.start_exc = (const int32_t*)ptr[10];
        ea.cpg_delta8 = (const uint8_t*)ptr[11]; ea.blk_call_off = (const uint32_t*)ptr[12];
        const size_t nb = (nR + 255) / 256;
        TRY(dev_reserve(c, c->exp_blocks, nb * 8 + 64, 0));
        TRY(dev_reserve(c, c->exp_tot, 16, 0));
        ea.n = (int64_t)nR;
        ea.r0 = r0; ea.i0 = i0; ea.w0 = w0;
        ea.lin_off = c->cur_lin_off;
        ea.start_out = (int32_t*)c->a_start.p; ea.end_out = (int32_t*)c->a_end.p; ea.meta_out = (uint32_t*)c->a_meta.p;
        ea.off_out = (uint32_t*)c->a_off.p; ea.pos_out = (int32_t*)c->a_pos.p; ea.rel_out = lp ? (uint16_t*)c->a_rel.p : nullptr;
        ea.meth_out = (uint64_t*)c->a_meth.p; ea.moff_out = c->has_meth_off ? (uint32_t*)c->a_moff.p : nullptr;
        ea.block_calls = (uint32_t*)c->exp_blocks.p; ea.block_rel = (uint32_t*)c->exp_blocks.p + nb;
        ea.err = &((RegionScalars*)c->scalars.p)->err;
        {
            ProfScope ps(c, "k_expand");
            ps.add(launch_expand(ea, (unsigned long long*)c->exp_tot.p, c->compute));
        }
        if (b->mem_kind == 0) {
            const int set = c->stage_next ^ 1;
            CUDA_TRY(c, cudaEventRecord(c->ev_stage_free[set], c->compute));
            c->stage_busy[set] = true;
        }
        c->R = r0 + (int64_t)nR;
        c->I = i0 + nI;
        c->W = w0 + (int64_t)nR;
        int64_t n_kept = nI;
        if (c->has_set) {
            TRY(filter_batch(c, r0, (int64_t)nR, i0, nI, &n_kept));
            c->I = i0 + n_kept;
        }
        TRY(run_ingest(c, b->tid, r0, (int64_t)nR, i0, n_kept, lp ? (const uint16_t*)c->a_rel.p : nullptr));
        x0 += nI;
        e0 += nE;
    }
    return MTH_OK;
}

}  // extern "C"

// ---- per-region processing -------------------------------------------------------------------------
static int reserve_site_rows(mth_ctx* c, SiteRowsBuf& r, int64_t n, bool counts) {
    TRY(dev_reserve(c, r.tid, (size_t)n * 4 + 4, (size_t)r.n * 4));
    TRY(dev_reserve(c, r.pos, (size_t)n * 4 + 4, (size_t)r.n * 4));
    TRY(dev_reserve(c, r.value, (size_t)n * 4 + 4, (size_t)r.n * 4));
    if (counts) {
        TRY(dev_reserve(c, r.nc, (size_t)n * 4 + 4, (size_t)r.n * 4));
        TRY(dev_reserve(c, r.nd, (size_t)n * 4 + 4, (size_t)r.n * 4));
    }
    return MTH_OK;
}
static int reserve_quartet_rows(mth_ctx* c, QuartetRowsBuf& r, int64_t n, bool counts) {
    for (DevBuf* b : {&r.tid, &r.p1, &r.p2, &r.p3, &r.p4, &r.value}) TRY(dev_reserve(c, *b, (size_t)n * 4 + 4, (size_t)r.n * 4));
    if (counts) TRY(dev_reserve(c, r.counts, (size_t)n * 64 + 64, (size_t)r.n * 64));
    return MTH_OK;
}
static SiteRowsDev site_rows_dev(SiteRowsBuf& r) {
    return SiteRowsDev{(int32_t*)r.tid.p, (int32_t*)r.pos.p, (float*)r.value.p, (uint32_t*)r.nc.p, (uint32_t*)r.nd.p};
}
static QuartetRowsDev quartet_rows_dev(QuartetRowsBuf& r) {
    return QuartetRowsDev{(int32_t*)r.tid.p, (int32_t*)r.p1.p, (int32_t*)r.p2.p, (int32_t*)r.p3.p, (int32_t*)r.p4.p,
                          (float*)r.value.p, (uint32_t*)r.counts.p};
}

static int build_me_lut(mth_ctx* c) {
    if (c->me_lut.p) return MTH_OK;
    // lut[t*(t+1)/2 + k] = p * log2f(p), p = k as f32 / t as f32 (me.rs:47-50), from the HOST libm so that the
    // device result equals what the reference binary prints on this machine; totals > LUT_MAX use k_quartet's own log2.
    const int LUT_MAX = 1024;
    size_t n = (size_t)(LUT_MAX + 1) * (LUT_MAX + 2) / 2;
    std::vector<float> lut(n, 0.f);
    for (int t = 1; t <= LUT_MAX; t++)
        for (int k = 1; k <= t; k++) {
            volatile float p = (float)k / (float)t;
            volatile float lg = log2f(p);
            volatile float v = p * lg;
            lut[(size_t)t * (t + 1) / 2 + k] = v;
        }
    TRY(dev_reserve(c, c->me_lut, n * 4, 0));
    CUDA_TRY(c, cudaMemcpy(c->me_lut.p, lut.data(), n * 4, cudaMemcpyHostToDevice));
    c->me_lut_max = LUT_MAX;
    return MTH_OK;
}

// Fork / join of the side stream: work queued on the returned stream starts after everything queued on `s` so far and `s`
// continues after side_stream_end only when it is done.  Under MTH_FLAG_PROFILE the side stream IS `s` (serial, so that the
// per-kernel event times mean what they say).
static cudaStream_t side_stream_begin(mth_ctx* c, cudaStream_t s) {
    if (c->prm.flags & MTH_FLAG_PROFILE) return s;
    if (cudaEventRecord(c->ev_fork, s) != cudaSuccess || cudaStreamWaitEvent(c->aux, c->ev_fork, 0) != cudaSuccess) return s;
    return c->aux;
}
static int side_stream_end(mth_ctx* c, cudaStream_t s, cudaStream_t side) {
    if (side == s) return MTH_OK;
    CUDA_TRY(c, cudaEventRecord(c->ev_join, side));
    CUDA_TRY(c, cudaStreamWaitEvent(s, c->ev_join, 0));
    return MTH_OK;
}

static void swap_slot(mth_ctx* c, RegionSlot& o) {
    std::swap(c->region_active, o.region_active);
    std::swap(c->cur_lin_off, o.cur_lin_off);
    std::swap(c->R, o.R); std::swap(c->I, o.I); std::swap(c->W, o.W);
    std::swap(c->borrowed, o.borrowed);
    std::swap(c->bview, o.bview);
    std::swap(c->has_meth_off, o.has_meth_off);
    c->reg_lin_off.swap(o.reg_lin_off);
    c->reg_tid.swap(o.reg_tid);
    std::swap(c->bitmap, o.bitmap); std::swap(c->scalars, o.scalars); std::swap(c->a_flags, o.a_flags);
    std::swap(c->bitmap_words_valid, o.bitmap_words_valid);
    std::swap(c->h_scalars, o.h_scalars); std::swap(c->h_totals, o.h_totals);
}

// The processing of a region is three steps with a host round trip between them (the host sizes the per-site buffers from the
// number of sites, and the row buffers from the row totals):
//   region_close   : site count of the region's bitmap + its scalars on their way to the host
//   region_phase_a : [wait] site dictionary, the measure kernels, row counts and their scans; totals on their way to the host
//   region_phase_b : [wait] row emission
// process_region runs them back to back.  For a chain of large device-resident contigs mth_submit interleaves them with the
// NEXT regions' ingest passes, so that the GPU has work queued while the host waits (the waits are on events then).
static bool region_needs_sites(const mth_ctx* c) {
    const uint32_t M = c->prm.measures;
    return (M & (MTH_PDR | MTH_MHL | MTH_PM | MTH_ME | MTH_FDRP | MTH_QFDRP)) != 0 || ((M & MTH_LPMD) && c->prm.lpmd.want_pairs);
}

static int region_close(mth_ctx* c, RegionCarry& k) {
    cudaStream_t s = c->compute;
    RegionScalars* d_sc = (RegionScalars*)c->scalars.p;
    k.C = 0; k.done = false; k.q_set[0] = 0; k.q_set[1] = 1; k.nct = 0;
    const int64_t n_words = (int64_t)c->bitmap_words_valid;
    const int64_t nb = (n_words + 1023) / 1024 + 1;
    if (region_needs_sites(c)) {
        TRY(dev_reserve(c, c->block_sums, (size_t)nb * 4, 0));
        TRY(dev_reserve(c, c->word_prefix, (size_t)n_words * 4 + 4, 0));
        ProfScope ps(c, "k_sites_count");
        ps.add(launch_sites_count((const unsigned long long*)c->bitmap.p, n_words, (uint32_t*)c->block_sums.p, d_sc, s));
    }
    CUDA_TRY(c, cudaMemcpyAsync(c->h_scalars.p, d_sc, sizeof(RegionScalars), cudaMemcpyDeviceToHost, s));
    if (k.ev_a) CUDA_TRY(c, cudaEventRecord(k.ev_a, s));
    return MTH_OK;
}

static int region_end(mth_ctx* c, RegionCarry& k) {
    cudaStream_t s = c->compute;
    k.done = true;
    c->region_active = false;
    c->borrowed = false;
    c->R = c->I = c->W = 0;
    // The emit kernels queued above still read the arena; the next region's host->device copies (copy stream) overwrite it
    // from offset 0: order them behind this point.
    CUDA_TRY(c, cudaEventRecord(c->ev_compute, s));
    CUDA_TRY(c, cudaStreamWaitEvent(c->copy, c->ev_compute, 0));
    return MTH_OK;
}

static int region_phase_a(mth_ctx* c, RegionCarry& k) {
    cudaStream_t s = c->compute;
    const uint32_t M = c->prm.measures;
    ReadsView rv = make_view(c);
    RegionScalars* d_sc = (RegionScalars*)c->scalars.p;
    const bool want_pairs = (M & MTH_LPMD) && c->prm.lpmd.want_pairs;
    const bool need_sites = region_needs_sites(c);
    const int64_t n_words = (int64_t)c->bitmap_words_valid;
    if (k.ev_a) CUDA_TRY(c, cudaEventSynchronize(k.ev_a));
    else CUDA_TRY(c, cudaStreamSynchronize(s));
    RegionScalars sc = *(RegionScalars*)c->h_scalars.p;
    TRY(check_scalars_err(c, sc.err));
    for (int k = 0; k < 4; k++) c->lpmd_total_host[k] += (int64_t)sc.lpmd[k];
    if (sc.lmax > c->stats.max_ref_span) c->stats.max_ref_span = sc.lmax;
    const int64_t C = need_sites ? (int64_t)sc.n_sites : 0;
    c->stats.n_sites += C;
    c->stats.n_regions += 1;
    k.C = C;
    if (C <= 0) return region_end(c, k);

    {
        TRY(dev_reserve(c, c->site_pos, (size_t)C * 4, 0));
        {
            ProfScope ps(c, "k_sites_emit");
            ps.add(launch_sites_emit((const unsigned long long*)c->bitmap.p, n_words, (const uint32_t*)c->block_sums.p,
                                     (uint32_t*)c->word_prefix.p, (int32_t*)c->site_pos.p, s));
        }
        // contig table
        size_t nct = c->reg_tid.size();
        // one copy from pinned memory ([lin_off | tid]); the buffer is free again: the previous region's copy ran before its syncs
        TRY(dev_reserve(c, c->ct_lin, nct * 8, 0));
        TRY(host_reserve(c, c->h_ct, nct * 8));
        memcpy(c->h_ct.p, c->reg_lin_off.data(), nct * 4);
        memcpy((char*)c->h_ct.p + nct * 4, c->reg_tid.data(), nct * 4);
        CUDA_TRY(c, cudaMemcpyAsync(c->ct_lin.p, c->h_ct.p, nct * 8, cudaMemcpyHostToDevice, s));
        ContigTable ct{(int32_t)nct, (const int32_t*)c->ct_lin.p, (const int32_t*)c->ct_lin.p + nct};
        k.nct = nct;
        const int32_t* site_pos = (const int32_t*)c->site_pos.p;

        TRY(dev_reserve(c, c->gfallback, (size_t)C + 64, 0));
        TRY(dev_reserve(c, c->totals, 8 * M_COUNT, 0));
        TRY(host_reserve(c, c->h_totals, 8 * M_COUNT));
        CUDA_TRY(c, cudaMemsetAsync(c->totals.p, 0, 8 * M_COUNT, s));
        unsigned long long* d_tot = (unsigned long long*)c->totals.p;
        TRY(dev_reserve(c, c->scan_scratch, (size_t)((C + 2047) / 2048 + 2) * 4, 0));
        uint32_t* scratch = (uint32_t*)c->scan_scratch.p;
        for (int m = 0; m < M_COUNT; m++) {
            static const uint32_t bit[M_COUNT] = {MTH_PDR, MTH_MHL, MTH_FDRP, MTH_QFDRP, MTH_PM, MTH_ME, 0};
            if ((M & bit[m]) || (m == M_PAIRS && want_pairs)) TRY(dev_reserve(c, c->rowcnt[m], (size_t)C * 4 + 4, 0));
            if ((M & bit[m]) && (m == M_MHL || m == M_FDRP || m == M_QFDRP)) TRY(dev_reserve(c, c->value[m], (size_t)C * 4 + 4, 0));
        }

        // ---------------- phase A: measure kernels + row counts + scans ----------------
        if (M & MTH_PDR) {
            TRY(dev_reserve(c, c->cnt2, (size_t)C * 8 + 8, 0));
            // 1 = scatter only (no read spans > 150 bases: no flush possible), 2 = gather everywhere (testing flag),
            // 3 = scatter + segment-exact gather on the hazard sites only
            const bool force = (c->prm.flags & MTH_FLAG_FORCE_GATHER) != 0;
            c->stats.pdr_path = force ? 2 : (sc.lmax > 150 ? 3 : 1);
            if (force) {
                ProfScope ps(c, "k_pdr_gather");
                ps.add(launch_pdr_gather(rv, site_pos, C, d_sc, (uint32_t*)c->cnt2.p, c->prm.pdr, nullptr, s));
            } else {
                CUDA_TRY(c, cudaMemsetAsync(c->cnt2.p, 0, (size_t)C * 8, s));
                {
                    ProfScope ps(c, "k_pdr_scatter");
                    ps.add(launch_pdr_scatter(rv.cpg_pos, (const uint8_t*)c->a_flags.p, rv.I, (const unsigned long long*)c->bitmap.p, n_words,
                                              (const uint32_t*)c->word_prefix.p, d_sc, (uint32_t*)c->cnt2.p, s));
                }
                if (sc.lmax > 150) {
                    ProfScope ps(c, "k_pdr_gather");
                    CUDA_TRY(c, cudaMemsetAsync(c->gfallback.p, 0, (size_t)C, s));
                    ps.add(launch_pdr_hazard(rv, (const unsigned long long*)c->bitmap.p, (const uint32_t*)c->word_prefix.p,
                                             (uint8_t*)c->gfallback.p, s));
                    ps.add(launch_pdr_gather(rv, site_pos, C, d_sc, (uint32_t*)c->cnt2.p, c->prm.pdr, (const uint8_t*)c->gfallback.p, s));
                }
            }
            ProfScope ps(c, "pdr_rows_count");
            ps.add(launch_pdr_rowcnt((const uint32_t*)c->cnt2.p, C, c->prm.pdr.min_depth, (uint32_t*)c->rowcnt[M_PDR].p, s));
            ps.add(launch_exclusive_scan_u32((uint32_t*)c->rowcnt[M_PDR].p, C, scratch, d_tot + M_PDR, s));
        }
        if (M & MTH_MHL) {
            {
                ProfScope ps(c, "k_mhl");
                CUDA_TRY(c, cudaMemsetAsync(c->gfallback.p, 0, (size_t)C, s));
                {
                    ProfScope p2(c, "~mhl_tile");
                    p2.add0(launch_mhl_site(rv, site_pos, C, d_sc, c->prm.mhl, (float*)c->value[M_MHL].p, (uint32_t*)c->rowcnt[M_MHL].p,
                                            (uint8_t*)c->gfallback.p, s));
                }
                ps.add(1);
                if (c->prm.flags & MTH_FLAG_PROFILE) ps.add(launch_count_flags((const uint8_t*)c->gfallback.p, C, &d_sc->fallback_sites[0], s));
                {
                    ProfScope p2(c, "~mhl_fallback");
                    p2.add0(launch_mhl(rv, site_pos, C, d_sc, c->prm.mhl, (float*)c->value[M_MHL].p, (uint32_t*)c->rowcnt[M_MHL].p,
                                       (const uint8_t*)c->gfallback.p, s));
                }
                ps.add(1);
            }
            ProfScope ps(c, "mhl_rows_count");
            ps.add(launch_exclusive_scan_u32((uint32_t*)c->rowcnt[M_MHL].p, C, scratch, d_tot + M_MHL, s));
        }
        {
            // FDRP and qFDRP share the pile, the overlap test and the Hamming distance (fdrp.rs:124-145, qfdrp.rs:137-157): with
            // equal thresholds (the reference's defaults are equal) ONE pile build + ONE pair loop emits both values.
            const mth_fdrp_params &pf = c->prm.fdrp, &pq = c->prm.qfdrp;
            const bool both = (M & MTH_FDRP) && (M & MTH_QFDRP) && pf.min_qual == pq.min_qual && pf.min_depth == pq.min_depth &&
                              pf.max_depth == pq.max_depth && pf.min_overlap == pq.min_overlap;
            for (int q = 0; q < 2; q++) {
                uint32_t bit = q ? MTH_QFDRP : MTH_FDRP;
                int m = q ? M_QFDRP : M_FDRP;
                if (!(M & bit)) continue;
                if (both && q == 1) {  // rows of qFDRP: value / rowcnt were written by the fused pass
                    ProfScope ps(c, "qfdrp_rows_count");
                    ps.add(launch_exclusive_scan_u32((uint32_t*)c->rowcnt[m].p, C, scratch, d_tot + m, s));
                    continue;
                }
                mth_fdrp_params fp = q ? c->prm.qfdrp : c->prm.fdrp;
                const int mode = both ? 2 : q;
                size_t sb = fdrp_scratch_bytes(fp, q);
                TRY(dev_reserve(c, c->fdrp_scratch, sb, 0));
                {
                    ProfScope ps(c, both ? "k_fdrp_qfdrp" : (q ? "k_qfdrp" : "k_fdrp"));
                    CUDA_TRY(c, cudaMemsetAsync(c->gfallback.p, 0, (size_t)C, s));
                    float* vq = both ? (float*)c->value[M_QFDRP].p : nullptr;
                    uint32_t* rq = both ? (uint32_t*)c->rowcnt[M_QFDRP].p : nullptr;
                    {
                        ProfScope p2(c, "~fdrp_tile");
                        p2.add0(launch_fdrp_tile(rv, site_pos, C, (const unsigned long long*)c->bitmap.p, n_words, (const uint32_t*)c->word_prefix.p,
                                                 d_sc, fp, mode, c->prm.seed, ct, (float*)c->value[m].p, (uint32_t*)c->rowcnt[m].p, vq, rq,
                                                 (uint8_t*)c->gfallback.p, s));
                    }
                    ps.add(1);
                    if (c->prm.flags & MTH_FLAG_PROFILE) ps.add(launch_count_flags((const uint8_t*)c->gfallback.p, C, &d_sc->fallback_sites[1], s));
                    {
                        ProfScope p2(c, "~fdrp_fallback");
                        p2.add0(launch_fdrp(rv, site_pos, C, d_sc, fp, mode, c->prm.seed, ct, c->fdrp_scratch.p, sb, (float*)c->value[m].p,
                                            (uint32_t*)c->rowcnt[m].p, vq, rq, (const uint8_t*)c->gfallback.p, s));
                    }
                    ps.add(1);
                }
                ProfScope ps(c, q ? "qfdrp_rows_count" : "fdrp_rows_count");
                ps.add(launch_exclusive_scan_u32((uint32_t*)c->rowcnt[m].p, C, scratch, d_tot + m, s));
            }
        }
        int* const q_set = k.q_set;  // which histogram / mixed-site set PM and ME use
        for (int q = 0; q < 2; q++) {
            uint32_t bit = q ? MTH_ME : MTH_PM;
            int m = q ? M_ME : M_PM;
            if (!(M & bit)) continue;
            mth_quartet_params qp = q ? c->prm.me : c->prm.pm;
            // PM and ME with identical thresholds share the histogram and the count pass
            if (q == 1 && (M & MTH_PM) && c->prm.pm.min_depth == qp.min_depth && c->prm.pm.min_qual == qp.min_qual) {
                CUDA_TRY(c, cudaMemcpyAsync(c->rowcnt[M_ME].p, c->rowcnt[M_PM].p, (size_t)C * 4, cudaMemcpyDeviceToDevice, s));
                CUDA_TRY(c, cudaMemcpyAsync(d_tot + M_ME, d_tot + M_PM, 8, cudaMemcpyDeviceToDevice, s));
                q_set[1] = 0;
                continue;
            }
            q_set[q] = q;
            TRY(dev_reserve(c, c->qcnt[q], (size_t)(C + 4) * 16, 0));  // [site][4 slots] u32
            TRY(dev_reserve(c, c->qmixed[q], (size_t)C + 64, 0));
            TRY(dev_reserve(c, c->qobs_site[q], (size_t)rv.I * 4 + 64, 0));  // at most one observation per call
            TRY(dev_reserve(c, c->qobs_vp[q], (size_t)rv.I + 64, 0));
            TRY(dev_reserve(c, c->qobs_n, 16, 0));
            TRY(dev_reserve(c, c->qmix_list[q], (size_t)C * 4 + 64, 0));
            TRY(dev_reserve(c, c->qmix_n, 16, 0));
            CUDA_TRY(c, cudaMemsetAsync(c->qcnt[q].p, 0, (size_t)(C + 4) * 16, s));
            CUDA_TRY(c, cudaMemsetAsync(c->qmixed[q].p, 0, (size_t)C + 64, s));
            CUDA_TRY(c, cudaMemsetAsync((unsigned long long*)c->qobs_n.p + q, 0, 8, s));
            {
                ProfScope ps(c, q ? "k_me_scatter" : "k_pm_scatter");
                ps.add(launch_quartet_scatter(rv.cpg_pos, (const uint8_t*)c->a_flags.p, rv.I, (const unsigned long long*)c->bitmap.p, n_words,
                                              (const uint32_t*)c->word_prefix.p, d_sc, q ? CF_ME_OK : CF_PM_OK, (uint32_t*)c->qcnt[q].p,
                                              (uint8_t*)c->qmixed[q].p, (uint32_t*)c->qobs_site[q].p, (uint8_t*)c->qobs_vp[q].p,
                                              (unsigned long long*)c->qobs_n.p + q, s));
            }
            {
                // The gather over the mixed sites is a handful of sites, each a chain of dependent loads: a launch lasts as long as
                // one site takes.  It writes rowcnt[] of the mixed sites only, the canonical count those of the others: the two run
                // side by side (not under MTH_FLAG_PROFILE, whose per-kernel times want them one after the other).
                ProfScope ps(c, q ? "k_me_count" : "k_pm_count");
                cudaStream_t gs = side_stream_begin(c, s);
                ps.add(launch_quartet_canon_count((const uint32_t*)c->qcnt[q].p, (const uint8_t*)c->qmixed[q].p, C, qp.min_depth,
                                                  (uint32_t*)c->rowcnt[m].p, s));
                CUDA_TRY(c, cudaMemsetAsync((unsigned long long*)c->qmix_n.p + q, 0, 8, gs));
                ps.add(launch_quartet_mixed_list((const uint8_t*)c->qmixed[q].p, C, (uint32_t*)c->qmix_list[q].p, (unsigned long long*)c->qmix_n.p + q, gs));
                ps.add(launch_quartet_count(rv, site_pos, C, d_sc, qp, (const uint8_t*)c->qmixed[q].p, (const uint32_t*)c->qmix_list[q].p,
                                            (const unsigned long long*)c->qmix_n.p + q, (uint32_t*)c->rowcnt[m].p, gs));
                TRY(side_stream_end(c, s, gs));
            }
            ProfScope ps(c, q ? "me_rows_count" : "pm_rows_count");
            ps.add(launch_exclusive_scan_u32((uint32_t*)c->rowcnt[m].p, C, scratch, d_tot + m, s));
        }
        if (M & MTH_ME) TRY(build_me_lut(c));
        const uint16_t* rel_all = c->borrowed ? c->bview.cpg_rel : (const uint16_t*)c->a_rel.p;
        if (want_pairs) {
            {
                ProfScope ps(c, "k_lpmd_pairs_count");
                ps.add(launch_lpmd_pairs_count(rv, rel_all, site_pos, C, d_sc, c->prm.lpmd, (uint32_t*)c->rowcnt[M_PAIRS].p, s));
            }
            ProfScope ps(c, "pairs_rows_count");
            ps.add(launch_exclusive_scan_u32((uint32_t*)c->rowcnt[M_PAIRS].p, C, scratch, d_tot + M_PAIRS, s));
        }

        CUDA_TRY(c, cudaMemcpyAsync(c->h_totals.p, d_tot, 8 * M_COUNT, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(c, cudaMemcpyAsync(c->h_scalars.p, d_sc, sizeof(RegionScalars), cudaMemcpyDeviceToHost, s));
        if (k.ev_b) CUDA_TRY(c, cudaEventRecord(k.ev_b, s));
        CUDA_TRY(c, cudaGetLastError());
    }
    return MTH_OK;
}

static int region_phase_b(mth_ctx* c, RegionCarry& k) {
    if (k.done) return MTH_OK;
    cudaStream_t s = c->compute;
    const uint32_t M = c->prm.measures;
    ReadsView rv = make_view(c);
    RegionScalars* d_sc = (RegionScalars*)c->scalars.p;
    const bool want_pairs = (M & MTH_LPMD) && c->prm.lpmd.want_pairs;
    const int64_t C = k.C;
    const int* const q_set = k.q_set;
    const int32_t* site_pos = (const int32_t*)c->site_pos.p;
    ContigTable ct{(int32_t)k.nct, (const int32_t*)c->ct_lin.p, (const int32_t*)c->ct_lin.p + k.nct};
    const uint16_t* rel_all = c->borrowed ? c->bview.cpg_rel : (const uint16_t*)c->a_rel.p;
    {
        if (k.ev_b) CUDA_TRY(c, cudaEventSynchronize(k.ev_b));
        else CUDA_TRY(c, cudaStreamSynchronize(s));
        TRY(check_scalars_err(c, ((RegionScalars*)c->h_scalars.p)->err));
        c->stats.fdrp_pair_ops += (int64_t)((RegionScalars*)c->h_scalars.p)->fdrp_pairs;
        c->stats.fallback_sites_mhl += (int64_t)((RegionScalars*)c->h_scalars.p)->fallback_sites[0];
        c->stats.fallback_sites_fdrp += (int64_t)((RegionScalars*)c->h_scalars.p)->fallback_sites[1];
        const unsigned long long* tot = (const unsigned long long*)c->h_totals.p;

        // ---------------- phase B: row emission ----------------
        if (M & MTH_PDR) {
            TRY(reserve_site_rows(c, c->rows_pdr, c->rows_pdr.n + (int64_t)tot[M_PDR], true));
            ProfScope ps(c, "k_pdr_emit");
            ps.add(launch_pdr_emit((const uint32_t*)c->cnt2.p, (const uint32_t*)c->rowcnt[M_PDR].p, site_pos, C, c->prm.pdr.min_depth,
                                   ct, site_rows_dev(c->rows_pdr), c->rows_pdr.n, s));
            c->rows_pdr.n += (int64_t)tot[M_PDR];
        }
        struct { uint32_t bit; int m; SiteRowsBuf* r; const char* name; } simple[3] = {
            {MTH_MHL, M_MHL, &c->rows_mhl, "k_mhl_emit"}, {MTH_FDRP, M_FDRP, &c->rows_fdrp, "k_fdrp_emit"},
            {MTH_QFDRP, M_QFDRP, &c->rows_qfdrp, "k_qfdrp_emit"}};
        for (auto& e : simple) {
            if (!(M & e.bit)) continue;
            TRY(reserve_site_rows(c, *e.r, e.r->n + (int64_t)tot[e.m], false));
            ProfScope ps(c, e.name);
            ps.add(launch_site_emit((const float*)c->value[e.m].p, (const uint32_t*)c->rowcnt[e.m].p, tot[e.m], site_pos, C, ct,
                                    site_rows_dev(*e.r), e.r->n, s));
            e.r->n += (int64_t)tot[e.m];
        }
        {
            const bool counts = (c->prm.flags & MTH_FLAG_QUARTET_COUNTS) != 0;
            const bool shared = (M & MTH_PM) && (M & MTH_ME) && q_set[1] == 0;  // identical thresholds: one histogram, identical rows
            for (int q = 0; q < 2; q++) {
                uint32_t bit = q ? MTH_ME : MTH_PM;
                int m = q ? M_ME : M_PM;
                if (!(M & bit)) continue;
                QuartetRowsBuf& r = q ? c->rows_me : c->rows_pm;
                TRY(reserve_quartet_rows(c, r, r.n + (int64_t)tot[m], counts));
            }
            for (int q = 0; q < 2; q++) {
                uint32_t bit = q ? MTH_ME : MTH_PM;
                int m = q ? M_ME : M_PM;
                if (!(M & bit)) continue;
                QuartetRowsBuf& r = q ? c->rows_me : c->rows_pm;
                QuartetRowsDev rd = quartet_rows_dev(r);
                if (!counts) rd.counts = nullptr;
                const mth_quartet_params qp = q ? c->prm.me : c->prm.pm;
                const int qs = q_set[q];
                const uint8_t* qm = (const uint8_t*)c->qmixed[qs].p;
                if (shared && q == 1) continue;  // histogram and rows were written together with PM's
                // rows of the mixed sites (gather, latency-bound) beside the histogram + canonical rows (streaming): disjoint rows
                cudaStream_t gs = side_stream_begin(c, s);
                {  // row histograms from the observation list (once per histogram set)
                    ProfScope ps(c, q ? "k_me_hist" : "k_pm_hist");
                    TRY(dev_reserve(c, c->qhrows[qs], (size_t)(tot[m] + 1) * 64, 0));
                    CUDA_TRY(c, cudaMemsetAsync(c->qhrows[qs].p, 0, (size_t)(tot[m] + 1) * 64, s));
                    ps.add(launch_quartet_hist((const uint32_t*)c->qobs_site[qs].p, (const uint8_t*)c->qobs_vp[qs].p,
                                               (const unsigned long long*)c->qobs_n.p + qs, rv.I, (const uint32_t*)c->qcnt[qs].p, qm,
                                               (const uint32_t*)c->rowcnt[m].p, qp.min_depth, (uint32_t*)c->qhrows[qs].p, s));
                }
                ProfScope ps(c, shared ? "k_pm_me_emit" : (q ? "k_me_emit" : "k_pm_emit"));
                if (shared) {  // PM and ME rows from one read of the histograms / one gather pass over the mixed sites
                    QuartetRowsDev rd2 = quartet_rows_dev(c->rows_me);
                    if (!counts) rd2.counts = nullptr;
                    ps.add(launch_quartet_canon_emit((const uint32_t*)c->qcnt[qs].p, qm, (const uint32_t*)c->qhrows[qs].p, site_pos, C, qp.min_depth, 2,
                                                     (const uint32_t*)c->rowcnt[m].p, (const float*)c->me_lut.p, c->me_lut_max, ct, rd, r.n, rd2,
                                                     c->rows_me.n, s));
                    ps.add(launch_quartet_emit(rv, site_pos, C, d_sc, qp, 2, qm, (const uint32_t*)c->qmix_list[qs].p,
                                               (const unsigned long long*)c->qmix_n.p + qs, (const uint32_t*)c->rowcnt[m].p, (const float*)c->me_lut.p,
                                               c->me_lut_max, ct, rd, r.n, rd2, c->rows_me.n, gs));
                } else {
                    ps.add(launch_quartet_canon_emit((const uint32_t*)c->qcnt[qs].p, qm, (const uint32_t*)c->qhrows[qs].p, site_pos, C, qp.min_depth, q,
                                                     (const uint32_t*)c->rowcnt[m].p, (const float*)c->me_lut.p, c->me_lut_max, ct, rd, r.n, rd, r.n, s));
                    ps.add(launch_quartet_emit(rv, site_pos, C, d_sc, qp, q, qm, (const uint32_t*)c->qmix_list[qs].p,
                                               (const unsigned long long*)c->qmix_n.p + qs, (const uint32_t*)c->rowcnt[m].p, (const float*)c->me_lut.p,
                                               c->me_lut_max, ct, rd, r.n, rd, r.n, gs));
                }
                TRY(side_stream_end(c, s, gs));
            }
            if (M & MTH_PM) c->rows_pm.n += (int64_t)tot[M_PM];
            if (M & MTH_ME) c->rows_me.n += (int64_t)tot[M_ME];
        }
        if (want_pairs) {
            PairRowsBuf& r = c->rows_pairs;
            int64_t n = r.n + (int64_t)tot[M_PAIRS];
            for (DevBuf* b : {&r.tid, &r.pos1, &r.pos2, &r.lpmd, &r.nc, &r.nd}) TRY(dev_reserve(c, *b, (size_t)n * 4 + 4, (size_t)r.n * 4));
            PairRowsDev rd{(int32_t*)r.tid.p, (int32_t*)r.pos1.p, (int32_t*)r.pos2.p, (float*)r.lpmd.p, (int32_t*)r.nc.p, (int32_t*)r.nd.p};
            ProfScope ps(c, "k_lpmd_pairs_emit");
            ps.add(launch_lpmd_pairs_emit(rv, rel_all, site_pos, C, d_sc, c->prm.lpmd, (const uint32_t*)c->rowcnt[M_PAIRS].p, ct, rd,
                                          r.n, s));
            r.n = n;
        }
        CUDA_TRY(c, cudaGetLastError());
    }
    return region_end(c, k);
}

// phase B of the parked region (rows are appended in region order: it goes first)
static int finish_pending(mth_ctx* c) {
    if (!c->has_pend) return MTH_OK;
    swap_slot(c, c->pend.st);
    const int rc = region_phase_b(c, c->pend.k);
    swap_slot(c, c->pend.st);
    c->has_pend = false;
    return rc;
}

static int process_region(mth_ctx* c) {
    TRY(finish_pending(c));
    if (!c->region_active) return MTH_OK;
    RegionCarry k;
    TRY(region_close(c, k));
    TRY(region_phase_a(c, k));
    return region_phase_b(c, k);
}

static int fetch_site_rows(mth_ctx* c, SiteRowsBuf& r, bool counts, bool to_host, mth_site_rows* out) {
    memset(out, 0, sizeof(*out));
    out->n = r.n;
    if (!to_host || r.n == 0) return MTH_OK;
    size_t nb = (size_t)r.n * 4;
    TRY(host_reserve(c, r.h_tid, nb)); TRY(host_reserve(c, r.h_pos, nb)); TRY(host_reserve(c, r.h_value, nb));
    cudaStream_t s = c->compute;
    CUDA_TRY(c, cudaMemcpyAsync(r.h_tid.p, r.tid.p, nb, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaMemcpyAsync(r.h_pos.p, r.pos.p, nb, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaMemcpyAsync(r.h_value.p, r.value.p, nb, cudaMemcpyDeviceToHost, s));
    c->stats.d2h_bytes += (int64_t)nb * 3;
    out->tid = (const int32_t*)r.h_tid.p; out->pos = (const int32_t*)r.h_pos.p; out->value = (const float*)r.h_value.p;
    if (counts) {
        TRY(host_reserve(c, r.h_nc, nb)); TRY(host_reserve(c, r.h_nd, nb));
        CUDA_TRY(c, cudaMemcpyAsync(r.h_nc.p, r.nc.p, nb, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(c, cudaMemcpyAsync(r.h_nd.p, r.nd.p, nb, cudaMemcpyDeviceToHost, s));
        c->stats.d2h_bytes += (int64_t)nb * 2;
        out->n_conc = (const uint32_t*)r.h_nc.p; out->n_disc = (const uint32_t*)r.h_nd.p;
    }
    return MTH_OK;
}
static int fetch_quartet_rows(mth_ctx* c, QuartetRowsBuf& r, bool counts, bool to_host, mth_quartet_rows* out) {
    memset(out, 0, sizeof(*out));
    out->n = r.n;
    if (!to_host || r.n == 0) return MTH_OK;
    size_t nb = (size_t)r.n * 4;
    cudaStream_t s = c->compute;
    DevBuf* d[6] = {&r.tid, &r.p1, &r.p2, &r.p3, &r.p4, &r.value};
    HostBuf* h[6] = {&r.h_tid, &r.h_p1, &r.h_p2, &r.h_p3, &r.h_p4, &r.h_value};
    for (int i = 0; i < 6; i++) {
        TRY(host_reserve(c, *h[i], nb));
        CUDA_TRY(c, cudaMemcpyAsync(h[i]->p, d[i]->p, nb, cudaMemcpyDeviceToHost, s));
    }
    c->stats.d2h_bytes += (int64_t)nb * 6;
    out->tid = (const int32_t*)r.h_tid.p; out->p1 = (const int32_t*)r.h_p1.p; out->p2 = (const int32_t*)r.h_p2.p;
    out->p3 = (const int32_t*)r.h_p3.p; out->p4 = (const int32_t*)r.h_p4.p; out->value = (const float*)r.h_value.p;
    if (counts) {
        TRY(host_reserve(c, r.h_counts, nb * 16));
        CUDA_TRY(c, cudaMemcpyAsync(r.h_counts.p, r.counts.p, nb * 16, cudaMemcpyDeviceToHost, s));
        c->stats.d2h_bytes += (int64_t)nb * 16;
        out->counts = (const uint32_t*)r.h_counts.p;
    }
    return MTH_OK;
}

static int fetch_pair_rows(mth_ctx* c, PairRowsBuf& r, bool to_host, mth_pair_rows* out) {
    memset(out, 0, sizeof(*out));
    out->n = r.n;
    if (!to_host || r.n == 0) return MTH_OK;
    size_t nb = (size_t)r.n * 4;
    DevBuf* d[6] = {&r.tid, &r.pos1, &r.pos2, &r.lpmd, &r.nc, &r.nd};
    HostBuf* h[6] = {&r.h_tid, &r.h_pos1, &r.h_pos2, &r.h_lpmd, &r.h_nc, &r.h_nd};
    for (int i = 0; i < 6; i++) {
        TRY(host_reserve(c, *h[i], nb));
        CUDA_TRY(c, cudaMemcpyAsync(h[i]->p, d[i]->p, nb, cudaMemcpyDeviceToHost, c->compute));
    }
    c->stats.d2h_bytes += (int64_t)nb * 6;
    out->tid = (const int32_t*)r.h_tid.p; out->pos1 = (const int32_t*)r.h_pos1.p; out->pos2 = (const int32_t*)r.h_pos2.p;
    out->lpmd = (const float*)r.h_lpmd.p; out->n_conc = (const int32_t*)r.h_nc.p; out->n_disc = (const int32_t*)r.h_nd.p;
    return MTH_OK;
}

static void lpmd_from_totals(const int64_t* t, mth_lpmd_result* out) {
    out->n_read = t[0]; out->n_valid_read = t[1]; out->n_conc = t[2]; out->n_disc = t[3];
    // lpmd.rs:51-55: n_discordant as f32 / (n_concordant + n_discordant) as f32
    volatile float num = (float)t[3];
    volatile float den = (float)(t[2] + t[3]);
    out->lpmd = num / den;
}

extern "C" {

int mth_finish(mth_ctx* c, mth_results* out) {
    if (!c || !out) return MTH_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (!c->finished) {
        TRY(process_region(c));
        c->finished = 1;
    }
    memset(out, 0, sizeof(*out));
    const bool to_host = !(c->prm.flags & MTH_FLAG_KEEP_ON_DEVICE);
    const bool qc = (c->prm.flags & MTH_FLAG_QUARTET_COUNTS) != 0;
    TRY(fetch_site_rows(c, c->rows_pdr, true, to_host, &out->pdr));
    TRY(fetch_site_rows(c, c->rows_mhl, false, to_host, &out->mhl));
    TRY(fetch_site_rows(c, c->rows_fdrp, false, to_host, &out->fdrp));
    TRY(fetch_site_rows(c, c->rows_qfdrp, false, to_host, &out->qfdrp));
    TRY(fetch_quartet_rows(c, c->rows_pm, qc, to_host, &out->pm));
    TRY(fetch_quartet_rows(c, c->rows_me, qc, to_host, &out->me));
    TRY(fetch_pair_rows(c, c->rows_pairs, to_host, &out->lpmd_pairs));
    lpmd_from_totals(c->lpmd_total_host, &out->lpmd);
    TRY(dev_reserve(c, c->lpmd_total, 32, 0));
    CUDA_TRY(c, cudaMemcpyAsync(c->lpmd_total.p, c->lpmd_total_host, 32, cudaMemcpyHostToDevice, c->compute));
    CUDA_TRY(c, cudaStreamSynchronize(c->copy));
    CUDA_TRY(c, cudaStreamSynchronize(c->compute));
    resolve_spans(c);
    return MTH_OK;
}

int mth_results_device(mth_ctx* c, mth_results* out) {
    if (!c || !out) return MTH_ERR_INVALID;
    if (!c->finished) return fail(c, MTH_ERR_STATE, "mth_results_device before mth_finish");
    memset(out, 0, sizeof(*out));
    auto site = [](SiteRowsBuf& r, bool counts, mth_site_rows* o) {
        o->n = r.n; o->tid = (const int32_t*)r.tid.p; o->pos = (const int32_t*)r.pos.p; o->value = (const float*)r.value.p;
        o->n_conc = counts ? (const uint32_t*)r.nc.p : nullptr; o->n_disc = counts ? (const uint32_t*)r.nd.p : nullptr;
    };
    auto quart = [](QuartetRowsBuf& r, bool counts, mth_quartet_rows* o) {
        o->n = r.n; o->tid = (const int32_t*)r.tid.p; o->p1 = (const int32_t*)r.p1.p; o->p2 = (const int32_t*)r.p2.p;
        o->p3 = (const int32_t*)r.p3.p; o->p4 = (const int32_t*)r.p4.p; o->value = (const float*)r.value.p;
        o->counts = counts ? (const uint32_t*)r.counts.p : nullptr;
    };
    const bool qc = (c->prm.flags & MTH_FLAG_QUARTET_COUNTS) != 0;
    site(c->rows_pdr, true, &out->pdr); site(c->rows_mhl, false, &out->mhl);
    site(c->rows_fdrp, false, &out->fdrp); site(c->rows_qfdrp, false, &out->qfdrp);
    quart(c->rows_pm, qc, &out->pm); quart(c->rows_me, qc, &out->me);
    {
        PairRowsBuf& r = c->rows_pairs;
        out->lpmd_pairs.n = r.n; out->lpmd_pairs.tid = (const int32_t*)r.tid.p; out->lpmd_pairs.pos1 = (const int32_t*)r.pos1.p;
        out->lpmd_pairs.pos2 = (const int32_t*)r.pos2.p; out->lpmd_pairs.lpmd = (const float*)r.lpmd.p;
        out->lpmd_pairs.n_conc = (const int32_t*)r.nc.p; out->lpmd_pairs.n_disc = (const int32_t*)r.nd.p;
    }
    lpmd_from_totals(c->lpmd_total_host, &out->lpmd);
    return MTH_OK;
}

int mth_lpmd_counters_device(mth_ctx* c, void** dev_ptr) {
    if (!c || !dev_ptr) return MTH_ERR_INVALID;
    if (!c->finished) return fail(c, MTH_ERR_STATE, "mth_lpmd_counters_device before mth_finish");
    *dev_ptr = c->lpmd_total.p;
    return MTH_OK;
}

int mth_lpmd_refresh(mth_ctx* c, mth_lpmd_result* out) {
    if (!c || !out) return MTH_ERR_INVALID;
    if (!c->finished) return fail(c, MTH_ERR_STATE, "mth_lpmd_refresh before mth_finish");
    CUDA_TRY(c, cudaSetDevice(c->device));
    int64_t t[4];
    CUDA_TRY(c, cudaMemcpy(t, c->lpmd_total.p, 32, cudaMemcpyDeviceToHost));
    lpmd_from_totals(t, out);
    return MTH_OK;
}

int mth_get_stats(mth_ctx* c, mth_stats* out) {
    if (!c || !out) return MTH_ERR_INVALID;
    *out = c->stats;
    return MTH_OK;
}

}  // extern "C"

// ---- multi-GPU: NCCL, loaded at run time -------------------------------------------------------------------------
namespace {
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};
NcclApi g_nccl;

// The copy of NCCL already in the process wins (under PyTorch that is the one torch.distributed uses, so there is one NCCL
// per process); otherwise the system library.
bool nccl_load() {
    if (g_nccl.h) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        g_nccl.err = std::string("cannot load libnccl.so.2: ") + dlerror();
        return false;
    }
    bool ok = true;
    auto sym = [&](const char* name) {
        void* p = dlsym(h, name);
        if (!p) { ok = false; g_nccl.err = std::string("libnccl lacks ") + name; }
        return p;
    };
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
    g_nccl.CommInitAll = (decltype(g_nccl.CommInitAll))sym("ncclCommInitAll");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
    if (!ok) return false;
    g_nccl.h = h;
    return true;
}
int nccl_fail(mth_ctx* c, const char* what, ncclResult_t r) {
    std::string m = std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error");
    if (c) c->err = m; else g_create_err = m;
    return MTH_ERR_CUDA;
}
// enqueue the sum of the context's device-side LPMD totals (int64[4]) on its compute stream
int enqueue_allreduce(mth_ctx* c) {
    if (!c->comm) return fail(c, MTH_ERR_STATE, "mth_allreduce without a communicator (mth_comm_init_rank / mth_comm_init_all)");
    if (!c->finished) return fail(c, MTH_ERR_STATE, "mth_allreduce before mth_finish");
    ncclResult_t r = g_nccl.AllReduce(c->lpmd_total.p, c->lpmd_total.p, 4, ncclInt64, ncclSum, c->comm, c->compute);
    if (r != ncclSuccess) return nccl_fail(c, "ncclAllReduce", r);
    c->stats.kernel_launches += 1;
    return MTH_OK;
}
int fetch_allreduced(mth_ctx* c) {
    CUDA_TRY(c, cudaMemcpyAsync(c->lpmd_total_host, c->lpmd_total.p, 32, cudaMemcpyDeviceToHost, c->compute));
    CUDA_TRY(c, cudaStreamSynchronize(c->compute));
    return MTH_OK;
}
}  // namespace

extern "C" {

int mth_comm_unique_id(void* id128) {
    if (!id128) return MTH_ERR_INVALID;
    if (!nccl_load()) { g_create_err = g_nccl.err; return MTH_ERR_CUDA; }
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return nccl_fail(nullptr, "ncclGetUniqueId", r);
    static_assert(sizeof(id) == MTH_COMM_ID_BYTES, "ncclUniqueId size");
    memcpy(id128, &id, sizeof(id));
    return MTH_OK;
}

int mth_comm_init_rank(mth_ctx* c, int n_ranks, int rank, const void* id128) {
    if (!c || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return MTH_ERR_INVALID;
    if (c->comm) return fail(c, MTH_ERR_STATE, "context already has a communicator");
    if (!nccl_load()) return fail(c, MTH_ERR_CUDA, g_nccl.err);
    CUDA_TRY(c, cudaSetDevice(c->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, n_ranks, id, rank);
    if (r != ncclSuccess) { c->comm = nullptr; return nccl_fail(c, "ncclCommInitRank", r); }
    c->comm_ranks = n_ranks;
    return MTH_OK;
}

int mth_comm_init_all(mth_ctx** ctxs, int n) {
    if (!ctxs || n < 1) return MTH_ERR_INVALID;
    for (int i = 0; i < n; i++)
        if (!ctxs[i] || ctxs[i]->comm) return MTH_ERR_INVALID;
    if (!nccl_load()) return fail(ctxs[0], MTH_ERR_CUDA, g_nccl.err);
    std::vector<int> devs(n);
    std::vector<ncclComm_t> comms(n, nullptr);
    for (int i = 0; i < n; i++) devs[i] = ctxs[i]->device;
    ncclResult_t r = g_nccl.CommInitAll(comms.data(), n, devs.data());
    if (r != ncclSuccess) return nccl_fail(ctxs[0], "ncclCommInitAll", r);
    for (int i = 0; i < n; i++) { ctxs[i]->comm = comms[i]; ctxs[i]->comm_ranks = n; }
    return MTH_OK;
}

int mth_allreduce(mth_ctx* c) {
    if (!c) return MTH_ERR_INVALID;
    CUDA_TRY(c, cudaSetDevice(c->device));
    TRY(enqueue_allreduce(c));
    return fetch_allreduced(c);
}

int mth_allreduce_group(mth_ctx** ctxs, int n) {
    if (!ctxs || n < 1) return MTH_ERR_INVALID;
    for (int i = 0; i < n; i++)
        if (!ctxs[i]) return MTH_ERR_INVALID;
    if (!g_nccl.h) return fail(ctxs[0], MTH_ERR_STATE, "mth_allreduce_group without a communicator");
    ncclResult_t r = g_nccl.GroupStart();
    if (r != ncclSuccess) return nccl_fail(ctxs[0], "ncclGroupStart", r);
    int rc = MTH_OK;
    for (int i = 0; i < n && rc == MTH_OK; i++) {
        cudaSetDevice(ctxs[i]->device);
        rc = enqueue_allreduce(ctxs[i]);
    }
    r = g_nccl.GroupEnd();
    if (rc != MTH_OK) return rc;
    if (r != ncclSuccess) return nccl_fail(ctxs[0], "ncclGroupEnd", r);
    for (int i = 0; i < n; i++) {
        CUDA_TRY(ctxs[i], cudaSetDevice(ctxs[i]->device));
        TRY(fetch_allreduced(ctxs[i]));
    }
    return MTH_OK;
}

int mth_comm_destroy(mth_ctx* c) {
    if (!c) return MTH_ERR_INVALID;
    if (c->comm && g_nccl.CommDestroy) {
        cudaSetDevice(c->device);
        g_nccl.CommDestroy(c->comm);
    }
    c->comm = nullptr;
    c->comm_ranks = 0;
    return MTH_OK;
}

int mth_comm_n_ranks(mth_ctx* c) { return c ? c->comm_ranks : 0; }

}  // extern "C"
