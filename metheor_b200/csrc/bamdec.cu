// bamdec.cu — BGZF inflate + BAM record decode ON THE DEVICE (include/metheor_b200.h mth_bamdec_*).
//
// Replaces, for BAM input on one GPU, what the reference does per record on one host thread: htslib's bgzf inflate +
// bam_read1 (bamutil.rs:4-11) and BismarkRead::new + get_cpgs (readutil.rs:24-53, 323-345).  The host only walks the BGZF
// member headers of the memory-mapped file (18 bytes per <= 64 KiB member) and ships the COMPRESSED bytes; everything else
// happens here, window by window (a window = a run of whole members, ~100-200 MB of output):
//   k_bgzf_inflate   one warp per member: raw DEFLATE -> the window's uncompressed byte stream (inflate.cuh)
//   k_rec_entry      records follow each other as [block_size][block_size bytes]: finding the record boundaries is a serial
//                    pointer chase.  It is made parallel by SPECULATION + VERIFICATION: for every 32 KiB chunk a CTA looks
//                    for the first offset that parses as a plausible record header whose successors are plausible too ...
//   k_rec_walk       ... one thread per chunk then walks the chain from its entry to the end of the chunk and checks that it
//                    lands exactly on the next chunk's entry.  Chunk 0 starts at a known boundary, so when every link
//                    verifies the whole chain is exact by induction; a link that does not verify is repaired by re-walking
//                    (k_rec_repair, sequential, practically never needed).  Second walk: record offsets.
//   k_rec_decode     one thread per record, two passes (count, emit): fixed fields, aux walk to XM:Z, the CIGAR walk zipped
//                    with the XM string exactly like BismarkRead::new — first / last aligned position, strand shift
//                    (flags 0 / 99 / 147 are forward, readutil.rs:332), z / Z calls with their query index — written as
//                    the SoA batch layout of mth_batch (device memory), reads without any call dropped and counted.
// The result is a list of device-resident mth_batch runs (one per contig present in the window) that the host hands to
// mth_submit (mem_kind 1): the reads never exist on the host.
// Limits of this path (the host falls back to its CPU decoder, host/decode.cpp, for a file that exceeds them): at most 64
// CpG calls per read, query index < 65536, at most 1024 contig changes per window.
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "inflate.cuh"
#include "kernels.h"

using namespace mth;

namespace {

constexpr int INF_WARPS = 8;                 // members per CTA of k_bgzf_inflate
constexpr int STAGE_THREADS = 4;             // host threads per upload of a compressed window (mth_bamdec_stage)
constexpr uint32_t REC_CHUNK = 32u << 10;    // bytes of the uncompressed stream per speculative chain segment
constexpr int MAX_RUN_MARKS = 1024;
constexpr uint32_t NO_ENTRY = 0xffffffffu;

struct MemberDesc {
    unsigned long long in_off;   // raw DEFLATE payload within the compressed window
    uint32_t in_len, isize;
    unsigned long long out_off;  // where its output goes in the uncompressed stream
    uint32_t crc, check_crc;     // CRC-32 of the output (gzip trailer); check_crc 0: not given
};

__global__ void __launch_bounds__(INF_WARPS * 32) k_bgzf_inflate(const uint8_t* __restrict__ comp, const MemberDesc* __restrict__ m, int64_t n,
                                                                 uint8_t* out, int* __restrict__ status, int* __restrict__ any_bad) {
    extern __shared__ __align__(16) unsigned char inf_raw[];
    __shared__ uint32_t s_crc_tab[256];
    InflateTables* T = reinterpret_cast<InflateTables*>(inf_raw) + (threadIdx.x >> 5);
    crc_table_init(s_crc_tab);
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * INF_WARPS + (threadIdx.x >> 5);
    if (i >= n) return;
    const MemberDesc d = m[i];
    uint32_t produced = 0;
    int err = warp_inflate(comp + d.in_off, d.in_len, out + d.out_off, d.isize, *T, &produced);
    if (err == INF_OK && produced != d.isize) err = INF_ERR_SIZE;
    if (err == INF_OK && d.check_crc) {  // what htslib's bgzf reader verifies for every block
        __syncwarp();
        if (warp_crc32(out + d.out_off, d.isize, s_crc_tab) != d.crc) err = INF_ERR_CRC;
    }
    if (lane_id() == 0) {
        status[i] = err;
        if (err) atomicOr(any_bad, 1);
    }
}

// ---- record boundaries ------------------------------------------------------------------------------------------
struct RefTable {
    int32_t n_ref;
    const int64_t* len;
};

__device__ __forceinline__ uint32_t ld32(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
__device__ __forceinline__ uint32_t ld16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

// Does a BAM record header parse at offset o of u[0, end)?  Sets *next to the offset behind the record.
__device__ __forceinline__ bool plausible(const uint8_t* __restrict__ u, uint64_t o, uint64_t end, const RefTable& rt, uint64_t* next) {
    if (o + 36 > end) return false;
    const uint8_t* p = u + o;
    const uint32_t bs = ld32(p);
    if (bs < 32u || bs > (64u << 20)) return false;
    const int32_t tid = (int32_t)ld32(p + 4), pos = (int32_t)ld32(p + 8);
    if (tid < -1 || tid >= rt.n_ref || pos < -1) return false;
    if (tid >= 0 && (int64_t)pos > rt.len[tid]) return false;
    const uint32_t l_name = p[12], n_cig = ld16(p + 16);
    const int32_t l_seq = (int32_t)ld32(p + 20), mtid = (int32_t)ld32(p + 24), mpos = (int32_t)ld32(p + 28);
    if (l_name == 0 || l_seq < 0 || mtid < -1 || mtid >= rt.n_ref || mpos < -1) return false;
    const uint64_t fixed = 32ull + l_name + 4ull * n_cig + ((uint64_t)l_seq + 1) / 2 + (uint64_t)l_seq;
    if (fixed > bs) return false;
    const uint64_t name_end = o + 36 + l_name - 1;
    if (name_end < end && u[name_end] != 0) return false;  // read names are NUL-terminated
    *next = o + 4 + bs;
    return true;
}

// entry[c] = first offset >= max(c * REC_CHUNK, first) where a chain of plausible records starts (NO_ENTRY: none before `end`)
__global__ void __launch_bounds__(256) k_rec_entry(const uint8_t* __restrict__ u, uint64_t end, uint64_t first, RefTable rt, int64_t n_chunks,
                                                   uint32_t* __restrict__ entry) {
    __shared__ unsigned int s_best;
    const int64_t c = blockIdx.x;
    if (c >= n_chunks) return;
    const uint64_t lo = max((uint64_t)c * REC_CHUNK, first);
    if (c == 0) {  // known exactly
        if (threadIdx.x == 0) entry[0] = (uint32_t)first;
        return;
    }
    if (threadIdx.x == 0) s_best = NO_ENTRY;
    __syncthreads();
    for (uint64_t base = lo; base < end; base += blockDim.x) {
        const uint64_t o = base + threadIdx.x;
        bool ok = false;
        if (o < end) {
            uint64_t nx = o, cur = o;
            ok = true;
            for (int k = 0; k < 4 && ok; k++) {  // the candidate and up to three successors must parse
                if (cur + 36 > end) break;       // the chain runs into the tail of the window: accept what was seen
                ok = plausible(u, cur, end, rt, &nx);
                if (ok && nx > end) break;       // a record that continues in the next window
                cur = nx;
            }
            if (o + 36 > end) ok = false;
        }
        if (ok) atomicMin(&s_best, (unsigned int)o);
        __syncthreads();
        if (s_best != NO_ENTRY) break;
        __syncthreads();
        if (base - lo > (8u << 20)) break;  // give up: the verification walk decides
    }
    __syncthreads();
    if (threadIdx.x == 0) entry[c] = s_best;
}

// One thread per chunk: follow the chain from entry[c] while it is below the end of the chunk.
//   COUNT: n_rec[c] = records that START in the chunk and END inside the window; landing[c] = first offset reached at or behind
//          the end of the chunk (or the start of the trailing partial record); *mismatch |= landing[c] != entry[c + 1]
//   EMIT : rec_off[base[c] + k] = offset of the k-th such record
template <bool EMIT>
__global__ void __launch_bounds__(128) k_rec_walk(const uint8_t* __restrict__ u, uint64_t end, int64_t n_chunks, const uint32_t* __restrict__ entry,
                                                  uint32_t* __restrict__ n_rec, uint32_t* __restrict__ landing, int* __restrict__ mismatch,
                                                  const uint32_t* __restrict__ base, uint32_t* __restrict__ rec_off) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    const uint64_t chunk_end = min((uint64_t)(c + 1) * REC_CHUNK, end);
    uint64_t o = entry[c];
    uint32_t n = 0;
    const uint32_t b0 = EMIT ? base[c] : 0u;
    if (o != NO_ENTRY) {
        while (o < chunk_end) {
            if (o + 4 > end) break;
            const uint64_t nx = o + 4 + (uint64_t)ld32(u + o);
            if (nx > end) break;  // partial record at the end of the window: carried over
            if (EMIT) rec_off[b0 + n] = (uint32_t)o;
            n++;
            o = nx;
        }
    }
    if (!EMIT) {
        n_rec[c] = n;
        landing[c] = o == NO_ENTRY ? NO_ENTRY : (uint32_t)o;
        if (c + 1 < n_chunks) {
            const uint32_t want = entry[c + 1];
            // a chunk the chain jumps over entirely (a record longer than the chunk) has its entry behind its own end
            if ((uint32_t)o != want && !(o < chunk_end && want == NO_ENTRY)) atomicOr(mismatch, 1);
        }
    }
}

// Where the trailing partial record starts: the last landing that exists (else `first`: no record ends in this window).
__global__ void k_rec_tail(const uint32_t* __restrict__ landing, int64_t n_chunks, uint64_t first, unsigned long long* __restrict__ out) {
    if (blockIdx.x || threadIdx.x) return;
    unsigned long long v = first;
    for (int64_t c = n_chunks - 1; c >= 0; c--)
        if (landing[c] != NO_ENTRY) { v = landing[c]; break; }
    *out = v;
}

// Sequential repair of the speculation (one thread): entry[c] := landing[c - 1] wherever they differ, re-walking the chunk.
__global__ void k_rec_repair(const uint8_t* __restrict__ u, uint64_t end, int64_t n_chunks, uint32_t* entry, uint32_t* n_rec, uint32_t* landing) {
    if (blockIdx.x || threadIdx.x) return;
    for (int64_t c = 1; c < n_chunks; c++) {
        const uint32_t e = landing[c - 1];
        if (e == entry[c]) continue;
        entry[c] = e;
        const uint64_t chunk_end = min((uint64_t)(c + 1) * REC_CHUNK, end);
        uint64_t o = e;
        uint32_t n = 0;
        if (e != NO_ENTRY) {
            while (o < chunk_end) {
                if (o + 4 > end) break;
                const uint64_t nx = o + 4 + (uint64_t)ld32(u + o);
                if (nx > end) break;
                n++;
                o = nx;
            }
        }
        n_rec[c] = n;
        landing[c] = e == NO_ENTRY ? NO_ENTRY : (uint32_t)o;
    }
}

// ---- record -> BismarkRead (readutil.rs:24-53, 323-345) -------------------------------------------------------------
struct DecodeArgs {
    const uint8_t* u;
    const uint32_t* rec_off;
    int64_t n_rec;
    RefTable rt;
    uint32_t lpmd_order, min_qual;
    // COUNT out
    uint32_t* keep;      // 1: the read has CpG calls and goes to the engine
    uint32_t* ncpg;      // calls of the read (0 when dropped)
    unsigned long long* counters;  // [0] dropped, [1] dropped with mapq >= min_qual, [2] max span, [3] max calls, [4] first bad record + 1 (atomicMin on ~), [5] unsupported
    unsigned long long* marks;  // contig changes: (tid << 32) | record index
    int* n_marks;
    // EMIT in
    const uint32_t* keep_scan;  // exclusive scans of keep / ncpg
    const uint32_t* call_scan;
    int64_t rec_lo, rec_hi;     // records of this run
    uint32_t read_base, call_base;  // first kept read / call of the run
    uint32_t run_reads;             // kept reads of the run
    int32_t* start; int32_t* end; uint32_t* meta; uint32_t* off; int32_t* pos; uint16_t* rel; unsigned long long* meth;
};

struct RecView {
    int32_t tid, pos;
    uint32_t mapq, flag, n_cig;
    const uint8_t* cig;
    const uint8_t* xm;  // nullptr: no XM:Z
    uint32_t xm_len;
    bool corrupt;
};

__device__ __forceinline__ RecView parse_record(const uint8_t* __restrict__ u, uint32_t o) {
    RecView r;
    const uint8_t* p = u + o + 4;
    const uint32_t bs = ld32(u + o);
    r.tid = (int32_t)ld32(p);
    r.pos = (int32_t)ld32(p + 4);
    const uint32_t l_name = p[8];
    r.mapq = p[9];
    r.n_cig = ld16(p + 12);
    r.flag = ld16(p + 14);
    const int32_t l_seq = (int32_t)ld32(p + 16);
    uint64_t q = 32ull + l_name;
    r.cig = p + q;
    q += 4ull * r.n_cig + ((uint64_t)(uint32_t)l_seq + 1) / 2 + (uint64_t)(uint32_t)l_seq;
    r.xm = nullptr;
    r.xm_len = 0;
    r.corrupt = l_seq < 0 || q > bs;
    if (r.corrupt) return r;
    while (q + 3 <= bs) {  // aux fields: tag[2], type, value (host/decode.cpp decode_bam)
        const uint8_t t0 = p[q], t1 = p[q + 1], ty = p[q + 2];
        q += 3;
        uint64_t len;
        switch (ty) {
            case 'A': case 'c': case 'C': len = 1; break;
            case 's': case 'S': len = 2; break;
            case 'i': case 'I': case 'f': len = 4; break;
            case 'd': len = 8; break;
            case 'Z': case 'H': {
                uint64_t e = q;
                while (e < bs && p[e] != 0) e++;
                if (e >= bs) { r.corrupt = true; return r; }
                len = e - q + 1;
                break;
            }
            case 'B': {
                if (q + 5 > bs) { r.corrupt = true; return r; }
                const uint8_t st = p[q];
                const uint64_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
                len = 5 + es * (uint64_t)ld32(p + q + 1);
                break;
            }
            default: r.corrupt = true; return r;
        }
        if (len > bs - q) { r.corrupt = true; return r; }
        if (t0 == 'X' && t1 == 'M') {
            if (ty == 'Z') { r.xm = p + q; r.xm_len = (uint32_t)(len - 1); }
            break;  // a non-string XM panics like a missing one (readutil.rs:45-47)
        }
        q += len;
    }
    return r;
}

template <bool EMIT>
__global__ void __launch_bounds__(128) k_rec_decode(DecodeArgs a) {
    const int64_t i = (EMIT ? a.rec_lo : 0) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (EMIT ? a.rec_hi : a.n_rec)) return;
    if (EMIT && !a.keep[i]) return;
    const uint32_t o = a.rec_off[i];
    const RecView r = parse_record(a.u, o);
    if (!EMIT) {
        a.keep[i] = 0;
        a.ncpg[i] = 0;
        const int32_t prev_tid = i ? (int32_t)ld32(a.u + a.rec_off[i - 1] + 4) : INT32_MIN;
        if (i == 0 || prev_tid != r.tid) {
            const int k = atomicAdd(a.n_marks, 1);
            if (k < MAX_RUN_MARKS) a.marks[k] = ((unsigned long long)(uint32_t)r.tid << 32) | (unsigned long long)i;
        }
    }
    const bool mapq_ok = r.mapq >= a.min_qual;
    if (a.lpmd_order && !mapq_ok) {  // lpmd.rs:176-181: counted, then skipped before BismarkRead::new looks at XM
        if (!EMIT) atomicAdd(&a.counters[0], 1ull);
        return;
    }
    if (r.corrupt || !r.xm) {
        if (!EMIT) atomicMin(&a.counters[4], ((unsigned long long)i << 1) | (r.corrupt ? 1ull : 0ull));
        return;
    }
    const bool fwd = r.flag == 0 || r.flag == 99 || r.flag == 147;  // readutil.rs:332
    int64_t ref = r.pos;
    uint32_t qi = 0, n = 0;
    int32_t start = -1, end = -1;
    unsigned long long mw = 0;
    const uint32_t call0 = EMIT ? (a.call_scan[i] - a.call_base) : 0u;
    for (uint32_t k = 0; k < r.n_cig; k++) {
        const uint32_t v = ld32(r.cig + 4 * k), len = v >> 4, op = v & 15u;
        if (len == 0) continue;
        if (op == 0 || op == 7 || op == 8) {  // M = X: one reference position per query base
            if (start == -1) start = (int32_t)ref;
            end = (int32_t)(ref + len - 1);
            if (qi < r.xm_len) {
                const uint32_t stop = min(r.xm_len, qi + len);  // zip() stops at the shorter side
                for (uint32_t x = qi; x < stop; x++) {
                    const uint8_t ch = r.xm[x];
                    if ((ch | 0x20) != 'z') continue;  // readutil.rs:327-329
                    if (EMIT) {
                        if (n < 64) {
                            a.pos[call0 + n] = (int32_t)(ref + (int64_t)(x - qi)) - (fwd ? 0 : 1);
                            a.rel[call0 + n] = (uint16_t)x;
                            mw |= (unsigned long long)(ch == 'Z') << n;
                        }
                    } else if (x > 65535u) {
                        a.counters[5] = 1ull;  // query index beyond the engine's 16 bits: host path
                    }
                    n++;
                }
            }
            ref += len;
            qi += len;
        } else if (op == 1 || op == 4) {
            qi += len;   // I S: query only
        } else if (op == 2 || op == 3) {
            ref += len;  // D N: reference only
        }
    }
    if (!EMIT) {
        if (n == 0 || r.tid < 0) {  // nothing to ship (an unplaced read cannot carry calls: it has no aligned base)
            atomicAdd(&a.counters[0], 1ull);
            if (mapq_ok) atomicAdd(&a.counters[1], 1ull);
            return;
        }
        a.keep[i] = 1;
        a.ncpg[i] = n;
        atomicMax(&a.counters[2], (unsigned long long)((int64_t)end - start + 1));
        atomicMax(&a.counters[3], (unsigned long long)n);
        if (n > 64) a.counters[5] = 1ull;
        return;
    }
    const uint32_t j = a.keep_scan[i] - a.read_base;
    a.start[j] = start;
    a.end[j] = end;
    a.meta[j] = r.mapq | ((uint32_t)fwd << 8);
    a.off[j] = call0;
    a.meth[j] = mw;
    if (j + 1 == a.run_reads) a.off[j + 1] = call0 + n;  // last kept read of the run: terminal offset
}

struct DBuf {
    void* p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct mth_bamdec {
    int device = 0;
    cudaStream_t s = nullptr, s_stage[2] = {nullptr, nullptr}, s_part[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
    size_t staged_bytes[2] = {0, 0};
    std::string err;
    int32_t n_ref = 0;
    uint32_t lpmd_order = 0, min_qual = 0;
    DBuf ref_len, comp_slot[2], comp, members, status, u, entry, n_rec, landing, base, rec_off, keep, ncpg, scan_scratch, small;
    DBuf o_start, o_end, o_meta, o_off, o_pos, o_rel, o_meth;
    void* h_small = nullptr;  // pinned mirror of `small`
    size_t carry = 0;         // bytes at the front of `u` carried over from the previous window (a partial record)
    std::vector<mth_batch> runs;
    double ms_inflate = 0, ms_boundaries = 0, ms_decode = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int64_t repairs = 0;
};

namespace {

std::string g_dec_err;

int dfail(mth_bamdec* d, int code, const std::string& m) {
    if (d) d->err = m; else g_dec_err = m;
    return code;
}
#define DTRY(d, call)                                                                      \
    do {                                                                                   \
        cudaError_t _e = (call);                                                           \
        if (_e != cudaSuccess) return dfail(d, MTH_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); \
    } while (0)

int reserve(mth_bamdec* d, DBuf& b, size_t bytes, size_t keep = 0) {
    if (bytes <= b.cap) return MTH_OK;
    size_t ncap = bytes + bytes / 4 + 256;
    void* np = nullptr;
    DTRY(d, cudaStreamSynchronize(d->s));
    DTRY(d, cudaMalloc(&np, ncap));
    if (b.p && keep) DTRY(d, cudaMemcpy(np, b.p, keep, cudaMemcpyDeviceToDevice));
    if (b.p) cudaFree(b.p);
    b.p = np;
    b.cap = ncap;
    return MTH_OK;
}

// small device scratch (mirrored in pinned host memory): indices in units of 8 bytes
enum { SM_ANYBAD = 0, SM_MISMATCH = 1, SM_NMARKS = 2, SM_TOT_REC = 3, SM_TOT_KEEP = 4, SM_TOT_CALLS = 5, SM_COUNTERS = 8, SM_MARKS = 16,
       SM_CARRY = 6, SM_WORDS = 16 + MAX_RUN_MARKS + 8 };

}  // namespace

extern "C" {

const char* mth_bamdec_last_error(mth_bamdec* d) { return d ? d->err.c_str() : g_dec_err.c_str(); }

int mth_bamdec_create(mth_bamdec** out, int device, int32_t n_ref, const int64_t* ref_len, uint32_t lpmd_order, uint32_t min_qual) {
    if (!out || n_ref < 0 || (n_ref && !ref_len)) return dfail(nullptr, MTH_ERR_INVALID, "null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return dfail(nullptr, MTH_ERR_CUDA, "no CUDA device available (no CPU fallback in the engine)");
    if (device < 0 || device >= ndev) return dfail(nullptr, MTH_ERR_INVALID, "device index out of range");
    mth_bamdec* d = new mth_bamdec();
    d->device = device;
    d->n_ref = n_ref;
    d->lpmd_order = lpmd_order;
    d->min_qual = min_qual;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&d->s, cudaStreamNonBlocking) != cudaSuccess) {
        delete d;
        return dfail(nullptr, MTH_ERR_CUDA, "CUDA initialisation failed");
    }
    for (auto& e : d->ev) cudaEventCreate(&e);
    for (auto& st : d->s_stage) cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    for (auto& row : d->s_part)
        for (auto& st : row) cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    if (reserve(d, d->ref_len, (size_t)std::max(1, n_ref) * 8) != MTH_OK || reserve(d, d->small, SM_WORDS * 8) != MTH_OK ||
        cudaHostAlloc(&d->h_small, SM_WORDS * 8, cudaHostAllocDefault) != cudaSuccess) {
        g_dec_err = d->err;
        mth_bamdec_destroy(d);
        return MTH_ERR_CUDA;
    }
    if (n_ref) cudaMemcpy(d->ref_len.p, ref_len, (size_t)n_ref * 8, cudaMemcpyHostToDevice);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_bgzf_inflate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(INF_WARPS * sizeof(InflateTables)));
        attr = true;
    }
    *out = d;
    return MTH_OK;
}

int mth_bamdec_destroy(mth_bamdec* d) {
    if (!d) return MTH_OK;
    cudaSetDevice(d->device);
    if (d->s) cudaStreamSynchronize(d->s);
    for (auto& st : d->s_stage)
        if (st) cudaStreamDestroy(st);
    for (auto& row : d->s_part)
        for (auto& st : row)
            if (st) cudaStreamDestroy(st);
    for (DBuf* b : {&d->ref_len, &d->comp_slot[0], &d->comp_slot[1], &d->comp, &d->members, &d->status, &d->u, &d->entry, &d->n_rec, &d->landing, &d->base, &d->rec_off, &d->keep,
                    &d->ncpg, &d->scan_scratch, &d->small, &d->o_start, &d->o_end, &d->o_meta, &d->o_off, &d->o_pos, &d->o_rel, &d->o_meth})
        if (b->p) cudaFree(b->p);
    if (d->h_small) cudaFreeHost(d->h_small);
    for (auto& e : d->ev)
        if (e) cudaEventDestroy(e);
    if (d->s) cudaStreamDestroy(d->s);
    delete d;
    return MTH_OK;
}

// Inflate only (tests, profiles): members of `comp` -> `out_host` (concatenated in order), per-member status.
int mth_bgzf_inflate(int device, const uint8_t* comp, size_t comp_bytes, const mth_bgzf_member* members, int64_t n, uint8_t* out_host,
                     size_t out_bytes, int32_t* status_host, double* kernel_ms) {
    if (!comp || !members || n < 0 || !out_host) return dfail(nullptr, MTH_ERR_INVALID, "null argument");
    mth_bamdec* d = nullptr;
    int rc = mth_bamdec_create(&d, device, 0, nullptr, 0, 0);
    if (rc != MTH_OK) return rc;
    std::vector<MemberDesc> md((size_t)n);
    size_t uo = 0;
    for (int64_t i = 0; i < n; i++) {
        if (members[i].offset + members[i].size > comp_bytes) { mth_bamdec_destroy(d); return dfail(nullptr, MTH_ERR_INVALID, "member outside the buffer"); }
        md[(size_t)i] = MemberDesc{members[i].offset, members[i].size, members[i].isize, uo, members[i].crc, members[i].flags & 1u};
        uo += members[i].isize;
    }
    if (uo > out_bytes) { mth_bamdec_destroy(d); return dfail(nullptr, MTH_ERR_INVALID, "output buffer too small"); }
    auto run = [&]() -> int {
        if (reserve(d, d->comp, comp_bytes + 64) || reserve(d, d->members, (size_t)n * sizeof(MemberDesc) + 64) ||
            reserve(d, d->status, (size_t)n * 4 + 64) || reserve(d, d->u, uo + 64))
            return MTH_ERR_CUDA;
        DTRY(d, cudaMemcpyAsync(d->comp.p, comp, comp_bytes, cudaMemcpyHostToDevice, d->s));
        DTRY(d, cudaMemcpyAsync(d->members.p, md.data(), (size_t)n * sizeof(MemberDesc), cudaMemcpyHostToDevice, d->s));
        DTRY(d, cudaMemsetAsync(d->small.p, 0, SM_WORDS * 8, d->s));
        DTRY(d, cudaEventRecord(d->ev[0], d->s));
        if (n) k_bgzf_inflate<<<(unsigned)((n + INF_WARPS - 1) / INF_WARPS), INF_WARPS * 32, INF_WARPS * sizeof(InflateTables), d->s>>>(
                (const uint8_t*)d->comp.p, (const MemberDesc*)d->members.p, n, (uint8_t*)d->u.p, (int*)d->status.p, (int*)d->small.p + 2 * SM_ANYBAD);
        DTRY(d, cudaEventRecord(d->ev[1], d->s));
        DTRY(d, cudaMemcpyAsync(out_host, d->u.p, uo, cudaMemcpyDeviceToHost, d->s));
        if (status_host) DTRY(d, cudaMemcpyAsync(status_host, d->status.p, (size_t)n * 4, cudaMemcpyDeviceToHost, d->s));
        DTRY(d, cudaStreamSynchronize(d->s));
        DTRY(d, cudaGetLastError());
        float ms = 0;
        cudaEventElapsedTime(&ms, d->ev[0], d->ev[1]);
        if (kernel_ms) *kernel_ms = ms;
        return MTH_OK;
    };
    rc = run();
    if (rc != MTH_OK) g_dec_err = d->err;
    mth_bamdec_destroy(d);
    return rc;
}

// Upload the compressed bytes of the NEXT window into staging slot 0 / 1.  Blocking (the source is ordinary pageable memory,
// typically the memory-mapped file), on its own stream: meant to be called from a helper thread while mth_bamdec_window works on
// the other slot.
int mth_bamdec_stage(mth_bamdec* d, int slot, const uint8_t* comp, size_t comp_bytes) {
    if (!d || slot < 0 || slot > 1 || (comp_bytes && !comp)) return MTH_ERR_INVALID;
    DTRY(d, cudaSetDevice(d->device));
    DBuf& b = d->comp_slot[slot];
    if (comp_bytes + 64 > b.cap) {  // (not reserve(): that synchronises the decode stream, which another thread may be using)
        if (b.p) DTRY(d, cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
        const size_t ncap = comp_bytes + comp_bytes / 4 + (1u << 20);
        DTRY(d, cudaMalloc(&b.p, ncap));
        b.cap = ncap;
    }
    // A copy out of pageable memory is staged by the driver on the calling thread (~9 GB/s); a few threads in parallel come
    // closer to what the link can do.
    const int nt = comp_bytes > (32u << 20) ? STAGE_THREADS : 1;
    std::vector<std::thread> th;
    std::vector<cudaError_t> rcs((size_t)nt, cudaSuccess);
    for (int t = 0; t < nt; t++) {
        const size_t a = (comp_bytes * (size_t)t / (size_t)nt) & ~(size_t)255, e = t + 1 == nt ? comp_bytes : (comp_bytes * (size_t)(t + 1) / (size_t)nt) & ~(size_t)255;
        th.emplace_back([=, &rcs, &b] {
            cudaSetDevice(d->device);
            cudaStream_t st = d->s_part[slot][t];
            cudaError_t r = e > a ? cudaMemcpyAsync((uint8_t*)b.p + a, comp + a, e - a, cudaMemcpyHostToDevice, st) : cudaSuccess;
            if (r == cudaSuccess) r = cudaStreamSynchronize(st);
            rcs[(size_t)t] = r;
        });
    }
    for (auto& t : th) t.join();
    for (cudaError_t r : rcs)
        if (r != cudaSuccess) return dfail(d, MTH_ERR_CUDA, std::string("upload of the compressed window: ") + cudaGetErrorString(r));
    d->staged_bytes[slot] = comp_bytes;
    return MTH_OK;
}

// comp == nullptr: the window's compressed bytes are the ones staged in slot `comp_bytes` (0 / 1) by mth_bamdec_stage.
int mth_bamdec_window(mth_bamdec* d, const uint8_t* comp, size_t comp_bytes, const mth_bgzf_member* members, int64_t n_members, uint64_t skip,
                      int last, mth_bamdec_result* out) {
    if (!d || !out || (n_members && !members) || n_members < 0) return MTH_ERR_INVALID;
    const uint8_t* comp_dev = nullptr;
    if (!comp) {
        if (comp_bytes > 1) return MTH_ERR_INVALID;
        comp_dev = (const uint8_t*)d->comp_slot[comp_bytes].p;
        comp_bytes = d->staged_bytes[comp_bytes];
    }
    memset(out, 0, sizeof(*out));
    out->bad_record = -1;
    DTRY(d, cudaSetDevice(d->device));
    cudaStream_t s = d->s;
    // ---- stage + inflate ----
    std::vector<MemberDesc> md((size_t)n_members);
    size_t uo = d->carry;
    for (int64_t i = 0; i < n_members; i++) {
        if (members[i].offset + members[i].size > comp_bytes) return dfail(d, MTH_ERR_INVALID, "BGZF member outside the compressed window");
        md[(size_t)i] = MemberDesc{members[i].offset, members[i].size, members[i].isize, uo, members[i].crc, members[i].flags & 1u};
        uo += members[i].isize;
    }
    const size_t u_end = uo;
    if (u_end >= 0xfffffff0ull) return dfail(d, MTH_ERR_UNSUPPORTED, "window larger than 4 GiB of uncompressed BAM");
    if (skip > u_end - d->carry || (skip && d->carry)) return dfail(d, MTH_ERR_INVALID, "skip outside the first window");
    if ((!comp_dev && reserve(d, d->comp, comp_bytes + 64)) || reserve(d, d->members, (size_t)n_members * sizeof(MemberDesc) + 64) ||
        reserve(d, d->status, (size_t)n_members * 4 + 64) || reserve(d, d->u, u_end + 64, d->carry))
        return MTH_ERR_CUDA;
    DTRY(d, cudaMemsetAsync(d->small.p, 0, SM_WORDS * 8, s));
    if (n_members) {
        if (!comp_dev) {
            DTRY(d, cudaMemcpyAsync(d->comp.p, comp, comp_bytes, cudaMemcpyHostToDevice, s));
            comp_dev = (const uint8_t*)d->comp.p;
        }
        DTRY(d, cudaMemcpyAsync(d->members.p, md.data(), (size_t)n_members * sizeof(MemberDesc), cudaMemcpyHostToDevice, s));
        DTRY(d, cudaEventRecord(d->ev[0], s));
        k_bgzf_inflate<<<(unsigned)((n_members + INF_WARPS - 1) / INF_WARPS), INF_WARPS * 32, INF_WARPS * sizeof(InflateTables), s>>>(
            comp_dev, (const MemberDesc*)d->members.p, n_members, (uint8_t*)d->u.p, (int*)d->status.p,
            (int*)((unsigned long long*)d->small.p + SM_ANYBAD));
    } else {
        DTRY(d, cudaEventRecord(d->ev[0], s));
    }
    DTRY(d, cudaEventRecord(d->ev[1], s));
    // ---- record boundaries: speculate, verify, (repair) ----
    unsigned long long* sm = (unsigned long long*)d->small.p;
    const uint8_t* u = (const uint8_t*)d->u.p;
    const RefTable rt{d->n_ref, (const int64_t*)d->ref_len.p};
    const int64_t n_chunks = (int64_t)((u_end + REC_CHUNK - 1) / REC_CHUNK);
    if (reserve(d, d->entry, (size_t)(n_chunks + 1) * 4) || reserve(d, d->n_rec, (size_t)(n_chunks + 1) * 4) ||
        reserve(d, d->landing, (size_t)(n_chunks + 1) * 4) || reserve(d, d->base, (size_t)(n_chunks + 1) * 4) ||
        reserve(d, d->scan_scratch, (size_t)((std::max<int64_t>(n_chunks, 1) + 2047) / 2048 + 2) * 4 + 64))
        return MTH_ERR_CUDA;
    uint64_t n_records = 0, carry_from = u_end;
    if (n_chunks > 0 && u_end > skip) {
        k_rec_entry<<<(unsigned)n_chunks, 256, 0, s>>>(u, u_end, skip, rt, n_chunks, (uint32_t*)d->entry.p);
        k_rec_walk<false><<<(unsigned)((n_chunks + 127) / 128), 128, 0, s>>>(u, u_end, n_chunks, (const uint32_t*)d->entry.p, (uint32_t*)d->n_rec.p,
                                                                            (uint32_t*)d->landing.p, (int*)(sm + SM_MISMATCH), nullptr, nullptr);
        DTRY(d, cudaMemcpyAsync(d->h_small, d->small.p, 64, cudaMemcpyDeviceToHost, s));
        DTRY(d, cudaStreamSynchronize(s));
        const unsigned long long* hs = (const unsigned long long*)d->h_small;
        if ((int)hs[SM_ANYBAD]) return dfail(d, MTH_ERR_INVALID, "BGZF inflate / CRC check failed on the device (corrupt member)");
        if ((int)hs[SM_MISMATCH]) {
            d->repairs++;
            k_rec_repair<<<1, 32, 0, s>>>(u, u_end, n_chunks, (uint32_t*)d->entry.p, (uint32_t*)d->n_rec.p, (uint32_t*)d->landing.p);
        }
        DTRY(d, cudaMemcpyAsync(d->base.p, d->n_rec.p, (size_t)n_chunks * 4, cudaMemcpyDeviceToDevice, s));
        launch_exclusive_scan_u32((uint32_t*)d->base.p, n_chunks, (uint32_t*)d->scan_scratch.p, sm + SM_TOT_REC, s);
        k_rec_tail<<<1, 32, 0, s>>>((const uint32_t*)d->landing.p, n_chunks, skip, sm + SM_CARRY);
        DTRY(d, cudaMemcpyAsync(d->h_small, d->small.p, 64, cudaMemcpyDeviceToHost, s));
        DTRY(d, cudaStreamSynchronize(s));
        n_records = ((const unsigned long long*)d->h_small)[SM_TOT_REC];
        carry_from = ((const unsigned long long*)d->h_small)[SM_CARRY];
        if (reserve(d, d->rec_off, (size_t)(n_records + 1) * 4)) return MTH_ERR_CUDA;
        k_rec_walk<true><<<(unsigned)((n_chunks + 127) / 128), 128, 0, s>>>(u, u_end, n_chunks, (const uint32_t*)d->entry.p, nullptr, nullptr, nullptr,
                                                                           (const uint32_t*)d->base.p, (uint32_t*)d->rec_off.p);
    } else if (u_end > skip) {
        carry_from = skip;
    }
    DTRY(d, cudaEventRecord(d->ev[2], s));
    if (last && carry_from < u_end) return dfail(d, MTH_ERR_INVALID, "truncated BAM: the file ends inside a record");
    // ---- decode ----
    d->runs.clear();
    const int64_t R = (int64_t)n_records;
    if (R > 0) {
        if (reserve(d, d->keep, (size_t)(R + 1) * 4) || reserve(d, d->ncpg, (size_t)(R + 1) * 4) ||
            reserve(d, d->scan_scratch, (size_t)((R + 2047) / 2048 + 2) * 4 + 64))
            return MTH_ERR_CUDA;
        DecodeArgs a;
        memset(&a, 0, sizeof(a));
        a.u = u; a.rec_off = (const uint32_t*)d->rec_off.p; a.n_rec = R; a.rt = rt; a.lpmd_order = d->lpmd_order; a.min_qual = d->min_qual;
        a.keep = (uint32_t*)d->keep.p; a.ncpg = (uint32_t*)d->ncpg.p; a.counters = sm + SM_COUNTERS; a.marks = sm + SM_MARKS;
        a.n_marks = (int*)(sm + SM_NMARKS);
        DTRY(d, cudaMemsetAsync(sm + SM_COUNTERS + 4, 0xff, 8, s));  // first bad record: atomicMin
        k_rec_decode<false><<<(unsigned)((R + 127) / 128), 128, 0, s>>>(a);
        // keep / ncpg -> exclusive scans (in place; the flags are recovered as differences by the emit pass through a copy)
        if (reserve(d, d->entry, (size_t)(R + 1) * 4) || reserve(d, d->landing, (size_t)(R + 1) * 4)) return MTH_ERR_CUDA;
        uint32_t* keep_scan = (uint32_t*)d->entry.p;    // reuse: the chain buffers are done
        uint32_t* call_scan = (uint32_t*)d->landing.p;
        DTRY(d, cudaMemcpyAsync(keep_scan, d->keep.p, (size_t)R * 4, cudaMemcpyDeviceToDevice, s));
        DTRY(d, cudaMemcpyAsync(call_scan, d->ncpg.p, (size_t)R * 4, cudaMemcpyDeviceToDevice, s));
        launch_exclusive_scan_u32(keep_scan, R, (uint32_t*)d->scan_scratch.p, sm + SM_TOT_KEEP, s);
        launch_exclusive_scan_u32(call_scan, R, (uint32_t*)d->scan_scratch.p, sm + SM_TOT_CALLS, s);
        DTRY(d, cudaMemcpyAsync(d->h_small, d->small.p, SM_WORDS * 8, cudaMemcpyDeviceToHost, s));
        DTRY(d, cudaStreamSynchronize(s));
        const unsigned long long* hs = (const unsigned long long*)d->h_small;
        const unsigned long long* cnt = hs + SM_COUNTERS;
        out->n_dropped = (int64_t)cnt[0];
        out->n_dropped_mapq_ok = (int64_t)cnt[1];
        out->max_span = (int64_t)cnt[2];
        out->max_cpgs = (int32_t)cnt[3];
        if (cnt[4] != ~0ull) {
            out->bad_record = (int64_t)(cnt[4] >> 1);
            out->bad_is_corrupt = (int32_t)(cnt[4] & 1ull);
        }
        if (cnt[5]) return dfail(d, MTH_ERR_UNSUPPORTED, "a read has more than 64 CpG calls or a query index beyond 65535 (device decode limit)");
        const int n_marks = (int)hs[SM_NMARKS];
        if (n_marks > MAX_RUN_MARKS) return dfail(d, MTH_ERR_UNSUPPORTED, "more than 1024 contig changes in one window (device decode limit)");
        const int64_t n_keep = (int64_t)hs[SM_TOT_KEEP], n_calls = (int64_t)hs[SM_TOT_CALLS];
        if (out->bad_record < 0 && n_keep > 0) {
            std::vector<unsigned long long> mk(hs + SM_MARKS, hs + SM_MARKS + n_marks);
            std::sort(mk.begin(), mk.end(), [](unsigned long long x, unsigned long long y) { return (uint32_t)x < (uint32_t)y; });
            std::vector<uint32_t> marks;
            std::vector<int32_t> run_tid;
            for (unsigned long long x : mk) { marks.push_back((uint32_t)x); run_tid.push_back((int32_t)(uint32_t)(x >> 32)); }
            marks.push_back((uint32_t)R);
            // exclusive scan values at the marks (a handful of 4-byte reads)
            std::vector<uint32_t> ks(marks.size()), cs(marks.size());
            for (size_t k = 0; k + 1 < marks.size(); k++) {
                DTRY(d, cudaMemcpyAsync(&ks[k], keep_scan + marks[k], 4, cudaMemcpyDeviceToHost, s));
                DTRY(d, cudaMemcpyAsync(&cs[k], call_scan + marks[k], 4, cudaMemcpyDeviceToHost, s));
            }
            DTRY(d, cudaStreamSynchronize(s));
            ks.back() = (uint32_t)n_keep;
            cs.back() = (uint32_t)n_calls;
            const size_t n_runs = marks.size() - 1;
            if (reserve(d, d->o_start, (size_t)n_keep * 4 + 64) || reserve(d, d->o_end, (size_t)n_keep * 4 + 64) ||
                reserve(d, d->o_meta, (size_t)n_keep * 4 + 64) || reserve(d, d->o_off, (size_t)(n_keep + n_runs + 1) * 4 + 64) ||
                reserve(d, d->o_meth, (size_t)n_keep * 8 + 64) || reserve(d, d->o_pos, (size_t)n_calls * 4 + 64 * n_runs + 64) ||
                reserve(d, d->o_rel, (size_t)n_calls * 2 + 32 * n_runs + 64))
                return MTH_ERR_CUDA;
            a.keep_scan = keep_scan;
            a.call_scan = call_scan;
            for (size_t k = 0; k < n_runs; k++) {
                const int64_t nr = (int64_t)ks[k + 1] - ks[k], nc = (int64_t)cs[k + 1] - cs[k];
                if (nr <= 0) continue;
                // every run gets 16-byte aligned call arrays (the engine's TMA staging wants them) and its own offset slots
                const size_t call_at = (((size_t)cs[k] + 15 * k) + 15) & ~(size_t)15;
                a.rec_lo = marks[k]; a.rec_hi = marks[k + 1]; a.read_base = ks[k]; a.call_base = cs[k]; a.run_reads = (uint32_t)nr;
                a.start = (int32_t*)d->o_start.p + ks[k]; a.end = (int32_t*)d->o_end.p + ks[k]; a.meta = (uint32_t*)d->o_meta.p + ks[k];
                a.off = (uint32_t*)d->o_off.p + ks[k] + k; a.meth = (unsigned long long*)d->o_meth.p + ks[k];
                a.pos = (int32_t*)d->o_pos.p + call_at; a.rel = (uint16_t*)d->o_rel.p + call_at;
                const int64_t nrec = a.rec_hi - a.rec_lo;
                k_rec_decode<true><<<(unsigned)((nrec + 127) / 128), 128, 0, s>>>(a);
                mth_batch b;
                memset(&b, 0, sizeof(b));
                b.tid = run_tid[k]; b.mem_kind = 2; b.n_reads = nr; b.n_cpg = nc; b.n_meth_words = nr;
                b.start = a.start; b.end = a.end; b.meta = a.meta; b.cpg_off = a.off; b.cpg_pos = a.pos; b.cpg_rel = a.rel;
                b.meth = (const uint64_t*)a.meth; b.meth_off = nullptr;
                d->runs.push_back(b);
            }
        }
    }
    DTRY(d, cudaEventRecord(d->ev[3], s));
    // ---- carry the trailing partial record to the front of the buffer for the next window ----
    const size_t new_carry = u_end - carry_from;
    if (new_carry && carry_from) {
        // the decode kernels read `u`: the move happens behind them on the same stream; source and destination may overlap
        // only if the partial record is longer than what precedes it — go through the compressed staging buffer then
        if (new_carry <= carry_from) {
            DTRY(d, cudaMemcpyAsync(d->u.p, (const uint8_t*)d->u.p + carry_from, new_carry, cudaMemcpyDeviceToDevice, s));
        } else {
            if (reserve(d, d->comp, new_carry + 64)) return MTH_ERR_CUDA;
            DTRY(d, cudaMemcpyAsync(d->comp.p, (const uint8_t*)d->u.p + carry_from, new_carry, cudaMemcpyDeviceToDevice, s));
            DTRY(d, cudaMemcpyAsync(d->u.p, d->comp.p, new_carry, cudaMemcpyDeviceToDevice, s));
        }
    }
    d->carry = new_carry;
    DTRY(d, cudaStreamSynchronize(s));
    DTRY(d, cudaGetLastError());
    float ms = 0;
    if (cudaEventElapsedTime(&ms, d->ev[0], d->ev[1]) == cudaSuccess) d->ms_inflate += ms;
    if (cudaEventElapsedTime(&ms, d->ev[1], d->ev[2]) == cudaSuccess) d->ms_boundaries += ms;
    if (cudaEventElapsedTime(&ms, d->ev[2], d->ev[3]) == cudaSuccess) d->ms_decode += ms;
    out->n_records = (int64_t)n_records;
    out->n_runs = (int32_t)d->runs.size();
    out->runs = d->runs.data();
    out->uncompressed_bytes = (uint64_t)u_end;
    out->ms_inflate = d->ms_inflate; out->ms_boundaries = d->ms_boundaries; out->ms_decode = d->ms_decode;
    out->chain_repairs = d->repairs;
    return MTH_OK;
}

}  // extern "C"
