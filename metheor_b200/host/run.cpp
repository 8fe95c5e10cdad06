// run.cpp — one reference subcommand end to end: decode windows of the alignment file on all host cores into pinned
// SoA batches, stream them to the GPU engine(s) through the C ABI of include/metheor_b200.h, fetch the rows and write
// the reference's TSV formats (pdr.rs:102-116, mhl.rs:122-131, fdrp.rs:169-172, qfdrp.rs:181-184, pm.rs:53-60,
// me.rs:57-65, lpmd.rs:89-122,145-147).
//
// Pipeline per window (~64 MiB of uncompressed records):
//   inflate BGZF blocks (all cores) -> walk record boundaries -> cut at contig changes -> decode records (all cores,
//   one SoaChunk per task) -> gather the chunks into one pinned batch (all cores) -> mth_submit (asynchronous H2D on
//   the engine's copy stream + ingest kernels) -> next window while the GPU works.
// Multi-GPU (--gpus N): contigs are assigned to N engine contexts (longest-first balancing on the header lengths);
// a contig's reads all go to one GPU, so every per-CpG / per-quartet row is final on its GPU and only LPMD's four
// counters are summed across contexts.
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <future>
#include <memory>

#include <zlib.h>

#include "../../include/metheor_b200.h"
#include "../../include/metheor_host.h"
#include "decode.hpp"
#include "input.hpp"

namespace mthh {

int format_f32(float v, char* buf, int cap) {
    // Rust `{}` for f32: shortest digits that round-trip, positional notation, "NaN", "inf", "-0"
    if (std::isnan(v)) return snprintf(buf, (size_t)cap, "NaN");
    if (std::isinf(v)) return snprintf(buf, (size_t)cap, v < 0 ? "-inf" : "inf");
    auto r = std::to_chars(buf, buf + cap - 1, v, std::chars_format::fixed);
    *r.ptr = 0;
    return (int)(r.ptr - buf);
}

namespace {

const size_t WINDOW_BYTES = 64u << 20;
const int64_t SHARD_HALO = 65536;  // >= the engine's longest accepted reference span (65 024) + 2
const size_t TASK_RECORDS = 4096;
const int RING = 3;

struct PinnedBatch {
    enum { START, END, META, OFF, POS, REL, METH, MOFF, SPAN, MAPQ, NCPG, FLAGS, DELTA, BITS, RELX, SOFF16, BLKSTART, STARTEXC, DELTA8,
           BLKCALL, DELTA16, NARR };
    void* p[NARR] = {nullptr};
    size_t cap[NARR] = {0};
    ~PinnedBatch() {
        for (void* x : p) mth_host_free(x);
    }
    void reserve(int k, size_t bytes) {
        if (bytes <= cap[k]) return;
        mth_host_free(p[k]);
        size_t ncap = bytes + bytes / 4 + 4096;
        p[k] = mth_host_alloc(ncap);
        if (!p[k]) throw HostError{1, "cannot allocate pinned host memory"};
        cap[k] = ncap;
    }
};

using Interval = ShardInterval;  // sites [lo, hi) of contig tid are OWNED by one GPU

struct Gpu {
    mth_ctx* ctx = nullptr;
    std::vector<Interval> own;  // ascending (tid, lo)
    PinnedBatch ring[RING];
    int next_slot = 0;
    int64_t submitted = 0;
    mth_results res;
    mth_stats stats;
    int rc = 0;
    std::string err;
};

[[noreturn]] void engine_fail(mth_ctx* ctx, int rc, const char* what) {
    std::string m = std::string("metheor_b200 engine: ") + what + " failed (" + std::to_string(rc) + "): " + mth_last_error(ctx);
    if (rc == MTH_ERR_UNSORTED) m += "\nThe GPU engine needs coordinate-sorted input (samtools sort).";
    throw HostError{1, m};
}

uint32_t measure_bit(int m) {
    static const uint32_t bits[] = {MTH_PDR, MTH_LPMD, MTH_MHL, MTH_PM, MTH_ME, MTH_FDRP, MTH_QFDRP};
    return bits[m];
}

// Gathers the decoded chunks [c0, c1) (all of one contig) into the pinned arrays of `pb` and fills `b`.
// own_lo / own_hi: reads starting outside [own_lo, own_hi) are halo copies for this GPU (MTH_META_HALO).
void assemble(ThreadPool& pool, std::vector<SoaChunk>& chunks, size_t c0, size_t c1, int32_t tid, bool want_rel, int max_cpgs,
              int64_t own_lo, int64_t own_hi, PinnedBatch& pb, mth_batch* b) {
    const size_t nc = c1 - c0;
    std::vector<size_t> r_off(nc + 1, 0), i_off(nc + 1, 0), w_off(nc + 1, 0);
    const bool multiword = max_cpgs > 64;
    for (size_t k = 0; k < nc; k++) {
        const SoaChunk& ch = chunks[c0 + k];
        r_off[k + 1] = r_off[k] + ch.start.size();
        i_off[k + 1] = i_off[k] + ch.cpg_pos.size();
        size_t w = ch.start.size();
        if (multiword) {
            w = 0;
            for (uint32_t n : ch.n_cpg) w += std::max<size_t>(1, (n + 63) / 64);
        }
        w_off[k + 1] = w_off[k] + w;
    }
    const size_t R = r_off[nc], I = i_off[nc], W = w_off[nc];
    if (I >= 0xFFFFFFF0ull) throw HostError{1, "more than 2^32 CpG calls in one window"};
    pb.reserve(PinnedBatch::START, R * 4); pb.reserve(PinnedBatch::END, R * 4); pb.reserve(PinnedBatch::META, R * 4);
    pb.reserve(PinnedBatch::OFF, (R + 1) * 4); pb.reserve(PinnedBatch::POS, I * 4 + 64); pb.reserve(PinnedBatch::METH, W * 8 + 8);
    if (want_rel) pb.reserve(PinnedBatch::REL, I * 2 + 64);
    if (multiword) pb.reserve(PinnedBatch::MOFF, (R + 1) * 4);
    int32_t* start = (int32_t*)pb.p[PinnedBatch::START];
    int32_t* end = (int32_t*)pb.p[PinnedBatch::END];
    uint32_t* meta = (uint32_t*)pb.p[PinnedBatch::META];
    uint32_t* off = (uint32_t*)pb.p[PinnedBatch::OFF];
    int32_t* pos = (int32_t*)pb.p[PinnedBatch::POS];
    uint16_t* rel = (uint16_t*)pb.p[PinnedBatch::REL];
    uint64_t* meth = (uint64_t*)pb.p[PinnedBatch::METH];
    uint32_t* moff = (uint32_t*)pb.p[PinnedBatch::MOFF];
    pool.run((int64_t)nc, [&](int64_t k, int) {
        const SoaChunk& ch = chunks[c0 + (size_t)k];
        const size_t r0 = r_off[(size_t)k], i0 = i_off[(size_t)k], n = ch.start.size();
        if (!n) return;
        memcpy(start + r0, ch.start.data(), n * 4);
        memcpy(end + r0, ch.end.data(), n * 4);
        for (size_t r = 0; r < n; r++)
            meta[r0 + r] = (ch.meta[r] & ~SOA_META_COMPLEX) | ((ch.start[r] < own_lo || ch.start[r] >= own_hi) ? (uint32_t)MTH_META_HALO : 0u);
        memcpy(pos + i0, ch.cpg_pos.data(), ch.cpg_pos.size() * 4);
        if (want_rel) memcpy(rel + i0, ch.cpg_rel.data(), ch.cpg_rel.size() * 2);
        size_t io = i0, wo = w_off[(size_t)k];
        const uint8_t* m = ch.cpg_meth.data();
        for (size_t r = 0; r < n; r++) {
            const uint32_t nr = ch.n_cpg[r];
            off[r0 + r] = (uint32_t)io;
            if (multiword) moff[r0 + r] = (uint32_t)wo;
            const size_t nw = multiword ? std::max<size_t>(1, (nr + 63) / 64) : 1;
            for (size_t w = 0; w < nw; w++) {
                uint64_t bits = 0;
                const uint32_t lo = (uint32_t)w * 64, hi = std::min(nr, lo + 64);
                for (uint32_t x = lo; x < hi; x++) bits |= (uint64_t)m[x] << (x - lo);
                meth[wo + w] = bits;
            }
            m += nr;
            io += nr;
            wo += nw;
        }
    });
    off[R] = (uint32_t)I;
    if (multiword) moff[R] = (uint32_t)W;
    memset(b, 0, sizeof(*b));
    b->tid = tid;
    b->mem_kind = 0;
    b->n_reads = (int64_t)R;
    b->n_cpg = (int64_t)I;
    b->n_meth_words = (int64_t)W;
    b->start = start; b->end = end; b->meta = meta; b->cpg_off = off; b->cpg_pos = pos;
    b->cpg_rel = want_rel ? rel : nullptr;
    b->meth = meth;
    b->meth_off = multiword ? moff : nullptr;
}

// Same reads in the compact wire format (mth_batch_compact): 9 B per read + 2.125 B per call over PCIe instead of 24 + 6.
// dense = also the block encodings MTH_CENC_START16 | MTH_CENC_DELTA8 (7 B per read + 1.125 B per call on short-read data).
void assemble_compact(ThreadPool& pool, std::vector<SoaChunk>& all_chunks, size_t c0, size_t c1, int32_t tid, bool want_rel, bool dense,
                      int64_t own_lo, int64_t own_hi, PinnedBatch& pb, mth_batch_compact* b) {
    const size_t nc = c1 - c0;
    SoaChunk* chunks = all_chunks.data() + c0;
    std::vector<size_t> r_off(nc + 1, 0), i_off(nc + 1, 0), e_off(nc + 1, 0);
    for (size_t k = 0; k < nc; k++) {
        const SoaChunk& ch = chunks[k];
        r_off[k + 1] = r_off[k] + ch.start.size();
        i_off[k + 1] = i_off[k] + ch.cpg_pos.size();
        size_t e = 0;
        if (want_rel)
            for (size_t r = 0; r < ch.meta.size(); r++)
                if (ch.meta[r] & SOA_META_COMPLEX) e += ch.n_cpg[r];
        e_off[k + 1] = e_off[k] + e;
    }
    const size_t R = r_off[nc], I = i_off[nc], E = e_off[nc];
    pb.reserve(PinnedBatch::START, R * 4); pb.reserve(PinnedBatch::SPAN, R * 2); pb.reserve(PinnedBatch::MAPQ, R);
    pb.reserve(PinnedBatch::NCPG, R); pb.reserve(PinnedBatch::FLAGS, R); pb.reserve(PinnedBatch::DELTA, I * 2 + 64);
    pb.reserve(PinnedBatch::BITS, (I + 7) / 8 + 64); pb.reserve(PinnedBatch::RELX, E * 2 + 64);
    int32_t* start = (int32_t*)pb.p[PinnedBatch::START];
    uint16_t* span = (uint16_t*)pb.p[PinnedBatch::SPAN];
    uint8_t* mapq = (uint8_t*)pb.p[PinnedBatch::MAPQ];
    uint8_t* ncpg = (uint8_t*)pb.p[PinnedBatch::NCPG];
    uint8_t* flags = (uint8_t*)pb.p[PinnedBatch::FLAGS];
    uint16_t* delta = (uint16_t*)pb.p[PinnedBatch::DELTA];
    uint8_t* bits = (uint8_t*)pb.p[PinnedBatch::BITS];
    uint16_t* relx = (uint16_t*)pb.p[PinnedBatch::RELX];
    for (size_t k = 0; k <= nc; k++) bits[i_off[k] >> 3] = 0;  // bytes shared by two chunks are OR-ed atomically below
    pool.run((int64_t)nc, [&](int64_t k, int) {
        const SoaChunk& ch = chunks[(size_t)k];
        const size_t r0 = r_off[(size_t)k], n = ch.start.size();
        if (!n) return;
        memcpy(start + r0, ch.start.data(), n * 4);
        size_t x = i_off[(size_t)k], e = e_off[(size_t)k];
        const size_t x_end = i_off[(size_t)k + 1];
        const size_t first_byte = x >> 3, last_byte = x_end >> 3;
        for (size_t by = first_byte + 1; by < last_byte; by++) bits[by] = 0;
        size_t c = 0;  // call index within the chunk
        for (size_t r = 0; r < n; r++) {
            const uint32_t m = ch.meta[r], nr = ch.n_cpg[r];
            const int32_t s = ch.start[r];
            span[r0 + r] = (uint16_t)(ch.end[r] - s);
            mapq[r0 + r] = (uint8_t)(m & 0xFFu);
            ncpg[r0 + r] = (uint8_t)nr;
            const bool cx = want_rel && (m & SOA_META_COMPLEX);
            flags[r0 + r] = (uint8_t)(((m >> 8) & 1u) | (cx ? MTH_CFLAG_REL_EXPLICIT : 0u) | ((s < own_lo || s >= own_hi) ? MTH_CFLAG_HALO : 0u));
            int32_t prev = s - 1;
            for (uint32_t q = 0; q < nr; q++, c++, x++) {
                // dense: delta from the previous call of the read (MTH_CENC_DELTA8); plain: offset from start - 1
                delta[x] = (uint16_t)(ch.cpg_pos[c] - (dense ? prev : s - 1));
                prev = ch.cpg_pos[c];
                if (ch.cpg_meth[c]) {
                    const size_t by = x >> 3;
                    const uint8_t bit = (uint8_t)(1u << (x & 7));
                    if (by == first_byte || by == last_byte) __atomic_fetch_or(&bits[by], bit, __ATOMIC_RELAXED);
                    else bits[by] |= bit;
                }
                if (cx) relx[e++] = ch.cpg_rel[c];
            }
        }
    });
    memset(b, 0, sizeof(*b));
    b->tid = tid;
    b->mem_kind = 0;
    b->n_reads = (int64_t)R;
    b->n_cpg = (int64_t)I;
    b->n_rel = (int64_t)E;
    b->start = start; b->span = span; b->mapq = mapq; b->n_cpg8 = ncpg; b->flags = flags;
    b->cpg_delta = delta; b->meth_bits = bits; b->rel_exc = relx;
    if (!dense || R == 0) return;

    // ---- block encodings: blocks of MTH_CBLOCK reads; `start` (32-bit) and `delta` (16-bit, chained) are the wide forms ----
    const size_t nb = (R + MTH_CBLOCK - 1) / MTH_CBLOCK;
    pb.reserve(PinnedBatch::SOFF16, R * 2); pb.reserve(PinnedBatch::BLKSTART, nb * 4); pb.reserve(PinnedBatch::BLKCALL, nb * 4);
    pb.reserve(PinnedBatch::DELTA8, I + 64); pb.reserve(PinnedBatch::DELTA16, I * 2 + 64);
    uint16_t* soff = (uint16_t*)pb.p[PinnedBatch::SOFF16];
    int32_t* blk_start = (int32_t*)pb.p[PinnedBatch::BLKSTART];
    uint32_t* blk_call = (uint32_t*)pb.p[PinnedBatch::BLKCALL];
    uint8_t* d8 = (uint8_t*)pb.p[PinnedBatch::DELTA8];
    uint16_t* d16 = (uint16_t*)pb.p[PinnedBatch::DELTA16];
    std::vector<size_t> call0(nb + 1, 0);       // first call of each block
    std::vector<uint8_t> wide_s(nb, 0), wide_c(nb, 0);
    {
        size_t x = 0;
        for (size_t blk = 0; blk < nb; blk++) {
            call0[blk] = x;
            const size_t r1 = std::min(R, (blk + 1) * (size_t)MTH_CBLOCK);
            for (size_t r = blk * MTH_CBLOCK; r < r1; r++) x += ncpg[r];
        }
        call0[nb] = x;
    }
    pool.run((int64_t)nb, [&](int64_t blk, int) {
        const size_t r0 = (size_t)blk * MTH_CBLOCK, r1 = std::min(R, r0 + (size_t)MTH_CBLOCK);
        wide_s[(size_t)blk] = (int64_t)start[r1 - 1] - (int64_t)start[r0] > 65535;  // reads are sorted: first = smallest
        uint16_t mx = 0;
        for (size_t x = call0[(size_t)blk]; x < call0[(size_t)blk + 1]; x++) mx = std::max(mx, delta[x]);
        wide_c[(size_t)blk] = mx > 255;
    });
    size_t n_exc = 0, n8 = 0, n16 = 0;
    std::vector<size_t> off8(nb), off16(nb), exc_idx(nb);
    for (size_t blk = 0; blk < nb; blk++) {
        const size_t ncall = call0[blk + 1] - call0[blk];
        exc_idx[blk] = n_exc;
        if (wide_s[blk]) n_exc++;
        off8[blk] = n8; off16[blk] = n16;
        if (wide_c[blk]) n16 += ncall; else n8 += ncall;
    }
    pb.reserve(PinnedBatch::STARTEXC, n_exc * MTH_CBLOCK * 4 + 64);
    int32_t* start_exc = (int32_t*)pb.p[PinnedBatch::STARTEXC];
    pool.run((int64_t)nb, [&](int64_t blk_, int) {
        const size_t blk = (size_t)blk_;
        const size_t r0 = blk * MTH_CBLOCK, r1 = std::min(R, r0 + (size_t)MTH_CBLOCK);
        if (wide_s[blk]) {
            blk_start[blk] = -(int32_t)(1 + exc_idx[blk]);
            int32_t* e = start_exc + exc_idx[blk] * MTH_CBLOCK;
            for (size_t r = r0; r < r0 + MTH_CBLOCK; r++) e[r - r0] = r < r1 ? start[r] : 0;
            for (size_t r = r0; r < r1; r++) soff[r] = 0;
        } else {
            blk_start[blk] = start[r0];
            for (size_t r = r0; r < r1; r++) soff[r] = (uint16_t)(start[r] - start[r0]);
        }
        const size_t c0 = call0[blk], c1 = call0[blk + 1];
        if (wide_c[blk]) {
            blk_call[blk] = (uint32_t)off16[blk] | 0x80000000u;
            memcpy(d16 + off16[blk], delta + c0, (c1 - c0) * 2);
        } else {
            blk_call[blk] = (uint32_t)off8[blk];
            for (size_t x = c0; x < c1; x++) d8[off8[blk] + (x - c0)] = (uint8_t)delta[x];
        }
    });
    b->enc = MTH_CENC_START16 | MTH_CENC_DELTA8;
    b->start = nullptr;
    b->start_off16 = soff; b->blk_start = blk_start; b->start_exc = start_exc; b->n_start_exc = (int64_t)(n_exc * MTH_CBLOCK);
    b->cpg_delta8 = d8; b->blk_call_off = blk_call; b->cpg_delta = d16;
    b->n_delta8 = (int64_t)n8; b->n_delta16 = (int64_t)n16;
}

// ---- TSV ----------------------------------------------------------------------------------------------------
// Output table: plain text like the reference (pdr.rs:95-101), or — engine extension --format *.gz — the same bytes as BGZF
// (concatenated <= 64 KiB gzip members + the 28-byte EOF marker: readable by gzip, indexable by tabix).
struct OutFile {
    FILE* f = nullptr;
    bool bgzf = false;
    std::string pend;  // bytes not yet compressed
    OutFile(const char* path, bool compress = false) : bgzf(compress) {
        // pdr.rs:95-101: create + truncate; .unwrap() panics when the path cannot be opened
        f = fopen(path, "w");
        if (!f) throw HostError{101, std::string("called `Result::unwrap()` on an `Err` value: cannot open output file ") + path + ": " + strerror(errno)};
        setvbuf(f, nullptr, _IOFBF, 1 << 20);
    }
    ~OutFile() {
        if (f) {
            try { close(); } catch (...) {}
        }
    }
    void raw(const void* p, size_t n) {
        if (fwrite(p, 1, n, f) != n) throw HostError{101, "Error writing to output file."};
    }
    void member(const char* p, size_t n) {  // one BGZF member holding p[0, n), n <= 0xff00
        uint8_t hdr[18] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 'B', 'C', 2, 0, 0, 0};
        std::vector<uint8_t> buf(compressBound((uLong)n) + 64);
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (deflateInit2(&zs, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw HostError{101, "Error writing to output file."};
        zs.next_in = (Bytef*)p; zs.avail_in = (uInt)n; zs.next_out = buf.data(); zs.avail_out = (uInt)buf.size();
        const int rc = deflate(&zs, Z_FINISH);
        const size_t clen = buf.size() - zs.avail_out;
        deflateEnd(&zs);
        if (rc != Z_STREAM_END || clen + 26 > 65536) throw HostError{101, "Error writing to output file."};
        const uint32_t bsize = (uint32_t)(clen + 25), crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef*)p, (uInt)n), isize = (uint32_t)n;
        hdr[16] = (uint8_t)bsize; hdr[17] = (uint8_t)(bsize >> 8);
        uint8_t tail[8];
        for (int k = 0; k < 4; k++) { tail[k] = (uint8_t)(crc >> (8 * k)); tail[4 + k] = (uint8_t)(isize >> (8 * k)); }
        raw(hdr, 18); raw(buf.data(), clen); raw(tail, 8);
    }
    void write(const std::string& s) {
        if (!bgzf) { raw(s.data(), s.size()); return; }
        pend += s;
        size_t o = 0;
        while (pend.size() - o >= 0xff00) { member(pend.data() + o, 0xff00); o += 0xff00; }
        pend.erase(0, o);
    }
    void close() {
        if (!f) return;
        FILE* g = f;
        if (bgzf) {
            if (!pend.empty()) member(pend.data(), pend.size());
            pend.clear();
            static const uint8_t eof[28] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            raw(eof, 28);
        }
        f = nullptr;
        if (fclose(g) != 0) throw HostError{101, "Error writing to output file."};
    }
};

inline void put_i32(std::string& s, int64_t v) {
    char b[24];
    auto r = std::to_chars(b, b + sizeof(b), v);
    s.append(b, (size_t)(r.ptr - b));
}
inline void put_f32(std::string& s, float v) {
    char b[128];
    s.append(b, (size_t)format_f32(v, b, sizeof(b)));
}

// rows [0,n) formatted by `row(i, string&)` on all cores, written in order
template <class RowFn>
void write_rows(ThreadPool& pool, OutFile& out, int64_t n, RowFn&& row) {
    const int64_t CH = 1 << 16;
    const int64_t nch = (n + CH - 1) / CH;
    const int64_t GROUP = (int64_t)pool.size() * 4;
    std::vector<std::string> bufs((size_t)std::min<int64_t>(nch, GROUP));
    for (int64_t g0 = 0; g0 < nch; g0 += GROUP) {
        const int64_t g1 = std::min(nch, g0 + GROUP);
        pool.run(g1 - g0, [&](int64_t k, int) {
            std::string& s = bufs[(size_t)k];
            s.clear();
            const int64_t a = (g0 + k) * CH, b = std::min(n, a + CH);
            for (int64_t i = a; i < b; i++) row(i, s);
        });
        for (int64_t k = 0; k < g1 - g0; k++) out.write(bufs[(size_t)k]);
    }
}

// Row sources of several GPUs merged by ownership: a GPU's rows are sorted by (tid, first position); the rows it OWNS are
// those whose first position lies in one of its intervals, i.e. one contiguous sub-run per interval, and the global order
// is the order of the intervals.
struct Run { int gpu; int64_t lo, hi; int32_t tid; int64_t pos_lo; };
template <class TidOf, class PosOf>
std::vector<Run> owned_runs(const std::vector<std::unique_ptr<Gpu>>& gpus, const std::vector<int64_t>& n_rows, TidOf&& tid_of, PosOf&& pos_of) {
    std::vector<Run> runs;
    for (int g = 0; g < (int)gpus.size(); g++) {
        const int64_t n = n_rows[(size_t)g];
        auto lower = [&](int32_t t, int64_t x) {  // first row with (tid, pos) >= (t, x)
            int64_t a = 0, b = n;
            while (a < b) {
                const int64_t m = (a + b) / 2;
                const int32_t tm = tid_of(g, m);
                if (tm < t || (tm == t && (int64_t)pos_of(g, m) < x)) a = m + 1; else b = m;
            }
            return a;
        };
        for (const Interval& iv : gpus[(size_t)g]->own) {
            // a reverse-strand read at position 0 calls the site at -1 (readutil.rs:338): it belongs to the contig's first interval
            const int64_t a = lower(iv.tid, iv.lo == 0 ? -1 : iv.lo), b = lower(iv.tid, iv.hi);
            if (b > a) runs.push_back(Run{g, a, b, iv.tid, iv.lo});
        }
    }
    std::stable_sort(runs.begin(), runs.end(), [](const Run& x, const Run& y) { return x.tid != y.tid ? x.tid < y.tid : x.pos_lo < y.pos_lo; });
    return runs;
}

void json_escape(FILE* f, const char* s) {
    for (; *s; s++) {
        if (*s == '"' || *s == '\\') fputc('\\', f);
        fputc(*s, f);
    }
}

}  // namespace

// ---- multi-GPU sharding plans (mirrors metheor_b200/shard.py) ------------------------------------------
// bins: the linearised genome is cut into `world` contiguous ranges of equal length; rank r owns the sites in its range.
// region_cost (bases): every contig a rank touches is a region of its own on the GPU, with a fixed cost next to the cost per
// base; each contig gets that many bases of padding in front and the PADDED coordinate is cut evenly (a cut inside the padding
// moves to the contig's first base).  0 = length alone.  `metheor --gpus N` takes it from METHEOR_SHARD_REGION_COST.
std::vector<std::vector<ShardInterval>> plan_bins(const std::vector<int64_t>& ref_len, int world, int64_t region_cost) {
    std::vector<std::vector<ShardInterval>> out((size_t)world);
    const int n_ref = (int)ref_len.size();
    if (n_ref == 0) return out;
    const int64_t P = region_cost > 0 ? region_cost : 0;
    std::vector<int64_t> base((size_t)n_ref + 1, 0), first((size_t)n_ref + 1, 0);  // base: first base of contig t (padded); first: start of its padding
    base[0] = P;
    for (int t = 0; t < n_ref; t++) base[(size_t)t + 1] = base[(size_t)t] + ref_len[(size_t)t] + P;
    for (int t = 0; t <= n_ref; t++) first[(size_t)t] = base[(size_t)t] - P;
    const int64_t total = base[(size_t)n_ref] - P;
    std::vector<std::pair<int, int64_t>> cuts;  // rank r covers [cuts[r], cuts[r+1]) in (tid, pos) order
    cuts.push_back({0, 0});
    for (int r = 1; r < world; r++) {
        const int64_t x = (int64_t)((__int128)total * r / world);
        int tid = (int)(std::upper_bound(first.begin(), first.end(), x) - first.begin()) - 1;
        if (tid > n_ref - 1) tid = n_ref - 1;
        cuts.push_back({tid, std::max<int64_t>(0, x - base[(size_t)tid])});
    }
    cuts.push_back({n_ref - 1, ref_len[(size_t)n_ref - 1]});
    for (size_t r = 1; r < cuts.size(); r++)  // keep the cut list monotone
        if (cuts[r] < cuts[r - 1]) cuts[r] = cuts[r - 1];
    for (int r = 0; r < world; r++) {
        const auto a = cuts[(size_t)r], b = cuts[(size_t)r + 1];
        for (int tid = a.first; tid <= b.first; tid++) {
            const int64_t lo = tid == a.first ? a.second : 0, hi = tid == b.first ? b.second : ref_len[(size_t)tid];
            if (hi > lo) out[(size_t)r].push_back(ShardInterval{tid, lo, hi});
        }
    }
    return out;
}
int64_t shard_region_cost_env() {
    const char* e = getenv("METHEOR_SHARD_REGION_COST");
    return e ? std::max<long long>(0, atoll(e)) : 0;
}
// contigs: whole contigs, longest first onto the least loaded rank
std::vector<std::vector<ShardInterval>> plan_contigs(const std::vector<int64_t>& ref_len, int world) {
    std::vector<std::vector<ShardInterval>> out((size_t)world);
    std::vector<size_t> order(ref_len.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return ref_len[a] > ref_len[b]; });
    std::vector<int64_t> load((size_t)world, 0);
    for (size_t t : order) {
        const size_t g = (size_t)(std::min_element(load.begin(), load.end()) - load.begin());
        out[g].push_back(ShardInterval{(int32_t)t, 0, ref_len[t]});
        load[g] += ref_len[t];
    }
    for (auto& v : out) std::sort(v.begin(), v.end(), [](const ShardInterval& a, const ShardInterval& b) { return a.tid < b.tid; });
    return out;
}

// ---- device-side decode: the host only walks BGZF member headers and ships COMPRESSED bytes ---------------
// Windows of whole members go through mth_bamdec_window (inflate, record boundaries, BismarkRead decode on the GPU) and come
// back as device-resident batches for mth_submit.  Returns false when the file needs the host decoder (a read with more than
// 64 CpG calls, ...): the caller resets the engine and runs the CPU path.
struct DeviceFeedStats {
    double s_stage = 0, s_window = 0, s_submit = 0, s_create = 0, s_reserve = 0, s_wait_stage = 0, s_tail = 0;
    double ms_inflate = 0, ms_boundaries = 0, ms_decode = 0;
    int64_t windows = 0, chain_repairs = 0;
    uint64_t bytes_compressed = 0, bytes_uncompressed = 0;
};

static bool feed_device(const mthh_options& o, const RecordStream& in, mth_ctx* ctx, int device, uint32_t lpmd_order, DecodeCounters* total,
                        int64_t* n_batches, int64_t* n_reads, int64_t* n_cpg, DeviceFeedStats* fs) {
    const Header& hdr = in.header();
    const uint8_t* d = in.file().data();
    const size_t fsize = in.file().size();
    // uncompressed bytes per window: one warp inflates one <= 64 KiB member and 148 SMs x 32 warps of k_bgzf_inflate are
    // resident at a time, so 148 x 32 full members are one wave of that kernel.  A member takes a warp ~11 ms however few
    // are in flight (Huffman decoding is a serial dependency chain), so smaller windows cost the same 11 ms each (measured:
    // 7 windows of half a wave 76 ms, 3 windows of 1.24 waves 47 ms, one launch of 3 waves 35 ms for a 918 MB BAM).
    const size_t WINDOW_U = (size_t)148 * 32 * 65280;
    struct Staged {
        size_t bytes = 0;
        std::vector<mth_bgzf_member> members;
        uint64_t skip = 0;
        bool last = false, ok = true;
        std::string err;
        double seconds = 0;
    };
    Staged st[2];
    size_t coff = 0;           // next compressed offset
    uint64_t hdr_left = in.bam_header_bytes();  // header bytes not skipped yet
    mth_bamdec* dec = nullptr;
    // stage the next window: member walk + upload of its compressed bytes (straight from the page cache) into a device slot
    auto stage = [&](Staged& w, int slot) {
        const double t0 = now_s();
        w.members.clear();
        w.bytes = 0; w.skip = 0; w.ok = true;
        size_t c0 = coff, u = 0, c1 = coff;
        while (coff < fsize && u < WINDOW_U) {
            size_t cdata, clen;
            uint32_t usize, crc;
            const size_t total = bgzf_member_info(d, fsize, coff, &cdata, &clen, &usize, &crc);
            if (!total) { w.ok = false; w.err = std::string("Error opening BAM file. corrupt or truncated BGZF block at offset ") + std::to_string(coff) + ": " + o.input; return; }
            coff += total;
            if (usize == 0) continue;  // empty member (the EOF marker)
            if (hdr_left >= usize) { hdr_left -= usize; c0 = coff; continue; }  // a member that holds only header bytes
            w.members.push_back(mth_bgzf_member{(uint64_t)(cdata - c0), (uint32_t)clen, usize, crc, 1u});
            c1 = cdata + clen;
            u += usize;
        }
        w.skip = hdr_left;
        hdr_left = 0;
        w.last = coff >= fsize;
        if (!w.members.empty()) {
            w.bytes = c1 - c0;
            const int rc = mth_bamdec_stage(dec, slot, d + c0, w.bytes);
            if (rc != MTH_OK) { w.ok = false; w.err = std::string("metheor_b200 engine: ") + mth_bamdec_last_error(dec); }
        }
        w.seconds = now_s() - t0;
    };
    double tc0 = now_s();
    int rc = mth_bamdec_create(&dec, device, (int32_t)hdr.lengths.size(), hdr.lengths.data(), lpmd_order, o.min_qual);
    if (rc != MTH_OK) throw HostError{1, std::string("metheor_b200 engine: ") + mth_bamdec_last_error(nullptr)};
    fs->s_create = now_s() - tc0;
    struct DecGuard { mth_bamdec* d; ~DecGuard() { mth_bamdec_destroy(d); } } guard{dec};
    int cur = 0;
    stage(st[cur], cur);
    bool first = true;
    for (;;) {
        Staged& w = st[cur];
        if (!w.ok) throw HostError{101, w.err};
        fs->s_stage += w.seconds;
        if (w.members.empty() && !first) break;
        const bool last = w.last;
        // the next window is staged by a helper thread while the GPU works on this one
        std::future<void> nxt;
        if (!last) nxt = std::async(std::launch::async, [&, cur] { stage(st[cur ^ 1], cur ^ 1); });
        double t0 = now_s();
        rc = mth_sync_copies(ctx);  // the previous window's device batches have been copied into the arena
        if (rc != MTH_OK) engine_fail(ctx, rc, "mth_sync_copies");
        mth_bamdec_result res;
        rc = mth_bamdec_window(dec, nullptr, (size_t)cur, w.members.data(), (int64_t)w.members.size(), w.skip, last ? 1 : 0, &res);
        if (rc == MTH_ERR_UNSUPPORTED) {
            if (nxt.valid()) nxt.wait();
            return false;
        }
        if (rc != MTH_OK) {
            if (nxt.valid()) nxt.wait();
            throw HostError{101, std::string("Error opening BAM file. ") + mth_bamdec_last_error(dec) + ": " + o.input};
        }
        fs->s_window += now_s() - t0;
        if (res.bad_record >= 0) {
            if (nxt.valid()) nxt.wait();
            if (res.bad_is_corrupt) throw HostError{101, "Error opening BAM file. corrupt BAM record"};
            throw HostError{101, "Error reading XM tag in BAM record. Make sure the reads are aligned using Bismark!"};  // readutil.rs:45-51
        }
        total->n_records += res.n_records;
        total->n_dropped += res.n_dropped;
        total->n_dropped_mapq_ok += res.n_dropped_mapq_ok;
        if (res.max_cpgs > total->max_cpgs) total->max_cpgs = res.max_cpgs;
        if (res.max_span > total->max_span) total->max_span = res.max_span;
        fs->ms_inflate = res.ms_inflate; fs->ms_boundaries = res.ms_boundaries; fs->ms_decode = res.ms_decode;
        fs->chain_repairs = res.chain_repairs;
        fs->bytes_uncompressed += res.uncompressed_bytes;
        fs->bytes_compressed += w.bytes;
        fs->windows++;
        if (first && !last && w.bytes) {  // size the (still empty) arena once: reads per compressed byte of the first window x file size (+15 %)
            int64_t r1 = 0, c1 = 0;
            for (int k = 0; k < res.n_runs; k++) { r1 += res.runs[k].n_reads; c1 += res.runs[k].n_cpg; }
            const double scale = 1.15 * (double)fsize / (double)w.bytes;
            const double tr0 = now_s();
            if (scale > 1.5) mth_reserve(ctx, (int64_t)((double)r1 * scale) + 4096, (int64_t)((double)c1 * scale) + 4096);
            fs->s_reserve += now_s() - tr0;
        }
        t0 = now_s();
        for (int k = 0; k < res.n_runs; k++) {
            const mth_batch& b = res.runs[k];
            if (b.tid < 0 || (size_t)b.tid >= hdr.lengths.size())
                throw HostError{101, "metheor_b200: a read with CpG calls has no valid reference id (tid " + std::to_string(b.tid) + ")"};
            rc = mth_submit(ctx, &b);
            if (rc != MTH_OK) engine_fail(ctx, rc, "mth_submit");
            (*n_batches)++;
            *n_reads += b.n_reads;
            *n_cpg += b.n_cpg;
        }
        fs->s_submit += now_s() - t0;
        first = false;
        const double tw0 = now_s();
        if (nxt.valid()) nxt.wait();
        fs->s_wait_stage += now_s() - tw0;
        if (last) break;
        cur ^= 1;
    }
    const double tt0 = now_s();
    rc = mth_sync_copies(ctx);
    if (rc != MTH_OK) engine_fail(ctx, rc, "mth_sync_copies");
    fs->s_tail = now_s() - tt0;
    return true;
}

void run(const mthh_options& o) {
    const double t_begin = now_s();
    if (!o.input || !o.output) throw HostError{2, "error: the following required arguments were not provided:\n  --input <INPUT>\n  --output <OUTPUT>"};
    if (o.measure < 0 || o.measure > MTHH_QFDRP) throw HostError{2, "error: unknown measure"};
    int n_threads = o.threads > 0 ? o.threads : (int)std::thread::hardware_concurrency();
    if (n_threads < 1) n_threads = 1;
    ThreadPool pool(n_threads);
    // the CUDA runtime takes a few hundred ms to come up: start it now, it overlaps with opening and inflating the file
    std::future<int> cuda_up = std::async(std::launch::async, [] { return mth_device_count(); });

    // bamutil.rs:4-11: the input is opened first; a missing / non-BAM file panics with "Error opening BAM file. ..."
    RecordStream in(o.input, n_threads, WINDOW_BYTES);  // inflates ahead on its own pool while `pool` decodes
    const Header& hdr = in.header();
    CpgSet cpg_set;
    if (o.cpg_set) cpg_set.load(o.cpg_set, hdr);  // readutil.rs:347-374

    const int n_dev = cuda_up.get();
    if (n_dev <= 0) throw HostError{1, "metheor_b200: no CUDA device found (this engine has no CPU path)"};
    const int n_gpus = std::max(1, o.n_gpus);
    if (o.device < 0 || o.device + n_gpus > n_dev)
        throw HostError{1, "metheor_b200: --device/--gpus outside the " + std::to_string(n_dev) + " visible CUDA device(s)"};

    mth_params prm;
    mth_params_default(&prm);
    prm.measures = measure_bit(o.measure);
    prm.seed = o.seed;
    if (o.stats_json) prm.flags |= MTH_FLAG_PROFILE;
    prm.pdr = {o.min_depth, o.min_cpgs, o.min_qual};
    prm.mhl = {o.min_depth, o.min_cpgs, o.min_qual};
    prm.pm = {o.min_depth, o.min_qual};
    prm.me = {o.min_depth, o.min_qual};
    prm.fdrp = {o.min_qual, o.min_depth, o.max_depth, o.min_overlap};
    prm.qfdrp = prm.fdrp;
    prm.lpmd = {o.min_distance, o.max_distance, o.min_qual, o.pairs ? 1u : 0u};
    if ((o.measure == MTHH_FDRP || o.measure == MTHH_QFDRP) && o.max_depth == 0)
        throw HostError{1, "metheor_b200: --max-depth must be at least 1"};

    const std::vector<int64_t>& ref_len = hdr.lengths;
    std::vector<std::unique_ptr<Gpu>> gpus;
    for (int g = 0; g < n_gpus; g++) {
        gpus.emplace_back(new Gpu());
        int rc = mth_ctx_create(&gpus.back()->ctx, o.device + g, &prm, (int32_t)ref_len.size(), ref_len.data());
        if (rc != MTH_OK) throw HostError{1, std::string("metheor_b200 engine: ") + mth_last_error(nullptr)};
    }
    struct CtxGuard {
        std::vector<std::unique_ptr<Gpu>>& g;
        ~CtxGuard() {
            for (auto& x : g) mth_ctx_destroy(x->ctx);
        }
    } guard{gpus};

    // Sharding: position bins of the linearised genome (default) or whole contigs.  A GPU receives the reads
    // that start in its interval plus a halo (reads starting up to SHARD_HALO before it, or exactly at its end); it OWNS the
    // sites inside the interval: every contributor and every flush trigger of an owned site is among its reads, so the rows
    // it reports for them are final.  Halo copies carry MTH_META_HALO so that LPMD counts each read once.
    {
        auto plan = (n_gpus > 1 && o.shard_contigs) ? plan_contigs(ref_len, n_gpus) : plan_bins(ref_len, n_gpus, shard_region_cost_env());
        for (int g = 0; g < n_gpus; g++) gpus[(size_t)g]->own = plan[(size_t)g];
        if (n_gpus > 1) {  // the contexts are joined by one NCCL communicator: LPMD's counters are summed at the end
            std::vector<mth_ctx*> cs;
            for (auto& gp : gpus) cs.push_back(gp->ctx);
            int rc = mth_comm_init_all(cs.data(), n_gpus);
            if (rc != MTH_OK) engine_fail(cs[0], rc, "mth_comm_init_all");
        }
    }

    DecodeOptions dopt;
    // --cpg-set (readutil.rs:87-95, 347-374): the BED positions go to the engine once and the filter runs on the device,
    // before anything else is computed from a read; the decoder ships unfiltered calls.  METHEOR_CPGSET_HOST=1 keeps the
    // filter in the host decoder instead (kernel-variant experiments).
    const bool set_on_host = o.cpg_set && getenv("METHEOR_CPGSET_HOST") != nullptr;
    if (o.cpg_set && !set_on_host) {
        std::vector<int32_t> st, sp;
        for (size_t t = 0; t < cpg_set.by_tid.size(); t++)
            for (int32_t p : cpg_set.by_tid[t]) { st.push_back((int32_t)t); sp.push_back(p); }
        for (auto& gp : gpus) {
            int rc = mth_set_cpg_set(gp->ctx, (int64_t)st.size(), st.data(), sp.data());
            if (rc != MTH_OK) engine_fail(gp->ctx, rc, "mth_set_cpg_set");
        }
    }
    dopt.cpg_set = set_on_host ? &cpg_set : nullptr;
    dopt.min_qual = o.min_qual;
    dopt.lpmd_order = o.measure == MTHH_LPMD;
    const bool want_rel = o.measure == MTHH_LPMD;

    std::vector<RecordRef> recs;
    std::vector<SoaChunk> chunks;
    std::vector<DecodeCounters> task_cnt;
    DecodeCounters total;
    double s_decode = 0, s_assemble = 0, s_submit = 0;
    int64_t n_batches = 0, n_shipped_reads = 0, n_shipped_cpg = 0;
    bool reserved = false;
    const Format fmt = in.format();

    // --region chr[:beg-end] (engine extension): the region is ONE ownership interval — exactly the multi-GPU
    // machinery with a single shard: reads starting up to SHARD_HALO before it are decoded as halo copies so that every site
    // inside has all its contributors and flush triggers, rows outside are dropped, LPMD counts the reads that start inside.
    bool has_region = false;
    Interval region{0, 0, 0};
    bool region_empty = false;
    if (o.region) {
        if (n_gpus != 1) throw HostError{2, "error: --region cannot be combined with --gpus > 1"};
        std::string r = o.region, name = r;
        int64_t beg = 1, end = INT64_MAX;
        if (hdr.tid_of(name) < 0) {
            const size_t c = r.rfind(':');
            if (c == std::string::npos) throw HostError{101, "unknown chromosome '" + r + "' in --region"};
            name = r.substr(0, c);
            std::string span = r.substr(c + 1);
            span.erase(std::remove(span.begin(), span.end(), ','), span.end());
            const size_t dsh = span.find('-');
            char* ep = nullptr;
            beg = strtoll(span.c_str(), &ep, 10);
            if (ep == span.c_str() || beg < 1) throw HostError{2, "error: invalid --region '" + r + "' (expected chr[:beg-end], 1-based)"};
            if (dsh != std::string::npos) {
                const char* e0 = span.c_str() + dsh + 1;
                end = strtoll(e0, &ep, 10);
                if (ep == e0 || end < beg) throw HostError{2, "error: invalid --region '" + r + "' (expected chr[:beg-end], 1-based)"};
            }
            if (hdr.tid_of(name) < 0) throw HostError{101, "unknown chromosome '" + name + "' in --region"};
        }
        region.tid = hdr.tid_of(name);
        region.lo = beg - 1;
        region.hi = std::min<int64_t>(end, ref_len[(size_t)region.tid]);
        has_region = true;
        gpus[0]->own.assign(1, region);
        // the .bai's linear index (one virtual offset per 16 kb window) says where to start reading
        if (fmt == Format::BAM) {
            std::string p1 = std::string(o.input) + ".bai", p2 = o.input;
            if (p2.size() > 4 && p2.compare(p2.size() - 4, 4, ".bam") == 0) p2 = p2.substr(0, p2.size() - 4) + ".bai";
            MappedFile bai;
            bool have = false;
            for (const std::string& p : {p1, p2}) {
                try { bai.open(p); have = true; break; } catch (const HostError&) {}
            }
            if (have) {
                const uint8_t* b = bai.data();
                const size_t n = bai.size();
                size_t q = 8;
                bool ok = n >= 8 && memcmp(b, "BAI\1", 4) == 0 && le32(b + 4) == (int32_t)ref_len.size();
                for (int32_t t = 0; ok && t <= region.tid; t++) {
                    if (q + 4 > n) { ok = false; break; }
                    const int32_t n_bin = le32(b + q);
                    q += 4;
                    for (int32_t k = 0; ok && k < n_bin; k++) {
                        if (q + 8 > n) { ok = false; break; }
                        const int32_t n_chunk = le32(b + q + 4);
                        q += 8 + (size_t)n_chunk * 16;
                    }
                    if (!ok || q + 4 > n) { ok = false; break; }
                    const int32_t n_intv = le32(b + q);
                    q += 4;
                    if (q + (size_t)n_intv * 8 > n) { ok = false; break; }
                    if (t == region.tid) {
                        auto iv = [&](int32_t w) { uint64_t v; memcpy(&v, b + q + (size_t)w * 8, 8); return v; };
                        if (n_intv == 0) { region_empty = true; break; }
                        int32_t w = (int32_t)std::min<int64_t>(std::max<int64_t>(0, region.lo - SHARD_HALO) >> 14, n_intv - 1);
                        uint64_t v = 0;
                        for (int32_t x = w; x >= 0 && !v; x--) v = iv(x);   // the last window at or before the start that holds a read
                        for (int32_t x = w + 1; x < n_intv && !v; x++) v = iv(x);  // nothing before: the contig's reads start later
                        if (v) in.seek_virtual(v); else region_empty = true;
                    }
                    q += (size_t)n_intv * 8;
                }
                if (!ok) fprintf(stderr, "metheor_b200: ignoring a malformed index next to %s (reading from the start)\n", o.input);
            }
        }
    }
    const bool sharded = n_gpus > 1 || has_region;

    // BAM on one GPU: inflate + record decode on the device (the host ships compressed bytes); everything else — SAM text,
    // several GPUs, a file beyond the device decoder's limits, --decode host — goes through the CPU decoder below.
    bool used_device = false;
    DeviceFeedStats dfs;
    if (fmt == Format::BAM && n_gpus == 1 && !o.decode_host && !set_on_host && !has_region && !getenv("METHEOR_DECODE_HOST")) {
        used_device = feed_device(o, in, gpus[0]->ctx, o.device, dopt.lpmd_order ? 1u : 0u, &total, &n_batches, &n_shipped_reads, &n_shipped_cpg, &dfs);
        if (!used_device) {  // start over on the host decoder
            int rc = mth_reset(gpus[0]->ctx);
            if (rc != MTH_OK) engine_fail(gpus[0]->ctx, rc, "mth_reset");
            total = DecodeCounters();
            n_batches = n_shipped_reads = n_shipped_cpg = 0;
        }
    }
    bool past_region = region_empty;
    while (!used_device && !past_region && in.next(&recs)) {
        // cut the window at contig changes (a batch carries one tid)
        size_t seg0 = 0;
        while (seg0 < recs.size()) {
            const int32_t tid = record_tid(fmt, hdr, recs[seg0]);
            size_t seg1 = seg0 + 1;
            if (record_tid(fmt, hdr, recs.back()) == tid) {
                seg1 = recs.size();
            } else {
                size_t a = seg0, b = recs.size();  // reads are grouped by contig: first record with another tid
                while (b - a > 1) {
                    size_t m = (a + b) / 2;
                    if (record_tid(fmt, hdr, recs[m]) == tid) a = m; else b = m;
                }
                seg1 = b;
            }
            const size_t n_tasks = (seg1 - seg0 + TASK_RECORDS - 1) / TASK_RECORDS;
            if (chunks.size() < n_tasks) chunks.resize(n_tasks);
            task_cnt.assign(n_tasks, DecodeCounters());
            double t0 = now_s();
            std::atomic<int> failed{0};
            HostError first_err{0, ""};
            std::mutex err_m;
            pool.run((int64_t)n_tasks, [&](int64_t k, int) {
                SoaChunk& ch = chunks[(size_t)k];
                ch.clear();
                if (failed.load()) return;
                const size_t a = seg0 + (size_t)k * TASK_RECORDS, b = std::min(seg1, a + TASK_RECORDS);
                try {
                    decode_records(fmt, hdr, recs.data(), a, b, dopt, &ch, &task_cnt[(size_t)k]);
                } catch (const HostError& e) {
                    std::lock_guard<std::mutex> g(err_m);
                    if (!failed.exchange(1)) first_err = e;
                }
            });
            if (failed.load()) throw first_err;
            for (size_t k = 0; k < n_tasks; k++)  // the cut above assumed reads grouped by contig
                for (int32_t t : chunks[k].tid)
                    if (t != tid) throw HostError{1, "metheor_b200: input is not coordinate-sorted (contigs interleave); sort it with samtools sort"};
            DecodeCounters seg;
            for (auto& c : task_cnt) seg.add(c);
            total.add(seg);
            s_decode += now_s() - t0;

            size_t kept = 0, kept_calls = 0;
            for (size_t k = 0; k < n_tasks; k++) { kept += chunks[k].start.size(); kept_calls += chunks[k].cpg_pos.size(); }
            if (!reserved && kept && fmt == Format::BAM && in.compressed_consumed() > 0 && n_gpus == 1) {
                // size the device arena once from the first window: reads per compressed byte x file size (+15 %)
                const double scale = 1.15 * (double)in.file_size() / (double)in.compressed_consumed();
                if (scale > 1.5) mth_reserve(gpus[0]->ctx, (int64_t)((double)kept * scale) + 4096, (int64_t)((double)kept_calls * scale) + 4096);
                reserved = true;
            }
            if (kept && (tid < 0 || (size_t)tid >= ref_len.size()))
                throw HostError{101, "metheor_b200: a read with CpG calls has no valid reference id (tid " + std::to_string(tid) + ")"};
            // coordinate order within the contig (the engine checks it again on the device; chunk routing below relies on it)
            {
                int32_t prev = INT32_MIN;
                for (size_t k = 0; k < n_tasks; k++) {
                    const auto& st = chunks[k].start;
                    for (size_t r = 0; r < st.size(); r++) {
                        if (st[r] < prev) throw HostError{1, "metheor_b200: input is not coordinate-sorted; sort it with samtools sort"};
                        prev = st[r];
                    }
                }
            }
            for (int g = 0; kept && g < n_gpus; g++) {
                Gpu& G = *gpus[(size_t)g];
                for (const Interval& iv : G.own) {
                    if (iv.tid != tid) continue;
                    // chunks (ascending starts) that hold a read starting in [lo - halo, hi]
                    size_t c0 = n_tasks, c1 = 0;
                    const int64_t need_lo = sharded ? iv.lo - SHARD_HALO : INT64_MIN, need_hi = sharded ? iv.hi : INT64_MAX;
                    for (size_t k = 0; k < n_tasks; k++) {
                        const auto& st = chunks[k].start;
                        if (st.empty() || (int64_t)st.back() < need_lo || (int64_t)st.front() > need_hi) continue;
                        c0 = std::min(c0, k);
                        c1 = std::max(c1, k + 1);
                    }
                    if (c0 >= c1) continue;
                    const int64_t own_lo = sharded ? iv.lo : INT64_MIN, own_hi = sharded ? iv.hi : INT64_MAX;
                    t0 = now_s();
                    // the slot was last used RING submits ago: its copy has long been issued, wait for it to have completed
                    if (G.submitted >= RING) {
                        int rc = mth_sync_copies(G.ctx);
                        if (rc != MTH_OK) engine_fail(G.ctx, rc, "mth_sync_copies");
                    }
                    PinnedBatch& pb = G.ring[G.next_slot];
                    G.next_slot = (G.next_slot + 1) % RING;
                    int64_t nr_b, nc_b;
                    int rc;
                    if (seg.max_cpgs <= 64 && seg.max_span <= 65024) {  // compact wire format: ~1/3 of the PCIe bytes
                        mth_batch_compact b;
                        assemble_compact(pool, chunks, c0, c1, tid, want_rel, /*dense=*/true, own_lo, own_hi, pb, &b);
                        s_assemble += now_s() - t0;
                        t0 = now_s();
                        rc = mth_submit_compact(G.ctx, &b);
                        nr_b = b.n_reads; nc_b = b.n_cpg;
                    } else {
                        mth_batch b;
                        assemble(pool, chunks, c0, c1, tid, want_rel, seg.max_cpgs, own_lo, own_hi, pb, &b);
                        s_assemble += now_s() - t0;
                        t0 = now_s();
                        rc = mth_submit(G.ctx, &b);
                        nr_b = b.n_reads; nc_b = b.n_cpg;
                    }
                    if (rc != MTH_OK) engine_fail(G.ctx, rc, "mth_submit");
                    G.submitted++;
                    s_submit += now_s() - t0;
                    n_batches++;
                    n_shipped_reads += nr_b;
                    n_shipped_cpg += nc_b;
                }
            }
            if (has_region && (tid > region.tid || (tid == region.tid && !chunks.empty() && n_tasks && !chunks[0].start.empty() &&
                                                    (int64_t)chunks[0].start.front() > region.hi)))
                past_region = true;  // coordinate-sorted input: nothing further can matter
            seg0 = seg1;
        }
    }
    const double t_decoded = now_s();

    // reads that carried no CpG call never reach the GPU; LPMD still counts them (lpmd.rs:176, :189)
    mth_add_skipped_reads(gpus[0]->ctx, total.n_dropped, total.n_dropped_mapq_ok);

    // finish every GPU (concurrently: mth_finish blocks on its device)
    {
        std::vector<std::thread> th;
        for (auto& gp : gpus) th.emplace_back([&gp] {
            Gpu& G = *gp;
            G.rc = mth_finish(G.ctx, &G.res);
            if (G.rc != MTH_OK) G.err = mth_last_error(G.ctx);
            mth_get_stats(G.ctx, &G.stats);
        });
        for (auto& t : th) t.join();
        for (auto& gp : gpus)
            if (gp->rc != MTH_OK) engine_fail(gp->ctx, gp->rc, "mth_finish");
    }
    const double t_finished = now_s();

    // ---- output ----
    // engine extensions: --format tsv | tsv.gz | bedgraph | bedgraph.gz  (on-disk formats behind the path)
    const bool bedgraph = (o.out_format & 2) != 0, out_gz = (o.out_format & 1) != 0;
    if (bedgraph && (o.measure == MTHH_LPMD || o.measure == MTHH_PM || o.measure == MTHH_ME))
        throw HostError{2, "error: --format bedgraph needs a per-CpG measure (pdr, mhl, fdrp, qfdrp)"};
    OutFile out(o.output, out_gz);
    auto chrom = [&](int32_t tid) -> const std::string& { return hdr.names[(size_t)tid]; };
    int64_t n_rows_total = 0;
    if (o.measure == MTHH_LPMD) {
        // lpmd.rs:190-191 sums over the whole file: with several GPUs that is the path's ONE collective — an NCCL all-reduce
        // of the four int64 counters inside the library; afterwards every context holds the global result.
        mth_lpmd_result lr = gpus[0]->res.lpmd;
        if (n_gpus > 1) {
            std::vector<mth_ctx*> cs;
            for (auto& gp : gpus) cs.push_back(gp->ctx);
            int rc = mth_allreduce_group(cs.data(), n_gpus);
            if (rc != MTH_OK) engine_fail(cs[0], rc, "mth_allreduce_group");
            rc = mth_lpmd_refresh(gpus[0]->ctx, &lr);
            if (rc != MTH_OK) engine_fail(gpus[0]->ctx, rc, "mth_lpmd_refresh");
        }
        const float lpmd = lr.lpmd;
        std::string s = "name\tlpmd\n";  // lpmd.rs:145-147
        s += o.input;
        s += '\t';
        put_f32(s, lpmd);
        s += '\n';
        out.write(s);
        if (o.pairs) {
            OutFile pf(o.pairs);
            pf.write("chrom\tcpg1\tcpg2\tlpmd\tn_concordant\tn_discordant\n");  // lpmd.rs:104
            std::vector<int64_t> nr;
            for (auto& gp : gpus) nr.push_back(gp->res.lpmd_pairs.n);
            auto runs = owned_runs(gpus, nr, [&](int g, int64_t i) { return gpus[(size_t)g]->res.lpmd_pairs.tid[i]; },
                                   [&](int g, int64_t i) { return gpus[(size_t)g]->res.lpmd_pairs.pos1[i]; });
            for (const Run& r : runs) {
                const mth_pair_rows& R = gpus[(size_t)r.gpu]->res.lpmd_pairs;
                write_rows(pool, pf, r.hi - r.lo, [&](int64_t k, std::string& s2) {
                    const int64_t i = r.lo + k;
                    s2 += chrom(R.tid[i]); s2 += '\t'; put_i32(s2, R.pos1[i]); s2 += '\t'; put_i32(s2, R.pos2[i]); s2 += '\t';
                    put_f32(s2, R.lpmd[i]); s2 += '\t'; put_i32(s2, R.n_conc[i]); s2 += '\t'; put_i32(s2, R.n_disc[i]); s2 += '\n';
                });
                n_rows_total += r.hi - r.lo;
            }
        }
    } else if (o.measure == MTHH_PM || o.measure == MTHH_ME) {
        auto rows_of = [&](Gpu& G) -> const mth_quartet_rows& { return o.measure == MTHH_PM ? G.res.pm : G.res.me; };
        std::vector<int64_t> nr;
        for (auto& gp : gpus) nr.push_back(rows_of(*gp).n);
        auto runs = owned_runs(gpus, nr, [&](int g, int64_t i) { return rows_of(*gpus[(size_t)g]).tid[i]; },
                               [&](int g, int64_t i) { return rows_of(*gpus[(size_t)g]).p1[i]; });
        for (const Run& r : runs) {
            const mth_quartet_rows& R = rows_of(*gpus[(size_t)r.gpu]);
            write_rows(pool, out, r.hi - r.lo, [&](int64_t k, std::string& s) {  // pm.rs:56-59, me.rs:61-64
                const int64_t i = r.lo + k;
                s += chrom(R.tid[i]); s += '\t'; put_i32(s, R.p1[i]); s += '\t'; put_i32(s, R.p2[i]); s += '\t';
                put_i32(s, R.p3[i]); s += '\t'; put_i32(s, R.p4[i]); s += '\t'; put_f32(s, R.value[i]); s += '\n';
            });
            n_rows_total += r.hi - r.lo;
        }
    } else {
        auto rows_of = [&](Gpu& G) -> const mth_site_rows& {
            switch (o.measure) {
                case MTHH_PDR: return G.res.pdr;
                case MTHH_MHL: return G.res.mhl;
                case MTHH_FDRP: return G.res.fdrp;
                default: return G.res.qfdrp;
            }
        };
        const bool counts = o.measure == MTHH_PDR && !bedgraph;  // bedGraph: chrom, start, end, value
        std::vector<int64_t> nr;
        for (auto& gp : gpus) nr.push_back(rows_of(*gp).n);
        auto runs = owned_runs(gpus, nr, [&](int g, int64_t i) { return rows_of(*gpus[(size_t)g]).tid[i]; },
                               [&](int g, int64_t i) { return rows_of(*gpus[(size_t)g]).pos[i]; });
        for (const Run& r : runs) {
            const mth_site_rows& R = rows_of(*gpus[(size_t)r.gpu]);
            write_rows(pool, out, r.hi - r.lo, [&](int64_t k, std::string& s) {  // pdr.rs:105-114, mhl.rs:123-130, fdrp.rs:171
                const int64_t i = r.lo + k;
                s += chrom(R.tid[i]); s += '\t'; put_i32(s, R.pos[i]); s += '\t'; put_i32(s, (int64_t)R.pos[i] + 2); s += '\t';
                put_f32(s, R.value[i]);
                if (counts) { s += '\t'; put_i32(s, R.n_conc[i]); s += '\t'; put_i32(s, R.n_disc[i]); }
                s += '\n';
            });
            n_rows_total += r.hi - r.lo;
        }
    }
    const double t_end = now_s();

    if (o.stats_json) {
        FILE* f = fopen(o.stats_json, "w");
        if (f) {
            const double wall = t_end - t_begin;
            fprintf(f, "{\"input\": \""); json_escape(f, o.input);
            fprintf(f, "\", \"format\": \"%s\", \"threads\": %d, \"gpus\": %d, \"records\": %lld, \"reads_shipped\": %lld, "
                       "\"cpg_calls_shipped\": %lld, \"batches\": %lld, \"rows\": %lld, \"bytes_uncompressed\": %llu, \"zlib_fallbacks\": %lld, "
                       "\"seconds\": {\"total\": %.6f, \"stream\": %.6f, \"inflate\": %.6f, \"walk\": %.6f, \"decode\": %.6f, "
                       "\"assemble\": %.6f, \"submit\": %.6f, \"finish\": %.6f, \"write\": %.6f}, \"reads_per_sec\": %.1f, \"decode\": \"%s\", "
                       "\"device_decode\": {\"windows\": %lld, \"create_s\": %.6f, \"reserve_s\": %.6f, \"wait_stage_s\": %.6f, \"tail_s\": %.6f, \"stage_s\": %.6f, \"window_s\": %.6f, \"submit_s\": %.6f, \"inflate_ms\": %.3f, "
                       "\"boundaries_ms\": %.3f, \"decode_ms\": %.3f, \"chain_repairs\": %lld, \"bytes_compressed\": %llu, \"bytes_uncompressed\": %llu}, \"gpu\": [",
                    fmt == Format::BAM ? "bam" : "sam", n_threads, n_gpus, (long long)total.n_records, (long long)n_shipped_reads,
                    (long long)n_shipped_cpg, (long long)n_batches, (long long)n_rows_total,
                    (unsigned long long)(used_device ? dfs.bytes_uncompressed : in.bytes_uncompressed),
                    (long long)g_zlib_fallbacks.load(), wall, t_decoded - t_begin, in.seconds_inflate, in.seconds_walk, s_decode, s_assemble, s_submit,
                    t_finished - t_decoded, t_end - t_finished, (double)total.n_records / wall, used_device ? "device" : "host",
                    (long long)dfs.windows, dfs.s_create, dfs.s_reserve, dfs.s_wait_stage, dfs.s_tail, dfs.s_stage, dfs.s_window, dfs.s_submit, dfs.ms_inflate, dfs.ms_boundaries, dfs.ms_decode,
                    (long long)dfs.chain_repairs, (unsigned long long)dfs.bytes_compressed, (unsigned long long)dfs.bytes_uncompressed);
            for (size_t g = 0; g < gpus.size(); g++) {
                const mth_stats& st = gpus[g]->stats;
                fprintf(f, "%s{\"reads\": %lld, \"cpg_calls\": %lld, \"sites\": %lld, \"regions\": %lld, \"kernel_launches\": %lld, "
                           "\"h2d_bytes\": %lld, \"d2h_bytes\": %lld, \"kernels\": {",
                        g ? ", " : "", (long long)st.n_reads, (long long)st.n_cpg, (long long)st.n_sites, (long long)st.n_regions,
                        (long long)st.kernel_launches, (long long)st.h2d_bytes, (long long)st.d2h_bytes);
                for (int k = 0; k < st.n_kernel_stats; k++)
                    fprintf(f, "%s\"%s\": {\"launches\": %lld, \"ms\": %.4f}", k ? ", " : "", st.kernel[k].name,
                            (long long)st.kernel[k].launches, st.kernel[k].ms);
                fprintf(f, "}}");
            }
            fprintf(f, "]}\n");
            fclose(f);
        }
    }
}

}  // namespace mthh
