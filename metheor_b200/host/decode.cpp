// decode.cpp — see decode.hpp.
#include "decode.hpp"

#ifdef __AVX2__
#include <immintrin.h>
#endif

#include <algorithm>
#include <cerrno>
#include <cstdlib>

namespace mthh {

static const char* XM_PANIC = "Error reading XM tag in BAM record. Make sure the reads are aligned using Bismark!";

void SoaChunk::clear() {
    tid.clear(); start.clear(); end.clear(); meta.clear(); n_cpg.clear();
    cpg_pos.clear(); cpg_rel.clear(); cpg_meth.clear();
}

// ---- --cpg-set -------------------------------------------------------------------------------------------
void CpgSet::load(const std::string& path, const Header& h) {
    MappedFile f;
    try {
        f.open(path);
    } catch (const HostError&) {
        throw HostError{101, "Could not read target CpG file.: " + path};  // readutil.rs:356 expect(...)
    }
    by_tid.assign(h.names.size(), {});
    const char* d = (const char*)f.data();
    size_t n = f.size(), o = 0;
    std::string last_chrom;
    int last_tid = -1;
    while (o < n) {
        const char* nl = (const char*)memchr(d + o, '\n', n - o);
        size_t e = nl ? (size_t)(nl - d) : n;
        size_t len = e - o;
        if (len && d[o + len - 1] == '\r') len--;  // str::lines() also strips "\r\n"
        const char* line = d + o;
        o = e + 1;
        if (len == 0 && o >= n) break;
        const char* t1 = (const char*)memchr(line, '\t', len);
        // tokens[1] on a line without a tab is an index panic in the reference (readutil.rs:362)
        if (!t1) throw HostError{101, "malformed line in target CpG file (expected chrom<TAB>pos): " + path};
        std::string chrom(line, (size_t)(t1 - line));
        const char* p0 = t1 + 1;
        const char* t2 = (const char*)memchr(p0, '\t', len - (size_t)(p0 - line));
        std::string ps(p0, t2 ? (size_t)(t2 - p0) : len - (size_t)(p0 - line));
        if (chrom != last_chrom) {
            last_tid = h.tid_of(chrom);
            last_chrom = chrom;
        }
        // bamutil.rs:24 header.tid(chrom).unwrap(): unknown contig panics
        if (last_tid < 0) throw HostError{101, "unknown chromosome '" + chrom + "' in target CpG file: " + path};
        char* endp = nullptr;
        errno = 0;
        long long v = strtoll(ps.c_str(), &endp, 10);
        if (ps.empty() || *endp || errno || v < INT32_MIN || v > INT32_MAX)  // parse::<i32>().unwrap()
            throw HostError{101, "invalid position '" + ps + "' in target CpG file: " + path};
        by_tid[(size_t)last_tid].push_back((int32_t)v);
    }
    for (auto& v : by_tid) {
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
    }
}

bool CpgSet::contains(int32_t tid, int32_t pos) const {
    if (tid < 0 || (size_t)tid >= by_tid.size()) return false;
    const auto& v = by_tid[(size_t)tid];
    return std::binary_search(v.begin(), v.end(), pos);
}

// ---- shared tail: CIGAR walk zipped with the XM string ---------------------------------------------------------
namespace {

struct CigarOp { uint32_t len; uint8_t op; };  // op index into "MIDNSHP=X"

struct ReadFields {
    int32_t tid, pos;
    uint32_t mapq, flag;
    const char* xm;   // nullptr: no usable XM:Z tag
    size_t xm_len;
};

template <class NextOp>
inline void emit_read(const ReadFields& r, NextOp&& next_op, const DecodeOptions& opt, SoaChunk* out, DecodeCounters* cnt) {
    cnt->n_records++;
    const bool mapq_ok = r.mapq >= opt.min_qual;
    if (opt.lpmd_order && !mapq_ok) {  // lpmd.rs:176-181: counted, then skipped before BismarkRead::new looks at XM
        cnt->n_dropped++;
        return;
    }
    if (!r.xm) throw HostError{101, XM_PANIC};
    // readutil.rs:332: forward iff flags is exactly 0, 99 or 147; every other value shifts by -1 (readutil.rs:338)
    const bool fwd = r.flag == 0 || r.flag == 99 || r.flag == 147;
    int64_t ref = r.pos;
    size_t qi = 0;
    int32_t start = -1, end = -1;
    bool plain = true;  // only M/=/X (and H/P): the query index of a call is implied by its position
    const size_t i0 = out->cpg_pos.size();
    CigarOp c;
    while (next_op(&c)) {
        if (c.len == 0) continue;
        switch (c.op) {
            case 0: case 7: case 8: {  // M = X: one reference position per query base
                if (start == -1) start = (int32_t)ref;
                end = (int32_t)(ref + c.len - 1);
                if (qi < r.xm_len) {
                    size_t stop = std::min(r.xm_len, qi + (size_t)c.len);  // zip() stops at the shorter side
                    auto call = [&](size_t x) {
                        const char ch = r.xm[x];
                        int32_t p = (int32_t)(ref + (int64_t)(x - qi)) - (fwd ? 0 : 1);
                        if (opt.cpg_set && !opt.cpg_set->contains(r.tid, p)) return;  // readutil.rs:87-95
                        if (x > 65535) throw HostError{101, "read longer than 65535 query bases is not supported"};
                        out->cpg_pos.push_back(p);
                        out->cpg_rel.push_back((uint16_t)x);
                        out->cpg_meth.push_back(ch == 'Z');
                    };
                    size_t x = qi;
#ifdef __AVX2__
                    // readutil.rs:327-329 keeps 'z' / 'Z': 32 XM characters per compare, then only the hits are visited
                    const __m256i fold = _mm256_set1_epi8(0x20), zed = _mm256_set1_epi8('z');
                    for (; x + 32 <= stop; x += 32) {
                        const __m256i v = _mm256_loadu_si256((const __m256i*)(r.xm + x));
                        uint32_t hits = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_or_si256(v, fold), zed));
                        while (hits) {
                            call(x + (size_t)__builtin_ctz(hits));
                            hits &= hits - 1;
                        }
                    }
#endif
                    for (; x < stop; x++)
                        if ((r.xm[x] | 0x20) == 'z') call(x);
                }
                ref += c.len;
                qi += c.len;
                break;
            }
            case 1: case 4: qi += c.len; plain = false; break;   // I S: query only (None in reference_positions_full)
            case 2: case 3: ref += c.len; plain = false; break;  // D N: reference only
            default: break;                        // H P
        }
    }
    const size_t n = out->cpg_pos.size() - i0;
    if (n == 0 && !opt.keep_empty) {
        cnt->n_dropped++;
        if (mapq_ok) cnt->n_dropped_mapq_ok++;
        return;
    }
    if ((int32_t)n > cnt->max_cpgs) cnt->max_cpgs = (int32_t)n;
    if ((int64_t)end - start + 1 > cnt->max_span) cnt->max_span = (int64_t)end - start + 1;
    out->tid.push_back(r.tid);
    out->start.push_back(start);
    out->end.push_back(end);
    out->meta.push_back(r.mapq | ((uint32_t)fwd << 8) | (plain ? 0u : SOA_META_COMPLEX));
    out->n_cpg.push_back((uint32_t)n);
}

void decode_bam(const RecordRef& rec, const DecodeOptions& opt, SoaChunk* out, DecodeCounters* cnt) {
    const uint8_t* p = rec.p;
    const size_t bs = rec.len;
    ReadFields r;
    r.tid = le32(p);
    r.pos = le32(p + 4);
    const uint32_t l_read_name = p[8];
    r.mapq = p[9];
    const uint32_t n_cigar = le16(p + 12);
    r.flag = le16(p + 14);
    const int32_t l_seq = le32(p + 16);
    size_t q = 32 + l_read_name;
    const uint8_t* cig = p + q;
    q += 4 * (size_t)n_cigar + ((size_t)l_seq + 1) / 2 + (size_t)l_seq;
    if (l_seq < 0 || q > bs) throw HostError{101, "Error opening BAM file. corrupt BAM record"};
    r.xm = nullptr;
    r.xm_len = 0;
    while (q + 3 <= bs) {  // aux fields: tag[2], type, value
        const uint8_t t0 = p[q], t1 = p[q + 1], ty = p[q + 2];
        q += 3;
        size_t len;
        switch (ty) {
            case 'A': case 'c': case 'C': len = 1; break;
            case 's': case 'S': len = 2; break;
            case 'i': case 'I': case 'f': len = 4; break;
            case 'd': len = 8; break;  // double: htslib accepts it
            case 'Z': case 'H': {
                const void* z = memchr(p + q, 0, bs - q);
                if (!z) throw HostError{101, "Error opening BAM file. corrupt BAM record (unterminated string in an auxiliary field)"};
                len = (size_t)((const uint8_t*)z - (p + q)) + 1;
                break;
            }
            case 'B': {
                if (q + 5 > bs) throw HostError{101, "Error opening BAM file. corrupt BAM record (truncated array in an auxiliary field)"};
                const uint8_t st = p[q];
                const size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
                len = 5 + es * (size_t)(uint32_t)le32(p + q + 1);
                break;
            }
            default: throw HostError{101, "Error opening BAM file. unknown auxiliary field type in BAM record"};
        }
        if (len > bs - q) throw HostError{101, "Error opening BAM file. corrupt BAM record (auxiliary field runs past the record)"};
        if (t0 == 'X' && t1 == 'M') {
            if (ty == 'Z') { r.xm = (const char*)p + q; r.xm_len = len ? len - 1 : 0; }
            break;  // a non-string XM panics like a missing one (readutil.rs:45-47)
        }
        q += len;
    }
    uint32_t k = 0;
    emit_read(r, [&](CigarOp* c) {
        if (k >= n_cigar) return false;
        uint32_t v = (uint32_t)le32(cig + 4 * (size_t)k++);
        c->len = v >> 4;
        c->op = (uint8_t)(v & 15);
        return true;
    }, opt, out, cnt);
}

struct SamNameCache { std::string name; int32_t tid = -2; };

void decode_sam(const Header& h, const RecordRef& rec, const DecodeOptions& opt, SoaChunk* out, DecodeCounters* cnt,
                SamNameCache* cache) {
    const char* s = (const char*)rec.p;
    const char* e = s + rec.len;
    const char* f[12];
    size_t fl[12];
    int nf = 0;
    const char* cur = s;
    while (nf < 11 && cur <= e) {
        const char* t = (const char*)memchr(cur, '\t', (size_t)(e - cur));
        const char* fe = t ? t : e;
        f[nf] = cur; fl[nf] = (size_t)(fe - cur); nf++;
        cur = fe + 1;
        if (!t) break;
    }
    if (nf < 11) throw HostError{101, "Error opening BAM file. SAM line with fewer than 11 fields"};
    ReadFields r;
    r.flag = (uint32_t)strtoul(std::string(f[1], fl[1]).c_str(), nullptr, 10);
    if (fl[2] == 1 && f[2][0] == '*') {
        r.tid = -1;
    } else {
        if (cache->tid == -2 || cache->name.size() != fl[2] || memcmp(cache->name.data(), f[2], fl[2]) != 0) {
            cache->name.assign(f[2], fl[2]);
            cache->tid = h.tid_of(cache->name);
        }
        r.tid = cache->tid;
    }
    r.pos = (int32_t)strtol(std::string(f[3], fl[3]).c_str(), nullptr, 10) - 1;
    r.mapq = (uint32_t)strtoul(std::string(f[4], fl[4]).c_str(), nullptr, 10) & 0xFFu;
    r.xm = nullptr;
    r.xm_len = 0;
    while (cur < e) {  // optional fields TAG:TYPE:VALUE
        const char* t = (const char*)memchr(cur, '\t', (size_t)(e - cur));
        const char* fe = t ? t : e;
        if (fe - cur >= 5 && cur[0] == 'X' && cur[1] == 'M' && cur[2] == ':') {
            if (cur[3] == 'Z') { r.xm = cur + 5; r.xm_len = (size_t)(fe - cur - 5); }
            break;
        }
        cur = fe + 1;
    }
    const char* c = f[5];
    const char* ce = f[5] + fl[5];
    if (fl[5] == 1 && *c == '*') c = ce;
    emit_read(r, [&](CigarOp* op) {
        if (c >= ce) return false;
        uint32_t len = 0;
        while (c < ce && *c >= '0' && *c <= '9') len = len * 10 + (uint32_t)(*c++ - '0');
        if (c >= ce) return false;
        static const char OPS[] = "MIDNSHP=X";
        const char* w = (const char*)memchr(OPS, *c++, 9);
        if (!w) throw HostError{101, "Error opening BAM file. invalid CIGAR operation in SAM line"};
        op->len = len;
        op->op = (uint8_t)(w - OPS);
        return true;
    }, opt, out, cnt);
}

}  // namespace

void decode_records(Format fmt, const Header& h, const RecordRef* recs, size_t begin, size_t end, const DecodeOptions& opt,
                    SoaChunk* out, DecodeCounters* cnt) {
    if (fmt == Format::BAM) {
        for (size_t i = begin; i < end; i++) decode_bam(recs[i], opt, out, cnt);
    } else {
        SamNameCache cache;
        for (size_t i = begin; i < end; i++) decode_sam(h, recs[i], opt, out, cnt, &cache);
    }
}

int32_t record_tid(Format fmt, const Header& h, const RecordRef& r) {
    if (fmt == Format::BAM) return le32(r.p);
    const char* s = (const char*)r.p;
    const char* e = s + r.len;
    const char* t1 = (const char*)memchr(s, '\t', (size_t)(e - s));
    if (!t1) return -1;
    const char* t2 = (const char*)memchr(t1 + 1, '\t', (size_t)(e - t1 - 1));
    if (!t2) return -1;
    const char* t3 = (const char*)memchr(t2 + 1, '\t', (size_t)(e - t2 - 1));
    if (!t3) return -1;
    std::string name(t2 + 1, (size_t)(t3 - t2 - 1));
    return name == "*" ? -1 : h.tid_of(name);
}

}  // namespace mthh
