// tag.cpp — `metheor tag -i in.bam -o out.sam -g genome.fa`: the host side of XM synthesis.
//
// Mirrors tag::run of the reference (src/tag.rs:386-443) step by step — open the input, check the output directory, open
// the writer, load every contig of the header from the FASTA, then per record: compute the XM string, append it as the
// last aux field, write SAM text (the reference's writer is always bam::Format::Sam, tag.rs:406) — with the per-record
// work (determine_xm_tag_string, tag.rs:130-384) done by the engine: genome resident in HBM, one batch of records per
// window through mth_tag (include/metheor_b200.h), SAM lines formatted on all cores.  No CPU fallback.
#include <sys/stat.h>

#include <cerrno>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/metheor_b200.h"
#include "../../include/metheor_host.h"
#include "input.hpp"
#include "util.hpp"

namespace mthh {

namespace {

struct GenomeHandle {
    mth_genome* g = nullptr;
    ~GenomeHandle() { if (g) mth_genome_destroy(g); }
};

[[noreturn]] void panic(const std::string& msg) { throw HostError{101, msg}; }

// std::path::Path::parent of the output path (tag.rs:396-397): "out.sam" -> "", "d/out.sam" -> "d", "/out.sam" -> "/"
std::string parent_dir(const std::string& path) {
    std::string p = path;
    while (p.size() > 1 && p.back() == '/') p.pop_back();
    const size_t k = p.rfind('/');
    if (k == std::string::npos) return "";
    if (k == 0) return "/";
    std::string d = p.substr(0, k);
    while (d.size() > 1 && d.back() == '/') d.pop_back();
    return d;
}

bool is_dir(const std::string& p) {
    struct stat st;
    return !p.empty() && stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

// Loads every contig of the header from the FASTA into the engine (tag.rs:419-428).  One pass over the mapped file; the
// sequence of a wanted contig is gathered without line ends / blanks (faidx counts isgraph() bytes only) and uploaded.
void load_genome(const std::string& path, const Header& h, mth_genome* g) {
    struct stat st;
    if (stat(path.c_str(), &st) != 0) panic("Error opening reference genome file: file not found: " + path);
    MappedFile f;
    try {
        f.open(path);
    } catch (const HostError& e) {
        panic("Error opening reference genome file: " + e.msg);
    }
    const uint8_t* d = f.data();
    const size_t n = f.size();
    std::vector<char> have(h.names.size(), 0);
    std::vector<uint8_t> seq;
    size_t o = 0;
    while (o < n) {
        if (d[o] != '>') {  // stray text before the first header line
            const void* nl = memchr(d + o, '\n', n - o);
            o = nl ? (size_t)((const uint8_t*)nl - d) + 1 : n;
            continue;
        }
        size_t e = o + 1;
        while (e < n && d[e] != '\n' && d[e] != ' ' && d[e] != '\t' && d[e] != '\r') e++;
        const std::string name((const char*)d + o + 1, e - o - 1);
        const void* nl = memchr(d + o, '\n', n - o);
        o = nl ? (size_t)((const uint8_t*)nl - d) + 1 : n;
        const int tid = h.tid_of(name);
        const bool want = tid >= 0 && !have[(size_t)tid];
        seq.clear();
        if (want) seq.reserve((size_t)h.lengths[(size_t)tid] + 64);
        while (o < n && d[o] != '>') {
            const void* q = memchr(d + o, '\n', n - o);
            const size_t le = q ? (size_t)((const uint8_t*)q - d) : n;
            if (want) {
                size_t a = o, b = le;
                while (b > a && (d[b - 1] == '\r' || d[b - 1] == ' ' || d[b - 1] == '\t')) b--;
                bool clean = true;
                for (size_t k = a; k < b; k++)
                    if (d[k] <= ' ' || d[k] > '~') { clean = false; break; }
                if (clean) seq.insert(seq.end(), d + a, d + b);
                else
                    for (size_t k = a; k < b; k++)
                        if (d[k] > ' ' && d[k] <= '~') seq.push_back(d[k]);
            }
            o = le + 1;
        }
        if (want) {
            if (mth_genome_set_contig(g, tid, seq.data(), (int64_t)seq.size()) != MTH_OK)
                panic(std::string("metheor_b200 engine: ") + mth_genome_last_error(g));
            have[(size_t)tid] = 1;
        }
    }
    for (size_t t = 0; t < have.size(); t++)
        if (!have[t]) panic("Error fetching reference genome sequence.");  // tag.rs:425 (.expect)
}

struct TagBatchHost {
    std::vector<int32_t> tid, pos, l_seq;
    std::vector<uint8_t> rc, seq4;
    std::vector<uint32_t> cigar_off, cigar;
    std::vector<uint64_t> seq_off;
    void clear() {
        tid.clear(); pos.clear(); l_seq.clear(); rc.clear(); seq4.clear(); cigar.clear();
        cigar_off.assign(1, 0);
        seq_off.assign(1, 0);
    }
    void finish_read() {
        cigar_off.push_back((uint32_t)cigar.size());
        seq_off.push_back((uint64_t)seq4.size());
    }
};

// tag.rs:15-18 for pairs, tag.rs:141-144 otherwise
inline bool reverse_complement(uint32_t flag, bool paired) {
    const bool rev = flag & 16, first = flag & 64, last = flag & 128;
    if (paired) return !((!rev && first) || (rev && last));
    return rev;
}

struct Nt16 {
    uint8_t code[256];
    Nt16() {  // htslib seq_nt16_table
        memset(code, 15, sizeof(code));
        const char* s = "=ACMGRSVTWYHKDBN";
        for (int i = 0; i < 16; i++) {
            code[(uint8_t)s[i]] = (uint8_t)i;
            if (s[i] >= 'A' && s[i] <= 'Z') code[(uint8_t)(s[i] + 32)] = (uint8_t)i;
        }
        code['0'] = 1; code['1'] = 2; code['2'] = 4; code['3'] = 8;
    }
};
const Nt16 NT16;

struct Fields {
    const char* b[12];
    const char* e[12];
    int n;
};
inline Fields split11(const RecordRef& r) {
    Fields f;
    f.n = 0;
    const char* p = (const char*)r.p;
    const char* end = p + r.len;
    while (f.n < 11) {
        const char* t = (const char*)memchr(p, '\t', (size_t)(end - p));
        f.b[f.n] = p;
        f.e[f.n] = t ? t : end;
        f.n++;
        if (!t) break;
        p = t + 1;
    }
    f.b[11] = f.n == 11 && f.e[10] < end ? f.e[10] + 1 : end;  // aux part
    f.e[11] = end;
    return f;
}

void pack_sam(const RecordRef& r, const Header& h, bool paired, TagBatchHost* b, std::string* last_name, int* last_tid) {
    const Fields f = split11(r);
    if (f.n < 11) panic("Error opening BAM file. truncated SAM line (fewer than 11 fields)");
    const uint32_t flag = (uint32_t)strtoul(std::string(f.b[1], f.e[1]).c_str(), nullptr, 10);
    int tid = -1;
    if (!(f.e[2] - f.b[2] == 1 && *f.b[2] == '*')) {
        const std::string name(f.b[2], f.e[2]);
        if (name == *last_name) tid = *last_tid;
        else { tid = h.tid_of(name); *last_name = name; *last_tid = tid; }
    }
    const int64_t pos1 = strtoll(std::string(f.b[3], f.e[3]).c_str(), nullptr, 10);
    b->tid.push_back(tid);
    b->pos.push_back((int32_t)(pos1 - 1));
    b->rc.push_back(reverse_complement(flag, paired));
    if (!(f.e[5] - f.b[5] == 1 && *f.b[5] == '*')) {
        const char* c = f.b[5];
        while (c < f.e[5]) {
            uint32_t len = 0;
            while (c < f.e[5] && *c >= '0' && *c <= '9') len = len * 10 + (uint32_t)(*c++ - '0');
            if (c >= f.e[5]) break;
            static const char OPS[] = "MIDNSHP=X";
            const char* q = strchr(OPS, *c++);
            if (!q || !*q) panic("Error opening BAM file. unrecognised CIGAR operator in SAM line");
            b->cigar.push_back((len << 4) | (uint32_t)(q - OPS));
        }
    }
    int32_t l_seq = 0;
    if (!(f.e[9] - f.b[9] == 1 && *f.b[9] == '*')) {
        l_seq = (int32_t)(f.e[9] - f.b[9]);
        const uint8_t* s = (const uint8_t*)f.b[9];
        for (int32_t i = 0; i + 1 < l_seq; i += 2) b->seq4.push_back((uint8_t)((NT16.code[s[i]] << 4) | NT16.code[s[i + 1]]));
        if (l_seq & 1) b->seq4.push_back((uint8_t)(NT16.code[s[l_seq - 1]] << 4));
    }
    b->l_seq.push_back(l_seq);
    b->finish_read();
}

void pack_bam(const RecordRef& r, bool paired, TagBatchHost* b) {
    const uint8_t* p = r.p;
    if (r.len < 32) panic("Error opening BAM file. corrupt BAM record");
    const uint32_t l_name = p[8], n_cigar = le16(p + 12), flag = le16(p + 14);
    const int32_t l_seq = le32(p + 16);
    const size_t o_cigar = 32 + (size_t)l_name, o_seq = o_cigar + 4 * (size_t)n_cigar;
    if (l_seq < 0 || o_seq + ((size_t)l_seq + 1) / 2 + (size_t)l_seq > r.len) panic("Error opening BAM file. corrupt BAM record");
    b->tid.push_back(le32(p));
    b->pos.push_back(le32(p + 4));
    b->rc.push_back(reverse_complement(flag, paired));
    b->l_seq.push_back(l_seq);
    const size_t c0 = b->cigar.size();
    b->cigar.resize(c0 + n_cigar);
    if (n_cigar) memcpy(b->cigar.data() + c0, p + o_cigar, 4 * (size_t)n_cigar);
    b->seq4.insert(b->seq4.end(), p + o_seq, p + o_seq + ((size_t)l_seq + 1) / 2);
    b->finish_read();
}

// ---- SAM text of a BAM record (what htslib's sam_format1 prints) ----
inline void put_int(std::string* s, long long v) {
    char buf[24];
    char* e = buf + sizeof(buf);
    char* p = e;
    unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
    do {
        *--p = (char)('0' + u % 10);
        u /= 10;
    } while (u);
    if (v < 0) *--p = '-';
    s->append(p, (size_t)(e - p));
}
inline void put_g(std::string* s, double v) {
    char buf[40];
    int n = snprintf(buf, sizeof(buf), "%g", v);
    s->append(buf, (size_t)n);
}

// Returns false when the record already carries an XM tag.
bool format_bam_record(const RecordRef& r, const Header& h, std::string* s) {
    const uint8_t* p = r.p;
    const int32_t tid = le32(p), pos = le32(p + 4);
    const uint32_t l_name = p[8], mapq = p[9], n_cigar = le16(p + 12), flag = le16(p + 14);
    const int32_t l_seq = le32(p + 16), ntid = le32(p + 20), npos = le32(p + 24), tlen = le32(p + 28);
    const size_t o_cigar = 32 + (size_t)l_name, o_seq = o_cigar + 4 * (size_t)n_cigar, o_qual = o_seq + ((size_t)l_seq + 1) / 2;
    size_t q = o_qual + (size_t)l_seq;
    if (l_name > 1) s->append((const char*)p + 32, strnlen((const char*)p + 32, l_name - 1));
    else s->push_back('*');
    s->push_back('\t'); put_int(s, flag);
    s->push_back('\t');
    if (tid >= 0 && (size_t)tid < h.names.size()) s->append(h.names[(size_t)tid]); else s->push_back('*');
    s->push_back('\t'); put_int(s, (long long)pos + 1);
    s->push_back('\t'); put_int(s, mapq);
    s->push_back('\t');
    if (n_cigar) {
        for (uint32_t k = 0; k < n_cigar; k++) {
            const uint32_t v = (uint32_t)le32(p + o_cigar + 4 * (size_t)k);
            put_int(s, v >> 4);
            s->push_back("MIDNSHP=XB??????"[v & 15]);
        }
    } else {
        s->push_back('*');
    }
    s->push_back('\t');
    if (ntid < 0) s->push_back('*');
    else if (ntid == tid) s->push_back('=');
    else if ((size_t)ntid < h.names.size()) s->append(h.names[(size_t)ntid]);
    else s->push_back('*');
    s->push_back('\t'); put_int(s, (long long)npos + 1);
    s->push_back('\t'); put_int(s, tlen);
    s->push_back('\t');
    if (l_seq) {
        const size_t at = s->size();
        s->resize(at + (size_t)l_seq);
        char* dst = &(*s)[at];
        for (int32_t i = 0; i < l_seq; i++) dst[i] = "=ACMGRSVTWYHKDBN"[(p[o_seq + ((size_t)i >> 1)] >> ((~i & 1) << 2)) & 15];
    } else {
        s->push_back('*');
    }
    s->push_back('\t');
    if (l_seq && p[o_qual] != 0xff) {
        const size_t at = s->size();
        s->resize(at + (size_t)l_seq);
        char* dst = &(*s)[at];
        for (int32_t i = 0; i < l_seq; i++) dst[i] = (char)(p[o_qual + (size_t)i] + 33);
    } else {
        s->push_back('*');
    }
    bool has_xm = false;
    while (q + 3 <= r.len) {
        const uint8_t t0 = p[q], t1 = p[q + 1], ty = p[q + 2];
        q += 3;
        if (t0 == 'X' && t1 == 'M') has_xm = true;
        s->push_back('\t'); s->push_back((char)t0); s->push_back((char)t1); s->push_back(':');
        auto need = [&](size_t nbytes) { if (q + nbytes > r.len) panic("Error opening BAM file. corrupt auxiliary field in BAM record"); };
        switch (ty) {
            case 'A': need(1); s->append("A:"); s->push_back((char)p[q]); q += 1; break;
            case 'c': need(1); s->append("i:"); put_int(s, (int8_t)p[q]); q += 1; break;
            case 'C': need(1); s->append("i:"); put_int(s, p[q]); q += 1; break;
            case 's': need(2); s->append("i:"); put_int(s, (int16_t)le16(p + q)); q += 2; break;
            case 'S': need(2); s->append("i:"); put_int(s, le16(p + q)); q += 2; break;
            case 'i': need(4); s->append("i:"); put_int(s, le32(p + q)); q += 4; break;
            case 'I': need(4); s->append("i:"); put_int(s, (uint32_t)le32(p + q)); q += 4; break;
            case 'f': { need(4); float v; memcpy(&v, p + q, 4); s->append("f:"); put_g(s, v); q += 4; break; }
            case 'd': { need(8); double v; memcpy(&v, p + q, 8); s->append("d:"); put_g(s, v); q += 8; break; }
            case 'Z': case 'H': {
                const void* z = memchr(p + q, 0, r.len - q);
                if (!z) panic("Error opening BAM file. unterminated string in auxiliary field");
                const size_t len = (size_t)((const uint8_t*)z - (p + q));
                s->push_back((char)ty); s->push_back(':');
                s->append((const char*)p + q, len);
                q += len + 1;
                break;
            }
            case 'B': {
                need(5);
                const uint8_t st = p[q];
                const uint32_t cnt = (uint32_t)le32(p + q + 1);
                q += 5;
                const size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
                need(es * (size_t)cnt);
                s->append("B:"); s->push_back((char)st);
                for (uint32_t k = 0; k < cnt; k++, q += es) {
                    s->push_back(',');
                    switch (st) {
                        case 'c': put_int(s, (int8_t)p[q]); break;
                        case 'C': put_int(s, p[q]); break;
                        case 's': put_int(s, (int16_t)le16(p + q)); break;
                        case 'S': put_int(s, le16(p + q)); break;
                        case 'i': put_int(s, le32(p + q)); break;
                        case 'I': put_int(s, (uint32_t)le32(p + q)); break;
                        case 'f': { float v; memcpy(&v, p + q, 4); put_g(s, v); break; }
                        default: panic("Error opening BAM file. unknown array type in auxiliary field");
                    }
                }
                break;
            }
            default: panic("Error opening BAM file. unknown auxiliary field type in BAM record");
        }
    }
    return !has_xm;
}

bool sam_line_has_xm(const RecordRef& r) {
    const Fields f = split11(r);
    const char* p = f.b[11];
    while (p < f.e[11]) {
        if (f.e[11] - p >= 3 && p[0] == 'X' && p[1] == 'M' && p[2] == ':') return true;
        const char* t = (const char*)memchr(p, '\t', (size_t)(f.e[11] - p));
        if (!t) break;
        p = t + 1;
    }
    return false;
}

const char* status_text(uint8_t st) {
    switch (st) {
        case MTH_TAG_NO_COMPLEMENT: return "no reverse complement for a base of the read or the reference (tag.rs:23)";
        case MTH_TAG_NO_CONTEXT: return "cytosine context runs past the end of the alignment (tag.rs:301)";
        case MTH_TAG_BAD_CONTIG: return "read is unmapped or names a contig outside the header (tag.rs:152)";
        case MTH_TAG_PAST_END: return "alignment ends past the contig or the genome sequence (tag.rs:158-172)";
        case MTH_TAG_SHORT_SEQ: return "SEQ is shorter than the CIGAR's aligned bases";
        default: return "unknown";
    }
}

}  // namespace

void run_tag(const char* input, const char* output, const char* genome, int device, int threads, const char* stats_json) {
    const double t_begin = now_s();
    int n_threads = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (n_threads < 1) n_threads = 1;
    RecordStream rs(input, n_threads, (size_t)64 << 20);  // tag.rs:387 (get_reader panics when the file is missing)
    const Header& h = rs.header();
    const std::string dir = parent_dir(output);
    if (!is_dir(dir)) panic("No such directory for output alignment file: " + dir);  // tag.rs:396-404
    std::vector<char> obuf((size_t)1 << 22);
    FILE* out = fopen(output, "wb");
    if (!out) panic(std::string("Error opening alignment file to write: ") + strerror(errno));  // tag.rs:407-410
    struct Closer { FILE* f; ~Closer() { if (f) fclose(f); } } closer{out};
    setvbuf(out, obuf.data(), _IOFBF, obuf.size());

    {  // tag.rs:412-417: faidx::Reader::from_path fails before any sequence is fetched
        struct stat st;
        if (stat(genome, &st) != 0) panic(std::string("Error opening reference genome file: file not found: ") + genome);
    }
    GenomeHandle G;
    if (mth_genome_create(&G.g, device, (int32_t)h.lengths.size(), h.lengths.data()) != MTH_OK)
        panic(std::string("metheor_b200 engine: ") + mth_genome_last_error(nullptr));
    printf("Parsing reference genome...\n");  // tag.rs:418
    fflush(stdout);
    const double t_g0 = now_s();
    load_genome(genome, h, G.g);
    const double t_g1 = now_s();
    printf("Done!\n");  // tag.rs:429
    fflush(stdout);

    // header: Header::from_template(view) re-serialises the text, bam::Writer prints it followed by the records
    std::string text = h.text;
    while (!text.empty() && (text.back() == '\0' || text.back() == '\n')) text.pop_back();
    if (!text.empty()) {
        text.push_back('\n');
        fwrite(text.data(), 1, text.size(), out);
    }

    ThreadPool pool(n_threads);
    std::vector<RecordRef> recs;
    TagBatchHost b;
    bool first = true, paired = false;
    int64_t n_records = 0, n_batches = 0;
    double s_pack = 0, s_gpu = 0, s_format = 0, s_write = 0;
    std::future<bool> writing;  // declared after `closer`: a pending write is joined before the file is closed
    std::string last_name;
    int last_tid = -1;
    const bool bam = rs.format() == Format::BAM;
    while (rs.next(&recs)) {
        if (recs.empty()) continue;
        double t0 = now_s();
        if (first) {  // bamutil.rs:27-38: paired iff the FIRST record has flag bit 1
            uint32_t flag;
            if (bam) flag = le16(recs[0].p + 14);
            else {
                const Fields f = split11(recs[0]);
                flag = f.n >= 2 ? (uint32_t)strtoul(std::string(f.b[1], f.e[1]).c_str(), nullptr, 10) : 0;
            }
            paired = flag & 1;
            first = false;
        }
        b.clear();
        for (const RecordRef& r : recs) {
            if (bam) pack_bam(r, paired, &b);
            else pack_sam(r, h, paired, &b, &last_name, &last_tid);
        }
        double t1 = now_s();
        mth_tag_batch tb;
        tb.n_reads = (int64_t)recs.size();
        tb.tid = b.tid.data(); tb.pos = b.pos.data(); tb.rc = b.rc.data(); tb.l_seq = b.l_seq.data();
        tb.cigar_off = b.cigar_off.data(); tb.cigar = b.cigar.data(); tb.seq_off = b.seq_off.data(); tb.seq4 = b.seq4.data();
        mth_tag_result res;
        if (mth_tag(G.g, &tb, &res) != MTH_OK) panic(std::string("metheor_b200 engine: ") + mth_genome_last_error(G.g));
        if (res.n_failed) {
            for (int64_t i = 0; i < res.n_reads; i++)
                if (res.status[i] != MTH_TAG_OK)
                    panic(std::string("Error determining XM tag for record ") + std::to_string(n_records + i + 1) + ": " + status_text(res.status[i]));
        }
        double t2 = now_s();
        // format in parallel, in slices; write in order
        const int64_t n = (int64_t)recs.size();
        const int64_t n_slices = std::min<int64_t>((int64_t)n_threads * 4, std::max<int64_t>(1, n / 256));
        std::vector<std::string> parts((size_t)n_slices);
        std::atomic<int64_t> dup{-1};
        pool.run(n_slices, [&](int64_t sidx, int) {
            const int64_t lo = n * sidx / n_slices, hi = n * (sidx + 1) / n_slices;
            std::string& s = parts[(size_t)sidx];
            s.reserve((size_t)(hi - lo) * 512);
            for (int64_t i = lo; i < hi; i++) {
                bool ok;
                if (bam) ok = format_bam_record(recs[(size_t)i], h, &s);
                else {
                    ok = !sam_line_has_xm(recs[(size_t)i]);
                    s.append((const char*)recs[(size_t)i].p, recs[(size_t)i].len);
                }
                if (!ok) {
                    int64_t want = -1;
                    dup.compare_exchange_strong(want, i);
                }
                s.append("\tXM:Z:");
                s.append((const char*)res.xm + res.xm_off[i], res.xm_len[i]);
                s.push_back('\n');
            }
        });
        if (dup.load() >= 0)  // rust-htslib's push_aux refuses a tag that is already there (tag.rs:417-421)
            panic("Error adding XM tag to alignment record. the record already carries an XM tag");
        double t3 = now_s();
        // the text of this window is written on a second thread while the next window is inflated, packed and tagged
        if (writing.valid() && !writing.get()) panic("Error writing to output file.");  // tag.rs:423
        auto owned = std::make_shared<std::vector<std::string>>(std::move(parts));
        writing = std::async(std::launch::async, [owned, out, &s_write] {
            const double w0 = now_s();
            bool ok = true;
            for (const std::string& s : *owned)
                if (!s.empty() && fwrite(s.data(), 1, s.size(), out) != s.size()) ok = false;
            s_write += now_s() - w0;
            return ok;
        });
        s_pack += t1 - t0; s_gpu += t2 - t1; s_format += t3 - t2;
        n_records += n;
        n_batches++;
    }
    if (writing.valid() && !writing.get()) panic("Error writing to output file.");
    if (fflush(out) != 0) panic("Error writing to output file.");
    const double t_end = now_s();
    if (stats_json) {
        FILE* f = fopen(stats_json, "w");
        if (f) {
            fprintf(f, "{\"records\": %lld, \"batches\": %lld, \"threads\": %d, \"seconds\": {\"total\": %.6f, \"genome\": %.6f, \"inflate\": %.6f, "
                       "\"walk\": %.6f, \"pack\": %.6f, \"gpu\": %.6f, \"format\": %.6f, \"write\": %.6f}, \"reads_per_sec\": %.1f}\n",
                    (long long)n_records, (long long)n_batches, n_threads, t_end - t_begin, t_g1 - t_g0, rs.seconds_inflate, rs.seconds_walk,
                    s_pack, s_gpu, s_format, s_write, (double)n_records / (t_end - t_begin));
            fclose(f);
        }
    }
}

}  // namespace mthh
