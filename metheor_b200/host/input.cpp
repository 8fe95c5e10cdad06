// input.cpp — see input.hpp.  BGZF/BAM layout per the SAM specification.
#include "input.hpp"

#include <errno.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cstdlib>

#include "inflate_fast.hpp"

namespace mthh {

std::atomic<int64_t> g_zlib_fallbacks{0};

int Header::tid_of(const std::string& n) const {
    for (size_t i = 0; i < names.size(); i++)
        if (names[i] == n) return (int)i;
    return -1;
}

static HostError open_error(const std::string& what) { return HostError{101, "Error opening BAM file. " + what}; }

MappedFile::~MappedFile() {
    if (data_ && size_) munmap((void*)data_, size_);
    if (fd_ >= 0) close(fd_);
}

void MappedFile::open(const std::string& path) {
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) {
        // rust-htslib: Error::FileNotFound { path } => "file not found: {path}" (tests/pdr-cli.rs:19-33 checks both parts)
        if (errno == ENOENT) throw open_error("file not found: " + path);
        throw open_error("unable to open " + path + ": " + strerror(errno));
    }
    struct stat st;
    if (fstat(fd_, &st) != 0 || !S_ISREG(st.st_mode)) throw open_error("not a regular file: " + path);
    size_ = (size_t)st.st_size;
    if (size_ == 0) throw open_error("invalid (empty) file: " + path);
    void* p = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
    if (p == MAP_FAILED) throw open_error("mmap failed for " + path + ": " + strerror(errno));
    madvise(p, size_, MADV_SEQUENTIAL);
    data_ = (const uint8_t*)p;
}

namespace {

struct Inflater {
    z_stream zs;
    bool ready = false;
    ~Inflater() {
        if (ready) inflateEnd(&zs);
    }
    // raw DEFLATE payload of one BGZF member -> exactly usize bytes; CRC32 checked like htslib's bgzf reader does
    bool run(const uint8_t* in, size_t clen, uint8_t* out, uint32_t usize, uint32_t crc_expected) {
        static const bool zlib_only = getenv("METHEOR_ZLIB_INFLATE") != nullptr;  // A/B switch for measurements
        if (!zlib_only) {
            if (inflate_fast(in, clen, out, usize) && crc32_fast(out, usize) == crc_expected) return true;
            g_zlib_fallbacks.fetch_add(1, std::memory_order_relaxed);  // rejected or wrong: let zlib have the last word
        }
        if (!ready) {
            memset(&zs, 0, sizeof(zs));
            if (inflateInit2(&zs, -15) != Z_OK) return false;
            ready = true;
        } else if (inflateReset(&zs) != Z_OK) {
            return false;
        }
        zs.next_in = (Bytef*)in;
        zs.avail_in = (uInt)clen;
        zs.next_out = out;
        zs.avail_out = usize;
        int rc = inflate(&zs, Z_FINISH);
        if (rc != Z_STREAM_END || zs.avail_out != 0) return false;
        return crc32_fast(out, usize) == crc_expected;
    }
};

// Parses one BGZF member header at `o`; returns the member size or 0 if it is not a valid BGZF member.
size_t bgzf_member(const uint8_t* d, size_t size, size_t o, size_t* cdata, size_t* clen, uint32_t* usize, uint32_t* crc) {
    if (o + 18 > size) return 0;
    if (d[o] != 31 || d[o + 1] != 139 || d[o + 2] != 8 || !(d[o + 3] & 4)) return 0;
    uint16_t xlen = le16(d + o + 10);
    size_t x = o + 12, xend = x + xlen;
    if (xend > size) return 0;
    int bsize = -1;
    while (x + 4 <= xend) {
        uint16_t slen = le16(d + x + 2);
        if (d[x] == 'B' && d[x + 1] == 'C' && slen == 2 && x + 6 <= xend) bsize = le16(d + x + 4);
        x += 4 + (size_t)slen;
    }
    if (bsize < 0) return 0;
    size_t total = (size_t)bsize + 1;
    if (total < (size_t)xlen + 20 || o + total > size) return 0;
    *cdata = o + 12 + xlen;
    *clen = total - xlen - 12 - 8;
    *crc = (uint32_t)le32(d + o + total - 8);
    *usize = (uint32_t)le32(d + o + total - 4);
    return total;
}

}  // namespace

size_t bgzf_member_info(const uint8_t* d, size_t size, size_t o, size_t* cdata, size_t* clen, uint32_t* usize, uint32_t* crc) {
    return bgzf_member(d, size, o, cdata, clen, usize, crc);
}

size_t RecordStream::bam_header_bytes() const { return bam_header_len_; }

void RecordStream::seek_virtual(uint64_t voffset) {
    if (format_ != Format::BAM || prefetching_ || inflating_) throw HostError{1, "metheor_b200: seek on a stream that is already being read"};
    coff_ = (size_t)(voffset >> 16);
    skip_first_ = (size_t)(voffset & 0xffffu);
    carry_.clear();
    consumed_ = coff_;
    if (coff_ > file_.size()) throw HostError{101, "Error opening BAM file. index offset beyond the end of the file: " + path_};
}

RecordStream::RecordStream(const std::string& path, int n_threads, size_t window_bytes)
    : path_(path), pool_(n_threads), window_bytes_(window_bytes) {
    file_.open(path);
    const uint8_t* d = file_.data();
    if (file_.size() >= 2 && d[0] == 31 && d[1] == 139) {
        format_ = Format::BAM;
        parse_bam_header();
    } else if (d[0] == '@') {
        format_ = Format::SAM;
        parse_sam_header();
    } else {
        // htslib also accepts header-less SAM; require at least one plausible alignment line (11 tab-separated fields)
        size_t e = 0, tabs = 0;
        while (e < file_.size() && d[e] != '\n' && e < 65536) tabs += d[e++] == '\t';
        if (tabs < 10) throw open_error("unrecognised format (neither BGZF/BAM nor SAM): " + path);
        format_ = Format::SAM;
    }
}

RecordStream::~RecordStream() {
    if (prefetching_) next_.wait();
    if (inflating_) inflate_fut_.wait();
}

// Index the next members (up to window_bytes_ of output), inflate them on the reader's pool.  Touches only coff_,
// which nobody else uses while a prefetch is in flight.
RecordStream::Inflated RecordStream::inflate_next(size_t headroom, size_t max_bytes) {
    if (!max_bytes) max_bytes = window_bytes_;
    Inflated out;
    const uint8_t* d = file_.data();
    std::vector<Block> blocks;
    size_t utotal = 0;
    while (coff_ < file_.size() && utotal < max_bytes) {
        Block b;
        uint32_t crc;
        size_t total = bgzf_member(d, file_.size(), coff_, &b.cdata, &b.clen, &b.usize, &crc);
        if (!total) {
            out.ok = false;
            out.err = "corrupt or truncated BGZF block at offset " + std::to_string(coff_) + ": " + path_;
            return out;
        }
        b.uoff = utotal;
        utotal += b.usize;
        if (b.usize) blocks.push_back(b);
        coff_ += total;
    }
    out.headroom = headroom;
    out.len = utotal;
    out.coff_end = coff_;
    if (blocks.empty()) return out;
    out.data.reset(new uint8_t[headroom + utotal]);  // deliberately uninitialised
    double t0 = now_s();
    std::vector<Inflater> infl((size_t)pool_.size());
    std::atomic<int> bad{0};
    uint8_t* dst = out.data.get() + headroom;
    pool_.run((int64_t)blocks.size(), [&](int64_t i, int w) {
        const Block& b = blocks[(size_t)i];
        uint32_t crc = (uint32_t)le32(d + b.cdata + b.clen);
        if (!infl[(size_t)w].run(d + b.cdata, b.clen, dst + b.uoff, b.usize, crc)) bad.store(1);
    });
    if (bad.load()) {
        out.ok = false;
        out.err = "BGZF inflate / CRC failure: " + path_;
    }
    out.seconds = now_s() - t0;
    return out;
}

RecordStream::Window RecordStream::produce() {
    Window w;
    for (;;) {
        // two-deep: the NEXT run is already inflating on the pool while this one is walked below
        if (!inflating_) {
            inflate_fut_ = std::async(std::launch::async, [this] { return inflate_next(kHeadroom); });
            inflating_ = true;
        }
        Inflated in = inflate_fut_.get();
        inflating_ = false;
        if (in.ok && in.len) {
            inflate_fut_ = std::async(std::launch::async, [this] { return inflate_next(kHeadroom); });
            inflating_ = true;
        }
        if (in.ok && carry_.size() > in.headroom) {  // rare: the carried-over record is larger than the headroom
            std::unique_ptr<uint8_t[]> nb(new uint8_t[carry_.size() + in.len]);
            if (in.len) memcpy(nb.get() + carry_.size(), in.data.get() + in.headroom, in.len);
            in.data = std::move(nb);
            in.headroom = carry_.size();
        }
        w.s_inflate += in.seconds;
        if (!in.ok) { w.ok = false; w.err = in.err; return w; }
        const bool last = in.len == 0;  // no members left: only the carried-over bytes remain
        if (last) {
            if (carry_.empty()) { w.eof = true; return w; }
            in.data.reset(new uint8_t[carry_.size()]);
        }
        w.inflated += in.len;
        w.coff_end = in.coff_end;
        const size_t carry = carry_.size();
        uint8_t* front = in.data.get() + (last ? 0 : in.headroom - carry);
        if (carry) memcpy(front, carry_.data(), carry);
        const uint8_t* b = front;
        const size_t n = carry + in.len;
        double t0 = now_s();
        size_t o = 0;
        if (skip_first_) {  // seek_virtual: the first record starts this far into the first inflated member
            if (skip_first_ > n) { w.ok = false; w.err = "index points past the end of a BGZF block: " + path_; return w; }
            o = skip_first_;
            skip_first_ = 0;
        }
        w.recs.reserve(n / 256);
        while (o + 4 <= n) {
            int32_t bs = le32(b + o);
            if (bs < 32) { w.ok = false; w.err = "corrupt BAM record (block_size < 32): " + path_; return w; }
            if (o + 4 + (size_t)bs > n) break;
            w.recs.push_back(RecordRef{b + o + 4, (uint32_t)bs});
            o += 4 + (size_t)bs;
        }
        w.s_walk += now_s() - t0;
        if (last && o != n) { w.ok = false; w.err = "truncated BAM record at end of file: " + path_; return w; }
        carry_.assign(b + o, b + n);  // the partial record (if any) moves in front of the next run
        if (!w.recs.empty()) {
            w.data = std::move(in.data);
            return w;
        }
        // a single record larger than one run: keep accumulating
    }
}

void RecordStream::parse_bam_header() {
    // the header may span several runs: keep appending until it is complete
    for (;;) {
        Inflated in = inflate_next(carry_.size(), 1u << 18);  // small runs: the header is usually a few KB
        if (!in.ok) throw open_error(in.err);
        seconds_inflate += in.seconds;
        bytes_uncompressed += in.len;
        consumed_ = in.coff_end;
        const size_t carry = carry_.size();
        if (in.len == 0 && carry == 0) throw open_error("not a BAM file (empty BGZF stream): " + path_);
        std::vector<uint8_t> all(carry + in.len);
        if (carry) memcpy(all.data(), carry_.data(), carry);
        if (in.len) memcpy(all.data() + carry, in.data.get() + in.headroom, in.len);
        const uint8_t* b = all.data();
        const size_t n = all.size();
        if (n < 12 || memcmp(b, "BAM\1", 4) != 0) throw open_error("not a BAM file (bad magic): " + path_);
        bool ok = false;
        size_t o = 4;
        do {
            int32_t l_text = le32(b + o);
            if (l_text < 0) throw open_error("corrupt BAM header: " + path_);
            const size_t o_text = o + 4;
            o += 4 + (size_t)l_text;
            if (o + 4 > n) break;
            int32_t n_ref = le32(b + o);
            o += 4;
            Header h;
            h.text.assign((const char*)b + o_text, strnlen((const char*)b + o_text, (size_t)l_text));
            bool complete = true;
            for (int32_t i = 0; i < n_ref; i++) {
                if (o + 4 > n) { complete = false; break; }
                int32_t l_name = le32(b + o);
                o += 4;
                if (l_name < 1 || o + (size_t)l_name + 4 > n) { complete = false; break; }
                h.names.emplace_back((const char*)b + o, (size_t)l_name - 1);
                o += (size_t)l_name;
                h.lengths.push_back((int64_t)(uint32_t)le32(b + o));
                o += 4;
            }
            if (!complete) break;
            header_ = std::move(h);
            ok = true;
        } while (false);
        if (ok) {
            bam_header_len_ = o;
            carry_.assign(b + o, b + n);  // the records behind the header start the first window
            return;
        }
        if (in.len == 0) throw open_error("truncated BAM header: " + path_);
        carry_.swap(all);
    }
}

void RecordStream::parse_sam_header() {
    const uint8_t* d = file_.data();
    size_t n = file_.size(), o = 0;
    while (o < n && d[o] == '@') {
        size_t e = o;
        while (e < n && d[e] != '\n') e++;
        if (e - o >= 3 && d[o + 1] == 'S' && d[o + 2] == 'Q') {
            std::string sn;
            int64_t ln = 0;
            size_t f = o;
            while (f < e) {
                size_t g = f;
                while (g < e && d[g] != '\t') g++;
                if (g - f > 3 && d[f + 2] == ':') {
                    if (d[f] == 'S' && d[f + 1] == 'N') sn.assign((const char*)d + f + 3, g - f - 3);
                    if (d[f] == 'L' && d[f + 1] == 'N') ln = atoll(std::string((const char*)d + f + 3, g - f - 3).c_str());
                }
                f = g + 1;
            }
            if (!sn.empty() && sn.back() == '\r') sn.pop_back();
            header_.names.push_back(sn);
            header_.lengths.push_back(ln);
        }
        o = e + 1;
    }
    if (o > n) o = n;
    header_.text.assign((const char*)d, o);
    coff_ = o;
}

bool RecordStream::next(std::vector<RecordRef>* recs) {
    recs->clear();
    if (format_ == Format::SAM) {
        const uint8_t* d = file_.data();
        size_t n = file_.size();
        if (coff_ >= n) return false;
        double t0 = now_s();
        size_t stop = std::min(n, coff_ + window_bytes_);
        size_t o = coff_;
        while (o < n && (o < stop)) {
            const uint8_t* nl = (const uint8_t*)memchr(d + o, '\n', n - o);
            size_t e = nl ? (size_t)(nl - d) : n;
            size_t len = e - o;
            if (len && d[e - 1] == '\r') len--;
            if (len && d[o] != '@') recs->push_back(RecordRef{d + o, (uint32_t)len});
            o = e + 1;
        }
        bytes_uncompressed += o - coff_;
        coff_ = o;
        seconds_walk += now_s() - t0;
        return true;
    }
    if (eof_) return false;
    if (!prefetching_) {
        next_ = std::async(std::launch::async, [this] { return produce(); });
        prefetching_ = true;
    }
    Window w = next_.get();
    prefetching_ = false;
    if (!w.ok) throw open_error(w.err);
    seconds_inflate += w.s_inflate;
    seconds_walk += w.s_walk;
    bytes_uncompressed += w.inflated;
    if (w.coff_end) consumed_ = w.coff_end;
    if (w.eof) {
        eof_ = true;
        return false;
    }
    cur_ = std::move(w);  // keeps the buffer the records point into alive until the next call
    recs->swap(cur_.recs);
    next_ = std::async(std::launch::async, [this] { return produce(); });  // overlaps with the caller's decode
    prefetching_ = true;
    return true;
}

}  // namespace mthh
