// input.hpp — alignment-file input without htslib: memory-mapped file, BGZF block windows inflated on all cores
// (zlib raw inflate), BAM header, record boundaries; SAM text as the second format bam::Reader::from_path
// auto-detects (reference: src/bamutil.rs:4-25 over rust-htslib 0.50.0 / htslib).
#pragma once
#include <future>
#include <memory>
#include <string>
#include <vector>

#include "util.hpp"

namespace mthh {

extern std::atomic<int64_t> g_zlib_fallbacks;  // BGZF members the fast decoder rejected (input.cpp)


struct Header {  // bamutil.rs:13-25 — the binary reference list gives tid <-> name
    std::vector<std::string> names;
    std::vector<int64_t> lengths;
    std::string text;  // the SAM header text (BAM: the l_text bytes; SAM: the '@' lines), what `tag` copies to its output
    int tid_of(const std::string& n) const;
};

// A read-only memory map of the input file.
class MappedFile {
public:
    ~MappedFile();
    void open(const std::string& path);  // throws HostError{101, "Error opening BAM file. ..."} (bamutil.rs:8)
    const uint8_t* data() const { return data_; }
    size_t size() const { return size_; }

private:
    const uint8_t* data_ = nullptr;
    size_t size_ = 0;
    int fd_ = -1;
};

enum class Format { BAM, SAM };

// One BGZF member header at offset `o` of d[0, size): returns the member's total size (0: not a BGZF member) and where its raw
// DEFLATE payload lies, its uncompressed size (ISIZE) and CRC-32.
size_t bgzf_member_info(const uint8_t* d, size_t size, size_t o, size_t* cdata, size_t* clen, uint32_t* usize, uint32_t* crc);

// One record of the input in a format-neutral form that points into the window buffer (no copies).
struct RecordRef {
    const uint8_t* p;  // BAM: first byte after block_size; SAM: first byte of the line
    uint32_t len;      // BAM: block_size; SAM: line length without '\n'
};

// Streams the file as "windows": each window is a contiguous buffer of whole records.
class RecordStream {
public:
    RecordStream(const std::string& path, int n_threads, size_t window_bytes);
    ~RecordStream();
    const Header& header() const { return header_; }
    Format format() const { return format_; }
    // Fills `recs` with the records of the next window (pointers valid until the next call); false at end of file.
    bool next(std::vector<RecordRef>* recs);
    size_t file_size() const { return file_.size(); }
    const MappedFile& file() const { return file_; }
    // BAM: uncompressed bytes in front of the first alignment record (magic, header text, reference list)
    size_t bam_header_bytes() const;
    // BAM only, before the first next(): continue at a BGZF virtual offset (compressed offset << 16 | offset within the member's
    // output) that points at a record boundary, e.g. from the linear index of a .bai
    void seek_virtual(uint64_t voffset);
    size_t compressed_consumed() const { return consumed_; }  // compressed bytes behind the records handed out so far
    double seconds_inflate = 0, seconds_walk = 0;
    uint64_t bytes_compressed = 0, bytes_uncompressed = 0;

private:
    struct Block { size_t cdata, clen; uint32_t usize; size_t uoff; };
    // One inflated run of BGZF members (behind `headroom` free bytes).
    struct Inflated {
        std::unique_ptr<uint8_t[]> data;
        size_t headroom = 0, len = 0, coff_end = 0;
        bool ok = true;
        std::string err;
        double seconds = 0;
    };
    // One window of whole records.  The NEXT window is always produced (members inflated on the reader's own thread
    // pool, record boundaries walked) on a background thread while the caller decodes the current one.
    struct Window {
        std::unique_ptr<uint8_t[]> data;
        std::vector<RecordRef> recs;
        bool ok = true, eof = false;
        std::string err;
        double s_inflate = 0, s_walk = 0;
        size_t inflated = 0, coff_end = 0;
    };
    Inflated inflate_next(size_t headroom, size_t max_bytes = 0);  // max_bytes 0: one window
    Window produce();         // background thread: inflate + walk; owns coff_ and carry_
    void parse_bam_header();
    void parse_sam_header();
    const std::string path_;
    ThreadPool pool_;
    std::future<Window> next_;
    bool prefetching_ = false;
    std::future<Inflated> inflate_fut_;   // the run after the one being walked (producer thread only)
    bool inflating_ = false;
    static constexpr size_t kHeadroom = 4u << 20;  // room in front of a run for the record carried over from the previous one
    size_t window_bytes_;
    MappedFile file_;
    Format format_ = Format::BAM;
    Header header_;
    size_t coff_ = 0;               // next compressed offset (BAM) / next text offset (SAM)
    std::vector<uint8_t> carry_;      // bytes of the record that straddles into the next run (producer only)
    Window cur_;                      // the window whose records were handed out last
    size_t consumed_ = 0;
    size_t bam_header_len_ = 0;
    size_t skip_first_ = 0;
    bool eof_ = false;
};

}  // namespace mthh
