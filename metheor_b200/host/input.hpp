// input.hpp — alignment-file input without htslib: memory-mapped file, BGZF block windows inflated on all cores
// (zlib raw inflate), BAM header, record boundaries; SAM text as the second format bam::Reader::from_path
// auto-detects (reference: src/bamutil.rs:4-25 over rust-htslib 0.50.0 / htslib).
#pragma once
#include <string>
#include <vector>

#include "util.hpp"

namespace mthh {

struct Header {  // bamutil.rs:13-25 — the binary reference list gives tid <-> name
    std::vector<std::string> names;
    std::vector<int64_t> lengths;
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

// One record of the input in a format-neutral form that points into the window buffer (no copies).
struct RecordRef {
    const uint8_t* p;  // BAM: first byte after block_size; SAM: first byte of the line
    uint32_t len;      // BAM: block_size; SAM: line length without '\n'
};

// Streams the file as "windows": each window is a contiguous buffer of whole records.
class RecordStream {
public:
    RecordStream(const std::string& path, ThreadPool& pool, size_t window_bytes);
    const Header& header() const { return header_; }
    Format format() const { return format_; }
    // Fills `recs` with the records of the next window (pointers valid until the next call); false at end of file.
    bool next(std::vector<RecordRef>* recs);
    double seconds_inflate = 0, seconds_walk = 0;
    uint64_t bytes_compressed = 0, bytes_uncompressed = 0;

private:
    struct Block { size_t cdata, clen; uint32_t usize; size_t uoff; };
    bool fill_bam_window();   // inflates the next blocks behind the carried-over tail; false when nothing is left
    void parse_bam_header();
    void parse_sam_header();
    const std::string path_;
    ThreadPool& pool_;
    size_t window_bytes_;
    MappedFile file_;
    Format format_ = Format::BAM;
    Header header_;
    size_t coff_ = 0;               // next compressed offset (BAM) / next text offset (SAM)
    std::vector<uint8_t> buf_;      // uncompressed window: [carry | newly inflated blocks]
    size_t buf_len_ = 0, buf_pos_ = 0;
    std::vector<Block> blocks_;
    bool eof_ = false;
};

}  // namespace mthh
