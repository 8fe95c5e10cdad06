// decode.hpp — one alignment record -> the BismarkRead of the reference (src/readutil.rs:24-53, 323-345), written
// straight into structure-of-arrays chunks; optional --cpg-set filter (readutil.rs:87-95, 347-374).
#pragma once
#include <string>
#include <vector>

#include "input.hpp"

namespace mthh {

// --cpg-set: BED columns 0 (chrom) and 1 (start), tab separated, no header (readutil.rs:347-374); kept as one
// sorted position list per contig.
struct CpgSet {
    std::vector<std::vector<int32_t>> by_tid;
    void load(const std::string& path, const Header& h);  // throws HostError{101, ...} where the reference panics
    bool contains(int32_t tid, int32_t pos) const;
};

// host-internal meta bit: the alignment has indels / clips / skips, so query indices must be shipped explicitly
constexpr uint32_t SOA_META_COMPLEX = 1u << 10;

// Decoded reads of a contiguous run of records, in file order.
struct SoaChunk {
    std::vector<int32_t> tid, start, end;
    std::vector<uint32_t> meta;      // mapq | forward << 8 | SOA_META_COMPLEX
    std::vector<uint32_t> n_cpg;     // CpG calls kept per read
    std::vector<int32_t> cpg_pos;    // strand-adjusted positions (readutil.rs:332-339)
    std::vector<uint16_t> cpg_rel;   // query index (readutil.rs:335)
    std::vector<uint8_t> cpg_meth;   // 'Z' -> 1 (readutil.rs:258)
    void clear();
};

struct DecodeOptions {
    const CpgSet* cpg_set = nullptr;
    bool keep_empty = false;    // keep reads without any retained CpG call (decode-only API); the engine path drops them
    bool lpmd_order = false;    // lpmd.rs:176-181: LPMD tests mapq BEFORE building the read: low-mapq records are only
    uint32_t min_qual = 0;      // counted (and may lack XM); every other measure builds the read first and aborts
};

struct DecodeCounters {
    int64_t n_records = 0;           // every record seen
    int64_t n_dropped = 0;           // records without a retained CpG call (not shipped to the GPU)
    int64_t n_dropped_mapq_ok = 0;   // ... of which mapq >= min_qual (LPMD n_valid_read, lpmd.rs:176-189)
    int32_t max_cpgs = 0;
    int64_t max_span = 0;            // longest end - start + 1 among the kept reads
    void add(const DecodeCounters& o) {
        n_records += o.n_records; n_dropped += o.n_dropped; n_dropped_mapq_ok += o.n_dropped_mapq_ok;
        if (o.max_cpgs > max_cpgs) max_cpgs = o.max_cpgs;
        if (o.max_span > max_span) max_span = o.max_span;
    }
};

// Appends the reads of recs[begin, end) to `out`.  Throws HostError{101, "Error reading XM tag ..."} where
// BismarkRead::new panics (readutil.rs:45-51).
void decode_records(Format fmt, const Header& h, const RecordRef* recs, size_t begin, size_t end, const DecodeOptions& opt,
                    SoaChunk* out, DecodeCounters* cnt);

// tid of a record without decoding it (used to cut batches at contig boundaries)
int32_t record_tid(Format fmt, const Header& h, const RecordRef& r);

}  // namespace mthh
