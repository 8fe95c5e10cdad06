// oracle/metheor_oracle.cpp — CPU restatement of dohlee/metheor v0.1.9 (reference commit 33248b1).
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may build, load or execute anything in oracle/.
// The product (metheor_b200/) never links or calls it.
//
// What it is: a single-threaded C++17 restatement that keeps the reference's streaming structure
// (per-read hash-map accumulate + early flush + overwrite, same filter order, same f32 expressions).
// Every function cites the reference file:line it follows (paths relative to the reference root).
// The Rust binary cannot be built in this image (no cargo/rustc, crates not vendored), so this
// restatement is the checker; it is pinned against every value the reference's own unit tests hold
// for this path (tests/test_oracle_golden.py, values cited from src/*.rs #[cfg(test)] modules).
//
// Where the reference is non-deterministic the oracle takes a canonical choice, stated here:
//   * MHL f32 sum order (mhl.rs:50 iterates a HashMap): ascending l.
//   * PM / ME row order (pm.rs:76, me.rs:81 iterate a HashMap): sorted by (tid,p1,p2,p3,p4).
//   * FDRP/qFDRP reservoir sampling once depth > max_depth (fdrp.rs:90, qfdrp.rs:90 use the
//     unseeded rand::thread_rng): a seeded counter-based draw (orc_reservoir_draw below).
//     PARITY UNPINNED for piles deeper than max_depth — the reference itself is not reproducible there.
//   * Third-party behaviour restated from its published semantics (not under /root/reference):
//     rust-htslib 0.50.0 Record::reference_positions_full (M/=/X -> Some(pos) per base, I/S -> None per
//     base, D/N advance the reference only, H/P ignored); itertools 0.10.5 combinations(2)
//     (lexicographic i<j); libm log2f (called directly here).  Only pure-M forward-strand records are
//     pinned by the reference's fixtures; indel / clip / reverse-strand decoding is PARITY UNPINNED.
//
// Build: see oracle/Makefile (g++ -O2 -std=c++17 -ffp-contract=off ... -lz).

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <zlib.h>

namespace orc {

// ---------------------------------------------------------------------------------------------
// Key types — readutil.rs:279-315 (CpGPosition, is_before, Ord), :246-261 (CpG), :227-233 (Quartet)
// ---------------------------------------------------------------------------------------------
struct CpGPosition {
    int32_t tid, pos;
    // readutil.rs:290-296
    bool is_before(const CpGPosition& o, int32_t distance) const {
        if (tid > o.tid) return false;
        if (tid < o.tid) return true;
        return pos + distance < o.pos;
    }
    // readutil.rs:311-314
    bool operator<(const CpGPosition& o) const { return tid != o.tid ? tid < o.tid : pos < o.pos; }
    bool operator==(const CpGPosition& o) const { return tid == o.tid && pos == o.pos; }
};
struct CpGPositionHash {
    size_t operator()(const CpGPosition& p) const {
        uint64_t x = ((uint64_t)(uint32_t)p.tid << 32) | (uint32_t)p.pos;
        x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33;
        return (size_t)x;
    }
};
struct CpG {
    int32_t relpos;      // query index, readutil.rs:335
    CpGPosition abspos;  // strand-adjusted, readutil.rs:332-339
    bool methylated;     // c == 'Z', readutil.rs:258
};
struct Quartet {
    CpGPosition p1, p2, p3, p4;
    bool operator<(const Quartet& o) const {
        if (!(p1 == o.p1)) return p1 < o.p1;
        if (!(p2 == o.p2)) return p2 < o.p2;
        if (!(p3 == o.p3)) return p3 < o.p3;
        return p4 < o.p4;
    }
};

// ---------------------------------------------------------------------------------------------
// BAM / SAM record model (what rust-htslib hands to readutil.rs) and the BismarkRead built from it
// ---------------------------------------------------------------------------------------------
struct Record {
    int32_t tid = -1, pos = -1;
    uint8_t mapq = 0;
    uint16_t flag = 0;
    std::vector<uint32_t> cigar;  // BAM encoding len<<4|op, op index into "MIDNSHP=X"
    bool has_xm = false;
    std::string xm;
};
struct Header {
    std::vector<std::string> names;
    std::vector<int64_t> lengths;
    int tid_of(const std::string& n) const {
        for (size_t i = 0; i < names.size(); i++) if (names[i] == n) return (int)i;
        return -1;
    }
};

struct BismarkRead {
    int32_t start_pos = -1, end_pos = -1;
    uint8_t mapq = 0;
    std::vector<CpG> cpgs;
    // readutil.rs:55-60
    bool first_cpg(CpGPosition* out) const {
        if (cpgs.empty()) return false;
        *out = cpgs[0].abspos;
        return true;
    }
};

// rust-htslib 0.50.0 bam/ext.rs reference_positions_full (published semantics, see header comment):
// one entry per query base; -1 stands for None.
static void reference_positions_full(const Record& r, std::vector<int64_t>* out) {
    out->clear();
    int64_t ref = r.pos;
    for (uint32_t c : r.cigar) {
        uint32_t len = c >> 4, op = c & 15;
        switch (op) {
            case 0: case 7: case 8:  // M = X
                for (uint32_t i = 0; i < len; i++) out->push_back(ref++);
                break;
            case 1: case 4:  // I S
                for (uint32_t i = 0; i < len; i++) out->push_back(-1);
                break;
            case 2: case 3:  // D N
                ref += len;
                break;
            default:  // H P
                break;
        }
    }
}

// readutil.rs:24-53 (BismarkRead::new) + :323-345 (get_cpgs).  Returns false where the reference panics
// ("Error reading XM tag in BAM record...").
static bool bismark_read_new(const Record& r, BismarkRead* br) {
    std::vector<int64_t> positions;
    reference_positions_full(r, &positions);
    br->start_pos = -1;
    br->end_pos = -1;
    br->mapq = r.mapq;
    br->cpgs.clear();
    for (int64_t p : positions) {  // readutil.rs:28-33 (.flatten() skips None)
        if (p < 0) continue;
        if (br->start_pos == -1) br->start_pos = (int32_t)p;
        br->end_pos = (int32_t)p;
    }
    if (!r.has_xm) return false;  // readutil.rs:45-51
    size_t n = std::min(positions.size(), r.xm.size());  // zip, readutil.rs:326
    bool fwd = (r.flag == 0) || (r.flag == 99) || (r.flag == 147);  // readutil.rs:332
    for (size_t rel = 0; rel < n; rel++) {
        char c = r.xm[rel];
        if (c != 'z' && c != 'Z') continue;  // readutil.rs:327-329
        if (positions[rel] < 0) continue;     // readutil.rs:331
        int32_t pos = fwd ? (int32_t)positions[rel] : (int32_t)(positions[rel] - 1);  // :334, :338
        br->cpgs.push_back(CpG{(int32_t)rel, CpGPosition{r.tid, pos}, c == 'Z'});
    }
    return true;
}

typedef std::unordered_set<CpGPosition, CpGPositionHash> CpGSet;

// readutil.rs:87-95
static void filter_isin(BismarkRead* br, const CpGSet& target) {
    std::vector<CpG> keep;
    for (const CpG& c : br->cpgs) if (target.count(c.abspos)) keep.push_back(c);
    br->cpgs.swap(keep);
}

// readutil.rs:134-145
static bool is_discordant_read(const BismarkRead& br) {
    bool init = br.cpgs[0].methylated;
    bool disc = false;
    for (const CpG& c : br.cpgs) if (c.methylated != init) disc = true;
    return disc;
}

// readutil.rs:147-164 — map l -> count (ordered map: canonical ascending-l iteration)
static void get_stretch_info(const BismarkRead& br, std::map<int32_t, int32_t>* info) {
    info->clear();
    int32_t cur = 0;
    for (const CpG& c : br.cpgs) {
        if (c.methylated) {
            cur += 1;
            for (int32_t l = 1; l < cur + 1; l++) (*info)[l] += 1;
        } else {
            cur = 0;
        }
    }
}

struct PairObs { CpGPosition a, b; bool concordant; };

// readutil.rs:166-224
static void pairwise_concordance(const BismarkRead& br, int32_t min_distance, int32_t max_distance,
                                 int32_t* n_conc, int32_t* n_disc, std::vector<PairObs>* pairs) {
    std::vector<CpG> anchors;
    int32_t min_anchor_pos = -1;
    *n_conc = 0;
    *n_disc = 0;
    for (const CpG& cpg : br.cpgs) {
        if (min_anchor_pos != -1) {
            while ((cpg.relpos - min_anchor_pos > max_distance) && !anchors.empty()) {
                anchors.erase(anchors.begin());
                min_anchor_pos = anchors.empty() ? -1 : anchors[0].relpos;
            }
        }
        for (const CpG& a : anchors) {
            if (cpg.relpos - a.relpos < min_distance) continue;
            bool conc = (a.methylated == cpg.methylated);
            if (conc) *n_conc += 1; else *n_disc += 1;
            if (pairs) pairs->push_back(PairObs{a.abspos, cpg.abspos, conc});
        }
        if (min_anchor_pos == -1) min_anchor_pos = cpg.relpos;
        anchors.push_back(cpg);
    }
}

// ---------------------------------------------------------------------------------------------
// Read stream: either decoded records (BAM/SAM) or BismarkRead-level SoA handed in by the tests.
// ---------------------------------------------------------------------------------------------
struct ReadSet {
    Header header;
    std::vector<BismarkRead> reads;  // in file order; cpgs NOT yet cpg-set filtered
    std::vector<uint8_t> xm_ok;      // 0 where the reference would panic on a missing XM tag
    std::string error;
};

// ---------------------------------------------------------------------------------------------
// PDR — pdr.rs:119-212
// ---------------------------------------------------------------------------------------------
struct PdrRow { CpGPosition pos; float pdr; uint32_t n_conc, n_disc; };

static float f32div(float a, float b) { volatile float r = a / b; return r; }

static void pdr_compute(const ReadSet& rs, uint32_t min_depth, size_t min_cpgs, uint8_t min_qual,
                        const CpGSet* target, std::vector<PdrRow>* out) {
    struct PDRResult { uint32_t n_conc = 0, n_disc = 0; };  // pdr.rs:12-50
    std::unordered_map<CpGPosition, PDRResult, CpGPositionHash> cpg2reads;
    std::map<CpGPosition, PdrRow> result;  // BTreeMap, pdr.rs:136
    auto emit = [&](const CpGPosition& cpg, const PDRResult& r) {
        // pdr.rs:47-49
        float pdr = f32div((float)r.n_disc, (float)r.n_conc + (float)r.n_disc);
        result[cpg] = PdrRow{cpg, pdr, r.n_conc, r.n_disc};  // insert overwrites, pdr.rs:164
    };
    for (const BismarkRead& src : rs.reads) {
        BismarkRead br = src;
        if (target) filter_isin(&br, *target);       // pdr.rs:142-144
        if (br.cpgs.size() < min_cpgs) continue;      // pdr.rs:147
        if (br.mapq < min_qual) continue;             // pdr.rs:150
        if (br.cpgs.empty()) continue;                // pdr.rs:155
        CpGPosition first = br.cpgs[0].abspos;
        for (auto it = cpg2reads.begin(); it != cpg2reads.end();) {  // retain, pdr.rs:160-177
            if (it->first.is_before(first, 150)) {
                if (it->second.n_conc + it->second.n_disc >= min_depth) emit(it->first, it->second);
                it = cpg2reads.erase(it);
            } else {
                ++it;
            }
        }
        bool disc = is_discordant_read(br);  // pdr.rs:185
        for (const CpG& c : br.cpgs) {       // pdr.rs:180-191
            PDRResult& r = cpg2reads[c.abspos];
            if (disc) r.n_disc += 1; else r.n_conc += 1;
        }
    }
    for (auto& kv : cpg2reads)  // pdr.rs:199-210
        if (kv.second.n_conc + kv.second.n_disc >= min_depth) emit(kv.first, kv.second);
    out->clear();
    for (auto& kv : result) out->push_back(kv.second);
}

// ---------------------------------------------------------------------------------------------
// LPMD — lpmd.rs:154-202, :51-55, :89-122
// ---------------------------------------------------------------------------------------------
struct LpmdPairRow { CpGPosition a, b; float lpmd; int32_t n_conc, n_disc; };
struct LpmdResult {
    int32_t n_read = 0, n_valid_read = 0, n_conc = 0, n_disc = 0;
    float lpmd = 0;
    std::vector<LpmdPairRow> pairs;
};

static void lpmd_compute(const ReadSet& rs, int32_t min_distance, int32_t max_distance, uint8_t min_qual,
                         const CpGSet* target, bool want_pairs, LpmdResult* res) {
    std::map<std::pair<CpGPosition, CpGPosition>, std::pair<int32_t, int32_t>> pair2n;
    *res = LpmdResult();
    std::vector<PairObs> obs;
    for (const BismarkRead& src : rs.reads) {
        res->n_read += 1;                  // lpmd.rs:176
        if (src.mapq < min_qual) continue;  // lpmd.rs:177
        BismarkRead br = src;
        if (target) filter_isin(&br, *target);
        int32_t c, d;
        obs.clear();
        pairwise_concordance(br, min_distance, max_distance, &c, &d, want_pairs ? &obs : nullptr);
        res->n_valid_read += 1;
        res->n_conc += c;  // i32, lpmd.rs:190-191
        res->n_disc += d;
        for (const PairObs& o : obs) {  // lpmd.rs:70-87
            auto& e = pair2n[std::make_pair(o.a, o.b)];
            if (o.concordant) e.first += 1; else e.second += 1;
        }
    }
    // lpmd.rs:51-55: i32 add, then cast
    res->lpmd = f32div((float)res->n_disc, (float)(res->n_conc + res->n_disc));
    for (auto& kv : pair2n) {  // sorted by key, lpmd.rs:94; value lpmd.rs:111
        float v = f32div((float)kv.second.second, (float)kv.second.first + (float)kv.second.second);
        res->pairs.push_back(LpmdPairRow{kv.first.first, kv.first.second, v, kv.second.first, kv.second.second});
    }
}

// ---------------------------------------------------------------------------------------------
// MHL — mhl.rs:135-208, AssociatedReads mhl.rs:12-81
// ---------------------------------------------------------------------------------------------
struct SiteRow { CpGPosition pos; float value; };

struct MhlReads {
    std::map<int32_t, int32_t> stretch_info;  // canonical ascending l
    std::vector<int32_t> num_cpgs;
    size_t max_num_cpgs = 0;
    // mhl.rs:43-73
    float compute_mhl() const {
        volatile float mhl = 0.0f;
        volatile float l_sum = 0.0f;
        for (size_t l = 1; l < max_num_cpgs + 1; l++) l_sum = l_sum + (float)l;
        for (auto& kv : stretch_info) {
            int32_t l = kv.first;
            float dom = (float)kv.second;
            volatile float denom = 0.0f;
            for (int32_t n : num_cpgs)
                if (n >= l) denom = denom + (float)(n - l + 1);
            volatile float num = (float)l * dom;
            volatile float term = num / denom;
            mhl = mhl + term;
        }
        mhl = mhl / l_sum;
        return mhl;
    }
};

static void mhl_compute(const ReadSet& rs, uint32_t min_depth, size_t min_cpgs, uint8_t min_qual,
                        const CpGSet* target, std::vector<SiteRow>* out) {
    std::unordered_map<CpGPosition, MhlReads, CpGPositionHash> cpg2reads;
    std::map<CpGPosition, float> result;
    std::map<int32_t, int32_t> info;
    for (const BismarkRead& src : rs.reads) {
        BismarkRead br = src;
        if (target) filter_isin(&br, *target);
        CpGPosition first;
        if (br.first_cpg(&first)) {  // mhl.rs:162-173: flush from ANY read with >=1 CpG
            for (auto it = cpg2reads.begin(); it != cpg2reads.end();) {
                if (it->first < first) {
                    if (it->second.num_cpgs.size() >= min_depth) result[it->first] = it->second.compute_mhl();
                    it = cpg2reads.erase(it);
                } else {
                    ++it;
                }
            }
        }
        if (br.mapq < min_qual) continue;         // mhl.rs:176
        if (br.cpgs.size() < min_cpgs) continue;   // mhl.rs:181
        get_stretch_info(br, &info);               // mhl.rs:191 (recomputed per CpG there; same value)
        for (const CpG& c : br.cpgs) {             // mhl.rs:185-192
            MhlReads& r = cpg2reads[c.abspos];
            r.num_cpgs.push_back((int32_t)br.cpgs.size());  // mhl.rs:75-80
            if (br.cpgs.size() >= r.max_num_cpgs) r.max_num_cpgs = br.cpgs.size();
            for (auto& kv : info) r.stretch_info[kv.first] += kv.second;  // mhl.rs:36-41
        }
    }
    for (auto& kv : cpg2reads)  // mhl.rs:201-205
        if (kv.second.num_cpgs.size() >= min_depth) result[kv.first] = kv.second.compute_mhl();
    out->clear();
    for (auto& kv : result) out->push_back(SiteRow{kv.first, kv.second});
}

// ---------------------------------------------------------------------------------------------
// PM / ME — pm.rs:85-128 + :42-51, me.rs:90-132 + :42-55, quartets readutil.rs:97-132
// ---------------------------------------------------------------------------------------------
struct QuartetRow { Quartet q; uint32_t counts[16]; float pm, me; uint32_t depth; };

static float compute_pm(const uint32_t* counts) {  // pm.rs:42-51
    uint32_t total = 0;
    for (int k = 0; k < 16; k++) total += counts[k];
    volatile float pm = 1.0f;
    for (int k = 0; k < 16; k++) {
        volatile float a = (float)counts[k] / (float)total;
        volatile float b = (float)counts[k] / (float)total;
        volatile float sq = a * b;
        pm = pm - sq;
    }
    return pm;
}
static float compute_me(const uint32_t* counts) {  // me.rs:42-55
    volatile float me = 0.0f;
    uint32_t total = 0;
    for (int k = 0; k < 16; k++) total += counts[k];
    for (int k = 0; k < 16; k++) {
        volatile float p = (float)counts[k] / (float)total;
        if (counts[k] > 0) {
            volatile float lg = log2f(p);  // f32::log2 -> libm log2f
            volatile float t = p * lg;
            me = me + t;
        }
    }
    me = me * -0.25f;
    return me;
}

static void quartet_compute(const ReadSet& rs, uint32_t min_depth, uint8_t min_qual, const CpGSet* target,
                            std::vector<QuartetRow>* out) {
    std::map<Quartet, QuartetRow> q2s;  // canonical sorted order (reference: HashMap, unordered)
    for (const BismarkRead& src : rs.reads) {
        BismarkRead br = src;
        if (target) filter_isin(&br, *target);
        if (br.mapq < min_qual) continue;  // pm.rs:111, me.rs:115
        if (br.cpgs.size() < 4) continue;  // readutil.rs:101
        for (size_t i = 0; i < br.cpgs.size() - 3; i++) {  // readutil.rs:105-129
            Quartet q{br.cpgs[i].abspos, br.cpgs[i + 1].abspos, br.cpgs[i + 2].abspos, br.cpgs[i + 3].abspos};
            int p = 0;
            if (br.cpgs[i].methylated) p += 8;
            if (br.cpgs[i + 1].methylated) p += 4;
            if (br.cpgs[i + 2].methylated) p += 2;
            if (br.cpgs[i + 3].methylated) p += 1;
            auto it = q2s.find(q);
            if (it == q2s.end()) {
                QuartetRow row;
                row.q = q;
                memset(row.counts, 0, sizeof(row.counts));
                it = q2s.emplace(q, row).first;
            }
            it->second.counts[p] += 1;  // pm.rs:38-40
        }
    }
    out->clear();
    for (auto& kv : q2s) {
        QuartetRow row = kv.second;
        row.depth = 0;
        for (int k = 0; k < 16; k++) row.depth += row.counts[k];
        if (row.depth < min_depth) continue;  // pm.rs:77, me.rs:82
        row.pm = compute_pm(row.counts);
        row.me = compute_me(row.counts);
        out->push_back(row);
    }
}

// ---------------------------------------------------------------------------------------------
// FDRP / qFDRP — fdrp.rs:10-145,176-246 ; qfdrp.rs:97-157,188-258
// ---------------------------------------------------------------------------------------------
static const int32_t MAX_READ_LEN = 201;  // fdrp.rs:10
static const int WIN = MAX_READ_LEN * 2 + 1;

// Seeded stand-in for rand::thread_rng().gen_range(1..total+1) (fdrp.rs:90).  Same formula is used by the
// CUDA engine (metheor_b200/csrc) so that depth > max_depth runs are reproducible between the two.
static inline uint32_t reservoir_draw(uint64_t seed, int32_t tid, int32_t pos, uint32_t total) {
    uint64_t x = seed ^ ((uint64_t)(uint32_t)tid * 0x9E3779B97F4A7C15ULL) ^
                 ((uint64_t)(uint32_t)pos * 0xBF58476D1CE4E5B9ULL) ^ ((uint64_t)total * 0x94D049BB133111EBULL);
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return 1u + (uint32_t)(((x >> 32) * (uint64_t)total) >> 32);  // in 1..=total
}

struct Pile {
    CpGPosition pos;
    std::vector<std::vector<uint8_t>> reads;  // each WIN bytes, fdrp.rs:22
    int32_t num_total_read = 0, num_sampled_read = 0;
    size_t max_depth = 0;
    bool oob = false;  // the reference would panic (index out of bounds), see add_read

    // fdrp.rs:51-95
    void add_read(const BismarkRead& br, uint64_t seed) {
        std::vector<uint8_t> nr(WIN, 0);
        int32_t s = MAX_READ_LEN + (br.start_pos - pos.pos);
        int32_t e = MAX_READ_LEN + (br.end_pos - pos.pos);
        if (s < 0) return;
        if (e > MAX_READ_LEN * 2) return;
        for (int32_t p = s; p < e + 1; p++) nr[p] |= 1;
        for (const CpG& c : br.cpgs) {
            int64_t rel = (int64_t)MAX_READ_LEN + ((int64_t)c.abspos.pos - pos.pos);
            if (rel < 0 || rel >= WIN) { oob = true; continue; }  // reference: panic
            nr[rel] |= 2;
            if (c.methylated) nr[rel] |= 4;
        }
        if (num_total_read < (int32_t)max_depth) {
            num_sampled_read += 1;
            num_total_read += 1;
            reads.push_back(nr);
        } else {
            num_total_read += 1;
            uint32_t j = reservoir_draw(seed, pos.tid, pos.pos, (uint32_t)num_total_read);
            if (j <= (uint32_t)max_depth) reads[j - 1] = nr;
        }
    }
    int32_t overlap_bases(size_t i, size_t j) const {  // fdrp.rs:97-107
        int32_t n = 0;
        for (int p = 0; p < WIN; p++) n += (reads[i][p] & reads[j][p]) & 1;
        return n;
    }
    int32_t overlap_cpgs(size_t i, size_t j) const {  // qfdrp.rs:109-119
        int32_t n = 0;
        for (int p = 0; p < WIN; p++) n += ((reads[i][p] >> 1) & (reads[j][p] >> 1)) & 1;
        return n;
    }
    int32_t hamming(size_t i, size_t j) const {  // qfdrp.rs:121-135 (fdrp.rs:109-122 is hamming > 0)
        int32_t d = 0;
        for (int p = 0; p < WIN; p++)
            if (((reads[i][p] & reads[j][p]) & 3) == 3 && (((reads[i][p] ^ reads[j][p]) & 4) >> 2) == 1) d += 1;
        return d;
    }
    float compute_fdrp(int32_t min_overlap) const {  // fdrp.rs:124-145
        size_t n = (size_t)num_sampled_read;
        volatile float fdrp = 0.0f;
        for (size_t i = 0; i < n; i++)
            for (size_t j = i + 1; j < n; j++) {  // combinations(2): lexicographic
                if (overlap_bases(i, j) < min_overlap) continue;
                if (hamming(i, j) > 0) fdrp = fdrp + 1.0f;
            }
        volatile float den = (float)(n * (n - 1)) / 2.0f;
        fdrp = fdrp / den;
        return fdrp;
    }
    float compute_qfdrp(int32_t min_overlap) const {  // qfdrp.rs:137-157
        size_t n = (size_t)num_sampled_read;
        volatile float q = 0.0f;
        for (size_t i = 0; i < n; i++)
            for (size_t j = i + 1; j < n; j++) {
                int32_t ob = overlap_bases(i, j);
                int32_t oc = overlap_cpgs(i, j);
                if (ob < min_overlap) continue;
                volatile float t = (float)hamming(i, j) / (float)oc;
                q = q + t;
            }
        volatile float den = (float)(n * (n - 1)) / 2.0f;
        q = q / den;
        return q;
    }
};

static void fdrp_compute(const ReadSet& rs, bool quantitative, uint8_t min_qual, size_t min_depth, size_t max_depth,
                         int32_t min_overlap, const CpGSet* target, uint64_t seed, std::vector<SiteRow>* out,
                         int* oob_flag) {
    std::map<CpGPosition, Pile> cpg2reads;  // BTreeMap, fdrp.rs:193
    std::map<CpGPosition, float> result;
    auto value = [&](const Pile& p) { return quantitative ? p.compute_qfdrp(min_overlap) : p.compute_fdrp(min_overlap); };
    for (const BismarkRead& src : rs.reads) {
        BismarkRead br = src;
        if (target) filter_isin(&br, *target);
        if (br.mapq < min_qual) continue;  // fdrp.rs:205
        if (br.cpgs.empty()) continue;     // fdrp.rs:208
        CpGPosition first = br.cpgs[0].abspos;
        for (auto it = cpg2reads.begin(); it != cpg2reads.end();) {  // fdrp.rs:213-222 (strict <)
            if (it->first < first) {
                // min_depth==0 with an empty pile underflows in the reference (fdrp.rs:143); skip such piles.
                if ((size_t)it->second.num_sampled_read >= min_depth && it->second.num_sampled_read > 0)
                    result[it->first] = value(it->second);
                it = cpg2reads.erase(it);
            } else {
                ++it;
            }
        }
        for (const CpG& c : br.cpgs) {  // fdrp.rs:225-231
            auto it = cpg2reads.find(c.abspos);
            if (it == cpg2reads.end()) {
                Pile p;
                p.pos = c.abspos;
                p.max_depth = max_depth;
                it = cpg2reads.emplace(c.abspos, p).first;
            }
            it->second.add_read(br, seed);
            if (it->second.oob && oob_flag) *oob_flag = 1;
        }
    }
    for (auto& kv : cpg2reads)  // fdrp.rs:239-243
        if ((size_t)kv.second.num_sampled_read >= min_depth && kv.second.num_sampled_read > 0)
            result[kv.first] = value(kv.second);
    out->clear();
    for (auto& kv : result) out->push_back(SiteRow{kv.first, kv.second});
}

// ---------------------------------------------------------------------------------------------
// Input: BGZF/BAM and SAM text (rust-htslib bam::Reader::from_path auto-detects, bamutil.rs:4-11)
// ---------------------------------------------------------------------------------------------
static bool read_file(const std::string& path, std::vector<uint8_t>* buf) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf->resize((size_t)n);
    size_t got = n ? fread(buf->data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return got == (size_t)n;
}

static bool bgzf_inflate_all(const std::vector<uint8_t>& in, std::vector<uint8_t>* out, std::string* err) {
    size_t off = 0;
    out->clear();
    while (off + 18 <= in.size()) {
        if (in[off] != 31 || in[off + 1] != 139) { *err = "not a gzip member"; return false; }
        uint16_t xlen = in[off + 10] | (in[off + 11] << 8);
        size_t x = off + 12, xend = x + xlen;
        int bsize = -1;
        while (x + 4 <= xend) {
            uint16_t slen = in[x + 2] | (in[x + 3] << 8);
            if (in[x] == 66 && in[x + 1] == 67 && slen == 2) bsize = in[x + 4] | (in[x + 5] << 8);
            x += 4 + slen;
        }
        if (bsize < 0) { *err = "gzip member without BGZF BC field"; return false; }
        size_t total = (size_t)bsize + 1;
        if (off + total > in.size()) { *err = "truncated BGZF block"; return false; }
        size_t cdata = off + 12 + xlen, clen = total - xlen - 12 - 8;
        uint32_t isize = in[off + total - 4] | (in[off + total - 3] << 8) | (in[off + total - 2] << 16) |
                         ((uint32_t)in[off + total - 1] << 24);
        size_t old = out->size();
        out->resize(old + isize);
        if (isize) {
            z_stream zs;
            memset(&zs, 0, sizeof(zs));
            if (inflateInit2(&zs, -15) != Z_OK) { *err = "inflateInit2"; return false; }
            zs.next_in = (Bytef*)&in[cdata];
            zs.avail_in = (uInt)clen;
            zs.next_out = out->data() + old;
            zs.avail_out = isize;
            int rc = inflate(&zs, Z_FINISH);
            inflateEnd(&zs);
            if (rc != Z_STREAM_END) { *err = "inflate failed"; return false; }
        }
        off += total;
    }
    return true;
}

static inline int32_t le32(const uint8_t* p) { return (int32_t)(p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24)); }

static bool parse_bam(const std::vector<uint8_t>& d, Header* h, std::vector<Record>* recs, std::string* err) {
    if (d.size() < 12 || memcmp(d.data(), "BAM\1", 4) != 0) { *err = "bad BAM magic"; return false; }
    size_t o = 4;
    int32_t l_text = le32(&d[o]); o += 4 + (size_t)l_text;
    int32_t n_ref = le32(&d[o]); o += 4;
    for (int i = 0; i < n_ref; i++) {
        int32_t l_name = le32(&d[o]); o += 4;
        h->names.push_back(std::string((const char*)&d[o], (size_t)l_name - 1)); o += (size_t)l_name;
        h->lengths.push_back(le32(&d[o])); o += 4;
    }
    while (o + 4 <= d.size()) {
        int32_t bs = le32(&d[o]); o += 4;
        if (o + (size_t)bs > d.size()) { *err = "truncated BAM record"; return false; }
        const uint8_t* p = &d[o];
        Record r;
        r.tid = le32(p); r.pos = le32(p + 4);
        uint8_t l_read_name = p[8]; r.mapq = p[9];
        uint16_t n_cigar = p[12] | (p[13] << 8);
        r.flag = p[14] | (p[15] << 8);
        int32_t l_seq = le32(p + 16);
        size_t q = 32 + l_read_name;
        for (int i = 0; i < n_cigar; i++) { r.cigar.push_back((uint32_t)le32(p + q)); q += 4; }
        q += (size_t)(l_seq + 1) / 2 + (size_t)l_seq;
        while (q + 3 <= (size_t)bs) {  // aux fields
            char t0 = p[q], t1 = p[q + 1], ty = p[q + 2];
            q += 3;
            size_t len = 0;
            switch (ty) {
                case 'A': case 'c': case 'C': len = 1; break;
                case 's': case 'S': len = 2; break;
                case 'i': case 'I': case 'f': len = 4; break;
                case 'Z': case 'H': { size_t e = q; while (e < (size_t)bs && p[e]) e++; len = e - q + 1; break; }
                case 'B': {
                    char st = p[q]; int32_t cnt = le32(p + q + 1);
                    size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
                    len = 5 + es * (size_t)cnt; break;
                }
                default: *err = "unknown aux type"; return false;
            }
            if (t0 == 'X' && t1 == 'M') {
                if (ty == 'Z') { r.has_xm = true; r.xm.assign((const char*)p + q, len - 1); }
                // non-string XM: reference panics (readutil.rs:45-47) -> has_xm stays false
            }
            q += len;
        }
        recs->push_back(std::move(r));
        o += (size_t)bs;
    }
    return true;
}

static bool parse_sam(const std::vector<uint8_t>& d, Header* h, std::vector<Record>* recs, std::string* err) {
    size_t o = 0;
    while (o < d.size()) {
        size_t e = o;
        while (e < d.size() && d[e] != '\n') e++;
        std::string line((const char*)&d[o], e - o);
        o = e + 1;
        if (line.empty()) continue;
        std::vector<std::string> f;
        size_t s = 0;
        while (true) {
            size_t t = line.find('\t', s);
            if (t == std::string::npos) { f.push_back(line.substr(s)); break; }
            f.push_back(line.substr(s, t - s));
            s = t + 1;
        }
        if (line[0] == '@') {
            if (f[0] == "@SQ") {
                std::string sn; int64_t ln = 0;
                for (auto& x : f) {
                    if (x.rfind("SN:", 0) == 0) sn = x.substr(3);
                    if (x.rfind("LN:", 0) == 0) ln = atoll(x.c_str() + 3);
                }
                h->names.push_back(sn); h->lengths.push_back(ln);
            }
            continue;
        }
        if (f.size() < 11) { *err = "short SAM line"; return false; }
        Record r;
        r.flag = (uint16_t)atoi(f[1].c_str());
        r.tid = f[2] == "*" ? -1 : h->tid_of(f[2]);
        r.pos = atoi(f[3].c_str()) - 1;
        r.mapq = (uint8_t)atoi(f[4].c_str());
        if (f[5] != "*") {
            const char* c = f[5].c_str();
            while (*c) {
                uint32_t len = 0;
                while (*c >= '0' && *c <= '9') len = len * 10 + (uint32_t)(*c++ - '0');
                const char* ops = "MIDNSHP=X";
                const char* w = strchr(ops, *c++);
                if (!w) { *err = "bad CIGAR"; return false; }
                r.cigar.push_back((len << 4) | (uint32_t)(w - ops));
            }
        }
        for (size_t i = 11; i < f.size(); i++)
            if (f[i].rfind("XM:Z:", 0) == 0) { r.has_xm = true; r.xm = f[i].substr(5); }
        recs->push_back(std::move(r));
    }
    return true;
}

static bool load_alignment_file(const std::string& path, ReadSet* rs) {
    std::vector<uint8_t> raw;
    if (!read_file(path, &raw)) { rs->error = "Error opening BAM file. file not found: " + path; return false; }
    std::vector<Record> recs;
    std::string err;
    bool ok;
    if (raw.size() >= 2 && raw[0] == 31 && raw[1] == 139) {
        std::vector<uint8_t> d;
        ok = bgzf_inflate_all(raw, &d, &err) && parse_bam(d, &rs->header, &recs, &err);
    } else if (!raw.empty() && raw[0] == '@') {
        ok = parse_sam(raw, &rs->header, &recs, &err);
    } else {
        ok = false; err = "unrecognised format";
    }
    if (!ok) { rs->error = "Error opening BAM file. " + err + ": " + path; return false; }
    for (const Record& r : recs) {
        BismarkRead br;
        bool xm = bismark_read_new(r, &br);
        rs->reads.push_back(br);
        rs->xm_ok.push_back(xm ? 1 : 0);
    }
    return true;
}

// readutil.rs:347-374 — BED columns 0,1, tab split, no header tolerated
static bool load_cpg_set(const std::string& path, const Header& h, CpGSet* set, std::string* err) {
    std::vector<uint8_t> raw;
    if (!read_file(path, &raw)) { *err = "Could not read target CpG file."; return false; }
    size_t o = 0;
    while (o < raw.size()) {
        size_t e = o;
        while (e < raw.size() && raw[e] != '\n') e++;
        std::string line((const char*)&raw[o], e - o);
        o = e + 1;
        if (!line.empty() && line.back() == '\r') line.pop_back();  // str::lines() strips \r\n
        if (line.empty() && o >= raw.size()) break;
        size_t t1 = line.find('\t');
        if (t1 == std::string::npos) { *err = "malformed BED line"; return false; }
        size_t t2 = line.find('\t', t1 + 1);
        std::string chrom = line.substr(0, t1);
        std::string ps = line.substr(t1 + 1, t2 == std::string::npos ? std::string::npos : t2 - t1 - 1);
        int tid = h.tid_of(chrom);
        if (tid < 0) { *err = "unknown chromosome in CpG set"; return false; }
        char* endp = nullptr;
        long pos = strtol(ps.c_str(), &endp, 10);
        if (endp == ps.c_str() || *endp) { *err = "bad position in CpG set"; return false; }
        set->insert(CpGPosition{(int32_t)tid, (int32_t)pos});
    }
    return true;
}

// Rust `{}` for f32: shortest round-trip digits, positional notation, "NaN", "inf", "-0".
static std::string fmt_f32(float v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    char buf[128];
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}

}  // namespace orc

// =============================================================================================
// C ABI (ctypes from tests / bench) — handle-based
// =============================================================================================
using namespace orc;

struct orc_handle {
    ReadSet rs;
    CpGSet target;
    bool has_target = false;
    std::vector<PdrRow> pdr;
    std::vector<SiteRow> site;
    std::vector<QuartetRow> quart;
    LpmdResult lpmd;
    int oob = 0;
};

extern "C" {

orc_handle* orc_open(const char* path) {
    orc_handle* h = new orc_handle();
    load_alignment_file(path, &h->rs);
    return h;
}
// BismarkRead-level SoA (what the engine's mth_batch carries): per read tid/start/end/mapq, CSR of CpGs.
orc_handle* orc_from_soa(int64_t n_reads, const int32_t* tid, const int32_t* start, const int32_t* end,
                         const uint8_t* mapq, const int64_t* cpg_off, const int32_t* cpg_pos,
                         const int32_t* cpg_rel, const uint8_t* cpg_meth) {
    orc_handle* h = new orc_handle();
    h->rs.reads.resize((size_t)n_reads);
    for (int64_t i = 0; i < n_reads; i++) {
        BismarkRead& br = h->rs.reads[(size_t)i];
        br.start_pos = start[i]; br.end_pos = end[i]; br.mapq = mapq[i];
        for (int64_t k = cpg_off[i]; k < cpg_off[i + 1]; k++)
            br.cpgs.push_back(CpG{cpg_rel ? cpg_rel[k] : (int32_t)(k - cpg_off[i]), CpGPosition{tid[i], cpg_pos[k]}, cpg_meth[k] != 0});
    }
    h->rs.xm_ok.assign((size_t)n_reads, 1);
    return h;
}
void orc_close(orc_handle* h) { delete h; }
const char* orc_error(orc_handle* h) { return h->rs.error.c_str(); }
int64_t orc_n_reads(orc_handle* h) { return (int64_t)h->rs.reads.size(); }
int64_t orc_n_cpgs(orc_handle* h) { int64_t n = 0; for (auto& r : h->rs.reads) n += (int64_t)r.cpgs.size(); return n; }
int orc_n_ref(orc_handle* h) { return (int)h->rs.header.names.size(); }
const char* orc_ref_name(orc_handle* h, int i) { return h->rs.header.names[(size_t)i].c_str(); }
int64_t orc_ref_len(orc_handle* h, int i) { return h->rs.header.lengths[(size_t)i]; }
int orc_all_xm_ok(orc_handle* h) { for (uint8_t x : h->rs.xm_ok) if (!x) return 0; return 1; }

// export decoded reads (used to cross-check the product's host decoder and to build fixtures)
void orc_export_reads(orc_handle* h, int32_t* tid, int32_t* start, int32_t* end, uint8_t* mapq, int64_t* cpg_off,
                      int32_t* cpg_pos, int32_t* cpg_rel, uint8_t* cpg_meth) {
    int64_t k = 0;
    for (size_t i = 0; i < h->rs.reads.size(); i++) {
        const BismarkRead& br = h->rs.reads[i];
        tid[i] = br.cpgs.empty() ? -1 : br.cpgs[0].abspos.tid;
        start[i] = br.start_pos; end[i] = br.end_pos; mapq[i] = br.mapq;
        cpg_off[i] = k;
        for (const CpG& c : br.cpgs) { cpg_pos[k] = c.abspos.pos; cpg_rel[k] = c.relpos; cpg_meth[k] = c.methylated; k++; }
    }
    cpg_off[h->rs.reads.size()] = k;
}

int orc_set_cpg_set_file(orc_handle* h, const char* path) {
    std::string err;
    h->target.clear();
    if (!load_cpg_set(path, h->rs.header, &h->target, &err)) { h->rs.error = err; return -1; }
    h->has_target = true;
    return 0;
}
void orc_set_cpg_set(orc_handle* h, int64_t n, const int32_t* tid, const int32_t* pos) {
    h->target.clear();
    for (int64_t i = 0; i < n; i++) h->target.insert(CpGPosition{tid[i], pos[i]});
    h->has_target = true;
}
void orc_clear_cpg_set(orc_handle* h) { h->target.clear(); h->has_target = false; }

int64_t orc_pdr(orc_handle* h, uint32_t min_depth, uint32_t min_cpgs, uint32_t min_qual) {
    pdr_compute(h->rs, min_depth, min_cpgs, (uint8_t)min_qual, h->has_target ? &h->target : nullptr, &h->pdr);
    return (int64_t)h->pdr.size();
}
void orc_pdr_rows(orc_handle* h, int32_t* tid, int32_t* pos, float* pdr, uint32_t* nc, uint32_t* nd) {
    for (size_t i = 0; i < h->pdr.size(); i++) {
        tid[i] = h->pdr[i].pos.tid; pos[i] = h->pdr[i].pos.pos; pdr[i] = h->pdr[i].pdr;
        nc[i] = h->pdr[i].n_conc; nd[i] = h->pdr[i].n_disc;
    }
}
int64_t orc_mhl(orc_handle* h, uint32_t min_depth, uint32_t min_cpgs, uint32_t min_qual) {
    mhl_compute(h->rs, min_depth, min_cpgs, (uint8_t)min_qual, h->has_target ? &h->target : nullptr, &h->site);
    return (int64_t)h->site.size();
}
int64_t orc_fdrp(orc_handle* h, int quantitative, uint32_t min_qual, uint32_t min_depth, uint32_t max_depth,
                 int32_t min_overlap, uint64_t seed) {
    h->oob = 0;
    fdrp_compute(h->rs, quantitative != 0, (uint8_t)min_qual, min_depth, max_depth, min_overlap,
                 h->has_target ? &h->target : nullptr, seed, &h->site, &h->oob);
    return (int64_t)h->site.size();
}
int orc_fdrp_oob(orc_handle* h) { return h->oob; }
void orc_site_rows(orc_handle* h, int32_t* tid, int32_t* pos, float* value) {
    for (size_t i = 0; i < h->site.size(); i++) {
        tid[i] = h->site[i].pos.tid; pos[i] = h->site[i].pos.pos; value[i] = h->site[i].value;
    }
}
int64_t orc_quartets(orc_handle* h, uint32_t min_depth, uint32_t min_qual) {
    quartet_compute(h->rs, min_depth, (uint8_t)min_qual, h->has_target ? &h->target : nullptr, &h->quart);
    return (int64_t)h->quart.size();
}
void orc_quartet_rows(orc_handle* h, int32_t* tid, int32_t* p1, int32_t* p2, int32_t* p3, int32_t* p4, float* pm,
                      float* me, uint32_t* counts16) {
    for (size_t i = 0; i < h->quart.size(); i++) {
        const QuartetRow& r = h->quart[i];
        tid[i] = r.q.p1.tid; p1[i] = r.q.p1.pos; p2[i] = r.q.p2.pos; p3[i] = r.q.p3.pos; p4[i] = r.q.p4.pos;
        pm[i] = r.pm; me[i] = r.me;
        if (counts16) memcpy(counts16 + 16 * i, r.counts, sizeof(r.counts));
    }
}
// out4 = {n_read, n_valid_read, n_conc, n_disc}; returns number of pair rows (0 unless want_pairs)
int64_t orc_lpmd(orc_handle* h, int32_t min_distance, int32_t max_distance, uint32_t min_qual, int want_pairs,
                 int32_t* out4, float* lpmd) {
    lpmd_compute(h->rs, min_distance, max_distance, (uint8_t)min_qual, h->has_target ? &h->target : nullptr,
                 want_pairs != 0, &h->lpmd);
    out4[0] = h->lpmd.n_read; out4[1] = h->lpmd.n_valid_read; out4[2] = h->lpmd.n_conc; out4[3] = h->lpmd.n_disc;
    *lpmd = h->lpmd.lpmd;
    return (int64_t)h->lpmd.pairs.size();
}
void orc_lpmd_pair_rows(orc_handle* h, int32_t* tid, int32_t* pos1, int32_t* pos2, float* lpmd, int32_t* nc, int32_t* nd) {
    for (size_t i = 0; i < h->lpmd.pairs.size(); i++) {
        const LpmdPairRow& r = h->lpmd.pairs[i];
        tid[i] = r.a.tid; pos1[i] = r.a.pos; pos2[i] = r.b.pos; lpmd[i] = r.lpmd; nc[i] = r.n_conc; nd[i] = r.n_disc;
    }
}
int orc_fmt_f32(float v, char* buf, int cap) {
    std::string s = fmt_f32(v);
    if ((int)s.size() + 1 > cap) return -1;
    memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
}
uint32_t orc_reservoir_draw(uint64_t seed, int32_t tid, int32_t pos, uint32_t total) {
    return reservoir_draw(seed, tid, pos, total);
}
float orc_compute_pm(const uint32_t* counts16) { return compute_pm(counts16); }
float orc_compute_me(const uint32_t* counts16) { return compute_me(counts16); }

}  // extern "C"

// =============================================================================================
// CLI: metheor_oracle <measure> -i in -o out [flags]   (same flags/defaults as lib.rs:24-231)
// =============================================================================================
#ifdef ORC_MAIN
static const char* arg(int argc, char** argv, const char* s, const char* l, const char* dflt) {
    for (int i = 2; i + 1 < argc; i++)
        if (!strcmp(argv[i], s) || !strcmp(argv[i], l)) return argv[i + 1];
    return dflt;
}
int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: metheor_oracle <pdr|pm|me|fdrp|qfdrp|mhl|lpmd> -i in -o out [flags]\n"); return 2; }
    std::string cmd = argv[1];
    const char* in = arg(argc, argv, "-i", "--input", nullptr);
    const char* outp = arg(argc, argv, "-o", "--output", nullptr);
    if (!in || !outp) { fprintf(stderr, "error: the following required arguments were not provided: --input --output\n"); return 2; }
    orc_handle* h = orc_open(in);
    if (!h->rs.error.empty()) { fprintf(stderr, "%s\n", h->rs.error.c_str()); return 101; }
    if (cmd != "lpmd" && !orc_all_xm_ok(h)) {
        fprintf(stderr, "Error reading XM tag in BAM record. Make sure the reads are aligned using Bismark!\n");
        return 101;
    }
    const char* cs = arg(argc, argv, "-c", "--cpg-set", nullptr);
    if (cs && orc_set_cpg_set_file(h, cs) != 0) { fprintf(stderr, "%s\n", h->rs.error.c_str()); return 101; }
    FILE* out = fopen(outp, "w");
    if (!out) { fprintf(stderr, "cannot open output\n"); return 101; }
    auto chrom = [&](int32_t tid) { return h->rs.header.names[(size_t)tid].c_str(); };
    uint32_t q = (uint32_t)atoi(arg(argc, argv, "-q", "--min-qual", "10"));
    uint32_t d = (uint32_t)atoi(arg(argc, argv, "-d", "--min-depth", "10"));
    if (cmd == "pdr") {
        uint32_t p = (uint32_t)atoi(arg(argc, argv, "-p", "--min-cpgs", "4"));
        orc_pdr(h, d, p, q);
        for (auto& r : h->pdr)
            fprintf(out, "%s\t%d\t%d\t%s\t%u\t%u\n", chrom(r.pos.tid), r.pos.pos, r.pos.pos + 2, fmt_f32(r.pdr).c_str(), r.n_conc, r.n_disc);
    } else if (cmd == "mhl") {
        uint32_t p = (uint32_t)atoi(arg(argc, argv, "-p", "--min-cpgs", "4"));
        orc_mhl(h, d, p, q);
        for (auto& r : h->site) fprintf(out, "%s\t%d\t%d\t%s\n", chrom(r.pos.tid), r.pos.pos, r.pos.pos + 2, fmt_f32(r.value).c_str());
    } else if (cmd == "fdrp" || cmd == "qfdrp") {
        uint32_t D = (uint32_t)atoi(arg(argc, argv, "-D", "--max-depth", "40"));
        int32_t l = atoi(arg(argc, argv, "-l", "--min-overlap", "35"));
        uint64_t seed = strtoull(arg(argc, argv, "--seed", "--seed", "0"), nullptr, 10);
        orc_fdrp(h, cmd == "qfdrp", q, d, D, l, seed);
        for (auto& r : h->site) fprintf(out, "%s\t%d\t%d\t%s\n", chrom(r.pos.tid), r.pos.pos, r.pos.pos + 2, fmt_f32(r.value).c_str());
    } else if (cmd == "pm" || cmd == "me") {
        orc_quartets(h, d, q);
        for (auto& r : h->quart)
            fprintf(out, "%s\t%d\t%d\t%d\t%d\t%s\n", chrom(r.q.p1.tid), r.q.p1.pos, r.q.p2.pos, r.q.p3.pos, r.q.p4.pos,
                    fmt_f32(cmd == "pm" ? r.pm : r.me).c_str());
    } else if (cmd == "lpmd") {
        int32_t m = atoi(arg(argc, argv, "-m", "--min-distance", "2"));
        int32_t M = atoi(arg(argc, argv, "-M", "--max-distance", "16"));
        const char* pairs = arg(argc, argv, "-p", "--pairs", nullptr);
        int32_t o4[4]; float v;
        orc_lpmd(h, m, M, q, pairs != nullptr, o4, &v);
        fprintf(out, "name\tlpmd\n%s\t%s\n", in, fmt_f32(v).c_str());
        if (pairs) {
            FILE* pf = fopen(pairs, "w");
            fprintf(pf, "chrom\tcpg1\tcpg2\tlpmd\tn_concordant\tn_discordant\n");
            for (auto& r : h->lpmd.pairs)
                fprintf(pf, "%s\t%d\t%d\t%s\t%d\t%d\n", chrom(r.a.tid), r.a.pos, r.b.pos, fmt_f32(r.lpmd).c_str(), r.n_conc, r.n_disc);
            fclose(pf);
        }
    } else {
        fprintf(stderr, "error: unrecognized subcommand '%s'\n", cmd.c_str());
        return 2;
    }
    fclose(out);
    orc_close(h);
    return 0;
}
#endif
