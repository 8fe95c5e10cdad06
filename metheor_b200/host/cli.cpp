// cli.cpp — the `metheor` command line: same subcommands, flags, defaults and value types as the reference's
// clap-derive definition (src/lib.rs:11-231), same exit statuses (2 for usage errors, 101 where the reference
// panics, 0 otherwise) and the stderr key phrases its CLI tests look for (tests/cli_error_handling.rs,
// tests/*-cli.rs).  Plus the C entry points of include/metheor_host.h.
#include <cerrno>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/metheor_b200.h"
#include "../../include/metheor_host.h"
#include "decode.hpp"
#include "inflate_fast.hpp"
#include "input.hpp"

namespace mthh {
void run(const mthh_options& o);
void run_tag(const char* input, const char* output, const char* genome, int device, int threads, const char* stats_json);
int format_f32(float v, char* buf, int cap);
}  // namespace mthh

using mthh::HostError;

namespace {

enum ArgKind { A_STR, A_U32, A_USIZE, A_U8, A_I32, A_U64 };
struct ArgSpec {
    char shortc;       // 0: long only
    const char* longn; // without "--"
    ArgKind kind;
    bool required;
    const char* dflt;  // printed in help, nullptr = none
    const char* help;
    const char* value_name;
};
struct CmdSpec {
    const char* name;
    const char* about;
    int measure;  // -1: tag
    std::vector<ArgSpec> args;
};

const ArgSpec IN1 = {'i', "input", A_STR, true, nullptr, "Input BAM file", "INPUT"};
const ArgSpec IN2 = {'i', "input", A_STR, true, nullptr, "Path to input BAM file", "INPUT"};
const ArgSpec CPGSET = {'c', "cpg-set", A_STR, false, nullptr, "(Optional) Specify a predefined set of CpGs (in BED file) to be analyzed", "CPG_SET"};
// engine options: long flags only, none collides with the reference's -i -o -d -p -q -c -D -l -m -M -g
const ArgSpec E_DEVICE = {0, "device", A_I32, false, "0", "[engine] first CUDA device to use", "DEVICE"};
const ArgSpec E_GPUS = {0, "gpus", A_I32, false, "1", "[engine] shard the genome over this many GPUs (position bins + halo reads)", "GPUS"};
const ArgSpec E_SHARD = {0, "shard", A_STR, false, "bins", "[engine] multi-GPU sharding: bins (equal position ranges) or contigs (whole contigs)", "SHARD"};
const ArgSpec E_DECODE = {0, "decode", A_STR, false, "gpu", "[engine] where BAM is inflated and decoded: gpu (one GPU, BAM input) or host", "DECODE"};
const ArgSpec E_FORMAT = {0, "format", A_STR, false, "tsv", "[engine] output format: tsv (the reference's), tsv.gz (BGZF), bedgraph, bedgraph.gz", "FORMAT"};
const ArgSpec E_REGION = {0, "region", A_STR, false, nullptr, "[engine] only the sites of chr[:beg-end] (1-based, inclusive); uses <input>.bai when present", "REGION"};
const ArgSpec E_THREADS = {0, "threads", A_I32, false, "0", "[engine] host decode threads (0 = all cores)", "THREADS"};
const ArgSpec E_SEED = {0, "seed", A_U64, false, "0", "[engine] reservoir-sampling seed once a pile exceeds --max-depth", "SEED"};
const ArgSpec E_STATS = {0, "stats", A_STR, false, nullptr, "[engine] write timing / throughput statistics as JSON", "STATS"};

std::vector<CmdSpec> commands() {
    auto out = [](const char* what) {
        const char* h = "Path to output table file";
        if (!strcmp(what, "PDR")) h = "Path to output table file summarizing the result of PDR calculation";
        else if (!strcmp(what, "PM")) h = "Path to output table file summarizing the result of PM calculation";
        else if (!strcmp(what, "FDRP")) h = "Path to output table file summarizing the result of FDRP calculation";
        else if (!strcmp(what, "MHL")) h = "Path to output table file summarizing the result of MHL calculation";
        else if (!strcmp(what, "LPMD")) h = "Path to output table file summarizing the result of LPMD calculation";
        return ArgSpec{'o', "output", A_STR, true, nullptr, h, "OUTPUT"};
    };
    const ArgSpec q = {'q', "min-qual", A_U8, false, "10", "Minimum quality for a read to be considered", "MIN_QUAL"};
    std::vector<CmdSpec> c;
    c.push_back({"pdr", "Compute proportion of discordant reads (PDR)", MTHH_PDR,
                 {IN1, out("PDR"), {'d', "min-depth", A_U32, false, "10", "Minimum depth of CpG stretches to consider", "MIN_DEPTH"},
                  {'p', "min-cpgs", A_USIZE, false, "4", "Minimum number of consecutive CpGs in a CpG stretch to consider", "MIN_CPGS"}, q, CPGSET}});
    c.push_back({"pm", "Compute epipolymorphism", MTHH_PM,
                 {IN1, out("PM"), {'d', "min-depth", A_U32, false, "10", "Minimum depth of CpG quartets to consider", "MIN_DEPTH"}, q, CPGSET}});
    c.push_back({"me", "Compute methylation entropy", MTHH_ME,
                 {IN1, out("PDR"), {'d', "min-depth", A_U32, false, "10", "Minimum depth of CpG quartets to consider", "MIN_DEPTH"}, q, CPGSET}});
    for (int k = 0; k < 2; k++)
        c.push_back({k ? "qfdrp" : "fdrp",
                     k ? "Compute quantitative fraction of discordant read pairs (qFDRP)" : "Compute fraction of discordant read pairs (FDRP)",
                     k ? MTHH_QFDRP : MTHH_FDRP,
                     {IN2, out("FDRP"), q,
                      {'d', "min-depth", A_USIZE, false, "10", "Minimum number of reads mapped to a CpG in order to be considered", "MIN_DEPTH"},
                      {'D', "max-depth", A_USIZE, false, "40", "Maximum number of reads to consider", "MAX_DEPTH"},
                      {'l', "min-overlap", A_I32, false, "35", "Minimum overlap between two reads to consider in bp", "MIN_OVERLAP"}, CPGSET}});
    c.push_back({"mhl", "Compute methylation haplotype load (MHL)", MTHH_MHL,
                 {IN1, out("MHL"), {'d', "min-depth", A_U32, false, "10", "Minimum depth of CpG stretches to consider", "MIN_DEPTH"},
                  {'p', "min-cpgs", A_USIZE, false, "4", "Minimum number of consecutive CpGs in a CpG stretch to consider", "MIN_CPGS"}, q, CPGSET}});
    c.push_back({"lpmd", "Compute local pairwise methylation discordance (LPMD)", MTHH_LPMD,
                 {IN2, out("LPMD"), {'p', "pairs", A_STR, false, nullptr, "(Optional) Concordance information for all CpG pairs", "PAIRS"},
                  {'m', "min-distance", A_I32, false, "2", "Minimum distance between CpG pairs to consider", "MIN_DISTANCE"},
                  {'M', "max-distance", A_I32, false, "16", "Maximum distance between CpG pairs to consider", "MAX_DISTANCE"}, q, CPGSET}});
    c.push_back({"tag", "Add bismark XM tag to BAM file", -1,
                 {{'i', "input", A_STR, true, nullptr, "", "INPUT"}, {'o', "output", A_STR, true, nullptr, "", "OUTPUT"},
                  {'g', "genome", A_STR, true, nullptr, "", "GENOME"}, E_DEVICE, E_THREADS, E_STATS}});
    for (auto& cmd : c)
        if (cmd.measure >= 0) {
            cmd.args.push_back(E_DEVICE); cmd.args.push_back(E_GPUS); cmd.args.push_back(E_SHARD); cmd.args.push_back(E_DECODE); cmd.args.push_back(E_FORMAT); cmd.args.push_back(E_REGION); cmd.args.push_back(E_THREADS); cmd.args.push_back(E_STATS);
            if (cmd.measure == MTHH_FDRP || cmd.measure == MTHH_QFDRP) cmd.args.push_back(E_SEED);
        }
    return c;
}

const char* VERSION_LINE = "metheor 0.1.9";  // lib.rs:14; the engine build is reported by --version as a second line

void print_main_help(FILE* f) {
    fprintf(f, "Summarizes the heterogeneity of DNA methylation states using BAM files.\n\nUsage: metheor <COMMAND>\n\nCommands:\n");
    for (auto& c : commands()) fprintf(f, "  %-6s %s\n", c.name, c.about);
    fprintf(f, "  %-6s %s\n\nOptions:\n  -h, --help     Print help\n  -V, --version  Print version\n", "help",
            "Print this message or the help of the given subcommand(s)");
}

std::string usage_line(const CmdSpec& c) {
    std::string u = std::string("Usage: metheor ") + c.name + " [OPTIONS]";
    for (auto& a : c.args)
        if (a.required) u += std::string(" --") + a.longn + " <" + a.value_name + ">";
    return u;
}

void print_cmd_help(FILE* f, const CmdSpec& c) {
    fprintf(f, "%s\n\n%s\n\nOptions:\n", c.about, usage_line(c).c_str());
    for (auto& a : c.args) {
        std::string left = a.shortc ? std::string("  -") + a.shortc + ", --" + a.longn : std::string("      --") + a.longn;
        left += std::string(" <") + a.value_name + ">";
        fprintf(f, "%-34s %s", left.c_str(), a.help);
        if (a.dflt) fprintf(f, " [default: %s]", a.dflt);
        fprintf(f, "\n");
    }
    fprintf(f, "  -h, --help                       Print help\n");
}

bool parse_int(const char* s, long long lo, unsigned long long hi, bool is_signed, unsigned long long* out_u, long long* out_s,
               std::string* why) {
    if (!*s) { *why = "cannot parse integer from empty string"; return false; }
    errno = 0;
    char* e = nullptr;
    if (is_signed) {
        long long v = strtoll(s, &e, 10);
        if (*e) { *why = "invalid digit found in string"; return false; }
        if (errno || v < lo || v > (long long)hi) { *why = v < 0 ? "number too small to fit in target type" : "number too large to fit in target type"; return false; }
        *out_s = v;
        return true;
    }
    if (s[0] == '-') { *why = "invalid digit found in string"; return false; }
    unsigned long long v = strtoull(s[0] == '+' ? s + 1 : s, &e, 10);
    if (*e) { *why = "invalid digit found in string"; return false; }
    if (errno || v > hi) { *why = "number too large to fit in target type"; return false; }
    *out_u = v;
    return true;
}

int usage_error(const CmdSpec* c, const std::string& msg) {
    fprintf(stderr, "error: %s\n\n%s\n\nFor more information, try '--help'.\n", msg.c_str(),
            c ? usage_line(*c).c_str() : "Usage: metheor <COMMAND>");
    return 2;
}

}  // namespace

extern "C" {

void mthh_options_default(mthh_options* o, int32_t measure) {
    memset(o, 0, sizeof(*o));
    o->measure = measure;
    o->min_depth = 10;   // lib.rs:37,66,86,113,183
    o->min_cpgs = 4;     // lib.rs:42,188
    o->min_qual = 10;
    o->max_depth = 40;   // lib.rs:117
    o->min_overlap = 35; // lib.rs:121
    o->min_distance = 2; // lib.rs:215
    o->max_distance = 16;
    o->n_gpus = 1;
    o->shard_contigs = 0;
    o->decode_host = 0;
    o->out_format = 0;
    o->region = nullptr;
}

int mthh_run(const mthh_options* o, char* err, size_t errcap) {
    try {
        mthh::run(*o);
        return 0;
    } catch (const HostError& e) {
        if (err && errcap) snprintf(err, errcap, "%s", e.msg.c_str());
        return e.status ? e.status : 1;
    } catch (const std::exception& e) {
        if (err && errcap) snprintf(err, errcap, "%s", e.what());
        return 1;
    }
}

int mthh_tag(const char* input, const char* output, const char* genome, int32_t device, int32_t threads, const char* stats_json,
             char* err, size_t errcap) {
    try {
        if (!input || !output || !genome) throw HostError{2, "mthh_tag: input, output and genome are required"};
        mthh::run_tag(input, output, genome, device, threads, stats_json);
        return 0;
    } catch (const HostError& e) {
        if (err && errcap) snprintf(err, errcap, "%s", e.msg.c_str());
        return e.status ? e.status : 1;
    } catch (const std::exception& e) {
        if (err && errcap) snprintf(err, errcap, "%s", e.what());
        return 1;
    }
}

int mthh_format_f32(float v, char* buf, int cap) { return mthh::format_f32(v, buf, cap); }

int mthh_plan_shards(int32_t n_ref, const int64_t* ref_len, int32_t world, int32_t by_contig, mthh_interval* out, int32_t cap) {
    if (n_ref < 0 || world < 1 || (n_ref && !ref_len)) return -1;
    std::vector<int64_t> rl(ref_len, ref_len + n_ref);
    auto plan = by_contig ? mthh::plan_contigs(rl, world) : mthh::plan_bins(rl, world, mthh::shard_region_cost_env());
    int32_t n = 0;
    for (int r = 0; r < world; r++)
        for (const auto& iv : plan[(size_t)r]) {
            if (out && n < cap) out[n] = mthh_interval{r, iv.tid, iv.lo, iv.hi};
            n++;
        }
    return n;
}

int mthh_inflate_raw(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
    return mthh::inflate_fast(in, in_len, out, out_len) ? 1 : 0;
}
uint32_t mthh_crc32(const uint8_t* p, size_t n) { return mthh::crc32_fast(p, n); }
int64_t mthh_zlib_fallbacks(void) { return mthh::g_zlib_fallbacks.load(); }

int mthh_main(int argc, char** argv) {
    const std::vector<CmdSpec> cmds = commands();
    if (argc < 2) {  // lib.rs:18 arg_required_else_help: usage on stderr, status 2
        print_main_help(stderr);
        return 2;
    }
    std::string a1 = argv[1];
    if (a1 == "-h" || a1 == "--help" || (a1 == "help" && argc == 2)) { print_main_help(stdout); return 0; }
    if (a1 == "-V" || a1 == "--version") {
        printf("%s\n", VERSION_LINE);
        fprintf(stderr, "engine: %s\n", mth_version());
        return 0;
    }
    if (a1 == "help") a1 = argv[2];
    const CmdSpec* cmd = nullptr;
    for (auto& c : cmds)
        if (a1 == c.name) cmd = &c;
    if (!cmd) {
        if (a1.size() && a1[0] == '-') return usage_error(nullptr, "unexpected argument '" + a1 + "' found");
        return usage_error(nullptr, "unrecognized subcommand '" + a1 + "'");
    }
    if (std::string(argv[1]) == "help") { print_cmd_help(stdout, *cmd); return 0; }
    if (argc == 2) {  // subcommands also carry arg_required_else_help
        print_cmd_help(stderr, *cmd);
        return 2;
    }
    std::vector<const char*> val(cmd->args.size(), nullptr);
    for (int i = 2; i < argc; i++) {
        std::string a = argv[i];
        if (a == "-h" || a == "--help") { print_cmd_help(stdout, *cmd); return 0; }
        int which = -1;
        const char* inline_val = nullptr;
        if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
            size_t eq = a.find('=');
            std::string name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
            for (size_t k = 0; k < cmd->args.size(); k++)
                if (name == cmd->args[k].longn) which = (int)k;
            if (eq != std::string::npos) inline_val = argv[i] + eq + 1;
        } else if (a.size() >= 2 && a[0] == '-') {
            for (size_t k = 0; k < cmd->args.size(); k++)
                if (cmd->args[k].shortc && a[1] == cmd->args[k].shortc) which = (int)k;
            if (which >= 0 && a.size() > 2) inline_val = argv[i] + (a[2] == '=' ? 3 : 2);
        }
        if (which < 0) return usage_error(cmd, "unexpected argument '" + a + "' found");
        const ArgSpec& sp = cmd->args[(size_t)which];
        if (val[(size_t)which]) return usage_error(cmd, std::string("the argument '--") + sp.longn + " <" + sp.value_name + ">' cannot be used multiple times");
        if (inline_val) {
            val[(size_t)which] = inline_val;
        } else {
            if (i + 1 >= argc)
                return usage_error(cmd, std::string("a value is required for '--") + sp.longn + " <" + sp.value_name + ">' but none was supplied");
            val[(size_t)which] = argv[++i];
        }
    }
    std::string missing;
    for (size_t k = 0; k < cmd->args.size(); k++)
        if (cmd->args[k].required && !val[k]) missing += std::string("\n  --") + cmd->args[k].longn + " <" + cmd->args[k].value_name + ">";
    if (!missing.empty()) return usage_error(cmd, "the following required arguments were not provided:" + missing);

    mthh_options o;
    mthh_options_default(&o, cmd->measure < 0 ? 0 : cmd->measure);
    for (size_t k = 0; k < cmd->args.size(); k++) {
        const ArgSpec& sp = cmd->args[k];
        const char* v = val[k];
        if (!v) continue;
        unsigned long long u = 0;
        long long s = 0;
        std::string why;
        bool ok = true;
        switch (sp.kind) {
            case A_STR: break;
            case A_U8: ok = parse_int(v, 0, 255, false, &u, &s, &why); if (ok) why.clear(); if (!ok && why.find("large") != std::string::npos) why = std::string(v) + " is not in 0..=255"; break;
            case A_U32: ok = parse_int(v, 0, UINT32_MAX, false, &u, &s, &why); break;
            case A_USIZE: ok = parse_int(v, 0, ULLONG_MAX, false, &u, &s, &why); break;
            case A_U64: ok = parse_int(v, 0, ULLONG_MAX, false, &u, &s, &why); break;
            case A_I32: ok = parse_int(v, INT32_MIN, INT32_MAX, true, &u, &s, &why); break;
        }
        if (!ok) return usage_error(cmd, std::string("invalid value '") + v + "' for '--" + sp.longn + " <" + sp.value_name + ">': " + why);
        const std::string n = sp.longn;
        auto clamp32 = [](unsigned long long x) { return (uint32_t)(x > UINT32_MAX ? UINT32_MAX : x); };
        if (n == "input") o.input = v;
        else if (n == "output") o.output = v;
        else if (n == "cpg-set") o.cpg_set = v;
        else if (n == "pairs") o.pairs = v;
        else if (n == "min-depth") o.min_depth = clamp32(u);
        else if (n == "min-cpgs") o.min_cpgs = clamp32(u);
        else if (n == "min-qual") o.min_qual = (uint32_t)u;
        else if (n == "max-depth") o.max_depth = clamp32(u);
        else if (n == "min-overlap") o.min_overlap = (int32_t)s;
        else if (n == "min-distance") o.min_distance = (int32_t)s;
        else if (n == "max-distance") o.max_distance = (int32_t)s;
        else if (n == "device") o.device = (int32_t)s;
        else if (n == "gpus") o.n_gpus = (int32_t)s;
        else if (n == "shard") {
            if (!strcmp(v, "contigs")) o.shard_contigs = 1;
            else if (!strcmp(v, "bins")) o.shard_contigs = 0;
            else return usage_error(cmd, std::string("invalid value '") + v + "' for '--shard <SHARD>': expected bins or contigs");
        }
        else if (n == "decode") {
            if (!strcmp(v, "host")) o.decode_host = 1;
            else if (!strcmp(v, "gpu")) o.decode_host = 0;
            else return usage_error(cmd, std::string("invalid value '") + v + "' for '--decode <DECODE>': expected gpu or host");
        }
        else if (n == "format") {
            if (!strcmp(v, "tsv")) o.out_format = 0;
            else if (!strcmp(v, "tsv.gz")) o.out_format = 1;
            else if (!strcmp(v, "bedgraph")) o.out_format = 2;
            else if (!strcmp(v, "bedgraph.gz")) o.out_format = 3;
            else return usage_error(cmd, std::string("invalid value '") + v + "' for '--format <FORMAT>': expected tsv, tsv.gz, bedgraph or bedgraph.gz");
        }
        else if (n == "region") o.region = v;
        else if (n == "threads") o.threads = (int32_t)s;
        else if (n == "seed") o.seed = u;
        else if (n == "stats") o.stats_json = v;
    }
    char err[4096];
    err[0] = 0;
    if (cmd->measure < 0) {  // tag: -i -o -g (src/lib.rs:219-230)
        const char* genome = nullptr;
        for (size_t k = 0; k < cmd->args.size(); k++)
            if (!strcmp(cmd->args[k].longn, "genome")) genome = val[k];
        int rc = mthh_tag(o.input, o.output, genome, o.device, o.threads, o.stats_json, err, sizeof(err));
        if (rc != 0) fprintf(stderr, "%s\n", err);
        return rc;
    }
    int rc = mthh_run(&o, err, sizeof(err));
    if (rc != 0) fprintf(stderr, "%s\n", err);
    return rc;
}

}  // extern "C"

// ---- decode-only API (no GPU): the records of a file as BismarkRead-level SoA ------------------------------------
namespace {
struct DecodedOwner {
    mthh_decoded view;
    mthh::SoaChunk soa;
    std::vector<int64_t> off;
    std::vector<std::string> names;
    std::vector<const char*> name_ptrs;
    std::vector<int64_t> lens;
};
}  // namespace

extern "C" {

int mthh_decode_file(const char* path, const char* cpg_set, int32_t threads, mthh_decoded** out, char* err, size_t errcap) {
    try {
        int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
        mthh::ThreadPool pool(nt < 1 ? 1 : nt);
        mthh::RecordStream in(path, nt < 1 ? 1 : nt, 8u << 20);
        mthh::CpgSet set;
        if (cpg_set) set.load(cpg_set, in.header());
        mthh::DecodeOptions opt;
        opt.cpg_set = cpg_set ? &set : nullptr;
        opt.keep_empty = true;
        auto* d = new DecodedOwner();
        std::vector<mthh::RecordRef> recs;
        std::vector<mthh::SoaChunk> chunks;
        std::vector<mthh::DecodeCounters> cnt;
        const size_t TASK = 2048;
        try {
            while (in.next(&recs)) {
                size_t nt2 = (recs.size() + TASK - 1) / TASK;
                if (chunks.size() < nt2) chunks.resize(nt2);
                cnt.assign(nt2, mthh::DecodeCounters());
                std::string first;
                int status = 0;
                std::mutex m;
                pool.run((int64_t)nt2, [&](int64_t k, int) {
                    chunks[(size_t)k].clear();
                    try {
                        mthh::decode_records(in.format(), in.header(), recs.data(), (size_t)k * TASK, std::min(recs.size(), ((size_t)k + 1) * TASK),
                                             opt, &chunks[(size_t)k], &cnt[(size_t)k]);
                    } catch (const HostError& e) {
                        std::lock_guard<std::mutex> g(m);
                        if (!status) { status = e.status; first = e.msg; }
                    }
                });
                if (status) throw HostError{status, first};
                for (size_t k = 0; k < nt2; k++) {
                    auto& c = chunks[k];
                    auto app = [](auto& dst, const auto& src) { dst.insert(dst.end(), src.begin(), src.end()); };
                    app(d->soa.tid, c.tid); app(d->soa.start, c.start); app(d->soa.end, c.end); app(d->soa.meta, c.meta);
                    app(d->soa.n_cpg, c.n_cpg); app(d->soa.cpg_pos, c.cpg_pos); app(d->soa.cpg_rel, c.cpg_rel); app(d->soa.cpg_meth, c.cpg_meth);
                }
            }
        } catch (...) {
            delete d;
            throw;
        }
        d->off.assign(d->soa.n_cpg.size() + 1, 0);
        for (size_t i = 0; i < d->soa.n_cpg.size(); i++) d->off[i + 1] = d->off[i] + d->soa.n_cpg[i];
        d->names = in.header().names;
        d->lens = in.header().lengths;
        for (auto& s : d->names) d->name_ptrs.push_back(s.c_str());
        mthh_decoded& v = d->view;
        v.n_reads = (int64_t)d->soa.start.size();
        v.n_cpg = (int64_t)d->soa.cpg_pos.size();
        v.n_ref = (int32_t)d->names.size();
        v.ref_name = d->name_ptrs.data();
        v.ref_len = d->lens.data();
        v.tid = d->soa.tid.data(); v.start = d->soa.start.data(); v.end = d->soa.end.data(); v.meta = d->soa.meta.data();
        v.cpg_off = d->off.data(); v.cpg_pos = d->soa.cpg_pos.data(); v.cpg_rel = d->soa.cpg_rel.data(); v.cpg_meth = d->soa.cpg_meth.data();
        *out = &d->view;
        return 0;
    } catch (const HostError& e) {
        if (err && errcap) snprintf(err, errcap, "%s", e.msg.c_str());
        return e.status ? e.status : 1;
    } catch (const std::exception& e) {
        if (err && errcap) snprintf(err, errcap, "%s", e.what());
        return 1;
    }
}

void mthh_decoded_free(mthh_decoded* d) {
    if (d) delete reinterpret_cast<DecodedOwner*>(d);  // `view` is the first member
}

}  // extern "C"
