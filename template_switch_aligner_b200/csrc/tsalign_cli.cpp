// tsalign_cli.cpp -- `tsalign-b200 align ...`: drop-in for the reference's `tsalign align` (default method
// a-star-template-switch) on top of the C ABI.  Mirrors tsalign/src/align.rs:57-432 (flags, input handling order:
// parse FASTA -> drop skip characters -> upper-case -> ranges -> alphabet check), align/fasta_parser.rs:24-174,
// util.rs:14-28 (config directory), the stdout block of alignment_result.rs:736-777 and the TOML layout of
// alignment_result.rs:32-81 as written by align/template_switch_distance_type_selectors.rs:442-449.
//
// Post-processing (extension beyond the range, equal-cost ranges: alignment_result.rs:247-573) runs inside the library call.
// Batch front-end (not in the reference, whose fasta_parser.rs:157-173 accepts exactly two records): `--pairs FILE` takes a
// multi-FASTA with an even number of records (records 2k, 2k+1 = reference, query of pair k) or a TSV (name, reference, query),
// aligns all pairs with ONE tsa_align_batch call and writes one TOML per pair (`-o DIR`) and / or a JSON-lines stream.
// Not part of this build: `show`, `preprocess` and the other alignment methods.
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "tsalign_b200.h"

namespace {

struct Record { std::string id, comment, seq; };

[[noreturn]] void die(const std::string& msg, int code = 1) {
    fprintf(stderr, "Error: %s\n", msg.c_str());
    exit(code);
}

// align/fasta_parser.rs:24-174
std::vector<Record> parse_fasta(const std::string& path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) die("Unable to open input file \"" + path + "\"");
    std::stringstream ss; ss << in.rdbuf();
    const std::string text = ss.str();
    std::vector<Record> recs;
    size_t pos = 0;
    bool newline = true;
    for (; pos < text.size(); pos++) {   // FileStart
        char c = text[pos];
        if (c == '\n' || c == '\r') newline = true;
        else if (c == '>') { if (!newline) die("First fasta record is not preceded by a newline character"); break; }
        else { newline = false; if (!isspace((unsigned char)c)) die(std::string("Found non-whitespace character before first fasta record: ") + c); }
    }
    if (pos >= text.size()) die("Input file \"" + path + "\" contains no fasta record");
    while (pos < text.size()) {
        Record r;
        pos++;  // '>'
        while (pos < text.size() && text[pos] != '\n' && text[pos] != '\r' && !isspace((unsigned char)text[pos])) r.id.push_back(text[pos++]);
        if (pos < text.size() && text[pos] != '\n' && text[pos] != '\r') {
            pos++;  // the whitespace that ends the id
            while (pos < text.size() && text[pos] != '\n' && text[pos] != '\r') r.comment.push_back(text[pos++]);
        }
        newline = false;
        for (; pos < text.size(); pos++) {
            char c = text[pos];
            if (c == '\n' || c == '\r') newline = true;
            else if (c == '>' && newline) break;
            else { r.seq.push_back(c); newline = false; }
        }
        recs.push_back(r);
    }
    return recs;
}

std::string fmt_f64(double v) {  // Rust `{}` for f64: shortest round-trip, always with a fractional part
    if (std::isnan(v)) return "nan";
    if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::fixed);
    std::string s(buf, res.ptr);
    // fixed shortest can be long for tiny values; fall back to general when it helps
    auto res2 = std::to_chars(buf, buf + sizeof(buf), v);
    std::string g(buf, res2.ptr);
    if (g.find('e') == std::string::npos) s = g;
    if (s.find('.') == std::string::npos) s += ".0";
    return s;
}

std::string toml_string(const std::string& s) {
    std::string out = "\"";
    for (char c : s) {
        if (c == '"' || c == '\\') { out.push_back('\\'); out.push_back(c); }
        else if (c == '\n') out += "\\n";
        else if (c == '\t') out += "\\t";
        else out.push_back(c);
    }
    return out + "\"";
}

const char* const OP_NAMES[] = {"PrimaryInsertion", "PrimaryDeletion", "PrimarySubstitution", "PrimaryMatch", "PrimaryFlankInsertion", "PrimaryFlankDeletion",
                                "PrimaryFlankSubstitution", "PrimaryFlankMatch", "SecondaryInsertion", "SecondaryDeletion", "SecondarySubstitution", "SecondaryMatch"};

std::string cigar(const tsa_result& r) {  // alignment.rs:95-110, template_switch_distance/display.rs:8-41
    std::string out;
    for (size_t i = 0; i < r.n_ops; i++) {
        const tsa_op& op = r.ops[i];
        if (op.type == TSA_OP_TS_ENTRANCE) {
            // EqualCostRange display (template_switch_distance/display.rs:80-94)
            const bool valid = op.min_start <= op.max_start && op.min_end <= op.max_end;
            const std::string rng = valid ? "[" + std::to_string(op.min_start) + "," + std::to_string(op.max_start) + "]:[" + std::to_string(op.min_end) + "," + std::to_string(op.max_end) + "]" : "[-]:[-]";
            out += std::string("[TS") + "RQ"[op.primary] + "RQ"[op.secondary] + "FR"[op.direction] + ":" + rng + ":" + std::to_string(op.value) + ":";
        } else if (op.type == TSA_OP_TS_EXIT) {
            out += ":" + std::to_string(op.value) + "]";
        } else {
            out += std::to_string(op.count) + "IDX="[op.type & 3];
        }
    }
    return out;
}

std::string complement_text(const std::string& s, bool rna) {
    std::string out(s.rbegin(), s.rend());
    for (char& c : out) {
        switch (c) {
        case 'A': c = rna ? 'U' : 'T'; break; case 'T': case 'U': c = 'A'; break; case 'C': c = 'G'; break; case 'G': c = 'C'; break;
        case 'R': c = 'Y'; break; case 'Y': c = 'R'; break; case 'K': c = 'M'; break; case 'M': c = 'K'; break;
        case 'B': c = 'V'; break; case 'V': c = 'B'; break; case 'D': c = 'H'; break; case 'H': c = 'D'; break;
        default: break;
        }
    }
    return out;
}

struct Cli {
    std::string pair_fasta, reference, query, output, config_dir = "sample_tsa_config", alphabet = "dna-n", skip, method = "a-star-template-switch";
    std::string rq_ranges, pairs_file, jsonl;
    int total_length_strategy = 0, descendant_strategy = 0;
    bool no_ts = false, embedded = false, dont_extend = false;
    long long ref_off = -1, ref_lim = -1, qry_off = -1, qry_lim = -1;
    unsigned long long cost_limit = UINT64_MAX, memory_limit = UINT64_MAX;
    int device = 0;
};

void usage() {
    fprintf(stderr,
            "Usage: tsalign-b200 align [-p PAIR_FASTA | -r REFERENCE -q QUERY] [-o OUTPUT] [-a ALPHABET] [--skip-characters S]\n"
            "       [-c CONFIGURATION_DIRECTORY] [--no-ts] [--cost-limit N] [--memory-limit BYTES] [--rq-ranges R<a>..<b>Q<c>..<d>]\n"
            "       [--reference-offset N] [--reference-limit N] [--query-offset N] [--query-limit N] [--use-embedded-rq-ranges]\n"
            "       [--dont-extend-beyond-range] [--device N] (search-heuristic flags of tsalign are accepted and ignored)\n"
            "   or: tsalign-b200 align --pairs MULTI_FASTA_OR_TSV [-o OUTPUT_DIRECTORY] [--output-jsonl FILE] [same options]   (batch: one GPU call)\n");
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 2 || std::string(argv[1]) == "-h" || std::string(argv[1]) == "--help") { usage(); return argc < 2 ? 2 : 0; }
    const std::string sub = argv[1];
    if (sub == "show" || sub == "preprocess") die("subcommand '" + sub + "' is not part of tsalign-b200 (alignment path only)", 2);
    if (sub != "align") { usage(); return 2; }
    Cli cli;
    // flags with a value that only steer the reference's search heuristics (never the optimal cost): accepted, ignored
    const char* ignored_with_value[] = {"-l", "--log-level", "--cache-directory", "-k", "--ts-node-ord-strategy", "--ts-min-length-strategy", "--ts-chaining-strategy",
                                        "--max-chaining-successors", "--max-exact-cost-function-cost", "--chaining-closed-list", "--chaining-open-list"};
    for (int i = 2; i < argc; i++) {
        std::string a = argv[i], val;
        bool has_inline = false;
        size_t eq = a.find('=');
        if (a.rfind("--", 0) == 0 && eq != std::string::npos) { val = a.substr(eq + 1); a = a.substr(0, eq); has_inline = true; }
        auto value = [&]() -> std::string {
            if (has_inline) return val;
            if (i + 1 >= argc) die("a value is required for '" + a + "'", 2);
            return argv[++i];
        };
        auto number = [&]() -> unsigned long long {
            std::string v = value();
            char* end = nullptr;
            unsigned long long x = strtoull(v.c_str(), &end, 10);
            if (v.empty() || *end) die("invalid value '" + v + "' for '" + a + "'", 2);
            return x;
        };
        bool ignored = false;
        for (const char* f : ignored_with_value) if (a == f) { value(); ignored = true; }
        if (ignored) continue;
        if (a == "-p" || a == "--pair-fasta") cli.pair_fasta = value();
        else if (a == "-r" || a == "--reference") cli.reference = value();
        else if (a == "-q" || a == "--query") cli.query = value();
        else if (a == "-o" || a == "--output") cli.output = value();
        else if (a == "-a" || a == "--alphabet") cli.alphabet = value();
        else if (a == "--skip-characters") cli.skip = value();
        else if (a == "-c" || a == "--configuration-directory") cli.config_dir = value();
        else if (a == "--alignment-method") cli.method = value();
        else if (a == "--ts-descendant-strategy") { const std::string v = value(); if (v == "allow-any") cli.descendant_strategy = 0; else if (v == "allow-only-all-equal") cli.descendant_strategy = 1; else die("invalid value '" + v + "' for '--ts-descendant-strategy'", 2); }
        else if (a == "--ts-total-length-strategy") { const std::string v = value(); if (v == "maximise") cli.total_length_strategy = 0; else if (v == "none") cli.total_length_strategy = 1; else die("invalid value '" + v + "' for '--ts-total-length-strategy'", 2); }
        else if (a == "--pairs") cli.pairs_file = value();
        else if (a == "--output-jsonl") cli.jsonl = value();
        else if (a == "--no-ts") cli.no_ts = true;
        else if (a == "--force-no-preprocessing" || a == "--force-label-correcting") {}
        else if (a == "--cost-limit") cli.cost_limit = number();
        else if (a == "--memory-limit") cli.memory_limit = number();
        else if (a == "--reference-offset") cli.ref_off = (long long)number();
        else if (a == "--reference-limit") cli.ref_lim = (long long)number();
        else if (a == "--query-offset") cli.qry_off = (long long)number();
        else if (a == "--query-limit") cli.qry_lim = (long long)number();
        else if (a == "--rq-ranges") cli.rq_ranges = value();
        else if (a == "--use-embedded-rq-ranges") cli.embedded = true;
        else if (a == "--dont-extend-beyond-range") cli.dont_extend = true;
        else if (a == "--device") cli.device = (int)number();
        else die("unexpected argument '" + a + "'", 2);
    }
    if (cli.method != "a-star-template-switch") die("--alignment-method " + cli.method + " is not part of tsalign-b200 (only a-star-template-switch)", 2);
    static const char* const ALPHABETS[] = {"dna", "dna-n", "rna", "rna-n", "dna-iupac", "rna-iupac"};
    int alphabet = -1;
    for (int k = 0; k < 6; k++) if (cli.alphabet == ALPHABETS[k]) alphabet = k;
    if (alphabet < 0) die("invalid value '" + cli.alphabet + "' for '--alphabet'", 2);

    if (cli.embedded && cli.skip.find('|') != std::string::npos) die("Using embedded RQ ranges, but '|' is part of the skip characters");
    // drop skip characters, upper-case (align.rs:304-335), then the ranges (align.rs:338-379, 516-599)
    auto prepare = [&](Record& ref, Record& qry, long long& ro, long long& rl, long long& qo, long long& ql) {
    for (Record* r : {&ref, &qry}) {
        std::string s;
        for (char c : r->seq) if (cli.skip.find(c) == std::string::npos) s.push_back((char)toupper((unsigned char)c));
        r->seq = s;
    }
    ro = 0; rl = -1; qo = 0; ql = -1;
    if (cli.embedded) {
        if (!cli.rq_ranges.empty() || cli.ref_off >= 0 || cli.ref_lim >= 0 || cli.qry_off >= 0 || cli.qry_lim >= 0) die("Redundant specification of RQ ranges");
        auto split = [&](Record& r, const char* what, long long& off, long long& lim) {
            size_t a = r.seq.find('|');
            if (a == std::string::npos) die(std::string("Using embedded RQ ranges, but ") + what + " sequence contains no '|' character.");
            size_t b = r.seq.find('|', a + 1);
            if (b == std::string::npos) die(std::string("Using embedded RQ ranges, but ") + what + " sequence contains only one '|' character.");
            if (r.seq.find('|', b + 1) != std::string::npos) die(std::string("Using embedded RQ ranges, but ") + what + " sequence contains more than two '|' characters");
            off = (long long)a; lim = (long long)b - 1;
            std::string s;
            for (char c : r.seq) if (c != '|') s.push_back(c);
            r.seq = s;
        };
        split(ref, "reference", ro, rl);
        split(qry, "query", qo, ql);
    } else {
        long long rr0 = 0, rr1 = (long long)ref.seq.size(), qq0 = 0, qq1 = (long long)qry.seq.size();
        if (!cli.rq_ranges.empty()) {
            const std::string& s = cli.rq_ranges;
            size_t p = 0;
            bool have_r = false, have_q = false;
            while (p < s.size()) {
                char which = s[p++];
                while (p < s.size() && isspace((unsigned char)s[p])) p++;
                size_t d0 = p; while (p < s.size() && isdigit((unsigned char)s[p])) p++;
                std::string off = s.substr(d0, p - d0);
                if (s.compare(p, 2, "..") != 0) die("malformed --rq-ranges '" + s + "'", 2);
                p += 2;
                d0 = p; while (p < s.size() && isdigit((unsigned char)s[p])) p++;
                std::string lim = s.substr(d0, p - d0);
                if (off.empty() || lim.empty()) die("malformed --rq-ranges '" + s + "'", 2);
                while (p < s.size() && isspace((unsigned char)s[p])) p++;
                if (which == 'R' && !have_r) { rr0 = atoll(off.c_str()); rr1 = atoll(lim.c_str()); have_r = true; }
                else if (which == 'Q' && !have_q) { qq0 = atoll(off.c_str()); qq1 = atoll(lim.c_str()); have_q = true; }
                else die("malformed --rq-ranges '" + s + "'", 2);
            }
            if ((have_r && (cli.ref_off >= 0 || cli.ref_lim >= 0)) || (have_q && (cli.qry_off >= 0 || cli.qry_lim >= 0))) die("Redundant specification of RQ ranges", 2);
        }
        ro = cli.ref_off >= 0 ? cli.ref_off : rr0; rl = cli.ref_lim >= 0 ? cli.ref_lim : rr1;
        qo = cli.qry_off >= 0 ? cli.qry_off : qq0; ql = cli.qry_lim >= 0 ? cli.qry_lim : qq1;
    }

    };
    // ---- cost model (util.rs:14-28) ----
    std::string cfg_path = cli.config_dir + "/config.tsa";
    std::ifstream cin_(cfg_path);
    if (!cin_) die("Unable to open config file \"" + cfg_path + "\"");
    std::stringstream cs; cs << cin_.rdbuf();
    const std::string cfg_text = cs.str();
    int status = 0;
    char err[512] = {0};
    tsa_config* cfg = tsa_config_parse(cfg_text.data(), cfg_text.size(), alphabet, &status, err, sizeof(err));
    if (!cfg) die(std::string("cannot parse ") + cfg_path + ": " + err);

    tsa_options opt;
    memset(&opt, 0, sizeof(opt));
    opt.no_ts = cli.no_ts; opt.device = cli.device; opt.cost_limit = cli.cost_limit; opt.memory_limit = cli.memory_limit;
    // a_star_aligner.rs:238-253: extension unless --dont-extend-beyond-range, equal-cost ranges always
    opt.postprocess = TSA_POST_EQUAL_COST_RANGES | (cli.dont_extend ? 0 : TSA_POST_EXTEND_BEYOND_RANGE);
    opt.total_length_strategy = cli.total_length_strategy; opt.descendant_strategy = cli.descendant_strategy;
    const bool rna = alphabet == TSA_ALPHABET_RNA || alphabet == TSA_ALPHABET_RNA_N || alphabet == TSA_ALPHABET_RNA_IUPAC;
    auto result_line_of = [](const tsa_result& res) -> std::string {
        switch (res.result_type) {
        case TSA_FOUND_TARGET: return "Reached target with cost " + std::to_string(res.cost);
        case TSA_EXCEEDED_COST_LIMIT: return "Exceeded cost limit of " + std::to_string(res.cost);
        case TSA_EXCEEDED_MEMORY_LIMIT: return "Exceeded memory limit, but reached a maximum cost of " + std::to_string(res.cost);
        default: return "Found no target";
        }
    };
    auto per_base_of = [](const tsa_result& res, const Record& ref, const Record& qry) {
        return (ref.seq.size() + qry.seq.size()) ? 2.0 * (double)res.cost / (double)(ref.seq.size() + qry.seq.size()) : 0.0;
    };
    // TOML result file (alignment_result.rs:32-81 as written by align/template_switch_distance_type_selectors.rs:442-449)
    auto write_toml = [&](const std::string& path, const tsa_result& res, const Record& ref, const Record& qry, long long ro, long long qo) {
        const bool found = res.result_type == TSA_FOUND_TARGET;
        const double cost = (double)res.cost, per_base = per_base_of(res, ref, qry);
        int ts_amount = 0;
        for (size_t i = 0; i < res.n_ops; i++) ts_amount += res.ops[i].type == TSA_OP_TS_EXIT;
        const std::string ref_name = ref.id + " " + ref.comment, qry_name = qry.id + " " + qry.comment;  // align.rs:418-419
        std::ofstream out(path);
        if (!out) die("Unable to open output file \"" + path + "\"");
        out << "type = " << (found ? "\"WithTarget\"" : "\"WithoutTarget\"") << "\n";
        if (found) {
            out << "alignment = [";
            for (size_t i = 0; i < res.n_ops; i++) {
                const tsa_op& op = res.ops[i];
                if (i) out << ", ";
                out << "[" << op.count << ", ";
                if (op.type == TSA_OP_TS_ENTRANCE)
                    out << "{ TemplateSwitchEntrance = { first_offset = " << op.value << ", equal_cost_range = { min_start = " << (int)op.min_start << ", max_start = " << (int)op.max_start
                        << ", min_end = " << (int)op.min_end << ", max_end = " << (int)op.max_end << " }, primary = \""
                        << (op.primary ? "Query" : "Reference") << "\", secondary = \"" << (op.secondary ? "Query" : "Reference") << "\", direction = \""
                        << (op.direction ? "Reverse" : "Forward") << "\" } }";
                else if (op.type == TSA_OP_TS_EXIT) out << "{ TemplateSwitchExit = { anti_primary_gap = " << op.value << " } }";
                else out << "\"" << OP_NAMES[op.type] << "\"";
                out << "]";
            }
            out << "]\n";
        }
        out << "reference_offset = " << ro << "\nquery_offset = " << qo << "\ncost = " << fmt_f64(cost) << "\ncost_per_base = " << fmt_f64(per_base)
            << "\nduration_seconds = " << fmt_f64(res.duration_seconds) << "\nopened_nodes = 0.0\nclosed_nodes = 0.0\nsuboptimal_opened_nodes = 0.0"
            << "\nsuboptimal_opened_nodes_ratio = 0.0\ntemplate_switch_amount = " << fmt_f64((double)ts_amount) << "\nruntime = 0.0\nmemory = 0.0\n\n[result]\n";
        switch (res.result_type) {
        case TSA_FOUND_TARGET: out << "astar_result_type = \"FoundTarget\"\ncost = " << res.cost << "\n"; break;
        case TSA_EXCEEDED_COST_LIMIT: out << "astar_result_type = \"ExceededCostLimit\"\ncost_limit = " << res.cost << "\n"; break;
        case TSA_EXCEEDED_MEMORY_LIMIT: out << "astar_result_type = \"ExceededMemoryLimit\"\nmax_cost = " << res.cost << "\n"; break;
        default: out << "astar_result_type = \"NoTarget\"\n";
        }
        out << "\n[sequences]\nreference_name = " << toml_string(ref_name) << "\nreference = " << toml_string(ref.seq) << "\nreference_rc = "
            << toml_string(complement_text(ref.seq, rna)) << "\nquery_name = " << toml_string(qry_name) << "\nquery = " << toml_string(qry.seq) << "\nquery_rc = "
            << toml_string(complement_text(qry.seq, rna)) << "\n";
    };

    // ---- batch front-end: --pairs ------------------------------------------------------------------------------------------
    if (!cli.pairs_file.empty()) {
        if (!cli.pair_fasta.empty() || !cli.reference.empty() || !cli.query.empty()) die("the argument '--pairs' cannot be used with '--pair-fasta' / '--reference' / '--query'", 2);
        std::vector<Record> refs, qrys;
        {
            std::ifstream probe(cli.pairs_file, std::ios::binary);
            if (!probe) die("Unable to open input file \"" + cli.pairs_file + "\"");
            int c = probe.peek();
            while (c == '\n' || c == '\r' || c == ' ') { probe.get(); c = probe.peek(); }
            if (c == '>') {
                std::vector<Record> recs = parse_fasta(cli.pairs_file);
                if (recs.size() % 2) die("Pair list fasta file must contain an even number of records, but it contains " + std::to_string(recs.size()));
                for (size_t k = 0; k + 1 < recs.size(); k += 2) { refs.push_back(recs[k]); qrys.push_back(recs[k + 1]); }
            } else {
                // TSV: name <TAB> reference <TAB> query (a two-column line has no name)
                std::string line;
                size_t ln = 0;
                while (std::getline(probe, line)) {
                    ln++;
                    if (!line.empty() && line.back() == '\r') line.pop_back();
                    if (line.empty() || line[0] == '#') continue;
                    std::vector<std::string> f;
                    size_t a = 0;
                    for (;;) { size_t b = line.find('\t', a); f.push_back(line.substr(a, b == std::string::npos ? b : b - a)); if (b == std::string::npos) break; a = b + 1; }
                    if (f.size() != 2 && f.size() != 3) die("line " + std::to_string(ln) + " of \"" + cli.pairs_file + "\": expected name<TAB>reference<TAB>query");
                    Record r, q;
                    r.id = f.size() == 3 ? f[0] : "pair" + std::to_string(refs.size()); q.id = r.id;
                    r.comment = "reference"; q.comment = "query";
                    r.seq = f[f.size() - 2]; q.seq = f[f.size() - 1];
                    refs.push_back(r); qrys.push_back(q);
                }
            }
        }
        const size_t n = refs.size();
        std::vector<tsa_pair> pairs(n);
        for (size_t k = 0; k < n; k++) {
            long long ro, rl, qo, ql;
            prepare(refs[k], qrys[k], ro, rl, qo, ql);
            pairs[k].reference = refs[k].seq.data(); pairs[k].reference_len = refs[k].seq.size();
            pairs[k].query = qrys[k].seq.data(); pairs[k].query_len = qrys[k].seq.size();
            pairs[k].reference_offset = ro; pairs[k].reference_limit = rl; pairs[k].query_offset = qo; pairs[k].query_limit = ql;
        }
        std::vector<tsa_result> results(n ? n : 1);
        int rc = tsa_align_batch(cfg, &opt, pairs.data(), n, results.data(), err, sizeof(err));
        if (rc != TSA_OK) die(std::string("alignment failed: ") + err);
        std::ofstream jl;
        if (!cli.jsonl.empty()) { jl.open(cli.jsonl); if (!jl) die("Unable to open output file \"" + cli.jsonl + "\""); }
        auto json_string = [](const std::string& s) { std::string o = "\""; for (char c : s) { if (c == '"' || c == '\\') { o.push_back('\\'); o.push_back(c); } else if ((unsigned char)c < 0x20) o += ' '; else o.push_back(c); } return o + "\""; };
        size_t failed = 0;
        // stdout: one line per pair: index, reference id, query id, result, cost, template switches, CIGAR
        for (size_t k = 0; k < n; k++) {
            const tsa_result& res = results[k];
            if (res.status != TSA_OK) {
                failed++;
                printf("%zu\t%s\t%s\tError\t-\t-\t%s\n", k, refs[k].id.c_str(), qrys[k].id.c_str(), res.message);
                if (jl.is_open()) jl << "{\"index\": " << k << ", \"reference_name\": " << json_string(refs[k].id) << ", \"query_name\": " << json_string(qrys[k].id) << ", \"error\": " << json_string(res.message) << "}\n";
                continue;
            }
            static const char* const KINDS[] = {"FoundTarget", "ExceededCostLimit", "ExceededMemoryLimit", "NoTarget"};
            const bool found = res.result_type == TSA_FOUND_TARGET;
            const std::string cg = found ? cigar(res) : std::string("-");
            printf("%zu\t%s\t%s\t%s\t%llu\t%d\t%s\n", k, refs[k].id.c_str(), qrys[k].id.c_str(), KINDS[res.result_type], (unsigned long long)res.cost, res.template_switches, cg.c_str());
            const long long ro = found ? res.reference_offset : pairs[k].reference_offset, qo = found ? res.query_offset : pairs[k].query_offset;
            if (jl.is_open())
                jl << "{\"index\": " << k << ", \"reference_name\": " << json_string(refs[k].id) << ", \"query_name\": " << json_string(qrys[k].id) << ", \"result\": \"" << KINDS[res.result_type]
                   << "\", \"cost\": " << res.cost << ", \"template_switches\": " << res.template_switches << ", \"reference_offset\": " << ro << ", \"query_offset\": " << qo
                   << ", \"reference_limit\": " << res.reference_limit << ", \"query_limit\": " << res.query_limit << ", \"cigar\": " << json_string(cg) << "}\n";
            if (!cli.output.empty()) write_toml(cli.output + "/" + std::to_string(k) + ".toml", res, refs[k], qrys[k], ro, qo);
        }
        fprintf(stderr, "%zu pairs aligned in one batch call (%zu with a per-pair error), %.3f s\n", n, failed, n ? results[0].duration_seconds * (double)n : 0.0);
        tsa_results_free(results.data(), n);
        tsa_config_free(cfg);
        return failed ? 1 : 0;
    }

    // ---- single pair (align.rs:304-335) ------------------------------------------------------------------------------------
    Record ref, qry;
    if (!cli.pair_fasta.empty()) {
        if (!cli.reference.empty() || !cli.query.empty()) die("the argument '--pair-fasta' cannot be used with '--reference' / '--query'", 2);
        std::vector<Record> recs = parse_fasta(cli.pair_fasta);
        if (recs.size() != 2) die("Pair fasta file must contain exactly two records, but it contains " + std::to_string(recs.size()));
        ref = recs[0]; qry = recs[1];
    } else if (!cli.reference.empty() && !cli.query.empty()) {
        std::vector<Record> a = parse_fasta(cli.reference), b = parse_fasta(cli.query);
        if (a.size() != 1 || b.size() != 1) die("Single fasta files must contain exactly one record");
        ref = a[0]; qry = b[0];
    } else die("No fasta input file given");
    long long ro, rl, qo, ql;
    prepare(ref, qry, ro, rl, qo, ql);
    tsa_pair pair;
    pair.reference = ref.seq.data(); pair.reference_len = ref.seq.size();
    pair.query = qry.seq.data(); pair.query_len = qry.seq.size();
    pair.reference_offset = ro; pair.reference_limit = rl; pair.query_offset = qo; pair.query_limit = ql;
    tsa_result res;
    int rc = tsa_align_batch(cfg, &opt, &pair, 1, &res, err, sizeof(err));
    if (rc != TSA_OK) die(std::string("alignment failed: ") + err);
    if (res.status == TSA_ERR_INVALID_CHAR) die(std::string(strstr(res.message, "reference") ? "Reference" : "Query") + " contains non-alphabet character: " + res.message);
    if (res.status != TSA_OK) die(std::string("alignment failed: ") + res.message);
    if (res.status == TSA_OK && res.result_type == TSA_FOUND_TARGET) { ro = res.reference_offset; qo = res.query_offset; }   // statistics follow the extended range
    const bool found = res.result_type == TSA_FOUND_TARGET;
    const double per_base = per_base_of(res, ref, qry);
    const std::string result_line = result_line_of(res);
    if (!cli.output.empty()) write_toml(cli.output, res, ref, qry, ro, qo);

    if (found) printf("CIGAR: %s\n", cigar(res).c_str()); else printf("No alignment found\n");
    printf("%s\nReference offset: %lld\nQuery offset: %lld\nCost per base: %.2f\nOpened nodes: 0\nClosed nodes: 0\nSuboptimal openend nodes: 0\n"
           "Suboptimal openend nodes per optimal opened node: 0.00\nDuration: %.2fs\n", result_line.c_str(), ro, qo, per_base, res.duration_seconds);
    tsa_results_free(&res, 1);
    tsa_config_free(cfg);
    return 0;
}
