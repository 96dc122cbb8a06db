// tsa_config.cpp -- see tsa_config.hpp.
#include "tsa_config.hpp"

#include <algorithm>
#include <climits>
#include <cstring>
#include <sstream>

namespace tsa {

// ------------------------------------------------------------------------------------------------ alphabets
// compact-genome 12.5.0 (Cargo.toml:27): DnaAlphabet index order A,C,G,T is pinned by
// lib_tsalign/src/costs/gap_affine/io/tests.rs:8-24; the complement A<->T, C<->G by the reference_rc / query_rc
// strings of test_files/*.toml.  N, U and the IUPAC codes follow the IUPAC standard (unpinned by the reference).
static const char* const ALPHABET_CHARS[ALPHA_COUNT] = {"ACGT", "ACGTN", "ACGU", "ACGUN", "ACGTRYSWKMBDHVN", "ACGURYSWKMBDHVN"};
static const char* const ALPHABET_NAMES[ALPHA_COUNT] = {"dna", "dna-n", "rna", "rna-n", "dna-iupac", "rna-iupac"};

const char* alphabet_chars(int a) { return (a >= 0 && a < ALPHA_COUNT) ? ALPHABET_CHARS[a] : ""; }
const char* alphabet_name(int a) { return (a >= 0 && a < ALPHA_COUNT) ? ALPHABET_NAMES[a] : "?"; }
int alphabet_from_name(const std::string& name) {
    for (int a = 0; a < ALPHA_COUNT; a++) if (name == ALPHABET_NAMES[a]) return a;
    return -1;
}
int alphabet_index(int a, unsigned char ascii) {
    const char* cs = alphabet_chars(a);
    const char* p = ascii ? strchr(cs, ascii) : nullptr;
    return p ? (int)(p - cs) : -1;
}
int alphabet_complement(int a, int index) {
    const char* cs = alphabet_chars(a);
    char c = cs[index], r;
    switch (c) {
    case 'A': r = (a == ALPHA_RNA || a == ALPHA_RNA_N || a == ALPHA_RNA_IUPAC) ? 'U' : 'T'; break;
    case 'T': case 'U': r = 'A'; break;
    case 'C': r = 'G'; break;
    case 'G': r = 'C'; break;
    case 'R': r = 'Y'; break;
    case 'Y': r = 'R'; break;
    case 'K': r = 'M'; break;
    case 'M': r = 'K'; break;
    case 'B': r = 'V'; break;
    case 'V': r = 'B'; break;
    case 'D': r = 'H'; break;
    case 'H': r = 'D'; break;
    default: r = c;  // S, W, N
    }
    return alphabet_index(a, (unsigned char)r);
}

// ------------------------------------------------------------------------------------------------ HostConfig
uint64_t HostConfig::evaluate(int k, int64_t x) const {
    const StepFunction& f = fn[k];
    size_t lo = 0, hi = f.size();
    while (hi - lo > 1) { size_t mid = (lo + hi) / 2; if (f[mid].first <= x) lo = mid; else hi = mid; }
    return f[lo].second;
}
int64_t HostConfig::min_length() const {
    for (const auto& p : fn[2]) if (p.second != COST_INF) return p.first;
    return -1;
}

// ------------------------------------------------------------------------------------------------ parser
namespace {

struct Cursor {
    const std::string& s;
    size_t pos = 0;
    std::string err;
    explicit Cursor(const std::string& text) : s(text) {}
    bool eof() const { return pos >= s.size(); }
    char peek() const { return eof() ? '\0' : s[pos]; }
    void skip_any_ws() { while (!eof() && (s[pos] == ' ' || s[pos] == '\t' || s[pos] == '\n' || s[pos] == '\r')) pos++; }
    void skip_ws() { while (!eof() && (s[pos] == ' ' || s[pos] == '\t')) pos++; }
    bool at_eol() const { return eof() || s[pos] == '\n' || s[pos] == '\r'; }
    bool tag(const char* t) {
        size_t n = strlen(t);
        if (s.compare(pos, n, t) == 0) { pos += n; return true; }
        return false;
    }
    bool fail(const std::string& what) {
        if (err.empty()) {
            size_t line = 1 + std::count(s.begin(), s.begin() + std::min(pos, s.size()), '\n');
            err = what + " (line " + std::to_string(line) + ")";
        }
        return false;
    }
};

// parse_inf_value (config/io.rs:181-221): [+-](inf|digits); -inf is the type minimum, inf the maximum.
bool parse_value(Cursor& c, bool is_signed, int64_t& sv, uint64_t& uv) {
    size_t start = c.pos;
    bool neg = false;
    if (c.peek() == '-') { neg = true; c.pos++; } else if (c.peek() == '+') c.pos++;
    if (c.tag("inf")) {
        if (is_signed) sv = neg ? INT64_MIN : INT64_MAX; else uv = neg ? 0 : COST_INF;
        return true;
    }
    size_t d0 = c.pos;
    unsigned __int128 acc = 0;
    while (!c.eof() && c.peek() >= '0' && c.peek() <= '9') { acc = acc * 10 + (unsigned)(c.peek() - '0'); if (acc > ((unsigned __int128)1 << 65)) acc = (unsigned __int128)1 << 65; c.pos++; }
    if (c.pos == d0) { c.pos = start; return c.fail("expected a number or inf"); }
    if (is_signed) {
        if (acc > (unsigned __int128)INT64_MAX + (neg ? 1 : 0)) return c.fail("number out of range");
        sv = neg ? (int64_t)(-(__int128)acc) : (int64_t)acc;
    } else {
        if (neg && acc != 0) return c.fail("negative value for an unsigned field");
        if (acc > (unsigned __int128)UINT64_MAX) return c.fail("number out of range");
        uv = (uint64_t)acc;
    }
    return true;
}

bool parse_section(Cursor& c, const char* name) {  // "# <name>" on its own line
    c.skip_any_ws();
    if (!c.tag("#")) return c.fail(std::string("expected section '# ") + name + "'");
    c.skip_ws();
    if (!c.tag(name)) return c.fail(std::string("expected section '# ") + name + "'");
    c.skip_ws();
    if (!c.at_eol()) return c.fail(std::string("trailing characters after section '# ") + name + "'");
    c.skip_any_ws();
    return true;
}

bool parse_key_value(Cursor& c, const char* key, bool is_signed, int64_t& sv, uint64_t& uv) {
    c.skip_any_ws();
    size_t k0 = c.pos;
    while (!c.eof() && (isalnum((unsigned char)c.peek()) || c.peek() == '_')) c.pos++;
    if (c.s.compare(k0, c.pos - k0, key) != 0 || c.pos - k0 != strlen(key)) { c.pos = k0; return c.fail(std::string("expected '") + key + " = <value>'"); }
    c.skip_ws();
    if (!c.tag("=")) return c.fail("expected '='");
    c.skip_ws();
    return parse_value(c, is_signed, sv, uv);
}

// CostFunction::parse_plain (costs/cost_function/io.rs:81-120)
bool parse_step_function(Cursor& c, const char* name, bool index_signed, StepFunction& out) {
    c.skip_any_ws();
    if (!c.tag(name)) return c.fail(std::string("expected cost function '") + name + "'");
    c.skip_any_ws();
    std::vector<int64_t> xs;
    while (!c.at_eol()) {
        int64_t sv = 0; uint64_t uv = 0;
        if (!parse_value(c, index_signed, sv, uv)) return false;
        xs.push_back(index_signed ? sv : (int64_t)std::min<uint64_t>(uv, (uint64_t)INT64_MAX));
        c.skip_ws();
    }
    c.skip_any_ws();
    std::vector<uint64_t> cs;
    while (!c.at_eol()) {
        int64_t sv = 0; uint64_t uv = 0;
        if (!parse_value(c, false, sv, uv)) return false;
        cs.push_back(uv);
        c.skip_ws();
    }
    const int64_t first = index_signed ? INT64_MIN : 0;
    bool ok = xs.size() == cs.size() && !xs.empty() && xs[0] == first;
    for (size_t i = 1; ok && i < xs.size(); i++) ok = xs[i - 1] < xs[i];
    if (!ok) return c.fail(std::string("malformed cost function '") + name + "' (first index must be the type minimum, indexes strictly increasing)");
    out.clear();
    for (size_t i = 0; i < xs.size(); i++) out.emplace_back(xs[i], cs[i]);
    return true;
}

bool parse_characters(Cursor& c, int alphabet, int A, std::vector<int>& order) {
    order.clear();
    for (int i = 0; i < A; i++) {
        c.skip_ws();
        int idx = c.eof() ? -1 : alphabet_index(alphabet, (unsigned char)c.peek());
        if (idx < 0) return c.fail("expected an alphabet character");
        c.pos++;
        order.push_back(idx);
    }
    std::vector<int> sorted = order;
    std::sort(sorted.begin(), sorted.end());
    if (std::unique(sorted.begin(), sorted.end()) != sorted.end()) return c.fail("duplicate alphabet character");
    return true;
}

bool parse_cost_vector(Cursor& c, const char* name, int alphabet, int A, std::vector<uint64_t>& out) {
    c.skip_any_ws();
    if (!c.tag(name)) return c.fail(std::string("expected '") + name + "'");
    c.skip_any_ws();
    std::vector<int> order;
    if (!parse_characters(c, alphabet, A, order)) return false;
    c.skip_any_ws();
    out.assign(A, 0);
    for (int i = 0; i < A; i++) {
        c.skip_ws();
        int64_t sv = 0; uint64_t uv = 0;
        if (!parse_value(c, false, sv, uv)) return false;
        out[order[i]] = uv;
    }
    return true;
}

// GapAffineAlignmentCostTable::parse_plain (costs/gap_affine/io.rs:156-359)
bool parse_table(Cursor& c, const char* name, int alphabet, EditTable& t) {
    const int A = (int)strlen(alphabet_chars(alphabet));
    if (!parse_section(c, name)) return false;
    c.skip_any_ws();
    if (!c.tag("SubstitutionCostTable")) return c.fail("expected 'SubstitutionCostTable'");
    c.skip_any_ws();
    if (!c.tag("|")) return c.fail("expected '|' before the column characters");
    std::vector<int> cols;
    if (!parse_characters(c, alphabet, A, cols)) return false;
    c.skip_any_ws();
    size_t dashes = 0;
    while (c.peek() == '-') { c.pos++; dashes++; }
    if (dashes == 0 || !c.tag("+")) return c.fail("expected the '---+---' separator line");
    dashes = 0;
    while (c.peek() == '-') { c.pos++; dashes++; }
    if (dashes == 0) return c.fail("expected the '---+---' separator line");
    t.name = name;
    t.sub.assign((size_t)A * A, 0);
    std::vector<bool> seen(A, false);
    for (int r = 0; r < A; r++) {
        c.skip_any_ws();
        int row = c.eof() ? -1 : alphabet_index(alphabet, (unsigned char)c.peek());
        if (row < 0) return c.fail("expected a row character");
        c.pos++;
        c.skip_ws();
        if (!c.tag("|")) return c.fail("expected '|' after the row character");
        for (int k = 0; k < A; k++) {
            c.skip_ws();
            int64_t sv = 0; uint64_t uv = 0;
            if (!parse_value(c, false, sv, uv)) return false;
            t.sub[(size_t)row * A + cols[k]] = uv;
        }
        seen[row] = true;
    }
    for (int r = 0; r < A; r++) if (!seen[r]) return c.fail("a substitution row is missing");
    if (!parse_cost_vector(c, "GapOpenCostVector", alphabet, A, t.open)) return false;
    if (!parse_cost_vector(c, "GapExtendCostVector", alphabet, A, t.ext)) return false;
    return true;
}

// CostFunction::is_v_shaped (costs/cost_function.rs:170-176)
bool is_v_shaped(const StepFunction& f) {
    for (size_t i = 1; i < f.size(); i++) {
        const auto& a = f[i - 1]; const auto& b = f[i];
        bool ok = (a.first < 0 && b.first > 0) || (a.first < 0 && a.second >= b.second) || (a.first >= 0 && a.second <= b.second);
        if (!ok) return false;
    }
    return true;
}

const char* const FN_NAMES[6] = {"RQQROffset", "RRQQOffset", "Length", "LengthDifference", "ForwardAntiPrimaryGap", "ReverseAntiPrimaryGap"};
const char* const TABLE_NAMES[5] = {"Primary Edit Costs", "Secondary Forward Edit Costs", "Secondary Reverse Edit Costs", "Left Flank Edit Costs", "Right Flank Edit Costs"};
const char* const BASE_NAMES[8] = {"rrf_cost", "rqf_cost", "qrf_cost", "qqf_cost", "rrr_cost", "rqr_cost", "qrr_cost", "qqr_cost"};

}  // namespace

int parse_config(const std::string& text, int alphabet, HostConfig& out, std::string& err) {
    if (alphabet < 0 || alphabet >= ALPHA_COUNT) { err = "unknown alphabet"; return CFG_ALPHABET; }
    Cursor c(text);
    HostConfig cfg;
    cfg.alphabet = alphabet;
    int64_t sv = 0; uint64_t uv = 0;
    bool ok = parse_section(c, "Limits") && parse_key_value(c, "left_flank_length", true, cfg.left_flank_length, uv) &&
              parse_key_value(c, "right_flank_length", true, cfg.right_flank_length, uv) && parse_section(c, "Base Cost");
    for (int k = 0; ok && k < 8; k++) ok = parse_key_value(c, BASE_NAMES[k], false, sv, cfg.base[k]);
    ok = ok && parse_section(c, "Jump Costs");
    for (int k = 0; ok && k < 6; k++) ok = parse_step_function(c, FN_NAMES[k], k != 2, cfg.fn[k]);
    for (int k = 0; ok && k < 5; k++) ok = parse_table(c, TABLE_NAMES[k], alphabet, cfg.table[k]);
    if (!ok) { err = c.err.empty() ? "parse error" : c.err; return CFG_PARSE; }
    if (!is_v_shaped(cfg.fn[0])) { err = "RQQROffset costs are not V-shaped"; return CFG_RQQR_NOT_V; }
    if (!is_v_shaped(cfg.fn[1])) { err = "RRQQOffset costs are not V-shaped"; return CFG_RRQQ_NOT_V; }
    if (!is_v_shaped(cfg.fn[3])) { err = "LengthDifference costs are not V-shaped"; return CFG_LENDIFF_NOT_V; }
    out = cfg;
    return CFG_OK;
}

HostConfig default_config(int alphabet) {
    HostConfig cfg;
    cfg.alphabet = alphabet;
    const int A = (int)strlen(alphabet_chars(alphabet));
    const uint64_t base[8] = {4, 4, 4, 4, 3, 2, 2, 3};
    memcpy(cfg.base, base, sizeof(base));
    for (int t = 0; t < 5; t++) {
        EditTable& e = cfg.table[t];
        e.name = TABLE_NAMES[t];
        e.sub.assign((size_t)A * A, 2);
        for (int x = 0; x < A; x++) e.sub[(size_t)x * A + x] = 0;
        e.open.assign(A, 3);
        e.ext.assign(A, 1);
    }
    cfg.fn[0] = {{INT64_MIN, COST_INF}, {-100, 0}, {101, COST_INF}};
    cfg.fn[1] = {{INT64_MIN, COST_INF}, {-100, 0}, {1, COST_INF}};
    cfg.fn[2] = {{0, COST_INF}, {5, 0}};
    cfg.fn[3] = {{INT64_MIN, COST_INF}, {-100, 0}, {101, COST_INF}};
    cfg.fn[4] = {{INT64_MIN, COST_INF}, {-100, 0}, {101, COST_INF}};
    cfg.fn[5] = {{INT64_MIN, COST_INF}, {-100, 0}, {101, COST_INF}};
    return cfg;
}

std::string write_config(const HostConfig& cfg) {
    std::ostringstream o;
    const char* cs = alphabet_chars(cfg.alphabet);
    const int A = (int)strlen(cs);
    auto cost = [](uint64_t v) { return v == COST_INF ? std::string("inf") : std::to_string(v); };
    o << "# Limits\n\nleft_flank_length = " << cfg.left_flank_length << "\nright_flank_length = " << cfg.right_flank_length << "\n\n# Base Cost\n\n";
    for (int k = 0; k < 8; k++) o << BASE_NAMES[k] << " = " << cost(cfg.base[k]) << "\n";
    o << "\n# Jump Costs\n";
    for (int k = 0; k < 6; k++) {
        o << "\n" << FN_NAMES[k] << "\n";
        for (const auto& p : cfg.fn[k]) {
            if (p.first == INT64_MIN) o << " -inf"; else if (p.first == INT64_MAX) o << " inf"; else o << " " << p.first;
        }
        o << "\n";
        for (const auto& p : cfg.fn[k]) o << " " << cost(p.second);
        o << "\n";
    }
    for (int t = 0; t < 5; t++) {
        const EditTable& e = cfg.table[t];
        o << "\n# " << TABLE_NAMES[t] << "\n\nSubstitutionCostTable\n  |";
        for (int x = 0; x < A; x++) o << " " << cs[x];
        o << "\n--+" << std::string(2 * A, '-') << "\n";
        for (int r = 0; r < A; r++) {
            o << cs[r] << " |";
            for (int x = 0; x < A; x++) o << " " << cost(e.sub[(size_t)r * A + x]);
            o << "\n";
        }
        o << "\nGapOpenCostVector\n";
        for (int x = 0; x < A; x++) o << " " << cs[x];
        o << "\n";
        for (int x = 0; x < A; x++) o << " " << cost(e.open[x]);
        o << "\n\nGapExtendCostVector\n";
        for (int x = 0; x < A; x++) o << " " << cs[x];
        o << "\n";
        for (int x = 0; x < A; x++) o << " " << cost(e.ext[x]);
        o << "\n";
    }
    return o.str();
}

// ------------------------------------------------------------------------------------------------ flatten
namespace {
const int64_t SPAN = 1 << 28;  // piece bounds are clipped to +-SPAN so that index arithmetic stays in int

int clamp_cost(uint64_t v) { return v >= (uint64_t)INF32 ? INF32 : (int)v; }

bool pieces_of(const StepFunction& f, std::vector<Piece>& out, std::string& err, const char* what) {
    out.clear();
    for (size_t i = 0; i < f.size(); i++) {
        if (f[i].second >= (uint64_t)INF32) continue;
        int64_t lo = std::max<int64_t>(f[i].first, -SPAN);
        int64_t hi = std::min<int64_t>(i + 1 < f.size() ? f[i + 1].first - 1 : INT64_MAX, SPAN);
        if (lo <= hi) out.push_back(Piece{(int)lo, (int)hi, (int)f[i].second});
    }
    if (out.size() > (size_t)MAX_PIECES) { err = std::string(what) + " has more than " + std::to_string(MAX_PIECES) + " finite pieces (unsupported)"; return false; }
    return true;
}
int min_cost(const std::vector<Piece>& v) { int m = INF32; for (const Piece& p : v) m = std::min(m, p.cost); return m; }
}  // namespace

bool flatten_config(const HostConfig& cfg, DevConfig& dev, std::vector<int>& lc_dense, std::string& err) {
    memset(&dev, 0, sizeof(dev));
    const int A = (int)strlen(alphabet_chars(cfg.alphabet));
    dev.A = A;
    for (int t = 0; t < 5; t++) {
        for (int x = 0; x < MAX_ALPHABET; x++) {
            for (int y = 0; y < MAX_ALPHABET; y++) dev.sub[t][x * MAX_ALPHABET + y] = (x < A && y < A) ? clamp_cost(cfg.table[t].sub[(size_t)x * A + y]) : INF32;
            dev.open[t][x] = x < A ? clamp_cost(cfg.table[t].open[x]) : INF32;
            dev.ext[t][x] = x < A ? clamp_cost(cfg.table[t].ext[x]) : INF32;
        }
    }
    for (int x = 0; x < A; x++) dev.comp[x] = (uint8_t)alphabet_complement(cfg.alphabet, x);
    if (cfg.left_flank_length < 0 || cfg.right_flank_length < 0 || cfg.left_flank_length > 1000 || cfg.right_flank_length > 1000) { err = "flank length out of range"; return false; }
    dev.left_flank = (int)cfg.left_flank_length;
    dev.right_flank = (int)cfg.right_flank_length;

    // Length costs: dense table up to the last breakpoint, constant tail after it.
    const StepFunction& lf = cfg.fn[2];
    dev.ml = -1; dev.lmax = -1;
    const int64_t last_x = lf.back().first;
    if (last_x > (1 << 20)) { err = "Length cost function breakpoints beyond 2^20 are unsupported"; return false; }
    lc_dense.assign((size_t)last_x + 1, INF32);
    int min_lc = INF32;
    for (size_t i = 0; i < lf.size(); i++) {
        int c = clamp_cost(lf[i].second);
        int64_t hi = i + 1 < lf.size() ? lf[i + 1].first - 1 : last_x;
        for (int64_t x = lf[i].first; x <= hi; x++) lc_dense[(size_t)x] = c;
        if (c < INF32) {
            if (dev.ml < 0) dev.ml = (int)lf[i].first;
            dev.lmax = i + 1 < lf.size() ? (int)hi : INT32_MAX / 2;
            min_lc = std::min(min_lc, c);
        }
    }
    dev.n_lc = (int)lc_dense.size();
    dev.lc_tail = clamp_cost(lf.back().second);

    std::vector<Piece> ld, apg[2], oc_raw[2];
    if (!pieces_of(cfg.fn[3], ld, err, "LengthDifference") || !pieces_of(cfg.fn[4], apg[0], err, "ForwardAntiPrimaryGap") ||
        !pieces_of(cfg.fn[5], apg[1], err, "ReverseAntiPrimaryGap") || !pieces_of(cfg.fn[0], oc_raw[0], err, "RQQROffset") ||
        !pieces_of(cfg.fn[1], oc_raw[1], err, "RRQQOffset"))
        return false;
    dev.n_ld = (int)ld.size();
    for (size_t i = 0; i < ld.size(); i++) dev.ld[i] = ld[i];
    dev.ld_lo = INT32_MAX; dev.ld_hi = INT32_MIN;
    for (const Piece& pc : ld) { dev.ld_lo = std::min(dev.ld_lo, pc.lo); dev.ld_hi = std::max(dev.ld_hi, pc.hi); }
    const bool exit_possible = cfg.evaluate(3, 0) < (uint64_t)INF32;  // exit needs ldc(0) finite (context.rs:622-633)

    dev.n_kinds = 0;
    dev.min_ts = INF32;
    for (int k = 0; k < 8 && dev.ml >= 0 && exit_possible; k++) {
        const int p = (k >> 1) & 1, s = k & 1, d = k >> 2;
        const int fnk = p == s ? 1 : 0;
        const uint64_t oc0 = cfg.evaluate(fnk, 0);
        if (cfg.base[k] >= (uint64_t)INF32 || oc0 >= (uint64_t)INF32) continue;  // context.rs:356-374
        KindDesc kd;
        memset(&kd, 0, sizeof(kd));
        kd.p = p; kd.s = s; kd.d = d; kd.base = (int)cfg.base[k]; kd.table = d == 0 ? 1 : 2;
        std::vector<Piece> oc;
        if (d == 1) oc = oc_raw[fnk];
        else {
            // Forward entrances start at offset +-1 and are charged oc(0); walking further out adds oc(o) - oc(+-1);
            // offset 0 is unreachable (identifier.rs:290-319, context.rs:392-462).
            for (int sign = -1; sign <= 1; sign += 2) {
                const uint64_t oc1 = cfg.evaluate(fnk, sign);
                oc.push_back(Piece{sign, sign, (int)oc0});
                if (oc1 >= (uint64_t)INF32) continue;
                for (const Piece& pc : oc_raw[fnk]) {
                    int lo = pc.lo, hi = pc.hi;
                    if (sign > 0) lo = std::max(lo, 2); else hi = std::min(hi, -2);
                    if (lo <= hi) oc.push_back(Piece{lo, hi, (int)oc0 + pc.cost - (int)oc1});
                }
            }
        }
        {   // merge adjacent pieces of equal cost (fewer window queries per row)
            std::sort(oc.begin(), oc.end(), [](const Piece& a, const Piece& b) { return a.lo < b.lo; });
            std::vector<Piece> merged;
            for (const Piece& pc : oc) {
                if (!merged.empty() && merged.back().cost == pc.cost && merged.back().hi + 1 == pc.lo) merged.back().hi = pc.hi;
                else merged.push_back(pc);
            }
            oc.swap(merged);
        }
        if (oc.size() > (size_t)MAX_PIECES) { err = "offset cost function has too many pieces (unsupported)"; return false; }
        if (oc.empty() || apg[d].empty() || ld.empty()) continue;
        kd.n_oc = (int)oc.size();
        kd.oc_lo = oc.front().lo; kd.oc_hi = oc.back().hi;
        kd.min_open = INF32;
        for (int x = 0; x < A; x++) kd.min_open = std::min(kd.min_open, dev.open[kd.table][x]);
        kd.oc_skip0 = 1;
        for (const Piece& pc : oc) if (pc.lo <= 0 && 0 <= pc.hi) kd.oc_skip0 = 0;
        for (size_t i = 0; i < oc.size(); i++) kd.oc[i] = oc[i];
        kd.n_apg = (int)apg[d].size();
        for (size_t i = 0; i < apg[d].size(); i++) kd.apg[i] = apg[d][i];
        kd.apg_nonpos = 1;
        for (const Piece& pc : apg[d]) if (pc.hi > 0) kd.apg_nonpos = 0;
        kd.apg_lo = INT32_MAX; kd.apg_hi = INT32_MIN;
        for (const Piece& pc : apg[d]) { kd.apg_lo = std::min(kd.apg_lo, pc.lo); kd.apg_hi = std::max(kd.apg_hi, pc.hi); }
        kd.min_ext = INF32;
        for (int x = 0; x < A; x++) kd.min_ext = std::min(kd.min_ext, dev.ext[kd.table][x]);
        kd.min_rest_nolc = min_cost(oc) + min_cost(ld) + min_cost(apg[d]);
        kd.min_rest = kd.min_rest_nolc + min_lc;
        dev.min_ts = std::min(dev.min_ts, kd.base + kd.min_rest);
        dev.kinds[dev.n_kinds++] = kd;
    }
    return true;
}

}  // namespace tsa
