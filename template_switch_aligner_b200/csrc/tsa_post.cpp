// tsa_post.cpp -- host post-processing of a found alignment, as the reference applies it after the search
// (lib_tsalign/src/a_star_aligner.rs:238-253): greedy extension beyond the alignment range and the equal-cost ranges of
// every template switch.  Pure host logic on the run-length encoded alignment, O(length^2) like the reference's.
#include "tsa_post.hpp"

#include <algorithm>

namespace tsa {

namespace {

bool is_entrance(const PostOp& o) { return o.type == 12; }
bool is_exit(const PostOp& o) { return o.type == 13; }
PostOp unit(int type) { PostOp o; o.count = 1; o.type = type; return o; }

// AlignmentStreamCoordinates::advance (alignment/stream.rs:188-257) over the first `upto` entries (iter_compact: the
// multiplicity of entrances and exits does not matter for coordinates).
struct Coords { int64_t r, q; int primary; bool ok; };
Coords advance(const std::vector<PostOp>& ops, size_t from, size_t upto, Coords c) {
    for (size_t k = from; k < upto && k < ops.size(); k++) {
        const PostOp& o = ops[k];
        const int64_t n = o.count;
        switch (o.type) {
        case 0: case 4: c.q += n; break;
        case 1: case 5: c.r += n; break;
        case 2: case 3: case 6: case 7: c.r += n; c.q += n; break;
        case 8: case 10: case 11: if (c.primary == 0) c.r += n; else if (c.primary == 1) c.q += n; else c.ok = false; break;
        case 9: break;
        case 12: if (c.primary >= 0) c.ok = false; c.primary = o.primary; break;
        case 13:
            if (c.primary < 0) { c.ok = false; break; }
            if (c.primary == 0) c.q += o.value; else c.r += o.value;
            if (c.q < 0 || c.r < 0) c.ok = false;
            c.primary = -1;
            break;
        default: c.ok = false;
        }
    }
    return c;
}

size_t find_exit(const std::vector<PostOp>& ops, size_t from) {
    for (size_t k = from; k < ops.size(); k++) if (is_exit(ops[k])) return k;
    return ops.size();
}

}  // namespace

// Alignment::compute_cost (alignment/template_switch_specifics.rs:591-835).  Where the reference would index out of bounds
// (a panic) or meets a flank operation (todo!()), the alignment is not one it can score: COST_INF.
uint64_t post_compute_cost(const HostConfig& cfg, const uint8_t* R, int64_t n, const uint8_t* Q, int64_t m, int64_t ro, int64_t qo,
                           const std::vector<PostOp>& ops) {
    const int A = (int)cfg.table[0].open.size();
    auto add = [](uint64_t a, uint64_t b) { return (a == COST_INF || b == COST_INF || a + b < a) ? COST_INF : a + b; };
    uint64_t cost = 0;
    int last = -1;
    int64_t ri = ro, qi = qo, pi = 0, si = 0;
    int p = 0, s = 0, d = 0;
    for (const PostOp& o : ops) {
        const int64_t reps = (o.type == 12 || o.type == 13) ? std::min<int64_t>(1, o.count) : o.count;   // alignment/iter.rs:62-90
        for (int64_t rep = 0; rep < reps; rep++) {
            uint64_t inc = 0;
            const uint8_t* P = p == 0 ? R : Q; const int64_t pl = p == 0 ? n : m;
            const uint8_t* S = s == 0 ? R : Q; const int64_t sl = s == 0 ? n : m;
            const EditTable& sec = cfg.table[d == 0 ? 1 : 2];
            auto sec_char = [&](int& out) {
                if (d == 0) { if (si < 0 || si >= sl) return false; out = S[si]; }
                else { if (si <= 0 || si > sl) return false; out = alphabet_complement(cfg.alphabet, S[si - 1]); }
                return true;
            };
            switch (o.type) {
            case 0:
                if (qi >= m) return COST_INF;
                inc = last == 0 ? cfg.table[0].ext[Q[qi]] : cfg.table[0].open[Q[qi]]; qi++; break;
            case 1:
                if (ri >= n) return COST_INF;
                inc = last == 1 ? cfg.table[0].ext[R[ri]] : cfg.table[0].open[R[ri]]; ri++; break;
            case 2: case 3:
                if (ri >= n || qi >= m) return COST_INF;
                inc = cfg.table[0].sub[(size_t)R[ri] * A + Q[qi]]; ri++; qi++; break;
            case 8:
                if (pi < 0 || pi >= pl) return COST_INF;
                inc = last == 8 ? sec.ext[P[pi]] : sec.open[P[pi]]; pi++; break;
            case 9: {
                int sc;
                if (!sec_char(sc)) return COST_INF;
                inc = last == 9 ? sec.ext[sc] : sec.open[sc];
                si += d == 0 ? 1 : -1;
                break;
            }
            case 10: case 11: {
                int sc;
                if (pi < 0 || pi >= pl || !sec_char(sc)) return COST_INF;
                inc = sec.sub[(size_t)P[pi] * A + sc];
                pi++; si += d == 0 ? 1 : -1;
                break;
            }
            case 12:
                p = o.primary; s = o.secondary; d = o.direction;
                inc = add(cfg.base[d * 4 + p * 2 + s], cfg.evaluate(p == s ? 1 : 0, o.value));
                if (inc == COST_INF) return COST_INF;
                pi = p == 0 ? ri : qi;
                si = (s == 0 ? ri : qi) + o.value;
                if (si < 0) return COST_INF;
                break;
            case 13: {
                int64_t length;
                if (p == 0) { length = pi - ri; ri = pi; qi += o.value; if (qi < 0) return COST_INF; }
                else { length = pi - qi; qi = pi; ri += o.value; if (ri < 0) return COST_INF; }
                if (length < 0) return COST_INF;
                inc = add(add(cfg.evaluate(d == 0 ? 4 : 5, o.value), cfg.evaluate(2, length)), cfg.evaluate(3, o.value - length));
                if (inc == COST_INF) return COST_INF;
                break;
            }
            default: return COST_INF;   // flank operations: todo!() in the reference
            }
            cost = add(cost, inc);
            if (cost == COST_INF) return COST_INF;
            last = o.type;   // Some(alignment_type) == last_alignment_type only matters for the gap types 0, 1, 8, 9
        }
    }
    return cost;
}

// AlignmentResult::extend_beyond_range_without_increasing_cost (alignment_result.rs:247-395).
int64_t post_extend_beyond_range(const HostConfig& cfg, const uint8_t* R, int64_t n, const uint8_t* Q, int64_t m, std::vector<PostOp>& ops,
                                 int64_t& ro, int64_t& rl, int64_t& qo, int64_t& ql) {
    if (cfg.left_flank_length > 0 || cfg.right_flank_length > 0) return 0;   // :265-268
    uint64_t current = post_compute_cost(cfg, R, n, Q, m, ro, qo, ops);
    int64_t steps = 0;
    while (ro > 0 && qo > 0) {                                                // move_offsets_left
        const int type = R[ro - 1] == Q[qo - 1] ? 3 : 2;
        if (!ops.empty() && ops.front().type == type) ops.front().count++;
        else ops.insert(ops.begin(), unit(type));
        const uint64_t now = post_compute_cost(cfg, R, n, Q, m, ro - 1, qo - 1, ops);
        if (now > current) {
            if (--ops.front().count == 0) ops.erase(ops.begin());
            break;
        }
        current = now; ro--; qo--; steps++;
    }
    while (rl < n && ql < m) {                                                // move_limits_right
        const int type = R[rl] == Q[ql] ? 3 : 2;
        if (!ops.empty() && ops.back().type == type) ops.back().count++;
        else ops.push_back(unit(type));
        const uint64_t now = post_compute_cost(cfg, R, n, Q, m, ro, qo, ops);
        if (now > current) {
            if (--ops.back().count == 0) ops.pop_back();
            break;
        }
        current = now; rl++; ql++; steps++;
    }
    return steps;
}

// Alignment::move_template_switch_start_backwards (template_switch_specifics.rs:30-170).
bool post_move_start_backwards(const uint8_t* R, int64_t n, const uint8_t* Q, int64_t m, int alphabet, int64_t ro, int64_t qo,
                               std::vector<PostOp>& ops, size_t& ci) {
    if (ci >= ops.size() || !is_entrance(ops[ci])) return false;
    const PostOp e = ops[ci];
    if (ci == 0 || !(ops[ci - 1].type == 2 || ops[ci - 1].type == 3)) return false;
    const Coords c = advance(ops, 0, ci, Coords{ro, qo, -1, true});
    if (!c.ok) return false;
    const int64_t pidx = e.primary == 0 ? c.r : c.q;
    if (pidx == 0) return false;
    const int64_t sidx = (e.secondary == 0 ? c.r : c.q) + e.value;
    if (sidx < 0) return false;
    const uint8_t* P = e.primary == 0 ? R : Q;
    const uint8_t* S = e.secondary == 0 ? R : Q;
    const int64_t sl = e.secondary == 0 ? n : m;
    if (e.direction == 0 && sidx == 0) return false;
    if (e.direction == 1 && sidx >= sl) return false;
    if (ops[ci - 1].count == 0) return false;
    if (--ops[ci - 1].count == 0) { ci--; ops.erase(ops.begin() + (long)ci); }
    const int pc = P[pidx - 1];
    const int sc = e.direction == 0 ? S[sidx - 1] : alphabet_complement(alphabet, S[sidx]);
    const int inner = pc == sc ? 11 : 10;
    if (ci + 1 < ops.size() && ops[ci + 1].type == inner) ops[ci + 1].count++;
    else ops.insert(ops.begin() + (long)ci + 1, unit(inner));
    if (e.direction == 1) ops[ci].value += 2;
    const size_t x = find_exit(ops, ci);
    if (x == ops.size()) return false;
    ops[x].value += 1;
    return true;
}

// Alignment::move_template_switch_start_forwards (:182-303).
bool post_move_start_forwards(const uint8_t* R, int64_t n, const uint8_t* Q, int64_t m, int64_t ro, int64_t qo, std::vector<PostOp>& ops, size_t& ci) {
    if (ci >= ops.size() || !is_entrance(ops[ci])) return false;
    const int direction = ops[ci].direction;
    if (ci != 0 && ops[ci - 1].type >= 4 && ops[ci - 1].type <= 7) return false;   // flanks
    if (!(ci + 1 < ops.size() && (ops[ci + 1].type == 10 || ops[ci + 1].type == 11))) return false;
    const Coords c = advance(ops, 0, ci, Coords{ro, qo, -1, true});
    if (!c.ok || c.r == n || c.q == m) return false;
    if (ops[ci + 1].count == 0) return false;
    if (--ops[ci + 1].count == 0) ops.erase(ops.begin() + (long)ci + 1);
    const int outer = R[c.r] == Q[c.q] ? 3 : 2;
    if (ci != 0 && ops[ci - 1].type == outer) ops[ci - 1].count++;
    else { ops.insert(ops.begin() + (long)ci, unit(outer)); ci++; }
    if (direction == 1) ops[ci].value -= 2;
    const size_t x = find_exit(ops, ci);
    if (x == ops.size()) return false;
    ops[x].value -= 1;
    return true;
}

// Alignment::move_template_switch_end_forwards (:305-465).
bool post_move_end_forwards(const uint8_t* R, int64_t n, const uint8_t* Q, int64_t m, int alphabet, int64_t ro, int64_t qo, std::vector<PostOp>& ops, size_t ci) {
    if (ci >= ops.size() || !is_entrance(ops[ci])) return false;
    const PostOp e = ops[ci];
    size_t x = find_exit(ops, ci);
    if (x == ops.size()) return false;
    int64_t sec_len = 0;
    for (size_t k = ci + 1; k < x; k++) {
        if (ops[k].type == 9 || ops[k].type == 10 || ops[k].type == 11) sec_len += ops[k].count;
        else if (ops[k].type != 8) return false;
    }
    if (!(x + 1 < ops.size() && (ops[x + 1].type == 2 || ops[x + 1].type == 3))) return false;
    const Coords at_entrance = advance(ops, 0, ci, Coords{ro, qo, -1, true});
    const Coords after_exit = advance(ops, ci, x + 1, at_entrance);
    if (!at_entrance.ok || !after_exit.ok) return false;
    const int64_t pidx = e.primary == 0 ? after_exit.r : after_exit.q;
    int64_t sidx = (e.secondary == 0 ? at_entrance.r : at_entrance.q) + e.value;
    if (sidx < 0) return false;
    const uint8_t* P = e.primary == 0 ? R : Q;
    const uint8_t* S = e.secondary == 0 ? R : Q;
    const int64_t pl = e.primary == 0 ? n : m, sl = e.secondary == 0 ? n : m;
    if (e.direction == 0) { sidx += sec_len; if (sidx >= sl) return false; }
    else { if (sidx < sec_len) return false; sidx -= sec_len; if (sidx == 0) return false; }
    if (ops[x + 1].count == 0) return false;
    if (--ops[x + 1].count == 0) ops.erase(ops.begin() + (long)x + 1);
    if (pidx < 0 || pidx >= pl || sidx > sl) return false;
    const int pc = P[pidx];
    const int sc = e.direction == 0 ? S[sidx] : alphabet_complement(alphabet, S[sidx - 1]);
    const int inner = pc == sc ? 11 : 10;
    if (ops[x - 1].type == inner) ops[x - 1].count++;
    else { ops.insert(ops.begin() + (long)x, unit(inner)); x++; }
    ops[x].value += 1;
    return true;
}

// Alignment::move_template_switch_end_backwards (:477-589).
bool post_move_end_backwards(const uint8_t* R, int64_t n, const uint8_t* Q, int64_t m, int64_t ro, int64_t qo, std::vector<PostOp>& ops, size_t ci) {
    (void)n; (void)m;
    if (ci >= ops.size() || !is_entrance(ops[ci])) return false;
    size_t x = find_exit(ops, ci);
    if (x == ops.size()) return false;
    if (x + 1 < ops.size() && ops[x + 1].type >= 4 && ops[x + 1].type <= 7) return false;   // flanks
    if (!(x >= 1 && (ops[x - 1].type == 10 || ops[x - 1].type == 11))) return false;
    const Coords c = advance(ops, 0, x + 1, Coords{ro, qo, -1, true});
    if (!c.ok || c.r == 0 || c.q == 0) return false;
    if (ops[x - 1].count == 0) return false;
    if (--ops[x - 1].count == 0) { x--; ops.erase(ops.begin() + (long)x); }
    const int outer = R[c.r - 1] == Q[c.q - 1] ? 3 : 2;
    if (x + 1 < ops.size() && ops[x + 1].type == outer) ops[x + 1].count++;
    else ops.insert(ops.begin() + (long)x + 1, unit(outer));
    const size_t x2 = find_exit(ops, ci);
    if (x2 == ops.size()) return false;
    ops[x2].value -= 1;
    return true;
}

// AlignmentResult::compute_ts_equal_cost_ranges (alignment_result.rs:398-573).
void post_equal_cost_ranges(const HostConfig& cfg, const uint8_t* R, int64_t n, const uint8_t* Q, int64_t m, std::vector<PostOp>& ops, int64_t ro, int64_t qo) {
    if (cfg.left_flank_length > 0 || cfg.right_flank_length > 0) return;   // :416-419
    for (size_t i = 0; i < ops.size(); i++) {
        if (!is_entrance(ops[i])) continue;
        int8_t ecr[4] = {0, 0, 0, 0};   // min_start, max_start, min_end, max_end (i8 in the reference: saturated here instead of wrapping)
        uint64_t current = post_compute_cost(cfg, R, n, Q, m, ro, qo, ops);
        {
            std::vector<PostOp> a = ops; size_t k = i;
            while (post_move_start_backwards(R, n, Q, m, cfg.alphabet, ro, qo, a, k)) {
                const uint64_t now = post_compute_cost(cfg, R, n, Q, m, ro, qo, a);
                if (now > current) break;
                current = now; if (ecr[0] == INT8_MIN) break; ecr[0]--;
            }
        }
        {
            std::vector<PostOp> a = ops; size_t k = i;
            while (post_move_start_forwards(R, n, Q, m, ro, qo, a, k)) {
                const uint64_t now = post_compute_cost(cfg, R, n, Q, m, ro, qo, a);
                if (now > current) break;
                current = now; if (ecr[1] == INT8_MAX) break; ecr[1]++;
            }
        }
        {
            std::vector<PostOp> a = ops;
            while (post_move_end_backwards(R, n, Q, m, ro, qo, a, i)) {
                const uint64_t now = post_compute_cost(cfg, R, n, Q, m, ro, qo, a);
                if (now > current) break;
                current = now; if (ecr[2] == INT8_MIN) break; ecr[2]--;
            }
        }
        {
            std::vector<PostOp> a = ops;
            while (post_move_end_forwards(R, n, Q, m, cfg.alphabet, ro, qo, a, i)) {
                const uint64_t now = post_compute_cost(cfg, R, n, Q, m, ro, qo, a);
                if (now > current) break;
                current = now; if (ecr[3] == INT8_MAX) break; ecr[3]++;
            }
        }
        for (int t = 0; t < 4; t++) ops[i].ecr[t] = ecr[t];
        ops[i].ecr_valid = true;
    }
}

}  // namespace tsa
