// astar_oracle.cpp -- TEST INFRASTRUCTURE ONLY (see tsa_oracle.h).
//
// CPU restatement of the reference's A* template-switch aligner:
//   search loop ........ generic_a_star/src/lib.rs:316-552
//   closed-list rule ... generic_a_star/src/closed_lists.rs:45-88
//   heap order ......... generic_a_star/src/comparator.rs:10-17 +
//                        lib_tsalign/.../strategies/node_ord.rs:41-69 (AntiDiagonalNodeOrd)
//   graph .............. lib_tsalign/.../template_switch_distance/context.rs:112-761,
//                        identifier.rs:12-442, ../template_switch_distance.rs:89-761
//   min-length bound ... strategies/template_switch_min_length.rs:137-235,651-684 (Lookahead)
//   total TS length .... strategies/template_switch_total_length.rs:69-109 (label-correcting)
//   no-ts .............. strategies/template_switch_count.rs:41-63 (budget 0)
//   backtrack + RLE .... lib_tsalign/src/a_star_aligner.rs:100-122, alignment_type.rs:101-139
//   rescoring .......... alignment_result/alignment/template_switch_specifics.rs:591-835
//
// Written from the behaviour of those files; not a translation of their code
// structure (index-pool nodes, one flat identifier struct).  The descendant
// strategy is AllowAny only (descendant.rs:22-36), as is the CLI/Python default.
#include "tsa_oracle.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <unordered_map>
#include <vector>

namespace {

typedef uint64_t u64;
typedef int64_t i64;
static const u64 INF = UINT64_MAX;

enum NodeType : uint8_t { PRIMARY = 0, PRIMARY_REENTRY = 1, TS_ENTRANCE = 2, SECONDARY = 3, TS_EXIT = 4 };
enum Gap : uint8_t { GAP_INS = 0, GAP_DEL = 1, GAP_NONE = 2 };  // identifier.rs:62-67
enum { EDGE_ROOT = 100, EDGE_SECONDARY_ROOT = 101, EDGE_PRIMARY_REENTRY = 102 };

// identifier.rs:12-60, flattened. Field use per type:
//   PRIMARY/PRIMARY_REENTRY: a=reference_index b=query_index c=flank_index gap
//   TS_ENTRANCE:             a=entrance_ref b=entrance_qry p,s,d c=first_offset
//   SECONDARY:               a,b entrance; p,s,d; c=length d_=primary_index e=secondary_index gap
//   TS_EXIT:                 a,b entrance; p,s,d; d_=primary_index c=anti_primary_gap
struct Id {
    int32_t a, b, c, d_, e;   // 32-bit coordinates: 28-byte identifiers keep the node pool small (sequences are far below 2^31)
    uint8_t type, gap, p, s, d;
    bool operator==(const Id& o) const {
        return a == o.a && b == o.b && c == o.c && d_ == o.d_ && e == o.e && type == o.type && gap == o.gap && p == o.p && s == o.s && d == o.d;
    }
};

struct IdHash {
    size_t operator()(const Id& i) const {
        u64 h = 0x9E3779B97F4A7C15ull;
        auto mix = [&](u64 v) { h = (h ^ v) * 0xff51afd7ed558ccdull; h ^= h >> 29; };
        mix((u64)(uint32_t)i.a | ((u64)(uint32_t)i.b << 32)); mix((u64)(uint32_t)i.c | ((u64)(uint32_t)i.d_ << 32)); mix((u64)(uint32_t)i.e);
        mix((u64)i.type | ((u64)i.gap << 8) | ((u64)i.p << 16) | ((u64)i.s << 24) | ((u64)i.d << 32));
        return (size_t)h;
    }
};

struct Edge {
    uint8_t type, p, s, d;
    i64 value;
};

static const uint32_t NO_PRED = UINT32_MAX;

// Nodes live in one pool per search (index = identity of an opened node); the predecessor is the pool index of the node that
// was expanded.  Backtracking still goes through the closed list by *identifier* (a_star_aligner.rs:100-122 looks the
// predecessor identifier up in the closed list, which holds the best node seen for it), see tsao_astar_align.
struct Node {
    Id id;
    uint32_t pred_idx = NO_PRED;
    Edge edge;
    u64 cost, lb;
    uint32_t ts_total_length;  // MaxTemplateSwitchTotalLengthStrategy memory
    uint32_t ts_count;         // MaxTemplateSwitchCountStrategy memory
    u64 f() const { return cost + lb; }
    u64 anti_diagonal() const { return id.type == PRIMARY ? (u64)(id.a + id.b) : UINT64_MAX; }  // identifier.rs:424-441
};

struct Config {
    const tsao_config* c;
    int A;
    u64 min_length;  // config/io.rs:82-84

    u64 sub(int t, int x, int y) const { return c->sub[(size_t)t * A * A + (size_t)x * A + y]; }
    u64 open(int t, int x) const { return c->open[(size_t)t * A + x]; }
    u64 ext(int t, int x) const { return c->ext[(size_t)t * A + x]; }
    u64 gap(int t, int x, bool first) const { return first ? open(t, x) : ext(t, x); }
    // cost_function.rs:39-47
    u64 fn(int k, i64 x) const {
        const i64* xs = c->fn_x[k];
        int n = c->fn_len[k];
        int lo = 0, hi = n;  // last index with xs[idx] <= x
        while (hi - lo > 1) { int mid = (lo + hi) / 2; if (xs[mid] <= x) lo = mid; else hi = mid; }
        return c->fn_c[k][lo];
    }
    // cost_function.rs:67-128 with range = x..
    u64 fn_min_from(int k, i64 x) const {
        const i64* xs = c->fn_x[k];
        int n = c->fn_len[k];
        u64 best = INF;
        for (int i = 0; i < n; i++) {
            bool last_right_of_start = (i + 1 == n) || (x <= xs[i + 1] - 1);
            if (last_right_of_start) best = std::min(best, c->fn_c[k][i]);
        }
        return best;
    }
    u64 base(int p, int s, int d) const { return c->base[d * 4 + p * 2 + s]; }  // config.rs:145-194
    int offset_fn(int p, int s) const { return p == s ? TSAO_FN_RRQQ_OFFSET : TSAO_FN_RQQR_OFFSET; }  // config.rs:108-127
    int sec_table(int d) const { return d == 0 ? TSAO_TAB_SEC_FWD : TSAO_TAB_SEC_REV; }
    int apg_fn(int d) const { return d == 0 ? TSAO_FN_FWD_APG : TSAO_FN_REV_APG; }
};

static u64 checked_add(u64 a, u64 b) {
    // cost.rs:93-99: panicking add. Costs never get near 2^64 for real inputs; saturate instead of aborting the test process.
    u64 r = a + b;
    return r < a ? INF : r;
}

struct Problem {
    Config cfg;
    const uint8_t* R; i64 n;
    const uint8_t* Q; i64 m;
    i64 ro, rl, qo, ql;
    tsao_options opt;
    std::unordered_map<Id, u64, IdHash> lookahead_memo;  // template_switch_min_length.rs:112-135
    bool label_setting() const { return !opt.total_length_maximise && !opt.force_label_correcting; }  // context.rs:758-760
};

// template_switch_distance.rs:704-761 + strategies' generate_successor
static Node make_successor(const Problem& pb, const Node& from, const Id& id, u64 inc, const Edge& edge) {
    Node s;
    s.id = id;
    s.pred_idx = NO_PRED;   // set by the search loop: the pool index of `from`
    s.edge = edge;
    s.cost = checked_add(from.cost, inc);
    s.lb = from.lb > inc ? from.lb - inc : 0;
    u64 len_inc = (edge.type == TSAO_OP_SECONDARY_MATCH || edge.type == TSAO_OP_SECONDARY_SUBSTITUTION || edge.type == TSAO_OP_SECONDARY_INSERTION) ? 1 : 0;
    s.ts_total_length = pb.opt.total_length_maximise ? from.ts_total_length + (uint32_t)len_inc : 0;  // template_switch_total_length.rs:94-108
    s.ts_count = from.ts_count;
    return s;
}

struct AStar;
static bool lookahead(Problem& pb, Node& secondary_root);

// context.rs:125-729
static void generate_successors(Problem& pb, const Node& node, std::vector<Node>& out) {
    const Config& cfg = pb.cfg;
    const Id& id = node.id;
    const i64 LF = cfg.c->left_flank_length;
    switch (id.type) {
    case PRIMARY:
    case PRIMARY_REENTRY: {
        const i64 ri = id.a, qi = id.b, flank = id.c;
        const uint8_t gap = id.gap;
        const bool can_ts = !pb.opt.no_ts;  // template_switch_count.rs:52-62: budget 0 => never
        auto primary_id = [&](i64 r, i64 q, i64 fl, uint8_t g) {
            Id s{}; s.type = PRIMARY; s.a = r; s.b = q; s.c = fl; s.gap = g; return s;  // identifier.rs:141-149 (successors are always Primary)
        };
        if (ri < pb.n && qi < pb.m) {  // NoPrune ranges: primary_range.rs:31-49
            int r = pb.R[ri], q = pb.Q[qi];
            bool is_match = r == q;
            if (flank == 0) {
                u64 inc = cfg.sub(TSAO_TAB_PRIMARY, r, q);
                if (inc != INF) {
                    Edge e{(uint8_t)(is_match ? TSAO_OP_PRIMARY_MATCH : TSAO_OP_PRIMARY_SUBSTITUTION), 0, 0, 0, 0};
                    out.push_back(make_successor(pb, node, primary_id(ri + 1, qi + 1, 0, GAP_NONE), inc, e));
                }
            }
            if ((flank < LF && can_ts) || flank < 0) {
                int t = flank < 0 ? TSAO_TAB_RIGHT_FLANK : TSAO_TAB_LEFT_FLANK;
                u64 inc = cfg.sub(t, r, q);
                if (inc != INF) {
                    Edge e{(uint8_t)(is_match ? TSAO_OP_PRIMARY_FLANK_MATCH : TSAO_OP_PRIMARY_FLANK_SUBSTITUTION), 0, 0, 0, 0};
                    out.push_back(make_successor(pb, node, primary_id(ri + 1, qi + 1, flank + 1, GAP_NONE), inc, e));
                }
            }
        }
        if (ri < pb.n) {
            int r = pb.R[ri];
            bool first = gap != GAP_DEL;
            if (flank == 0) {
                u64 inc = cfg.gap(TSAO_TAB_PRIMARY, r, first);
                if (inc != INF) out.push_back(make_successor(pb, node, primary_id(ri + 1, qi, 0, GAP_DEL), inc, Edge{TSAO_OP_PRIMARY_DELETION, 0, 0, 0, 0}));
            }
            if (flank >= 0 && flank < LF && can_ts) {
                u64 inc = cfg.gap(TSAO_TAB_LEFT_FLANK, r, first);
                if (inc != INF) out.push_back(make_successor(pb, node, primary_id(ri + 1, qi, flank + 1, GAP_DEL), inc, Edge{TSAO_OP_PRIMARY_FLANK_DELETION, 0, 0, 0, 0}));
            } else if (flank < 0) {
                u64 inc = cfg.gap(TSAO_TAB_RIGHT_FLANK, r, first);
                if (inc != INF) out.push_back(make_successor(pb, node, primary_id(ri + 1, qi, flank + 1, GAP_DEL), inc, Edge{TSAO_OP_PRIMARY_FLANK_DELETION, 0, 0, 0, 0}));
            }
        }
        if (qi < pb.m) {
            int q = pb.Q[qi];
            bool first = gap != GAP_INS;
            if (flank == 0) {
                u64 inc = cfg.gap(TSAO_TAB_PRIMARY, q, first);
                if (inc != INF) out.push_back(make_successor(pb, node, primary_id(ri, qi + 1, 0, GAP_INS), inc, Edge{TSAO_OP_PRIMARY_INSERTION, 0, 0, 0, 0}));
            }
            if (flank >= 0 && flank < LF && can_ts) {
                u64 inc = cfg.gap(TSAO_TAB_LEFT_FLANK, q, first);
                if (inc != INF) out.push_back(make_successor(pb, node, primary_id(ri, qi + 1, flank + 1, GAP_INS), inc, Edge{TSAO_OP_PRIMARY_FLANK_INSERTION, 0, 0, 0, 0}));
            } else if (flank < 0) {
                u64 inc = cfg.gap(TSAO_TAB_RIGHT_FLANK, q, first);
                if (inc != INF) out.push_back(make_successor(pb, node, primary_id(ri, qi + 1, flank + 1, GAP_INS), inc, Edge{TSAO_OP_PRIMARY_FLANK_INSERTION, 0, 0, 0, 0}));
            }
        }
        // context.rs:356-374 + template_switch_distance.rs:221-299 + identifier.rs:241-327
        if (flank == LF && can_ts) {
            u64 rq0 = cfg.fn(TSAO_FN_RQQR_OFFSET, 0), rr0 = cfg.fn(TSAO_FN_RRQQ_OFFSET, 0);
            if (rq0 != INF || rr0 != INF) {
                static const int P[8] = {0, 0, 1, 1, 0, 0, 1, 1}, S[8] = {0, 1, 0, 1, 0, 1, 0, 1}, D[8] = {0, 0, 0, 0, 1, 1, 1, 1};
                for (int k = 0; k < 8; k++) {
                    u64 base = cfg.base(P[k], S[k], D[k]);
                    u64 oc = P[k] == S[k] ? rr0 : rq0;
                    if (base == INF || oc == INF) continue;
                    int n_off = D[k] == 0 ? 2 : 1;
                    for (int t = 0; t < n_off; t++) {
                        i64 first_offset = D[k] == 0 ? (t == 0 ? -1 : 1) : 0;
                        Id s{}; s.type = TS_ENTRANCE; s.a = ri; s.b = qi; s.p = P[k]; s.s = S[k]; s.d = D[k]; s.c = first_offset;
                        Edge e{TSAO_OP_TS_ENTRANCE, (uint8_t)P[k], (uint8_t)S[k], (uint8_t)D[k], first_offset};
                        out.push_back(make_successor(pb, node, s, checked_add(base, oc), e));
                    }
                }
            }
        }
        break;
    }
    case TS_ENTRANCE: {
        // context.rs:377-489
        const i64 sec_entrance = id.s == 0 ? id.a : id.b;
        const i64 sec_len = id.s == 0 ? pb.n : pb.m;
        const i64 off = id.c;
        const i64 sidx = sec_entrance + off;
        const i64 ml = (i64)std::min<u64>(cfg.min_length, (u64)INT64_MAX / 4);
        const int ofn = cfg.offset_fn(id.p, id.s);
        auto walk = [&](i64 new_off) {
            u64 new_cost = cfg.fn(ofn, new_off);
            if (new_cost == INF) return;
            u64 old_cost = cfg.fn(ofn, off);
            if (new_cost < old_cost) abort();  // assert!(new_cost >= old_cost), context.rs:420,450
            Id s = id; s.c = new_off;
            Edge e{TSAO_OP_TS_ENTRANCE, id.p, id.s, id.d, new_off};
            out.push_back(make_successor(pb, node, s, new_cost - old_cost, e));
        };
        if (off >= 0 && (id.d == 0 ? (sidx + ml < sec_len) : (sidx < sec_len))) walk(off + 1);
        if (off <= 0 && (id.d == 0 ? (sidx > 0) : (sidx > ml))) walk(off - 1);
        bool can_start = id.d == 0 ? (sidx >= 0 && sidx + ml <= sec_len) : (sidx >= ml && sidx <= sec_len);
        if (can_start) {
            // template_switch_distance.rs:346-412
            Id s{}; s.type = SECONDARY; s.a = id.a; s.b = id.b; s.p = id.p; s.s = id.s; s.d = id.d;
            s.c = 0; s.d_ = id.p == 0 ? id.a : id.b; s.e = sidx; s.gap = GAP_NONE;
            Node root = make_successor(pb, node, s, 0, Edge{EDGE_SECONDARY_ROOT, 0, 0, 0, 0});
            if (!pb.opt.min_length_lookahead || lookahead(pb, root)) out.push_back(root);
        }
        break;
    }
    case SECONDARY: {
        // context.rs:491-634
        const uint8_t* Pseq = id.p == 0 ? pb.R : pb.Q; const i64 plen = id.p == 0 ? pb.n : pb.m;
        const uint8_t* Sseq = id.s == 0 ? pb.R : pb.Q; const i64 slen = id.s == 0 ? pb.n : pb.m;
        const i64 length = id.c, pi = id.d_, si = id.e;
        const int t = cfg.sec_table(id.d);
        if (cfg.fn_min_from(TSAO_FN_LENGTH, length) != INF) {
            bool has_sec = id.d == 0 ? si < slen : si > 0;
            int sc = 0;
            if (has_sec) sc = id.d == 0 ? Sseq[si] : pb.cfg.c->complement[Sseq[si - 1]];
            i64 si_next = id.d == 0 ? si + 1 : si - 1;
            if (pi < plen && has_sec) {
                int pc = Pseq[pi];
                u64 inc = cfg.sub(t, pc, sc);
                if (inc != INF) {
                    Id s = id; s.c = length + 1; s.d_ = pi + 1; s.e = si_next; s.gap = GAP_NONE;
                    out.push_back(make_successor(pb, node, s, inc, Edge{(uint8_t)(pc == sc ? TSAO_OP_SECONDARY_MATCH : TSAO_OP_SECONDARY_SUBSTITUTION), 0, 0, 0, 0}));
                }
            }
            if (has_sec) {  // AllowSecondaryDeletionStrategy
                u64 inc = cfg.gap(t, sc, id.gap != GAP_DEL);
                if (inc != INF) {
                    Id s = id; s.e = si_next; s.gap = GAP_DEL;
                    out.push_back(make_successor(pb, node, s, inc, Edge{TSAO_OP_SECONDARY_DELETION, 0, 0, 0, 0}));
                }
            }
            if (pi < plen) {
                int pc = Pseq[pi];
                u64 inc = cfg.gap(t, pc, id.gap != GAP_INS);
                if (inc != INF) {
                    Id s = id; s.c = length + 1; s.d_ = pi + 1; s.gap = GAP_INS;
                    out.push_back(make_successor(pb, node, s, inc, Edge{TSAO_OP_SECONDARY_INSERTION, 0, 0, 0, 0}));
                }
            }
        }
        u64 lc = cfg.fn(TSAO_FN_LENGTH, length), ldc0 = cfg.fn(TSAO_FN_LENGTH_DIFFERENCE, 0);
        if (lc != INF && ldc0 != INF) {
            Id s{}; s.type = TS_EXIT; s.a = id.a; s.b = id.b; s.p = id.p; s.s = id.s; s.d = id.d; s.d_ = pi; s.c = length;
            out.push_back(make_successor(pb, node, s, checked_add(lc, ldc0), Edge{TSAO_OP_TS_EXIT, 0, 0, 0, length}));
        }
        break;
    }
    case TS_EXIT: {
        // context.rs:636-722
        const i64 anti_start = 0, anti_end = id.p == 0 ? pb.m : pb.n;
        const i64 entrance_primary = id.p == 0 ? id.a : id.b;
        const i64 pi = id.d_, apg = id.c;
        const i64 ld = apg - (pi - entrance_primary);
        auto walk = [&](i64 new_ld, i64 new_apg) {
            u64 new_cost = cfg.fn(TSAO_FN_LENGTH_DIFFERENCE, new_ld);
            if (new_cost == INF) return;
            u64 old_cost = cfg.fn(TSAO_FN_LENGTH_DIFFERENCE, ld);
            if (new_cost < old_cost) abort();
            Id s = id; s.c = new_apg;
            out.push_back(make_successor(pb, node, s, new_cost - old_cost, Edge{TSAO_OP_TS_EXIT, 0, 0, 0, new_apg}));
        };
        if (ld >= 0 && pi + ld < anti_end) walk(ld + 1, apg + 1);
        if (ld <= 0 && pi + ld > anti_start) walk(ld - 1, apg - 1);
        u64 apg_cost = cfg.fn(cfg.apg_fn(id.d), apg);
        if (apg_cost != INF) {
            // template_switch_distance.rs:579-644
            i64 ri, qi;
            bool ok = true;
            if (id.p == 0) { qi = id.b + apg; ri = pi; if (qi < 0 || qi >= pb.m) ok = false; }
            else { ri = id.a + apg; qi = pi; if (ri < 0 || ri >= pb.n) ok = false; }
            if (ok) {
                Id s{}; s.type = PRIMARY_REENTRY; s.a = ri; s.b = qi; s.c = -cfg.c->right_flank_length; s.gap = GAP_NONE;
                Node r = make_successor(pb, node, s, apg_cost, Edge{EDGE_PRIMARY_REENTRY, 0, 0, 0, 0});
                r.ts_count += 1;
                out.push_back(r);
            }
        }
        break;
    }
    }
}

// Heap order: best = min (f, cost), then larger anti-diagonal, then larger secondary score
// (node_ord.rs:41-69 wrapped by comparator.rs:10-17).
struct HeapItem {
    u64 f, cost;
    uint32_t ad, score;   // anti-diagonal (UINT32_MAX for non-primary nodes, identifier.rs:424-441), total TS length
    uint32_t idx;
};

// Closed list (generic_a_star/src/closed_lists.rs:45-88): identifier -> pool index of the node kept for it.  Open addressing
// over 32-bit pool indices; the identifiers are read from the pool.
struct Store {
    std::vector<Node> pool;
    std::vector<uint32_t> slots;   // pool index + 1, 0 = empty
    size_t count = 0, mask = 0;
    Store() { slots.assign(1 << 12, 0); mask = slots.size() - 1; }
    size_t size() const { return count; }
    // returns the slot of `id` (occupied) or the empty slot where it belongs
    uint32_t* locate(const Id& id) {
        size_t h = IdHash()(id) & mask;
        for (;;) {
            uint32_t& sl = slots[h];
            if (sl == 0 || pool[sl - 1].id == id) return &sl;
            h = (h + 1) & mask;
        }
    }
    const Node* find(const Id& id) { const uint32_t sl = *locate(id); return sl ? &pool[sl - 1] : nullptr; }
    void put(uint32_t idx) {
        uint32_t* sl = locate(pool[idx].id);
        if (*sl == 0) {
            count++;
            *sl = idx + 1;
            if (count * 10 > slots.size() * 6) grow();
        } else *sl = idx + 1;
    }
    void grow() {
        std::vector<uint32_t> old;
        old.swap(slots);
        slots.assign(old.size() * 2, 0);
        mask = slots.size() - 1;
        for (uint32_t v : old) if (v) {
            size_t h = IdHash()(pool[v - 1].id) & mask;
            while (slots[h]) h = (h + 1) & mask;
            slots[h] = v;
        }
    }
};
struct HeapLess {  // "a is worse than b" for std::priority_queue (top = best)
    bool operator()(const HeapItem& a, const HeapItem& b) const {
        if (a.f != b.f) return a.f > b.f;
        if (a.cost != b.cost) return a.cost > b.cost;
        if (a.ad != b.ad) return a.ad < b.ad;
        return a.score < b.score;
    }
};

struct SearchResult {
    int type;
    u64 cost;
    Id target;
    u64 opened = 0, closed = 0, suboptimal = 0;
};

// generic_a_star/src/lib.rs:316-552.  `is_target` and label mode are parameters so the same loop serves the
// top-level search and the min-length lookahead (template_switch_min_length.rs:651-684).
template <class IsTarget>
static SearchResult search(Problem& pb, const Node& root, bool label_setting, IsTarget is_target, Store& closed) {
    std::vector<Node>& pool = closed.pool;
    std::priority_queue<HeapItem, std::vector<HeapItem>, HeapLess> open;
    auto push = [&](const Node& nd) {
        if (pool.size() >= (size_t)NO_PRED - 1) abort();   // 32-bit pool indices
        pool.push_back(nd);
        const u64 ad = nd.anti_diagonal();
        open.push(HeapItem{nd.f(), nd.cost, ad == UINT64_MAX ? UINT32_MAX : (uint32_t)ad, nd.ts_total_length, (uint32_t)(pool.size() - 1)});
    };
    SearchResult res;
    const u64 cost_limit = pb.opt.cost_limit;
    bool applied_cost_limit = false;
    const double node_bytes = 160.0;  // size_of::<Node>() + Box; parity-unpinned (layout of the Rust struct)
    const u64 node_count_limit = pb.opt.memory_limit == UINT64_MAX ? UINT64_MAX : (u64)((double)pb.opt.memory_limit / node_bytes / 2.3 + 0.5);

    push(root);
    bool have_target = false;
    Id target_id{};
    u64 target_cost = INF, target_score = 0;
    std::vector<Node> succ;

    auto node_better = [](const Node& x, const Node& y) {  // AStarNodeComparator.compare(x, y) == Greater
        if (x.f() != y.f()) return x.f() < y.f();
        if (x.cost != y.cost) return x.cost < y.cost;
        u64 ax = x.anti_diagonal(), ay = y.anti_diagonal();
        if (ax != ay) return ax > ay;
        return x.ts_total_length > y.ts_total_length;
    };

    for (;;) {
        if (open.empty()) {
            if (applied_cost_limit) { res.type = TSAO_EXCEEDED_COST_LIMIT; res.cost = cost_limit; return res; }
            if (have_target) break;
            res.type = TSAO_NO_TARGET; res.cost = 0; return res;
        }
        const uint32_t node_idx = open.top().idx;
        const Node node = pool[node_idx];
        open.pop();
        if (node.f() > cost_limit) { res.type = TSAO_EXCEEDED_COST_LIMIT; res.cost = cost_limit; return res; }
        if ((u64)closed.size() + (u64)open.size() > node_count_limit) { res.type = TSAO_EXCEEDED_MEMORY_LIMIT; res.cost = node.cost; return res; }
        if (node.f() > target_cost) break;

        const bool tgt = is_target(node);
        const Node* it = closed.find(node.id);
        bool skip = false;
        if (it) skip = label_setting ? true : !node_better(node, *it);
        if (skip) {
            res.suboptimal++;
            const u64 existing_cost = it->cost, existing_score = it->ts_total_length;
            if (tgt && (node.cost < std::min(target_cost, existing_cost) ||
                        (node.cost == std::min(target_cost, existing_cost) && node.ts_total_length > std::max(target_score, existing_score)))) {
                have_target = true; target_id = node.id; target_cost = node.cost; target_score = node.ts_total_length;
                if (label_setting) { closed.put(node_idx); res.closed++; break; }
            } else if (tgt && (existing_cost < target_cost || (existing_cost == target_cost && node.ts_total_length > existing_score))) {
                have_target = true; target_id = it->id; target_cost = it->cost; target_score = it->ts_total_length;
                if (label_setting) { res.closed++; break; }
            }
            continue;
        }

        succ.clear();
        generate_successors(pb, node, succ);
        for (Node& s : succ) {
            s.pred_idx = node_idx;
            if (s.f() <= cost_limit) { push(s); res.opened++; }
            else applied_cost_limit = true;
        }

        if (tgt && (node.cost < target_cost || (node.cost == target_cost && node.ts_total_length > target_score))) {
            have_target = true; target_id = node.id; target_cost = node.cost; target_score = node.ts_total_length;
            if (label_setting) { closed.put(node_idx); res.closed++; break; }
        }
        closed.put(node_idx);
        res.closed++;
    }
    if (!have_target) { res.type = TSAO_NO_TARGET; res.cost = 0; return res; }
    res.type = TSAO_FOUND_TARGET;
    res.target = target_id;
    res.cost = closed.find(target_id)->cost;
    return res;
}

// template_switch_min_length.rs:137-235
static bool lookahead(Problem& pb, Node& root) {
    Id key{}; key.type = SECONDARY; key.p = root.id.p; key.s = root.id.s; key.d = root.id.d; key.d_ = root.id.d_; key.e = root.id.e;
    auto it = pb.lookahead_memo.find(key);
    if (it != pb.lookahead_memo.end()) {
        if (it->second == INF) return false;  // not cached by the reference (it re-searches); same outcome
        root.lb = std::max(root.lb, it->second);
        return true;
    }
    Store closed;
    const u64 ml = pb.cfg.min_length;
    tsao_options saved = pb.opt;
    pb.opt.min_length_lookahead = 0;  // nested graph never reaches an entrance
    SearchResult r = search(pb, root, /*label_setting=*/true, [&](const Node& nd) { return nd.id.type == SECONDARY && (u64)nd.id.c == ml; }, closed);
    pb.opt = saved;
    if (r.type != TSAO_FOUND_TARGET) {
        if (r.type == TSAO_NO_TARGET) pb.lookahead_memo[key] = INF;
        return false;
    }
    u64 lb = r.cost - root.cost;
    pb.lookahead_memo[key] = lb;
    root.lb = std::max(root.lb, lb);
    return true;
}

// alignment_type.rs:101-139
static bool is_repeated(const Edge& self, const Edge& prev) {
    auto grp = [](int t) {
        switch (t) {
        case TSAO_OP_PRIMARY_INSERTION: case TSAO_OP_PRIMARY_FLANK_INSERTION: return 1;
        case TSAO_OP_PRIMARY_DELETION: case TSAO_OP_PRIMARY_FLANK_DELETION: return 2;
        case TSAO_OP_PRIMARY_SUBSTITUTION: case TSAO_OP_PRIMARY_FLANK_SUBSTITUTION: return 3;
        case TSAO_OP_PRIMARY_MATCH: case TSAO_OP_PRIMARY_FLANK_MATCH: return 4;
        default: return 0;
        }
    };
    if (grp(self.type) && grp(self.type) == grp(prev.type)) return true;
    if (self.type == TSAO_OP_TS_ENTRANCE && prev.type == TSAO_OP_TS_ENTRANCE) return self.p == prev.p && self.s == prev.s;
    if (self.type == TSAO_OP_TS_EXIT && prev.type == TSAO_OP_TS_EXIT) return true;
    return self.type == prev.type && self.p == prev.p && self.s == prev.s && self.d == prev.d && self.value == prev.value;
}

}  // namespace

extern "C" int tsao_astar_align(const tsao_config* c, const uint8_t* reference, int64_t n, const uint8_t* query, int64_t m,
                                int64_t ro, int64_t rl, int64_t qo, int64_t ql, const tsao_options* opt, tsao_result* out) {
    Problem pb;
    pb.cfg.c = c;
    pb.cfg.A = c->alphabet_size;
    pb.cfg.min_length = UINT64_MAX;
    for (int i = 0; i < c->fn_len[TSAO_FN_LENGTH]; i++)
        if (c->fn_c[TSAO_FN_LENGTH][i] != INF) { pb.cfg.min_length = (u64)c->fn_x[TSAO_FN_LENGTH][i]; break; }  // cost_function.rs:49-60
    pb.R = reference; pb.n = n; pb.Q = query; pb.m = m;
    pb.ro = ro; pb.rl = rl; pb.qo = qo; pb.ql = ql;
    pb.opt = *opt;

    Node root{};  // context.rs:112-123
    root.id.type = PRIMARY; root.id.a = ro; root.id.b = qo; root.id.c = 0; root.id.gap = GAP_NONE;
    root.pred_idx = NO_PRED; root.edge = Edge{EDGE_ROOT, 0, 0, 0, 0};

    Store closed;
    SearchResult r = search(pb, root, pb.label_setting(),
                            [&](const Node& nd) { return (nd.id.type == PRIMARY || nd.id.type == PRIMARY_REENTRY) && nd.id.a == rl && nd.id.b == ql; },  // context.rs:731-748
                            closed);
    memset(out, 0, sizeof(*out));
    out->result_type = r.type;
    out->cost = r.cost;
    out->opened_nodes = r.opened; out->closed_nodes = r.closed; out->suboptimal_opened_nodes = r.suboptimal;
    if (r.type != TSAO_FOUND_TARGET) return 0;

    // a_star_aligner.rs:100-122: walk predecessor edges target -> root, drop internal ops, merge runs.
    std::vector<std::pair<i64, Edge>> rle;
    const Node* cur = closed.find(r.target);
    out->ts_total_length = cur->ts_total_length;
    while (cur->pred_idx != NO_PRED) {
        const Edge& e = cur->edge;
        if (e.type < 100) {
            if (!rle.empty() && is_repeated(e, rle.back().second)) rle.back().first++;
            else rle.push_back({1, e});
        }
        cur = closed.find(closed.pool[cur->pred_idx].id);   // the closed list's node for the predecessor identifier
    }
    std::reverse(rle.begin(), rle.end());
    out->n_ops = (i64)rle.size();
    out->ops = (tsao_op*)malloc(sizeof(tsao_op) * std::max<size_t>(1, rle.size()));
    for (size_t i = 0; i < rle.size(); i++) {
        const Edge& e = rle[i].second;
        out->ops[i] = tsao_op{rle[i].first, e.type, e.p, e.s, e.d, e.value};
    }
    return 0;
}

// template_switch_specifics.rs:591-835 (flat iteration clamps non-repeatable ops to multiplicity 1: iter.rs:62-90)
extern "C" uint64_t tsao_rescore_mode(const tsao_config* c, const uint8_t* R, int64_t n, const uint8_t* Q, int64_t m,
                                      int64_t ro, int64_t qo, const tsao_op* ops, int64_t n_ops,
                                      int64_t* end_ref, int64_t* end_qry, int32_t* ok, int32_t as_searched) {
    Config cfg; cfg.c = c; cfg.A = c->alphabet_size; cfg.min_length = 0;
    u64 cost = 0;
    int last = -1, last_group = -1;
    i64 right_flank_remaining = 0;
    i64 ri = ro, qi = qo, pi = 0, si = 0;
    int p = 0, s = 0, d = 0;
    *ok = 1;
    auto fail = [&]() { *ok = 0; if (end_ref) *end_ref = ri; if (end_qry) *end_qry = qi; return INF; };
    for (i64 k = 0; k < n_ops; k++) {
        const tsao_op& op = ops[k];
        i64 reps = (op.type == TSAO_OP_TS_ENTRANCE || op.type == TSAO_OP_TS_EXIT) ? std::min<i64>(1, op.count) : op.count;
        for (i64 rep = 0; rep < reps; rep++) {
            u64 inc = 0;
            switch (op.type) {
            case TSAO_OP_PRIMARY_INSERTION: case TSAO_OP_PRIMARY_FLANK_INSERTION:
            case TSAO_OP_PRIMARY_DELETION: case TSAO_OP_PRIMARY_FLANK_DELETION:
            case TSAO_OP_PRIMARY_SUBSTITUTION: case TSAO_OP_PRIMARY_MATCH:
            case TSAO_OP_PRIMARY_FLANK_SUBSTITUTION: case TSAO_OP_PRIMARY_FLANK_MATCH: {
                // Non-flank ops follow :614-649.  Flank ops are todo!() in the reference (:650-655); here they are
                // charged as the search charges them (context.rs:225-353): the first right_flank_length flank ops
                // after an exit use the right-flank table, all other flank ops the left-flank table, and a gap
                // continues across the flank boundary.
                bool flank = op.type >= TSAO_OP_PRIMARY_FLANK_INSERTION;
                int t = TSAO_TAB_PRIMARY;
                if (flank) { if (right_flank_remaining > 0) { t = TSAO_TAB_RIGHT_FLANK; right_flank_remaining--; } else t = TSAO_TAB_LEFT_FLANK; }
                int grp = op.type & 3;  // 0 ins, 1 del, 2 sub, 3 match
                if (grp == 0) {
                    if (qi >= m) return fail();
                    inc = last_group == 0 ? cfg.ext(t, Q[qi]) : cfg.open(t, Q[qi]);
                    qi++;
                } else if (grp == 1) {
                    if (ri >= n) return fail();
                    inc = last_group == 1 ? cfg.ext(t, R[ri]) : cfg.open(t, R[ri]);
                    ri++;
                } else {
                    if (ri >= n || qi >= m) return fail();
                    inc = cfg.sub(t, R[ri], Q[qi]);
                    ri++; qi++;
                }
                break;
            }
            case TSAO_OP_SECONDARY_INSERTION: {
                const uint8_t* P = p == 0 ? R : Q; i64 pl = p == 0 ? n : m;
                if (pi >= pl) return fail();
                int t = cfg.sec_table(d);
                inc = last == op.type ? cfg.ext(t, P[pi]) : cfg.open(t, P[pi]);
                pi++;
                break;
            }
            case TSAO_OP_SECONDARY_DELETION: {
                const uint8_t* S = s == 0 ? R : Q; i64 sl = s == 0 ? n : m;
                int sc;
                if (d == 0) { if (si >= sl) return fail(); sc = S[si]; } else { if (si <= 0 || si > sl) return fail(); sc = c->complement[S[si - 1]]; }
                int t = cfg.sec_table(d);
                inc = last == op.type ? cfg.ext(t, sc) : cfg.open(t, sc);
                si += d == 0 ? 1 : -1;
                break;
            }
            case TSAO_OP_SECONDARY_SUBSTITUTION: case TSAO_OP_SECONDARY_MATCH: {
                const uint8_t* P = p == 0 ? R : Q; i64 pl = p == 0 ? n : m;
                const uint8_t* S = s == 0 ? R : Q; i64 sl = s == 0 ? n : m;
                if (pi >= pl) return fail();
                int sc;
                if (d == 0) { if (si >= sl) return fail(); sc = S[si]; } else { if (si <= 0 || si > sl) return fail(); sc = c->complement[S[si - 1]]; }
                inc = cfg.sub(cfg.sec_table(d), P[pi], sc);
                pi++;
                si += d == 0 ? 1 : -1;
                break;
            }
            case TSAO_OP_TS_ENTRANCE: {
                p = op.primary; s = op.secondary; d = op.direction;
                inc = checked_add(cfg.base(p, s, d), cfg.fn(cfg.offset_fn(p, s), op.value));
                if (as_searched && d == 0) {
                    // The search charges a forward entrance oc(0) and then walks from offset +-1
                    // (identifier.rs:290-319, context.rs:392-462): oc(0) + oc(o) - oc(+-1).  compute_cost charges
                    // oc(o); the two agree whenever oc is flat around 0 (all shipped configs).
                    int fnk = cfg.offset_fn(p, s);
                    i64 sign = op.value < 0 ? -1 : 1;
                    u64 walk = (op.value == sign) ? 0 : cfg.fn(fnk, op.value) - cfg.fn(fnk, sign);
                    inc = checked_add(cfg.base(p, s, d), checked_add(cfg.fn(fnk, 0), walk));
                }
                if (inc == INF) { *ok = 1; return INF; }
                pi = p == 0 ? ri : qi;
                si = (s == 0 ? ri : qi) + op.value;
                if (si < 0) return fail();
                break;
            }
            case TSAO_OP_TS_EXIT: {
                i64 apg = op.value, length;
                if (p == 0) { length = pi - ri; ri = pi; qi += apg; if (qi < 0) return fail(); }
                else { length = pi - qi; qi = pi; ri += apg; if (ri < 0) return fail(); }
                if (length < 0) return fail();
                i64 ld = apg - length;
                inc = cfg.fn(cfg.apg_fn(d), apg);
                inc = checked_add(inc, cfg.fn(TSAO_FN_LENGTH, length));
                inc = checked_add(inc, cfg.fn(TSAO_FN_LENGTH_DIFFERENCE, ld));
                if (inc == INF) return INF;
                break;
            }
            default: return fail();
            }
            cost = checked_add(cost, inc);
            if (cost == INF) return INF;
            last = op.type;
            last_group = op.type < 8 ? (op.type & 3) : -1;
            if (op.type == TSAO_OP_TS_EXIT) right_flank_remaining = c->right_flank_length;
        }
    }
    if (end_ref) *end_ref = ri;
    if (end_qry) *end_qry = qi;
    return cost;
}

extern "C" uint64_t tsao_rescore(const tsao_config* c, const uint8_t* R, int64_t n, const uint8_t* Q, int64_t m,
                                 int64_t ro, int64_t qo, const tsao_op* ops, int64_t n_ops,
                                 int64_t* end_ref, int64_t* end_qry, int32_t* ok) {
    return tsao_rescore_mode(c, R, n, Q, m, ro, qo, ops, n_ops, end_ref, end_qry, ok, 0);
}

extern "C" void tsao_result_free(tsao_result* r) {
    if (r && r->ops) { free(r->ops); r->ops = nullptr; r->n_ops = 0; }
}
