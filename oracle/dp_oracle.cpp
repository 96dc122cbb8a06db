// dp_oracle.cpp -- TEST INFRASTRUCTURE ONLY (see tsa_oracle.h).
//
// Scalar statement of the layered min-plus DP the CUDA kernels implement.  It solves
// the same shortest-path problem as the reference's A* (restated in astar_oracle.cpp): the graph of
//   lib_tsalign/src/a_star_aligner/template_switch_distance/context.rs:125-729
// but filled densely, layer by layer, where layer k holds the states reached after exactly k completed
// template switches.  tests/test_dp_vs_astar.py proves it equal to astar_oracle.cpp on the reference's
// test_files and on random inputs/configs; the GPU path is then diffed against this file at sizes the A*
// cannot reach.
//
//   primary moves ........ context.rs:135-354  (fill_layer)
//   TS entrance/offset ... context.rs:356-489, identifier.rs:241-327  (jump: OffsetPieces)
//   inner alignment ...... context.rs:491-634  (Chain: cost-to-go rows V_l(x))
//   exit / reentry ....... context.rs:636-722, template_switch_distance.rs:579-644  (jump: exit windows)
//   target ............... context.rs:731-748
//   RLE ops .............. a_star_aligner.rs:100-122, alignment_type.rs:101-139
#include "tsa_oracle.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <vector>

namespace {

typedef int64_t i64;
typedef uint64_t u64;
static const i64 INF = INT64_MAX / 4;  // every finite cost in the tests is far below this

static inline i64 sat(u64 c) { return c >= (u64)INF ? INF : (i64)c; }
static inline i64 add(i64 a, i64 b) { return (a >= INF || b >= INF) ? INF : std::min(INF, a + b); }

struct Cfg {
    const tsao_config* c;
    int A;
    i64 sub(int t, int x, int y) const { return sat(c->sub[(size_t)t * A * A + (size_t)x * A + y]); }
    i64 open(int t, int x) const { return sat(c->open[(size_t)t * A + x]); }
    i64 ext(int t, int x) const { return sat(c->ext[(size_t)t * A + x]); }
    i64 gap(int t, int x, bool first) const { return first ? open(t, x) : ext(t, x); }
    i64 fn(int k, i64 x) const {  // cost_function.rs:39-47
        const i64* xs = c->fn_x[k];
        int lo = 0, hi = c->fn_len[k];
        while (hi - lo > 1) { int mid = (lo + hi) / 2; if (xs[mid] <= x) lo = mid; else hi = mid; }
        return sat(c->fn_c[k][lo]);
    }
    i64 base(int p, int s, int d) const { return sat(c->base[d * 4 + p * 2 + s]); }
};

// A maximal interval [lo, hi] on which a step function is constant and finite.
struct Piece { i64 lo, hi, cost; };

// Finite constant pieces of step function k, clipped to [clip_lo, clip_hi].
static std::vector<Piece> pieces_of(const Cfg& cfg, int k, i64 clip_lo, i64 clip_hi) {
    std::vector<Piece> out;
    int n = cfg.c->fn_len[k];
    for (int i = 0; i < n; i++) {
        i64 c = sat(cfg.c->fn_c[k][i]);
        if (c >= INF) continue;
        i64 lo = cfg.c->fn_x[k][i];
        i64 hi = (i + 1 < n) ? cfg.c->fn_x[k][i + 1] - 1 : INT64_MAX;
        lo = std::max(lo, clip_lo); hi = std::min(hi, clip_hi);
        if (lo <= hi) out.push_back(Piece{lo, hi, c});
    }
    return out;
}

// b[q] = min_{x in [q+lo, q+hi] and 0 <= x < |a|} a[x]  (monotonic deque), arg[q] = that x or -1.
static void window_min(const std::vector<i64>& a, i64 lo, i64 hi, size_t nq, std::vector<i64>& b, std::vector<i64>& arg) {
    b.assign(nq, INF); arg.assign(nq, -1);
    std::deque<i64> dq;
    i64 next = 0, na = (i64)a.size();
    for (i64 q = 0; q < (i64)nq; q++) {
        i64 wl = q + lo, wh = std::min(q + hi, na - 1);
        while (next <= wh) {
            if (next >= 0) {
                while (!dq.empty() && a[dq.back()] > a[next]) dq.pop_back();
                dq.push_back(next);
            }
            next++;
        }
        while (!dq.empty() && dq.front() < wl) dq.pop_front();
        if (!dq.empty() && a[dq.front()] < INF) { b[q] = a[dq.front()]; arg[q] = dq.front(); }
    }
}

enum { G_INS = 0, G_DEL = 1, G_NONE = 2 };
enum { MV_DIAG = 0, MV_DIAG_FLANK = 1, MV_DEL = 2, MV_DEL_FLANK = 3, MV_INS = 4, MV_INS_FLANK = 5, MV_SEED = 6 };

struct SeedFrom { int8_t kind; int32_t len, entr_anti, e; };  // provenance of a reentry seed

struct Layer {
    std::vector<uint8_t> pred;     // [g][fi][cell]
    std::vector<i64> dmin;         // [cell] min_g cost at f == L_f (TS entrance cost)
    std::vector<uint8_t> dmin_g;   // argmin g
    std::vector<i64> seed;         // [cell] incoming seeds (state f = -R_f, g = None); layer 0: root only
    std::vector<SeedFrom> from;    // provenance of seed
    i64 target = INF; int tf = 0, tg = 0;
};

struct Solver {
    Cfg cfg;
    const uint8_t* R; i64 n;
    const uint8_t* Q; i64 m;
    i64 ro, rl, qo, ql;
    tsao_options opt;
    i64 LF, RF; int F;
    i64 ml;        // template_switch_min_length, -1 if none (config/io.rs:82-84)
    i64 Lmax;      // largest length with finite Length cost, capped by sequence lengths
    size_t cells;
    std::vector<Layer> layers;

    size_t cell(i64 i, i64 j) const { return (size_t)i * (m + 1) + j; }
    size_t st(int g, i64 f, size_t c) const { return ((size_t)g * F + (size_t)(f + RF)) * cells + c; }

    // ---- primary relaxation of one layer (context.rs:135-354) ------------------------------------------------
    void fill_layer(Layer& L, bool is_root_layer) {
        std::vector<i64> C((size_t)3 * F * cells, INF);
        L.pred.assign((size_t)3 * F * cells, 255);
        const bool can_ts = !opt.no_ts;
        const i64 seed_f = is_root_layer ? 0 : -RF;
        auto relax = [&](int g, i64 f, size_t c, i64 v, int mv, int gprev) {
            size_t s = st(g, f, c);
            if (v < C[s]) { C[s] = v; L.pred[s] = (uint8_t)(mv * 3 + gprev); }
        };
        for (i64 i = 0; i <= n; i++) for (i64 j = 0; j <= m; j++) {
            size_t c = cell(i, j);
            if (L.seed[c] < INF) relax(G_NONE, seed_f, c, L.seed[c], MV_SEED, G_NONE);
            for (i64 f = -RF; f <= LF; f++) for (int g = 0; g < 3; g++) {
                i64 v = C[st(g, f, c)];
                if (v >= INF) continue;
                if (i < n && j < m) {
                    int r = R[i], q = Q[j];
                    if (f == 0) relax(G_NONE, 0, cell(i + 1, j + 1), add(v, cfg.sub(TSAO_TAB_PRIMARY, r, q)), MV_DIAG, g);
                    if ((f < LF && can_ts) || f < 0) {
                        // NB context.rs:225-259: for f >= 0 this needs f < L_f; for f < 0 always the right-flank table
                        int t = f < 0 ? TSAO_TAB_RIGHT_FLANK : TSAO_TAB_LEFT_FLANK;
                        if (f >= 0 || true) relax(G_NONE, f + 1, cell(i + 1, j + 1), add(v, cfg.sub(t, r, q)), MV_DIAG_FLANK, g);
                    }
                }
                if (i < n) {
                    int r = R[i]; bool first = g != G_DEL;
                    if (f == 0) relax(G_DEL, 0, cell(i + 1, j), add(v, cfg.gap(TSAO_TAB_PRIMARY, r, first)), MV_DEL, g);
                    if (f >= 0 && f < LF && can_ts) relax(G_DEL, f + 1, cell(i + 1, j), add(v, cfg.gap(TSAO_TAB_LEFT_FLANK, r, first)), MV_DEL_FLANK, g);
                    else if (f < 0) relax(G_DEL, f + 1, cell(i + 1, j), add(v, cfg.gap(TSAO_TAB_RIGHT_FLANK, r, first)), MV_DEL_FLANK, g);
                }
                if (j < m) {
                    int q = Q[j]; bool first = g != G_INS;
                    if (f == 0) relax(G_INS, 0, cell(i, j + 1), add(v, cfg.gap(TSAO_TAB_PRIMARY, q, first)), MV_INS, g);
                    if (f >= 0 && f < LF && can_ts) relax(G_INS, f + 1, cell(i, j + 1), add(v, cfg.gap(TSAO_TAB_LEFT_FLANK, q, first)), MV_INS_FLANK, g);
                    else if (f < 0) relax(G_INS, f + 1, cell(i, j + 1), add(v, cfg.gap(TSAO_TAB_RIGHT_FLANK, q, first)), MV_INS_FLANK, g);
                }
            }
        }
        // Target: any flank, any gap (context.rs:731-748).  Tie order: smallest f, then g = None, Del, Ins.
        L.target = INF;
        static const int gorder[3] = {G_NONE, G_DEL, G_INS};
        size_t tc = cell(rl, ql);
        for (i64 f = -RF; f <= LF; f++) for (int gi = 0; gi < 3; gi++) {
            i64 v = C[st(gorder[gi], f, tc)];
            if (v < L.target) { L.target = v; L.tf = (int)f; L.tg = gorder[gi]; }
        }
        L.dmin.assign(cells, INF); L.dmin_g.assign(cells, G_NONE);
        for (size_t c = 0; c < cells; c++) for (int gi = 0; gi < 3; gi++) {
            i64 v = C[st(gorder[gi], LF, c)];
            if (v < L.dmin[c]) { L.dmin[c] = v; L.dmin_g[c] = (uint8_t)gorder[gi]; }
        }
    }

    // ---- one inner chain: cost-to-go rows for a fixed primary end e (context.rs:491-634 read backwards) -----
    // T[y] is the secondary in reading order (forward: S; reverse: reverse complement of S), y in [0, ns].
    // V[l][g][y] = min cost of consuming exactly the primary characters P[e-l .. e) starting at secondary
    // position y with incoming gap state g (exit is allowed in any gap state, so V[0] = 0).
    struct Chain {
        i64 ns, rows;
        std::vector<i64> V;  // [l][g][y]
        i64& at(i64 l, int g, i64 y) { return V[((size_t)l * 3 + g) * (ns + 1) + y]; }
    };
    void run_chain(Chain& ch, const uint8_t* P, i64 e, const std::vector<uint8_t>& T, int table, i64 rows) {
        i64 ns = (i64)T.size();
        ch.ns = ns; ch.rows = rows;
        ch.V.assign((size_t)(rows + 1) * 3 * (ns + 1), INF);
        for (int g = 0; g < 3; g++) for (i64 y = 0; y <= ns; y++) ch.at(0, g, y) = 0;
        for (i64 l = 1; l <= rows; l++) {
            int pc = P[e - l];
            for (i64 y = ns; y >= 0; y--) {
                for (int g = 0; g < 3; g++) {
                    i64 best = INF;
                    if (y < ns) best = std::min(best, add(cfg.sub(table, pc, T[y]), ch.at(l - 1, G_NONE, y + 1)));             // diag
                    if (y < ns) best = std::min(best, add(cfg.gap(table, T[y], g != G_DEL), ch.at(l, G_DEL, y + 1)));           // deletion (secondary only)
                    best = std::min(best, add(cfg.gap(table, pc, g != G_INS), ch.at(l - 1, G_INS, y)));                        // insertion (primary only)
                    ch.at(l, g, y) = best;
                }
            }
        }
    }

    struct Kind { int p, s, d; i64 base; int idx; };
    std::vector<Kind> kinds() const {
        std::vector<Kind> out;
        i64 rq0 = cfg.fn(TSAO_FN_RQQR_OFFSET, 0), rr0 = cfg.fn(TSAO_FN_RRQQ_OFFSET, 0);
        if (ml < 0) return out;
        for (int k = 0; k < 8; k++) {
            int p = (k >> 1) & 1, s = k & 1, d = k >> 2;
            i64 b = cfg.base(p, s, d), oc0 = p == s ? rr0 : rq0;
            if (b >= INF || oc0 >= INF) continue;  // context.rs:356-374
            out.push_back(Kind{p, s, d, b, k});
        }
        return out;
    }

    // Effective entrance cost as a function of the first offset o: base excluded (A.3).
    //   reverse: oc(o);   forward: o = +-1 -> oc(0), |o| >= 2 -> oc(0) + oc(o) - oc(+-1), o = 0 unreachable.
    std::vector<Piece> offset_pieces(const Kind& k) const {
        int fnk = k.p == k.s ? TSAO_FN_RRQQ_OFFSET : TSAO_FN_RQQR_OFFSET;
        i64 span = n + m + 2;
        std::vector<Piece> raw = pieces_of(cfg, fnk, -span, span), out;
        if (k.d == 1) return raw;
        i64 oc0 = cfg.fn(fnk, 0);
        for (int sign = -1; sign <= 1; sign += 2) {
            i64 oc1 = cfg.fn(fnk, sign);
            out.push_back(Piece{sign, sign, oc0});
            if (oc1 >= INF) continue;
            for (const Piece& pc : raw) {
                i64 lo = pc.lo, hi = pc.hi;
                if (sign > 0) lo = std::max<i64>(lo, 2); else hi = std::min<i64>(hi, -2);
                if (lo <= hi) out.push_back(Piece{lo, hi, oc0 + pc.cost - oc1});
            }
        }
        return out;
    }

    void secondary_string(const Kind& k, std::vector<uint8_t>& T) const {
        const uint8_t* S = k.s == 0 ? R : Q; i64 ns = k.s == 0 ? n : m;
        T.resize(ns);
        for (i64 y = 0; y < ns; y++) T[y] = k.d == 0 ? S[y] : cfg.c->complement[S[ns - 1 - y]];
    }

    // jump-in for one chain row: E[j] = min_o OC(o) + V_l^None(start(sE + o)) with the A.3 start bounds.
    // For p == s the entrance secondary coordinate is the primary one (a scalar); otherwise it is j.
    void jump_in(const Kind& k, Chain& ch, i64 l, i64 ip, i64 na, const std::vector<Piece>& ocp,
                 std::vector<i64>& E, std::vector<i64>& Eoff) {
        i64 ns = ch.ns;
        // start cost per secondary boundary index b (reference coordinates); INF where the start is not allowed
        std::vector<i64> start(ns + 1, INF);
        for (i64 b = 0; b <= ns; b++) {
            bool ok = k.d == 0 ? (b + ml <= ns) : (b >= ml);
            if (!ok) continue;
            i64 y = k.d == 0 ? b : ns - b;
            start[b] = ch.at(l, G_NONE, y);
        }
        E.assign(na + 1, INF); Eoff.assign(na + 1, 0);
        std::vector<i64> wb, wa;
        for (const Piece& pc : ocp) {
            if (k.p == k.s) {
                for (i64 o = std::max(pc.lo, -ip); o <= std::min(pc.hi, ns - ip); o++) {
                    i64 v = add(pc.cost, start[ip + o]);
                    if (v < E[0] || (v == E[0] && v < INF && std::llabs(o) < std::llabs(Eoff[0]))) { E[0] = v; Eoff[0] = o; }
                }
            } else {
                window_min(start, pc.lo, pc.hi, (size_t)na + 1, wb, wa);
                for (i64 j = 0; j <= na; j++) {
                    i64 v = add(pc.cost, wb[j]);
                    if (v < E[j]) { E[j] = v; Eoff[j] = wa[j] - j; }
                }
            }
        }
        if (k.p == k.s) for (i64 j = 1; j <= na; j++) { E[j] = E[0]; Eoff[j] = Eoff[0]; }
    }

    // ---- TS transitions layer k -> seeds of layer k+1 --------------------------------------------------------
    void jump(const Layer& from, Layer& to, i64 bound) {
        to.seed.assign(cells, INF);
        to.from.assign(cells, SeedFrom{-1, 0, 0, 0});
        if (opt.no_ts) return;
        i64 ldc0 = cfg.fn(TSAO_FN_LENGTH_DIFFERENCE, 0);
        if (ldc0 >= INF) return;  // exit needs ldc(0) finite (context.rs:622-633)
        std::vector<uint8_t> T;
        Chain ch;
        std::vector<i64> E, Eoff, X, wb, wa;
        for (const Kind& k : kinds()) {
            const uint8_t* P = k.p == 0 ? R : Q; i64 np = k.p == 0 ? n : m;
            i64 na = k.p == 0 ? m : n;  // anti-primary length
            secondary_string(k, T);
            int table = k.d == 0 ? TSAO_TAB_SEC_FWD : TSAO_TAB_SEC_REV;
            std::vector<Piece> ocp = offset_pieces(k);
            for (i64 e = 0; e <= np; e++) {
                i64 rows = std::min(e, Lmax);
                if (rows < ml) continue;
                run_chain(ch, P, e, T, table, rows);
                // A.5 walk bounds (mixed coordinates, as written in context.rs:662-663,685-687)
                i64 ld_min = -e, ld_max = std::max<i64>(0, na - e);
                std::vector<Piece> ldp = pieces_of(cfg, TSAO_FN_LENGTH_DIFFERENCE, ld_min, ld_max);
                std::vector<Piece> apgp = pieces_of(cfg, k.d == 0 ? TSAO_FN_FWD_APG : TSAO_FN_REV_APG, -(n + m + 2), n + m + 2);
                for (i64 l = std::max<i64>(ml, 0); l <= rows; l++) {
                    i64 lc = cfg.fn(TSAO_FN_LENGTH, l);
                    if (lc >= INF) continue;
                    i64 ip = e - l;
                    jump_in(k, ch, l, ip, na, ocp, E, Eoff);
                    X.assign(na + 1, INF);
                    bool any = false;
                    for (i64 j = 0; j <= na; j++) {
                        size_t c = k.p == 0 ? cell(ip, j) : cell(j, ip);
                        i64 v = add(add(from.dmin[c], k.base), add(E[j], lc));
                        if (v < bound) { X[j] = v; any = true; }
                    }
                    if (!any) continue;
                    for (const Piece& lp : ldp) for (const Piece& ap : apgp) {
                        // reentry anti coordinate j2 = j + apg, apg = l + ld:  ld in lp, apg in ap
                        i64 dlo = std::max(lp.lo + l, ap.lo), dhi = std::min(lp.hi + l, ap.hi);  // range of j2 - j
                        if (dlo > dhi) continue;
                        window_min(X, -dhi, -dlo, (size_t)na, wb, wa);  // j2 in [0, na): reentry at na is rejected
                        for (i64 j2 = 0; j2 < na; j2++) {
                            i64 v = add(wb[j2], lp.cost + ap.cost);
                            size_t c = k.p == 0 ? cell(e, j2) : cell(j2, e);
                            if (v < to.seed[c]) { to.seed[c] = v; to.from[c] = SeedFrom{(int8_t)k.idx, (int32_t)l, (int32_t)wa[j2], (int32_t)e}; }
                        }
                    }
                }
            }
        }
    }
};

struct OpOut { std::vector<tsao_op> ops; };

static int group_of(int t) {
    switch (t) {
    case TSAO_OP_PRIMARY_INSERTION: case TSAO_OP_PRIMARY_FLANK_INSERTION: return 1;
    case TSAO_OP_PRIMARY_DELETION: case TSAO_OP_PRIMARY_FLANK_DELETION: return 2;
    case TSAO_OP_PRIMARY_SUBSTITUTION: case TSAO_OP_PRIMARY_FLANK_SUBSTITUTION: return 3;
    case TSAO_OP_PRIMARY_MATCH: case TSAO_OP_PRIMARY_FLANK_MATCH: return 4;
    default: return 0;
    }
}

// Forward-order RLE append; a merged run keeps the label of its last op (a_star_aligner.rs:100-122).
static void push_op(std::vector<tsao_op>& ops, tsao_op op) {
    if (!ops.empty()) {
        tsao_op& b = ops.back();
        bool rep = false;
        if (group_of(op.type) && group_of(op.type) == group_of(b.type)) rep = true;
        else if (op.type == b.type && op.type >= TSAO_OP_SECONDARY_INSERTION && op.type <= TSAO_OP_SECONDARY_MATCH) rep = true;
        if (rep) { b.count += op.count; b.type = op.type; return; }
    }
    ops.push_back(op);
}

}  // namespace

extern "C" int tsao_dp_align(const tsao_config* c, const uint8_t* reference, int64_t n, const uint8_t* query, int64_t m,
                             int64_t ro, int64_t rl, int64_t qo, int64_t ql, const tsao_options* opt, tsao_result* out) {
    Solver S;
    S.cfg.c = c; S.cfg.A = c->alphabet_size;
    S.R = reference; S.n = n; S.Q = query; S.m = m;
    S.ro = ro; S.rl = rl; S.qo = qo; S.ql = ql;
    S.opt = *opt;
    S.LF = c->left_flank_length; S.RF = c->right_flank_length; S.F = (int)(S.LF + S.RF + 1);
    S.cells = (size_t)(n + 1) * (m + 1);
    S.ml = -1; S.Lmax = -1;
    {
        int len = c->fn_len[TSAO_FN_LENGTH];
        for (int i = 0; i < len; i++) if (sat(c->fn_c[TSAO_FN_LENGTH][i]) < INF) {
            if (S.ml < 0) S.ml = c->fn_x[TSAO_FN_LENGTH][i];
            S.Lmax = (i + 1 < len) ? c->fn_x[TSAO_FN_LENGTH][i + 1] - 1 : std::max(n, m);
        }
        S.Lmax = std::min(S.Lmax, std::max(n, m));
    }
    memset(out, 0, sizeof(*out));

    S.layers.emplace_back();
    S.layers[0].seed.assign(S.cells, INF);
    S.layers[0].from.assign(S.cells, SeedFrom{-1, 0, 0, 0});
    S.layers[0].seed[S.cell(ro, qo)] = 0;
    S.fill_layer(S.layers[0], true);
    i64 best = S.layers[0].target; int best_layer = 0;
    for (int k = 0; !opt->no_ts && k < 64; k++) {
        Layer next;
        S.jump(S.layers[k], next, best);
        i64 mn = INF;
        for (i64 v : next.seed) mn = std::min(mn, v);
        if (mn >= best) break;  // costs are non-negative: nothing in later layers can be cheaper
        S.layers.push_back(std::move(next));
        S.fill_layer(S.layers[k + 1], false);
        if (S.layers[k + 1].target < best) { best = S.layers[k + 1].target; best_layer = k + 1; }
    }

    if (best >= INF) { out->result_type = TSAO_NO_TARGET; return 0; }
    if (opt->cost_limit != UINT64_MAX && (u64)best > opt->cost_limit) {
        out->result_type = TSAO_EXCEEDED_COST_LIMIT; out->cost = opt->cost_limit; return 0;
    }
    out->result_type = TSAO_FOUND_TARGET;
    out->cost = (u64)best;

    // ---- traceback (tie-break documented in DESIGN.md: lowest layer; diag before del before ins) ----------
    std::vector<std::vector<tsao_op>> segments;  // built back to front
    int k = best_layer;
    i64 i = rl, j = ql, f = S.layers[k].tf; int g = S.layers[k].tg;
    u64 total_len = 0;
    for (;;) {
        Layer& L = S.layers[k];
        std::vector<tsao_op> rev;  // ops of this layer segment in reverse path order
        for (;;) {
            uint8_t pc = L.pred[S.st(g, f, S.cell(i, j))];
            if (pc == 255) return -2;
            int mv = pc / 3, gp = pc % 3;
            if (mv == MV_SEED) break;
            bool flank = mv & 1;
            int type;
            if (mv == MV_DIAG || mv == MV_DIAG_FLANK) {
                i--; j--;
                bool match = reference[i] == query[j];
                type = flank ? (match ? TSAO_OP_PRIMARY_FLANK_MATCH : TSAO_OP_PRIMARY_FLANK_SUBSTITUTION) : (match ? TSAO_OP_PRIMARY_MATCH : TSAO_OP_PRIMARY_SUBSTITUTION);
            } else if (mv == MV_DEL || mv == MV_DEL_FLANK) {
                i--; type = flank ? TSAO_OP_PRIMARY_FLANK_DELETION : TSAO_OP_PRIMARY_DELETION;
            } else {
                j--; type = flank ? TSAO_OP_PRIMARY_FLANK_INSERTION : TSAO_OP_PRIMARY_INSERTION;
            }
            if (flank) f--;
            g = gp;
            rev.push_back(tsao_op{1, type, 0, 0, 0, 0});
        }
        std::reverse(rev.begin(), rev.end());
        segments.push_back(rev);
        if (k == 0) break;
        // the seed at (i, j) of layer k: re-run its chain to recover offset and inner path
        SeedFrom sf = L.from[S.cell(i, j)];
        if (sf.kind < 0) return -3;
        Solver::Kind kd{(sf.kind >> 1) & 1, sf.kind & 1, sf.kind >> 2, S.cfg.base((sf.kind >> 1) & 1, sf.kind & 1, sf.kind >> 2), sf.kind};
        const uint8_t* P = kd.p == 0 ? reference : query;
        i64 na = kd.p == 0 ? m : n;
        std::vector<uint8_t> T; S.secondary_string(kd, T);
        Solver::Chain ch;
        int table = kd.d == 0 ? TSAO_TAB_SEC_FWD : TSAO_TAB_SEC_REV;
        S.run_chain(ch, P, sf.e, T, table, sf.len);
        std::vector<i64> E, Eoff;
        std::vector<Piece> ocp = S.offset_pieces(kd);
        i64 ip = sf.e - sf.len;
        S.jump_in(kd, ch, sf.len, ip, na, ocp, E, Eoff);
        i64 o = Eoff[sf.entr_anti];
        i64 anti2 = kd.p == 0 ? j : i;
        i64 apg = anti2 - sf.entr_anti;
        i64 ld = apg - sf.len;
        std::vector<tsao_op> ts;
        ts.push_back(tsao_op{kd.d == 1 ? std::llabs(o) + 1 : std::llabs(o), TSAO_OP_TS_ENTRANCE, kd.p, kd.s, kd.d, o});
        // inner path: follow the cost-to-go table forwards
        {
            i64 sE = kd.p == kd.s ? ip : sf.entr_anti;
            i64 b = sE + o, ns = ch.ns;
            i64 y = kd.d == 0 ? b : ns - b;
            i64 l = sf.len; int gs = G_NONE;
            while (l > 0) {
                i64 v = ch.at(l, gs, y);
                int pcx = P[sf.e - l];
                if (y < ns && v == add(S.cfg.sub(table, pcx, T[y]), ch.at(l - 1, G_NONE, y + 1))) {
                    ts.push_back(tsao_op{1, pcx == T[y] ? TSAO_OP_SECONDARY_MATCH : TSAO_OP_SECONDARY_SUBSTITUTION, 0, 0, 0, 0});
                    l--; y++; gs = G_NONE; total_len++;
                } else if (y < ns && v == add(S.cfg.gap(table, T[y], gs != G_DEL), ch.at(l, G_DEL, y + 1))) {
                    ts.push_back(tsao_op{1, TSAO_OP_SECONDARY_DELETION, 0, 0, 0, 0});
                    y++; gs = G_DEL;
                } else if (v == add(S.cfg.gap(table, pcx, gs != G_INS), ch.at(l - 1, G_INS, y))) {
                    ts.push_back(tsao_op{1, TSAO_OP_SECONDARY_INSERTION, 0, 0, 0, 0});
                    l--; gs = G_INS; total_len++;
                } else return -4;
            }
        }
        ts.push_back(tsao_op{std::llabs(ld) + 1, TSAO_OP_TS_EXIT, 0, 0, 0, apg});
        segments.push_back(ts);
        // continue in layer k-1 at the entrance cell, state f = L_f, g = argmin
        if (kd.p == 0) { i = ip; j = sf.entr_anti; } else { i = sf.entr_anti; j = ip; }
        k--;
        f = S.LF; g = S.layers[k].dmin_g[S.cell(i, j)];
    }
    std::vector<tsao_op> ops;
    for (size_t sidx = segments.size(); sidx-- > 0;)
        for (const tsao_op& op : segments[sidx]) {
            if (op.type == TSAO_OP_TS_ENTRANCE || op.type == TSAO_OP_TS_EXIT) ops.push_back(op);
            else push_op(ops, op);
        }
    out->ts_total_length = total_len;
    out->n_ops = (i64)ops.size();
    out->ops = (tsao_op*)malloc(sizeof(tsao_op) * std::max<size_t>(1, ops.size()));
    for (size_t t = 0; t < ops.size(); t++) out->ops[t] = ops[t];
    return 0;
}
