// tsa_capi.cpp -- the extern "C" boundary declared in include/tsalign_b200.h.
#include "tsalign_b200.h"

#include <chrono>
#include <exception>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "tsa_config.hpp"
#include "tsa_engine.hpp"
#include "tsa_long.hpp"
#include "tsa_post.hpp"
#include "tsa_rt.hpp"

using namespace tsa;

#include <atomic>
#include <mutex>
#include <thread>

// A parsed cost model plus the engines (device buffers, stream) that tsa_align_batch reuses from call to call.
struct tsa_config {
    HostConfig host;
    std::mutex lock[16];                   // one call at a time per (config, device); calls on different devices run concurrently
    std::unique_ptr<Engine> engine[16];
    std::unique_ptr<Engine> engine2[16];   // second engine (own stream and buffers) for the other half of a large batch
};

namespace {

void set_err(char* err, size_t cap, const std::string& msg) {
    if (!err || cap == 0) return;
    size_t n = std::min(cap - 1, msg.size());
    memcpy(err, msg.data(), n);
    err[n] = '\0';
}

struct Encoded {
    std::vector<uint8_t> pool;
    std::vector<PairView> views;
    std::vector<int> pair_status;       // TSA_OK or per-pair input error
    std::vector<std::string> pair_msg;
    std::vector<size_t> live;           // indices of pairs handed to the engine
};

// tsalign/src/align.rs:389-405 (VectorGenome::from_slice_u8) + AlignmentRange checks.
// Result assembly (run-length encoding, post-processing) of a batch on the host cores.
template <class F>
void parallel_for(size_t n, F&& body, size_t threads_hint = 0) {
    // host threads of this process: TSA_B200_THREADS, else the cores divided among the ranks of a torchrun launch
    static const size_t hw = []() -> size_t {
        size_t cores = std::max<size_t>(1, std::thread::hardware_concurrency());
        if (const char* t = getenv("TSA_B200_THREADS")) { const long v = atol(t); if (v > 0) return (size_t)v; }
        if (const char* w = getenv("LOCAL_WORLD_SIZE")) { const long v = atol(w); if (v > 1) cores = std::max<size_t>(1, cores / (size_t)v); }
        return cores;
    }();
    const size_t nt = std::min<size_t>(std::min<size_t>(hw, 32), std::max((n + 255) / 256, std::min(threads_hint, (n + 63) / 64)));
    if (nt <= 1) { for (size_t i = 0; i < n; i++) body(i); return; }
    std::vector<std::thread> th;
    std::atomic<size_t> next(0);
    for (size_t t = 0; t < nt; t++) th.emplace_back([&]() { for (;;) { const size_t lo = next.fetch_add(64); if (lo >= n) break; for (size_t i = lo; i < std::min(n, lo + 64); i++) body(i); } });
    for (auto& x : th) x.join();
}

void encode_pairs(const HostConfig& cfg, const tsa_pair* pairs, size_t n, Encoded& e) {
    int lut[256];
    for (int c = 0; c < 256; c++) lut[c] = alphabet_index(cfg.alphabet, (unsigned char)c);
    size_t total = 0;
    for (size_t i = 0; i < n; i++) total += pairs[i].reference_len + pairs[i].query_len;
    e.pool.resize(total + 1);
    e.pair_status.assign(n, TSA_OK);
    e.pair_msg.assign(n, std::string());
    std::vector<size_t> offs(n + 1, 0);
    for (size_t i = 0; i < n; i++) offs[i + 1] = offs[i] + pairs[i].reference_len + pairs[i].query_len;
    std::vector<PairView> all(n);
    parallel_for(n, [&](size_t i) {   // (every pair writes its own slice of the pool and its own status / message)
        const tsa_pair& p = pairs[i];
        size_t off = offs[i];
        PairView v;
        v.ref = &e.pool[off];
        for (size_t k = 0; k < p.reference_len; k++) {
            int x = lut[(unsigned char)p.reference[k]];
            if (x < 0 && e.pair_status[i] == TSA_OK) { e.pair_status[i] = TSA_ERR_INVALID_CHAR; e.pair_msg[i] = std::string("reference character '") + p.reference[k] + "' is not part of alphabet " + alphabet_name(cfg.alphabet); }
            e.pool[off++] = (uint8_t)(x < 0 ? 0 : x);
        }
        v.qry = &e.pool[off];
        for (size_t k = 0; k < p.query_len; k++) {
            int x = lut[(unsigned char)p.query[k]];
            if (x < 0 && e.pair_status[i] == TSA_OK) { e.pair_status[i] = TSA_ERR_INVALID_CHAR; e.pair_msg[i] = std::string("query character '") + p.query[k] + "' is not part of alphabet " + alphabet_name(cfg.alphabet); }
            e.pool[off++] = (uint8_t)(x < 0 ? 0 : x);
        }
        v.n = (int)p.reference_len; v.m = (int)p.query_len;
        int64_t rl = p.reference_limit < 0 ? (int64_t)p.reference_len : p.reference_limit;
        int64_t ql = p.query_limit < 0 ? (int64_t)p.query_len : p.query_limit;
        if (e.pair_status[i] == TSA_OK && (p.reference_offset < 0 || p.query_offset < 0 || p.reference_offset > rl || p.query_offset > ql ||
                                           rl > (int64_t)p.reference_len || ql > (int64_t)p.query_len)) {
            e.pair_status[i] = TSA_ERR_INVALID_RANGE; e.pair_msg[i] = "alignment range outside the sequences";
        }
        if (e.pair_status[i] == TSA_OK && (p.reference_len > (size_t)1 << 30 || p.query_len > (size_t)1 << 30)) { e.pair_status[i] = TSA_ERR_UNSUPPORTED; e.pair_msg[i] = "sequence too long"; }
        v.ro = (int)p.reference_offset; v.rl = (int)rl; v.qo = (int)p.query_offset; v.ql = (int)ql;
        all[i] = v;
    }, total >> 20);   // (long sequences: one thread per MB rather than per 256 pairs)
    for (size_t i = 0; i < n; i++) if (e.pair_status[i] == TSA_OK) { e.live.push_back(i); e.views.push_back(all[i]); }
}

// Unit ops -> the run-length encoded list the reference emits (a_star_aligner.rs:100-122, alignment_type.rs:101-139).
// Entrance / exit multiplicities are the artefacts of the reference's +-1 walks: |first_offset| + 1
// (reverse) or |first_offset| (forward) for an entrance, |length_difference| + 1 for an exit.
void assemble_ops(tsa_result& r, const PairCost& pc, bool keep_flank_runs) {
    std::vector<tsa_op> out;
    size_t rec = 0;
    for (uint8_t u : pc.ops) {
        if (u == 12 || u == 13) {
            if (rec >= pc.recs.size() + (u == 13 ? 1 : 0) && u == 12) break;
            const TsRecord& t = pc.recs[u == 12 ? rec : rec - 1];
            tsa_op op;
            memset(&op, 0, sizeof(op));
            if (u == 12) {
                const int d = t.kind >> 2;
                const long long o = t.first_offset;
                op.type = TSA_OP_TS_ENTRANCE; op.primary = (t.kind >> 1) & 1; op.secondary = t.kind & 1; op.direction = d;
                op.value = o; op.count = d == 1 ? std::llabs(o) + 1 : std::llabs(o);
                rec++;
            } else {
                const long long ld = (long long)t.anti_primary_gap - t.length;
                op.type = TSA_OP_TS_EXIT; op.value = t.anti_primary_gap; op.count = std::llabs(ld) + 1;
            }
            out.push_back(op);
            continue;
        }
        // alignment_type.rs:101-121: flank and non-flank variants of the same primary operation repeat each other; a
        // merged run carries the label of its last operation (a_star_aligner.rs:100-122 walks the path backwards)
        if (!out.empty()) {
            const int32_t pt = out.back().type;
            const bool same = pt == (int32_t)u || (!keep_flank_runs && pt < 8 && u < 8 && (pt & 3) == (u & 3));
            if (same) { out.back().count++; out.back().type = u; continue; }
        }
        tsa_op op;
        memset(&op, 0, sizeof(op));
        op.type = u; op.count = 1;
        out.push_back(op);
    }
    for (tsa_op& op : out) if (op.type == TSA_OP_TS_ENTRANCE) { op.min_start = 1; op.max_start = -1; op.min_end = 1; op.max_end = -1; }   // EqualCostRange::new_invalid()
    r.n_ops = out.size();
    r.ops = (tsa_op*)malloc(sizeof(tsa_op) * std::max<size_t>(1, out.size()));
    memcpy(r.ops, out.data(), sizeof(tsa_op) * out.size());
}

void fill_result(tsa_result& r, const PairCost& pc, const tsa_options& opt) {
    memset(&r, 0, sizeof(r));
    switch (pc.status) {
    case PAIR_OK:
        r.status = TSA_OK;
        if (opt.cost_limit != UINT64_MAX && (uint64_t)pc.cost > opt.cost_limit) {
            // generic_a_star/src/lib.rs:372-377,654-660: with an exact fill, "f > limit" means "optimum > limit"
            r.result_type = TSA_EXCEEDED_COST_LIMIT; r.cost = opt.cost_limit;
        } else {
            r.result_type = TSA_FOUND_TARGET; r.cost = (uint64_t)pc.cost; r.template_switches = pc.layers;
            if (pc.trace_status == TRACE_OK) assemble_ops(r, pc, (opt.flags & TSA_FLAG_KEEP_FLANK_RUNS) != 0);
            else if (pc.trace_status != TRACE_SKIPPED) { r.status = TSA_ERR_INTERNAL; snprintf(r.message, sizeof(r.message), "traceback failed (code %d)", pc.trace_status); }
        }
        break;
    case PAIR_NO_TARGET: r.status = TSA_OK; r.result_type = TSA_NO_TARGET; break;
    case PAIR_MEMORY_LIMIT: r.status = TSA_OK; r.result_type = TSA_EXCEEDED_MEMORY_LIMIT; snprintf(r.message, sizeof(r.message), "the pair alone needs more resident memory than --memory-limit"); break;
    case PAIR_ERR_TOO_LONG: r.status = TSA_ERR_UNSUPPORTED; snprintf(r.message, sizeof(r.message), "offset / length hulls of the cost model exceed the 1056-column windows"); break;
    case PAIR_ERR_COST_RANGE: r.status = TSA_ERR_UNSUPPORTED; snprintf(r.message, sizeof(r.message), "alignment cost exceeds the kernels' integer range (2^14 with template switches, else 2^26)"); break;
    case PAIR_ERR_LAYER_CAP: r.status = TSA_ERR_UNSUPPORTED; snprintf(r.message, sizeof(r.message), "not proven optimal within max_template_switches template switches: refused"); break;
    case PAIR_ERR_FLANKS: r.status = TSA_ERR_UNSUPPORTED; snprintf(r.message, sizeof(r.message), "flank lengths above 255 are not supported"); break;
    default: r.status = TSA_ERR_ARGUMENT; break;
    }
}

std::vector<PostOp> to_post(const tsa_op* ops, size_t n) {
    std::vector<PostOp> v(n);
    for (size_t i = 0; i < n; i++) {
        v[i].count = ops[i].count; v[i].type = ops[i].type; v[i].primary = ops[i].primary; v[i].secondary = ops[i].secondary;
        v[i].direction = ops[i].direction; v[i].value = ops[i].value;
        v[i].ecr[0] = ops[i].min_start; v[i].ecr[1] = ops[i].max_start; v[i].ecr[2] = ops[i].min_end; v[i].ecr[3] = ops[i].max_end;
    }
    return v;
}
void from_post(const std::vector<PostOp>& v, tsa_op* ops) {
    for (size_t i = 0; i < v.size(); i++) {
        memset(&ops[i], 0, sizeof(tsa_op));
        ops[i].count = v[i].count; ops[i].type = v[i].type; ops[i].primary = v[i].primary; ops[i].secondary = v[i].secondary;
        ops[i].direction = v[i].direction; ops[i].value = v[i].value;
        ops[i].min_start = v[i].ecr[0]; ops[i].max_start = v[i].ecr[1]; ops[i].min_end = v[i].ecr[2]; ops[i].max_end = v[i].ecr[3];
    }
}

// a_star_aligner.rs:238-253 on one found alignment (encoded sequences of the pair)
void postprocess_result(const HostConfig& cfg, const PairView& pv, int32_t flags, tsa_result& r) {
    if (!flags || r.status != TSA_OK || r.result_type != TSA_FOUND_TARGET || !r.ops) return;
    std::vector<PostOp> ops = to_post(r.ops, r.n_ops);
    int64_t ro = r.reference_offset, rl = r.reference_limit, qo = r.query_offset, ql = r.query_limit;
    if (flags & TSA_POST_EXTEND_BEYOND_RANGE) post_extend_beyond_range(cfg, pv.ref, pv.n, pv.qry, pv.m, ops, ro, rl, qo, ql);
    if (flags & TSA_POST_EQUAL_COST_RANGES) post_equal_cost_ranges(cfg, pv.ref, pv.n, pv.qry, pv.m, ops, ro, qo);
    if (ops.size() != r.n_ops) { free(r.ops); r.ops = (tsa_op*)malloc(sizeof(tsa_op) * std::max<size_t>(1, ops.size())); r.n_ops = ops.size(); }
    from_post(ops, r.ops);
    r.reference_offset = ro; r.reference_limit = rl; r.query_offset = qo; r.query_limit = ql;
}

AlignOptions engine_options(const tsa_options& o) {
    AlignOptions a;
    a.no_ts = o.no_ts != 0;
    if (o.max_template_switches > 0) a.max_layers = o.max_template_switches;
    if (o.first_threshold > 0) a.first_threshold = o.first_threshold;
    a.traceback = o.no_traceback == 0;
    a.scout_round = (o.reserved & 1) != 0;   // bit 0 of `reserved`: developer knob, enables the scouting round
    a.no_windows = (o.reserved & 2) != 0;    // bit 1: developer knob, medium pairs skip the column-window stage
    a.test_small_windows = (o.reserved & 4) != 0;   // bit 2: honoured by emulator builds only
    a.cta_fill = (o.reserved & 512) != 0;           // bit 9: developer knob, no grid-pipelined primary fill
    a.pair_major_wave = (o.reserved & 256) != 0;    // bit 8: developer knob, --no-ts strips in (pair, strip) ticket order
    a.fused_windows = (o.reserved & 32) != 0;       // bit 5: developer knob, first window stage through the fused jump kernel
    a.narrow_fill = (o.reserved & 128) != 0;        // bit 7: developer knob, one warp per pair in the primary fill whatever the length
    a.test_tiled = (o.reserved & 64) != 0;          // bit 6: developer knob, pairs wider than 31 run only the tiled window stage
    if (o.reserved & 8) a.wave_checkpoints = 1;     // bit 3: developer knob, --no-ts alignments always through checkpoints (parity tests on short pairs)
    if (o.reserved & 16) a.wave_checkpoints = -1;   // bit 4: developer knob, --no-ts alignments always through the code matrix
    if (a.traceback) a.max_layers = std::min(a.max_layers, (int)MAX_TRACE_LAYERS);
    if (o.memory_limit != UINT64_MAX) { a.chunk_bytes = (size_t)std::max<uint64_t>(o.memory_limit, (uint64_t)1 << 20); a.memory_limit_strict = true; }
    return a;
}

tsa_options default_options() {
    tsa_options o;
    memset(&o, 0, sizeof(o));
    o.cost_limit = UINT64_MAX; o.memory_limit = UINT64_MAX;
    return o;
}

void long_result(tsa_result& r, const PairView& pv, int status, long long cost, std::vector<uint8_t>&& ops_path_order, bool with_ops, const tsa_options& o);

}  // namespace

extern "C" {

tsa_config* tsa_config_parse(const char* text, size_t len, int alphabet, int* status, char* err, size_t errcap) {
    std::unique_ptr<tsa_config> cfg(new tsa_config);
    std::string msg;
    int rc = parse_config(std::string(text ? text : "", text ? len : 0), alphabet, cfg->host, msg);
    static const int map[] = {TSA_OK, TSA_ERR_CONFIG_PARSE, TSA_ERR_NOT_V_SHAPED_RQQR, TSA_ERR_NOT_V_SHAPED_RRQQ, TSA_ERR_NOT_V_SHAPED_LENDIFF, TSA_ERR_ALPHABET};
    if (status) *status = map[rc];
    if (rc != CFG_OK) { set_err(err, errcap, msg); return nullptr; }
    return cfg.release();
}

tsa_config* tsa_config_default(int alphabet) {
    if (alphabet < 0 || alphabet >= ALPHA_COUNT) return nullptr;
    tsa_config* cfg = new tsa_config;
    cfg->host = default_config(alphabet);
    return cfg;
}

size_t tsa_config_write(const tsa_config* cfg, char* out, size_t cap) {
    if (!cfg) return 0;
    std::string s = write_config(cfg->host);
    if (out && cap) { size_t n = std::min(cap - 1, s.size()); memcpy(out, s.data(), n); out[n] = '\0'; }
    return s.size() + 1;
}

void tsa_config_free(tsa_config* cfg) { delete cfg; }
int tsa_config_alphabet(const tsa_config* cfg) { return cfg ? cfg->host.alphabet : -1; }

}  // extern "C"

struct tsa_batch {
    HostConfig host;
    std::unique_ptr<Engine> engine;
    Encoded enc;
    tsa_options opt;
    size_t n = 0;
    std::vector<PairCost> costs;
};

extern "C" {

tsa_batch* tsa_batch_create(const tsa_config* cfg, const tsa_options* opt, const tsa_pair* pairs, size_t n, int* status, char* err, size_t errcap) try {
    int dummy; if (!status) status = &dummy;
    if (!cfg || (!pairs && n)) { *status = TSA_ERR_ARGUMENT; set_err(err, errcap, "null argument"); return nullptr; }
    std::unique_ptr<tsa_batch> b(new tsa_batch);
    b->opt = opt ? *opt : default_options();
    b->n = n;
    b->host = cfg->host;
    b->engine.reset(new Engine(cfg->host, b->opt.device));
    if (!b->engine->ok()) {
        *status = b->engine->error().find("CUDA") != std::string::npos ? TSA_ERR_NO_DEVICE : TSA_ERR_UNSUPPORTED;
        set_err(err, errcap, b->engine->error());
        return nullptr;
    }
    encode_pairs(cfg->host, pairs, n, b->enc);
    AlignOptions ao = engine_options(b->opt);
    ao.chunk_bytes = std::max(ao.chunk_bytes, (size_t)1 << 40);  // staged form: one resident chunk, caller sizes the batch
    if (!b->engine->stage(b->enc.views.data(), b->enc.views.size(), ao)) { *status = TSA_ERR_UNSUPPORTED; set_err(err, errcap, "batch does not fit one chunk"); return nullptr; }
    *status = TSA_OK;
    return b.release();
} catch (const std::exception& e) {
    if (status) *status = TSA_ERR_INTERNAL;
    set_err(err, errcap, e.what());
    return nullptr;
}

int tsa_batch_run(tsa_batch* b) try {
    if (!b) return TSA_ERR_ARGUMENT;
    b->engine->run_staged();
    return TSA_OK;
} catch (const std::exception& e) {
    fprintf(stderr, "tsalign_b200: %s\n", e.what());
    return TSA_ERR_INTERNAL;
}

int tsa_batch_fetch(tsa_batch* b, tsa_result* out) try {
    if (!b || !out) return TSA_ERR_ARGUMENT;
    b->costs.resize(b->enc.views.size());
    b->engine->fetch_staged(b->costs.data());
    for (size_t i = 0; i < b->n; i++) {
        memset(&out[i], 0, sizeof(tsa_result));
        out[i].status = b->enc.pair_status[i];
        snprintf(out[i].message, sizeof(out[i].message), "%s", b->enc.pair_msg[i].c_str());
    }
    parallel_for(b->enc.live.size(), [&](size_t k) {
        tsa_result& r = out[b->enc.live[k]];
        fill_result(r, b->costs[k], b->opt);
        const PairView& pv = b->enc.views[k];
        r.reference_offset = pv.ro; r.reference_limit = pv.rl; r.query_offset = pv.qo; r.query_limit = pv.ql;
        postprocess_result(b->host, pv, b->opt.postprocess, r);
    });
    return TSA_OK;
} catch (const std::exception& e) {
    fprintf(stderr, "tsalign_b200: %s\n", e.what());
    return TSA_ERR_INTERNAL;
}

void tsa_batch_stats(const tsa_batch* b, int64_t* launches, int64_t* jump_launches, int64_t* fill_launches, int32_t* layers, int64_t* h2d_bytes, int64_t* d2h_bytes) {
    if (!b) return;
    const EngineStats& s = b->engine->stats();
    if (launches) *launches = s.launches;
    if (jump_launches) *jump_launches = s.jump_launches;
    if (fill_launches) *fill_launches = s.fill_launches;
    if (layers) *layers = s.layers_run;
    if (h2d_bytes) *h2d_bytes = s.h2d_bytes;
    if (d2h_bytes) *d2h_bytes = s.d2h_bytes;
}

void tsa_batch_work(const tsa_batch* b, int64_t* chains_started, int64_t* chains_run, int64_t* rows_filled, int64_t* rows_jumped) {
    if (!b) return;
    const EngineStats& s = b->engine->stats();
    if (chains_started) *chains_started = s.chains_started;
    if (chains_run) *chains_run = s.chains_run;
    if (rows_filled) *rows_filled = s.rows_filled;
    if (rows_jumped) *rows_jumped = s.rows_jumped;
}

void tsa_batch_timing(const tsa_batch* b, double* jump_ms, double* fill_ms) {
    if (!b) return;
    if (jump_ms) *jump_ms = b->engine->stats().jump_ms;
    if (fill_ms) *fill_ms = b->engine->stats().fill_ms;
}

int tsa_measure_addmin_peak(int device, double* s16x2_lane_ops_per_s, double* s32_lane_ops_per_s) {
    std::string err;
    double a = 0, c = 0;
    if (!measure_addmin_peak(device, &a, &c, err)) return TSA_ERR_NO_DEVICE;
    if (s16x2_lane_ops_per_s) *s16x2_lane_ops_per_s = a;
    if (s32_lane_ops_per_s) *s32_lane_ops_per_s = c;
    return TSA_OK;
}

void tsa_batch_free(tsa_batch* b) { delete b; }

int tsa_align_batch(const tsa_config* cfg, const tsa_options* opt, const tsa_pair* pairs, size_t n, tsa_result* out, char* err, size_t errcap) try {
    if (!cfg || (!pairs && n) || (!out && n)) { set_err(err, errcap, "null argument"); return TSA_ERR_ARGUMENT; }
    const tsa_options o = opt ? *opt : default_options();
    auto t0 = std::chrono::steady_clock::now();
    tsa_config* mcfg = const_cast<tsa_config*>(cfg);
    if (o.device < 0 || o.device >= 16 || o.device >= std::max(1, tsa_device_count())) {
        set_err(err, errcap, tsa_device_count() ? "invalid CUDA device index" : "no CUDA device available: tsalign_b200 has no CPU path");
        return TSA_ERR_NO_DEVICE;
    }
    const int slot = o.device;
    std::lock_guard<std::mutex> guard(mcfg->lock[slot]);   // one call at a time per (config, device)
    if (!mcfg->engine[slot] || !mcfg->engine[slot]->ok()) mcfg->engine[slot].reset(new Engine(cfg->host, o.device));
    Engine& engine = *mcfg->engine[slot];
    if (!engine.ok()) {
        set_err(err, errcap, engine.error());
        const int rc = engine.error().find("CUDA") != std::string::npos ? TSA_ERR_NO_DEVICE : TSA_ERR_UNSUPPORTED;
        mcfg->engine[slot].reset();
        return rc;
    }
    Encoded enc;
    encode_pairs(cfg->host, pairs, n, enc);
    auto t1 = std::chrono::steady_clock::now();
    std::vector<PairCost> costs(enc.views.size());
    const size_t live = enc.views.size();
    for (size_t i = 0; i < n; i++) {
        memset(&out[i], 0, sizeof(tsa_result));
        out[i].status = enc.pair_status[i];
        snprintf(out[i].message, sizeof(out[i].message), "%s", enc.pair_msg[i].c_str());
    }
    // result assembly + post-processing of the pairs [lo, hi) of the encoded list (host cores)
    std::vector<char> done(enc.live.size(), 0);
    auto finish = [&](size_t lo, size_t hi) {
        parallel_for(hi - lo, [&](size_t t) {
            const size_t k = lo + t;
            if (done[k]) return;
            tsa_result& r = out[enc.live[k]];
            fill_result(r, costs[k], o);
            const PairView& pv = enc.views[k];
            r.reference_offset = pv.ro; r.reference_limit = pv.rl; r.query_offset = pv.qo; r.query_limit = pv.ql;
            postprocess_result(cfg->host, pv, o.postprocess, r);
            done[k] = 1;
        });
    };
    const bool pipelined = o.descendant_strategy != 1;      // (the better of two searches is only known after both)
    auto run_engines = [&](const AlignOptions& ao, std::vector<PairCost>& costs) -> int {
#ifndef TSA_EMUL
    // (without template switches: batches of long pairs -- staging, the copy of the operations and the result assembly of one half
    // then overlap the wavefront kernels of the other)
    double job_cells = 0;
    if (o.no_ts) for (size_t k = 0; k < live; k++) job_cells += (double)(enc.views[k].rl - enc.views[k].ro + 1) * (double)(enc.views[k].ql - enc.views[k].qo + 1);
    const bool split = (live >= 4096 && !o.no_ts) || (o.no_ts && live >= 64 && job_cells >= 1.6e10);
#else
    const bool split = false;   // the emulator is single-threaded
#endif
    if (split) {
        // Two halves on two engines (two streams, two sets of buffers) driven by two host threads: staging, result copies and
        // the latency-bound tails of one half overlap the kernels of the other.
        if (!mcfg->engine2[slot] || !mcfg->engine2[slot]->ok()) mcfg->engine2[slot].reset(new Engine(cfg->host, o.device));
        Engine& second = *mcfg->engine2[slot];
        if (!second.ok()) { set_err(err, errcap, second.error()); mcfg->engine2[slot].reset(); return TSA_ERR_NO_DEVICE; }
        // Each half in `parts` pieces: the results of a piece are assembled and post-processed on the host cores right after its
        // kernels, while the kernels of the other engine keep the GPU busy.
        const size_t half = live / 2;
        static const size_t parts_env = []() -> size_t { const char* e = getenv("TSA_B200_PARTS"); const long v = e ? atol(e) : 0; return v > 0 ? (size_t)v : 1; }();   // measured on B200 (16 384 read pairs): 1 / 2 / 4 pieces per half -> 2.05 / 2.00 / 1.84 GCUPS end to end
        const size_t parts = pipelined ? std::max<size_t>(1, std::min(parts_env, half / 2048)) : 1;
        auto run_half = [&](Engine& eng, size_t lo, size_t hi) {
            for (size_t p = 0; p < parts; p++) {
                const size_t a = lo + (hi - lo) * p / parts, b2 = lo + (hi - lo) * (p + 1) / parts;
                eng.align_costs(enc.views.data() + a, b2 - a, ao, costs.data() + a);
                if (pipelined) finish(a, b2);
            }
        };
        std::exception_ptr failed;
        std::thread other([&]() {
            try { run_half(second, half, live); }
            catch (...) { failed = std::current_exception(); }
        });
        try { run_half(engine, 0, half); }
        catch (...) { other.join(); throw; }
        other.join();
        if (failed) std::rethrow_exception(failed);
    } else {
        engine.align_costs(enc.views.data(), live, ao, costs.data());
    }
        return TSA_OK;
    };
    AlignOptions ao = engine_options(o);
    if (o.descendant_strategy == 1 && !o.no_ts) {
        // --ts-descendant-strategy allow-only-all-equal (strategies/descendant.rs:22-104): every template switch of an alignment has
        // the same primary.  As an explicit state dimension that is the better of two searches, one with the kinds whose primary is
        // the reference and one with the kinds whose primary is the query (ties: the reference).
        std::vector<PairCost> other(costs.size());
        ao.primary_filter = 1;
        int rc1 = run_engines(ao, costs);
        if (rc1 != TSA_OK) return rc1;
        ao.primary_filter = 2;
        rc1 = run_engines(ao, other);
        if (rc1 != TSA_OK) return rc1;
        for (size_t k = 0; k < costs.size(); k++) {
            const auto is_err = [](const PairCost& p) { return p.status != PAIR_OK && p.status != PAIR_NO_TARGET; };
            if (is_err(costs[k])) continue;                                       // a refusal of either search: the pair's answer is not proven
            if (is_err(other[k])) { costs[k] = std::move(other[k]); continue; }
            const bool a_ok = costs[k].status == PAIR_OK, b_ok = other[k].status == PAIR_OK;
            if (b_ok && (!a_ok || other[k].cost < costs[k].cost)) costs[k] = std::move(other[k]);
        }
    } else {
        const int rc1 = run_engines(ao, costs);
        if (rc1 != TSA_OK) return rc1;
    }
    auto t2 = std::chrono::steady_clock::now();
    // --memory-limit without template switches: a pair whose code matrix does not fit is aligned by the checkpointed path of
    // tsa_long.cu (checkpoint rows + recomputed tiles under the same limit): the limit bounds what is resident, the alignment is
    // still produced; ExceededMemoryLimit only if not even the checkpoints fit.
    if (o.no_ts) for (size_t k = 0; k < enc.live.size(); k++) if (costs[k].status == PAIR_MEMORY_LIMIT) {
        const PairView& pv = enc.views[k];
        tsa_result& r = out[enc.live[k]];
        const int dev = o.device;
        LongResult lr = align_long(cfg->host, &dev, 1, pv.ref + pv.ro, pv.rl - pv.ro, pv.qry + pv.qo, pv.ql - pv.qo, 0, 0, (size_t)o.memory_limit, o.no_traceback == 0);
        if (lr.memory_limit_hit) {
            r.status = TSA_OK; r.result_type = TSA_EXCEEDED_MEMORY_LIMIT; r.cost = 0;
            snprintf(r.message, sizeof(r.message), "%s", lr.message.c_str());
            r.reference_offset = pv.ro; r.reference_limit = pv.rl; r.query_offset = pv.qo; r.query_limit = pv.ql;
        } else {
            long_result(r, pv, lr.status, lr.cost, std::move(lr.ops), o.no_traceback == 0, o);
            postprocess_result(cfg->host, pv, o.postprocess, r);
        }
        done[k] = 1;
    }
    finish(0, enc.live.size());
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (getenv("TSA_B200_DEBUG"))
        fprintf(stderr, "[tsalign_b200] align_batch n=%zu: encode %.2f ms, engine %.2f ms, results %.2f ms\n", n,
                1e3 * std::chrono::duration<double>(t1 - t0).count(), 1e3 * std::chrono::duration<double>(t2 - t1).count(),
                1e3 * (secs - std::chrono::duration<double>(t2 - t0).count()));
    for (size_t i = 0; i < n; i++) out[i].duration_seconds = n ? secs / (double)n : 0.0;
    return TSA_OK;
} catch (const std::exception& e) {
    // e.g. cudaMalloc failure: drop the cached engine (its buffers) and report; nothing is thrown across the ABI
    if (cfg && opt && opt->device >= 0 && opt->device < 16) {
        tsa_config* mcfg = const_cast<tsa_config*>(cfg);
        std::lock_guard<std::mutex> guard(mcfg->lock[opt->device]);
        mcfg->engine[opt->device].reset(); mcfg->engine2[opt->device].reset();
    }
    set_err(err, errcap, e.what());
    return TSA_ERR_INTERNAL;
}

void tsa_results_free(tsa_result* results, size_t n) {
    if (!results) return;
    for (size_t i = 0; i < n; i++) { free(results[i].ops); results[i].ops = nullptr; results[i].n_ops = 0; }
}

int tsa_postprocess(const tsa_config* cfg, const tsa_pair* pair, int32_t postprocess, tsa_op* ops, size_t* n_ops, size_t cap,
                    int64_t* reference_offset, int64_t* reference_limit, int64_t* query_offset, int64_t* query_limit, uint64_t* cost) try {
    if (!cfg || !pair || !ops || !n_ops || !reference_offset || !reference_limit || !query_offset || !query_limit) return TSA_ERR_ARGUMENT;
    Encoded enc;
    encode_pairs(cfg->host, pair, 1, enc);
    if (enc.views.empty()) return TSA_ERR_ARGUMENT;
    const PairView& pv = enc.views[0];
    std::vector<PostOp> v = to_post(ops, *n_ops);
    int64_t ro = *reference_offset, rl = *reference_limit, qo = *query_offset, ql = *query_limit;
    if (postprocess & TSA_POST_EXTEND_BEYOND_RANGE) post_extend_beyond_range(cfg->host, pv.ref, pv.n, pv.qry, pv.m, v, ro, rl, qo, ql);
    if (postprocess & TSA_POST_EQUAL_COST_RANGES) post_equal_cost_ranges(cfg->host, pv.ref, pv.n, pv.qry, pv.m, v, ro, qo);
    if (cost) *cost = post_compute_cost(cfg->host, pv.ref, pv.n, pv.qry, pv.m, ro, qo, v);
    if (v.size() > cap) return TSA_ERR_ARGUMENT;
    from_post(v, ops);
    *n_ops = v.size();
    *reference_offset = ro; *reference_limit = rl; *query_offset = qo; *query_limit = ql;
    return TSA_OK;
} catch (const std::exception&) {
    return TSA_ERR_INTERNAL;
}

int tsa_post_move(const tsa_config* cfg, const tsa_pair* pair, int which, tsa_op* ops, size_t* n_ops, size_t cap,
                  int64_t reference_offset, int64_t query_offset, size_t* compact_index, uint64_t* cost) try {
    if (!cfg || !pair || !ops || !n_ops || !compact_index || which < 0 || which > 3) return -TSA_ERR_ARGUMENT;
    Encoded enc;
    encode_pairs(cfg->host, pair, 1, enc);
    if (enc.views.empty()) return -TSA_ERR_ARGUMENT;
    const PairView& pv = enc.views[0];
    std::vector<PostOp> v = to_post(ops, *n_ops);
    size_t ci = *compact_index;
    bool ok = false;
    switch (which) {
    case 0: ok = post_move_start_backwards(pv.ref, pv.n, pv.qry, pv.m, cfg->host.alphabet, reference_offset, query_offset, v, ci); break;
    case 1: ok = post_move_start_forwards(pv.ref, pv.n, pv.qry, pv.m, reference_offset, query_offset, v, ci); break;
    case 2: ok = post_move_end_backwards(pv.ref, pv.n, pv.qry, pv.m, reference_offset, query_offset, v, ci); break;
    default: ok = post_move_end_forwards(pv.ref, pv.n, pv.qry, pv.m, cfg->host.alphabet, reference_offset, query_offset, v, ci); break;
    }
    if (cost) *cost = post_compute_cost(cfg->host, pv.ref, pv.n, pv.qry, pv.m, reference_offset, query_offset, v);
    if (v.size() > cap) return -TSA_ERR_ARGUMENT;
    from_post(v, ops);
    *n_ops = v.size();
    *compact_index = ci;
    return ok ? 1 : 0;
} catch (const std::exception&) {
    return -TSA_ERR_INTERNAL;
}

int tsa_device_count(void) {
#ifndef TSA_EMUL
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) return 0;
    return count;
#else
    return 1;
#endif
}

const char* tsa_version(void) {
#ifndef TSA_EMUL
    return "tsalign_b200 0.1.0 (sm_100a)";
#else
    return "tsalign_b200 0.1.0 (SIMT emulator, tests only)";
#endif
}

}  // extern "C"

// ---- one long pair without template switches, column-banded (tsa_long.hpp) ------------------------------------------------------
namespace {

// Encodes the pair and cuts out its alignment range.  Returns TSA_OK or the per-pair input error.
int encode_long(const HostConfig& host, const tsa_pair* pair, Encoded& enc, std::string& msg) {
    encode_pairs(host, pair, 1, enc);
    if (enc.pair_status[0] != TSA_OK) { msg = enc.pair_msg[0]; return enc.pair_status[0]; }
    return TSA_OK;
}

void long_stats_out(const BandStats& b, tsa_long_stats& o) {
    o.forward_ms = b.forward_ms; o.trace_ms = b.trace_ms; o.tiles = b.tiles; o.tile_cells = b.tile_cells;
    o.boundary_bytes_out = b.boundary_bytes_out; o.resident_bytes = b.resident_bytes; o.interval = b.interval; o.group = b.group;
    o.speculated_tiles = b.speculated_tiles; o.speculated_used = b.speculated_used; o.speculate_ms = b.speculate_ms;
}

// unit ops in path order -> tsa_result (run-length encoded), range of the pair
void long_result(tsa_result& r, const PairView& pv, int status, long long cost, std::vector<uint8_t>&& ops_path_order, bool with_ops, const tsa_options& o) {
    PairCost pc;
    pc.status = status; pc.cost = cost; pc.layers = 0;
    pc.trace_status = with_ops ? TRACE_OK : TRACE_SKIPPED;
    pc.ops = std::move(ops_path_order);
    fill_result(r, pc, o);
    r.reference_offset = pv.ro; r.reference_limit = pv.rl; r.query_offset = pv.qo; r.query_limit = pv.ql;
}

}  // namespace

struct tsa_long {
    HostConfig host;
    Encoded enc;
    tsa_options opt;
    BandPlan plan;
    std::unique_ptr<LongPair> band;
    void* mapped = nullptr;    // the next rank's boundary buffer (cudaIpcOpenMemHandle)
};

extern "C" {

int tsa_align_long(const tsa_config* cfg, const tsa_options* opt, const tsa_pair* pair, const int32_t* devices, int32_t n_devices,
                   int32_t interval, int32_t group, tsa_result* out, tsa_long_stats* stats, char* err, size_t errcap) try {
    if (!cfg || !pair || !out || n_devices < 0 || n_devices > 64) { set_err(err, errcap, "null or invalid argument"); return TSA_ERR_ARGUMENT; }
    const tsa_options o = opt ? *opt : default_options();
    if (!o.no_ts) { set_err(err, errcap, "tsa_align_long aligns without template switches: set no_ts"); return TSA_ERR_UNSUPPORTED; }
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<int> devs;
    for (int k = 0; k < n_devices; k++) devs.push_back(devices ? devices[k] : k);
    if (devs.empty()) devs.push_back(o.device);
    const int count = tsa_device_count();
    for (int d : devs) if (d < 0 || d >= std::max(1, count) || count == 0) {
        set_err(err, errcap, count ? "invalid CUDA device index" : "no CUDA device available: tsalign_b200 has no CPU path");
        return TSA_ERR_NO_DEVICE;
    }
    memset(out, 0, sizeof(*out));
    Encoded enc;
    std::string msg;
    const int in_rc = encode_long(cfg->host, pair, enc, msg);
    if (in_rc != TSA_OK) { out->status = in_rc; snprintf(out->message, sizeof(out->message), "%s", msg.c_str()); return TSA_OK; }
    const PairView& pv = enc.views[0];
    const int nn = pv.rl - pv.ro, mm = pv.ql - pv.qo;
    const size_t limit = o.memory_limit == UINT64_MAX ? 0 : (size_t)o.memory_limit;
    LongResult lr = align_long(cfg->host, devs.data(), (int)devs.size(), pv.ref + pv.ro, nn, pv.qry + pv.qo, mm, interval, group, limit, o.no_traceback == 0);
    if (lr.memory_limit_hit) {
        out->status = TSA_OK; out->result_type = TSA_EXCEEDED_MEMORY_LIMIT; out->cost = 0;
        snprintf(out->message, sizeof(out->message), "%s", lr.message.c_str());
        out->reference_offset = pv.ro; out->reference_limit = pv.rl; out->query_offset = pv.qo; out->query_limit = pv.ql;
    } else {
        long_result(*out, pv, lr.status, lr.cost, std::move(lr.ops), o.no_traceback == 0, o);
        postprocess_result(cfg->host, pv, o.postprocess, *out);
    }
    if (stats) for (size_t k = 0; k < lr.stats.size() && k < (size_t)n_devices; k++) long_stats_out(lr.stats[k], stats[k]);
    out->duration_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return TSA_OK;
} catch (const std::exception& e) {
    set_err(err, errcap, e.what());
    return TSA_ERR_INTERNAL;
}

tsa_long* tsa_long_create(const tsa_config* cfg, const tsa_options* opt, const tsa_pair* pair, int32_t rank, int32_t world,
                          int32_t interval, int32_t group, int* status, char* err, size_t errcap) try {
    int dummy; if (!status) status = &dummy;
    if (!cfg || !pair || rank < 0 || world < 1 || rank >= world) { *status = TSA_ERR_ARGUMENT; set_err(err, errcap, "null or invalid argument"); return nullptr; }
    std::unique_ptr<tsa_long> b(new tsa_long);
    b->opt = opt ? *opt : default_options();
    b->host = cfg->host;
    if (!b->opt.no_ts) { *status = TSA_ERR_UNSUPPORTED; set_err(err, errcap, "column bands align without template switches: set no_ts"); return nullptr; }
    std::string msg;
    const int in_rc = encode_long(cfg->host, pair, b->enc, msg);
    if (in_rc != TSA_OK) { *status = in_rc; set_err(err, errcap, msg); return nullptr; }
    const PairView& pv = b->enc.views[0];
    const int nn = pv.rl - pv.ro, mm = pv.ql - pv.qo;
    const size_t limit = b->opt.memory_limit == UINT64_MAX ? 0 : (size_t)b->opt.memory_limit;
    b->plan = plan_bands(nn, mm, world, interval, group, limit, b->opt.no_traceback == 0);
    if (!b->plan.ok) { *status = TSA_ERR_UNSUPPORTED; set_err(err, errcap, b->plan.why); return nullptr; }
    b->band.reset(new LongPair(cfg->host, b->opt.device, pv.ref + pv.ro, pv.qry + pv.qo, b->plan, rank, b->opt.no_traceback == 0));
    if (!b->band->ok()) {
        *status = b->band->error().find("CUDA") != std::string::npos ? TSA_ERR_NO_DEVICE : TSA_ERR_UNSUPPORTED;
        set_err(err, errcap, b->band->error());
        return nullptr;
    }
    *status = TSA_OK;
    return b.release();
} catch (const std::exception& e) {
    if (status) *status = TSA_ERR_INTERNAL;
    set_err(err, errcap, e.what());
    return nullptr;
}

int tsa_long_ipc_export(const tsa_long* b, void* handle64) {
    if (!b || !handle64) return TSA_ERR_ARGUMENT;
#ifndef TSA_EMUL
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaIpcMemHandle_t h;
    if (cudaSetDevice(b->band->device()) != cudaSuccess || cudaIpcGetMemHandle(&h, b->band->incoming_boundary()) != cudaSuccess) { cudaGetLastError(); return TSA_ERR_INTERNAL; }
    memcpy(handle64, &h, 64);
#else
    void* p = b->band->incoming_boundary();   // one process: the "handle" is the pointer
    memset(handle64, 0, 64); memcpy(handle64, &p, sizeof(p));
#endif
    return TSA_OK;
}

int tsa_long_ipc_connect(tsa_long* b, const void* handle64) {
    if (!b || !handle64) return TSA_ERR_ARGUMENT;
#ifndef TSA_EMUL
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    if (cudaSetDevice(b->band->device()) != cudaSuccess || cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return TSA_ERR_INTERNAL; }
    b->mapped = p;
#else
    void* p = nullptr; memcpy(&p, handle64, sizeof(p));
#endif
    b->band->set_outgoing_boundary(p);
    return TSA_OK;
}

int tsa_long_forward(tsa_long* b) try {
    if (!b) return TSA_ERR_ARGUMENT;
    b->band->forward_launch();
    b->band->forward_wait();
    if (b->opt.no_traceback == 0) {
        // tiles around the diagonal recomputed ahead of the walk, under what the memory limit leaves (tsa_long.hpp: speculate)
        size_t budget = (size_t)16 << 30;
        if (b->opt.memory_limit != UINT64_MAX) {
            const long long used = b->plan.resident_bytes(b->band->rank(), true);
            budget = (long long)b->opt.memory_limit > used ? std::min<size_t>(budget, (size_t)((long long)b->opt.memory_limit - used)) : 0;
        }
        b->band->speculate(budget);
    }
    return TSA_OK;
} catch (const std::exception& e) {
    fprintf(stderr, "tsalign_b200: %s\n", e.what());
    return TSA_ERR_INTERNAL;
}

int tsa_long_cost(const tsa_long* b, uint64_t* cost, int32_t* result_type) {
    if (!b || !b->band->has_target()) return TSA_ERR_ARGUMENT;
    const long long c = b->band->cost();
    if (result_type) *result_type = c >= INF32 ? TSA_NO_TARGET : TSA_FOUND_TARGET;
    if (cost) *cost = (uint64_t)c;
    if (c < INF32 && b->band->saturated() && c >= (1 << 26) - 1) return TSA_ERR_UNSUPPORTED;
    return TSA_OK;
}

int tsa_long_owner(const tsa_long* b, int64_t column) { return (!b || column < 0) ? -1 : b->plan.owner_of_column((int)column); }

int tsa_long_walk(tsa_long* b, tsa_long_walk_state* state, uint8_t* ops, size_t cap, size_t* n_ops) try {
    if (!b || !state || !n_ops) return TSA_ERR_ARGUMENT;
    BandWalk in;
    in.i = state->i; in.j = state->j; in.g = state->g; in.need = state->need; in.cost = state->cost; in.status = 0;
    std::vector<uint8_t> rev;
    const BandWalk st = b->band->walk(in, rev);
    state->i = st.i; state->j = st.j; state->g = st.g; state->need = st.need; state->cost = st.cost; state->status = st.status;
    *n_ops = rev.size();
    if (rev.size() > cap || (!ops && !rev.empty())) return TSA_ERR_ARGUMENT;
    if (!rev.empty()) memcpy(ops, rev.data(), rev.size());
    return TSA_OK;
} catch (const std::exception& e) {
    fprintf(stderr, "tsalign_b200: %s\n", e.what());
    return TSA_ERR_INTERNAL;
}

int tsa_long_result(const tsa_long* b, uint64_t cost, const uint8_t* ops_walk_order, size_t n_ops, tsa_result* out) try {
    if (!b || !out || (!ops_walk_order && n_ops)) return TSA_ERR_ARGUMENT;
    std::vector<uint8_t> path(n_ops);
    for (size_t k = 0; k < n_ops; k++) path[k] = ops_walk_order[n_ops - 1 - k];
    const PairView& pv = b->enc.views[0];
    long_result(*out, pv, PAIR_OK, (long long)cost, std::move(path), true, b->opt);
    postprocess_result(b->host, pv, b->opt.postprocess, *out);
    return TSA_OK;
} catch (const std::exception&) {
    return TSA_ERR_INTERNAL;
}

void tsa_long_get_stats(const tsa_long* b, tsa_long_stats* stats) { if (b && stats) long_stats_out(b->band->stats(), *stats); }
void tsa_long_dims(const tsa_long* b, int64_t* rows, int64_t* columns) {
    if (!b) return;
    if (rows) *rows = b->plan.nn;
    if (columns) *columns = b->plan.mm;
}

void tsa_long_free(tsa_long* b) {
    if (!b) return;
#ifndef TSA_EMUL
    if (b->mapped) { cudaSetDevice(b->band->device()); cudaIpcCloseMemHandle(b->mapped); }
#endif
    delete b;
}

}  // extern "C"
