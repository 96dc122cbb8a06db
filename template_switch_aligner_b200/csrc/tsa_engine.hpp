// tsa_engine.hpp -- host side of the hot path: batches of independent pairs -> chunks resident in HBM ->
// layer loop (primary fill, TS jump) -> costs (and alignments).  One Engine per (config, device).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "tsa_config.hpp"
#include "tsa_types.hpp"

namespace tsa {

struct AlignOptions {
    bool no_ts = false;            // --no-ts: MaxTemplateSwitchCount(0), strategies/template_switch_count.rs:41-63
    int max_layers = 64;           // cap on the number of template switches per alignment
    bool traceback = true;         // produce alignments (ops), not only costs
    bool scout_round = false;      // optional first round with the reverse kinds only; measured slower on read pairs (profiles/)
    int first_threshold = 12;      // first pruning threshold of the iterative deepening (doubles per round)
    bool no_windows = false;       // developer knob: pairs of 545..1055 characters skip the column-window stage (parity tests)
    int wave_checkpoints = 0;      // --no-ts with alignments: 0 = checkpoint rows + recomputed tiles for chunks of long pairs, codes of the whole
                                   // matrix otherwise; 1 = always checkpoints; -1 = never (developer knob, parity tests)
    bool cta_fill = false;         // developer knob: very long pairs keep the one-CTA-per-pair primary fill (timing)
    bool pair_major_wave = false;  // developer knob: --no-ts strips handed out in (pair, strip) order instead of strip-major (timing)
    bool fused_windows = false;    // developer knob: the first window stage runs the fused jump kernel instead of row queue + evaluation (parity tests, timing)
    bool narrow_fill = false;      // developer knob: the primary fill of long pairs stays one warp per pair (parity tests, timing)
    bool test_tiled = false;       // developer knob: pairs wider than 31 run only the tiled window stage with narrow sub-ranges (parity tests)
    bool test_small_windows = false;  // emulator builds only: pairs wider than 48 use 96 / 160-column windows (CPU tests of the window logic)
    size_t chunk_bytes = 0;        // HBM budget of one resident chunk of pairs; 0 = half of the free device memory
    int primary_filter = 0;        // 0: every kind; 1 / 2: only template switches whose primary (descendant) is the reference / the query
    bool memory_limit_strict = false;   // --memory-limit given: a pair that does not fit the budget on its own is reported (PAIR_MEMORY_LIMIT), not staged
};

// One pair, already encoded as alphabet indices.
struct PairView {
    const uint8_t* ref; int n;
    const uint8_t* qry; int m;
    int ro, rl, qo, ql;            // AlignmentRange (offset / limit)
};

enum PairStatus {
    PAIR_OK = 0,
    PAIR_NO_TARGET = 1,            // AStarResult::NoTarget
    PAIR_ERR_TOO_LONG = 2,         // TS-enabled pair longer than the jump kernel's widest instantiation
    PAIR_ERR_COST_RANGE = 3,       // costs do not fit the packed s16 lanes of the jump kernel
    PAIR_ERR_LAYER_CAP = 4,        // still improving after max_layers template switches
    PAIR_MEMORY_LIMIT = 6,         // AStarResult::ExceededMemoryLimit: the pair alone needs more than --memory-limit resident
    PAIR_ERR_FLANKS = 5,           // flank lengths > 0 with TS enabled: not built yet
};

struct PairCost {
    int status = PAIR_OK;
    int64_t cost = 0;              // optimal cost (status == PAIR_OK)
    int layers = 0;                // number of template switches on the optimal path found (first layer reaching cost)
    int trace_status = TRACE_SKIPPED;
    std::vector<uint8_t> ops;      // unit ops in path order (TraceOut encoding), when traceback was requested
    std::vector<TsRecord> recs;    // template switches in path order
};

struct EngineStats {
    long long launches = 0;        // kernels launched by the last run()
    long long fill_launches = 0, jump_launches = 0, trace_launches = 0;
    int layers_run = 0;            // jump rounds of the last run (max over chunks)
    int rounds_run = 0;            // deepening rounds of the last run
    double jump_ms = 0, fill_ms = 0;  // device time of the two kernel families (CUDA events; 0 in the emulator)
    long long h2d_bytes = 0, d2h_bytes = 0;
    long long chains_started = 0, chains_run = 0, rows_filled = 0, rows_jumped = 0;  // jump-kernel work of the last run
};

class Engine {
public:
    Engine(const HostConfig& cfg, int device);
    ~Engine();
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;

    bool ok() const { return ok_; }
    const std::string& error() const { return err_; }
    const DevConfig& dev_config() const { return dev_; }

    // Whole job: chunk, upload, run, fetch.  out[i] matches pairs[i].
    void align_costs(const PairView* pairs, size_t n, const AlignOptions& opt, PairCost* out);

    // Split form used by the benchmark: make one chunk resident, run it any number of times, fetch.
    // (All pairs must fit one chunk; returns false otherwise.)
    bool stage(const PairView* pairs, size_t n, const AlignOptions& opt);
    void run_staged();
    void fetch_staged(PairCost* out);
    void run_trace();
    void run_wave();

    const EngineStats& stats() const { return stats_; }
    static size_t bytes_per_pair(int n, int m);

private:
    struct Impl;
    Impl* impl_;
    HostConfig host_;
    DevConfig dev_;
    std::vector<int> lc_;
    bool ok_ = false;
    std::string err_;
    EngineStats stats_;
};

// Measured issue rate of the DPX add-min instructions (lanes per second), the denominator of the integer roofline.
bool measure_addmin_peak(int device, double* s16x2_lane_ops_per_s, double* s32_lane_ops_per_s, std::string& err);

}  // namespace tsa
