// tsa_long.hpp -- host side of tsa_band.cuh: one column band of one long pair without template switches on one device.
//
// world == 1: the whole pair on one GPU under a memory limit (checkpoint rows / boundary columns + recomputed tiles instead of a
// resident code matrix).  world > 1: rank g owns a band of the query columns; the boundary column streams into the next rank's
// memory (NVLink P2P), the traceback is handed from right to left.  SURVEY.md 8(e), BASELINE config 5.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "tsa_config.hpp"
#include "tsa_types.hpp"

namespace tsa {

struct BandWalk {            // mirror of WalkState (tsa_band.cuh) for the callers
    int i = 0, j = 0, g = 0, need = 1;
    long long cost = 0;
    int status = 0;          // 0: left this band to the left (hand to rank - 1), 1: reached the root, < 0: error
    int pad = 0;
};

struct BandStats {
    double forward_ms = 0, trace_ms = 0;         // device time of the forward launch (CUDA events); host wall time of the walks
    long long tiles = 0, tile_cells = 0;         // tiles recomputed by the traceback and their cells
    long long speculated_tiles = 0, speculated_used = 0;   // tiles recomputed ahead of the walk, all at once (speculate()), and how many the walk used
    double speculate_ms = 0;
    long long boundary_bytes_out = 0;            // bytes stored into the next rank's memory by the forward pass (8 per row)
    long long resident_bytes = 0;                // device memory held by this band
    int interval = 0, group = 0;
};

// Geometry of the bands, the same on every rank: groups of `group` strips, dealt out contiguously.
struct BandPlan {
    int nn = 0, mm = 0, world = 1;
    int interval = 0, group = 0;
    int s_total = 0, n_groups = 0;
    bool ok = false;             // false: no (interval, group) fits the memory limit, or fewer strips than ranks
    std::string why;
    int group_first(int rank) const { return (int)((long long)rank * n_groups / world); }
    int strip_first(int rank) const { return group_first(rank) * group; }
    int strip_last(int rank) const { const int s = group_first(rank + 1) * group; return (s < s_total ? s : s_total) - 1; }
    int owner_of_column(int j) const;
    long long resident_bytes(int rank, bool traceback) const;
};
// interval / group: 0 = chosen (4096 rows / 16 strips, or what the memory limit allows); memory_limit: bytes per device, 0 = none.
BandPlan plan_bands(int nn, int mm, int world, int interval, int group, size_t memory_limit, bool traceback);

class LongPair {
public:
    // R / Q: the alignment range already cut out and encoded as alphabet indices (nn / mm characters).
    LongPair(const HostConfig& cfg, int device, const uint8_t* R, const uint8_t* Q, const BandPlan& plan, int rank, bool traceback);
    ~LongPair();
    LongPair(const LongPair&) = delete;
    LongPair& operator=(const LongPair&) = delete;

    bool ok() const { return ok_; }
    const std::string& error() const { return err_; }
    int rank() const { return rank_; }
    int device() const { return device_; }
    int col_first() const { return col_first_; }      // first column of this band (range-relative)

    void* incoming_boundary() const;                  // device buffer the band on the left writes (rank > 0): (nn + 1) 8-byte entries
    size_t boundary_bytes() const { return (size_t)(plan_.nn + 1) * 8; }
    void set_outgoing_boundary(void* remote);         // the next rank's incoming buffer, mapped into this process (rank < world - 1)

    void forward_launch();                            // fill of the band; asynchronous (the bands of all ranks run as one pipeline)
    void forward_wait();
    // Recompute, with codes, the tiles around the straight line from the root to the target that lie in this band -- all of them in ONE
    // launch (every tile only needs the checkpoint row above it and the boundary column to its left, both kept by the forward pass), as
    // many as `budget_bytes` of code buffers allow.  The walk then only recomputes a tile on its own when the path leaves that set:
    // the traceback of a long pair costs one tile latency instead of one per tile crossed.  Asynchronous; walk() waits.
    void speculate(size_t budget_bytes);
    bool has_target() const { return rank_ == plan_.world - 1; }
    long long cost() const { return cost_; }          // has_target() only; >= INF32: no target
    bool saturated() const { return saturated_; }

    // Continue the traceback inside this band from `in` (the first call on the last rank starts at the target with
    // BandWalk{nn, mm, 0, 1, cost}).  Appends unit ops in walk order (reverse path order) and returns where the walk left.
    BandWalk walk(const BandWalk& in, std::vector<uint8_t>& ops_rev);

    const BandStats& stats() const { return stats_; }

private:
    struct Impl;
    Impl* impl_;
    DevConfig dev_;
    std::vector<int> lc_;
    BandPlan plan_;
    int device_, rank_, col_first_ = 0;
    bool traceback_, ok_ = false, saturated_ = false;
    long long cost_ = -1;
    std::string err_;
    BandStats stats_;
};

// All bands of one pair in this process (one host thread, `n_devices` devices with peer access between neighbours).
struct LongResult {
    int status = 0;              // PairStatus
    bool memory_limit_hit = false;
    long long cost = 0;
    std::vector<uint8_t> ops;    // unit ops in path order (empty without traceback)
    std::vector<BandStats> stats;
    BandPlan plan;
    std::string message;
};
LongResult align_long(const HostConfig& cfg, const int* devices, int n_devices, const uint8_t* R, int nn, const uint8_t* Q, int mm,
                      int interval, int group, size_t memory_limit, bool traceback);

}  // namespace tsa
