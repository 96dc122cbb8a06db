// tsa_post.hpp -- host post-processing of found alignments (see tsa_post.cpp).
#pragma once
#include <cstdint>
#include <vector>

#include "tsa_config.hpp"

namespace tsa {

// One run of the run-length encoded alignment (the fields of tsa_op plus the equal-cost range of an entrance).
struct PostOp {
    int64_t count = 1;
    int type = 0;                       // TSA_OP_*
    int primary = 0, secondary = 0, direction = 0;
    int64_t value = 0;                  // entrance: first_offset; exit: anti_primary_gap
    int8_t ecr[4] = {1, -1, 1, -1};     // EqualCostRange::new_invalid(): min_start, max_start, min_end, max_end
    bool ecr_valid = false;
};

uint64_t post_compute_cost(const HostConfig& cfg, const uint8_t* R, int64_t n, const uint8_t* Q, int64_t m, int64_t ro, int64_t qo,
                           const std::vector<PostOp>& ops);
int64_t post_extend_beyond_range(const HostConfig& cfg, const uint8_t* R, int64_t n, const uint8_t* Q, int64_t m, std::vector<PostOp>& ops,
                                 int64_t& ro, int64_t& rl, int64_t& qo, int64_t& ql);
void post_equal_cost_ranges(const HostConfig& cfg, const uint8_t* R, int64_t n, const uint8_t* Q, int64_t m, std::vector<PostOp>& ops, int64_t ro, int64_t qo);
bool post_move_start_backwards(const uint8_t* R, int64_t n, const uint8_t* Q, int64_t m, int alphabet, int64_t ro, int64_t qo, std::vector<PostOp>& ops, size_t& ci);
bool post_move_start_forwards(const uint8_t* R, int64_t n, const uint8_t* Q, int64_t m, int64_t ro, int64_t qo, std::vector<PostOp>& ops, size_t& ci);
bool post_move_end_forwards(const uint8_t* R, int64_t n, const uint8_t* Q, int64_t m, int alphabet, int64_t ro, int64_t qo, std::vector<PostOp>& ops, size_t ci);
bool post_move_end_backwards(const uint8_t* R, int64_t n, const uint8_t* Q, int64_t m, int64_t ro, int64_t qo, std::vector<PostOp>& ops, size_t ci);

}  // namespace tsa
