// tsa_config.hpp -- host cost-model layer: alphabets, config.tsa parser, flattening to DevConfig.
//
// Mirrors lib_tsalign's TemplateSwitchConfig (lib_tsalign/src/config.rs:24-49), its plain-text reader
// (config/io.rs:33-111, costs/cost_function/io.rs:81-120, costs/gap_affine/io.rs:156-359) and the V-shape
// validation (config.rs:72-85, costs/cost_function.rs:170-176).
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "tsa_types.hpp"

namespace tsa {

constexpr uint64_t COST_INF = UINT64_MAX;  // U64Cost::max_value()

enum Alphabet { ALPHA_DNA = 0, ALPHA_DNA_N = 1, ALPHA_RNA = 2, ALPHA_RNA_N = 3, ALPHA_DNA_IUPAC = 4, ALPHA_RNA_IUPAC = 5, ALPHA_COUNT = 6 };

const char* alphabet_chars(int alphabet);   // index -> ASCII
const char* alphabet_name(int alphabet);    // "dna", "dna-n", ...
int alphabet_from_name(const std::string& name);  // -1 if unknown
int alphabet_index(int alphabet, unsigned char ascii);  // -1 if the character is not in the alphabet
int alphabet_complement(int alphabet, int index);

struct EditTable {                 // GapAffineAlignmentCostTable (costs/gap_affine.rs:17-33)
    std::string name;
    std::vector<uint64_t> sub;     // [A*A] row-major: (first character, second character)
    std::vector<uint64_t> open, ext;
};

typedef std::vector<std::pair<int64_t, uint64_t>> StepFunction;  // CostFunction (costs/cost_function.rs:22-24)

struct HostConfig {                // TemplateSwitchConfig
    int alphabet = ALPHA_DNA_N;
    int64_t left_flank_length = 0, right_flank_length = 0;
    uint64_t base[8] = {0};        // rrf rqf qrf qqf rrr rqr qrr qqr (config.rs:52-69)
    StepFunction fn[6];            // RQQROffset RRQQOffset Length LengthDifference ForwardAntiPrimaryGap ReverseAntiPrimaryGap
    EditTable table[5];            // primary, secondary forward, secondary reverse, left flank, right flank

    uint64_t evaluate(int k, int64_t x) const;       // CostFunction::evaluate
    int64_t min_length() const;                      // template_switch_min_length (config/io.rs:82-84), -1 if none
};

// Error kinds mirror lib_tsalign/src/error.rs:5-49.
enum ConfigErrorKind { CFG_OK = 0, CFG_PARSE = 1, CFG_RQQR_NOT_V = 2, CFG_RRQQ_NOT_V = 3, CFG_LENDIFF_NOT_V = 4, CFG_ALPHABET = 5 };

// TemplateSwitchConfig::read_plain / from_str.  Returns CFG_OK or an error kind with a message.
int parse_config(const std::string& text, int alphabet, HostConfig& out, std::string& err);
// TemplateSwitchConfig::default() (config.rs:219-303).
HostConfig default_config(int alphabet);
// Display of the config in config.tsa syntax (config/io.rs:113-179), parseable by parse_config.
std::string write_config(const HostConfig& cfg);

// Flatten for the device.  Fails (returns false + message) if the model does not fit the kernels' limits.
bool flatten_config(const HostConfig& cfg, DevConfig& dev, std::vector<int>& lc_dense, std::string& err);

}  // namespace tsa
