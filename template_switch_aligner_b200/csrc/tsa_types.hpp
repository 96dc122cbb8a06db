// tsa_types.hpp -- plain structs shared by the host engine and the kernels.
#pragma once
#include <cstdint>

namespace tsa {

constexpr int INF16 = 0x3FFF;            // "infinite" in the packed s16 lanes: INF16 + INF16 still fits in s16
constexpr uint32_t INF16X2 = 0x3FFF3FFFu;
constexpr int INF32 = 0x3FFFFFFF;        // "infinite" in s32 lanes: INF32 + INF32 still fits in s32
constexpr int MAX_ALPHABET = 16;
constexpr int MAX_PIECES = 64;           // finite constant pieces per step function
constexpr int MAX_KINDS = 8;

// One finite constant piece of a step function: cost on [lo, hi].
struct Piece { int lo, hi, cost; };

// One template-switch kind (primary, secondary, direction) with finite base cost (config.rs:145-194).
struct KindDesc {
    int p, s, d;           // 0 = reference / forward, 1 = query / reverse
    int base;              // base cost
    int table;             // 1 = secondary forward, 2 = secondary reverse edit table
    int n_oc;              // effective first-offset cost pieces (forward quirk folded in)
    Piece oc[MAX_PIECES];
    int n_apg;             // anti-primary-gap pieces of this direction
    Piece apg[MAX_PIECES];
    int min_rest_nolc;     // min over finite (oc + ldc + apg): lower bound of everything but base, length and inner
    int min_rest;          // min_rest_nolc + cheapest finite length cost
};

// Flattened TemplateSwitchConfig (config.rs:24-49) in device memory.
struct DevConfig {
    int A;                                   // alphabet size
    int sub[5][MAX_ALPHABET * MAX_ALPHABET]; // [table][row = first char][col = second char], INF32 = inf
    int open[5][MAX_ALPHABET];
    int ext[5][MAX_ALPHABET];
    uint8_t comp[MAX_ALPHABET];              // complement by alphabet index
    int left_flank, right_flank;
    int ml;                                  // template_switch_min_length, -1: no finite Length cost
    int lmax;                                // largest length with finite cost (INT32_MAX/2 if unbounded)
    int n_lc;                                // dense length costs lc[0 .. n_lc) (INF32 = inf); beyond: lc_tail
    int lc_tail;
    int n_ld;
    Piece ld[MAX_PIECES];                    // LengthDifference pieces
    int n_kinds;
    KindDesc kinds[MAX_KINDS];
    int min_ts;                              // lower bound on the cost of any template switch
};

struct PairMeta {
    int n, m;                // |reference|, |query|
    int ro, rl, qo, ql;      // alignment range (offset / limit), a_star_aligner/alignment_geometry.rs
    long long seq_r, seq_q;  // byte offsets of the encoded sequences in the sequence pool
    long long mat;           // cell offset of this pair's (n+1) x (m+1) matrices in the matrix pools
    long long vec;           // int offset of this pair's row/col minima: rowmin[n+1] then colmin[m+1]
    long long scr;           // int offset of the column-tiling scratch of the primary fill: 3 * (n+1)
};

// Everything a kernel needs about the resident chunk of pairs.
struct Chunk {
    const PairMeta* pairs;
    const uint8_t* seq;      // alphabet indices
    const DevConfig* cfg;
    const int* lc;           // dense length costs (cfg->n_lc entries)
    int16_t* D;              // [pair][i][j]  min_g cost of layer k at flank == L_f, clamped to INF16
    int16_t* DT;             // [pair][j][i]  the same, transposed (primary = query kinds read rows of it)
    int* seedA;              // [pair][i][j]  reentry seeds of layer k+1 written by primary = reference kinds
    int* seedB;              // [pair][j][i]  reentry seeds written by primary = query kinds
    int* minvec;             // rowmin / colmin of D
    int* scratch;            // primary-fill column tiling
    int* best;               // [pair] best target cost over the layers filled so far
    int* best_layer;         // [pair] first layer that reached `best`
    int* active;             // [pair] layer k has seeds below best (fill it, then jump from it)
    int* next_active;        // [pair] set by the jump kernel when it writes a seed below best
    int* counters;           // [0] number of pairs with next_active
};

}  // namespace tsa
