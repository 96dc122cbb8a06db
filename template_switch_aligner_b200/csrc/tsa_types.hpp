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
    int min_open;          // cheapest gap-open cost of this kind's secondary edit table
    int oc_lo, oc_hi, oc_skip0; // hull of the reachable first offsets; oc_skip0: offset 0 is not reachable (forward kinds)
    int n_apg;             // anti-primary-gap pieces of this direction
    Piece apg[MAX_PIECES];
    int apg_nonpos;        // every finite anti-primary gap is <= 0: a switch never reenters right of its entrance
    int min_rest_nolc;     // min over finite (oc + ldc + apg): lower bound of everything but base, length and inner
    int min_rest;          // min_rest_nolc + cheapest finite length cost
    int min_ext;           // cheapest gap-extend cost of this kind's secondary edit table (column windows: bounds deletion runs)
    int apg_lo, apg_hi;    // hull of the finite anti-primary-gap pieces
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
    int ld_lo, ld_hi;                        // their hull
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
    long long tab;           // byte offset of this pair's per-column cost tables (k_prepare_tables), -1 if none
    int lw;                  // columns of one table row (32 * C of the pair's jump-kernel class; windowed pairs: the whole row, padded to 8)
    int win;                 // 1: the jump / traceback kernels run on column windows of this pair (k_ts_jump<C, true>)
    int cpf_words;           // words of one chain-pair bitmap (below)
    long long prog;          // int offset of this pair's progress counters in Chunk::fill_prog (grid-pipelined primary fill), one per column block
    long long cpf;           // int offset of this pair's two chain-pair bitmaps in Chunk::cpflag (bit kind * n_ep + chain pair: the windows of
                             // that chain pair overflowed the first / the second window stage of this layer), -1: none (not the long class)
};

// Traceback code of one cell of one layer (written by k_primary_fill, read by k_traceback).
constexpr int DIR_N_DIAG = 1;     // N state came from the diagonal move (else: reentry seed / root)
constexpr int DIR_DL_EXT = 2;     // Dl state extended a deletion (else: opened from N or I of the cell above)
constexpr int DIR_I_EXT = 4;      // I state extended an insertion (else: opened from N or Dl of the cell to the left)
constexpr int DIR_M_SHIFT = 3;    // bits 3-4: cheapest state of the cell: 0 = N, 1 = Dl, 2 = I (ties in that order)
constexpr int DIR_NI_IS_I = 32;   // min(N, I) is I
constexpr int DIR_ND_IS_DL = 64;  // min(N, Dl) is Dl
constexpr int DIR2_DL_SEED = 1;   // plane 0, flank mode: Dl state taken from the right-flank arrivals
constexpr int DIR2_I_SEED = 2;    //                       I state taken from the right-flank arrivals
// Flank planes: one byte per cell and plane: predecessor state (0 N, 1 Dl, 2 I, 3 none) of the N / Dl / I state in the
// previous plane (bits 0-1 / 2-3 / 4-5) and the cheapest state of the cell (bits 6-7).
constexpr int KEY_PLANES = 512;   // tgt_key = cost * KEY_PLANES + plane index
constexpr int KEY_INF = 0x7f7f7f7f;   // memset-able

// One candidate row of a chain pair, handed from the row kernel (k_ts_jump<C, false, true>) to the evaluation kernel (k_ts_eval):
// the start costs of the row's live lanes follow in Chunk::q_rows (a slot of 32 * C packed words, used from the front).
struct QueueHdr {
    int kk;                  // kind index | flags << 8 (bit 0 / 1: the low / high chain can still produce a seed below the bound) | entrance row << 12; < 0: reserved slot that was not used
    int b;                   // pair
    int e0;                  // primary end of the low chain (the high chain ends at e0 + 1)
    int lanes;               // lanes whose columns are stored (packed to the front of the slot's row): those with a start cost below the bound
    int nbw;                 // packed, negated pruning bounds of the two chains: -(slack - rowmin D(ip) - length cost), 0: no exit from this row
    int lcw;                 // packed length costs of the two chains at this row (INF16: no exit)
    int msw;                 // packed cheapest start costs of the two chains at this row (lower bound of every jump-in)
    int T;                   // pruning bound of the pair during this launch
};
constexpr int QUEUE_RESERVE = 8;   // slots a warp reserves per atomic: rows of one chain pair, evaluated by one warp of k_ts_eval

// Everything a kernel needs about the resident chunk of pairs.
struct Chunk {
    const PairMeta* pairs;
    const uint8_t* seq;      // alphabet indices
    const DevConfig* cfg;
    const int* lc;           // dense length costs (cfg->n_lc entries)
    unsigned char* tables;   // per pair, per (secondary, direction): sub[A][lw] u16, then open[lw], ext[lw] u32 (both halves)
    int16_t* D;              // [pair][i][j]  min_g cost of layer k at flank == L_f, clamped to INF16
    int16_t* DT;             // [pair][j][i]  the same, transposed (primary = query kinds read rows of it)
    uint8_t* dir;            // [pair][i][j]  traceback codes of the layer being filled (DIR_* bits), or null
    int* seedA;              // [pair][i][j]  reentry seeds of layer k+1 written by primary = reference kinds
    int* seedB;              // [pair][j][i]  reentry seeds written by primary = query kinds
    int* minvec;             // rowmin / colmin of D
    int* scratch;            // primary-fill column tiling
    int* best;               // [pair] best target cost over the layers filled so far
    int* best_layer;         // [pair] first layer that reached `best`
    int* active;             // [pair] layer k has seeds below best (fill it, then jump from it)
    int* next_active;        // [pair] set by the jump kernel when it writes a seed below best
    int* thr;                // [pair] exclusive cost threshold of the current deepening round (see k_resolve)
    int* ub;                 // [pair] pruning bound: candidates with cost >= ub are dropped (<= thr, tightened as targets are found)
    int* t0;                 // [pair] target cost of layer 0 (no template switch)
    int* resolved;           // [pair] optimum proven
    int* capped;             // [pair] still had seeds below the bound after max_layers template switches: refused (PAIR_ERR_LAYER_CAP)
    int round;               // deepening round (0 = first)
    unsigned kind_mask;      // kinds (index into DevConfig::kinds) evaluated by the jump kernel in this round
    // ---- flank planes (left/right flank lengths > 0); all null / 0 otherwise -------------------
    const int16_t* pl_in;    // [state N,Dl,I][cell]: plane 0 arrivals of the right-flank run (seeds of all three gap states)
    int16_t* pl_out;         // [state][cell]: plane 0 states, input of the left-flank run
    uint8_t* dir2;           // [cell] plane 0: DIR2_* bits (which states were taken from pl_in)
    long long cells_total;   // cells of the whole chunk (stride between the state planes of pl_in / pl_out)
    int* tgt_key;            // [pair] min over the planes of this layer of (target cost * 512 + plane index)
    int* best_plane;         // [pair] plane index (flank index + right flank length) of the best target
    int flank_mode;          // 1: k_layer_finish does the bookkeeping of k_primary_fill's epilogue
    // ---- column windows (pairs wider than a non-windowed jump class, see k_ts_jump) ------------------------------
    int* band;               // per pair at 2 * vec: rowlo[n+1], rowhi[n+1], collo[m+1], colhi[m+1]: first / last anti coordinate with
                             // D < thr - min_ts (the only cells a template switch below the threshold can start from); null: unused
    int* winflag;            // [pair] bit 0: a chain's window did not fit the first-stage class in this layer (redo in the second
                             // stage); bit 1: it did not fit the widest class either (the pair is refused)
    int* fill_prog;          // grid-pipelined primary fill: chunks finished per (pair, column block), zeroed before every launch
    int* cpflag;             // chain-pair bitmaps of the long class (PairMeta::cpf)
    int seeds_merged;        // 1: seedA already holds min(seedA, seedB transposed) (k_merge_seeds): the primary fill reads seedA only
    int win_stage;           // 0: not a windowed launch; 1: first stage; 2: second stage (only pairs with bit 0 set)
    // ---- row queue between the row kernel and the evaluation kernel (pairs that run without column windows) ---------------
    int* q_count;            // slots reserved in this launch (may exceed q_cap: the host then repeats the layer with smaller slices)
    int q_cap;               // slots of q_hdr / q_rows
    QueueHdr* q_hdr;
    uint32_t* q_rows;        // [slot][32 * C]: start costs of the row (both chains packed), "infinite" where no start is allowed
    int* counters;           // [0] pairs with next_active, [1..4] work statistics, [8 + class] compacted list sizes
};

// ---- traceback ------------------------------------------------------------------------------------------------
constexpr int MAX_TRACE_LAYERS = 64;
// Per-layer matrices kept for the traceback (layer k of every pair at the pair's `mat` offset).
struct TraceLayers {
    const uint8_t* dir[MAX_TRACE_LAYERS + 1];
    const int16_t* D[MAX_TRACE_LAYERS + 1];
    const uint8_t* dir2[MAX_TRACE_LAYERS + 1];   // flank mode: DIR2_* of plane 0
    const uint8_t* fd[MAX_TRACE_LAYERS + 1];     // flank mode: [step 1 .. RF+LF][cell] codes of the flank planes
    long long cells_total;
    int rf, lf;
    int wave;                                    // codes written by k_affine_wave: alignment range only, rows padded to 8 bytes
};
// One template switch of an alignment, in traceback order (last switch first).
struct TsRecord { int kind; int first_offset; int anti_primary_gap; int length; };
enum { TRACE_OK = 0, TRACE_ERR_OVERFLOW = 1, TRACE_ERR_NO_SOURCE = 2, TRACE_ERR_WALK = 3, TRACE_SKIPPED = 4 };
// Unit ops are written back to front, one byte each: 0..3 primary (insertion, deletion, substitution, match),
// 8..11 secondary (same order), 12 = template switch entrance, 13 = exit (payload in the TsRecord list).
struct TraceOut {
    uint8_t* ops;            // [pair] at ops_off: capacity ops_cap
    const long long* ops_off;
    const int* ops_cap;
    int* ops_len;            // [pair]
    TsRecord* recs;          // [pair][max_recs]
    int max_recs;
    int* n_recs;             // [pair]
    int* status;             // [pair] TRACE_*
    int16_t* rows;           // chain rows scratch: one block of rows_stride shorts per warp of the launch
    long long rows_stride;
};

}  // namespace tsa
