// tsa_band.cuh -- one very long pair without template switches (BASELINE config 5): the gap-affine wavefront of tsa_wave.cuh
// as a column band of one GPU, with checkpoints instead of a resident code matrix.
//
// Replaces, for this shape, what the reference does with `--memory-limit` (tsalign/src/align.rs:57-223 -> generic_a_star/src/
// lib.rs:332-335,380-389 gives up with ExceededMemoryLimit when the node store outgrows the limit): here the limit bounds what
// is resident, the alignment is still produced.
//   * Column bands: rank g of `world` owns a contiguous range of column groups (a group = `group` strips of 256 columns).
//     Strips hand their boundary column to the next strip through L2 exactly as in k_affine_wave (8-byte self-validating
//     entries, tag = global strip index).  The LAST strip of a band writes its entries into the NEXT rank's memory -- plain
//     8-byte stores through a peer mapping (NVLink P2P; cudaIpcOpenMemHandle between the processes of a torchrun launch) --
//     and the first strip of that rank polls them like any other boundary: no flag, no fence, no collective.
//   * Checkpoints: the forward pass keeps no traceback codes.  Every `interval` rows the three states (N, Dl, I) of the band's
//     columns are stored (12 B per column and checkpoint row); the boundary column entering every group is kept for all rows
//     (8 B per row and group: the entries the strips exchange anyway, written once, never reused).  The traceback recomputes, with
//     codes, only the (row block x column group) tiles the optimal path crosses -- from the checkpoint row above and the boundary
//     column to the left -- and walks them one after the other (k_band_walk); a walk that leaves the band to the left is handed
//     to the rank on the left.
// Costs are s32 (DPX __viaddmin_s32); boundary values saturate at 2^26 - 1 as in k_affine_wave.
#pragma once
#include "tsa_rt.hpp"
#include "tsa_types.hpp"
#include "tsa_wave.cuh"

namespace tsa {

struct BandArgs {
    const uint8_t* R;        // alignment range of the reference (encoded), nn characters
    const uint8_t* Q;        // alignment range of the query, mm characters
    int nn, mm;
    int s_lo, n_strips;      // strips of this launch (global strip indices s_lo .. s_lo + n_strips - 1), handed out by the ticket
    int s_total;             // strips of the whole pair
    int s_band_first;        // first strip of this rank's band: it reads colck[0] (written by the rank on the left)
    int s_band_last;         // last strip of this rank's band: it writes bnd_out
    int group;               // strips per column group
    int row0, row1;          // rows row0 + 1 .. row1 are computed from the state of row row0 (row0 == 0: row 0 too, from the root)
    int ck_col0;             // first column of the band: checkpoint entry of column j = j - ck_col0 + 1 (entry 0: the column left of the band)
    const int* ckpt_in;      // row0 > 0: [entry][3] of row row0: cheapest state M, deletion state Dl, min(N, I) (what the row below reads)
    int* ckpt_out;           // forward pass: [row / interval - 1][entry][3]; null: no checkpoints are stored
    int interval;            // checkpoint every `interval` rows
    long long ckpt_stride;   // ints per checkpoint row
    WaveBnd* colck;          // [group index - colck_g0][nn + 1]: the boundary column entering a group (group colck_g0 of a rank > 0: written
    int colck_g0;            //   by the rank on the left)
    int store_cols;          // 1: the last strip of a group writes the next group's boundary column (forward pass); 0: tile recomputation
    WaveBnd* bnd_local;      // [row] between the strips of a group, reused in place
    WaveBnd* bnd_out;        // [row] in the next rank's memory (null: nothing to hand over)
    uint8_t* dir;            // TRACE: codes of the tile, `dstride` bytes per row, first row = row_base, first column = s_lo * 256
    long long dstride;
    int row_base;            // first row that has codes (row0 + 1, or 0)
    int* ticket;             // next strip of this launch to hand out
    int* result;             // [0] cost of the target cell (the band that holds it), [1] saturation flag
};

// Walk state between tiles.  need: which code of the cell (i, j) still has to be looked up to know the state g
// (0: g is known, 1: cheapest state of the cell, 2: "min(N, I) is I", 3: "min(N, Dl) is Dl").
struct WalkState { int i, j, g, need; long long cost; int status; int pad; };   // status: 0 walking, 1 reached the root, < 0 error
enum { WALK_GOING = 0, WALK_DONE = 1, WALK_ERR = -1, WALK_OPS_FULL = -2 };

// Shared memory tables of the wavefront kernels: substitution costs [r][q] with row stride A + 1, then the gap costs by character.
TSA_DEV void band_stage_tables(const DevConfig* cfg, int* subP, int* openP, int* extP) {
    const int A = cfg->A, ws = A + 1;
    for (int t = (int)threadIdx.x; t < A * ws; t += (int)blockDim.x) {
        const int r = t / ws, q = t % ws;
        subP[t] = q < A ? imin(cfg->sub[0][r * MAX_ALPHABET + q], INF32) : INF32;
    }
    for (int t = (int)threadIdx.x; t < MAX_ALPHABET; t += (int)blockDim.x) {
        openP[t] = t < A ? imin(cfg->open[0][t], INF32) : INF32;
        extP[t] = t < A ? imin(cfg->ext[0][t], INF32) : INF32;
    }
    sync_block();
}

// One strip (global strip index s) of the launch described by `ba`, by one warp.
template <bool TRACE>
TSA_DEV void band_strip(const BandArgs& ba, int s, int A, const int* subP, const int* openP, const int* extP) {
    constexpr int CB = WAVE_CB;
    const int lane = lane_id();
    const int ws = A + 1;
    const int nn = ba.nn, mm = ba.mm, row0 = ba.row0, row1 = ba.row1;
    const uint8_t* R = ba.R;
    const uint8_t* Q = ba.Q;
    const long long col_rows = (long long)nn + 1;
    {
        const bool last_strip = s == ba.s_total - 1;
        const int j0 = s * WAVE_SW + lane * CB;
        // where the boundary column comes from and where this strip's last column goes
        const WaveBnd* bnd_rd = s == ba.s_band_first ? ba.colck : ((s % ba.group == 0) ? ba.colck + (long long)(s / ba.group - ba.colck_g0) * col_rows : ba.bnd_local);
        WaveBnd* bnd_wr = ba.bnd_local;
        if (last_strip) bnd_wr = nullptr;
        else if (s == ba.s_band_last) bnd_wr = ba.bnd_out;
        else if ((s + 1) % ba.group == 0) bnd_wr = ba.store_cols ? ba.colck + (long long)((s + 1) / ba.group - ba.colck_g0) * col_rows : nullptr;
        const uint32_t tag_in = (uint32_t)(s + WAVE_TAGS - 1) % WAVE_TAGS, tag_out = (uint32_t)s % WAVE_TAGS;
        uint8_t* dirp = TRACE ? ba.dir + (j0 - ba.s_lo * WAVE_SW) : nullptr;

        int qoff[CB], opQ[CB], exQ[CB], Mup[CB], Dlup[CB], NIup[CB];
#pragma unroll
        for (int c = 0; c < CB; c++) {
            const int j = j0 + c;
            const bool has = j >= 1 && j <= mm;
            const int qc = has ? (int)Q[j - 1] : A;
            qoff[c] = qc;
            opQ[c] = has ? openP[qc] : INF32;
            exQ[c] = has ? extP[qc] : INF32;
            Mup[c] = INF32; Dlup[c] = INF32; NIup[c] = INF32;
        }
        int diag_in = INF32;                                      // M(i - 1, j0 - 1)
        int out_nd = INF32, out_i = INF32, out_r = 0;
        int rchunk = 0, bnd_nd = INF32, bnd_i = INF32;
        int rnext = 0;
        WaveBnd raw_next = WaveBnd{0, 0};
        // chunks of 32 rows, relative to row0: chunk row t is the absolute row row0 + t
        auto request_chunk = [&](int first) {
            const int row = row0 + first + lane;
            rnext = (row >= 1 && row <= row1) ? (int)R[row - 1] : 0;
            if (s > 0 && row <= row1) raw_next = wave_bnd_load(bnd_rd + row);
        };
        auto accept_chunk = [&](int first) {
            const int row = row0 + first + lane;
            rchunk = rnext; bnd_nd = INF32; bnd_i = INF32;
            if (s > 0 && row0 + first <= row1) {
                for (;;) {
                    const bool ok = row > row1 || ((raw_next.nd >> 26) | ((raw_next.i >> 26) << 6)) == tag_in;
                    if (ballot(!ok) == 0) break;
                    spin_pause();
                    if (row <= row1) raw_next = wave_bnd_load(bnd_rd + row);
                }
                if (row <= row1) {
                    bnd_nd = (int)(raw_next.nd & (uint32_t)WAVE_SAT); bnd_i = (int)(raw_next.i & (uint32_t)WAVE_SAT);
                    if (bnd_nd == WAVE_SAT) bnd_nd = INF32;
                    if (bnd_i == WAVE_SAT) bnd_i = INF32;
                }
            }
        };
        request_chunk(0);
        int tgt = INF32;
        bool saturated = false;
        const bool is_root_lane = s == 0 && lane == 0;
        const int tcol = mm - j0;
        const int steps = (row1 - row0) + 32;
        int* const ck_out = ba.ckpt_out;
        const int ck_interval = imax(ba.interval, 1);
        const int ck_shift = (ck_interval & (ck_interval - 1)) == 0 ? 31 - clz_u32((uint32_t)ck_interval) : -1;   // power of two: no division per checkpoint row
        const long long dstride = ba.dstride;
        const int row_base = ba.row_base;
        const bool local_out = bnd_wr != nullptr && bnd_wr == ba.bnd_local;
        int ck_phase = row0 % ck_interval;
        for (int st = 0; st < steps; st++) {
            if ((st & 31) == 0) { accept_chunk(st); request_chunk(st + 32); }
            int rch = (int)shfl_up((uint32_t)out_r, 1);
            int lnd = (int)shfl_up((uint32_t)out_nd, 1);
            int li = (int)shfl_up((uint32_t)out_i, 1);
            const int r0 = (int)shfl_idx((uint32_t)rchunk, st & 31);
            const int n0 = (int)shfl_idx((uint32_t)bnd_nd, st & 31);
            const int i0 = (int)shfl_idx((uint32_t)bnd_i, st & 31);
            if (lane == 0) { rch = r0; lnd = n0; li = i0; }
            const int i = row0 + st - lane;
            if (i < row0 || i > row1) continue;
            // phase of row i within the checkpoint interval, kept incrementally (a lane visits the rows row0, row0 + 1, ... in order)
            const int ph = ck_phase;
            ck_phase = ck_phase + 1 == ck_interval ? 0 : ck_phase + 1;
            if (i == row0 && row0 > 0) {
                // the checkpoint row: what the next row needs of it is loaded instead of computed: per column the cheapest state M,
                // the deletion state Dl and min(N, I) -- exactly the registers the row below reads -- and M of the column to the left
                // (the diagonal predecessor of row0 + 1, column j0; entry 0 = the column left of the band)
#pragma unroll
                for (int c = 0; c < CB; c++) {
                    const int j = j0 + c;
                    int vm = INF32, vd = INF32, vni = INF32;
                    if (j <= mm) {
                        const int* e = ba.ckpt_in + (long long)(j - ba.ck_col0 + 1) * 3;
                        vm = e[0]; vd = e[1]; vni = e[2];
                    }
                    Mup[c] = vm; Dlup[c] = vd; NIup[c] = vni;
                }
                diag_in = j0 >= 1 && j0 - 1 <= mm ? ba.ckpt_in[(long long)(j0 - 1 - ba.ck_col0 + 1) * 3] : INF32;
                // (the neighbour lane and the next strip read nothing of this row but the tag of its boundary entry)
                out_nd = INF32; out_i = INF32; out_r = rch;
                if (lane == 31 && local_out)
                    wave_bnd_store(bnd_wr + i, WaveBnd{(uint32_t)WAVE_SAT | ((tag_out & 63u) << 26), (uint32_t)WAVE_SAT | ((tag_out >> 6) << 26)});
                continue;
            }
            const int opR = i > 0 ? openP[rch] : INF32;
            const int exR = i > 0 ? extP[rch] : INF32;
            const int* srow = subP + rch * ws;
            int prevM = i > 0 ? diag_in : INF32;
            int left_nd = lnd, left_i = li;
            uint32_t w0 = 0, w1 = 0;
            const bool store_ck = !TRACE && ck_out != nullptr && i > 0 && ph == 0;
#pragma unroll
            for (int c = 0; c < CB; c++) {
                int nn_ = addmin_s32(prevM, srow[qoff[c]], INF32);
                unsigned cd = nn_ < INF32 ? (unsigned)DIR_N_DIAG : 0u;
                if (c == 0 && is_root_lane && i == 0) { nn_ = 0; cd = 0; }
                const int op = addmin_s32(NIup[c], opR, INF32);
                const int dl = addmin_s32(Dlup[c], exR, op);
                prevM = Mup[c];
                const int nd = imin(nn_, dl);
                const int iop = addmin_s32(left_nd, opQ[c], INF32);
                const int iv = addmin_s32(left_i, exQ[c], iop);
                const int M = imin(nd, iv);
                if (TRACE) {
                    if (dl < op) cd |= DIR_DL_EXT;
                    if (iv < iop) cd |= DIR_I_EXT;
                    cd |= (nn_ <= M ? 0u : (dl <= M ? 1u : 2u)) << DIR_M_SHIFT;
                    if (iv < nn_) cd |= DIR_NI_IS_I;
                    if (dl < nn_) cd |= DIR_ND_IS_DL;
                    if (c < 4) w0 |= cd << (8 * c); else w1 |= cd << (8 * (c - 4));
                }
                Mup[c] = M; Dlup[c] = dl; NIup[c] = imin(nn_, iv);
                left_nd = nd; left_i = iv;
            }
            if (store_ck) {
                // checkpoint row: M, Dl, min(N, I) of every column, from the registers the next row reads
                int* ck_row = ck_out + (long long)((ck_shift >= 0 ? i >> ck_shift : i / ck_interval) - 1) * ba.ckpt_stride;
                if (lane == 0 && j0 == ba.ck_col0) { ck_row[0] = imin(lnd, li); ck_row[1] = INF32; ck_row[2] = INF32; }   // the column left of the band
#pragma unroll
                for (int c = 0; c < CB; c++) if (j0 + c <= mm) {
                    int* e = ck_row + (long long)(j0 + c - ba.ck_col0 + 1) * 3;
                    e[0] = Mup[c]; e[1] = Dlup[c]; e[2] = NIup[c];
                }
            }
            diag_in = imin(lnd, li);
            out_nd = left_nd; out_i = left_i; out_r = rch;
            if (TRACE && j0 <= mm) *reinterpret_cast<WaveCodes8*>(dirp + (long long)(i - row_base) * dstride) = WaveCodes8{w0, w1};
            if (lane == 31 && bnd_wr != nullptr) {
                if ((left_nd >= WAVE_SAT && left_nd < INF32) || (left_i >= WAVE_SAT && left_i < INF32)) saturated = true;
                wave_bnd_store(bnd_wr + i, WaveBnd{(uint32_t)imin(left_nd, WAVE_SAT) | ((tag_out & 63u) << 26), (uint32_t)imin(left_i, WAVE_SAT) | ((tag_out >> 6) << 26)});
            }
        }
        if (ballot(saturated) != 0 && lane == 0) atomic_or_s32(&ba.result[1], 1);
        if (last_strip && row1 == nn && !TRACE) {
#pragma unroll
            for (int c = 0; c < CB; c++) if (c == tcol) tgt = Mup[c];
            tgt = reduce_min_s32(tgt);
            if (lane == 0) ba.result[0] = tgt;
        }
    }
}

template <bool TRACE>
TSA_KERNEL void TSA_LAUNCH_BOUNDS(32 * WAVE_WARPS, TSA_WAVE_BLOCKS_PER_SM) k_affine_band(const DevConfig* cfg, BandArgs ba) {
    TSA_SHARED_DECL(smem_raw);
    int* subP = reinterpret_cast<int*>(smem_raw);
    int* openP = subP + MAX_ALPHABET * MAX_ALPHABET;
    int* extP = openP + MAX_ALPHABET;
    band_stage_tables(cfg, subP, openP, extP);
    const int lane = lane_id();
    for (;;) {
        int tk = 0;
        if (lane == 0) tk = atomic_add_s32(ba.ticket, 1);
        tk = (int)shfl_idx((uint32_t)tk, 0);
        if (tk >= ba.n_strips) break;
        band_strip<TRACE>(ba, ba.s_lo + tk, cfg->A, subP, openP, extP);
    }
}

// Walk of the optimal path inside one tile, backwards from `state`, by one warp.  Leaves the tile upwards when it arrives at row
// row_lo (> 0) and to the left when it arrives left of column col_base; the state then says which code of that cell the next tile
// has to look up (WalkState::need).  Unit ops are appended back to front (0 insertion, 1 deletion, 2 substitution, 3 match: the
// numbering of TraceOut).  Runs of diagonal moves are taken 32 at a time: lane l looks at the cell l steps up the diagonal, the
// run ends at the first cell whose cheapest state is not "N by the diagonal".  `cost` is decremented by every edge cost, so that
// the sum over all tiles must reach exactly 0 at the root (checked by the host).
struct WalkArgs {
    const uint8_t* R; const uint8_t* Q;
    const uint8_t* dir; long long dstride; int row_base, col_base;
    int row_lo;              // leave upwards at this row (0: the tile holds row 0, the walk ends at the root)
    uint8_t* ops; int ops_cap;
    int* ops_len;            // [1]
    WalkState* state;        // in / out
};

// The walk itself: all 32 lanes of one warp call it; returns the number of unit ops appended (wa.ops[0 .. pos)).
TSA_DEV int band_walk(const DevConfig* cfg, const WalkArgs& wa, WalkState& st) {
    const int lane = lane_id();
    int i = st.i, j = st.j, g = st.g, need = st.need, pos = 0;
    long long cost = st.cost;
    int status = WALK_GOING;
    const uint8_t* R = wa.R;
    const uint8_t* Q = wa.Q;
    for (;;) {
        if (j < wa.col_base || (i <= wa.row_lo && wa.row_lo > 0)) break;       // the cell belongs to another tile
        if (pos + 32 > wa.ops_cap) { status = WALK_OPS_FULL; break; }
        const int ii = i - lane, jj = j - lane;
        const bool inside = jj >= wa.col_base && ii >= 0 && (ii > wa.row_lo || wa.row_lo == 0);
        const int code = inside ? (int)wa.dir[(long long)(ii - wa.row_base) * wa.dstride + (jj - wa.col_base)] : 0;
        int gl = (code >> DIR_M_SHIFT) & 3;
        if (lane == 0 && need != 1) gl = need == 0 ? g : (need == 2 ? ((code & DIR_NI_IS_I) ? 2 : 0) : ((code & DIR_ND_IS_DL) ? 1 : 0));
        const bool diag = inside && gl == 0 && (code & DIR_N_DIAG) && ii > 0 && jj > 0;
        const uint32_t mask = ballot(diag);
        const int run = mask == 0xffffffffu ? 32 : ffs_u32(~mask) - 1;
        if (run > 0) {
            int c = 0;
            if (lane < run) {
                const int r = R[ii - 1], q = Q[jj - 1];
                wa.ops[pos + lane] = r == q ? 3 : 2;
                c = cfg->sub[0][r * MAX_ALPHABET + q];
            }
            cost -= (long long)reduce_add_s32(c);
            pos += run; i -= run; j -= run; g = 0; need = 1;
            continue;
        }
        // one step that is not a diagonal move, the same on every lane (lane 0's cell)
        const int code0 = (int)shfl_idx((uint32_t)code, 0);
        g = (int)shfl_idx((uint32_t)gl, 0);
        need = 0;
        if (g == 0) {
            status = (!(code0 & DIR_N_DIAG) && i == 0 && j == 0 && cost == 0) ? WALK_DONE : WALK_ERR;   // the root, or an inconsistency
            break;
        } else if (g == 1) {
            if (i <= 0) { status = WALK_ERR; break; }
            const int r = R[i - 1];
            if (lane == 0) wa.ops[pos] = 1;
            pos++; i--;
            if (code0 & DIR_DL_EXT) { cost -= cfg->ext[0][r]; g = 1; }
            else { cost -= cfg->open[0][r]; need = 2; }
        } else {
            if (j <= 0) { status = WALK_ERR; break; }
            const int q = Q[j - 1];
            if (lane == 0) wa.ops[pos] = 0;
            pos++; j--;
            if (code0 & DIR_I_EXT) { cost -= cfg->ext[0][q]; g = 2; }
            else { cost -= cfg->open[0][q]; need = 3; }
        }
    }
    st.i = i; st.j = j; st.g = g; st.need = need; st.cost = cost; st.status = status;
    return pos;
}

static TSA_KERNEL void k_band_walk(const DevConfig* cfg, WalkArgs wa) {
    if (threadIdx.x >= 32 || blockIdx.x != 0) return;
    WalkState st = *wa.state;
    const int pos = band_walk(cfg, wa, st);
    if (lane_id() == 0) { *wa.state = st; *wa.ops_len = pos; }
}

// ---------------------------------------------------------------------------------------------------- batches of long pairs
// BASELINE config 4 (many 10 kb pairs, --no-ts) WITH alignments and without a code matrix: the forward pass of every pair stores
// checkpoint rows and group boundary columns exactly as one band of a long pair does (one BandArgs per pair, strips of all pairs
// handed out through one ticket in (pair, strip) order), then one warp per pair recomputes, with codes, only the tiles its
// optimal path crosses -- strip after strip into a per-warp scratch tile -- and walks them.  1 B/cell of codes (and the 14
// instructions per cell that produce them in the forward pass) become ~6 % recomputed cells and 0.06 B/cell of checkpoints.
constexpr int BB_INTERVAL = 256;   // checkpoint every 256 rows, tiles of 2 strips: a 10 kb pair recomputes ~6 % of its cells
constexpr int BB_GROUP = 2;

struct BandBatch {
    const BandArgs* args;    // [n_pairs] forward description of every pair (s_lo = 0, all strips, row0 = 0, row1 = nn)
    const int* pair_of;      // [n_pairs] staged pair index (Chunk arrays, TraceOut)
    const int* prefix;       // [n_pairs + 1] first ticket of every pair
    int n_pairs;
    int* ticket;
    const int* order;        // [2 * tickets] (entry of args, strip) of every ticket, or null: tickets in (pair, strip) order
};

// TRACE = false: forward passes of many pairs.  TRACE = true: many TILES (of one pair or several) recomputed with codes at once --
// every entry of bb.args is then one tile (row block x column group) with its own code and boundary buffers.
template <bool TRACE>
TSA_KERNEL void TSA_LAUNCH_BOUNDS(32 * WAVE_WARPS, TSA_WAVE_BLOCKS_PER_SM) k_band_batch(const DevConfig* cfg, BandBatch bb) {
    TSA_SHARED_DECL(smem_raw);
    int* subP = reinterpret_cast<int*>(smem_raw);
    int* openP = subP + MAX_ALPHABET * MAX_ALPHABET;
    int* extP = openP + MAX_ALPHABET;
    band_stage_tables(cfg, subP, openP, extP);
    const int lane = lane_id();
    const int total = bb.prefix[bb.n_pairs];
    for (;;) {
        int tk = 0;
        if (lane == 0) tk = atomic_add_s32(bb.ticket, 1);
        tk = (int)shfl_idx((uint32_t)tk, 0);
        if (tk >= total) break;
        int lo = 0, s = 0;
        if (bb.order) { lo = bb.order[2 * tk]; s = bb.order[2 * tk + 1]; }
        else {
            int hi = bb.n_pairs - 1;                               // last pair whose first ticket is <= tk
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (bb.prefix[mid] <= tk) lo = mid; else hi = mid - 1; }
            s = tk - bb.prefix[lo];
        }
        band_strip<TRACE>(bb.args[lo], bb.args[lo].s_lo + s, cfg->A, subP, openP, extP);      // (the description stays in global memory: its fields are read once per strip)
    }
}

// results of the forward pass into the per-pair arrays the rest of the engine reads
static TSA_KERNEL void k_band_batch_finish(Chunk ck, BandBatch bb) {
    const int k = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (k >= bb.n_pairs) return;
    const int b = bb.pair_of[k];
    ck.best[b] = bb.args[k].result[0]; ck.best_layer[b] = 0; ck.active[b] = 0;
    ck.next_active[b] = bb.args[k].result[1];      // boundary values saturated (WAVE_SAT)
}

struct BandBatchTrace {
    uint8_t* tiles;          // per warp of the launch: codes of one tile, (interval + 1) rows of group * 256 bytes
    long long tile_bytes;
    WaveBnd* bnds;           // per warp: boundary between the strips of a tile, interval + 1 entries
};

static TSA_KERNEL void TSA_LAUNCH_BOUNDS(32 * WAVE_WARPS, 3) k_band_batch_trace(const DevConfig* cfg, Chunk ck, BandBatch bb, BandBatchTrace bt, TraceOut to) {
    TSA_SHARED_DECL(smem_raw);
    int* subP = reinterpret_cast<int*>(smem_raw);
    int* openP = subP + MAX_ALPHABET * MAX_ALPHABET;
    int* extP = openP + MAX_ALPHABET;
    band_stage_tables(cfg, subP, openP, extP);
    const int lane = lane_id();
    const int wid = (int)blockIdx.x * WAVE_WARPS + (int)(threadIdx.x >> 5), nw = (int)gridDim.x * WAVE_WARPS;
    uint8_t* tile = bt.tiles + (long long)wid * bt.tile_bytes;
    for (int k = wid; k < bb.n_pairs; k += nw) {
        const int b = bb.pair_of[k];
        const BandArgs fa = bb.args[k];
        const int best = ck.best[b];
        if (best >= INF32) { if (lane == 0) { to.status[b] = TRACE_SKIPPED; to.ops_len[b] = 0; to.n_recs[b] = 0; } continue; }
        const int IV = fa.interval, G = fa.group;
        WalkState st;
        st.i = fa.nn; st.j = fa.mm; st.g = 0; st.need = 1; st.cost = best; st.status = WALK_GOING; st.pad = 0;
        uint8_t* ops = to.ops + to.ops_off[b];
        const int cap = to.ops_cap[b];
        int pos = 0;
        while (st.status == WALK_GOING) {
            if (st.j < 0) { st.status = WALK_ERR; break; }
            const int kk = st.i == 0 ? 0 : (st.i - 1) / IV;
            const int row0 = kk * IV, row1 = st.i;
            const int s_hi = st.j / WAVE_SW, s_lo = (s_hi / G) * G;
            BandArgs ta = fa;
            ta.s_lo = s_lo; ta.n_strips = s_hi - s_lo + 1;
            ta.row0 = row0; ta.row1 = row1;
            ta.ckpt_in = kk > 0 ? fa.ckpt_out + (long long)(kk - 1) * fa.ckpt_stride : nullptr;
            ta.ckpt_out = nullptr; ta.store_cols = 0;
            ta.bnd_local = bt.bnds + (long long)wid * (IV + 1) - row0;      // indexed by the absolute row
            ta.bnd_out = nullptr;
            ta.dir = tile; ta.dstride = (long long)G * WAVE_SW; ta.row_base = kk > 0 ? row0 + 1 : 0;
            for (int s = s_lo; s <= s_hi; s++) {
                band_strip<true>(ta, s, cfg->A, subP, openP, extP);
                sync_warp();      // the next strip reads this strip's boundary entries, the walk reads the codes
            }
            WalkArgs wa;
            wa.R = fa.R; wa.Q = fa.Q; wa.dir = tile; wa.dstride = ta.dstride; wa.row_base = ta.row_base; wa.col_base = s_lo * WAVE_SW;
            wa.row_lo = row0; wa.ops = ops + pos; wa.ops_cap = cap - pos; wa.ops_len = nullptr; wa.state = nullptr;
            const WalkState before = st;
            pos += band_walk(cfg, wa, st);
            sync_warp();
            if (st.status == WALK_GOING && st.i == before.i && st.j == before.j && st.need == before.need && st.g == before.g) st.status = WALK_ERR;
        }
        if (lane == 0) {
            to.status[b] = st.status == WALK_DONE ? TRACE_OK : (st.status == WALK_OPS_FULL ? TRACE_ERR_OVERFLOW : TRACE_ERR_WALK);
            to.ops_len[b] = pos; to.n_recs[b] = 0;
        }
    }
}

}  // namespace tsa
