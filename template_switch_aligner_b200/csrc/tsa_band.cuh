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
    const int* ckpt_in;      // row0 > 0: [entry][3] states N, Dl, I of row row0
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

template <bool TRACE>
TSA_KERNEL void TSA_LAUNCH_BOUNDS(32 * WAVE_WARPS, TSA_WAVE_BLOCKS_PER_SM) k_affine_band(const DevConfig* cfg, BandArgs ba) {
    TSA_SHARED_DECL(smem_raw);
    constexpr int CB = WAVE_CB;
    const int lane = lane_id();
    const int A = cfg->A, ws = A + 1;
    int* subP = reinterpret_cast<int*>(smem_raw);
    int* openP = subP + MAX_ALPHABET * MAX_ALPHABET;
    int* extP = openP + MAX_ALPHABET;
    for (int t = (int)threadIdx.x; t < A * ws; t += (int)blockDim.x) {
        const int r = t / ws, q = t % ws;
        subP[t] = q < A ? imin(cfg->sub[0][r * MAX_ALPHABET + q], INF32) : INF32;
    }
    for (int t = (int)threadIdx.x; t < MAX_ALPHABET; t += (int)blockDim.x) {
        openP[t] = t < A ? imin(cfg->open[0][t], INF32) : INF32;
        extP[t] = t < A ? imin(cfg->ext[0][t], INF32) : INF32;
    }
    sync_block();
    const int nn = ba.nn, mm = ba.mm, row0 = ba.row0, row1 = ba.row1;
    const uint8_t* R = ba.R;
    const uint8_t* Q = ba.Q;
    const long long col_rows = (long long)nn + 1;

    for (;;) {
        int tk = 0;
        if (lane == 0) tk = atomic_add_s32(ba.ticket, 1);
        tk = (int)shfl_idx((uint32_t)tk, 0);
        if (tk >= ba.n_strips) break;
        const int s = ba.s_lo + tk;
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
            if (i == row0 && row0 > 0) {
                // the checkpoint row: its states are loaded instead of computed
                int left_nd = INF32, left_i = INF32;
#pragma unroll
                for (int c = 0; c < CB; c++) {
                    const int j = j0 + c;
                    int vn = INF32, vd = INF32, vi = INF32;
                    if (j <= mm) {
                        const int* e = ba.ckpt_in + (long long)(j - ba.ck_col0 + 1) * 3;
                        vn = e[0]; vd = e[1]; vi = e[2];
                    }
                    Mup[c] = imin(vn, imin(vd, vi)); Dlup[c] = vd; NIup[c] = imin(vn, vi);
                    left_nd = imin(vn, vd); left_i = vi;
                }
                // M of the column to the left (the diagonal predecessor of row0 + 1, column j0): the neighbour lane's last column, or
                // for lane 0 the checkpoint entry of column j0 - 1 (entry 0 = the column left of the band, stored by the forward pass)
                int dleft = imin(lnd, li);
                if (lane == 0) {
                    dleft = INF32;
                    if (j0 >= 1) { const int* e = ba.ckpt_in + (long long)(j0 - 1 - ba.ck_col0 + 1) * 3; dleft = imin(e[0], imin(e[1], e[2])); }
                }
                diag_in = dleft;
                out_nd = left_nd; out_i = left_i; out_r = rch;
                // the next strip of the group validates the entry of this row like any other
                if (lane == 31 && bnd_wr != nullptr && bnd_wr == ba.bnd_local)
                    wave_bnd_store(bnd_wr + i, WaveBnd{(uint32_t)imin(left_nd, WAVE_SAT) | ((tag_out & 63u) << 26), (uint32_t)imin(left_i, WAVE_SAT) | ((tag_out >> 6) << 26)});
                continue;
            }
            const int opR = i > 0 ? openP[rch] : INF32;
            const int exR = i > 0 ? extP[rch] : INF32;
            const int* srow = subP + rch * ws;
            int prevM = i > 0 ? diag_in : INF32;
            int left_nd = lnd, left_i = li;
            uint32_t w0 = 0, w1 = 0;
            const bool store_ck = !TRACE && ba.ckpt_out != nullptr && i > 0 && (i % ba.interval) == 0;
            int* ck_row = store_ck ? ba.ckpt_out + (long long)(i / ba.interval - 1) * ba.ckpt_stride : nullptr;
            if (store_ck && lane == 0 && j0 == ba.ck_col0) { ck_row[0] = lnd; ck_row[1] = INF32; ck_row[2] = li; }   // the column left of the band
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
                if (store_ck && j0 + c <= mm) {
                    int* e = ck_row + (long long)(j0 + c - ba.ck_col0 + 1) * 3;
                    e[0] = nn_; e[1] = dl; e[2] = iv;
                }
                Mup[c] = M; Dlup[c] = dl; NIup[c] = imin(nn_, iv);
                left_nd = nd; left_i = iv;
            }
            diag_in = imin(lnd, li);
            out_nd = left_nd; out_i = left_i; out_r = rch;
            if (TRACE && j0 <= mm) *reinterpret_cast<WaveCodes8*>(dirp + (long long)(i - ba.row_base) * ba.dstride) = WaveCodes8{w0, w1};
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

TSA_KERNEL void k_band_walk(const DevConfig* cfg, WalkArgs wa) {
    if (threadIdx.x >= 32 || blockIdx.x != 0) return;
    const int lane = lane_id();
    WalkState st = *wa.state;
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
    if (lane == 0) {
        st.i = i; st.j = j; st.g = g; st.need = need; st.cost = cost; st.status = status;
        *wa.state = st;
        *wa.ops_len = pos;
    }
}

}  // namespace tsa
