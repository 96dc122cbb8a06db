// tsa_wave.cuh -- gap-affine fill without template switches (--no-ts, BASELINE config 4) as an anti-diagonal wavefront.
//
// Replaces the A* search over the primary transitions only (context.rs:135-354 with MaxTemplateSwitchCount(0),
// strategies/template_switch_count.rs:41-63): states N / Dl / I per cell, exactly as k_primary_fill, but organised for
// throughput on long pairs:
//   * the (range of the) matrix is cut into column strips of 32 * CB columns; one warp owns a strip, lane l owns CB
//     adjacent columns of it;
//   * inside a strip the warp sweeps anti-diagonals: at step t lane l computes row t - l, so that the left-neighbour
//     dependency (the insertion state and the diagonal) is the value lane l - 1 produced one step earlier and arrives by
//     __shfl_up_sync; there is no scan and no dependency inside a step except along the CB cells of a lane;
//   * strips of one pair are chained through a boundary column in global memory (2 ints per row, written by lane 31,
//     prefetched 32 rows at a time by the next strip) guarded by a release / acquire progress flag per strip;
//   * strips are handed out through one atomic ticket in (pair, strip) order, so the strip a warp waits for was always
//     claimed earlier by a warp that is running: any grid size is deadlock free, long pairs use many warps, short pairs
//     one.
// Costs are s32 (DPX __viaddmin_s32).  TRACE writes the same 1-byte traceback codes as k_primary_fill (DIR_* bits),
// 8 cells per 64-bit store, into a row-padded matrix of the alignment range (TraceLayers::wave).
#pragma once
#include "tsa_rt.hpp"
#include "tsa_types.hpp"

namespace tsa {

#ifndef TSA_WAVE_BLOCKS_PER_SM
#define TSA_WAVE_BLOCKS_PER_SM 5
#endif
constexpr int WAVE_CB = 8;                 // columns per lane
constexpr int WAVE_SW = 32 * WAVE_CB;      // columns per strip
constexpr int WAVE_WARPS = 4;
constexpr int WAVE_SMEM_INTS = MAX_ALPHABET * MAX_ALPHABET + 2 * MAX_ALPHABET;

TSA_HOSTDEV inline int wave_strips(int cols) { return (cols + WAVE_SW - 1) / WAVE_SW; }        // cols = range width + 1
TSA_HOSTDEV inline long long wave_dir_stride(int cols) { return ((long long)cols + 7) & ~7LL; }  // bytes per row of the code matrix

struct alignas(8) WaveCodes8 { uint32_t lo, hi; };   // the codes of 8 adjacent cells, one 64-bit store

struct WaveArgs {
    const int* list;         // pairs of this launch
    int n_list;
    const int* strip_prefix; // [n_list + 1] first ticket of every pair
    int* progress;           // [tickets] rows of the boundary column published by that strip
    int* ticket;             // next strip to hand out
};

template <bool TRACE>
TSA_KERNEL void TSA_LAUNCH_BOUNDS(32 * WAVE_WARPS, TSA_WAVE_BLOCKS_PER_SM) k_affine_wave(Chunk ck, WaveArgs wa) {
    TSA_SHARED_DECL(smem_raw);
    constexpr int CB = WAVE_CB;
    const int lane = lane_id();
    // substitution costs [r][q] with row stride A + 1 (dense: the 25 entries of a 5-letter alphabet sit in 25 different
    // banks); column A = "no character" (infinite).  Then the gap costs by character.
    const DevConfig* cfg = ck.cfg;
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
    const int total = wa.strip_prefix[wa.n_list];

    for (;;) {
        int tk = 0;
        if (lane == 0) tk = atomic_add_s32(wa.ticket, 1);
        tk = (int)shfl_idx((uint32_t)tk, 0);
        if (tk >= total) break;
        int lo = 0, hi = wa.n_list - 1;                           // last pair whose first ticket is <= tk
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (wa.strip_prefix[mid] <= tk) lo = mid; else hi = mid - 1; }
        const int b = wa.list[lo], s = tk - wa.strip_prefix[lo];
        const PairMeta pm = ck.pairs[b];
        const int nn = pm.rl - pm.ro, mm = pm.ql - pm.qo;         // the alignment range is the matrix
        const uint8_t* R = ck.seq + pm.seq_r + pm.ro;
        const uint8_t* Q = ck.seq + pm.seq_q + pm.qo;
        const bool last_strip = s == wave_strips(mm + 1) - 1;
        const int j0 = s * WAVE_SW + lane * CB;
        int* bnd = ck.scratch + pm.scr;                           // [row][ND, I] of the column left of the next strip
        const long long dstride = wave_dir_stride(mm + 1);
        uint8_t* dirp = TRACE ? ck.dir + pm.mat + j0 : nullptr;

        int qoff[CB], opQ[CB], exQ[CB], Mup[CB], Dlup[CB], NIup[CB];
#pragma unroll
        for (int c = 0; c < CB; c++) {
            const int j = j0 + c;
            const bool has = j >= 1 && j <= mm;
            const int qc = has ? (int)Q[j - 1] : A;
            qoff[c] = qc;
            opQ[c] = has ? openP[qc] : INF32;
            exQ[c] = has ? extP[qc] : INF32;
            (void)0;
            Mup[c] = INF32; Dlup[c] = INF32; NIup[c] = INF32;
        }
        int diag_in = INF32;                 // M(i - 1, j0 - 1)
        int out_nd = INF32, out_i = INF32, out_r = 0;            // what lane + 1 needs next step: ND / I of my last column, my row's character
        int rchunk = 0, bnd_nd = INF32, bnd_i = INF32;
        int tgt = INF32;
        const bool is_root_lane = s == 0 && lane == 0;
        const int tcol = mm - j0;                                  // target column within this lane (0 .. CB-1) or outside
        const int steps = nn + 32;
        for (int st = 0; st < steps; st++) {
            if ((st & 31) == 0) {
                const int row = st + lane;                        // rows st .. st + 31 enter lane 0 during the next 32 steps
                rchunk = (row >= 1 && row <= nn) ? (int)R[row - 1] : 0;
                if (s > 0 && st <= nn) {
                    const int need = imin(st + 32, nn + 1);
                    while (ld_acquire_s32(wa.progress + tk - 1) < need) spin_pause();
                    bnd_nd = row <= nn ? ld_cg_s32(bnd + 2 * row) : INF32;
                    bnd_i = row <= nn ? ld_cg_s32(bnd + 2 * row + 1) : INF32;
                }
            }
            int rch = (int)shfl_up((uint32_t)out_r, 1);
            int lnd = (int)shfl_up((uint32_t)out_nd, 1);
            int li = (int)shfl_up((uint32_t)out_i, 1);
            const int r0 = (int)shfl_idx((uint32_t)rchunk, st & 31);
            const int n0 = (int)shfl_idx((uint32_t)bnd_nd, st & 31);
            const int i0 = (int)shfl_idx((uint32_t)bnd_i, st & 31);
            if (lane == 0) { rch = r0; lnd = n0; li = i0; }
            const int i = st - lane;
            if (i >= 0 && i <= nn) {
                const int opR = i > 0 ? openP[rch] : INF32;
                const int exR = i > 0 ? extP[rch] : INF32;
                const int* srow = subP + rch * ws;
                int prevM = i > 0 ? diag_in : INF32;
                int left_nd = lnd, left_i = li;
                uint32_t w0 = 0, w1 = 0;
#pragma unroll
                for (int c = 0; c < CB; c++) {
                    int nn_ = addmin_s32(prevM, srow[qoff[c]], INF32);                 // diagonal (context.rs:174-208)
                    unsigned cd = nn_ < INF32 ? (unsigned)DIR_N_DIAG : 0u;
                    if (c == 0 && is_root_lane && i == 0) { nn_ = 0; cd = 0; }         // the root of the search
                    const int op = addmin_s32(NIup[c], opR, INF32);                    // deletion opened from N / I above
                    const int dl = addmin_s32(Dlup[c], exR, op);                       //          or extended
                    prevM = Mup[c];
                    const int nd = imin(nn_, dl);
                    const int iop = addmin_s32(left_nd, opQ[c], INF32);                // insertion opened from N / Dl on the left
                    const int iv = addmin_s32(left_i, exQ[c], iop);                    //           or extended
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
                diag_in = imin(lnd, li);
                out_nd = left_nd; out_i = left_i; out_r = rch;
                if (TRACE && j0 <= mm) *reinterpret_cast<WaveCodes8*>(dirp + (long long)i * dstride) = WaveCodes8{w0, w1};
                if (lane == 31 && !last_strip) {
                    bnd[2 * i] = left_nd; bnd[2 * i + 1] = left_i;
                    if ((i & 31) == 31 || i == nn) st_release_s32(wa.progress + tk, i + 1);
                }
            }
        }
        if (last_strip) {
            // every lane's Mup holds row nn now; target: any gap state (context.rs:731-748)
#pragma unroll
            for (int c = 0; c < CB; c++) if (c == tcol) tgt = Mup[c];
            tgt = reduce_min_s32(tgt);
            if (lane == 0) { ck.best[b] = tgt; ck.best_layer[b] = 0; ck.active[b] = 0; }
        }
    }
}

}  // namespace tsa
