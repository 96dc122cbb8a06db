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
//   * strips of one pair are chained through a boundary column in global memory: one 8-byte entry per row (ND and I of the
//     strip's last column) written by lane 31 with a single store and prefetched 32 rows at a time, one chunk ahead, by
//     the next strip.  Every entry carries a 12-bit tag (strip index mod 4095) in the top bits of its two words, so an
//     entry validates itself: no flag, no fence -- the consumer polls a chunk until all 32 tags are its left neighbour's
//     (the buffer starts as all ones = tag 4095, which no strip writes, and is reused in place by all strips of the pair:
//     the latest writer of a row is some earlier strip, and two strips with equal tags are 4095 strips apart, more than
//     can be in flight).  The two values keep 26 bits: larger finite costs saturate to "infinite" and set a per-pair
//     flag; a result below 2^26 - 1 is exact regardless (costs are non-negative), anything else is reported as out of range;
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
    int* ticket;             // next strip to hand out
    const int* order;        // [2 * tickets] (pair index in list, strip) of every ticket, or null: tickets in (pair, strip) order
};

struct alignas(8) WaveBnd { uint32_t nd, i; };      // boundary entry: values in bits 0..25, tag bits 0-5 in nd[26..31], tag bits 6-11 in i[26..31]
constexpr int WAVE_SAT = 0x3FFFFFF;                 // 2^26 - 1: boundary values saturate here
constexpr uint32_t WAVE_TAGS = 4095;
TSA_DEV WaveBnd wave_bnd_load(const WaveBnd* p) {   // L2 load (another SM wrote it)
#ifndef TSA_EMUL
    const unsigned long long v = __ldcg(reinterpret_cast<const unsigned long long*>(p));
    return WaveBnd{(uint32_t)v, (uint32_t)(v >> 32)};
#else
    const volatile uint32_t* q = reinterpret_cast<const volatile uint32_t*>(p);
    return WaveBnd{q[0], q[1]};
#endif
}
TSA_DEV void wave_bnd_store(WaveBnd* p, WaveBnd v) {
#ifndef TSA_EMUL
    __stcg(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v.nd | ((unsigned long long)v.i << 32));
#else
    volatile uint32_t* q = reinterpret_cast<volatile uint32_t*>(p);
    q[0] = v.nd; q[1] = v.i;
#endif
}

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
        int lo = 0, s = 0;
        if (wa.order) { lo = wa.order[2 * tk]; s = wa.order[2 * tk + 1]; }
        else {
            int hi = wa.n_list - 1;                                // last pair whose first ticket is <= tk
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (wa.strip_prefix[mid] <= tk) lo = mid; else hi = mid - 1; }
            s = tk - wa.strip_prefix[lo];
        }
        const int b = wa.list[lo];
        const PairMeta pm = ck.pairs[b];
        const int nn = pm.rl - pm.ro, mm = pm.ql - pm.qo;         // the alignment range is the matrix
        const uint8_t* R = ck.seq + pm.seq_r + pm.ro;
        const uint8_t* Q = ck.seq + pm.seq_q + pm.qo;
        const bool last_strip = s == wave_strips(mm + 1) - 1;
        const int j0 = s * WAVE_SW + lane * CB;
        WaveBnd* bnd = reinterpret_cast<WaveBnd*>(ck.scratch + pm.scr);   // [row] ND / I of the column left of the next strip
        const uint32_t tag_in = (uint32_t)(s + WAVE_TAGS - 1) % WAVE_TAGS, tag_out = (uint32_t)s % WAVE_TAGS;   // tags of strip s - 1 / of this strip
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
        int rchunk = 0, bnd_nd = INF32, bnd_i = INF32;          // rows st .. st + 31 (lane k: row st + k)
        int rnext = 0;                                          // the chunk after that: characters and raw boundary entries,
        WaveBnd raw_next = WaveBnd{0, 0};                       // requested one chunk ahead and validated when they are needed
        // request: no wait -- the entries may not be written yet, the tags tell at validation time
        auto request_chunk = [&](int first) {
            const int row = first + lane;
            rnext = (row >= 1 && row <= nn) ? (int)R[row - 1] : 0;
            if (s > 0 && row <= nn) raw_next = wave_bnd_load(bnd + row);
        };
        // validate: all 32 tags must be the left neighbour's; entries that were requested too early are read again
        auto accept_chunk = [&](int first) {
            const int row = first + lane;
            rchunk = rnext; bnd_nd = INF32; bnd_i = INF32;
            if (s > 0 && first <= nn) {
                for (;;) {
                    const bool ok = row > nn || ((raw_next.nd >> 26) | ((raw_next.i >> 26) << 6)) == tag_in;
                    if (ballot(!ok) == 0) break;
                    spin_pause();
                    if (row <= nn) raw_next = wave_bnd_load(bnd + row);
                }
                if (row <= nn) {
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
        const int tcol = mm - j0;                                  // target column within this lane (0 .. CB-1) or outside
        const int steps = nn + 32;
        for (int st = 0; st < steps; st++) {
            if ((st & 31) == 0) {
                // rows st .. st + 31 enter lane 0 during the next 32 steps: validate their entries (requested 32 steps ago), then
                // request the chunk after them, so that the load latency overlaps these 32 steps.  A strip that runs too
                // close behind its left neighbour reads again here and thereby falls back until its requests succeed.
                accept_chunk(st);
                request_chunk(st + 32);
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
                    if ((left_nd >= WAVE_SAT && left_nd < INF32) || (left_i >= WAVE_SAT && left_i < INF32)) saturated = true;
                    wave_bnd_store(bnd + i, WaveBnd{(uint32_t)imin(left_nd, WAVE_SAT) | ((tag_out & 63u) << 26), (uint32_t)imin(left_i, WAVE_SAT) | ((tag_out >> 6) << 26)});
                }
            }
        }
        if (ballot(saturated) != 0 && lane == 0) atomic_or_s32(&ck.next_active[b], 1);   // (next_active is unused without template switches)
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
