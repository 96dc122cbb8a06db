// tsa_engine.cu -- see tsa_engine.hpp.  Compiled by nvcc for sm_100a (product) or by g++ with -DTSA_EMUL
// (tests/emul only).
#include "tsa_engine.hpp"

#include <algorithm>
#include <atomic>
#include <mutex>
#include <chrono>
#include <cstring>
#include <stdexcept>

#include "tsa_kernels.cuh"
#include "tsa_band.cuh"

namespace tsa {

namespace {

// Jump-kernel classes by pair width.  0..3: k_ts_jump<C, false> over the whole sequences (96 .. 544 columns).
// 4 ("medium", up to 1055): column windows of 544 columns first (k_ts_jump<17, true>), whole sequences with C = 33 for
// the pairs whose windows did not fit.  5 ("long", anything wider): windows of 544, then windows of 1056 columns, then (third
// stage) 1056-column windows over sub-ranges of the entrance columns, which fit any pair; only a cost model whose offset /
// length-difference hulls alone exceed the lane grid is refused (PAIR_ERR_TOO_LONG).
constexpr int N_CLASS = 6;
const int CLASS_C[N_CLASS] = {3, 5, 9, 17, 33, 33};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes) {
        if (bytes > cap) { rt::dev_free(p); p = nullptr; cap = 0; p = rt::dev_alloc(bytes); cap = bytes; }   // (an allocation that throws leaves the buffer empty)
    }
    void reset() { rt::dev_free(p); p = nullptr; cap = 0; }
    ~DevBuf() { rt::dev_free(p); }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

// Page-locked host staging buffer: device-to-host copies of results run at PCIe speed instead of through a bounce buffer.
struct HostBuf {
    void* p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes) {
        if (bytes > cap) { rt::host_free(p); p = nullptr; cap = 0; p = rt::host_alloc(bytes); cap = bytes; }
    }
    ~HostBuf() { rt::host_free(p); }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

#ifndef TSA_EMUL
// Function attributes (dynamic shared memory opt-in) are per device: one flag per (kernel instantiation, device).
struct PerDeviceOnce {
    std::mutex mu;
    bool done[64] = {};
    // Runs `init` once per device; a concurrent caller (the second engine thread of a call) waits until it has finished, so that
    // no kernel is launched before its attributes are set.
    template <class F> void run(F&& init) {
        int dev = 0;
        cudaGetDevice(&dev);
        dev &= 63;
        std::lock_guard<std::mutex> guard(mu);
        if (!done[dev]) { init(); done[dev] = true; }
    }
};
#endif

size_t jump_smem_per_warp(int A, int C) {
    const int LW = 32 * C;
    int KL = 0;
    while ((1 << (KL + 1)) <= LW) KL++;
    KL += 1;
    const size_t sub_bytes = ((size_t)A * LW * 2 + 15) & ~(size_t)15;
    return sub_bytes + (size_t)KL * LW * 4 + JUMP_GAP_BYTES;
}

// Row kernel (k_ts_jump<C, false, true>): substitution table + gap costs only; evaluation kernel (k_ts_eval): range-minimum levels.
size_t row_smem_per_warp(int A, int C) { return (((size_t)A * 32 * C * 2 + 15) & ~(size_t)15) + JUMP_GAP_BYTES; }
size_t eval_smem_per_warp(int C) {
    const int LW = 32 * C;
    int KL = 0;
    while ((1 << (KL + 1)) <= LW) KL++;
    return (size_t)(KL + 1) * LW * 4;
}
constexpr int MAX_Q_SLICES = 2048;

int jump_warps(int A, int C) {
    const size_t per = jump_smem_per_warp(A, C);
    int w = K2_WARPS;
    while (w > 1 && per * w > (size_t)200 * 1024) w >>= 1;
    return w;
}

template <int C, bool WIN>
void launch_jump(Chunk ck, int stage, const int* d_list, int n_list, int max_len, int A, int n_kinds, int ml, cudaStream_t stream, long long& launches, int gz = 1) {
    ck.win_stage = stage;
    const int warps = jump_warps(A, C);
    const size_t smem = jump_smem_per_warp(A, C) * warps;
#ifndef TSA_EMUL
    static PerDeviceOnce once;
    once.run([&] {
        rt::check(cudaFuncSetAttribute(k_ts_jump<C, WIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)227 * 1024)), "cudaFuncSetAttribute");
        rt::check(cudaFuncSetAttribute(k_ts_jump<C, WIN>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared), "cudaFuncSetAttribute");
        if (getenv("TSA_B200_DEBUG")) {
            int blocks = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k_ts_jump<C, WIN>, 32 * warps, smem);
            cudaFuncAttributes fa;
            cudaFuncGetAttributes(&fa, k_ts_jump<C, WIN>);
            fprintf(stderr, "[tsalign_b200] k_ts_jump<%d,%d>: %d regs, %zu B dynamic smem/block, %d blocks/SM resident\n", C, (int)WIN, fa.numRegs, smem, blocks);
        }
    });
#endif
    const int n_ep = (max_len - ml + 2) / 2;
    if (n_ep <= 0 || n_kinds <= 0) return;
    const int tasks = n_kinds * n_ep;
    const unsigned gx = (unsigned)((tasks + warps - 1) / warps);
    for (int off = 0; off < n_list; off += 65535) {
        const int cnt = std::min(65535, n_list - off);
        auto kern = k_ts_jump<C, WIN>;
        TSA_LAUNCH(kern, dim3(gx, (unsigned)cnt, (unsigned)gz), dim3(32 * warps), smem, stream, ck, d_list + off, cnt);
        launches++;
    }
}

// Pairs without column windows: row kernel -> row queue -> evaluation kernel, in slices of pairs whose candidate rows fit the queue.
// counts[slice] receives the slots each slice asked for (the caller compares them with the capacity after the layer).
template <int C, bool WIN = false>
void launch_jump_split(Chunk ck, int stage, const int* d_list, int n_list, int max_len, int A, int n_kinds, int ml, cudaStream_t stream, long long& launches,
                       int slice_pairs, int* d_counts, int& n_slices, std::vector<int>& slice_size) {
    ck.win_stage = stage;
    const int warps = jump_warps(A, C);
    const size_t smem_row = row_smem_per_warp(A, C) * warps, smem_eval = eval_smem_per_warp(C) * warps;
    int eval_blocks = 2;
#ifndef TSA_EMUL
    static PerDeviceOnce once;
    static std::atomic<int> resident[64];
    int dev = 0;
    cudaGetDevice(&dev);
    once.run([&] {
        auto row = k_ts_jump<C, WIN, true>;
        auto eval = k_ts_eval<C, WIN>;
        rt::check(cudaFuncSetAttribute(row, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)227 * 1024)), "cudaFuncSetAttribute");
        rt::check(cudaFuncSetAttribute(row, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared), "cudaFuncSetAttribute");
        rt::check(cudaFuncSetAttribute(eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)227 * 1024)), "cudaFuncSetAttribute");
        rt::check(cudaFuncSetAttribute(eval, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared), "cudaFuncSetAttribute");
        cudaDeviceProp prop;
        rt::check(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties");
        int per_sm = 0;
        rt::check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, eval, 32 * warps, smem_eval), "occupancy");
        resident[dev & 63] = std::max(1, per_sm) * prop.multiProcessorCount;
        if (getenv("TSA_B200_DEBUG")) {
            cudaFuncAttributes fa, fb;
            cudaFuncGetAttributes(&fa, row); cudaFuncGetAttributes(&fb, eval);
            int row_blocks = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&row_blocks, row, 32 * warps, smem_row);
            fprintf(stderr, "[tsalign_b200] C=%d windows=%d: row kernel %d regs, %d blocks/SM; eval kernel %d regs, %d blocks/SM\n", C, (int)WIN, fa.numRegs, row_blocks, fb.numRegs, per_sm);
        }
    });
    eval_blocks = resident[dev & 63];
#endif
    const int n_ep = (max_len - ml + 2) / 2;
    if (n_ep <= 0 || n_kinds <= 0) return;
    const int tasks = n_kinds * n_ep;
    const unsigned gx = (unsigned)((tasks + warps - 1) / warps);
    slice_pairs = std::max(1, std::min(slice_pairs, 65535));
    for (int off = 0; off < n_list; off += slice_pairs) {
        const int cnt = std::min(slice_pairs, n_list - off);
        if (n_slices >= MAX_Q_SLICES) throw std::runtime_error("row queue: too many slices in one layer");
        ck.q_count = d_counts + n_slices;
        slice_size.push_back(cnt);
        n_slices++;
        auto row = k_ts_jump<C, WIN, true>;
        auto eval = k_ts_eval<C, WIN>;
        TSA_LAUNCH(row, dim3(gx, (unsigned)cnt), dim3(32 * warps), smem_row, stream, ck, d_list + off, cnt);
        TSA_LAUNCH(eval, dim3((unsigned)eval_blocks), dim3(32 * warps), smem_eval, stream, ck);
        launches += 2;
    }
}

}  // namespace

struct Engine::Impl {
    int device = 0;
    cudaStream_t stream = 0;
    std::vector<DevBuf*> dirL, DL;       // per-layer traceback codes / D matrices (traceback only)
    std::vector<DevBuf*> fdL, dir2L;     // flank mode: per-layer codes of the flank planes / seed flags of plane 0
    size_t budget_hint = 0;              // chunk budget of the first large job (the buffers of that size stay allocated)
    int layers_seen = 8;                 // most layers a chunk of this engine has kept so far (chunk sizing)
    void release_layers() {              // after an allocation failure: give the per-layer buffers and the row queue back
        for (auto* v : {&dirL, &DL, &fdL, &dir2L}) { for (DevBuf* b : *v) delete b; v->clear(); }
        q_rows.reset(); q_hdr.reset();
    }
    DevBuf PA, PB, tgt_key, best_plane;  // flank mode: ping-pong state planes, per-layer target keys
    bool flank = false;
    DevBuf ops, ops_off, ops_cap, ops_len, recs, n_recs, tstatus, rows, work_a, work_b, tables;
    DevBuf band, winflag;                             // column windows: band vectors of the fill, overflow flags
    DevBuf cpflag;                                    // ... and the chain-pair bitmaps of the long class
    DevBuf fill_prog;                                 // grid-pipelined primary fill: progress counters per (pair, column block)
    size_t fill_prog_ints = 0;
    size_t cpflag_ints = 0;
    DevBuf q_hdr, q_rows, q_counts;                   // row queue between the row kernel and the evaluation kernel
    size_t q_cap = 0;                                 // slots
    double q_est[N_CLASS][2][3] = {};                 // learned demand of queue slots per pair, by class, first / later deepening round and
                                                      // layer (0, 1, 2+): the demand differs by an order of magnitude between them (sizes the slices)
    bool any_win = false;
    std::vector<int> h_winflag;
    DevBuf wave_prefix, wave_ticket;   // k_affine_wave: first ticket per pair, ticket counter
    DevBuf wave_order;                 // (pair, strip) of every ticket: strip-major inside groups of pairs
    bool wave_ck = false;              // --no-ts with alignments through checkpoints (tsa_band.cuh: k_band_batch_*) instead of a code matrix
    DevBuf bb_args, bb_ckpt, bb_colck, bb_res, bb_tiles, bb_bnds;
    size_t bb_colck_bytes = 0;
    int bb_trace_blocks = 0;
    size_t scratch_bytes = 0;
    int wave_tickets = 0;
    bool wave_ordered = false;
    std::vector<long long> h_ops_off; std::vector<int> h_ops_cap;
    size_t cells = 0, ops_total = 0;
    int max_m = 0, max_n = 0;            // longest query / reference of the staged chunk
    int max_recs = 0;
    DevBuf cfg, lc, meta, seq, D, DT, seedA, seedB, minvec, scratch, best, best_layer, active, next_active, counters, lists, thr, ub, t0, resolved, capped;
    std::vector<PairMeta> metas;
    std::vector<uint8_t> seqpool;
    std::vector<int> status;            // per staged pair (PairStatus)
    std::vector<int> list_all;          // staged pairs that run at all
    std::vector<int> class_list[N_CLASS];
    int class_maxlen[N_CLASS] = {0};
    int* d_list_all = nullptr;
    int* d_class_list[N_CLASS] = {nullptr};
    size_t npairs = 0;
    AlignOptions opt;
    Chunk ck = Chunk();
    bool ts_enabled = false;
    std::vector<int> h_best, h_layer, h_active, h_capped;
    HostBuf h_ops, h_recs, h_small;      // pinned staging of the traceback output
#ifndef TSA_EMUL
    cudaEvent_t ev[4];
#endif
};

size_t Engine::bytes_per_pair(int n, int m) {
    const size_t cells = (size_t)(n + 1) * (m + 1);
    return cells * (2 + 2 + 4 + 4) + (size_t)(n + m + 2) * 4 + (size_t)(n + 1) * 12 + (size_t)n + m + 256;
}

Engine::Engine(const HostConfig& cfg, int device) : impl_(new Impl), host_(cfg) {
    impl_->device = device;
    if (!flatten_config(cfg, dev_, lc_, err_)) return;
#ifndef TSA_EMUL
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { err_ = "no CUDA device available: tsalign_b200 has no CPU path"; return; }
    if (device < 0 || device >= count) { err_ = "invalid CUDA device index"; return; }
    rt::check(cudaSetDevice(device), "cudaSetDevice");
    rt::check(cudaStreamCreateWithFlags(&impl_->stream, cudaStreamNonBlocking), "cudaStreamCreate");
    for (auto& e : impl_->ev) rt::check(cudaEventCreateWithFlags(&e, rt::event_flags()), "cudaEventCreate");
#endif
    impl_->cfg.ensure(sizeof(DevConfig));
    rt::h2d(impl_->cfg.p, &dev_, sizeof(DevConfig), impl_->stream);
    impl_->lc.ensure(lc_.size() * sizeof(int));
    rt::h2d(impl_->lc.p, lc_.data(), lc_.size() * sizeof(int), impl_->stream);
    rt::stream_sync(impl_->stream);
    ok_ = true;
}

Engine::~Engine() {
    for (DevBuf* b : impl_->dirL) delete b;
    for (DevBuf* b : impl_->DL) delete b;
    for (DevBuf* b : impl_->fdL) delete b;
    for (DevBuf* b : impl_->dir2L) delete b;
#ifndef TSA_EMUL
    if (ok_) {
        cudaSetDevice(impl_->device);
        for (auto& e : impl_->ev) cudaEventDestroy(e);
        cudaStreamDestroy(impl_->stream);
    }
#endif
    delete impl_;
}

bool Engine::stage(const PairView* pairs, size_t n, const AlignOptions& opt) {
    Impl& I = *impl_;
#ifndef TSA_EMUL
    rt::check(cudaSetDevice(I.device), "cudaSetDevice");
#endif
    I.opt = opt;
    I.npairs = n;
    I.ts_enabled = !opt.no_ts && dev_.n_kinds > 0;
    I.flank = I.ts_enabled && (dev_.left_flank > 0 || dev_.right_flank > 0);
    I.metas.assign(n, PairMeta());
    I.status.assign(n, PAIR_OK);
    I.list_all.clear();
    I.any_win = false;
    for (auto& l : I.class_list) l.clear();
    for (auto& v : I.class_maxlen) v = 0;
    size_t seq_bytes = 0, cells = 0, vec = 0, scr = 0, tab = 0, cpf_ints = 0, prog_ints = 0;
    I.max_m = 0; I.max_n = 0;
    I.wave_ck = false;
    if (!I.ts_enabled && opt.traceback && n > 0 && opt.wave_checkpoints >= 0) {
        // long pairs: checkpoints; short pairs: the code matrix is small and its walk needs no recomputation
        double range_cells = 0;
        for (size_t i = 0; i < n; i++) range_cells += (double)(pairs[i].rl - pairs[i].ro + 1) * (double)(pairs[i].ql - pairs[i].qo + 1);
        I.wave_ck = opt.wave_checkpoints > 0 || range_cells / (double)n >= (double)(1 << 22);
    }
    for (size_t i = 0; i < n; i++) {
        const PairView& pv = pairs[i];
        PairMeta& pm = I.metas[i];
        I.max_m = std::max(I.max_m, pv.m); I.max_n = std::max(I.max_n, pv.n);
        pm.n = pv.n; pm.m = pv.m; pm.ro = pv.ro; pm.rl = pv.rl; pm.qo = pv.qo; pm.ql = pv.ql;
        pm.seq_r = (long long)seq_bytes; seq_bytes += (size_t)pv.n;
        pm.seq_q = (long long)seq_bytes; seq_bytes += (size_t)pv.m;
        pm.vec = (long long)vec; vec += (size_t)pv.n + pv.m + 2;
        // even: k_affine_wave keeps 8-byte entries there; the multi-warp primary fill keeps one boundary column per warp
#ifdef TSA_EMUL
        const bool wide_fill = I.ts_enabled;     // (the emulator tests run the pipeline on 32-column blocks)
#else
        const bool wide_fill = I.ts_enabled && pv.m + 1 > 32 * K1_CB;
#endif
        pm.scr = (long long)scr; scr += ((3 * ((size_t)pv.n + 1) + 1) & ~(size_t)1) * (wide_fill ? K1_WIDE_MAX : 1);
        pm.mat = (long long)cells;
        pm.tab = -1; pm.lw = 0; pm.cpf = -1; pm.cpf_words = 0;
        pm.prog = (long long)prog_ints;
#ifdef TSA_EMUL
        if (I.ts_enabled) prog_ints += (size_t)pv.m + 2;                                   // (32-column blocks and narrower in the emulator tests)
#else
        if (I.ts_enabled) prog_ints += ((size_t)pv.m + 1 + 32 * K1_GRID_CB - 1) / (32 * K1_GRID_CB);
#endif
        const int W = std::max(pv.n, pv.m) + 1;
        if (I.ts_enabled) {
            if (dev_.left_flank + dev_.right_flank + 1 >= KEY_PLANES) { I.status[i] = PAIR_ERR_FLANKS; continue; }
            int cls = N_CLASS - 1;
            for (int c = 0; c + 1 < N_CLASS; c++) if (W <= 32 * CLASS_C[c] - 1) { cls = c; break; }   // the last column stays "infinite" (RowTable)
#ifdef TSA_EMUL
            if (opt.test_small_windows && W > 48) cls = N_CLASS - 1;
#endif
            if (opt.test_tiled && W > 32) cls = N_CLASS - 1;
            I.class_list[cls].push_back((int)i);
            I.class_maxlen[cls] = std::max(I.class_maxlen[cls], W - 1);
            pm.lw = cls == N_CLASS - 1 ? (W + 7) & ~7 : 32 * CLASS_C[cls];
            pm.win = (cls == 4 && !opt.no_windows) || cls == 5 ? 1 : 0;
            if (cls == N_CLASS - 1 && dev_.ml >= 0 && W - 1 >= dev_.ml) {
                const size_t n_ep = (size_t)((W - 1 - dev_.ml + 2) / 2);
                pm.cpf_words = (int)(((size_t)dev_.n_kinds * n_ep + 31) / 32);
                pm.cpf = (long long)cpf_ints; cpf_ints += 2 * (size_t)pm.cpf_words;
            }
            if (pm.win) I.any_win = true;
            pm.tab = (long long)tab;
            tab += 4 * table_bytes(dev_.A, pm.lw);
        }
        if (I.ts_enabled) cells += (size_t)(pv.n + 1) * (pv.m + 1);
        else if (opt.traceback && !I.wave_ck) cells += (size_t)(pv.rl - pv.ro + 1) * (size_t)wave_dir_stride(pv.ql - pv.qo + 1);   // k_affine_wave: codes of the range, padded rows
        I.list_all.push_back((int)i);
    }
    if (!I.ts_enabled) {
        std::vector<int> prefix(I.list_all.size() + 1, 0);
        for (size_t k = 0; k < I.list_all.size(); k++) {
            const PairView& pv = pairs[I.list_all[k]];
            prefix[k + 1] = prefix[k] + wave_strips(pv.ql - pv.qo + 1);
        }
        I.wave_tickets = prefix.back();
        I.wave_prefix.ensure(prefix.size() * 4);
        I.wave_ticket.ensure(4);
        rt::h2d(I.wave_prefix.p, prefix.data(), prefix.size() * 4, I.stream);
        // Ticket order: strip-major inside groups of pairs (strip 0 of every pair of the group, then strip 1, ...).  The warps that run
        // at the same time then work on DIFFERENT pairs whenever the batch has enough of them, and a strip starts when its left
        // neighbour is (nearly) done instead of running 32 rows behind it: the boundary polls of the (pair, strip) order -- a fifth
        // of the issued instructions and a quarter of the stall samples of the forward pass, profiles/r02_k_band_batch_forward_* --
        // succeed at once.  The boundary column of a pair then lives in L2 / HBM between two of its strips (16 B per row and strip).
        std::vector<int> order;
        if (I.list_all.size() > 1 && !I.opt.pair_major_wave) {
            order.reserve((size_t)prefix.back() * 2);
            const size_t G = 8192;
            for (size_t g0 = 0; g0 < I.list_all.size(); g0 += G) {
                const size_t g1 = std::min(I.list_all.size(), g0 + G);
                int smax = 0;
                for (size_t k = g0; k < g1; k++) smax = std::max(smax, prefix[k + 1] - prefix[k]);
                for (int s = 0; s < smax; s++)
                    for (size_t k = g0; k < g1; k++) if (s < prefix[k + 1] - prefix[k]) { order.push_back((int)k); order.push_back(s); }
            }
            I.wave_order.ensure(order.size() * 4);
            rt::h2d(I.wave_order.p, order.data(), order.size() * 4, I.stream);
        }
        I.wave_ordered = !order.empty();
        rt::stream_sync(I.stream);   // `prefix`, `order` are locals
    }
    size_t bb_bytes = 0;
    std::vector<size_t> bb_ckpt_off, bb_colck_off;
    if (I.wave_ck) {
        // per pair: checkpoint rows (3 ints per column, every BB_INTERVAL rows) and the boundary column entering every column group
        size_t ck_ints = 0, col_entries = 0;
        for (size_t k = 0; k < I.list_all.size(); k++) {
            const PairView& pv = pairs[I.list_all[k]];
            const size_t nn = (size_t)(pv.rl - pv.ro), mm = (size_t)(pv.ql - pv.qo);
            bb_ckpt_off.push_back(ck_ints); bb_colck_off.push_back(col_entries);
            ck_ints += std::max<size_t>(1, nn / BB_INTERVAL) * (mm + 2) * 3;
            col_entries += (size_t)((wave_strips((int)mm + 1) + BB_GROUP - 1) / BB_GROUP) * (nn + 1);
        }
        bb_bytes = ck_ints * 4 + col_entries * 8;
        I.bb_ckpt.ensure(ck_ints * 4); I.bb_colck.ensure(col_entries * 8); I.bb_colck_bytes = col_entries * 8;
        I.bb_args.ensure(I.list_all.size() * sizeof(BandArgs)); I.bb_res.ensure(I.list_all.size() * 8);
    }
    I.cells = cells;
    I.max_recs = std::min(opt.max_layers, MAX_TRACE_LAYERS);
    if (opt.traceback) {
        I.h_ops_off.assign(n, 0); I.h_ops_cap.assign(n, 0);
        size_t off = 0;
        for (size_t i = 0; i < n; i++) {
            I.h_ops_off[i] = (long long)off;
            I.h_ops_cap[i] = 3 * (pairs[i].n + pairs[i].m) + 256;
            off += (size_t)I.h_ops_cap[i];
        }
        I.ops_total = off;
    }
    size_t need = seq_bytes + cells * (I.ts_enabled ? 12 : 0) + (opt.traceback ? cells * 3 + I.ops_total : 0) + vec * 4 + scr * 4 + n * (sizeof(PairMeta) + 32) + bb_bytes;
    if (need > opt.chunk_bytes && n > 1) return false;
    I.seqpool.resize(seq_bytes);
    for (size_t i = 0; i < n; i++) {
        if (pairs[i].n) memcpy(&I.seqpool[(size_t)I.metas[i].seq_r], pairs[i].ref, (size_t)pairs[i].n);
        if (pairs[i].m) memcpy(&I.seqpool[(size_t)I.metas[i].seq_q], pairs[i].qry, (size_t)pairs[i].m);
    }
    I.meta.ensure(n * sizeof(PairMeta));
    I.seq.ensure(seq_bytes);
    I.tables.ensure(tab);
    I.minvec.ensure(vec * 4);
    I.scratch.ensure(scr * 4);
    I.scratch_bytes = scr * 4;
    I.best.ensure(n * 4); I.best_layer.ensure(n * 4); I.active.ensure(n * 4); I.next_active.ensure(n * 4);
    I.counters.ensure(64);
    I.thr.ensure(n * 4); I.ub.ensure(n * 4); I.t0.ensure(n * 4); I.resolved.ensure(n * 4); I.capped.ensure(n * 4);
    if (I.ts_enabled) { I.D.ensure(cells * 2); I.DT.ensure(cells * 2); I.seedA.ensure(cells * 4); I.seedB.ensure(cells * 4); }
    if (I.any_win) { I.band.ensure(vec * 8); I.winflag.ensure(n * 4); I.cpflag.ensure(cpf_ints * 4 + 4); }
    I.cpflag_ints = cpf_ints;
    I.fill_prog_ints = prog_ints;
    if (prog_ints) I.fill_prog.ensure(prog_ints * 4);
    if (I.flank) { I.PA.ensure(cells * 6); I.PB.ensure(cells * 6); I.tgt_key.ensure(n * 4); I.best_plane.ensure(n * 4); }
    if (opt.traceback) {
        I.ops.ensure(I.ops_total); I.ops_off.ensure(n * 8); I.ops_cap.ensure(n * 4); I.ops_len.ensure(n * 4);
        I.recs.ensure(n * (size_t)I.max_recs * sizeof(TsRecord)); I.n_recs.ensure(n * 4); I.tstatus.ensure(n * 4);
        rt::h2d(I.ops_off.p, I.h_ops_off.data(), n * 8, I.stream);
        rt::h2d(I.ops_cap.p, I.h_ops_cap.data(), n * 4, I.stream);
    }
    // pair lists: all, then one per class
    std::vector<int> flat = I.list_all;
    size_t off_class[N_CLASS];
    for (int c = 0; c < N_CLASS; c++) { off_class[c] = flat.size(); flat.insert(flat.end(), I.class_list[c].begin(), I.class_list[c].end()); }
    I.lists.ensure(std::max<size_t>(1, flat.size()) * 4);
    rt::h2d(I.lists.p, flat.data(), flat.size() * 4, I.stream);
    I.d_list_all = I.lists.as<int>();
    for (int c = 0; c < N_CLASS; c++) I.d_class_list[c] = I.lists.as<int>() + off_class[c];
    rt::h2d(I.meta.p, I.metas.data(), n * sizeof(PairMeta), I.stream);
    rt::h2d(I.seq.p, I.seqpool.data(), seq_bytes, I.stream);
    stats_.h2d_bytes = (long long)(n * sizeof(PairMeta) + seq_bytes + flat.size() * 4);
    if (I.wave_ck) {
        std::vector<BandArgs> args(I.list_all.size());
        for (size_t k = 0; k < I.list_all.size(); k++) {
            const PairMeta& pm = I.metas[(size_t)I.list_all[k]];
            BandArgs& ba = args[k];
            memset(&ba, 0, sizeof(ba));
            ba.R = I.seq.as<uint8_t>() + pm.seq_r + pm.ro; ba.Q = I.seq.as<uint8_t>() + pm.seq_q + pm.qo;
            ba.nn = pm.rl - pm.ro; ba.mm = pm.ql - pm.qo;
            ba.s_total = wave_strips(ba.mm + 1);
            ba.s_lo = 0; ba.n_strips = ba.s_total; ba.s_band_first = 0; ba.s_band_last = ba.s_total - 1;
            ba.group = BB_GROUP; ba.row0 = 0; ba.row1 = ba.nn; ba.ck_col0 = 0;
            ba.ckpt_in = nullptr; ba.ckpt_out = I.bb_ckpt.as<int>() + bb_ckpt_off[k];
            ba.interval = BB_INTERVAL; ba.ckpt_stride = ((long long)ba.mm + 2) * 3;
            ba.colck = I.bb_colck.as<WaveBnd>() + bb_colck_off[k]; ba.colck_g0 = 0; ba.store_cols = 1;
            ba.bnd_local = reinterpret_cast<WaveBnd*>(I.scratch.as<int>() + pm.scr); ba.bnd_out = nullptr;
            ba.dir = nullptr; ba.dstride = 0; ba.row_base = 0;
            ba.ticket = nullptr; ba.result = I.bb_res.as<int>() + 2 * k;
        }
        rt::h2d(I.bb_args.p, args.data(), args.size() * sizeof(BandArgs), I.stream);
        rt::stream_sync(I.stream);   // `args` is a local
        stats_.h2d_bytes += (long long)(args.size() * sizeof(BandArgs));
    }

    Chunk& ck = I.ck;
    ck.pairs = I.meta.as<PairMeta>();
    ck.seq = I.seq.as<uint8_t>();
    ck.cfg = I.cfg.as<DevConfig>();
    ck.lc = I.lc.as<int>();
    ck.tables = I.tables.as<unsigned char>();
    ck.D = I.ts_enabled ? I.D.as<int16_t>() : nullptr;
    ck.DT = I.ts_enabled ? I.DT.as<int16_t>() : nullptr;
    ck.dir = nullptr;
    ck.seedA = I.seedA.as<int>();
    ck.seedB = I.seedB.as<int>();
    ck.minvec = I.minvec.as<int>();
    ck.scratch = I.scratch.as<int>();
    ck.best = I.best.as<int>();
    ck.best_layer = I.best_layer.as<int>();
    ck.active = I.active.as<int>();
    ck.next_active = I.next_active.as<int>();
    ck.counters = I.counters.as<int>();
    ck.thr = I.thr.as<int>(); ck.ub = I.ub.as<int>(); ck.t0 = I.t0.as<int>(); ck.resolved = I.resolved.as<int>(); ck.capped = I.capped.as<int>();
    ck.round = 0;
    ck.kind_mask = ~0u;
    ck.pl_in = nullptr; ck.pl_out = nullptr; ck.dir2 = nullptr;
    ck.cells_total = (long long)cells;
    ck.tgt_key = I.tgt_key.as<int>(); ck.best_plane = I.best_plane.as<int>();
    ck.flank_mode = I.flank ? 1 : 0;
    ck.band = I.any_win ? I.band.as<int>() : nullptr;
    ck.winflag = I.any_win ? I.winflag.as<int>() : nullptr;
    ck.cpflag = I.any_win && I.cpflag_ints ? I.cpflag.as<int>() : nullptr;
    ck.fill_prog = I.fill_prog_ints ? I.fill_prog.as<int>() : nullptr;
    ck.win_stage = 0;
    ck.seeds_merged = 0;
    rt::stream_sync(I.stream);
    return true;
}

void Engine::run_staged() {
    Impl& I = *impl_;
#ifndef TSA_EMUL
    rt::check(cudaSetDevice(I.device), "cudaSetDevice");
#endif
    stats_.launches = stats_.fill_launches = stats_.jump_launches = 0;
    stats_.chains_run = stats_.rows_filled = stats_.rows_jumped = stats_.chains_started = 0;
    stats_.layers_run = 0; stats_.rounds_run = 0;
    stats_.jump_ms = stats_.fill_ms = 0;
    const int n_all = (int)I.list_all.size();
    if (n_all == 0) return;
    const size_t k1_smem = (size_t)K1_WARPS * K1_SMEM_INTS * sizeof(int);
    rt::dev_memset(I.active.p, 0, I.npairs * 4, I.stream);
    rt::dev_memset(I.next_active.p, 0, I.npairs * 4, I.stream);
    rt::dev_memset(I.capped.p, 0, I.npairs * 4, I.stream);
    if (!I.ts_enabled) { run_wave(); run_trace(); return; }
    if (I.any_win) rt::dev_memset(I.winflag.p, 0, I.npairs * 4, I.stream);
    if (I.any_win && I.cpflag_ints) rt::dev_memset(I.cpflag.p, 0, I.cpflag_ints * 4, I.stream);
#ifndef TSA_EMUL
    auto mark = [&](int k) { rt::check(cudaEventRecord(I.ev[k], I.stream), "cudaEventRecord"); };
    auto span = [&](int a, int b) { float ms = 0; rt::check(cudaEventSynchronize(I.ev[b]), "cudaEventSynchronize"); cudaEventElapsedTime(&ms, I.ev[a], I.ev[b]); return (double)ms; };
#else
    auto mark = [&](int) {};
    auto span = [&](int, int) { return 0.0; };
#endif
    auto fill = [&](const int* d_list, int cnt, int layer) {
        if (I.opt.traceback) {
            // keep every layer's traceback codes (and D, the template-switch entrance costs) until the traceback
            while ((int)I.dirL.size() <= layer) { I.dirL.push_back(new DevBuf); I.DL.push_back(new DevBuf); }
            I.dirL[layer]->ensure(I.cells);
            I.ck.dir = I.dirL[layer]->as<uint8_t>();
            if (I.ts_enabled) { I.DL[layer]->ensure(I.cells * 2); I.ck.D = I.DL[layer]->as<int16_t>(); }
        }
        // Pairs of several column blocks: one CTA per pair, the blocks pipelined over its warps; the reentry seeds of both
        // orientations are merged into seedA first (coalesced tiles on all SMs) so that the fill reads one contiguous row.
        auto wide = [&](auto kern, int warps) {
            Chunk ckm = I.ck;
            if (layer > 0 && !I.ck.pl_in) {
                const int tiles = ((I.max_n + 32) / 32) * ((I.max_m + 32) / 32);
                for (int off = 0; off < cnt; off += 65535) {
                    const int c2 = std::min(65535, cnt - off);
                    TSA_LAUNCH(k_merge_seeds, dim3((unsigned)std::min(tiles, 2048), (unsigned)c2), dim3(256), (size_t)32 * 33 * sizeof(int), I.stream, I.ck, d_list + off, c2);
                    stats_.launches++;
                }
                ckm.seeds_merged = 1;
            }
            TSA_LAUNCH(kern, dim3((unsigned)cnt), dim3(32 * warps), (size_t)(warps * K1_SMEM_INTS + 1) * sizeof(int), I.stream, ckm, d_list, cnt, layer);
        };
        // One warp per column block over the whole device (very long pairs: a layer of one pair is otherwise filled by one SM)
        auto grid_pipe = [&](auto kern, int cb) {
            Chunk ckm = I.ck;
            if (layer > 0 && !I.ck.pl_in) {
                const int tiles = ((I.max_n + 32) / 32) * ((I.max_m + 32) / 32);
                for (int off = 0; off < cnt; off += 65535) {
                    const int c2 = std::min(65535, cnt - off);
                    TSA_LAUNCH(k_merge_seeds, dim3((unsigned)std::min(tiles, 2048), (unsigned)c2), dim3(256), (size_t)32 * 33 * sizeof(int), I.stream, I.ck, d_list + off, c2);
                    stats_.launches++;
                }
                ckm.seeds_merged = 1;
            }
            rt::dev_memset(I.fill_prog.p, 0, I.fill_prog_ints * 4, I.stream);
            const int nb = (I.max_m + 1 + 32 * cb - 1) / (32 * cb);
            for (int off = 0; off < cnt; off += 65535) {
                const int c2 = std::min(65535, cnt - off);
                TSA_LAUNCH(kern, dim3((unsigned)nb, (unsigned)c2), dim3(32), (size_t)(K1_SMEM_INTS + 1) * sizeof(int), I.stream, ckm, d_list + off, c2, layer);
            }
        };
#ifdef TSA_EMUL
        if ((I.opt.test_small_windows || I.opt.test_tiled) && I.ts_enabled && I.max_m + 1 > 128) grid_pipe(k_primary_fill<1, 0>, 1);   // CPU tests: 32-column blocks
        else if ((I.opt.test_small_windows || I.opt.test_tiled) && I.ts_enabled && I.max_m + 1 > 32) wide(k_primary_fill<1, 4>, 4);
        else
#endif
        if (I.max_m + 1 > 32 * K1_CB * K1_WIDE_MAX && !I.opt.narrow_fill && !I.opt.cta_fill) grid_pipe(k_primary_fill<K1_GRID_CB, 0>, K1_GRID_CB);
        else
        if (I.max_m + 1 <= 32 * 5) TSA_LAUNCH(k_primary_fill<5>, dim3((unsigned)((cnt + K1_WARPS - 1) / K1_WARPS)), dim3(32 * K1_WARPS), k1_smem, I.stream, I.ck, d_list, cnt, layer);
        else if (I.max_m + 1 > 32 * K1_CB * 4 && !I.opt.narrow_fill) wide(k_primary_fill<K1_CB, K1_WIDE_MAX>, K1_WIDE_MAX);
        else if (I.max_m + 1 > 32 * K1_CB && !I.opt.narrow_fill) wide(k_primary_fill<K1_CB, 4>, 4);
        else TSA_LAUNCH(k_primary_fill<K1_CB>, dim3((unsigned)((cnt + K1_WARPS - 1) / K1_WARPS)), dim3(32 * K1_WARPS), k1_smem, I.stream, I.ck, d_list, cnt, layer);
        stats_.launches++; stats_.fill_launches++;
    };
    // One layer = the primary fill plus, with flank lengths > 0, the right-flank planes before it (reentry seeds ->
    // ordinary plane) and the left-flank planes after it (ordinary plane -> template switch entrance plane).
    auto fill_layer = [&](const int* d_list, int cnt, int layer) {
        if (!I.flank) { fill(d_list, cnt, layer); return; }
        const int RF = dev_.right_flank, LF = dev_.left_flank;
        const size_t cells = I.cells;
        uint8_t* fd = nullptr;
        I.ck.dir2 = nullptr;
        if (I.opt.traceback) {
            while ((int)I.fdL.size() <= layer) { I.fdL.push_back(new DevBuf); I.dir2L.push_back(new DevBuf); }
            I.fdL[layer]->ensure((size_t)(RF + LF) * cells);
            I.dir2L[layer]->ensure(cells);
            fd = I.fdL[layer]->as<uint8_t>();
            I.ck.dir2 = I.dir2L[layer]->as<uint8_t>();
        }
        int16_t* A = I.PA.as<int16_t>();
        int16_t* B = I.PB.as<int16_t>();
        // blocks per pair: enough to fill the GPU when the batch is small (1 kb pairs: ~250 blocks each), 16 for read pairs
        const unsigned gx = (unsigned)std::min<size_t>(1024, std::max<size_t>(16, (cells / std::max<size_t>(1, I.npairs)) / 4096));
        // `count` flank moves from plane `plane0` (in *src) in launches of up to FLANK_FS moves; the result ends up in *src
        const int tiles_x = (I.max_m + 1 + FLANK_FT - 1) / FLANK_FT, tiles_y = (I.max_n + 1 + FLANK_FT - 1) / FLANK_FT;
        auto run = [&](int16_t*& src, int16_t*& dst, int table, int plane0, int count, int final_at_end, int no_report_plane) {
#ifndef TSA_EMUL
            static PerDeviceOnce once;
            once.run([&] { rt::check(cudaFuncSetAttribute(k_flank_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FLANK_SMEM), "cudaFuncSetAttribute"); });
#endif
            for (int done = 0; done < count; done += FLANK_FS) {
                const int ns = std::min(FLANK_FS, count - done);
                for (int off = 0; off < cnt; off += 65535) {
                    const int c2 = std::min(65535, cnt - off);
                    TSA_LAUNCH(k_flank_fused, dim3((unsigned)(tiles_x * tiles_y), (unsigned)c2), dim3(256), FLANK_SMEM, I.stream, I.ck, d_list + off, c2, src, dst, fd, table,
                               plane0 + done, ns, (final_at_end && done + ns == count) ? 1 : 0, no_report_plane, layer, tiles_x);
                    stats_.launches++;
                }
                std::swap(src, dst);
            }
        };
        I.ck.pl_in = nullptr;
        if (layer > 0 && RF > 0) {
            for (int off = 0; off < cnt; off += 65535) {
                const int c2 = std::min(65535, cnt - off);
                TSA_LAUNCH(k_seed_to_plane, dim3(gx, (unsigned)c2), dim3(256), 0, I.stream, I.ck, d_list + off, c2, A);
                stats_.launches++;
            }
            run(A, B, 4, 0, RF, 0, RF);   // the target in plane RF is reported by the primary fill
            I.ck.pl_in = A;
        }
        I.ck.pl_out = LF > 0 ? B : nullptr;
        fill(d_list, cnt, layer);
        if (LF > 0) {
            // the jump kernel reads D of the entrance plane L_f: written (with its row / column minima) by the last step
            for (int off = 0; off < cnt; off += 65535) {
                const int c2 = std::min(65535, cnt - off);
                TSA_LAUNCH(k_reset_minvec, dim3(4, (unsigned)c2), dim3(256), 0, I.stream, I.ck, d_list + off, c2, layer);
                stats_.launches++;
            }
            int16_t* src = B;
            int16_t* dst = A;
            run(src, dst, 3, RF, LF, 1, -1);
        }
        TSA_LAUNCH(k_layer_finish, dim3((unsigned)((cnt + 255) / 256)), dim3(256), 0, I.stream, I.ck, d_list, cnt, layer);
        stats_.launches++;
    };
    int n_ts = 0;
    for (int c = 0; c < N_CLASS; c++) n_ts += (int)I.class_list[c].size();
    // the TS-enabled pairs are exactly the union of the class lists, which are contiguous after list_all
    const int* d_ts_list = I.d_class_list[0];
    const unsigned ts_grid = (unsigned)((n_ts + 255) / 256);
    I.ck.round = -1;
    if (n_ts) { TSA_LAUNCH(k_resolve, dim3(ts_grid), dim3(256), 0, I.stream, I.ck, d_ts_list, n_ts, I.opt.first_threshold, (int*)nullptr, 0); stats_.launches++; }
    I.ck.round = 0;
    if (I.flank) rt::dev_memset(I.tgt_key.p, 0x7f, I.npairs * 4, I.stream);
    mark(0);
    fill_layer(I.d_list_all, n_all, 0);
    mark(1);
    if (!I.ts_enabled || n_ts == 0) {
        rt::dev_memset(I.active.p, 0, I.npairs * 4, I.stream);   // no jump follows: nothing stays active
        rt::stream_sync(I.stream); stats_.fill_ms += span(0, 1); run_trace(); return;
    }
    bool fill_pending = true;   // events 0..1 bracket a fill that has not been read yet
    for (int off = 0; off < n_ts; off += 65535) {   // per-column cost tables of every (pair, secondary, direction)
        const int cnt = std::min(65535, n_ts - off);
        TSA_LAUNCH(k_prepare_tables, dim3(4, (unsigned)cnt), dim3(128), 0, I.stream, I.ck, d_ts_list + off, cnt);
        stats_.launches++;
    }

    // Work lists per jump-kernel class, compacted on the device after every layer / round: cur[c] = the pairs of
    // class c that are still active.  Two scratch list buffers with the layout of the class lists, swapped per step.
    I.work_a.ensure((size_t)std::max(1, n_ts) * 4); I.work_b.ensure((size_t)std::max(1, n_ts) * 4);
    int class_off[N_CLASS], cur_n[N_CLASS];
    const int* cur[N_CLASS];
    {
        int off = 0;
        for (int c = 0; c < N_CLASS; c++) { class_off[c] = off; off += (int)I.class_list[c].size(); }
    }
    for (int c = 0; c < N_CLASS; c++) { cur[c] = I.d_class_list[c]; cur_n[c] = (int)I.class_list[c].size(); }
    int* spare[2] = {I.work_a.as<int>(), I.work_b.as<int>()};
    int which = 0;
    auto read_counts = [&](int* counts8) {
        int h[16] = {0};
        rt::d2h(h, I.counters.p, sizeof(h), I.stream);
        rt::stream_sync(I.stream);
        stats_.chains_run += h[1]; stats_.rows_filled += h[2]; stats_.rows_jumped += h[3]; stats_.chains_started += h[4];
        if (getenv("TSA_B200_DEBUG")) fprintf(stderr, "[tsalign_b200] layer: chains %d/%d rows %d queued %d dominated %d evaluated %d\n", h[1], h[4], h[2], h[5], h[6], h[3]);
        for (int c = 0; c < N_CLASS; c++) counts8[c] = h[8 + c];
        if (h[7]) throw std::runtime_error("primary fill: a column block waited for its left neighbour for too long");
    };
    // Row queue of the split jump (classes without column windows): sized from the learned demand per pair, at least the worst case
    // of one pair (every row of every chain queued), at most 6 GiB.
    int n_slices = 0, est_r = 0, est_l = 0;      // bucket of the learned demand the current layer uses
    std::vector<int> slice_size, slice_class, slice_lw, h_qcounts;
    auto est = [&](int c) -> double& { return I.q_est[c][est_r][est_l]; };
    {
        size_t want = 0, one = 0;
        int min_lw = 1 << 30;
        for (int c = 0; c < N_CLASS; c++) {
            if (I.class_list[c].empty()) continue;
            // (the first window stage of the medium and long classes queues 544-column rows; long pairs only use the queue there)
            const int LW = c == N_CLASS - 1 ? 32 * 17 : 32 * CLASS_C[c], mx = I.class_maxlen[c];
            const int n_ep = std::max(0, (mx - dev_.ml + 2) / 2);
            const double chains = (double)dev_.n_kinds * n_ep;
            const double rows_max = (double)std::min(dev_.lmax, mx) + 1 + QUEUE_RESERVE;
            for (auto& byround : I.q_est[c]) for (double& v : byround) if (v <= 0) v = chains * 2.0;
            // a long pair whose candidate rows do not fit falls back to the fused windowed kernel (jump_layer)
            one = std::max(one, c == N_CLASS - 1 ? std::min((size_t)(chains * rows_max) * LW * 4, (size_t)2 << 30) : (size_t)(chains * rows_max) * LW * 4);
            want += (size_t)((double)I.class_list[c].size() * I.q_est[c][0][0] * 1.3) * LW * 4;
            min_lw = std::min(min_lw, c >= 4 ? 32 * 17 : LW);
        }
        if (min_lw < (1 << 30)) {
            size_t cap_bytes = (size_t)16 << 30;      // fewer, larger slices are faster (measured: 6 / 12 / 24 GiB)
#ifndef TSA_EMUL
            { size_t free_b = 0, total_b = 0; if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) cap_bytes = std::min(cap_bytes, std::max<size_t>((size_t)1 << 28, (free_b + I.q_rows.cap) / 4)); }
#endif
            if (const char* qm = getenv("TSA_B200_QUEUE_MB")) cap_bytes = std::max<size_t>(1, (size_t)atoll(qm)) << 20;   // developer knob: slice size of the split jump
            const size_t bytes = std::max(one, std::min(want, cap_bytes)) + 4096;
            I.q_rows.ensure(bytes);
            I.q_hdr.ensure((I.q_rows.cap / ((size_t)min_lw * 4) + 1) * sizeof(QueueHdr));
            I.q_counts.ensure((size_t)MAX_Q_SLICES * 4);
        }
    }
    // cw: columns per lane of the instantiation (0: the class's own); win: the windowed kernels (k_ts_jump<C, true, true> / k_ts_eval<C, true>)
    auto jump_split = [&](int c, int stage, long long& l, int cw = 0, bool win = false) {
        Chunk ck = I.ck;
        const int CW = cw ? cw : CLASS_C[c];
        const int LW = 32 * CW;
        ck.q_cap = (int)std::min<size_t>(I.q_rows.cap / ((size_t)LW * 4), (size_t)1 << 30) & ~(QUEUE_RESERVE - 1);   // whole reservation blocks
        ck.q_hdr = I.q_hdr.as<QueueHdr>();
        ck.q_rows = I.q_rows.as<uint32_t>();
        const int slice_pairs = (int)std::max(1.0, std::min(65535.0, (double)ck.q_cap / (est(c) * 1.3)));
        const int ml_ = dev_.ml, A_ = dev_.A, nk = dev_.n_kinds, mx = I.class_maxlen[c];
        int* counts = I.q_counts.as<int>();
        if (win) {
            switch (CW) {
#ifdef TSA_EMUL
            case 3: launch_jump_split<3, true>(ck, stage, cur[c], cur_n[c], mx, A_, nk, ml_, I.stream, l, slice_pairs, counts, n_slices, slice_size); break;
#endif
            case 17: launch_jump_split<17, true>(ck, stage, cur[c], cur_n[c], mx, A_, nk, ml_, I.stream, l, slice_pairs, counts, n_slices, slice_size); break;
            default: throw std::runtime_error("no windowed split jump of this width");
            }
        } else
        switch (CW) {
        case 3: launch_jump_split<3>(ck, stage, cur[c], cur_n[c], mx, A_, nk, ml_, I.stream, l, slice_pairs, counts, n_slices, slice_size); break;
        case 5: launch_jump_split<5>(ck, stage, cur[c], cur_n[c], mx, A_, nk, ml_, I.stream, l, slice_pairs, counts, n_slices, slice_size); break;
        case 9: launch_jump_split<9>(ck, stage, cur[c], cur_n[c], mx, A_, nk, ml_, I.stream, l, slice_pairs, counts, n_slices, slice_size); break;
        case 17: launch_jump_split<17>(ck, stage, cur[c], cur_n[c], mx, A_, nk, ml_, I.stream, l, slice_pairs, counts, n_slices, slice_size); break;
        default: launch_jump_split<33>(ck, stage, cur[c], cur_n[c], mx, A_, nk, ml_, I.stream, l, slice_pairs, counts, n_slices, slice_size); break;
        }
        slice_class.resize(slice_size.size(), c);
        slice_lw.resize(slice_size.size(), LW);
    };
    bool force_fused[N_CLASS] = {false};      // the candidate rows of ONE pair of the class overflowed the queue: fused windowed kernel from then on
    auto jump_class = [&](int c) {
        long long l = 0;
        const int ml_ = dev_.ml, A_ = dev_.A, nk = dev_.n_kinds, mx = I.class_maxlen[c];
        switch (c) {
        case 0: case 1: case 2: case 3: jump_split(c, 0, l); break;
        case 4:   // medium: windows of 544 columns, then the whole sequences for the pairs that were flagged
            if (I.opt.no_windows) { jump_split(c, 0, l); break; }
            if (I.opt.fused_windows || force_fused[c]) launch_jump<17, true>(I.ck, 1, cur[c], cur_n[c], mx, A_, nk, ml_, I.stream, l);
            else jump_split(c, 1, l, 17, true);
            jump_split(c, 2, l);
            break;
        default:  // long: windows of 544, then of 1056 columns
            if (I.opt.test_tiled) {   // developer knob: only the tiled stage, with narrow sub-ranges
#ifdef TSA_EMUL
                launch_jump<5, true>(I.ck, 4, cur[c], cur_n[c], mx, A_, nk, ml_, I.stream, l, 2);
#else
                launch_jump<33, true>(I.ck, 4, cur[c], cur_n[c], mx, A_, nk, ml_, I.stream, l, 3);
#endif
                break;
            }
#ifdef TSA_EMUL
            if (I.opt.test_small_windows) {
                if (I.opt.fused_windows) launch_jump<3, true>(I.ck, 1, cur[c], cur_n[c], mx, A_, nk, ml_, I.stream, l);
                else jump_split(c, 1, l, 3, true);
                launch_jump<5, true>(I.ck, 2, cur[c], cur_n[c], mx, A_, nk, ml_, I.stream, l);
                launch_jump<5, true>(I.ck, 3, cur[c], cur_n[c], mx, A_, nk, ml_, I.stream, l, 2);
                break;
            }
#endif
            if (I.opt.fused_windows || force_fused[c]) launch_jump<17, true>(I.ck, 1, cur[c], cur_n[c], mx, A_, nk, ml_, I.stream, l);
            else jump_split(c, 1, l, 17, true);
            launch_jump<33, true>(I.ck, 2, cur[c], cur_n[c], mx, A_, nk, ml_, I.stream, l);
            // third stage: the entrance columns of a chain pair in sub-ranges (any width; seeds combine through atomicMin)
            launch_jump<33, true>(I.ck, 3, cur[c], cur_n[c], mx, A_, nk, ml_, I.stream, l, std::max(1, std::min(4, mx / 700)));
            break;
        }
        stats_.launches += l; stats_.jump_launches += l;
    };
    // The jump of one layer.  A slice of pairs that asked for more queue slots than there are dropped rows: the learned demand is
    // updated from the exact count and the jump of the layer is repeated with smaller slices (seeds are minima: repeating is exact).
    auto jump_layer = [&]() {
        for (int attempt = 0;; attempt++) {
            n_slices = 0; slice_size.clear(); slice_class.clear(); slice_lw.clear();
            if (I.q_counts.p) rt::dev_memset(I.q_counts.p, 0, (size_t)MAX_Q_SLICES * 4, I.stream);
            mark(2);
            for (int c = 0; c < N_CLASS; c++) if (cur_n[c]) jump_class(c);
            mark(3);
            if (n_slices == 0) return;
            h_qcounts.resize((size_t)n_slices);
            rt::d2h(h_qcounts.data(), I.q_counts.p, (size_t)n_slices * 4, I.stream);
            rt::stream_sync(I.stream);
            bool overflow = false;
            for (int s2 = 0; s2 < n_slices; s2++) {
                const int c = slice_class[(size_t)s2];
                const size_t cap = std::min<size_t>(I.q_rows.cap / ((size_t)slice_lw[(size_t)s2] * 4), (size_t)1 << 30);
                est(c) = std::max(est(c), (double)h_qcounts[(size_t)s2] / std::max(1, slice_size[(size_t)s2]));
                if ((size_t)h_qcounts[(size_t)s2] > cap) {
                    overflow = true;
                    if (slice_size[(size_t)s2] == 1 && slice_lw[(size_t)s2] < 32 * CLASS_C[c] && !force_fused[c]) force_fused[c] = true;
                    else if (slice_size[(size_t)s2] == 1 || attempt > 8) throw std::runtime_error("row queue: one pair does not fit the queue");
                }
            }
            if (getenv("TSA_B200_DEBUG")) fprintf(stderr, "[tsalign_b200] jump layer: attempt %d, %d slices, overflow %d, q_est %.0f\n", attempt, n_slices, (int)overflow, est(1));
            if (!overflow) return;
        }
    };
    // Scouting round: the reverse kinds (anti-diagonal geometry, no trivial self matches) are cheap to evaluate and
    // usually already contain the optimum.  Running them alone first gives the full rounds a tight upper bound, so
    // that the expensive forward kinds are pruned with ub = best + 1 instead of the deepening threshold.
    unsigned full_mask = 0, rev_mask = 0;
    for (int k = 0; k < dev_.n_kinds; k++) {
        if (I.opt.primary_filter != 0 && dev_.kinds[k].p != I.opt.primary_filter - 1) continue;   // descendant strategy: one primary only
        full_mask |= 1u << k; if (dev_.kinds[k].d == 1) rev_mask |= 1u << k;
    }
    const bool scout = I.opt.scout_round && rev_mask != 0 && rev_mask != full_mask;
    const unsigned clear_gx = (unsigned)std::min<size_t>(512, std::max<size_t>(8, (I.cells / std::max<size_t>(1, I.npairs)) / 8192));   // blocks per pair
    for (int round = 0;; round++) {
        I.ck.round = round;
        I.ck.kind_mask = (scout && round == 0) ? rev_mask : full_mask;
        if (round > 0) {   // layer 0 again for the unresolved pairs
            mark(0);
            for (int c = 0; c < N_CLASS; c++) if (cur_n[c]) fill_layer(cur[c], cur_n[c], 0);
            mark(1);
            fill_pending = true;
        }
        for (int layer = 0;; layer++) {
            rt::dev_memset(I.counters.p, 0, 64, I.stream);
            for (int c = 0; c < N_CLASS; c++)
                for (int off = 0; off < cur_n[c]; off += 65535) {
                    const int cnt = std::min(65535, cur_n[c] - off);
                    TSA_LAUNCH(k_clear_seeds, dim3(clear_gx, (unsigned)cnt), dim3(256), 0, I.stream, I.ck, cur[c] + off, cnt);
                    stats_.launches++;
                }
            est_r = round > 0 ? 1 : 0; est_l = std::min(layer, 2);
            jump_layer();
            int* out = spare[which];
            for (int c = 0; c < N_CLASS; c++) if (cur_n[c]) {
                TSA_LAUNCH(k_advance, dim3((unsigned)((cur_n[c] + 255) / 256)), dim3(256), 0, I.stream, I.ck, cur[c], cur_n[c], out + class_off[c], 8 + c);
                stats_.launches++;
            }
            int counts[N_CLASS];
            read_counts(counts);
            if (fill_pending) { stats_.fill_ms += span(0, 1); fill_pending = false; }
            stats_.jump_ms += span(2, 3);
            stats_.layers_run = std::max(stats_.layers_run, layer + 1);
            int total = 0;
            for (int c = 0; c < N_CLASS; c++) { cur[c] = out + class_off[c]; cur_n[c] = counts[c]; total += counts[c]; }
            which ^= 1;
            if (total == 0) break;
            if (layer + 1 >= I.opt.max_layers) {
                // layer cap: the still-active pairs are refused and leave the deepening loop; the other pairs of the chunk go on
                for (int c = 0; c < N_CLASS; c++) if (cur_n[c]) {
                    TSA_LAUNCH(k_cap, dim3((unsigned)((cur_n[c] + 255) / 256)), dim3(256), 0, I.stream, I.ck, cur[c], cur_n[c]);
                    stats_.launches++;
                }
                break;
            }
            mark(0);
            for (int c = 0; c < N_CLASS; c++) if (cur_n[c]) fill_layer(cur[c], cur_n[c], layer + 1);
            mark(1);
            fill_pending = true;
        }
        stats_.rounds_run = round + 1;
        if (scout && round == 0) {   // nothing is proven yet: every pair goes into the first full round, same threshold
            for (int c = 0; c < N_CLASS; c++) { cur[c] = I.d_class_list[c]; cur_n[c] = (int)I.class_list[c].size(); }
            continue;
        }
        // which pairs still have to prove their optimum with a larger threshold?
        rt::dev_memset(I.counters.p, 0, 64, I.stream);
        int* out = spare[which];
        for (int c = 0; c < N_CLASS; c++) {
            const int cnt = (int)I.class_list[c].size();
            if (!cnt) continue;
            TSA_LAUNCH(k_resolve, dim3((unsigned)((cnt + 255) / 256)), dim3(256), 0, I.stream, I.ck, I.d_class_list[c], cnt, 0, out + class_off[c], 8 + c);
            stats_.launches++;
        }
        int counts[N_CLASS];
        read_counts(counts);
        int total = 0;
        for (int c = 0; c < N_CLASS; c++) { cur[c] = out + class_off[c]; cur_n[c] = counts[c]; total += counts[c]; }
        which ^= 1;
        if (total == 0) break;
    }
    run_trace();
}

// --no-ts: one launch of the wavefront kernel over all strips of all pairs (tsa_wave.cuh).
void Engine::run_wave() {
    Impl& I = *impl_;
    const int n_all = (int)I.list_all.size();
    if (n_all == 0 || I.wave_tickets == 0) return;
    if (I.opt.traceback && !I.wave_ck) {
        while (I.dirL.empty()) { I.dirL.push_back(new DevBuf); I.DL.push_back(new DevBuf); }
        I.dirL[0]->ensure(I.cells);
        I.ck.dir = I.dirL[0]->as<uint8_t>();
    } else I.ck.dir = nullptr;
    rt::dev_memset(I.scratch.p, 0xff, I.scratch_bytes, I.stream);   // boundary entries: tag 7 = "not written"
    rt::dev_memset(I.wave_ticket.p, 0, 4, I.stream);
    WaveArgs wa;
    wa.list = I.d_list_all; wa.n_list = n_all; wa.strip_prefix = I.wave_prefix.as<int>();
    wa.ticket = I.wave_ticket.as<int>();
    wa.order = I.wave_ordered ? I.wave_order.as<int>() : nullptr;
    const size_t smem = (size_t)WAVE_SMEM_INTS * sizeof(int);
    int blocks = (I.wave_tickets + WAVE_WARPS - 1) / WAVE_WARPS;
#ifndef TSA_EMUL
    static std::atomic<int> resident(0);
    if (!resident) {
        cudaDeviceProp prop;
        rt::check(cudaGetDeviceProperties(&prop, I.device), "cudaGetDeviceProperties");
        int per_sm = 0;
        rt::check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_affine_wave<true>, 32 * WAVE_WARPS, smem), "occupancy");
        resident = std::max(1, per_sm) * prop.multiProcessorCount;
        if (getenv("TSA_B200_DEBUG")) {
            cudaFuncAttributes fa, fb;
            cudaFuncGetAttributes(&fa, k_affine_wave<true>); cudaFuncGetAttributes(&fb, k_affine_wave<false>);
            fprintf(stderr, "[tsalign_b200] k_affine_wave: %d / %d regs (trace / costs), %d blocks/SM resident\n", fa.numRegs, fb.numRegs, per_sm);
        }
    }
    blocks = std::min(blocks, resident.load());   // persistent: every warp takes strips from the ticket until none is left
    rt::check(cudaEventRecord(I.ev[0], I.stream), "cudaEventRecord");
#else
    blocks = std::min(blocks, 2);
#endif
    if (I.wave_ck) {
        // forward pass with checkpoints (tsa_band.cuh); boundary columns start as "not written"
        rt::dev_memset(I.bb_colck.p, 0xff, I.bb_colck_bytes, I.stream);
        std::vector<int> res_init(2 * (size_t)n_all, 0);
        for (int k = 0; k < n_all; k++) res_init[2 * (size_t)k] = INF32;
        rt::h2d(I.bb_res.p, res_init.data(), res_init.size() * 4, I.stream);
        rt::stream_sync(I.stream);   // `res_init` is a local
        BandBatch bb;
        bb.args = I.bb_args.as<BandArgs>(); bb.pair_of = I.d_list_all; bb.prefix = I.wave_prefix.as<int>(); bb.n_pairs = n_all; bb.ticket = I.wave_ticket.as<int>();
        bb.order = I.wave_ordered ? I.wave_order.as<int>() : nullptr;
        TSA_LAUNCH(k_band_batch<false>, dim3((unsigned)blocks), dim3(32 * WAVE_WARPS), smem, I.stream, I.cfg.as<DevConfig>(), bb);
        TSA_LAUNCH(k_band_batch_finish, dim3((unsigned)((n_all + 255) / 256)), dim3(256), 0, I.stream, I.ck, bb);
        stats_.launches += 2; stats_.fill_launches++;
    } else {
    if (I.opt.traceback) TSA_LAUNCH(k_affine_wave<true>, dim3((unsigned)blocks), dim3(32 * WAVE_WARPS), smem, I.stream, I.ck, wa);
    else TSA_LAUNCH(k_affine_wave<false>, dim3((unsigned)blocks), dim3(32 * WAVE_WARPS), smem, I.stream, I.ck, wa);
    stats_.launches++; stats_.fill_launches++;
    }
#ifndef TSA_EMUL
    rt::check(cudaEventRecord(I.ev[1], I.stream), "cudaEventRecord");
    rt::check(cudaEventSynchronize(I.ev[1]), "cudaEventSynchronize");
    float ms = 0; cudaEventElapsedTime(&ms, I.ev[0], I.ev[1]);
    stats_.fill_ms += ms;
#endif
}

template <int C, bool WIN>
static void launch_trace(const Chunk& ck, const TraceLayers& tl, TraceOut to, DevBuf& rows, const int* d_list, int n_list, int rows_max, int A,
                         cudaStream_t stream, long long& launches) {
    if (n_list <= 0) return;
    const int warps = jump_warps(A, C);
    const size_t smem = jump_smem_per_warp(A, C) * warps;
#ifndef TSA_EMUL
    static PerDeviceOnce once;
    once.run([&] { rt::check(cudaFuncSetAttribute(k_traceback<C, WIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)227 * 1024)), "cudaFuncSetAttribute"); });
#endif
    const long long stride = (long long)(rows_max + 1) * 3 * 32 * C;   // shorts per warp
    const long long budget = (long long)1 << 29;                        // 1 GiB of shorts-pairs scratch per slice
    const int slice = (int)std::max<long long>(warps, std::min<long long>(n_list, budget / std::max<long long>(1, stride)));
    rows.ensure((size_t)slice * (size_t)stride * 2);
    to.rows = rows.as<int16_t>();
    to.rows_stride = stride;
    for (int off = 0; off < n_list; off += slice) {
        const int cnt = std::min(slice, n_list - off);
        auto kern = k_traceback<C, WIN>;
        TSA_LAUNCH(kern, dim3((unsigned)((cnt + warps - 1) / warps)), dim3(32 * warps), smem, stream, ck, tl, to, d_list + off, cnt);
        launches++;
    }
}

void Engine::run_trace() {
    Impl& I = *impl_;
    if (!I.opt.traceback) return;
    TraceLayers tl;
    memset(&tl, 0, sizeof(tl));
    for (size_t k = 0; k < I.dirL.size() && k <= (size_t)MAX_TRACE_LAYERS; k++) { tl.dir[k] = I.dirL[k]->as<uint8_t>(); tl.D[k] = I.DL[k]->as<int16_t>(); }
    for (size_t k = 0; k < I.fdL.size() && k <= (size_t)MAX_TRACE_LAYERS; k++) { tl.fd[k] = I.fdL[k]->as<uint8_t>(); tl.dir2[k] = I.dir2L[k]->as<uint8_t>(); }
    tl.cells_total = (long long)I.cells;
    tl.rf = I.flank ? dev_.right_flank : 0; tl.lf = I.flank ? dev_.left_flank : 0;
    tl.wave = I.ts_enabled ? 0 : 1;
    TraceOut to;
    memset(&to, 0, sizeof(to));
    to.ops = I.ops.as<uint8_t>(); to.ops_off = I.ops_off.as<long long>(); to.ops_cap = I.ops_cap.as<int>(); to.ops_len = I.ops_len.as<int>();
    to.recs = I.recs.as<TsRecord>(); to.max_recs = I.max_recs; to.n_recs = I.n_recs.as<int>(); to.status = I.tstatus.as<int>();
    long long l = 0;
    if (!I.ts_enabled && I.wave_ck) {
        const int n_all = (int)I.list_all.size();
        int blocks = std::max(1, std::min((n_all + WAVE_WARPS - 1) / WAVE_WARPS, 4 * 148));
#ifdef TSA_EMUL
        blocks = std::min(blocks, 2);
#endif
        const size_t warps = (size_t)blocks * WAVE_WARPS;
        BandBatchTrace bt;
        bt.tile_bytes = (long long)(BB_INTERVAL + 1) * BB_GROUP * WAVE_SW;
        I.bb_tiles.ensure(warps * (size_t)bt.tile_bytes); I.bb_bnds.ensure(warps * (size_t)(BB_INTERVAL + 1) * sizeof(WaveBnd));
        bt.tiles = I.bb_tiles.as<uint8_t>(); bt.bnds = I.bb_bnds.as<WaveBnd>();
        BandBatch bb;
        bb.args = I.bb_args.as<BandArgs>(); bb.pair_of = I.d_list_all; bb.prefix = I.wave_prefix.as<int>(); bb.n_pairs = n_all; bb.ticket = I.wave_ticket.as<int>(); bb.order = nullptr;
        if (n_all) { TSA_LAUNCH(k_band_batch_trace, dim3((unsigned)blocks), dim3(32 * WAVE_WARPS), (size_t)WAVE_SMEM_INTS * sizeof(int), I.stream, I.cfg.as<DevConfig>(), I.ck, bb, bt, to); l++; }
    } else if (!I.ts_enabled) {
        launch_trace<3, false>(I.ck, tl, to, I.rows, I.d_list_all, (int)I.list_all.size(), 0, dev_.A, I.stream, l);
    } else {
        for (int c = 0; c < N_CLASS; c++) {
            const int cnt = (int)I.class_list[c].size();
            if (!cnt) continue;
            const int rows_max = std::min(dev_.lmax, I.class_maxlen[c]);
            switch (c) {
            case 0: launch_trace<3, false>(I.ck, tl, to, I.rows, I.d_class_list[c], cnt, rows_max, dev_.A, I.stream, l); break;
            case 1: launch_trace<5, false>(I.ck, tl, to, I.rows, I.d_class_list[c], cnt, rows_max, dev_.A, I.stream, l); break;
            case 2: launch_trace<9, false>(I.ck, tl, to, I.rows, I.d_class_list[c], cnt, rows_max, dev_.A, I.stream, l); break;
            case 3: launch_trace<17, false>(I.ck, tl, to, I.rows, I.d_class_list[c], cnt, rows_max, dev_.A, I.stream, l); break;
            case 4: launch_trace<33, false>(I.ck, tl, to, I.rows, I.d_class_list[c], cnt, rows_max, dev_.A, I.stream, l); break;
            default:
#ifdef TSA_EMUL
                if (I.opt.test_small_windows || I.opt.test_tiled) { launch_trace<5, true>(I.ck, tl, to, I.rows, I.d_class_list[c], cnt, rows_max, dev_.A, I.stream, l); break; }
#endif
                launch_trace<33, true>(I.ck, tl, to, I.rows, I.d_class_list[c], cnt, rows_max, dev_.A, I.stream, l); break;
            }
        }
    }
    stats_.launches += l; stats_.trace_launches = l;
    rt::stream_sync(I.stream);
}

void Engine::fetch_staged(PairCost* out) {
    Impl& I = *impl_;
    const size_t n = I.npairs;
    I.h_best.resize(n); I.h_layer.resize(n); I.h_active.resize(n);
    rt::d2h(I.h_best.data(), I.best.p, n * 4, I.stream);
    rt::d2h(I.h_layer.data(), I.best_layer.p, n * 4, I.stream);
    rt::d2h(I.h_active.data(), I.active.p, n * 4, I.stream);
    I.h_capped.resize(n);
    rt::d2h(I.h_capped.data(), I.capped.p, n * 4, I.stream);
    std::vector<int> h_sat;
    if (!I.ts_enabled) { h_sat.resize(n); rt::d2h(h_sat.data(), I.next_active.p, n * 4, I.stream); }   // k_affine_wave: boundary values saturated
    I.h_winflag.assign(n, 0);
    if (I.any_win && I.ts_enabled) rt::d2h(I.h_winflag.data(), I.winflag.p, n * 4, I.stream);
    rt::stream_sync(I.stream);
    stats_.d2h_bytes = (long long)(n * 12);
    for (size_t i = 0; i < n; i++) {
        PairCost& pc = out[i];
        pc = PairCost();
        pc.status = I.status[i];
        if (pc.status != PAIR_OK) continue;
        if (I.h_winflag[i] & 4) { pc.status = PAIR_ERR_TOO_LONG; continue; }   // the hulls of the cost model alone exceed the widest lane grid
        if (!I.ts_enabled && h_sat[i] && I.h_best[i] >= WAVE_SAT) { pc.status = PAIR_ERR_COST_RANGE; continue; }   // exact only below 2^26 - 1
        if (I.ts_enabled && (I.h_active[i] || I.h_capped[i])) { pc.status = PAIR_ERR_LAYER_CAP; continue; }
        if (I.h_best[i] >= INF32) { pc.status = PAIR_NO_TARGET; continue; }
        // The jump kernel computes in saturating s16: every path cheaper than INF16 is exact, so a result below
        // INF16 is the optimum; above it a cheaper template-switch path may have been saturated away.
        if (I.ts_enabled && I.h_best[i] >= INF16 - 1) { pc.status = PAIR_ERR_COST_RANGE; continue; }
        pc.status = PAIR_OK;
        pc.cost = I.h_best[i];
        pc.layers = I.h_layer[i];
    }
    if (!I.opt.traceback) return;
    const size_t n_recs_total = n * (size_t)I.max_recs;
    I.h_small.ensure(n * 12); I.h_ops.ensure(I.ops_total); I.h_recs.ensure(n_recs_total * sizeof(TsRecord));
    int* len = I.h_small.as<int>(); int* nrec = len + n; int* tst = nrec + n;
    uint8_t* ops = I.h_ops.as<uint8_t>();
    TsRecord* recs = I.h_recs.as<TsRecord>();
    rt::d2h(len, I.ops_len.p, n * 4, I.stream);
    rt::d2h(nrec, I.n_recs.p, n * 4, I.stream);
    rt::d2h(tst, I.tstatus.p, n * 4, I.stream);
    rt::d2h(ops, I.ops.p, I.ops_total, I.stream);
    rt::d2h(recs, I.recs.p, n_recs_total * sizeof(TsRecord), I.stream);
    rt::stream_sync(I.stream);
    stats_.d2h_bytes += (long long)(n * 12 + I.ops_total + n_recs_total * sizeof(TsRecord));
    for (size_t i = 0; i < n; i++) {
        PairCost& pc = out[i];
        if (pc.status != PAIR_OK) continue;
        pc.trace_status = tst[i];
        if (tst[i] != TRACE_OK) continue;
        const uint8_t* src = ops + I.h_ops_off[i];
        pc.ops.assign(std::reverse_iterator<const uint8_t*>(src + len[i]), std::reverse_iterator<const uint8_t*>(src));
        const TsRecord* rs = recs + i * (size_t)I.max_recs;
        pc.recs.assign(std::reverse_iterator<const TsRecord*>(rs + nrec[i]), std::reverse_iterator<const TsRecord*>(rs));
    }
}

void Engine::align_costs(const PairView* pairs, size_t n, const AlignOptions& opt, PairCost* out) {
    // Greedy chunking in input order under the HBM budget.
    size_t i = 0;
    EngineStats total;
    size_t budget = opt.chunk_bytes;
    if (budget == 0) {
        // small jobs fit anyway: ask the driver for the free memory (a slow call) only when it matters
        size_t rough = 0;
        const size_t per_cell = 40 + (size_t)(dev_.left_flank + dev_.right_flank) * 8;
        for (size_t k = 0; k < n && rough <= ((size_t)32 << 30); k++) rough += (size_t)(pairs[k].n + 1) * (pairs[k].m + 1) * per_cell + 4096;
        if (rough <= ((size_t)32 << 30)) budget = (size_t)64 << 30;
    }
    if (budget == 0) {
#ifndef TSA_EMUL
        size_t free_b = 0, total_b = 0;
        rt::check(cudaSetDevice(impl_->device), "cudaSetDevice");
        rt::check(cudaMemGetInfo(&free_b, &total_b), "cudaMemGetInfo");
        // a second engine may be working on the other half of the batch; what this engine sized its buffers for earlier is still its own
        // (a third of what is free once the row queues of two engines, up to 16 GiB each, are set aside)
        budget = std::max<size_t>(std::max<size_t>((size_t)1 << 30, (free_b - std::min<size_t>(free_b / 4, (size_t)32 << 30)) / 3), impl_->budget_hint);
        impl_->budget_hint = budget;
#else
        budget = (size_t)1 << 30;
#endif
    }
    // --no-ts with alignments: long pairs keep checkpoint rows + boundary columns instead of a code matrix (stage() decides the same
    // way per chunk; the decision of the whole job is handed down so that every chunk is sized for what it will allocate)
    bool job_ck = false;
    if ((opt.no_ts || dev_.n_kinds == 0) && opt.traceback && n > 0 && opt.wave_checkpoints >= 0) {
        double range_cells = 0;
        for (size_t k = 0; k < n; k++) range_cells += (double)(pairs[k].rl - pairs[k].ro + 1) * (double)(pairs[k].ql - pairs[k].qo + 1);
        job_ck = opt.wave_checkpoints > 0 || range_cells / (double)n >= (double)(1 << 22);
    }
    auto resident_bytes = [&](const PairView& p) {
        const size_t cells = (size_t)(p.n + 1) * (p.m + 1);
        size_t b = (!opt.no_ts && dev_.n_kinds > 0) ? bytes_per_pair(p.n, p.m) : (size_t)(p.n + p.m) * 20 + 512;
        if (job_ck) {
            const size_t nn = (size_t)(p.rl - p.ro), mm = (size_t)(p.ql - p.qo);
            return b + std::max<size_t>(1, nn / BB_INTERVAL) * (mm + 2) * 12 + (size_t)((wave_strips((int)mm + 1) + BB_GROUP - 1) / BB_GROUP) * (nn + 1) * 8
                     + 3 * (size_t)(p.n + p.m) + 256;
        }
        // codes (+ D) of every layer a pair of the chunk reaches (3 B per cell and layer; the deepest chunk so far decides), ops
        if (opt.traceback) b += cells * ((!opt.no_ts && dev_.n_kinds > 0) ? 3 * (size_t)(impl_->layers_seen + 1) : 1) + 3 * (size_t)(p.n + p.m) + 256;
        if (!opt.no_ts && dev_.n_kinds > 0 && (dev_.left_flank > 0 || dev_.right_flank > 0))   // flank planes, and their codes of ~4 layers
            b += cells * (12 + (opt.traceback ? (size_t)(dev_.left_flank + dev_.right_flank + 1) * 4 : 0));
        return b;
    };
    // chunks of about equal size: as few as the budget allows, none much smaller than the others (a small last chunk runs at
    // a fraction of the throughput of a full one)
    size_t target = budget;
    {
        size_t all = 0;
        for (size_t k = 0; k < n; k++) all += resident_bytes(pairs[k]);
        if (all > budget) target = all / ((all + budget - 1) / budget) + 1;
    }
    while (i < n) {
        if (opt.memory_limit_strict && resident_bytes(pairs[i]) > budget) {
            // generic_a_star/src/lib.rs:380-389: the search gives up when its store outgrows the limit.  (The C ABI re-runs pairs
            // without template switches through the checkpointed path of tsa_long.cu, which needs no code matrix.)
            out[i] = PairCost(); out[i].status = PAIR_MEMORY_LIMIT;
            i++;
            continue;
        }
        size_t j = i, bytes = 0;
        while (j < n) {
            const size_t b = resident_bytes(pairs[j]);
            if (j > i && bytes + b > budget) break;
            bytes += b; j++;
            if (bytes >= target) break;
        }
        const auto cs = std::chrono::steady_clock::now();
        AlignOptions o = opt;
        if (job_ck) o.wave_checkpoints = std::max(o.wave_checkpoints, 1);
        o.chunk_bytes = (size_t)1 << 62;   // the chunk was sized above; stage() must not refuse it
        std::chrono::steady_clock::time_point c0, c1;
        bool refused = false;
        for (;;) {
            // The layers a pair needs are only known afterwards: a chunk whose pairs go deeper than estimated may not fit.  It is
            // then redone at half the size with the per-layer buffers released (nothing of it was handed out yet).
            try {
                if (!stage(pairs + i, j - i, o)) { refused = true; break; }
                c0 = std::chrono::steady_clock::now();
                run_staged();
                c1 = std::chrono::steady_clock::now();
                fetch_staged(out + i);
                break;
            } catch (const std::runtime_error& e) {
                if (std::string(e.what()).find("out of memory") == std::string::npos || j - i < 2) throw;
#ifndef TSA_EMUL
                cudaGetLastError();
                cudaStreamSynchronize(impl_->stream);
#endif
                impl_->release_layers();
                if (getenv("TSA_B200_DEBUG")) fprintf(stderr, "[tsalign_b200] chunk of %zu pairs did not fit (%s): redone at half the size\n", j - i, e.what());
                j = i + (j - i) / 2;
            }
        }
        if (refused) {
            for (size_t k = i; k < j; k++) { out[k] = PairCost(); out[k].status = PAIR_ERR_TOO_LONG; }
            i = j;
            continue;
        }
        impl_->layers_seen = std::max(impl_->layers_seen, stats_.layers_run);
        if (getenv("TSA_B200_DEBUG"))
            fprintf(stderr, "[tsalign_b200] chunk of %zu pairs: stage %.2f ms, run %.2f ms, fetch %.2f ms\n", j - i,
                    1e3 * std::chrono::duration<double>(c0 - cs).count(), 1e3 * std::chrono::duration<double>(c1 - c0).count(),
                    1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - c1).count());
        total.launches += stats_.launches; total.fill_launches += stats_.fill_launches; total.jump_launches += stats_.jump_launches;
        total.layers_run = std::max(total.layers_run, stats_.layers_run);
        total.h2d_bytes += stats_.h2d_bytes; total.d2h_bytes += stats_.d2h_bytes;
        i = j;
    }
    stats_ = total;
}

// ---- integer roofline probe: back-to-back DPX add-min on every SM ---------------------------------------------
// 8 independent dependency chains per thread so that the issue rate, not the latency, is measured.
TSA_KERNEL void k_addmin_probe(uint32_t* out, int iters, int packed) {
    uint32_t a[8];
    const uint32_t t = (uint32_t)(blockIdx.x * blockDim.x + threadIdx.x);
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = (t * 2654435761u + (uint32_t)k * 40503u) & 0x0fff0fffu;
    const uint32_t b = (t & 7u) | ((t & 3u) << 16);
    if (packed) {
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int k = 0; k < 8; k++) a[k] = addmin_s16x2(a[k], b, a[(k + 3) & 7]);
        }
    } else {
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int k = 0; k < 8; k++) a[k] = (uint32_t)addmin_s32((int)a[k], (int)b, (int)a[(k + 3) & 7]);
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) r ^= a[k];
    if (r == 0x12345678u) out[t & 1023] = r;  // keeps the chains alive
}

bool measure_addmin_peak(int device, double* s16x2_lane_ops_per_s, double* s32_lane_ops_per_s, std::string& err) {
#ifndef TSA_EMUL
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) { err = "no CUDA device"; return false; }
    rt::check(cudaSetDevice(device), "cudaSetDevice");
    cudaDeviceProp prop;
    rt::check(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
    uint32_t* d = (uint32_t*)rt::dev_alloc(4096);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096, threads = 256, blocks = prop.multiProcessorCount * 8;
    for (int packed = 1; packed >= 0; packed--) {
        double best = 0;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            k_addmin_probe<<<blocks, threads>>>(d, iters, packed);
            cudaEventRecord(e1);
            rt::check(cudaEventSynchronize(e1), "probe");
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            const double ops = (double)blocks * threads * iters * 8 * (packed ? 2 : 1);
            if (rep > 0) best = std::max(best, ops / (ms * 1e-3));
        }
        *(packed ? s16x2_lane_ops_per_s : s32_lane_ops_per_s) = best;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    rt::dev_free(d);
    return true;
#else
    (void)device; *s16x2_lane_ops_per_s = 0; *s32_lane_ops_per_s = 0; err = "emulator"; return false;
#endif
}

}  // namespace tsa
