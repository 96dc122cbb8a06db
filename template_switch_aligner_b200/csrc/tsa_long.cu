// tsa_long.cu -- see tsa_long.hpp.  Compiled by nvcc for sm_100a (product) or by g++ with -DTSA_EMUL (tests/emul only).
#include "tsa_long.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>
#include <stdexcept>

#include "tsa_band.cuh"
#include "tsa_engine.hpp"

namespace tsa {

namespace {
// Device buffers of a LongPair come from a small per-device cache and go back to it: cudaMalloc / cudaFree of a few hundred MB cost
// tens to hundreds of milliseconds in a process that holds other large allocations -- as much as the alignment of the pair itself.
struct BufCache {
    struct Item { int device; void* p; size_t cap; };
    std::mutex lock;
    std::vector<Item> items;
    size_t total = 0;
    static BufCache& get() { static BufCache c; return c; }
    static int device() {
#ifndef TSA_EMUL
        int d = 0; cudaGetDevice(&d); return d;
#else
        return 0;
#endif
    }
    void* take(size_t bytes, size_t& cap) {
        const int dev = device();
        {
            std::lock_guard<std::mutex> g(lock);
            int best = -1;
            for (size_t k = 0; k < items.size(); k++)
                if (items[k].device == dev && items[k].cap >= bytes && items[k].cap <= 2 * bytes + (1 << 20) && (best < 0 || items[k].cap < items[(size_t)best].cap)) best = (int)k;
            if (best >= 0) { Item it = items[(size_t)best]; items.erase(items.begin() + best); total -= it.cap; cap = it.cap; return it.p; }
        }
        cap = bytes;
        return rt::dev_alloc(bytes);
    }
    void give(void* p, size_t cap) {
        if (!p) return;
        std::vector<Item> drop;
        {
            std::lock_guard<std::mutex> g(lock);
            items.push_back(Item{device(), p, cap});
            total += cap;
            while (total > ((size_t)2 << 30) && !items.empty()) { drop.push_back(items.front()); total -= items.front().cap; items.erase(items.begin()); }   // at most 2 GiB kept
        }
        for (const Item& it : drop) rt::dev_free(it.p);
    }
};
struct Buf {
    void* p = nullptr;
    size_t cap = 0, used = 0;      // capacity of the (possibly cached, larger) allocation; bytes this object asked for
    void ensure(size_t bytes) { if (bytes > cap) { BufCache::get().give(p, cap); p = BufCache::get().take(bytes, cap); } used = std::max(used, bytes); }
    ~Buf() { BufCache::get().give(p, cap); }
    template <class T> T* as() const { return static_cast<T*>(p); }
};
constexpr long long SMALL_INTS = 16;   // [0] ticket, [2..3] result, [4] ops_len, [8..] WalkState
}  // namespace

// Half width of the strip of tiles around the straight line from the root to the target that speculate() recomputes ahead of the walk.
static int spec_margin(int nn, int mm) { return std::max(nn, mm) / 64 + WAVE_SW; }

int BandPlan::owner_of_column(int j) const {
    const int gi = (j < 0 ? 0 : j / WAVE_SW) / group;
    int r = (int)(((long long)gi + 1) * world / n_groups);   // first guess, then correct
    r = std::min(std::max(r, 0), world - 1);
    while (r > 0 && group_first(r) > gi) r--;
    while (r + 1 < world && group_first(r + 1) <= gi) r++;
    return r;
}

long long BandPlan::resident_bytes(int rank, bool traceback) const {
    const long long strips = strip_last(rank) - strip_first(rank) + 1;
    const long long bw = std::min<long long>((long long)(strip_last(rank) + 1) * WAVE_SW, (long long)mm + 1) - (long long)strip_first(rank) * WAVE_SW;
    long long b = (long long)nn + mm + 4096 + sizeof(DevConfig);                       // sequences, config
    b += ((long long)nn + 1) * 8;                                                       // boundary between the strips of a group
    const long long groups = (strips + group - 1) / group;
    if (traceback) {
        b += groups * ((long long)nn + 1) * 8;                                          // boundary column entering every group
        b += (long long)(nn / interval) * (bw + 1) * 12;                                // checkpoint rows
        b += ((long long)interval + 1) * (long long)group * WAVE_SW;                    // codes of one tile
        b += (long long)nn + mm + 64;                                                   // unit ops
    } else {
        b += ((long long)nn + 1) * 8;                                                   // incoming boundary only
    }
    return b;
}

BandPlan plan_bands(int nn, int mm, int world, int interval, int group, size_t memory_limit, bool traceback) {
    BandPlan best;
    best.nn = nn; best.mm = mm; best.world = world;
    best.s_total = wave_strips(mm + 1);
    if (world < 1 || best.s_total < world) { best.why = "fewer 256-column strips than devices"; return best; }
    // candidates: the given values, or the grid of powers of two; among those that fit the limit on every rank the one whose
    // traceback recomputes the fewest cells (a path crosses ~nn / interval + columns / (256 group) tiles of interval x 256 group cells)
    std::vector<int> cand_i, cand_g;
    if (interval > 0) cand_i.push_back(interval); else for (int v = 256; v <= 16384; v <<= 1) cand_i.push_back(v);
    const int g_cap = std::max(1, best.s_total / world);
    if (group > 0) cand_g.push_back(std::min(group, g_cap)); else for (int v = 1; v <= 64; v <<= 1) if (v <= g_cap) cand_g.push_back(v);
    double best_score = -1;
    for (int ci : cand_i) for (int cg : cand_g) {
        BandPlan p = best;
        p.interval = ci; p.group = cg;
        p.n_groups = (p.s_total + cg - 1) / cg;
        if (p.n_groups < world) continue;
        long long worst = 0;
        for (int r = 0; r < world; r++) worst = std::max(worst, p.resident_bytes(r, traceback));
        if (memory_limit && (unsigned long long)worst > memory_limit) continue;
        // time model of the traceback (measured on a B200, profiles/r02_c5_*).  A path crosses ~nn / interval + columns / (256 group)
        // tiles.  If the code buffers of the tiles around the diagonal fit (speculate()), they are recomputed in one launch -- the
        // slower of one tile's wavefront latency and their cells at ~400 GCUPS -- and every crossed tile costs one host round trip of
        // the walk; else a crossed tile costs a round trip plus its own wavefront latency.
        const double crossed = (double)nn / ci + (double)mm / ((double)cg * WAVE_SW) + world;
        const double margin = (double)spec_margin(nn, mm);
        const double diag_tiles = ((double)nn / ci + 1.0) * (std::floor(((double)ci * mm / std::max(nn, 1) + 2.0 * margin) / ((double)cg * WAVE_SW)) + 2.0);
        const double spec_bytes = diag_tiles * ((double)ci + 1.0) * cg * WAVE_SW / world;
        const bool can_spec = spec_bytes <= (double)((size_t)16 << 30) && (memory_limit == 0 || (double)worst + spec_bytes <= (double)memory_limit);
        const double t_tile = 120.0 + std::max(0.16 * ((double)ci + 32.0 * cg), (double)ci * cg * WAVE_SW / 4.0e5);   // microseconds
        const double score = can_spec ? crossed * 150.0 + std::max(0.4 * ((double)ci + 40.0 * cg), spec_bytes / 4.0e5) : crossed * t_tile;
        if (best_score < 0 || score < best_score) { best_score = score; p.ok = true; best = p; }
    }
    if (!best.ok) best.why = "no checkpoint spacing fits the memory limit";
    return best;
}

struct LongPair::Impl {
    cudaStream_t stream = 0;
    Buf cfg, R, Q, colck, bnd_local, ckpt, tile, ops, small;
    Buf sp_codes, sp_bnds, sp_args, sp_prefix;          // speculate(): codes / boundaries / descriptions of the tiles recomputed ahead
    std::vector<std::pair<int, int>> sp_tiles;          // (row block, first strip of the group) of every speculated tile, in buffer order
    size_t sp_tile_bytes = 0;
    bool sp_pending = false;
    WaveBnd* bnd_out = nullptr;
    int s_first = 0, s_last = 0, g_first = 0, n_groups = 0, bw = 0;
    long long ckpt_stride = 0, dstride = 0;
    int resident_blocks = 1;
    size_t ops_cap = 0;
#ifndef TSA_EMUL
    cudaEvent_t ev[2];
#endif
};

LongPair::LongPair(const HostConfig& cfg, int device, const uint8_t* R, const uint8_t* Q, const BandPlan& plan, int rank, bool traceback)
    : impl_(new Impl), plan_(plan), device_(device), rank_(rank), traceback_(traceback) {
    Impl& I = *impl_;
    if (!plan.ok) { err_ = plan.why; return; }
    if (!flatten_config(cfg, dev_, lc_, err_)) return;
#ifndef TSA_EMUL
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { err_ = "no CUDA device available: tsalign_b200 has no CPU path"; return; }
    if (device < 0 || device >= count) { err_ = "invalid CUDA device index"; return; }
    rt::check(cudaSetDevice(device), "cudaSetDevice");
    rt::check(cudaStreamCreateWithFlags(&I.stream, cudaStreamNonBlocking), "cudaStreamCreate");
    for (auto& e : I.ev) rt::check(cudaEventCreateWithFlags(&e, rt::event_flags()), "cudaEventCreate");
    {
        cudaDeviceProp prop;
        rt::check(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
        int per_sm = 0;
        rt::check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_affine_band<true>, 32 * WAVE_WARPS, (size_t)WAVE_SMEM_INTS * sizeof(int)), "occupancy");
        I.resident_blocks = std::max(1, per_sm) * prop.multiProcessorCount;
    }
#else
    I.resident_blocks = 2;
#endif
    const int nn = plan.nn, mm = plan.mm;
    I.s_first = plan.strip_first(rank); I.s_last = plan.strip_last(rank);
    I.g_first = plan.group_first(rank); I.n_groups = plan.group_first(rank + 1) - I.g_first;
    col_first_ = I.s_first * WAVE_SW;
    I.bw = std::min((I.s_last + 1) * WAVE_SW, mm + 1) - col_first_;
    I.cfg.ensure(sizeof(DevConfig));
    I.R.ensure((size_t)nn + 16); I.Q.ensure((size_t)mm + 16);
    rt::h2d(I.cfg.p, &dev_, sizeof(DevConfig), I.stream);
    if (nn) rt::h2d(I.R.p, R, (size_t)nn, I.stream);
    if (mm) rt::h2d(I.Q.p, Q, (size_t)mm, I.stream);
    const size_t col_bytes = (size_t)(nn + 1) * 8;
    I.bnd_local.ensure(col_bytes);
    I.colck.ensure(col_bytes * (size_t)(traceback ? I.n_groups : 1));
    I.small.ensure(SMALL_INTS * 4 + sizeof(WalkState));
    if (traceback) {
        I.ckpt_stride = ((long long)I.bw + 1) * 3;
        I.ckpt.ensure((size_t)std::max(1, nn / plan.interval) * (size_t)I.ckpt_stride * 4);
        I.dstride = (long long)plan.group * WAVE_SW;
        I.tile.ensure((size_t)(plan.interval + 1) * (size_t)I.dstride);
        I.ops_cap = (size_t)nn + mm + 64;
        I.ops.ensure(I.ops_cap);
    }
    // boundary entries start as "not written" (tag 4095); the rank on the left may start writing as soon as every rank is built
    rt::dev_memset(I.colck.p, 0xff, I.colck.used, I.stream);
    rt::dev_memset(I.bnd_local.p, 0xff, col_bytes, I.stream);
    rt::stream_sync(I.stream);
    stats_.resident_bytes = (long long)(I.cfg.used + I.R.used + I.Q.used + I.bnd_local.used + I.colck.used + I.small.used + I.ckpt.used + I.tile.used + I.ops.used);
    stats_.interval = plan.interval; stats_.group = plan.group;
    ok_ = true;
}

LongPair::~LongPair() {
#ifndef TSA_EMUL
    if (impl_->stream) {
        cudaSetDevice(device_);
        for (auto& e : impl_->ev) cudaEventDestroy(e);
        cudaStreamDestroy(impl_->stream);
    }
#endif
    delete impl_;
}

void* LongPair::incoming_boundary() const { return impl_->colck.p; }
void LongPair::set_outgoing_boundary(void* remote) { impl_->bnd_out = static_cast<WaveBnd*>(remote); }

void LongPair::forward_launch() {
    Impl& I = *impl_;
#ifndef TSA_EMUL
    rt::check(cudaSetDevice(device_), "cudaSetDevice");
#endif
    int init[SMALL_INTS] = {0};
    init[2] = INF32;
    rt::h2d(I.small.p, init, sizeof(init), I.stream);
    rt::stream_sync(I.stream);   // `init` is a local
    BandArgs ba;
    memset(&ba, 0, sizeof(ba));
    ba.R = I.R.as<uint8_t>(); ba.Q = I.Q.as<uint8_t>(); ba.nn = plan_.nn; ba.mm = plan_.mm;
    ba.s_lo = I.s_first; ba.n_strips = I.s_last - I.s_first + 1; ba.s_total = plan_.s_total; ba.s_band_first = I.s_first; ba.s_band_last = I.s_last;
    ba.group = traceback_ ? plan_.group : (1 << 30);   // costs only: no boundary column but the incoming one is kept
    ba.row0 = 0; ba.row1 = plan_.nn;
    ba.ck_col0 = col_first_;
    ba.ckpt_in = nullptr; ba.ckpt_out = traceback_ ? I.ckpt.as<int>() : nullptr;
    ba.interval = plan_.interval; ba.ckpt_stride = I.ckpt_stride;
    ba.colck = I.colck.as<WaveBnd>(); ba.colck_g0 = traceback_ ? I.g_first : 0; ba.store_cols = traceback_ ? 1 : 0;
    ba.bnd_local = I.bnd_local.as<WaveBnd>(); ba.bnd_out = I.bnd_out;
    ba.dir = nullptr; ba.dstride = 0; ba.row_base = 0;
    ba.ticket = I.small.as<int>(); ba.result = I.small.as<int>() + 2;
    const int blocks = std::min(I.resident_blocks, (ba.n_strips + WAVE_WARPS - 1) / WAVE_WARPS);
#ifndef TSA_EMUL
    rt::check(cudaEventRecord(I.ev[0], I.stream), "cudaEventRecord");
#endif
    TSA_LAUNCH(k_affine_band<false>, dim3((unsigned)blocks), dim3(32 * WAVE_WARPS), (size_t)WAVE_SMEM_INTS * sizeof(int), I.stream, I.cfg.as<DevConfig>(), ba);
#ifndef TSA_EMUL
    rt::check(cudaEventRecord(I.ev[1], I.stream), "cudaEventRecord");
#endif
    if (I.bnd_out) stats_.boundary_bytes_out = ((long long)plan_.nn + 1) * 8;
}

void LongPair::forward_wait() {
    Impl& I = *impl_;
#ifndef TSA_EMUL
    rt::check(cudaSetDevice(device_), "cudaSetDevice");
    rt::check(cudaEventSynchronize(I.ev[1]), "forward pass");
    float ms = 0; cudaEventElapsedTime(&ms, I.ev[0], I.ev[1]);
    stats_.forward_ms = ms;
#endif
    int h[SMALL_INTS];
    rt::d2h(h, I.small.p, sizeof(h), I.stream);
    rt::stream_sync(I.stream);
    saturated_ = h[3] != 0;
    if (has_target()) cost_ = h[2];
}

void LongPair::speculate(size_t budget_bytes) {
    Impl& I = *impl_;
    if (!traceback_ || !ok_) return;
#ifndef TSA_EMUL
    rt::check(cudaSetDevice(device_), "cudaSetDevice");
#endif
    const int nn = plan_.nn, mm = plan_.mm, IV = plan_.interval, G = plan_.group;
    const size_t tile_bytes = (size_t)(IV + 1) * (size_t)I.dstride;
    const size_t per_tile = tile_bytes + (size_t)(IV + 1) * sizeof(WaveBnd) + sizeof(BandArgs) + 8;
    const size_t max_tiles = std::min<size_t>(budget_bytes / std::max<size_t>(1, per_tile), 4096);
    // tiles that the straight line from (0, 0) to (nn, mm) crosses, widened by 1/64 of the pair + one strip on either side
    std::vector<std::pair<int, int>> tiles;
    const int margin = spec_margin(nn, mm);
    for (int k = 0; k * IV < std::max(nn, 1); k++) {
        const long long r0 = (long long)k * IV, r1 = std::min<long long>(r0 + IV, nn);
        const long long j_lo = std::max<long long>(0, (nn ? r0 * mm / nn : 0) - margin), j_hi = std::min<long long>(mm, (nn ? r1 * mm / nn : mm) + margin);
        for (int g = (int)(j_lo / WAVE_SW) / G; g <= (int)(j_hi / WAVE_SW) / G; g++) {
            const int s_lo = g * G;
            if (s_lo < I.s_first || s_lo > I.s_last) continue;      // another band's tile
            tiles.emplace_back(k, s_lo);
        }
    }
    // the tiles nearest to the target first (the walk starts there), as many as the budget holds
    std::reverse(tiles.begin(), tiles.end());
    if (tiles.size() > max_tiles) tiles.resize(max_tiles);
    if (tiles.empty()) return;
    const size_t T = tiles.size();
    I.sp_tile_bytes = tile_bytes;
    I.sp_codes.ensure(T * tile_bytes);
    I.sp_bnds.ensure(T * (size_t)(IV + 1) * sizeof(WaveBnd));
    I.sp_args.ensure(T * sizeof(BandArgs));
    I.sp_prefix.ensure((T + 2) * sizeof(int));
    std::vector<BandArgs> args(T);
    std::vector<int> prefix(T + 2, 0);
    for (size_t t = 0; t < T; t++) {
        const int k = tiles[t].first, s_lo = tiles[t].second;
        BandArgs& ba = args[t];
        memset(&ba, 0, sizeof(ba));
        const int row0 = k * IV, row1 = std::min(row0 + IV, nn);
        const int s_hi = std::min(std::min(s_lo + G - 1, I.s_last), plan_.s_total - 1);
        ba.R = I.R.as<uint8_t>(); ba.Q = I.Q.as<uint8_t>(); ba.nn = nn; ba.mm = mm;
        ba.s_lo = s_lo; ba.n_strips = s_hi - s_lo + 1; ba.s_total = plan_.s_total; ba.s_band_first = I.s_first; ba.s_band_last = I.s_last;
        ba.group = G; ba.row0 = row0; ba.row1 = row1; ba.ck_col0 = col_first_;
        ba.ckpt_in = k > 0 ? I.ckpt.as<int>() + (long long)(k - 1) * I.ckpt_stride : nullptr;
        ba.ckpt_out = nullptr; ba.interval = IV; ba.ckpt_stride = I.ckpt_stride;
        ba.colck = I.colck.as<WaveBnd>(); ba.colck_g0 = I.g_first; ba.store_cols = 0;
        ba.bnd_local = I.sp_bnds.as<WaveBnd>() + (long long)t * (IV + 1) - row0;      // indexed by the absolute row
        ba.bnd_out = nullptr;
        ba.dir = I.sp_codes.as<uint8_t>() + t * tile_bytes; ba.dstride = I.dstride; ba.row_base = k > 0 ? row0 + 1 : 0;
        ba.ticket = nullptr; ba.result = I.small.as<int>() + 2;
        prefix[t + 1] = prefix[t] + ba.n_strips;
    }
    prefix[T + 1] = 0;      // the ticket
    rt::h2d(I.sp_args.p, args.data(), T * sizeof(BandArgs), I.stream);
    rt::h2d(I.sp_prefix.p, prefix.data(), (T + 2) * sizeof(int), I.stream);
    rt::dev_memset(I.sp_bnds.p, 0xff, T * (size_t)(IV + 1) * sizeof(WaveBnd), I.stream);
    rt::stream_sync(I.stream);   // `args`, `prefix` are locals
    BandBatch bb;
    bb.args = I.sp_args.as<BandArgs>(); bb.pair_of = nullptr; bb.prefix = I.sp_prefix.as<int>(); bb.n_pairs = (int)T; bb.ticket = I.sp_prefix.as<int>() + T + 1; bb.order = nullptr;
    const int blocks = std::min(I.resident_blocks, (prefix[T] + WAVE_WARPS - 1) / WAVE_WARPS);
#ifndef TSA_EMUL
    rt::check(cudaEventRecord(I.ev[0], I.stream), "cudaEventRecord");
#endif
    TSA_LAUNCH(k_band_batch<true>, dim3((unsigned)std::max(1, blocks)), dim3(32 * WAVE_WARPS), (size_t)WAVE_SMEM_INTS * sizeof(int), I.stream, I.cfg.as<DevConfig>(), bb);
#ifndef TSA_EMUL
    rt::check(cudaEventRecord(I.ev[1], I.stream), "cudaEventRecord");
#endif
    I.sp_tiles = tiles;
    I.sp_pending = true;
    stats_.speculated_tiles = (long long)T;
    stats_.resident_bytes += (long long)(I.sp_codes.used + I.sp_bnds.used + I.sp_args.used + I.sp_prefix.used);
    for (size_t t = 0; t < T; t++) stats_.tile_cells += (long long)(args[t].row1 - args[t].row0 + 1) * (long long)args[t].n_strips * WAVE_SW;
}

BandWalk LongPair::walk(const BandWalk& in, std::vector<uint8_t>& ops_rev) {
    Impl& I = *impl_;
    BandWalk st = in;
    if (!traceback_) { st.status = WALK_ERR; return st; }
#ifndef TSA_EMUL
    rt::check(cudaSetDevice(device_), "cudaSetDevice");
#endif
    const auto t0 = std::chrono::steady_clock::now();
    const int nn = plan_.nn, mm = plan_.mm, IV = plan_.interval, G = plan_.group;
    size_t pos = 0;
    while (st.status == WALK_GOING && st.j >= col_first_) {
        const int k = st.i == 0 ? 0 : (st.i - 1) / IV;
        const int row0 = k * IV, row1 = st.i;
        const int s_hi = st.j / WAVE_SW, s_lo = (s_hi / G) * G;
        if (s_hi > I.s_last || s_lo < I.s_first) { st.status = WALK_ERR; break; }
        int spec = -1;
        for (size_t t = 0; t < I.sp_tiles.size(); t++) if (I.sp_tiles[t].first == k && I.sp_tiles[t].second == s_lo) { spec = (int)t; break; }
        if (spec >= 0 && I.sp_pending) {
#ifndef TSA_EMUL
            rt::check(cudaEventSynchronize(I.ev[1]), "speculative tiles");
            float ms = 0; cudaEventElapsedTime(&ms, I.ev[0], I.ev[1]);
            stats_.speculate_ms = ms;
#endif
            rt::stream_sync(I.stream);
            I.sp_pending = false;
        }
        int init[SMALL_INTS] = {0};
        WalkState ws;
        ws.i = st.i; ws.j = st.j; ws.g = st.g; ws.need = st.need; ws.cost = st.cost; ws.status = WALK_GOING; ws.pad = 0;
        rt::h2d(I.small.p, init, sizeof(init), I.stream);
        rt::h2d(I.small.as<int>() + SMALL_INTS, &ws, sizeof(ws), I.stream);
        rt::dev_memset(I.bnd_local.as<WaveBnd>() + row0, 0xff, (size_t)(row1 - row0 + 1) * 8, I.stream);
        BandArgs ba;
        memset(&ba, 0, sizeof(ba));
        ba.R = I.R.as<uint8_t>(); ba.Q = I.Q.as<uint8_t>(); ba.nn = nn; ba.mm = mm;
        ba.s_lo = s_lo; ba.n_strips = s_hi - s_lo + 1; ba.s_total = plan_.s_total; ba.s_band_first = I.s_first; ba.s_band_last = I.s_last;
        ba.group = G; ba.row0 = row0; ba.row1 = row1; ba.ck_col0 = col_first_;
        ba.ckpt_in = k > 0 ? I.ckpt.as<int>() + (long long)(k - 1) * I.ckpt_stride : nullptr;
        ba.ckpt_out = nullptr; ba.interval = IV; ba.ckpt_stride = I.ckpt_stride;
        ba.colck = I.colck.as<WaveBnd>(); ba.colck_g0 = I.g_first; ba.store_cols = 0;
        ba.bnd_local = I.bnd_local.as<WaveBnd>(); ba.bnd_out = nullptr;
        ba.dir = I.tile.as<uint8_t>(); ba.dstride = I.dstride; ba.row_base = k > 0 ? row0 + 1 : 0;
        ba.ticket = I.small.as<int>(); ba.result = I.small.as<int>() + 2;
        const int blocks = std::min(I.resident_blocks, (ba.n_strips + WAVE_WARPS - 1) / WAVE_WARPS);
        if (spec >= 0) {
            ba.dir = I.sp_codes.as<uint8_t>() + (size_t)spec * I.sp_tile_bytes;      // recomputed ahead (the whole tile: a superset of what this walk needs)
            stats_.speculated_used++;
        } else {
            TSA_LAUNCH(k_affine_band<true>, dim3((unsigned)blocks), dim3(32 * WAVE_WARPS), (size_t)WAVE_SMEM_INTS * sizeof(int), I.stream, I.cfg.as<DevConfig>(), ba);
        }
        WalkArgs wa;
        memset(&wa, 0, sizeof(wa));
        wa.R = ba.R; wa.Q = ba.Q; wa.dir = ba.dir; wa.dstride = ba.dstride; wa.row_base = ba.row_base; wa.col_base = s_lo * WAVE_SW;
        wa.row_lo = row0;
        wa.ops = I.ops.as<uint8_t>() + pos; wa.ops_cap = (int)std::min<size_t>(I.ops_cap - pos, (size_t)1 << 30);
        wa.ops_len = I.small.as<int>() + 4;
        wa.state = reinterpret_cast<WalkState*>(I.small.as<int>() + SMALL_INTS);
        TSA_LAUNCH(k_band_walk, dim3(1), dim3(32), 0, I.stream, I.cfg.as<DevConfig>(), wa);
        int h[SMALL_INTS];
        rt::d2h(h, I.small.p, sizeof(h), I.stream);
        rt::d2h(&ws, I.small.as<int>() + SMALL_INTS, sizeof(ws), I.stream);
        rt::stream_sync(I.stream);
        const bool moved = ws.i != st.i || ws.j != st.j || ws.need != st.need || ws.g != st.g || ws.status != WALK_GOING;
        st.i = ws.i; st.j = ws.j; st.g = ws.g; st.need = ws.need; st.cost = ws.cost; st.status = ws.status;
        pos += (size_t)h[4];
        if (spec < 0) {
            stats_.tiles++;
            stats_.tile_cells += (long long)(row1 - row0 + 1) * (long long)ba.n_strips * WAVE_SW;
        }
        if (!moved) { st.status = WALK_ERR; break; }
    }
    if (pos) {
        const size_t at = ops_rev.size();
        ops_rev.resize(at + pos);
        rt::d2h(ops_rev.data() + at, I.ops.p, pos, I.stream);
        rt::stream_sync(I.stream);
    }
    stats_.trace_ms += 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return st;
}

LongResult align_long(const HostConfig& cfg, const int* devices, int n_devices, const uint8_t* R, int nn, const uint8_t* Q, int mm,
                      int interval, int group, size_t memory_limit, bool traceback) {
    LongResult res;
    int world = std::max(1, n_devices);
    // fewer strips than devices: use as many devices as there are strips
    world = std::min(world, std::max(1, wave_strips(mm + 1)));
    res.plan = plan_bands(nn, mm, world, interval, group, memory_limit, traceback);
    if (!res.plan.ok) { res.status = PAIR_OK; res.memory_limit_hit = true; res.message = res.plan.why; return res; }
    const auto dbg_t0 = std::chrono::steady_clock::now();
    std::vector<std::unique_ptr<LongPair>> lp;
    for (int r = 0; r < world; r++) {
        lp.emplace_back(new LongPair(cfg, devices ? devices[r] : 0, R, Q, res.plan, r, traceback));
        if (!lp.back()->ok()) throw std::runtime_error(lp.back()->error());
    }
#ifndef TSA_EMUL
    for (int r = 0; r + 1 < world; r++) {
        const int a = lp[r]->device(), b = lp[r + 1]->device();
        if (a != b) {
            int can = 0;
            rt::check(cudaDeviceCanAccessPeer(&can, a, b), "cudaDeviceCanAccessPeer");
            if (!can) throw std::runtime_error("devices of neighbouring bands have no peer access");
            rt::check(cudaSetDevice(a), "cudaSetDevice");
            const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) rt::check(e, "cudaDeviceEnablePeerAccess");
            cudaGetLastError();
        }
    }
#endif
    const auto dbg_t1 = std::chrono::steady_clock::now();
    for (int r = 0; r + 1 < world; r++) lp[r]->set_outgoing_boundary(lp[r + 1]->incoming_boundary());
    for (int r = 0; r < world; r++) lp[r]->forward_launch();
    for (int r = 0; r < world; r++) lp[r]->forward_wait();
    if (traceback) {
        // code buffers for tiles recomputed ahead of the walk: what the memory limit leaves on every device, at most 16 GiB
        for (int r = 0; r < world; r++) {
            size_t budget = (size_t)16 << 30;
            if (memory_limit) { const long long used = res.plan.resident_bytes(r, true); budget = (long long)memory_limit > used ? std::min<size_t>(budget, (size_t)((long long)memory_limit - used)) : 0; }
            lp[r]->speculate(budget);
        }
    }
    bool sat = false;
    for (int r = 0; r < world; r++) sat = sat || lp[r]->saturated();
    res.cost = lp[world - 1]->cost();
    if (res.cost >= INF32) res.status = PAIR_NO_TARGET;
    else if (sat && res.cost >= WAVE_SAT) res.status = PAIR_ERR_COST_RANGE;
    if (res.status == PAIR_OK && traceback) {
        BandWalk st;
        st.i = nn; st.j = mm; st.g = 0; st.need = 1; st.cost = res.cost; st.status = WALK_GOING;
        std::vector<uint8_t> rev;
        rev.reserve((size_t)nn + mm);
        int r = world - 1;
        for (;;) {
            st = lp[r]->walk(st, rev);
            if (st.status != WALK_GOING) break;
            r = st.j < 0 ? -1 : res.plan.owner_of_column(st.j);
            if (r < 0) { st.status = WALK_ERR; break; }
        }
        if (st.status != WALK_DONE) throw std::runtime_error("long pair: the traceback walk failed (status " + std::to_string(st.status) + ")");
        res.ops.assign(rev.rbegin(), rev.rend());
    }
    for (int r = 0; r < world; r++) res.stats.push_back(lp[r]->stats());
    const auto dbg_t2 = std::chrono::steady_clock::now();
    lp.clear();
    if (getenv("TSA_B200_DEBUG"))
        fprintf(stderr, "[tsalign_b200] align_long: setup %.1f ms, forward + traceback %.1f ms, release %.1f ms\n", 1e3 * std::chrono::duration<double>(dbg_t1 - dbg_t0).count(),
                1e3 * std::chrono::duration<double>(dbg_t2 - dbg_t1).count(), 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - dbg_t2).count());
    return res;
}

}  // namespace tsa
