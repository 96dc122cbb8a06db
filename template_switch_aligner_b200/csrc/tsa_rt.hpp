// tsa_rt.hpp -- thin runtime layer under the kernels and the engine.
//
// Product build (nvcc, sm_100a): everything maps 1:1 onto CUDA (cudaMalloc, <<<>>>, warp intrinsics, DPX).
// Test build (-DTSA_EMUL, plain g++): the same kernel source runs on a lock-step SIMT emulator (one fiber
// per CUDA thread, barriers at every warp primitive).  The emulator exists only so that the CPU test-suite
// (`pytest -m "not gpu"`) can exercise the kernel logic; it is built into tests/emul/_build, never into the
// shipped library, and the product has no CPU path (tsa_capi.cpp fails with TSA_ERR_NO_DEVICE).
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <algorithm>
#include <cstring>

#ifndef TSA_EMUL
// ------------------------------------------------------------------------------------------------ CUDA
#include <cuda_runtime.h>
#define TSA_DEV __device__ __forceinline__
#define TSA_KERNEL __global__
#define TSA_LAUNCH_BOUNDS(threads, blocks) __launch_bounds__(threads, blocks)
#define TSA_SHARED_DECL(name) extern __shared__ __align__(16) unsigned char name[]
#define TSA_HOSTDEV __host__ __device__

#include <stdexcept>
#include <string>
namespace tsa { namespace rt {
// CUDA failures (out of memory, launch errors) surface as exceptions; the C ABI turns them into status codes.
struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
inline void check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) {
        cudaGetLastError();  // clear the sticky-free error state for the next call
        throw CudaError(std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e));
    }
}
} }
#define TSA_LAUNCH(kernel, grid, block, smem, stream, ...)                                   \
    do {                                                                                      \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                           \
        ::tsa::rt::check(cudaGetLastError(), #kernel);                                        \
    } while (0)

namespace tsa {
TSA_DEV int lane_id() { return (int)(threadIdx.x & 31); }
TSA_DEV uint32_t shfl_up(uint32_t v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
TSA_DEV uint32_t shfl_down(uint32_t v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
TSA_DEV uint32_t shfl_xor(uint32_t v, int d) { return __shfl_xor_sync(0xffffffffu, v, d); }
TSA_DEV uint32_t shfl_idx(uint32_t v, int src) { return __shfl_sync(0xffffffffu, v, src); }
TSA_DEV int reduce_min_s32(int v) { return __reduce_min_sync(0xffffffffu, v); }
TSA_DEV int reduce_max_s32(int v) { return __reduce_max_sync(0xffffffffu, v); }
TSA_DEV int reduce_add_s32(int v) { return __reduce_add_sync(0xffffffffu, v); }
TSA_DEV uint32_t ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
TSA_DEV void sync_warp() { __syncwarp(); }
TSA_DEV void sync_block() { __syncthreads(); }
// DPX (sm_90+): fused add+min / three-way min, scalar s32 and packed s16x2
TSA_DEV int addmin_s32(int a, int b, int c) { return __viaddmin_s32(a, b, c); }               // min(a+b, c)
TSA_DEV int min3_s32(int a, int b, int c) { return __vimin3_s32(a, b, c); }
TSA_DEV uint32_t addmin_s16x2(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_s16x2(a, b, c); }
TSA_DEV uint32_t min3_s16x2(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_s16x2(a, b, c); }
TSA_DEV uint32_t min_s16x2(uint32_t a, uint32_t b) { return __vmins2(a, b); }
TSA_DEV uint32_t max_s16x2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
// per half: 0xffff where a + negb < 0 (a - b without overflow for the non-negative costs used here): VIADDMNMX + PRMT
TSA_DEV uint32_t ltmask_s16x2(uint32_t a, uint32_t negb) {
    uint32_t r;   // (__byte_perm ignores the sign-replication bit of the selector nibbles)
    asm("prmt.b32 %0, %1, 0, 0xbb99;" : "=r"(r) : "r"(__viaddmin_s16x2(a, negb, 0x7fff7fffu)));
    return r;
}
TSA_DEV int clamp0_s32(int v, int hi) { return __vimin_s32_relu(v, hi); }                            // max(min(v, hi), 0)
TSA_DEV uint32_t add_s16x2(uint32_t a, uint32_t b) { return __vadd2(a, b); }
TSA_DEV uint32_t cmplt_s16x2(uint32_t a, uint32_t b) { return __vcmplts2(a, b); }                 // per half: 0xffff where a < b (signed)
TSA_DEV int atomic_min_s32(int* p, int v) { return atomicMin(p, v); }
TSA_DEV int atomic_or_s32(int* p, int v) { return atomicOr(p, v); }
TSA_DEV int atomic_max_s32(int* p, int v) { return atomicMax(p, v); }
TSA_DEV int atomic_and_s32(int* p, int v) { return atomicAnd(p, v); }
TSA_DEV int atomic_add_s32(int* p, int v) { return atomicAdd(p, v); }
TSA_DEV int clz_u32(uint32_t v) { return __clz((int)v); }
TSA_DEV int popc_u32(uint32_t v) { return __popc(v); }
TSA_DEV int ffs_u32(uint32_t v) { return __ffs((int)v); }                                          // 1-based index of the lowest set bit, 0 if none
// producer / consumer flags between warps of one launch (k_affine_wave): release store, acquire load, L2 data load
TSA_DEV int ld_acquire_s32(const int* p) { int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
TSA_DEV void st_release_s32(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
TSA_DEV int ld_cg_s32(const int* p) { return __ldcg(p); }
TSA_DEV void thread_fence() { __threadfence(); }
#ifndef TSA_SPIN_NS
#define TSA_SPIN_NS 40
#endif
TSA_DEV void spin_pause() { __nanosleep(TSA_SPIN_NS); }
}  // namespace tsa

#else
// ------------------------------------------------------------------------------------------------ emulator
#include <vector>
#include <algorithm>
#define TSA_DEV inline
#define TSA_KERNEL
#define TSA_LAUNCH_BOUNDS(threads, blocks)
#define TSA_HOSTDEV
#define TSA_SHARED_DECL(name) unsigned char* name = ::tsa::emu::smem()

namespace tsa { namespace emu {
struct uint3e { unsigned x, y, z; };
struct Fiber;
Fiber* cur();
unsigned char* smem();
void warp_barrier();
void block_barrier();
void spin_yield();
uint32_t* warp_slots();
uint3e tid(); uint3e bid(); uint3e bdim(); uint3e gdim();
struct Dim { unsigned x = 1, y = 1, z = 1; Dim(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
void launch(void (*entry)(void*), void* args, Dim grid, Dim block, size_t smem_bytes);
} }
struct tsa_emu_idx { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };   // CUDA's 16-byte vector type
#define threadIdx (::tsa::emu::tid())
#define blockIdx (::tsa::emu::bid())
#define blockDim (::tsa::emu::bdim())
#define gridDim (::tsa::emu::gdim())
typedef ::tsa::emu::Dim dim3;
typedef int cudaStream_t;

// kernel<<<grid, block, smem>>>(args...) -> pack the arguments in a lambda, run every CUDA thread as a fiber
#define TSA_LAUNCH(kernel, grid, block, smem, stream, ...)                                   \
    do {                                                                                      \
        auto tsa_thunk = [&]() { kernel(__VA_ARGS__); };                                      \
        ::tsa::emu::launch([](void* p) { (*static_cast<decltype(tsa_thunk)*>(p))(); }, &tsa_thunk, (grid), (block), (smem)); \
        (void)(stream);                                                                       \
    } while (0)

namespace tsa {
inline int lane_id() { return (int)(threadIdx.x & 31); }
inline uint32_t emu_exchange(uint32_t v, int src_lane) {
    uint32_t* s = emu::warp_slots();
    s[lane_id()] = v;
    emu::warp_barrier();
    uint32_t r = s[src_lane];
    emu::warp_barrier();
    return r;
}
inline uint32_t shfl_up(uint32_t v, int d) { int l = lane_id(); return emu_exchange(v, l >= d ? l - d : l); }
inline uint32_t shfl_down(uint32_t v, int d) { int l = lane_id(); return emu_exchange(v, l + d < 32 ? l + d : l); }
inline uint32_t shfl_xor(uint32_t v, int d) { return emu_exchange(v, lane_id() ^ d); }
inline uint32_t shfl_idx(uint32_t v, int src) { return emu_exchange(v, src & 31); }
inline int reduce_min_s32(int v) {
    uint32_t* s = emu::warp_slots();
    s[lane_id()] = (uint32_t)v;
    emu::warp_barrier();
    int r = (int)s[0];
    for (int i = 1; i < 32; i++) r = std::min(r, (int)s[i]);
    emu::warp_barrier();
    return r;
}
inline int reduce_max_s32(int v) { return -reduce_min_s32(-v); }
inline int reduce_add_s32(int v) {
    uint32_t* s = emu::warp_slots();
    s[lane_id()] = (uint32_t)v;
    emu::warp_barrier();
    int r = 0;
    for (int i = 0; i < 32; i++) r += (int)s[i];
    emu::warp_barrier();
    return r;
}
inline uint32_t ballot(bool p) {
    uint32_t* s = emu::warp_slots();
    s[lane_id()] = p ? 1u : 0u;
    emu::warp_barrier();
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= s[i] << i;
    emu::warp_barrier();
    return r;
}
inline void sync_warp() { emu::warp_barrier(); }
inline void sync_block() { emu::block_barrier(); }
inline int addmin_s32(int a, int b, int c) { return std::min(a + b, c); }
inline int min3_s32(int a, int b, int c) { return std::min(a, std::min(b, c)); }
inline int16_t lo16(uint32_t v) { return (int16_t)(v & 0xffff); }
inline int16_t hi16(uint32_t v) { return (int16_t)(v >> 16); }
inline uint32_t pack16(int lo, int hi) { return ((uint32_t)(uint16_t)(int16_t)lo) | ((uint32_t)(uint16_t)(int16_t)hi << 16); }
inline uint32_t addmin_s16x2(uint32_t a, uint32_t b, uint32_t c) {
    // wrap-around add like the hardware instruction, then signed min
    return pack16(std::min<int>((int16_t)(lo16(a) + lo16(b)), lo16(c)), std::min<int>((int16_t)(hi16(a) + hi16(b)), hi16(c)));
}
inline uint32_t min3_s16x2(uint32_t a, uint32_t b, uint32_t c) {
    return pack16(std::min<int>(lo16(a), std::min<int>(lo16(b), lo16(c))), std::min<int>(hi16(a), std::min<int>(hi16(b), hi16(c))));
}
inline uint32_t min_s16x2(uint32_t a, uint32_t b) { return pack16(std::min<int>(lo16(a), lo16(b)), std::min<int>(hi16(a), hi16(b))); }
inline uint32_t max_s16x2(uint32_t a, uint32_t b) { return pack16(std::max<int>(lo16(a), lo16(b)), std::max<int>(hi16(a), hi16(b))); }
inline uint32_t ltmask_s16x2(uint32_t a, uint32_t negb) { return ((int16_t)(lo16(a) + lo16(negb)) < 0 ? 0xffffu : 0u) | ((int16_t)(hi16(a) + hi16(negb)) < 0 ? 0xffff0000u : 0u); }
inline int clamp0_s32(int v, int hi) { return std::max(std::min(v, hi), 0); }
inline uint32_t add_s16x2(uint32_t a, uint32_t b) { return pack16((int16_t)(lo16(a) + lo16(b)), (int16_t)(hi16(a) + hi16(b))); }
inline uint32_t cmplt_s16x2(uint32_t a, uint32_t b) { return (lo16(a) < lo16(b) ? 0xffffu : 0u) | (hi16(a) < hi16(b) ? 0xffff0000u : 0u); }
inline int atomic_min_s32(int* p, int v) { int o = *p; if (v < o) *p = v; return o; }
inline int atomic_or_s32(int* p, int v) { int o = *p; *p = o | v; return o; }
inline int atomic_max_s32(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
inline int atomic_and_s32(int* p, int v) { int o = *p; *p = o & v; return o; }
inline int atomic_add_s32(int* p, int v) { int o = *p; *p = o + v; return o; }
inline int clz_u32(uint32_t v) { return v ? __builtin_clz(v) : 32; }
inline int ffs_u32(uint32_t v) { return __builtin_ffs((int)v); }
inline int popc_u32(uint32_t v) { return __builtin_popcount(v); }
inline int ld_acquire_s32(const int* p) { return *(const volatile int*)p; }
inline void st_release_s32(int* p, int v) { *(volatile int*)p = v; }
inline int ld_cg_s32(const int* p) { return *(const volatile int*)p; }
inline void thread_fence() {}
inline void spin_pause() { emu::spin_yield(); }   // lets the other warps of the block run
}  // namespace tsa
#endif

namespace tsa {
TSA_DEV int imin(int a, int b) { return a < b ? a : b; }
TSA_DEV int imax(int a, int b) { return a > b ? a : b; }
}  // namespace tsa

// ---------------------------------------------------------------------------- device memory (both builds)
namespace tsa { namespace rt {
#ifndef TSA_EMUL
inline void* dev_alloc(size_t bytes) { void* p = nullptr; check(cudaMalloc(&p, bytes ? bytes : 1), "cudaMalloc"); return p; }
inline void dev_free(void* p) { if (p) cudaFree(p); }
inline void h2d(void* d, const void* h, size_t bytes, cudaStream_t s) { check(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s), "h2d"); }
inline void d2h(void* h, const void* d, size_t bytes, cudaStream_t s) { check(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s), "d2h"); }
inline void dev_memset(void* d, int byte, size_t bytes, cudaStream_t s) { check(cudaMemsetAsync(d, byte, bytes, s), "memset"); }
// Waiting host threads: the CUDA default spins on a core.  When several ranks share the host (one process per GPU, two engine
// threads each) the spinning threads take the cores the result assembly needs, so the waits then block on an event created with
// cudaEventBlockingSync instead (TSA_B200_BLOCKING_SYNC=0/1 overrides; automatic: ranks of the launch x 4 >= host cores; measured on one GPU: no difference).
inline bool blocking_sync() {
    static const bool on = []() {
        if (const char* e = getenv("TSA_B200_BLOCKING_SYNC")) return atoi(e) != 0;
        long ranks = 1;
        if (const char* w = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1L, atol(w));
        return ranks * 4 >= (long)std::max(1u, std::thread::hardware_concurrency());
    }();
    return on;
}
inline unsigned event_flags() { return blocking_sync() ? (unsigned)cudaEventBlockingSync : (unsigned)cudaEventDefault; }
inline void stream_sync(cudaStream_t s) {
    if (!blocking_sync()) { check(cudaStreamSynchronize(s), "stream sync"); return; }
    thread_local cudaEvent_t ev[64] = {};
    int dev = 0;
    check(cudaGetDevice(&dev), "cudaGetDevice");
    cudaEvent_t& e = ev[dev & 63];
    if (!e) check(cudaEventCreateWithFlags(&e, cudaEventBlockingSync | cudaEventDisableTiming), "cudaEventCreateWithFlags");
    check(cudaEventRecord(e, s), "cudaEventRecord");
    check(cudaEventSynchronize(e), "event sync");
}
inline void* host_alloc(size_t bytes) { void* p = nullptr; check(cudaMallocHost(&p, bytes ? bytes : 1), "cudaMallocHost"); return p; }
inline void host_free(void* p) { if (p) cudaFreeHost(p); }
#else
inline void* dev_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
inline void dev_free(void* p) { free(p); }
inline void h2d(void* d, const void* h, size_t bytes, cudaStream_t) { memcpy(d, h, bytes); }
inline void d2h(void* h, const void* d, size_t bytes, cudaStream_t) { memcpy(h, d, bytes); }
inline void dev_memset(void* d, int byte, size_t bytes, cudaStream_t) { memset(d, byte, bytes); }
inline void stream_sync(cudaStream_t) {}
inline void* host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
inline void host_free(void* p) { free(p); }
#endif
} }
