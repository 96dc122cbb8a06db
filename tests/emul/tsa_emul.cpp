// tsa_emul.cpp -- TEST INFRASTRUCTURE ONLY: lock-step SIMT emulator behind csrc/tsa_rt.hpp (-DTSA_EMUL).
//
// Every CUDA thread of a block is a fiber with its own stack; warp primitives (shuffles, ballots, reductions,
// __syncwarp) and __syncthreads are generation barriers, so the kernels execute with the same data exchange
// pattern as on the device.  A barrier that can never complete (divergent warp) aborts with a message.
// x86-64 SysV only (hand-written context switch).
#ifndef TSA_EMUL
#define TSA_EMUL
#endif
#include "tsa_rt.hpp"

#include <vector>

extern "C" void tsa_emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl tsa_emu_switch
.type tsa_emu_switch,@function
tsa_emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size tsa_emu_switch,.-tsa_emu_switch
)");

namespace tsa { namespace emu {

struct Fiber {
    void* sp = nullptr;
    unsigned char* stack = nullptr;
    bool done = false;
    uint3e tid{0, 0, 0};
    int warp = 0;
};
struct Barrier { int count = 0, gen = 0, expected = 0; };

static const size_t STACK_BYTES = 256 * 1024;
static std::vector<Fiber> g_fibers;
static std::vector<unsigned char*> g_stacks;
static std::vector<Barrier> g_warp_bar;
static std::vector<uint32_t> g_slots;
static Barrier g_block_bar;
static Fiber* g_cur = nullptr;
static void* g_sched_sp = nullptr;
static std::vector<unsigned char> g_smem;
static uint3e g_bid, g_bdim, g_gdim;
static void (*g_entry)(void*) = nullptr;
static void* g_args = nullptr;
static long g_progress = 0;

Fiber* cur() { return g_cur; }
unsigned char* smem() { return g_smem.data(); }
uint32_t* warp_slots() { return &g_slots[(size_t)g_cur->warp * 32]; }
uint3e tid() { return g_cur->tid; }
uint3e bid() { return g_bid; }
uint3e bdim() { return g_bdim; }
uint3e gdim() { return g_gdim; }

static void yield() { tsa_emu_switch(&g_cur->sp, g_sched_sp); }

static void wait_on(Barrier& b) {
    int gen = b.gen;
    g_progress++;
    if (++b.count == b.expected) { b.count = 0; b.gen++; }
    else while (b.gen == gen) yield();
}
void warp_barrier() { wait_on(g_warp_bar[g_cur->warp]); }
void spin_yield() { yield(); }   // a polling loop: no progress of its own, the deadlock check still sees the others
void block_barrier() { wait_on(g_block_bar); }

static void fiber_main() {
    g_entry(g_args);
    g_cur->done = true;
    g_progress++;
    for (;;) yield();
}

void launch(void (*entry)(void*), void* args, Dim grid, Dim block, size_t smem_bytes) {
    const unsigned nthreads = block.x * block.y * block.z;
    const unsigned nwarps = (nthreads + 31) / 32;
    while (g_stacks.size() < nthreads) g_stacks.push_back((unsigned char*)aligned_alloc(64, STACK_BYTES));
    g_entry = entry; g_args = args;
    g_bdim = uint3e{block.x, block.y, block.z};
    g_gdim = uint3e{grid.x, grid.y, grid.z};
    g_smem.assign(smem_bytes + 64, 0);
    for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) {
        g_bid = uint3e{bx, by, bz};
        g_fibers.assign(nthreads, Fiber());
        g_warp_bar.assign(nwarps, Barrier());
        g_slots.assign((size_t)nwarps * 32, 0);
        g_block_bar = Barrier(); g_block_bar.expected = (int)nthreads;
        for (unsigned t = 0; t < nthreads; t++) {
            Fiber& f = g_fibers[t];
            f.tid = uint3e{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
            f.warp = (int)(t / 32);
            g_warp_bar[f.warp].expected++;
            f.stack = g_stacks[t];
            uintptr_t top = ((uintptr_t)f.stack + STACK_BYTES) & ~(uintptr_t)15;
            void** sp = (void**)top;
            *--sp = nullptr;                 // fake return address of fiber_main
            *--sp = (void*)&fiber_main;      // popped by `ret` in tsa_emu_switch
            for (int r = 0; r < 6; r++) *--sp = nullptr;
            f.sp = sp;
        }
        unsigned alive = nthreads;
        while (alive > 0) {
            long before = g_progress;
            alive = 0;
            for (unsigned t = 0; t < nthreads; t++) {
                Fiber& f = g_fibers[t];
                if (f.done) continue;
                g_cur = &f;
                tsa_emu_switch(&g_sched_sp, f.sp);
                if (!f.done) alive++;
            }
            if (alive > 0 && g_progress == before) {
                fprintf(stderr, "tsa_emul: deadlock in block (%u,%u,%u): a barrier is waiting for threads that exited or diverged\n", bx, by, bz);
                abort();
            }
        }
        g_cur = nullptr;
    }
}

} }  // namespace tsa::emu
