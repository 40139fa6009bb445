// TEST AID ONLY: a small SIMT emulator so that the warp-cooperative kernels of csrc/relay2.cuh can be compiled with
// g++ and run on the CPU test box (no GPU there).  Every thread of a block is a ucontext fiber; warp collectives and
// block barriers are rendezvous points; TMA bulk copies complete at once.  The product library never includes this
// file -- it has no CPU path.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <ucontext.h>
#include <dlfcn.h>
#include <functional>
#include <vector>

#ifndef __CUDACC__
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
#endif

namespace simt {

enum Op { OP_NONE = 0, OP_BALLOT, OP_SHFL, OP_SHFL_UP, OP_SHFL_DOWN, OP_SHFL_XOR, OP_SYNCWARP, OP_RED_ADD, OP_RED_OR, OP_RED_MAX, OP_RED_MIN, OP_RED_AND, OP_MATCH };

struct WarpState {
    uint32_t arrived = 0, gen = 0, mask = 0; int op = OP_NONE;
    uint64_t in[32]; uint32_t aux[32]; uint64_t out[32];
    void* site = nullptr;
};

struct Fiber {
    ucontext_t ctx; void* stack = nullptr; bool done = false, started = false; unsigned tid = 0;
};

struct BlockRun {
    unsigned n_threads = 0, block = 0, grid = 0;
    std::vector<Fiber> fibers;
    std::vector<WarpState> warps;
    ucontext_t main_ctx;
    unsigned cur = 0;
    unsigned bar_arrived = 0, bar_gen = 0, bar_or = 0, bar_or_out = 0;
    unsigned alive = 0;
    uint64_t progress = 0;
    std::function<void()> body;
};

static const size_t kStack = 256 * 1024;
static BlockRun* g_run = nullptr;
alignas(128) static uint8_t g_dyn_smem[256 * 1024];

static inline unsigned tid() { return g_run->fibers[g_run->cur].tid; }
static inline unsigned bid() { return g_run->block; }
static inline unsigned nthreads() { return g_run->n_threads; }
static inline unsigned nblocks() { return g_run->grid; }

static inline void yield() { BlockRun* r = g_run; swapcontext(&r->fibers[r->cur].ctx, &r->main_ctx); }

static void trampoline() {
    BlockRun* r = g_run;
    r->body();
    r->fibers[r->cur].done = true; --r->alive; ++r->progress;
    swapcontext(&r->fibers[r->cur].ctx, &r->main_ctx);
}

// run `body` once per thread of every block of the grid (blocks one after the other)
static inline void launch(unsigned grid, unsigned block, const std::function<void()>& body) {
    static std::vector<void*> stacks;
    while (stacks.size() < block) {
        void* s = mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (s == MAP_FAILED) { perror("mmap"); abort(); }
        stacks.push_back(s);
    }
    for (unsigned b = 0; b < grid; ++b) {
        BlockRun run; run.n_threads = block; run.block = b; run.grid = grid; run.body = body;
        run.fibers.resize(block); run.warps.assign((block + 31) / 32, WarpState());
        run.alive = block;
        g_run = &run;
        for (unsigned t = 0; t < block; ++t) {
            Fiber& f = run.fibers[t]; f.tid = t; f.stack = stacks[t];
            getcontext(&f.ctx); f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = kStack; f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, trampoline, 0);
        }
        uint64_t last_progress = ~0ull; unsigned idle_rounds = 0;
        while (run.alive) {
            for (unsigned t = 0; t < block; ++t) {
                if (run.fibers[t].done) continue;
                run.cur = t;
                swapcontext(&run.main_ctx, &run.fibers[t].ctx);
            }
            if (run.progress == last_progress) { if (++idle_rounds > 4) { fprintf(stderr, "simt_emu: deadlock (block %u): a collective or barrier is waiting for threads that never arrive\n", b); abort(); } }
            else { idle_rounds = 0; last_progress = run.progress; }
        }
        g_run = nullptr;
    }
}

static inline uint64_t collective(int op, uint32_t mask, uint64_t val, uint32_t aux) {
    BlockRun* r = g_run; const unsigned t = r->fibers[r->cur].tid, lane = t & 31;
    WarpState& w = r->warps[t >> 5];
    if (!(mask & (1u << lane))) { fprintf(stderr, "simt_emu: lane %u not in mask %08x\n", lane, mask); abort(); }
    void* site = __builtin_return_address(0);
#if defined(SIMT_EMU_SITES)
    site = __builtin_return_address(1);
#endif
    if (w.arrived == 0) { w.mask = mask; w.op = op; w.site = site; }
    else if (w.mask != mask || w.op != op) {
        Dl_info di0{}, di1{};                      // offsets inside the test library: `addr2line -e tests/support/_host_relay2.so <offset>`
        dladdr(w.site, &di0); dladdr(site, &di1);
        fprintf(stderr, "simt_emu: divergent collective in warp %u lane %u (op %d/%d mask %08x/%08x) sites +0x%zx / +0x%zx (arrived %08x)\n", t >> 5, lane, w.op, op, w.mask, mask,
                (size_t)((char*)w.site - (char*)di0.dli_fbase), (size_t)((char*)site - (char*)di1.dli_fbase), w.arrived);
        abort();
    }
    w.in[lane] = val; w.aux[lane] = aux; w.arrived |= 1u << lane;
    const uint32_t gen = w.gen;
    ++r->progress;
    if (w.arrived == mask) {
        uint64_t acc = 0; bool first = true;
        for (unsigned l = 0; l < 32; ++l) if (mask & (1u << l)) {
            switch (op) {
            case OP_BALLOT: if (w.in[l]) acc |= 1ull << l; break;
            case OP_RED_ADD: acc = (uint32_t)(acc + w.in[l]); break;
            case OP_RED_OR: acc |= w.in[l]; break;
            case OP_RED_AND: acc = first ? w.in[l] : (acc & w.in[l]); break;
            case OP_RED_MAX: acc = first ? w.in[l] : (w.in[l] > acc ? w.in[l] : acc); break;
            case OP_RED_MIN: acc = first ? w.in[l] : (w.in[l] < acc ? w.in[l] : acc); break;
            default: break;
            }
            first = false;
        }
        for (unsigned l = 0; l < 32; ++l) if (mask & (1u << l)) {
            int src = -1;
            switch (op) {
            case OP_SHFL: src = (int)(w.aux[l] & 31); break;
            case OP_SHFL_UP: src = (int)l - (int)w.aux[l]; if (src < 0) src = (int)l; break;
            case OP_SHFL_DOWN: src = (int)l + (int)w.aux[l]; if (src > 31) src = (int)l; break;
            case OP_SHFL_XOR: src = (int)(l ^ w.aux[l]) & 31; break;
            case OP_MATCH: { uint64_t m = 0; for (unsigned k = 0; k < 32; ++k) if ((mask & (1u << k)) && w.in[k] == w.in[l]) m |= 1ull << k; w.out[l] = m; continue; }
            default: w.out[l] = acc; continue;
            }
            w.out[l] = (mask & (1u << src)) ? w.in[src] : w.in[l];
        }
        w.arrived = 0; ++w.gen;
    } else {
        while (w.gen == gen) yield();
    }
    return w.out[lane];
}

static inline unsigned block_barrier(unsigned pred) {
    BlockRun* r = g_run;
    const unsigned gen = r->bar_gen;
    r->bar_or |= pred ? 1u : 0u;
    ++r->bar_arrived; ++r->progress;
    unsigned done = 0; for (auto& f : r->fibers) done += f.done ? 1u : 0u;
    if (r->bar_arrived + done >= r->n_threads) { r->bar_or_out = r->bar_or; r->bar_or = 0; r->bar_arrived = 0; ++r->bar_gen; }
    else while (r->bar_gen == gen) yield();
    return r->bar_or_out;
}

}  // namespace simt

#ifndef __CUDACC__
static inline uint32_t __ballot_sync(uint32_t m, int p) { return (uint32_t)simt::collective(simt::OP_BALLOT, m, p ? 1 : 0, 0); }
static inline int __any_sync(uint32_t m, int p) { return (simt::collective(simt::OP_BALLOT, m, p ? 1 : 0, 0) != 0); }
static inline int __all_sync(uint32_t m, int p) { return ((uint32_t)simt::collective(simt::OP_BALLOT, m, p ? 1 : 0, 0) == m); }
static inline void __syncwarp(uint32_t m = 0xffffffffu) { simt::collective(simt::OP_SYNCWARP, m, 0, 0); }
template <class T> static inline T __shfl_sync(uint32_t m, T v, int src) { uint64_t x = 0; memcpy(&x, &v, sizeof(T)); x = simt::collective(simt::OP_SHFL, m, x, (uint32_t)src); T o; memcpy(&o, &x, sizeof(T)); return o; }
template <class T> static inline T __shfl_up_sync(uint32_t m, T v, unsigned d) { uint64_t x = 0; memcpy(&x, &v, sizeof(T)); x = simt::collective(simt::OP_SHFL_UP, m, x, d); T o; memcpy(&o, &x, sizeof(T)); return o; }
template <class T> static inline T __shfl_down_sync(uint32_t m, T v, unsigned d) { uint64_t x = 0; memcpy(&x, &v, sizeof(T)); x = simt::collective(simt::OP_SHFL_DOWN, m, x, d); T o; memcpy(&o, &x, sizeof(T)); return o; }
template <class T> static inline T __shfl_xor_sync(uint32_t m, T v, int d) { uint64_t x = 0; memcpy(&x, &v, sizeof(T)); x = simt::collective(simt::OP_SHFL_XOR, m, x, (uint32_t)d); T o; memcpy(&o, &x, sizeof(T)); return o; }
static inline uint32_t __reduce_add_sync(uint32_t m, uint32_t v) { return (uint32_t)simt::collective(simt::OP_RED_ADD, m, v, 0); }
static inline uint32_t __reduce_or_sync(uint32_t m, uint32_t v) { return (uint32_t)simt::collective(simt::OP_RED_OR, m, v, 0); }
static inline uint32_t __reduce_and_sync(uint32_t m, uint32_t v) { return (uint32_t)simt::collective(simt::OP_RED_AND, m, v, 0); }
static inline uint32_t __reduce_max_sync(uint32_t m, uint32_t v) { return (uint32_t)simt::collective(simt::OP_RED_MAX, m, v, 0); }
static inline uint32_t __reduce_min_sync(uint32_t m, uint32_t v) { return (uint32_t)simt::collective(simt::OP_RED_MIN, m, v, 0); }
static inline uint32_t __match_any_sync(uint32_t m, uint64_t v) { return (uint32_t)simt::collective(simt::OP_MATCH, m, v, 0); }
static inline void __syncthreads() { simt::block_barrier(0); }
static inline int __syncthreads_or(int p) { return (int)simt::block_barrier(p ? 1u : 0u); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) { s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo; }
static inline int __ffs(uint32_t x) { return x ? __builtin_ctz(x) + 1 : 0; }
static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
static inline uint32_t atomicOr(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o | v; return o; }
static inline uint32_t atomicAnd(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o & v; return o; }
static inline uint32_t atomicMax(uint32_t* p, uint32_t v) { uint32_t o = *p; if (v > o) *p = v; return o; }
static inline uint32_t atomicMin(uint32_t* p, uint32_t v) { uint32_t o = *p; if (v < o) *p = v; return o; }
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; if (v > o) *p = v; return o; }
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; if (v < o) *p = v; return o; }
static inline uint32_t atomicCAS(uint32_t* p, uint32_t cmp, uint32_t v) { uint32_t o = *p; if (o == cmp) *p = v; return o; }
static inline uint32_t atomicExch(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = v; return o; }
#endif
