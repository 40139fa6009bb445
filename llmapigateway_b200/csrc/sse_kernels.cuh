// Kernels of the SSE step (sm_100a).  See DESIGN.md for the data layout and the roofline of each.
//
//   k_prime    tile table (first chunk of every 8 KiB byte tile, one coalesced pass over the chunk offsets)
//              and, one thread per segment, the plan the bulk kernel works to; a fresh stream is
//              speculated to commit on its first non-empty chunk.
//   k_relay    bulk kernel: persistent blocks copy byte tiles in -> out with 16-byte vector accesses
//              (the re-emit), staging them in shared memory, and walk the events of the chunks that
//              start in the tile (event templates, window matcher, lean recogniser).
//   k_commit   one thread per segment: folds the bulk kernel's findings into the persistent
//              stream state; streams the bulk kernel flagged irregular (or whose speculation failed)
//              are redone sequentially with the exact machine.
//   k_general  (mode 1 / fix-up) the exact sequential machine over whole segments.
#pragma once
#include <cuda_runtime.h>
#include "stream_machine.cuh"
#include "lean_json.cuh"

namespace lgw {

struct DeviceTables {
    StreamState* state;
    uint8_t* carry_a; uint8_t* carry_b; uint8_t* detail; uint8_t* pending;
    uint32_t carry_cap, detail_cap, max_streams;
};

struct SegPlan {
    uint32_t resume_chunk;     // first chunk the sequential path would have to (re)do
    uint32_t relay_begin;      // byte offset where the bulk region starts (tap view); seg end when none
    uint32_t seg_end;          // byte offset of the end of the segment
    uint32_t emit_chunk_begin;
    uint32_t irregular;        // 1: the bulk kernel's findings are void, k_commit redoes [resume_chunk, end)
    uint32_t a_usage;          // handler bound `tokens_usage` (request_handler.py:134)
    unsigned long long last_usage;   // (1 + byte offset) << 32 | length of the last usage-bearing event (atomicMax)
    uint32_t n_events_a, n_events_b, n_usage_b;
    // priming by speculation: a fresh stream is assumed to commit on its first non-empty chunk
    // (request_handler.py:69-95); that chunk's own thread verifies it in the bulk kernel
    uint32_t kept_chunk;       // 0xFFFFFFFF: no speculation (stream was already committed)
    uint32_t kept_end;         // byte offset of the end of the kept chunk = where the handler's text starts
    uint32_t prime_ok;         // set by the kept chunk's thread when the speculation holds
    uint32_t _pad[2];
};
static_assert(sizeof(SegPlan) == 64, "SegPlan");

// Event templates that survive across launches (any validated event is a sound template wherever it
// came from, so sharing them between blocks, steps and slices is safe).  state: 0 empty, 1 being
// written, 2 ready.
#define LGW_TPLC_TEXT 520u
#define LGW_TPLC_MAP 512u
struct TemplateCache {
    uint32_t state[2], len[2], flags[2], cls[2];
    uint16_t sstart[2][32], send[2][32];
    uint8_t skind[2][32];
    uint8_t text[2][LGW_TPLC_TEXT];
    uint8_t map[2][LGW_TPLC_MAP];
};

struct StepScratch {
    SegPlan* plan;             // [max_streams]
    uint32_t* tile_chunk;      // [max tiles + 2] first chunk starting at or after each tile
    TemplateCache* tpl_cache;  // persistent across steps
    uint32_t max_tiles;
    uint32_t* long_q;          // [long_cap][2] (chunk, segment) pairs left out of the tile walk (k_relay -> k_relay_long)
    uint32_t* long_count;      // entries wanted (may exceed long_cap: the excess was walked in place)
    uint32_t long_cap;
};

struct StepArgs {
    DeviceTables t;
    const uint8_t* data; uint32_t n_bytes;     // n_bytes = end of the valid bytes (a slice of a pipelined step ends earlier)
    const uint32_t* chunk_off; uint32_t n_chunks;
    uint32_t tile_base;                        // byte offset of tile 0 (multiple of LGW_TILE_BYTES)
    uint32_t chunk_lo, chunk_hi;               // chunks of this launch: [chunk_lo, chunk_hi)
    const uint32_t* seg_chunk; const uint32_t* seg_slot; uint32_t n_segs;
    uint8_t* out; SegResult* seg_out;
    RowEvent* rowq; uint32_t* rowq_count; uint32_t rowq_cap;
    StepScratch s;
};

#ifndef LGW_TILE_BYTES
#define LGW_TILE_BYTES 8192u
#endif

static inline cudaError_t scratch_alloc(StepScratch& s, size_t max_streams, size_t max_chunks, size_t max_bytes) {
    cudaError_t r;
    if ((r = cudaMalloc((void**)&s.plan, max_streams * sizeof(SegPlan))) != cudaSuccess) return r;
    s.max_tiles = (uint32_t)((max_bytes + LGW_TILE_BYTES - 1) / LGW_TILE_BYTES);
    if ((r = cudaMalloc((void**)&s.tile_chunk, ((size_t)s.max_tiles + 2) * 4)) != cudaSuccess) return r;
    if ((r = cudaMalloc((void**)&s.tpl_cache, sizeof(TemplateCache))) != cudaSuccess) return r;
    if ((r = cudaMemset(s.tpl_cache, 0, sizeof(TemplateCache))) != cudaSuccess) return r;
    s.long_cap = (uint32_t)(max_chunks < (1u << 20) ? (max_chunks ? max_chunks : 1) : (1u << 20));
    if ((r = cudaMalloc((void**)&s.long_q, ((size_t)s.long_cap * 2 + 1) * 4)) != cudaSuccess) return r;
    s.long_count = s.long_q + (size_t)s.long_cap * 2;
    if ((r = cudaMemset(s.long_count, 0, 4)) != cudaSuccess) return r;
    return cudaSuccess;
}
static inline void scratch_free(StepScratch& s) { cudaFree(s.plan); cudaFree(s.tile_chunk); cudaFree(s.tpl_cache); cudaFree(s.long_q); s.long_q = nullptr; s.long_count = nullptr; s.plan = nullptr; s.tile_chunk = nullptr; s.tpl_cache = nullptr; }

__device__ __forceinline__ StepIO make_io(const StepArgs& a, uint32_t slot, StreamHdr* local_hdr) {
    StepIO io;
    io.st = local_hdr;
    io.rec = &a.t.state[slot].rec;
    io.pending = a.t.pending + (size_t)slot * LGW_PENDING_STRIDE;
    io.carry_a = a.t.carry_a + (size_t)slot * a.t.carry_cap;
    io.carry_b = a.t.carry_b + (size_t)slot * a.t.carry_cap;
    io.detail = a.t.detail + (size_t)slot * a.t.detail_cap;
    io.carry_cap = a.t.carry_cap; io.detail_cap = a.t.detail_cap;
    io.rowq = a.rowq; io.rowq_count = a.rowq_count; io.rowq_cap = a.rowq_cap; io.slot = slot;
    return io;
}

// ---- stream table maintenance -------------------------------------------------------------------
__global__ void k_streams_open(DeviceTables t, const uint32_t* slots, const int32_t* status, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    init_stream(t.state[slots[i]], status[i]);
}

__global__ void k_streams_gather(DeviceTables t, const uint32_t* slots, uint32_t n, StreamState* out, int free_after) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t slot = slots[i];
    StreamState* s = t.state + slot;
    if (s->h.flags & SF_PENDING) {            // extract the stashed usage event now (deferred by the bulk path)
        StreamHdr st = s->h;
        StepIO io;
        io.st = &st; io.rec = &s->rec; io.pending = t.pending + (size_t)slot * LGW_PENDING_STRIDE;
        io.carry_a = io.carry_b = io.detail = nullptr; io.carry_cap = io.detail_cap = 0;
        io.rowq = nullptr; io.rowq_count = nullptr; io.rowq_cap = 0; io.slot = slot;
        resolve_pending(io);
        s->h = st;
    }
    out[i] = *s;
    if (free_after) s->h.phase = PH_FREE;
}

// ---- exact sequential path ------------------------------------------------------------------------
__global__ void __launch_bounds__(64) k_general(StepArgs a) {
    const uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= a.n_segs) return;
    const uint32_t slot = a.seg_slot[seg];
    StreamHdr st = a.t.state[slot].h;
    const StepIO io = make_io(a, slot, &st);
    SegResult res;
    run_segment(io, a.data, a.chunk_off, a.seg_chunk[seg], a.seg_chunk[seg + 1], res);
    a.t.state[slot].h = st;
    a.seg_out[seg] = res;
}

// ---- plain re-emit (used with k_general): out[i] = in[i], 16 bytes per thread per trip ---------------
__global__ void __launch_bounds__(256) k_copy(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint32_t n) {
    const uint32_t nvec = n >> 4;
    const uint4* __restrict__ vi = reinterpret_cast<const uint4*>(in);
    uint4* __restrict__ vo = reinterpret_cast<uint4*>(out);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x) vo[i] = __ldg(vi + i);
    const uint32_t tail = nvec << 4;
    if (blockIdx.x == 0 && threadIdx.x < (n - tail)) out[tail + threadIdx.x] = in[tail + threadIdx.x];
}

#include "relay_kernels.cuh"

// Launch one step on `stream`.  ev[0..3] bracket prime / relay / commit.
static inline cudaError_t launch_step(const StepArgs& a, int mode, int sm_count, cudaStream_t stream, cudaEvent_t* ev, int* launched) {
    *launched = 0;
    cudaError_t r;
    if ((r = cudaEventRecord(ev[0], stream)) != cudaSuccess) return r;
    if (mode == 1) {
        if ((r = cudaEventRecord(ev[1], stream)) != cudaSuccess) return r;
        if (a.n_bytes) { k_copy<<<sm_count * 8, 256, 0, stream>>>(a.data, a.out, a.n_bytes); ++*launched; }
        if ((r = cudaEventRecord(ev[2], stream)) != cudaSuccess) return r;
        if (a.n_segs) { k_general<<<(a.n_segs + 63) / 64, 64, 0, stream>>>(a); ++*launched; }
        if ((r = cudaEventRecord(ev[3], stream)) != cudaSuccess) return r;
        return cudaGetLastError();
    }
    return launch_step_fast(a, sm_count, stream, ev, launched);
}

}  // namespace lgw
