// Argument and scratch structures of the SSE step kernels (shared by sse_kernels.cuh and, as a test aid, the host
// build of relay2.cuh in tests/support).
#pragma once
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
    uint32_t prime_ok;         // set by the warp that walks the kept chunk when the speculation holds
    uint32_t tail_start;       // where the event that is open at the end of the text begins (= new carry); 0xFFFFFFFF: not reached
    uint32_t cand_ps;          // 1 + start of the usage event whose field spans sit in usage_fields[9 * seg ..] (0: none)
};
static_assert(sizeof(SegPlan) == 64, "SegPlan");

struct TemplateCache2;     // relay2.cuh

struct StepScratch {
    SegPlan* plan;             // [max_streams]
    TemplateCache2* tpl_cache2;  // event templates, persistent across steps
    uint2* usage_fields;       // [max_streams][9] a template-following usage event of this step: (start, length | escapes << 31) of its eight
                               //   usage fields as the match located them, [8].x = the template's slot in the engine-wide cache
    uint32_t* tile_seg;        // [max tiles + 1] segment that holds the first byte of each 4 KiB tile of the launch
    uint32_t* counters;        // diagnostics since engine creation: [0] segments redone sequentially, [1] segments folded from the bulk
                               //   kernel's findings, [2] usage records read from template spans, [3] usage events stashed
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

LGW_HD StepIO make_io(const StepArgs& a, uint32_t slot, StreamHdr* local_hdr) {
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

}  // namespace lgw
