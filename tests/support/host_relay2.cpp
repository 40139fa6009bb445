// TEST AID ONLY: the bulk kernels of csrc/relay2.cuh (k_prime2, k_relay2, k_commit2) compiled for the
// host over a small SIMT emulator (simt_emu.h: fibers as lanes, rendezvous collectives, instant TMA), so that the CPU
// test-suite can run the very kernel code against the sequential machine, the goldens and the oracle without a GPU.
// Never linked into the product library; the product has no CPU path.
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <vector>

#include "simt_emu.h"
#include "../../include/llmgw_b200.h"
#include "../../llmapigateway_b200/csrc/step_types.cuh"

namespace lgw {
#include "../../llmapigateway_b200/csrc/relay2.cuh"
}
#include "../../llmapigateway_b200/csrc/transcript.cuh"

using namespace lgw;

struct HostEngine {
    uint32_t max_streams, carry_cap, detail_cap, rowq_cap, n_blocks;
    std::vector<StreamState> state;
    std::vector<uint8_t> carry_a, carry_b, detail, pending;
    std::vector<SegPlan> plan;
    std::vector<uint8_t> cache;          // TemplateCache2
    std::vector<uint2> fields;
    uint32_t counters[16] = {0};
    std::vector<uint32_t> tile_seg;
    std::vector<RowEvent> rowq; uint32_t rowq_count;
    int mode;
    // transcript tap (transcript.cuh)
    std::vector<TextTap> tap; std::vector<uint8_t> tcarry, tsparse, ttext;
    std::vector<uint32_t> piece_len, tseg_len, tseg_flags; std::vector<unsigned long long> tseg_off;
    std::vector<TextMark> markq; uint32_t markq_count = 0;
};

extern "C" {

void* lgwt_bulk_new(uint32_t max_streams, uint32_t carry_cap, uint32_t detail_cap, uint32_t rowq_cap, uint32_t n_blocks) {
    HostEngine* e = new HostEngine();
    e->max_streams = max_streams; e->carry_cap = carry_cap; e->detail_cap = detail_cap; e->rowq_cap = rowq_cap; e->n_blocks = n_blocks ? n_blocks : 1;
    e->state.assign(max_streams, StreamState{});
    e->carry_a.assign((size_t)max_streams * carry_cap, 0); e->carry_b.assign((size_t)max_streams * carry_cap, 0);
    e->detail.assign((size_t)max_streams * detail_cap, 0); e->pending.assign((size_t)max_streams * LGW_PENDING_STRIDE, 0);
    e->plan.assign(max_streams, SegPlan{});
    e->cache.assign(sizeof(TemplateCache2), 0);
    e->fields.assign((size_t)max_streams * 9, uint2{0, 0});
    e->rowq.assign(rowq_cap + 1, RowEvent{}); e->rowq_count = 0;
    e->mode = 0;
    return e;
}
void lgwt_bulk_free(void* h) { delete (HostEngine*)h; }
void lgwt_bulk_set_mode(void* h, int mode) { ((HostEngine*)h)->mode = mode; }
void lgwt_bulk_reset_templates(void* h) { HostEngine* e = (HostEngine*)h; memset(e->cache.data(), 0, e->cache.size()); }

void lgwt_bulk_open(void* h, const uint32_t* slots, const int32_t* status, uint32_t n) {
    HostEngine* e = (HostEngine*)h;
    for (uint32_t i = 0; i < n; ++i) init_stream(e->state[slots[i]], status[i]);
    if (!e->tap.empty()) simt::launch((n + 127) / 128, 128, [&] { k_text_open(e->tap.data(), slots, n); });
}

void lgwt_bulk_transcripts_enable(void* h) {
    HostEngine* e = (HostEngine*)h;
    e->tap.assign(e->max_streams, TextTap{});
    e->tcarry.assign((size_t)e->max_streams * e->carry_cap, 0);
    e->markq.assign(e->rowq_cap + 1, TextMark{});
}

// the transcript pass over the arrays of the step that has just run (seg_res = that step's results); text_out has room for
// 2 * (n_bytes + n_segs * carry_cap) bytes.  Returns the number of marks.
uint32_t lgwt_bulk_transcript(void* h, const uint8_t* data, uint32_t n_bytes, const uint32_t* chunk_off, uint32_t n_chunks,
                              const uint32_t* seg_chunk, const uint32_t* seg_slot, uint32_t n_segs, const lgw_seg_result* seg_res,
                              uint8_t* text_out, uint64_t* seg_text_off, uint32_t* seg_flags, lgw_text_mark* marks_out, uint32_t marks_cap) {
    HostEngine* e = (HostEngine*)h;
    // exactly sized scratch: the emulated kernels must not touch a byte beyond it
    e->tsparse.assign(2 * ((size_t)n_bytes + (size_t)n_segs * e->carry_cap), 0xEE);
    e->ttext.assign(2 * ((size_t)n_bytes + (size_t)n_segs * e->carry_cap) + 1, 0xEE);
    e->piece_len.assign(n_chunks + 1, 0xEEEEEEEEu); e->tseg_len.assign(n_segs + 1, 0); e->tseg_flags.assign(n_segs + 1, 0); e->tseg_off.assign(n_segs + 2, 0);
    e->markq_count = 0;
    TextArgs a{};
    a.data = data; a.chunk_off = chunk_off; a.seg_chunk = seg_chunk; a.seg_slot = seg_slot; a.n_segs = n_segs; a.seg_res = (const SegResult*)seg_res;
    a.tap = e->tap.data(); a.carry = e->tcarry.data(); a.carry_cap = e->carry_cap; a.sparse = e->tsparse.data();
    a.piece_len = e->piece_len.data(); a.seg_len = e->tseg_len.data(); a.seg_flags = e->tseg_flags.data(); a.seg_off = e->tseg_off.data(); a.text = e->ttext.data();
    a.markq = e->markq.data(); a.markq_count = &e->markq_count; a.markq_cap = e->rowq_cap;
    if (n_segs) {
        simt::launch((n_segs + TX_WARPS - 1) / TX_WARPS, TX_WARPS * 32, [&] { k_text_extract(a); });
        simt::launch(1, 1024, [&] { k_text_scan(a.seg_len, n_segs, a.seg_off); });
        simt::launch((n_segs + TX_WARPS - 1) / TX_WARPS, TX_WARPS * 32, [&] { k_text_pack(a); });
    }
    const unsigned long long total = e->tseg_off[n_segs];
    memcpy(text_out, e->ttext.data(), total);
    for (uint32_t i = 0; i <= n_segs; ++i) seg_text_off[i] = e->tseg_off[i];
    memcpy(seg_flags, e->tseg_flags.data(), (size_t)n_segs * 4);
    uint32_t cnt = e->markq_count < e->rowq_cap ? e->markq_count : e->rowq_cap;
    if (cnt > marks_cap) cnt = marks_cap;
    memcpy(marks_out, e->markq.data(), (size_t)cnt * sizeof(TextMark));
    return cnt;
}

static StepArgs make_args(HostEngine* e, const uint8_t* data, uint32_t n_bytes, const uint32_t* chunk_off, uint32_t n_chunks,
                          const uint32_t* seg_chunk, const uint32_t* seg_slot, uint32_t n_segs, uint8_t* out, lgw_seg_result* seg_out) {
    StepArgs a{};
    a.t.state = e->state.data(); a.t.carry_a = e->carry_a.data(); a.t.carry_b = e->carry_b.data(); a.t.detail = e->detail.data(); a.t.pending = e->pending.data();
    a.t.carry_cap = e->carry_cap; a.t.detail_cap = e->detail_cap; a.t.max_streams = e->max_streams;
    a.data = data; a.n_bytes = n_bytes; a.chunk_off = chunk_off; a.n_chunks = n_chunks; a.tile_base = 0; a.chunk_lo = 0; a.chunk_hi = n_chunks;
    a.seg_chunk = seg_chunk; a.seg_slot = seg_slot; a.n_segs = n_segs; a.out = out; a.seg_out = (SegResult*)seg_out;
    a.rowq = e->rowq.data(); a.rowq_count = &e->rowq_count; a.rowq_cap = e->rowq_cap;
    a.s.plan = e->plan.data(); a.s.tpl_cache2 = (TemplateCache2*)e->cache.data(); a.s.usage_fields = e->fields.data();
    a.s.counters = e->counters;
    e->tile_seg.assign(n_bytes / R2_TILE + 4, 0); a.s.tile_seg = e->tile_seg.data();
    return a;
}

// one step over a packed batch (same layout as lgw_sse_step).  tiles_per_warp = 0: spread the tiles over all warps.
int lgwt_bulk_step(void* h, const uint8_t* data, uint32_t n_bytes, const uint32_t* chunk_off, uint32_t n_chunks,
                   const uint32_t* seg_chunk, const uint32_t* seg_slot, uint32_t n_segs, uint8_t* out, lgw_seg_result* seg_out,
                   lgw_row_event* rows_out, uint32_t rows_cap, uint32_t* n_rows, uint32_t tiles_per_warp) {
    HostEngine* e = (HostEngine*)h;
    StepArgs a = make_args(e, data, n_bytes, chunk_off, n_chunks, seg_chunk, seg_slot, n_segs, out, seg_out);
    e->rowq_count = 0;
    if (e->mode == 1) {
        memcpy(out, data, n_bytes);
        for (uint32_t seg = 0; seg < n_segs; ++seg) {
            const uint32_t slot = seg_slot[seg];
            StreamHdr st = a.t.state[slot].h;
            const StepIO io = make_io(a, slot, &st);
            SegResult res;
            run_segment(io, a.data, a.chunk_off, a.seg_chunk[seg], a.seg_chunk[seg + 1], res);
            a.t.state[slot].h = st;
            a.seg_out[seg] = res;
        }
    } else {
        if (n_segs) simt::launch((n_segs + 127) / 128, 128, [&] { k_prime2(a); });
        const uint32_t n_tiles = (n_bytes + R2_TILE - 1) / R2_TILE;
        if (n_tiles) {
            const uint32_t max_warps = e->n_blocks * R2_WARPS;
            uint32_t tpw = tiles_per_warp ? tiles_per_warp : (n_tiles + max_warps - 1) / max_warps;
            uint32_t warps = (n_tiles + tpw - 1) / tpw;
            if (warps > max_warps) { tpw = (n_tiles + max_warps - 1) / max_warps; warps = (n_tiles + tpw - 1) / tpw; }
            const uint32_t blocks = (warps + R2_WARPS - 1) / R2_WARPS;
            simt::launch(blocks, R2_THREADS, [&] { k_relay2(a, n_tiles, tpw, 0u); });
        }
        if (n_segs) simt::launch((n_segs + R2_CWARPS - 1) / R2_CWARPS, R2_CWARPS * 32, [&] { k_commit2(a); });
    }
    uint32_t cnt = e->rowq_count < e->rowq_cap ? e->rowq_count : e->rowq_cap;
    if (cnt > rows_cap) cnt = rows_cap;
    memcpy(rows_out, e->rowq.data(), (size_t)cnt * sizeof(RowEvent));
    *n_rows = cnt;
    return 0;
}

void lgwt_bulk_state(void* h, const uint32_t* slots, uint32_t n, lgw_stream_state* out, int free_after) {
    HostEngine* e = (HostEngine*)h;
    for (uint32_t i = 0; i < n; ++i) {
        StreamState* s = &e->state[slots[i]];
        if (s->h.flags & SF_PENDING) {
            StreamHdr st = s->h;
            StepIO io;
            io.st = &st; io.rec = &s->rec; io.pending = e->pending.data() + (size_t)slots[i] * LGW_PENDING_STRIDE;
            io.carry_a = io.carry_b = io.detail = nullptr; io.carry_cap = io.detail_cap = 0;
            io.rowq = nullptr; io.rowq_count = nullptr; io.rowq_cap = 0; io.slot = slots[i];
            resolve_pending(io);
            s->h = st;
        }
        memcpy(&out[i], s, sizeof(StreamState));
        if (free_after) s->h.phase = PH_FREE;
    }
}

uint32_t lgwt_bulk_detail(void* h, uint32_t slot, uint8_t* buf, uint32_t cap) {
    HostEngine* e = (HostEngine*)h;
    uint32_t n = e->state[slot].h.detail_len < cap ? e->state[slot].h.detail_len : cap;
    memcpy(buf, e->detail.data() + (size_t)slot * e->detail_cap, n);
    return n;
}

void lgwt_bulk_counters(void* h, uint32_t* out) { memcpy(out, ((HostEngine*)h)->counters, 16); }

// diagnostics: out[0..15] = state[4], len[4], flags[4] (bit 31: usage_ok), hits[4]
void lgwt_bulk_templates(void* h, uint32_t* out) {
    HostEngine* e = (HostEngine*)h;
    const TemplateCache2* tc = (const TemplateCache2*)e->cache.data();
    for (int i = 0; i < 4; ++i) { out[i] = tc->state[i]; out[4 + i] = tc->tpl[i].m.len; out[8 + i] = tc->tpl[i].m.flags | (tc->tpl[i].m.usage_ok << 31); out[12 + i] = tc->hits[i]; }
}

}  // extern "C"
