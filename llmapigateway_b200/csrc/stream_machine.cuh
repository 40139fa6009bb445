// Per-stream relay/tap state machine: the exact sequential semantics of one upstream attempt.
//
//   priming   request_handler.py:21-98   stream_generator + priming loop (one shared carry: the
//             two loops see the same chunks and apply the same split rule, so their buffers are
//             always equal -- SURVEY.md Appendix A.1 items 4-7)
//   relay A   request_handler.py:100-144 combined_generator (its carry starts EMPTY after the
//             kept chunk)
//   tap B     chat_logging.py:87-150     ChunkProcessorThread.run over the emitted chunks (its
//             carry includes the kept chunk), get_token_usage :233-272
//
// Emitted bytes: once a stream commits, every later chunk is relayed verbatim (:141-142), so
// the kept chunks of a step are always a SUFFIX of the stream's segment; the engine reports the
// first kept chunk index and the bulk kernel copies bytes position-for-position.
//
// Host/device portable (tests compile it with g++ as a test aid; the product has no CPU path).
#pragma once
#include "json_machine.cuh"
#include "lean_json.cuh"

namespace lgw {

#define LGW_PENDING_CAP 1024u
#define LGW_PENDING_STRIDE (LGW_PENDING_CAP + 32u)     /* whole 16-byte vectors around the event text */

enum Phase : uint8_t { PH_FREE = 0, PH_PRIMING = 1, PH_COMMITTED = 2, PH_FAILED = 3 };
enum Verdict : uint8_t {
    VD_NONE = 0,
    VD_OK = 1,            // first real event accepted (request_handler.py:89-90)
    VD_FAIL_EVENT = 2,    // first real event has top-level error/detail (:50-54, :86-88); detail = the part
    VD_FAIL_PARSE = 3,    // first real event does not parse (:85 escapes to :183-187); detail = the part
    VD_FAIL_HTTP = 4      // upstream status >= 400 (:25-30); host owns the body
};
enum StreamFlag : uint16_t {
    SF_A_USAGE_BOUND = 1 << 0,   // request_handler.py:134 ran at least once (else :144 raises)
    SF_EMITTED_ANY = 1 << 1,     // at least one chunk was relayed (a tap thread exists, chat_logging.py:200)
    SF_CARRY_OVERFLOW = 1 << 2,  // an unterminated event outgrew the carry capacity (engine limit)
    SF_EXOTIC_SEEN = 1 << 3,     // some event had a shape the device does not model (counted)
    SF_SYNCED = 1 << 4,          // carry A == carry B (only carry A is stored)
    SF_REC_VALID = 1 << 5,       // `rec` holds a get_token_usage result
    SF_DETAIL_TRUNC = 1 << 6,
    SF_ROWQ_OVERFLOW = 1 << 7,
    SF_PENDING = 1 << 8          // a usage event's text is stashed; its values are extracted on demand
};

// get_token_usage's result (chat_logging.py:233-272).  Absent model/provider => key not in dict.
struct UsageRec {
    Val prompt, completion, total, reasoning, cached, cost;
    Val model_val, provider_val;       // kind KD_ABSENT: key missing; KD_STR: text below
    uint8_t model_len, provider_len;
    uint8_t str_flags;                 // bit0 model truncated, bit1 model lone surrogate, bit2/3 same for provider
    uint8_t exotic;                    // a value the device cannot represent (big int, string/array where a number goes ...)
    char model[LGW_STR_CAP];
    char provider[LGW_STR_CAP];
};

struct StreamHdr {            // the 64 hot bytes of a stream's state (kernels work on a local copy)
    uint8_t phase, verdict;
    uint16_t flags;
    uint32_t carry_a_len, carry_b_len, detail_len;
    uint32_t n_events_a;       // real events parsed by the handler loop
    uint32_t n_events_b;       // events the tap parsed as JSON
    uint32_t n_usage_b;        // of those, how many updated the usage record
    uint32_t n_exotic;
    uint32_t n_error_rows;     // tap "error" events => extra rows (chat_logging.py:137-139)
    uint32_t n_chunks_in, n_chunks_emitted;
    uint32_t pending_len;      // bytes of stashed usage event (SF_PENDING)
    uint64_t bytes_in, bytes_emitted;
};
struct StreamState {          // == lgw_stream_state
    StreamHdr h;
    UsageRec rec;
};
static_assert(sizeof(StreamHdr) == 64, "StreamHdr must be 64 bytes");

struct RowEvent {            // one write_log call that happened mid-stream (chat_logging.py:139)
    uint32_t slot, seq;
    UsageRec rec;
};

struct SegResult {           // per segment, per step
    uint32_t emit_chunk_begin;     // first relayed chunk of this step (== segment end when none)
    uint8_t phase, verdict;
    uint16_t flags;
    uint32_t detail_len;
};

// ---- UTF-8 (CPython's strict decoder: no overlongs, no surrogates, <= U+10FFFF) ---------------
LGW_HD bool utf8_valid(const uint8_t* p, uint32_t n) {
    uint32_t i = 0;
    while (i < n) {
        const uint32_t c = p[i];
        if (c < 0x80) { ++i; continue; }
        uint32_t need; uint32_t lo = 0x80, hi = 0xBF;
        if (c >= 0xC2 && c <= 0xDF) need = 1;
        else if (c == 0xE0) { need = 2; lo = 0xA0; }
        else if (c >= 0xE1 && c <= 0xEC) need = 2;
        else if (c == 0xED) { need = 2; hi = 0x9F; }
        else if (c >= 0xEE && c <= 0xEF) need = 2;
        else if (c == 0xF0) { need = 3; lo = 0x90; }
        else if (c >= 0xF1 && c <= 0xF3) need = 3;
        else if (c == 0xF4) { need = 3; hi = 0x8F; }
        else return false;
        if (i + need >= n) return false;
        uint32_t b = p[i + 1];
        if (b < lo || b > hi) return false;
        for (uint32_t k = 2; k <= need; ++k) { b = p[i + k]; if (b < 0x80 || b > 0xBF) return false; }
        i += need + 1;
    }
    return true;
}

// ---- two-piece rope: carry bytes then chunk bytes ----------------------------------------------
struct Rope {
    const uint8_t* a; uint32_t na;
    const uint8_t* b; uint32_t nb;
    LGW_HD uint32_t size() const { return na + nb; }
    LGW_HD uint32_t at(uint32_t i) const { return i < na ? a[i] : b[i - na]; }
};

enum PartClass : uint8_t { PC_NONE = 0, PC_DATA = 1 /* "data: {" */, PC_BRACE = 2 /* "{" (tap only) */ };

LGW_HD uint8_t classify_part(const Rope& r, uint32_t s, uint32_t e) {
    if (e <= s) return PC_NONE;
    const uint32_t c0 = r.at(s);
    if (c0 == '{') return PC_BRACE;
    if (c0 != 'd' || e - s < 7) return PC_NONE;
    return (r.at(s + 1) == 'a' && r.at(s + 2) == 't' && r.at(s + 3) == 'a' && r.at(s + 4) == ':' &&
            r.at(s + 5) == ' ' && r.at(s + 6) == '{') ? PC_DATA : PC_NONE;
}

// full machine over one classified part; returns TopKey|PartFlag bits
template <bool EXTRACT>
LGW_HD_NOINLINE uint32_t parse_part(const Rope& r, uint32_t s, uint32_t e, uint8_t cls, UsageRaw* raw) {
    JsonMachine<EXTRACT> m;
    m.reset(raw, cls == PC_DATA);
    uint32_t i = s + (cls == PC_DATA ? 6u : 0u);
    for (; i < e; ++i) {
        m.feed(r.at(i));
        if (m.failed()) break;
    }
    return m.finish();
}

// lean recogniser over one classified part: validity + error/detail/code/usage (lean_json.cuh)
LGW_HD uint32_t lean_part(const Rope& r, uint32_t s, uint32_t e, uint8_t cls) {
    const LeanTables& t = lean_tables();
    PlainEnv env{r.a, r.na, r.b, r.nb, t.cls, t.trans};
    return lean_parse(env, s + (cls == PC_DATA ? 6u : 0u), e, cls == PC_DATA);
}

// what the tap needs of one part: lean first; the full machine (choices walk, usage values) only
// for the rare events that carry "usage" or "error"
LGW_HD uint32_t tap_parse(const Rope& r, uint32_t s, uint32_t e, uint8_t cls, UsageRaw* raw) {
    uint32_t f = lean_part(r, s, e, cls);
    if ((f & PF_VALID_B) && (f & (TK_USAGE | TK_ERROR))) f = parse_part<true>(r, s, e, cls, raw);
    return f;
}

// ---- get_token_usage arithmetic (chat_logging.py:246-267) --------------------------------------
LGW_HD bool kind_unrepresentable(uint8_t k) { return k == KD_STR || k == KD_OBJ || k == KD_ARR || k == KD_BIG || k == KD_FLT_INEXACT; }

LGW_HD void copy_str(char* dst, uint8_t& dlen, const char* src, uint8_t n) { for (uint8_t i = 0; i < n; ++i) dst[i] = src[i]; dlen = n; }

LGW_HD void normalise_usage(const UsageRaw& raw, uint32_t part_flags, UsageRec& out) {
    const Val zero = {0, KD_INT};
    out.prompt = out.completion = out.total = out.reasoning = out.cached = out.cost = zero;   // :237-244
    out.model_val.kind = out.provider_val.kind = KD_ABSENT; out.model_val.bits = out.provider_val.bits = 0;
    out.model_len = out.provider_len = 0; out.str_flags = 0; out.exotic = 0;
    bool partial = false;
    if ((part_flags & TK_USAGE) && raw.usage_kind == KD_OBJ) {
        if (raw.prompt.kind != KD_ABSENT) out.prompt = raw.prompt;
        if (raw.completion.kind != KD_ABSENT) out.completion = raw.completion;
        if (raw.total.kind != KD_ABSENT) out.total = raw.total;
        if (raw.cost.kind != KD_ABSENT) out.cost = raw.cost;
        // "completion_tokens_details" in usage and "reasoning_tokens" in usage[...]        :256-258
        if (raw.ctd_kind != KD_ABSENT) {
            if (raw.ctd_kind == KD_OBJ) { if (raw.reasoning.kind != KD_ABSENT) out.reasoning = raw.reasoning; }
            else if (raw.ctd_kind == KD_STR || raw.ctd_kind == KD_ARR) out.exotic = 1;       // substring / membership test
            else partial = true;                                                             // TypeError: not iterable
        }
        if (!partial && !out.exotic && raw.ptd_kind != KD_ABSENT) {                          // :259-261
            if (raw.ptd_kind == KD_OBJ) { if (raw.cached.kind != KD_ABSENT) out.cached = raw.cached; }
            else if (raw.ptd_kind == KD_STR || raw.ptd_kind == KD_ARR) out.exotic = 1;
            else partial = true;
        }
        if (!partial && !out.exotic) {                                                       // :262-263
            const Val r = out.reasoning, c = out.completion;
            bool pos = false;
            if (r.kind == KD_INT) pos = r.bits > 0;
            else if (r.kind == KD_FLT) pos = bits2dbl((uint64_t)r.bits) > 0.0;
            else if (r.kind == KD_TRUE) pos = true;
            else if (r.kind == KD_FALSE) pos = false;
            else if (r.kind == KD_BIG) pos = r.bits > 0;
            else if (r.kind == KD_FLT_INEXACT) out.exotic = 1;
            else partial = true;                                   // None/str/list/dict > 0 raises
            if (pos && !partial) {
                const bool c_int = c.kind == KD_INT || c.kind == KD_TRUE || c.kind == KD_FALSE;
                const bool r_int = r.kind == KD_INT || r.kind == KD_TRUE;
                const int64_t ci = c.kind == KD_INT ? c.bits : (c.kind == KD_TRUE ? 1 : 0);
                const int64_t ri = r.kind == KD_INT ? r.bits : 1;
                if (c.kind == KD_NULL || c.kind == KD_STR || c.kind == KD_OBJ || c.kind == KD_ARR) partial = true;   // TypeError
                else if (c.kind == KD_BIG || r.kind == KD_BIG || c.kind == KD_FLT_INEXACT) out.exotic = 1;
                else if (c_int && r_int) {
                    int64_t d = (int64_t)((uint64_t)ci - (uint64_t)ri);
                    if (((ci ^ ri) & (ci ^ d)) < 0) out.exotic = 1;     // would need a Python big int
                    else { out.completion.kind = KD_INT; out.completion.bits = d; }
                } else {
                    const double cd = c_int ? (double)ci : bits2dbl((uint64_t)c.bits);
                    const double rd = r_int ? (double)ri : bits2dbl((uint64_t)r.bits);
                    out.completion.kind = KD_FLT; out.completion.bits = (int64_t)dbl2bits(cd - rd);
                }
            }
        }
    }
    if (kind_unrepresentable(out.prompt.kind) || kind_unrepresentable(out.completion.kind) || kind_unrepresentable(out.total.kind) ||
        kind_unrepresentable(out.cost.kind) || kind_unrepresentable(out.reasoning.kind) || kind_unrepresentable(out.cached.kind)) out.exotic = 1;
    if (!partial) {                                                                          // :264-267
        if (part_flags & TK_PROVIDER) {
            out.provider_val.kind = raw.provider_kind; out.provider_val.bits = raw.provider_val.bits;
            if (raw.provider_kind == KD_STR) { copy_str(out.provider, out.provider_len, raw.provider, raw.provider_len); out.str_flags |= (raw.provider_flags & 3) << 2; }
            else if (raw.provider_kind == KD_OBJ || raw.provider_kind == KD_ARR || raw.provider_kind == KD_BIG || raw.provider_kind == KD_FLT_INEXACT) out.exotic = 1;
        }
        if (part_flags & TK_MODEL) {
            out.model_val.kind = raw.model_kind; out.model_val.bits = raw.model_val.bits;
            if (raw.model_kind == KD_STR) { copy_str(out.model, out.model_len, raw.model, raw.model_len); out.str_flags |= (raw.model_flags & 3); }
            else if (raw.model_kind == KD_OBJ || raw.model_kind == KD_ARR || raw.model_kind == KD_BIG || raw.model_kind == KD_FLT_INEXACT) out.exotic = 1;
        }
    }
    if (out.str_flags) out.exotic = 1;
}

LGW_HD void default_usage(UsageRec& out) {       // chat_logging.py:77-84
    const Val zero = {0, KD_INT};
    out.prompt = out.completion = out.total = out.reasoning = out.cached = out.cost = zero;
    out.model_val.kind = out.provider_val.kind = KD_ABSENT; out.model_val.bits = out.provider_val.bits = 0;
    out.model_len = out.provider_len = 0; out.str_flags = 0; out.exotic = 0;
}

// ---- the split rule (request_handler.py:38-40 etc.) over carry ++ chunk --------------------------
// Calls on_part(start, end) for every complete piece that could be an event (pieces before the
// last separator), in order; on_part returns false to stop early (priming `break`).
// Returns the start of the tail piece (== size() when the buffer ends with the separator, in
// which case the carry resets) and whether it stopped early.
template <class F>
LGW_HD uint32_t split_scan(const Rope& r, F&& on_part, bool& stopped) {
    const uint32_t total = r.size();
    uint32_t pos = 0;
    uint32_t i = r.na > 0 ? r.na - 1 : 0;        // a carry never contains a separator
    stopped = false;
    while (i + 1 < total) {
        if (r.at(i) == '\n' && r.at(i + 1) == '\n') {
            if (!on_part(pos, i)) { stopped = true; return pos; }
            pos = i + 2; i += 2;
        } else ++i;
    }
    if (total >= 2 && r.at(total - 2) == '\n' && r.at(total - 1) == '\n') return total;   // endswith => carry ""
    return pos;
}

// store the tail [pos, total) of carry ++ chunk as the new carry
LGW_HD bool store_carry(uint8_t* carry, uint32_t& carry_len, uint32_t cap, const Rope& r, uint32_t pos) {
    const uint32_t total = r.size();
    const uint32_t n = total - pos;
    if (n > cap) { carry_len = 0; return false; }
    if (pos == 0) {                          // nothing consumed: append the chunk behind the old carry
        for (uint32_t k = 0; k < r.nb; ++k) carry[r.na + k] = r.b[k];
    } else {                                 // pos >= na (a separator cannot end inside the old carry... it can end at na)
        for (uint32_t k = 0; k < n; ++k) carry[k] = (uint8_t)r.at(pos + k);
    }
    carry_len = n;
    return true;
}

struct StepIO {                 // where one stream's step reads and writes
    StreamHdr* st;              // usually a local copy, written back by the caller
    UsageRec* rec;              // the stream's tap record (global memory)
    uint8_t* pending;           // stashed usage event text (LGW_PENDING_CAP bytes)
    uint8_t* carry_a; uint8_t* carry_b; uint8_t* detail;
    uint32_t carry_cap, detail_cap;
    RowEvent* rowq; uint32_t* rowq_count; uint32_t rowq_cap;
    uint32_t slot;
};

LGW_HD void save_detail(const StepIO& io, const Rope& r, uint32_t s, uint32_t e) {
    uint32_t n = e - s;
    if (n > io.detail_cap) { n = io.detail_cap; io.st->flags |= SF_DETAIL_TRUNC; }
    for (uint32_t k = 0; k < n; ++k) io.detail[k] = (uint8_t)r.at(s + k);
    io.st->detail_len = n;
}

LGW_HD void push_row(const StepIO& io) {
    StreamHdr& st = *io.st;
    ++st.n_error_rows;
#if defined(__CUDA_ARCH__)
    const uint32_t k = atomicAdd(io.rowq_count, 1u);
#else
    const uint32_t k = (*io.rowq_count)++;
#endif
    if (k >= io.rowq_cap) { st.flags |= SF_ROWQ_OVERFLOW; return; }
    RowEvent& ev = io.rowq[k];
    ev.slot = io.slot; ev.seq = st.n_error_rows;
    if (st.flags & SF_REC_VALID) ev.rec = *io.rec; else default_usage(ev.rec);
}

// tap handling of one parsed part (chat_logging.py:123-141).  The choices walk (and any shape of it
// the device does not model) is observable only through events that also carry "usage" or "error".
LGW_HD void tap_part(const StepIO& io, uint32_t f, const UsageRaw& raw) {
    StreamHdr& st = *io.st;
    if (!(f & PF_VALID_B)) return;
    ++st.n_events_b;
    if (!(f & (TK_USAGE | TK_ERROR))) return;
    if (f & PF_EXOTIC) { ++st.n_exotic; st.flags |= SF_EXOTIC_SEEN; return; }
    if ((f & TK_CHOICES) && (f & PF_TYPE_ERROR)) return;
    if (f & TK_USAGE) {
        normalise_usage(raw, f, *io.rec);
        st.flags |= SF_REC_VALID; ++st.n_usage_b;
        if (io.rec->exotic) { ++st.n_exotic; st.flags |= SF_EXOTIC_SEEN; }
    }
    if (f & TK_ERROR) push_row(io);
}

// The bulk path stashes the text of the stream's winning usage event instead of extracting its
// values on the spot (k_commit); this is the deferred chat_logging.py:134-135 for that event.
LGW_HD_NOINLINE void resolve_pending(const StepIO& io) {
    StreamHdr& st = *io.st;
    if (!(st.flags & SF_PENDING)) return;
    st.flags &= ~(uint16_t)SF_PENDING;
    const uint32_t n = st.pending_len & 0xFFFFu, off = st.pending_len >> 16;     // k_commit stores whole 16-byte vectors
    st.pending_len = 0;
    Rope r{nullptr, 0, io.pending + off, n};
    const uint8_t cls = classify_part(r, 0, n);
    UsageRaw raw;
    const uint32_t f = parse_part<true>(r, 0, n, cls, &raw);
    if (f & PF_EXOTIC) { ++st.n_exotic; st.flags |= SF_EXOTIC_SEEN; --st.n_usage_b; return; }
    if ((f & TK_CHOICES) && (f & PF_TYPE_ERROR)) { --st.n_usage_b; return; }
    normalise_usage(raw, f, *io.rec);
    st.flags |= SF_REC_VALID;
    if (io.rec->exotic) { ++st.n_exotic; st.flags |= SF_EXOTIC_SEEN; }
}

LGW_HD void handler_part(StreamHdr& st, uint32_t f) {       // request_handler.py:122-134
    ++st.n_events_a;
    if (!(f & PF_VALID_A)) return;
    if (f & TK_CODE) return;                                   // Appendix A.1 item 9
    if (f & TK_USAGE) st.flags |= SF_A_USAGE_BOUND;
}

// One chunk through the committed-phase loops.  `tap_only`: the kept chunk itself (the handler
// never splits it: combined_generator's buffer starts empty after it).
// While the two carries differ (SF_SYNCED clear) each loop scans with its own carry; they become
// equal for good as soon as both loops end a chunk on the same separator.
LGW_HD void relay_chunk(const StepIO& io, const uint8_t* p, uint32_t n, bool tap_only) {
    StreamHdr& st = *io.st;
    if (!utf8_valid(p, n)) return;            // both loops swallow the decode error; carries unchanged
    UsageRaw raw;
    bool stopped;
    if ((st.flags & SF_SYNCED) && !tap_only) {
        Rope r{io.carry_a, st.carry_a_len, p, n};
        const uint32_t tail = split_scan(r, [&](uint32_t s, uint32_t e) {
            const uint8_t cls = classify_part(r, s, e);
            if (cls == PC_NONE) return true;
            const uint32_t f = tap_parse(r, s, e, cls, &raw);
            if (cls == PC_DATA) handler_part(st, f);
            tap_part(io, f, raw);
            return true; }, stopped);
        if (!store_carry(io.carry_a, st.carry_a_len, io.carry_cap, r, tail)) st.flags |= SF_CARRY_OVERFLOW;
        return;
    }
    const uint32_t NONE = 0xFFFFFFFFu;
    uint32_t qa = NONE, qb = NONE;            // chunk position right after the last separator each loop consumed
    if (!tap_only) {
        Rope r{io.carry_a, st.carry_a_len, p, n};
        const uint32_t tail = split_scan(r, [&](uint32_t s, uint32_t e) {
            const uint8_t cls = classify_part(r, s, e);
            if (cls != PC_DATA) return true;
            handler_part(st, lean_part(r, s, e, cls));
            return true; }, stopped);
        if (tail != 0) qa = tail - st.carry_a_len;
        if (!store_carry(io.carry_a, st.carry_a_len, io.carry_cap, r, tail)) st.flags |= SF_CARRY_OVERFLOW;
    } else {
        st.carry_a_len = 0;
    }
    {
        Rope r{io.carry_b, st.carry_b_len, p, n};
        const uint32_t tail = split_scan(r, [&](uint32_t s, uint32_t e) {
            const uint8_t cls = classify_part(r, s, e);
            if (cls == PC_NONE) return true;
            const uint32_t f = tap_parse(r, s, e, cls, &raw);
            tap_part(io, f, raw);
            return true; }, stopped);
        if (tail != 0) qb = tail - st.carry_b_len;
        if (!store_carry(io.carry_b, st.carry_b_len, io.carry_cap, r, tail)) st.flags |= SF_CARRY_OVERFLOW;
    }
    if (!tap_only && qa != NONE && qa == qb) st.flags |= SF_SYNCED;
}

// One chunk while priming (request_handler.py:34-58 and :69-95).  Returns true when this chunk
// is the kept one (the stream committed on it).
LGW_HD bool prime_chunk(const StepIO& io, const uint8_t* p, uint32_t n) {
    StreamHdr& st = *io.st;
    if (!utf8_valid(p, n)) return false;                     // sniffer swallows, priming drops (:94-95)
    Rope r{io.carry_a, st.carry_a_len, p, n};
    bool stopped;
    bool kept = false;
    const uint32_t tail = split_scan(r, [&](uint32_t s, uint32_t e) {
        const uint8_t cls = classify_part(r, s, e);
        if (cls != PC_DATA) return true;
        const uint32_t f = lean_part(r, s, e, cls);
        ++st.n_events_a;
        if (!(f & PF_VALID_A)) { st.phase = PH_FAILED; st.verdict = VD_FAIL_PARSE; save_detail(io, r, s, e); }
        else if (f & (TK_ERROR | TK_DETAIL)) { st.phase = PH_FAILED; st.verdict = VD_FAIL_EVENT; save_detail(io, r, s, e); }
        else { st.phase = PH_COMMITTED; st.verdict = VD_OK; kept = true; }
        return false; }, stopped);
    if (!stopped) {
        if (!store_carry(io.carry_a, st.carry_a_len, io.carry_cap, r, tail)) st.flags |= SF_CARRY_OVERFLOW;
        return false;
    }
    st.carry_a_len = 0;
    return kept;
}

// One chunk (n > 0 bytes at p) of a PRIMING or COMMITTED stream.  Returns true when the stream
// committed on this chunk (it is the kept one).
LGW_HD bool step_chunk(const StepIO& io, const uint8_t* p, uint32_t n) {
    StreamHdr& st = *io.st;
    if (st.phase == PH_PRIMING) {
        ++st.n_chunks_in; st.bytes_in += n;
        if (prime_chunk(io, p, n)) {
            st.flags |= SF_EMITTED_ANY;
            st.flags &= ~(uint16_t)SF_SYNCED;
            st.carry_b_len = 0;
            ++st.n_chunks_emitted; st.bytes_emitted += n;
            relay_chunk(io, p, n, true);                    // the tap sees the kept chunk
            if (st.carry_b_len == 0) st.flags |= SF_SYNCED; // kept chunk ended on a separator
            return true;
        }
    } else if (st.phase == PH_COMMITTED) {
        ++st.n_chunks_in; st.bytes_in += n;
        ++st.n_chunks_emitted; st.bytes_emitted += n;
        relay_chunk(io, p, n, false);
    }
    return false;
}

// Chunks [c_from, c_to) of one stream, sequentially.  `emit_begin` tracks the first relayed chunk.
// With stop_at_commit the walk returns right after the kept chunk (the stream has just committed);
// the return value is the first chunk not consumed.
LGW_HD uint32_t run_chunks(const StepIO& io, const uint8_t* data, const uint32_t* chunk_off,
                           uint32_t c_from, uint32_t c_to, uint32_t& emit_begin, bool stop_at_commit) {
    StreamHdr& st = *io.st;
    if (st.flags & SF_PENDING) resolve_pending(io);          // the sequential path works on extracted records
    for (uint32_t c = c_from; c < c_to; ++c) {
        const uint32_t o = chunk_off[c], n = chunk_off[c + 1] - o;
        if (n == 0) continue;                                   // never yielded (request_handler.py:60-63)
        if (st.phase != PH_PRIMING && st.phase != PH_COMMITTED) return c_to;   // the generator is gone
        if (step_chunk(io, data + o, n)) {
            emit_begin = c;
            if (stop_at_commit) return c + 1;
        }
    }
    return c_to;
}

LGW_HD void fill_seg_result(const StreamHdr& st, uint32_t emit_begin, uint32_t c1, SegResult& res) {
    res.emit_chunk_begin = (st.phase == PH_COMMITTED) ? emit_begin : c1;
    res.phase = st.phase; res.verdict = st.verdict; res.flags = st.flags; res.detail_len = st.detail_len;
}

// A whole segment (this step's chunks of one stream), sequentially.
LGW_HD void run_segment(const StepIO& io, const uint8_t* data, const uint32_t* chunk_off,
                        uint32_t c0, uint32_t c1, SegResult& res) {
    uint32_t emit_begin = (io.st->phase == PH_COMMITTED) ? c0 : c1;
    run_chunks(io, data, chunk_off, c0, c1, emit_begin, false);
    fill_seg_result(*io.st, emit_begin, c1, res);
}

LGW_HD void init_stream(StreamState& s, int http_status) {
    StreamHdr& st = s.h;
    st.phase = http_status >= 400 ? PH_FAILED : PH_PRIMING;
    st.verdict = http_status >= 400 ? VD_FAIL_HTTP : VD_NONE;
    st.flags = 0; st.carry_a_len = st.carry_b_len = st.detail_len = 0;
    st.n_events_a = st.n_events_b = st.n_usage_b = st.n_exotic = st.n_error_rows = 0;
    st.n_chunks_in = st.n_chunks_emitted = 0; st.pending_len = 0; st.bytes_in = st.bytes_emitted = 0;
    default_usage(s.rec);
}

}  // namespace lgw
