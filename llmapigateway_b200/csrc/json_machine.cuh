// Streaming JSON recogniser + tracked-path extractor for one SSE event ("part").
//
// This is device code (sm_100a); every function is also __host__ so that tests/support can
// compile it with g++ and fuzz it against CPython on the CPU box (test aid only -- the product
// library has no CPU path).
//
// Grammar: RFC 8259 exactly as CPython's json.loads accepts it (strict=True): whitespace is
// SP/HT/LF/CR; strings reject raw control bytes < 0x20 and unknown escapes; \uXXXX accepts any
// four hex digits (lone surrogates included); numbers follow -?(0|[1-9]\d*)(\.\d+)?([eE][-+]?\d+)?
// plus the literals NaN, Infinity, -Infinity; duplicate keys are allowed and the LAST one wins.
// That is the strict-JSON subset of what the reference's `json5.loads` accepts;
// JSON5-only syntax is reported as invalid (documented in DESIGN.md).
//
// What is tracked (the only things the reference reads out of a parsed event):
//   top level   error, detail            request_handler.py:50,86 (first real event)
//               code, usage              request_handler.py:123,133
//               choices, usage, error    chat_logging.py:124,134,137
//               model, provider          chat_logging.py:264-267
//   usage.*     prompt_tokens, completion_tokens, total_tokens, cost,
//               completion_tokens_details.reasoning_tokens,
//               prompt_tokens_details.cached_tokens            chat_logging.py:246-261
//   choices[*]  delta.content / message.content                 chat_logging.py:125-133
// including the Python exception behaviour those reads have on odd shapes (a TypeError skips
// the whole event: chat_logging.py:140-141).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LGW_HD __host__ __device__ __forceinline__
#define LGW_HD_NOINLINE __host__ __device__ __noinline__
#else
#define LGW_HD inline
#define LGW_HD_NOINLINE inline
#endif

#include "decimal.cuh"

namespace lgw {

// ---- value kinds ---------------------------------------------------------------------------
enum Kind : uint8_t {
    KD_ABSENT = 0, KD_INT = 1, KD_FLT = 2, KD_NULL = 3, KD_TRUE = 4, KD_FALSE = 5,
    KD_STR = 6, KD_BIG = 7 /* integer outside int64 */, KD_OBJ = 8, KD_ARR = 9,
    KD_FLT_INEXACT = 10 /* decimal->double needed more than the on-device exact paths */
};

// ---- top-level key flags -------------------------------------------------------------------
enum TopKey : uint32_t {
    TK_ERROR = 1u << 0, TK_DETAIL = 1u << 1, TK_CODE = 1u << 2, TK_USAGE = 1u << 3,
    TK_CHOICES = 1u << 4, TK_MODEL = 1u << 5, TK_PROVIDER = 1u << 6
};

// ---- part outcome flags --------------------------------------------------------------------
enum PartFlag : uint32_t {
    PF_VALID_A = 1u << 8,      // json.loads(part[6:]) succeeds          (handler loops)
    PF_VALID_B = 1u << 9,      // json.loads(text after strip()) succeeds (tap loop)
    PF_TYPE_ERROR = 1u << 10,  // the choices walk raises (tap skips the event)
    PF_EXOTIC = 1u << 11,      // a shape whose Python behaviour is not modelled on device
    PF_TOO_DEEP = 1u << 12,    // nesting beyond LGW_MAX_DEPTH (reported as invalid)
    PF_CONTENT = 1u << 13      // at least one non-empty content string was appended
};

#define LGW_MAX_DEPTH 64
#define LGW_STR_CAP 120        // captured model / provider bytes (decoded UTF-8)

struct Val {
    int64_t bits;      // KD_INT: value; KD_FLT: IEEE-754 bits
    uint8_t kind;
};

// Usage as the tap reads it, before get_token_usage's arithmetic.
struct UsageRaw {
    Val prompt, completion, total, cost, reasoning, cached;
    uint8_t usage_kind;        // Kind of the top-level "usage" value
    uint8_t ctd_kind, ptd_kind;  // Kind of usage.completion_tokens_details / prompt_tokens_details
    uint8_t model_kind, provider_kind;
    Val model_val, provider_val;           // when not a string
    uint8_t model_len, provider_len;
    uint8_t model_flags, provider_flags;   // bit0: truncated, bit1: lone surrogate
    char model[LGW_STR_CAP];
    char provider[LGW_STR_CAP];
};

// ---- machine -------------------------------------------------------------------------------
enum St : uint8_t {
    S_VALUE, S_VALUE_OR_END, S_KEY_OR_END, S_KEY, S_COLON, S_AFTER,
    S_STR, S_STR_ESC, S_STR_U,
    S_NUM_MINUS, S_NUM_ZERO, S_NUM_INT, S_NUM_DOT, S_NUM_FRAC, S_NUM_E, S_NUM_ESIGN, S_NUM_EXP,
    S_LIT, S_DONE, S_TRAIL_B, S_ERR
};

enum Ctx : uint8_t {
    X_TOP, X_USAGE, X_CTD, X_PTD, X_CHOICES, X_CHOICE, X_DELTA, X_MESSAGE, X_OTHER
};

// tracked key slots (per context)
enum Slot : uint8_t {
    SL_NONE = 0,
    SL_ERROR, SL_DETAIL, SL_CODE, SL_USAGE, SL_CHOICES, SL_MODEL, SL_PROVIDER,       // X_TOP
    SL_PROMPT, SL_COMPLETION, SL_TOTAL, SL_COST, SL_CTD, SL_PTD,                     // X_USAGE
    SL_REASONING, SL_CACHED,                                                          // X_CTD / X_PTD
    SL_DELTA, SL_MESSAGE,                                                             // X_CHOICE
    SL_CONTENT                                                                        // X_DELTA / X_MESSAGE
};

LGW_HD constexpr uint64_t pk(const char* s, int off, int n) {
    uint64_t v = 0;
    for (int i = 0; i < 8 && off + i < n; ++i) v |= (uint64_t)(uint8_t)s[off + i] << (8 * i);
    return v;
}
#define LGW_KEYEQ(lit) (klen == (int)sizeof(lit) - 1 && k0 == pk(lit, 0, sizeof(lit) - 1) && \
                        k1 == pk(lit, 8, sizeof(lit) - 1) && k2 == pk(lit, 16, sizeof(lit) - 1) && \
                        k3 == pk(lit, 24, sizeof(lit) - 1))

struct ChoiceSide {     // what the tap needs to know about choice["delta"] / choice["message"]
    uint8_t present, kind, has_content, content_kind, content_truthy;
};

// Optional position tracker (template-driven usage extraction, relay2.cuh): where the values of the eight usage
// fields start and end in the text, how often each tracked key occurs, and the extent of the "choices" value.
enum UsageField : uint8_t { UF_PROMPT = 0, UF_COMPLETION, UF_TOTAL, UF_COST, UF_REASONING, UF_CACHED, UF_MODEL, UF_PROVIDER, UF_N };
struct ValueTrack {
    uint32_t pos;                  // position of the byte being fed (set by the caller before feed())
    uint32_t vstart;               // where the value being read began
    uint32_t fstart[UF_N], fend[UF_N];   // value extent per field: [fstart, fend) (numbers: fend = terminator; strings: fend = closing quote + 1)
    uint32_t dup;                  // some tracked key occurred more than once
    uint32_t seen;                 // bit per tracked slot
    uint32_t choices_lo, choices_hi;
    LGW_HD void reset() { pos = 0; vstart = 0; dup = 0; seen = 0; choices_lo = choices_hi = 0xFFFFFFFFu; for (int i = 0; i < UF_N; ++i) { fstart[i] = fend[i] = 0xFFFFFFFFu; } }
};

// TEXT (transcript tap, transcript.cuh): additionally reports WHERE the content strings the choices walk appends lie in the
// text (chat_logging.py:127-133), so that the caller can decode them into llm_response_accum.
template <bool EXTRACT, bool TEXT = false>
struct JsonMachine {
    // syntax
    uint8_t st, depth, ctx, slot;
    uint8_t other_ret, other_depth;
    uint8_t in_key, lit_id, lit_pos, ucount;
    uint8_t strip_mode;        // 1: text came after "data: " (tap strips Python whitespace)
    uint8_t cont_empty;        // the container being closed had no members
    uint64_t stack;            // bit d = 1 when the container at depth d+1 is an object
    uint32_t flags;            // TopKey | PartFlag bits
    // key accumulator (decoded key text, first 32 bytes)
    uint64_t k0, k1, k2, k3;
    int klen; uint8_t key_bad;
    // \uXXXX accumulator
    uint32_t ucode; uint32_t pending_high;
    // number accumulator
    DecAcc num;
    // string value truthiness
    uint32_t slen;
    // choices walk
    ChoiceSide cd, cm;
    uint8_t ch_stop;           // a TypeError / exotic shape already decided the walk
    uint32_t ret_slots;
    // capture target
    UsageRaw* rec;
    char* cap; uint8_t* cap_len; uint8_t* cap_flags;
    ValueTrack* trk;           // optional (nullptr: no tracking)
    // TEXT only: position of the byte being fed (set by the caller), raw extent of the last "content" string of the choice's
    // delta [0] / message [1] (between its quotes), and what the byte just fed decided: `emit_pending` -- the walk appends
    // text[emit_lo, emit_hi) (chat_logging.py:129,133); `rollback` -- a repeated "choices" key voids what an earlier one appended
    uint32_t tpos, c_lo[2], c_hi[2], emit_lo, emit_hi;
    uint8_t emit_pending, rollback;

    // (a container value -- "prompt_tokens":[1] -- has no extent of its own here: vstart is where its LAST inner value began, which
    // would alias that inner value's span; it is recorded as 0xFFFFFFFE = "present, not a scalar" and the template builder
    // refuses to read usage fields from such a template)
    LGW_HD void track_field(int f, uint8_t kind) {
        if (EXTRACT && trk) {
            if (kind == KD_ARR || kind == KD_OBJ) { trk->fstart[f] = trk->fend[f] = 0xFFFFFFFEu; return; }
            trk->fstart[f] = trk->vstart; trk->fend[f] = trk->pos + (kind == KD_STR ? 1u : 0u);
        }
    }

    LGW_HD void reset(UsageRaw* r, bool strip) {
        st = S_VALUE; depth = 0; ctx = X_TOP; slot = SL_NONE; other_ret = X_TOP; other_depth = 0;
        in_key = 0; lit_id = 0; lit_pos = 0; ucount = 0; strip_mode = strip ? 1 : 0; cont_empty = 0;
        stack = 0; flags = 0; k0 = k1 = k2 = k3 = 0; klen = 0; key_bad = 0; ucode = 0; pending_high = 0;
        num.reset(); slen = 0; cd = ChoiceSide{0, 0, 0, 0, 0}; cm = cd; ch_stop = 0; ret_slots = 0;
        rec = r; cap = nullptr; cap_len = nullptr; cap_flags = nullptr; trk = nullptr;
        if (TEXT) { tpos = 0; c_lo[0] = c_lo[1] = c_hi[0] = c_hi[1] = 0; emit_lo = emit_hi = 0; emit_pending = 0; rollback = 0; }
        if (EXTRACT && r) clear_usage(*r);
    }

    static LGW_HD void clear_val(Val& v) { v.bits = 0; v.kind = KD_ABSENT; }
    static LGW_HD void clear_usage_fields(UsageRaw& u) {
        clear_val(u.prompt); clear_val(u.completion); clear_val(u.total); clear_val(u.cost);
        clear_val(u.reasoning); clear_val(u.cached); u.ctd_kind = KD_ABSENT; u.ptd_kind = KD_ABSENT;
    }
    static LGW_HD void clear_usage(UsageRaw& u) {
        clear_usage_fields(u); u.usage_kind = KD_ABSENT; u.model_kind = KD_ABSENT; u.provider_kind = KD_ABSENT;
        clear_val(u.model_val); clear_val(u.provider_val);
        u.model_len = u.provider_len = 0; u.model_flags = u.provider_flags = 0;
    }

    LGW_HD bool failed() const { return st == S_ERR; }

    // ---- helpers -----------------------------------------------------------------------
    static LGW_HD bool is_ws(uint32_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
    // Python str.strip() whitespace that is ASCII (the multi-byte ones are handled as invalid;
    // see DESIGN.md "known gaps")
    static LGW_HD bool is_py_ws(uint32_t c) { return c == ' ' || (c >= 0x09 && c <= 0x0d) || (c >= 0x1c && c <= 0x1f); }
    static LGW_HD bool is_digit(uint32_t c) { return c - '0' < 10u; }

    LGW_HD bool top_is_obj() const { return (stack >> (depth - 1)) & 1ull; }

    LGW_HD void fail() { st = S_ERR; }

    // a complete value of kind `k` (scalar, string or just-closed container) arrived for `slot`
    LGW_HD void on_value(uint8_t kind, int64_t bits, bool truthy) {
        const uint8_t s = slot;
        slot = SL_NONE;
        switch (ctx) {
        case X_TOP:
            if (EXTRACT && rec) {
                if (s == SL_USAGE) rec->usage_kind = kind;
                else if (s == SL_MODEL) { rec->model_kind = kind; rec->model_val.kind = kind; rec->model_val.bits = bits; track_field(UF_MODEL, kind); }
                else if (s == SL_PROVIDER) { rec->provider_kind = kind; rec->provider_val.kind = kind; rec->provider_val.bits = bits; track_field(UF_PROVIDER, kind); }
            }
            if (EXTRACT && trk && s == SL_CHOICES) trk->choices_hi = trk->pos;
            if (s == SL_CHOICES) {
                // chat_logging.py:125 `for choice in chunk_json["choices"]`
                if (kind == KD_OBJ) { if (!cont_empty) flags |= PF_EXOTIC; }        // iterates keys
                else if (kind == KD_STR || kind == KD_ARR) {}                        // chars / handled per element
                else flags |= PF_TYPE_ERROR;                                        // not iterable
            }
            break;
        case X_USAGE:
            if (EXTRACT && rec) {
                Val v; v.kind = kind; v.bits = bits;
                if (s == SL_PROMPT) { rec->prompt = v; track_field(UF_PROMPT, kind); }
                else if (s == SL_COMPLETION) { rec->completion = v; track_field(UF_COMPLETION, kind); }
                else if (s == SL_TOTAL) { rec->total = v; track_field(UF_TOTAL, kind); }
                else if (s == SL_COST) { rec->cost = v; track_field(UF_COST, kind); }
                else if (s == SL_CTD) rec->ctd_kind = kind;
                else if (s == SL_PTD) rec->ptd_kind = kind;
            }
            break;
        case X_CTD:
            if (EXTRACT && rec && s == SL_REASONING) { rec->reasoning.kind = kind; rec->reasoning.bits = bits; track_field(UF_REASONING, kind); }
            break;
        case X_PTD:
            if (EXTRACT && rec && s == SL_CACHED) { rec->cached.kind = kind; rec->cached.bits = bits; track_field(UF_CACHED, kind); }
            break;
        case X_CHOICES:
            // a scalar / string / nested array element of the choices list (objects are
            // evaluated when they close, see close_choice)
            if (!ch_stop && kind != KD_OBJ) {
                if (kind == KD_STR) { if (slen) { flags |= PF_EXOTIC; ch_stop = 1; } }   // substring tests
                else if (kind == KD_ARR) { if (!cont_empty) { flags |= PF_EXOTIC; ch_stop = 1; } }
                else { flags |= PF_TYPE_ERROR; ch_stop = 1; }                            // `"delta" in 5`
            }
            break;
        case X_CHOICE: {
            ChoiceSide* side = (s == SL_DELTA) ? &cd : (s == SL_MESSAGE) ? &cm : nullptr;
            if (side) { side->present = 1; side->kind = kind; }
            break; }
        case X_DELTA: case X_MESSAGE:
            if (s == SL_CONTENT) {
                ChoiceSide* side = (ctx == X_DELTA) ? &cd : &cm;
                side->has_content = 1; side->content_kind = kind; side->content_truthy = truthy ? 1 : 0;
            }
            break;
        default: break;
        }
        st = (depth == 0) ? S_DONE : S_AFTER;
    }

    // chat_logging.py:126-133 for one choice object, once all its keys are known
    LGW_HD void close_choice() {
        if (ch_stop) return;
        for (int pass = 0; pass < 2; ++pass) {
            const ChoiceSide& sd = pass == 0 ? cd : cm;
            if (!sd.present) continue;
            bool test;
            if (sd.kind == KD_OBJ) test = sd.has_content;
            else if (sd.kind == KD_STR || sd.kind == KD_ARR) { flags |= PF_EXOTIC; ch_stop = 1; return; }
            else { flags |= PF_TYPE_ERROR; ch_stop = 1; return; }              // `"content" in None`
            if (!test) continue;
            if (sd.content_kind == KD_STR) {
                if (sd.content_truthy) { flags |= PF_CONTENT; if (TEXT) { emit_lo = c_lo[pass]; emit_hi = c_hi[pass]; emit_pending = 1; } }
            }
            else if (sd.content_truthy) { flags |= PF_TYPE_ERROR; ch_stop = 1; } // str += non-str
            return;                                                              // if / elif
        }
    }

    // The slot a container was opened for must be known again when it closes.  Pushes happen
    // only from tracked contexts (the root, usage, *_details, choices, choice, delta/message and
    // the one untracked container below them), so six 5-bit fields in one register suffice.
    LGW_HD void ret_slot_push(uint8_t s) { ret_slots = (ret_slots << 5) | s; }
    LGW_HD uint8_t ret_slot_pop() { uint8_t s = ret_slots & 31u; ret_slots >>= 5; return s; }

    LGW_HD void open_container(bool is_obj) {
        if (depth >= LGW_MAX_DEPTH) { flags |= PF_TOO_DEEP; fail(); return; }
        if (ctx != X_OTHER) {
            const uint8_t s = slot;
            uint8_t nctx = X_OTHER;
            if (depth == 0) nctx = X_TOP;                                   // the event's own object
            else if (ctx == X_TOP && s == SL_USAGE && is_obj) { nctx = X_USAGE; if (EXTRACT && rec) clear_usage_fields(*rec); }
            else if (ctx == X_USAGE && s == SL_CTD && is_obj) { nctx = X_CTD; if (EXTRACT && rec) clear_val(rec->reasoning); }
            else if (ctx == X_USAGE && s == SL_PTD && is_obj) { nctx = X_PTD; if (EXTRACT && rec) clear_val(rec->cached); }
            else if (ctx == X_TOP && s == SL_CHOICES && !is_obj) nctx = X_CHOICES;
            else if (ctx == X_CHOICES && is_obj) { nctx = X_CHOICE; cd = ChoiceSide{0, 0, 0, 0, 0}; cm = cd; }
            else if (ctx == X_CHOICE && s == SL_DELTA && is_obj) nctx = X_DELTA;
            else if (ctx == X_CHOICE && s == SL_MESSAGE && is_obj) nctx = X_MESSAGE;
            ret_slot_push(s);
            if (nctx == X_OTHER) { other_ret = ctx; other_depth = depth + 1; }
            ctx = nctx;
        }
        if (is_obj) stack |= (1ull << depth); else stack &= ~(1ull << depth);
        ++depth;
        slot = SL_NONE;
        st = is_obj ? S_KEY_OR_END : S_VALUE_OR_END;
    }

    LGW_HD void close_container(bool is_obj, bool empty) {
        if (depth == 0 || top_is_obj() != is_obj) { fail(); return; }
        --depth;
        if (ctx == X_OTHER) {
            if (depth >= other_depth) { st = S_AFTER; return; }      // still inside the untracked subtree
            ctx = other_ret;
        } else {
            const uint8_t closing = ctx;
            if (closing == X_CHOICE) close_choice();
            ctx = (closing == X_USAGE || closing == X_CHOICES) ? X_TOP
                : (closing == X_CTD || closing == X_PTD) ? X_USAGE
                : (closing == X_CHOICE) ? X_CHOICES
                : (closing == X_DELTA || closing == X_MESSAGE) ? X_CHOICE : X_TOP;
        }
        slot = ret_slot_pop();
        cont_empty = empty ? 1 : 0;
        on_value(is_obj ? KD_OBJ : KD_ARR, 0, !empty);
    }

    // ---- key end ------------------------------------------------------------------------
    LGW_HD void end_key() {
        uint8_t s = SL_NONE;
        if (!key_bad) {
            switch (ctx) {
            case X_TOP:
                if (LGW_KEYEQ("error")) { s = SL_ERROR; flags |= TK_ERROR; }
                else if (LGW_KEYEQ("detail")) { s = SL_DETAIL; flags |= TK_DETAIL; }
                else if (LGW_KEYEQ("code")) { s = SL_CODE; flags |= TK_CODE; }
                else if (LGW_KEYEQ("usage")) { s = SL_USAGE; flags |= TK_USAGE; if (EXTRACT && rec) { clear_usage_fields(*rec); rec->usage_kind = KD_ABSENT; } }
                else if (LGW_KEYEQ("choices")) {
                    s = SL_CHOICES; flags |= TK_CHOICES;
                    flags &= ~(uint32_t)(PF_TYPE_ERROR | PF_EXOTIC | PF_CONTENT); ch_stop = 0;   // last duplicate wins
                    if (TEXT) rollback = 1;
                }
                else if (LGW_KEYEQ("model")) { s = SL_MODEL; flags |= TK_MODEL; }
                else if (LGW_KEYEQ("provider")) { s = SL_PROVIDER; flags |= TK_PROVIDER; }
                break;
            case X_USAGE:
                if (LGW_KEYEQ("prompt_tokens")) s = SL_PROMPT;
                else if (LGW_KEYEQ("completion_tokens")) s = SL_COMPLETION;
                else if (LGW_KEYEQ("total_tokens")) s = SL_TOTAL;
                else if (LGW_KEYEQ("cost")) s = SL_COST;
                else if (LGW_KEYEQ("completion_tokens_details")) { s = SL_CTD; if (EXTRACT && rec) clear_val(rec->reasoning); }
                else if (LGW_KEYEQ("prompt_tokens_details")) { s = SL_PTD; if (EXTRACT && rec) clear_val(rec->cached); }
                break;
            case X_CTD: if (LGW_KEYEQ("reasoning_tokens")) s = SL_REASONING; break;
            case X_PTD: if (LGW_KEYEQ("cached_tokens")) s = SL_CACHED; break;
            case X_CHOICE:
                if (LGW_KEYEQ("delta")) { s = SL_DELTA; cd = ChoiceSide{0, 0, 0, 0, 0}; }
                else if (LGW_KEYEQ("message")) { s = SL_MESSAGE; cm = ChoiceSide{0, 0, 0, 0, 0}; }
                break;
            case X_DELTA: case X_MESSAGE: if (LGW_KEYEQ("content")) s = SL_CONTENT; break;
            default: break;
            }
        }
        if (EXTRACT && trk && s != SL_NONE) {
            // one bit per (context, key): the same key name in two different contexts is not a duplicate, but the
            // second "usage" object of a document re-enters X_USAGE -- its keys then count as repeats (conservative)
            const uint32_t bit = 1u << s;
            if (trk->seen & bit) trk->dup = 1;
            trk->seen |= bit;
            if (s == SL_CHOICES) trk->choices_lo = trk->pos;
        }
        slot = s;
        st = S_COLON;
    }

    // ---- decoded string character ---------------------------------------------------------
    LGW_HD void cap_byte(uint32_t b) {
        if (*cap_len < LGW_STR_CAP) cap[(*cap_len)++] = (char)b; else *cap_flags |= 1;
    }
    LGW_HD void cap_cp(uint32_t cp) {       // append a code point as UTF-8
        if (cp < 0x80) cap_byte(cp);
        else if (cp < 0x800) { cap_byte(0xC0 | (cp >> 6)); cap_byte(0x80 | (cp & 63)); }
        else if (cp < 0x10000) { cap_byte(0xE0 | (cp >> 12)); cap_byte(0x80 | ((cp >> 6) & 63)); cap_byte(0x80 | (cp & 63)); }
        else { cap_byte(0xF0 | (cp >> 18)); cap_byte(0x80 | ((cp >> 12) & 63)); cap_byte(0x80 | ((cp >> 6) & 63)); cap_byte(0x80 | (cp & 63)); }
    }
    LGW_HD void flush_high() {
        if (pending_high) { if (EXTRACT && cap) *cap_flags |= 2; pending_high = 0; }   // lone surrogate
    }
    // `raw` = a literal byte of the string (may be a UTF-8 continuation); `cp` = an escape's value
    LGW_HD void str_raw(uint32_t b) {
        ++slen;
        if (in_key) {
            if (b >= 0x80) key_bad = 1;
            else if (klen < 32) {
                const uint64_t v = (uint64_t)b << ((klen & 7) * 8);
                const int w = klen >> 3;
                if (w == 0) k0 |= v; else if (w == 1) k1 |= v; else if (w == 2) k2 |= v; else k3 |= v;
            }
            ++klen;
        } else if (EXTRACT && cap) { flush_high(); cap_byte(b); }
    }
    LGW_HD void str_escape_cp(uint32_t cp) {
        if (cp < 0x80) { str_raw(cp); return; }
        ++slen;
        if (in_key) { key_bad = 1; ++klen; return; }
        if (EXTRACT && cap) {
            if (cp >= 0xD800 && cp <= 0xDBFF) { flush_high(); pending_high = cp; }
            else if (cp >= 0xDC00 && cp <= 0xDFFF) {
                if (pending_high) { cap_cp(0x10000 + ((pending_high - 0xD800) << 10) + (cp - 0xDC00)); pending_high = 0; }
                else *cap_flags |= 2;
            } else { flush_high(); cap_cp(cp); }
        }
    }

    LGW_HD void begin_string(bool key) {
        in_key = key ? 1 : 0; st = S_STR; slen = 0; pending_high = 0;
        if (key) { k0 = k1 = k2 = k3 = 0; klen = 0; key_bad = 0; }
        else if (EXTRACT && rec) {
            cap = nullptr;
            if (ctx == X_TOP && slot == SL_MODEL) { cap = rec->model; cap_len = &rec->model_len; cap_flags = &rec->model_flags; }
            else if (ctx == X_TOP && slot == SL_PROVIDER) { cap = rec->provider; cap_len = &rec->provider_len; cap_flags = &rec->provider_flags; }
            if (cap) { *cap_len = 0; *cap_flags = 0; }
        }
        if (TEXT && !key && slot == SL_CONTENT && (ctx == X_DELTA || ctx == X_MESSAGE)) c_lo[ctx == X_MESSAGE] = tpos + 1;
    }
    LGW_HD void end_string() {
        if (in_key) { end_key(); return; }
        if (EXTRACT && cap) { flush_high(); cap = nullptr; }
        if (TEXT && slot == SL_CONTENT && (ctx == X_DELTA || ctx == X_MESSAGE)) c_hi[ctx == X_MESSAGE] = tpos;
        on_value(KD_STR, 0, slen != 0);
    }

    LGW_HD void begin_value(uint32_t c) {
        if (EXTRACT && trk) trk->vstart = trk->pos;
        if (c == '"') { begin_string(false); }
        else if (c == '{') open_container(true);
        else if (c == '[') open_container(false);
        else if (c == '-') { num.reset(); num.neg = 1; st = S_NUM_MINUS; }
        else if (c == '0') { num.reset(); st = S_NUM_ZERO; }
        else if (c - '1' < 9u) { num.reset(); num.digit(c - '0', false); st = S_NUM_INT; }
        else if (c == 't') { st = S_LIT; lit_id = 0; lit_pos = 1; }
        else if (c == 'f') { st = S_LIT; lit_id = 1; lit_pos = 1; }
        else if (c == 'n') { st = S_LIT; lit_id = 2; lit_pos = 1; }
        else if (c == 'N') { st = S_LIT; lit_id = 3; lit_pos = 1; }
        else if (c == 'I') { st = S_LIT; lit_id = 4; lit_pos = 1; num.neg = 0; }
        else fail();
    }

    LGW_HD void end_number() {
        uint8_t kind; int64_t bits; bool truthy;
        // the value is needed when extracting, and for a float "content" (its truthiness after
        // rounding decides between skip and TypeError, chat_logging.py:128)
        if (EXTRACT || (num.is_float && slot == SL_CONTENT)) num.finish(kind, bits, truthy);
        else { kind = num.is_float ? KD_FLT : KD_INT; bits = 0; truthy = num.nonzero; }
        on_value(kind, bits, truthy);
    }

    static LGW_HD const char* lit_text(int id) {
        return id == 0 ? "true" : id == 1 ? "false" : id == 2 ? "null" : id == 3 ? "NaN" : "Infinity";
    }

    // ---- one byte -------------------------------------------------------------------------
    LGW_HD void feed(uint32_t c) {
        for (;;) {
            switch (st) {
            case S_STR:
                if (c == '"') end_string();
                else if (c == '\\') st = S_STR_ESC;
                else if (c < 0x20) fail();
                else str_raw(c);
                return;
            case S_STR_ESC:
                st = S_STR;
                switch (c) {
                case '"': str_raw('"'); break;  case '\\': str_raw('\\'); break;  case '/': str_raw('/'); break;
                case 'b': str_raw(8); break;    case 'f': str_raw(12); break;     case 'n': str_raw(10); break;
                case 'r': str_raw(13); break;   case 't': str_raw(9); break;
                case 'u': st = S_STR_U; ucount = 0; ucode = 0; break;
                default: fail();
                }
                return;
            case S_STR_U: {
                uint32_t d;
                if (c - '0' < 10u) d = c - '0';
                else if ((c | 0x20) - 'a' < 6u) d = (c | 0x20) - 'a' + 10;
                else { fail(); return; }
                ucode = (ucode << 4) | d;
                if (++ucount == 4) { st = S_STR; str_escape_cp(ucode); }
                return; }
            case S_VALUE:
                if (is_ws(c)) return;
                begin_value(c); return;
            case S_VALUE_OR_END:
                if (is_ws(c)) return;
                if (c == ']') { close_container(false, true); return; }
                begin_value(c); return;
            case S_KEY_OR_END:
                if (is_ws(c)) return;
                if (c == '}') { close_container(true, true); return; }
                if (c == '"') { begin_string(true); return; }
                fail(); return;
            case S_KEY:
                if (is_ws(c)) return;
                if (c == '"') { begin_string(true); return; }
                fail(); return;
            case S_COLON:
                if (is_ws(c)) return;
                if (c == ':') { st = S_VALUE; return; }
                fail(); return;
            case S_AFTER:
                if (is_ws(c)) return;
                if (c == ',') { st = top_is_obj() ? S_KEY : S_VALUE; return; }
                if (c == '}') { close_container(true, false); return; }
                if (c == ']') { close_container(false, false); return; }
                fail(); return;
            case S_NUM_MINUS:
                if (c == '0') { st = S_NUM_ZERO; return; }
                if (c - '1' < 9u) { num.digit(c - '0', false); st = S_NUM_INT; return; }
                if (c == 'I') { st = S_LIT; lit_id = 4; lit_pos = 1; return; }     // -Infinity
                fail(); return;
            case S_NUM_ZERO:
                if (c == '.') { st = S_NUM_DOT; return; }
                if ((c | 0x20) == 'e') { st = S_NUM_E; return; }
                end_number(); continue;
            case S_NUM_INT:
                if (is_digit(c)) { num.digit(c - '0', false); return; }
                if (c == '.') { st = S_NUM_DOT; return; }
                if ((c | 0x20) == 'e') { st = S_NUM_E; return; }
                end_number(); continue;
            case S_NUM_DOT:
                if (is_digit(c)) { num.is_float = 1; num.digit(c - '0', true); st = S_NUM_FRAC; return; }
                fail(); return;
            case S_NUM_FRAC:
                if (is_digit(c)) { num.digit(c - '0', true); return; }
                if ((c | 0x20) == 'e') { st = S_NUM_E; return; }
                end_number(); continue;
            case S_NUM_E:
                num.is_float = 1;
                if (c == '+') { st = S_NUM_ESIGN; return; }
                if (c == '-') { num.exp_neg = 1; st = S_NUM_ESIGN; return; }
                if (is_digit(c)) { num.exp_digit(c - '0'); st = S_NUM_EXP; return; }
                fail(); return;
            case S_NUM_ESIGN:
                if (is_digit(c)) { num.exp_digit(c - '0'); st = S_NUM_EXP; return; }
                fail(); return;
            case S_NUM_EXP:
                if (is_digit(c)) { num.exp_digit(c - '0'); return; }
                end_number(); continue;
            case S_LIT: {
                const char* t = lit_text(lit_id);
                if ((uint32_t)(uint8_t)t[lit_pos] != c) { fail(); return; }
                ++lit_pos;
                if (t[lit_pos] == 0) {
                    if (lit_id == 0) on_value(KD_TRUE, 0, true);
                    else if (lit_id == 1) on_value(KD_FALSE, 0, false);
                    else if (lit_id == 2) on_value(KD_NULL, 0, false);
                    else if (lit_id == 3) on_value(KD_FLT, 0x7ff8000000000000ll, true);
                    else on_value(KD_FLT, num.neg ? (int64_t)0xfff0000000000000ull : 0x7ff0000000000000ll, true);
                }
                return; }
            case S_DONE:
                if (is_ws(c)) return;
                if (strip_mode && is_py_ws(c)) { st = S_TRAIL_B; return; }   // A: extra data, B: stripped
                fail(); return;
            case S_TRAIL_B:
                if (is_py_ws(c)) return;
                fail(); return;
            default: return;     // S_ERR absorbs
            }
        }
    }

    // end of the event text: returns TopKey | PartFlag bits
    LGW_HD uint32_t finish() {
        // a number can only end the text at depth 0, which a '{'-rooted document never reaches
        uint32_t f = flags;
        if (st == S_DONE) f |= PF_VALID_A | PF_VALID_B;
        else if (st == S_TRAIL_B) f |= PF_VALID_B;
        return f;
    }
};

}  // namespace lgw
