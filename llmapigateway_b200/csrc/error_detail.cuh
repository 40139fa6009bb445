// Error detail of a failing non-streaming upstream response (request_handler.py:167-169):
//     error_detail = response_json.get("error", {}).get("message") or response_json.get("detail")
// for a document that is KNOWN to be valid JSON with an object root (the response plan's parse said so).  A purpose-built walk
// over the top level (string- and depth-aware skipping of values, keys compared after unescaping): which of the two values wins
// by Python's rules, and its text when it is a string.  Values whose str() the device does not reproduce (numbers other than
// zero, non-empty containers) are reported as exotic, never guessed.  Host/device code: the CPU suite fuzzes it against CPython.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LGW_ED_HD __host__ __device__
#else
#define LGW_ED_HD
#endif

namespace lgw {

enum EdKind : uint8_t { ED_ABSENT = 0, ED_STR, ED_NULL, ED_TRUE, ED_FALSE, ED_NUM_ZERO, ED_NUM, ED_OBJ_EMPTY, ED_OBJ, ED_ARR_EMPTY, ED_ARR };
enum EdResult : uint8_t {
    EDR_NONE = 0,          // error_detail is None
    EDR_TEXT = 1,          // a str: `text[0..text_len)` (UTF-8, unescaped)
    EDR_TRUE = 2, EDR_FALSE = 3,                       // the bools themselves
    EDR_ERROR_NOT_OBJECT = 4,                          // `.get` on a non-dict "error" value raises: AttributeError text from error_kind
    EDR_EXOTIC = 5,        // a value whose str() is not modelled (number, container), a lone surrogate, or text longer than the buffer
    EDR_NUMBER = 6         // a number: `text` holds its literal as spelt in the document (the caller makes the int / float of it)
};

struct DocError {                 // == lgw_doc_error
    uint8_t result;               // EdResult
    uint8_t error_kind, message_kind, detail_kind;     // EdKind
    uint32_t text_len;
};

LGW_ED_HD inline bool ed_ws(uint32_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
LGW_ED_HD inline uint32_t ed_skip_ws(const uint8_t* p, uint32_t i, uint32_t n) { while (i < n && ed_ws(p[i])) ++i; return i; }
LGW_ED_HD inline int ed_hex(uint32_t c) { return c - '0' < 10u ? (int)(c - '0') : ((c | 0x20u) - 'a' < 6u ? (int)((c | 0x20u) - 'a' + 10) : -1); }

// string starting at the opening quote p[i]: returns the index behind the closing quote
LGW_ED_HD inline uint32_t ed_skip_string(const uint8_t* p, uint32_t i, uint32_t n) {
    ++i;
    while (i < n && p[i] != '"') i += p[i] == '\\' ? 2u : 1u;
    return i < n ? i + 1u : n;
}

// value starting at p[i] (no leading whitespace): its kind, returns the index behind it
LGW_ED_HD inline uint32_t ed_skip_value(const uint8_t* p, uint32_t i, uint32_t n, uint8_t& kind) {
    const uint32_t c = i < n ? p[i] : 0u;
    if (c == '"') { kind = ED_STR; return ed_skip_string(p, i, n); }
    if (c == '{' || c == '[') {
        const uint32_t j = ed_skip_ws(p, i + 1, n);
        const bool empty = j < n && p[j] == (c == '{' ? '}' : ']');
        kind = c == '{' ? (empty ? ED_OBJ_EMPTY : ED_OBJ) : (empty ? ED_ARR_EMPTY : ED_ARR);
        uint32_t depth = 0;
        while (i < n) {
            const uint32_t ch = p[i];
            if (ch == '"') { i = ed_skip_string(p, i, n); continue; }
            if (ch == '{' || ch == '[') ++depth;
            else if (ch == '}' || ch == ']') { if (--depth == 0) return i + 1u; }
            ++i;
        }
        return n;
    }
    uint32_t j = i;
    while (j < n && !ed_ws(p[j]) && p[j] != ',' && p[j] != '}' && p[j] != ']') ++j;
    if (c == 'n') kind = ED_NULL;
    else if (c == 't') kind = ED_TRUE;
    else if (c == 'f') kind = ED_FALSE;
    else {                                                   // a number: zero when every digit of its mantissa is 0 (0, -0, 0.0, 0e5 ...)
        bool zero = true;
        for (uint32_t k = i; k < j; ++k) { const uint32_t d = p[k]; if (d == 'e' || d == 'E') break; if (d - '1' < 9u) zero = false; }
        kind = zero ? ED_NUM_ZERO : ED_NUM;
    }
    return j;
}

// unescape the JSON string whose opening quote is p[i] into dst[0..cap): returns the length, 0xFFFFFFFF for a lone surrogate or overflow
LGW_ED_HD inline uint32_t ed_unescape(const uint8_t* p, uint32_t i, uint32_t n, uint8_t* dst, uint32_t cap) {
    uint32_t o = 0;
    ++i;
    while (i < n && p[i] != '"') {
        uint32_t cp = p[i];
        if (cp != '\\') { if (o >= cap) return 0xFFFFFFFFu; dst[o++] = (uint8_t)cp; ++i; continue; }
        const uint32_t e = i + 1 < n ? p[i + 1] : 0u;
        i += 2;
        if (e == 'u') {
            if (i + 4 > n) return 0xFFFFFFFFu;
            cp = (uint32_t)((ed_hex(p[i]) << 12) | (ed_hex(p[i + 1]) << 8) | (ed_hex(p[i + 2]) << 4) | ed_hex(p[i + 3]));
            i += 4;
            if (cp >= 0xD800u && cp < 0xDC00u) {             // high surrogate: needs its low half right behind
                if (i + 6 <= n && p[i] == '\\' && p[i + 1] == 'u') {
                    const uint32_t lo = (uint32_t)((ed_hex(p[i + 2]) << 12) | (ed_hex(p[i + 3]) << 8) | (ed_hex(p[i + 4]) << 4) | ed_hex(p[i + 5]));
                    if (lo >= 0xDC00u && lo < 0xE000u) { cp = 0x10000u + ((cp - 0xD800u) << 10) + (lo - 0xDC00u); i += 6; }
                    else return 0xFFFFFFFFu;
                } else return 0xFFFFFFFFu;
            } else if (cp >= 0xDC00u && cp < 0xE000u) return 0xFFFFFFFFu;
        } else if (e == 'n') cp = '\n'; else if (e == 't') cp = '\t'; else if (e == 'r') cp = '\r'; else if (e == 'b') cp = '\b'; else if (e == 'f') cp = '\f';
        else cp = e;                                         // \" \\ \/
        if (cp < 0x80u) { if (o + 1 > cap) return 0xFFFFFFFFu; dst[o++] = (uint8_t)cp; }
        else if (cp < 0x800u) { if (o + 2 > cap) return 0xFFFFFFFFu; dst[o++] = (uint8_t)(0xC0u | (cp >> 6)); dst[o++] = (uint8_t)(0x80u | (cp & 0x3Fu)); }
        else if (cp < 0x10000u) { if (o + 3 > cap) return 0xFFFFFFFFu; dst[o++] = (uint8_t)(0xE0u | (cp >> 12)); dst[o++] = (uint8_t)(0x80u | ((cp >> 6) & 0x3Fu)); dst[o++] = (uint8_t)(0x80u | (cp & 0x3Fu)); }
        else { if (o + 4 > cap) return 0xFFFFFFFFu; dst[o++] = (uint8_t)(0xF0u | (cp >> 18)); dst[o++] = (uint8_t)(0x80u | ((cp >> 12) & 0x3Fu)); dst[o++] = (uint8_t)(0x80u | ((cp >> 6) & 0x3Fu)); dst[o++] = (uint8_t)(0x80u | (cp & 0x3Fu)); }
    }
    return o;
}

// the members of the object whose '{' is p[i]: position (of the value's first byte) and kind of the LAST member named k1 / k2
// (a dict keeps the last of duplicate keys); returns the index behind the closing '}'
LGW_ED_HD inline uint32_t ed_walk_object(const uint8_t* p, uint32_t i, uint32_t n, const char* k1, uint32_t l1, uint32_t& pos1, uint8_t& kind1,
                                         const char* k2, uint32_t l2, uint32_t& pos2, uint8_t& kind2) {
    i = ed_skip_ws(p, i + 1, n);
    while (i < n && p[i] != '}') {
        if (p[i] == ',') { i = ed_skip_ws(p, i + 1, n); continue; }
        uint8_t key[16];
        const uint32_t kl = ed_unescape(p, i, n, key, 16);   // (longer than 16 bytes: neither of the names)
        i = ed_skip_ws(p, ed_skip_string(p, i, n), n);
        if (i < n && p[i] == ':') i = ed_skip_ws(p, i + 1, n);
        uint8_t kind = ED_ABSENT;
        const uint32_t vs = i;
        i = ed_skip_ws(p, ed_skip_value(p, i, n, kind), n);
        bool m1 = kl == l1, m2 = k2 && kl == l2;
        for (uint32_t q = 0; q < kl && q < 16; ++q) { if (m1 && key[q] != (uint8_t)k1[q]) m1 = false; if (m2 && key[q] != (uint8_t)k2[q]) m2 = false; }
        if (m1) { pos1 = vs; kind1 = kind; }
        if (m2) { pos2 = vs; kind2 = kind; }
    }
    return i < n ? i + 1u : n;
}

LGW_ED_HD inline bool ed_falsy(uint8_t k) { return k == ED_ABSENT || k == ED_NULL || k == ED_FALSE || k == ED_NUM_ZERO || k == ED_OBJ_EMPTY || k == ED_ARR_EMPTY; }

LGW_ED_HD inline void error_detail_of(const uint8_t* p, uint32_t n, DocError& o, uint8_t* text, uint32_t cap) {
    o.result = EDR_NONE; o.error_kind = o.message_kind = o.detail_kind = ED_ABSENT; o.text_len = 0;
    uint32_t i = ed_skip_ws(p, 0, n);
    if (i >= n || p[i] != '{') { o.result = EDR_EXOTIC; return; }
    uint32_t epos = 0, dpos = 0, mpos = 0, unused = 0; uint8_t unused_k = ED_ABSENT;
    ed_walk_object(p, i, n, "error", 5, epos, o.error_kind, "detail", 6, dpos, o.detail_kind);
    if (o.error_kind != ED_ABSENT) {
        if (o.error_kind != ED_OBJ && o.error_kind != ED_OBJ_EMPTY) {
            o.result = EDR_ERROR_NOT_OBJECT;
            if (o.error_kind == ED_NUM || o.error_kind == ED_NUM_ZERO) {          // int or float: the literal tells
                uint8_t k; const uint32_t end = ed_skip_value(p, epos, n, k);
                for (uint32_t q = epos; q < end && o.text_len < cap; ++q) text[o.text_len++] = p[q];
            }
            return;
        }
        ed_walk_object(p, epos, n, "message", 7, mpos, o.message_kind, nullptr, 0, unused, unused_k);
    }
    uint32_t pos = 0; uint8_t kind;
    bool message_wins = false;
    if (o.message_kind == ED_STR) { uint8_t probe[1]; message_wins = ed_unescape(p, mpos, n, probe, 1) != 0u; }    // (a non-empty string is truthy)
    else message_wins = !ed_falsy(o.message_kind);
    if (message_wins) { pos = mpos; kind = o.message_kind; } else { pos = dpos; kind = o.detail_kind; }
    if (kind == ED_ABSENT || kind == ED_NULL) { o.result = EDR_NONE; return; }
    if (kind == ED_TRUE) { o.result = EDR_TRUE; return; }
    if (kind == ED_FALSE) { o.result = EDR_FALSE; return; }
    if (kind == ED_NUM_ZERO || kind == ED_NUM) {
        uint8_t k; const uint32_t end = ed_skip_value(p, pos, n, k);
        if (end - pos > cap) { o.result = EDR_EXOTIC; return; }
        for (uint32_t q = pos; q < end; ++q) text[o.text_len++] = p[q];
        o.result = EDR_NUMBER; return;
    }
    if (kind != ED_STR) { o.result = EDR_EXOTIC; return; }
    const uint32_t len = ed_unescape(p, pos, n, text, cap);
    if (len == 0xFFFFFFFFu) { o.result = EDR_EXOTIC; return; }
    o.result = EDR_TEXT; o.text_len = len;
}

}  // namespace lgw
