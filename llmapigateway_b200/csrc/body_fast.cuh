// Data-parallel request-body rewrite: one thread block per body.
//
// The sequential machine of body_machine.cuh is exact but walks a body byte by byte.  A chat body is
// ~90 % string content, and every output byte is a function of (input byte, a few neighbours, the
// token it belongs to, the render mode) -- so the block does the work in bulk-synchronous phases:
//
//   load      body -> shared memory
//   utf8      whole-body UTF-8 validity (chat.py:32)
//   quotes    per-thread chunk: parity of unescaped quotes        -> block prefix scan -> in-string state
//   tokens    per-thread chunk: structural characters, string and scalar starts -> block prefix scan
//             -> token table in shared memory (position, type, brackets-before)
//   brackets  thread 0: bracket stack over the (few) bracket tokens -> container type / depth after each
//   check     per token: grammar (which token may follow which, in which container), scalar grammar,
//             key limits, duplicate-key suspicion (hash table), top-level key <-> plan matching
//   members   per token: the top-level member it belongs to (block max-scan) -> replaced values skipped
//   size      per-thread chunk: output bytes                        -> block prefix scan -> offsets
//   write     per-thread chunk: output bytes
//
// Contract: the fast path either produces EXACTLY what rewrite_body() produces with status BS_OK /
// BS_OVERFLOW, or it reports "irregular" and the caller runs rewrite_body() (invalid JSON, escapes in
// keys, duplicate-key suspicion, NaN under allow_nan=False, limits, anything unusual).  The CPU suite
// checks that contract on the host build of this very code (phases run thread by thread).
#pragma once
#include <string.h>
#include "body_machine.cuh"

namespace lgw {

#ifndef LGW_FAST_THREADS
#define LGW_FAST_THREADS 384
#endif
#define LGW_FAST_MAXB 6144
#define LGW_FAST_MAXT 1536
#define LGW_FAST_MAXBR 512
#define LGW_FAST_HT 1024
#define LGW_FAST_IRREGULAR 0xFFFFFFFFu

enum FastTok : uint8_t { TK_OBJ_OPEN = 0, TK_OBJ_CLOSE = 1, TK_ARR_OPEN = 2, TK_ARR_CLOSE = 3, TK_COMMA = 4, TK_COLON = 5, TK_STR = 6, TK_SCALAR = 7 };
enum FastFlag : uint8_t { TF_KEY = 1, TF_UNQUOTED = 2, TF_SKIP = 4, TF_REPLACE = 8, TF_FLOAT = 16, TF_NEGZERO = 32, TF_TOPKEY = 64 };

struct alignas(16) FastShared {
    uint8_t text[LGW_FAST_MAXB + 96];
    uint16_t tok_pos[LGW_FAST_MAXT], tok_end[LGW_FAST_MAXT], tok_br[LGW_FAST_MAXT];
    uint8_t tok_type[LGW_FAST_MAXT], tok_flags[LGW_FAST_MAXT], tok_depth[LGW_FAST_MAXT], tok_ctx[LGW_FAST_MAXT], tok_aux[LGW_FAST_MAXT];
    uint16_t br_tok[LGW_FAST_MAXBR], br_open_after[LGW_FAST_MAXBR];
    uint8_t br_ctx_after[LGW_FAST_MAXBR], br_depth_after[LGW_FAST_MAXBR];
    uint32_t br_nkeys[LGW_FAST_MAXBR];
    uint32_t ht[LGW_FAST_HT];
    uint32_t scan[LGW_FAST_THREADS + 1];
    uint16_t th_tok0[LGW_FAST_THREADS], th_br0[LGW_FAST_THREADS];
    uint8_t th_instr[LGW_FAST_THREADS];
    uint32_t n_tok, n_br, irregular, matched, top_members, out_len, n_keys;
};

struct FastCtx {
    FastShared* s;
    const uint8_t* in; uint32_t n;
    int mode; const BodyOp* ops; uint32_t n_ops; const uint8_t* blob;
    uint8_t* out; uint32_t cap; uint32_t chunk;
};

// ---- small portable primitives ---------------------------------------------------------------
LGW_HD uint32_t fast_atomic_or(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
    return atomicOr(p, v);
#else
    const uint32_t o = *p; *p |= v; return o;
#endif
}
LGW_HD uint32_t fast_atomic_add(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
    return atomicAdd(p, v);
#else
    const uint32_t o = *p; *p += v; return o;
#endif
}
LGW_HD uint32_t fast_atomic_cas(uint32_t* p, uint32_t cmp, uint32_t v) {
#ifdef __CUDA_ARCH__
    return atomicCAS(p, cmp, v);
#else
    const uint32_t o = *p; if (o == cmp) *p = v; return o;
#endif
}

// four bytes at a time (p is a multiple of 4; the text buffer is 16-byte aligned)
LGW_HD uint32_t fj_word(const uint8_t* t, uint32_t p) {
#ifdef __CUDA_ARCH__
    return *reinterpret_cast<const uint32_t*>(t + p);
#else
    uint32_t v; memcpy(&v, t + p, 4); return v;
#endif
}
LGW_HD uint32_t fj_has_byte(uint32_t v, uint32_t b) { const uint32_t x = v ^ (b * 0x01010101u); return (x - 0x01010101u) & ~x & 0x80808080u; }
LGW_HD uint32_t fj_has_less(uint32_t v, uint32_t n) { return (v - n * 0x01010101u) & ~v & 0x80808080u; }
LGW_HD uint32_t fj_has_quote_or_backslash(uint32_t v) { return fj_has_byte(v, '"') | fj_has_byte(v, '\\'); }

LGW_HD bool fj_ws(uint32_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
LGW_HD bool fj_struct(uint32_t c) { return c == '{' || c == '}' || c == '[' || c == ']' || c == ',' || c == ':'; }
LGW_HD bool fj_digit(uint32_t c) { return c - '0' < 10u; }
LGW_HD bool fj_scalar_byte(uint32_t c) { return !fj_ws(c) && !fj_struct(c) && c != '"'; }
LGW_HD int fj_hex(uint32_t c) { return c - '0' < 10u ? (int)(c - '0') : ((c | 0x20) - 'a' < 6u ? (int)((c | 0x20) - 'a' + 10) : -1); }
LGW_HD int fj_hex4(const uint8_t* t) {
    const int a = fj_hex(t[0]), b = fj_hex(t[1]), c = fj_hex(t[2]), d = fj_hex(t[3]);
    return (a | b | c | d) < 0 ? -1 : (a << 12) | (b << 8) | (c << 4) | d;
}
// is the byte at position q an ACTIVE backslash (one that escapes the byte after it)?
LGW_HD bool fj_active_backslash(const uint8_t* t, uint32_t q) {
    if (t[q] != '\\') return false;
    uint32_t k = 1;
    while (q >= k && t[q - k] == '\\') ++k;
    return (k & 1u) != 0;
}

// rendering of one decoded character; returns the number of bytes put in buf (<= 12), -1 = the encoder raises
LGW_HD int fj_render_cp(uint32_t cp, bool lone, int mode, uint8_t* buf) {
    const bool ascii_only = mode != RM_HTTPX028;
    int o = 0;
    auto hex4 = [&](uint32_t cu) {
        buf[o++] = '\\'; buf[o++] = 'u';
        for (int s = 12; s >= 0; s -= 4) { const uint32_t h = (cu >> s) & 15u; buf[o++] = (uint8_t)(h < 10 ? '0' + h : 'a' + h - 10); }
    };
    auto two = [&](char x) { buf[o++] = '\\'; buf[o++] = (uint8_t)x; };
    if (cp == '"') two('"');
    else if (cp == '\\') two('\\');
    else if (cp == '\n') two('n');
    else if (cp == '\r') two('r');
    else if (cp == '\t') two('t');
    else if (cp == 8) two('b');
    else if (cp == 12) two('f');
    else if (mode == RM_JSON5 && cp == 11) two('v');
    else if (mode == RM_JSON5 && cp == 0) two('0');
    else if (cp < 0x20) hex4(cp);
    else if (cp < 0x7f) buf[o++] = (uint8_t)cp;
    else if (cp == 0x7f) { if (ascii_only) hex4(cp); else buf[o++] = 0x7f; }
    else if (lone) { if (ascii_only) hex4(cp); else return -1; }
    else if (ascii_only) {
        if (cp >= 0x10000) { const uint32_t v = cp - 0x10000; hex4(0xD800 + (v >> 10)); hex4(0xDC00 + (v & 0x3FF)); }
        else hex4(cp);
    }
    else if (cp < 0x800) { buf[o++] = (uint8_t)(0xC0 | (cp >> 6)); buf[o++] = (uint8_t)(0x80 | (cp & 63)); }
    else if (cp < 0x10000) { buf[o++] = (uint8_t)(0xE0 | (cp >> 12)); buf[o++] = (uint8_t)(0x80 | ((cp >> 6) & 63)); buf[o++] = (uint8_t)(0x80 | (cp & 63)); }
    else { buf[o++] = (uint8_t)(0xF0 | (cp >> 18)); buf[o++] = (uint8_t)(0x80 | ((cp >> 12) & 63)); buf[o++] = (uint8_t)(0x80 | ((cp >> 6) & 63)); buf[o++] = (uint8_t)(0x80 | (cp & 63)); }
    return o;
}

// ---- phases (each is called once per thread id; the driver separates them with barriers) -----------
LGW_HD void fph_load(FastCtx& c, int tid) {
    FastShared* s = c.s;
#ifdef __CUDA_ARCH__
    #pragma unroll 8
#endif
    for (uint32_t i = tid; i < c.n + 32; i += LGW_FAST_THREADS) s->text[i] = i < c.n ? c.in[i] : (uint8_t)' ';
    for (uint32_t i = tid; i < LGW_FAST_HT; i += LGW_FAST_THREADS) s->ht[i] = 0;
    if (tid == 0) { s->n_tok = 0; s->n_br = 0; s->irregular = 0; s->matched = 0; s->top_members = 0; s->out_len = 0; s->n_keys = 0; }
}

LGW_HD uint32_t fj_seq_len(uint32_t c) { return c >= 0xC2 && c <= 0xDF ? 2u : (c >= 0xE0 && c <= 0xEF ? 3u : (c >= 0xF0 && c <= 0xF4 ? 4u : 0u)); }

LGW_HD void fph_utf8(FastCtx& c, int tid) {
    FastShared* s = c.s;
    const uint32_t a = tid * c.chunk, b = a + c.chunk < c.n ? a + c.chunk : c.n;
    if (a >= c.n) return;
    const uint8_t* t = s->text;
    uint32_t any = 0, p0 = a;
    for (; p0 + 4 <= b; p0 += 4) any |= fj_word(t, p0);
    for (; p0 < b; ++p0) any |= (uint32_t)t[p0] << 24 >> 24;
    if (!(any & 0x80808080u)) return;
    uint32_t p = a;
    // continuation bytes at the start of the chunk: some lead byte before the chunk must cover them
    while (p < b && (t[p] & 0xC0u) == 0x80u) {
        uint32_t k = 1;
        while (k <= 3 && p >= k && (t[p - k] & 0xC0u) == 0x80u) ++k;
        if (k > 3 || p < k || fj_seq_len(t[p - k]) <= k) { s->irregular = 1; return; }
        ++p;
    }
    while (p < b) {
        const uint32_t ch = t[p];
        if (ch < 0x80) { ++p; continue; }
        const uint32_t len = fj_seq_len(ch);
        if (len == 0 || p + len > c.n) { s->irregular = 1; return; }
        uint32_t lo = 0x80, hi = 0xBF;
        if (ch == 0xE0) lo = 0xA0; else if (ch == 0xED) hi = 0x9F; else if (ch == 0xF0) lo = 0x90; else if (ch == 0xF4) hi = 0x8F;
        if (t[p + 1] < lo || t[p + 1] > hi) { s->irregular = 1; return; }
        for (uint32_t k = 2; k < len; ++k) if ((t[p + k] & 0xC0u) != 0x80u) { s->irregular = 1; return; }
        p += len;
    }
}

// does an escape that starts before `a` swallow bytes at/after `a`?  returns how many
LGW_HD uint32_t fj_swallowed(const uint8_t* t, uint32_t a) {
    for (uint32_t d = 1; d <= 5 && d <= a; ++d) {
        const uint32_t q = a - d;
        if (fj_active_backslash(t, q)) {
            const uint32_t elen = t[q + 1] == 'u' ? 6u : 2u;
            return q + elen > a ? q + elen - a : 0u;
        }
    }
    return 0;
}

LGW_HD void fph_quotes(FastCtx& c, int tid) {
    FastShared* s = c.s;
    const uint32_t a = tid * c.chunk, b = a + c.chunk < c.n ? a + c.chunk : c.n;
    uint32_t par = 0;
    if (a < c.n) {
        const uint8_t* t = s->text;
        bool esc = a > 0 && fj_active_backslash(t, a - 1);
        for (uint32_t p = a; p < b; ++p) {
            if (!esc && (p & 3u) == 0 && p + 4 <= b && !fj_has_quote_or_backslash(fj_word(t, p))) { p += 3; continue; }
            const uint32_t ch = t[p];
            if (esc) { esc = false; continue; }
            if (ch == '\\') esc = true; else if (ch == '"') par ^= 1u;
        }
    }
    s->scan[tid] = par;
}

// shared by the counting and the writing pass
template <bool WRITE>
LGW_HD void fph_tokens_impl(FastCtx& c, int tid) {
    FastShared* s = c.s;
    const uint32_t a = tid * c.chunk, b = a + c.chunk < c.n ? a + c.chunk : c.n;
    uint32_t ntok = 0, nbr = 0;
    bool in_str;
    if (!WRITE) { in_str = (s->scan[tid] & 1u) != 0; s->th_instr[tid] = in_str; }
    else {
        in_str = s->th_instr[tid] != 0;
        ntok = s->scan[tid] & 0xffffu; nbr = s->scan[tid] >> 16;
        s->th_tok0[tid] = (uint16_t)ntok; s->th_br0[tid] = (uint16_t)nbr;
        if (s->n_tok > LGW_FAST_MAXT || s->n_br > LGW_FAST_MAXBR) return;
    }
    if (a < c.n) {
        const uint8_t* t = s->text;
        bool esc = a > 0 && fj_active_backslash(t, a - 1);
        for (uint32_t p = a; p < b; ++p) {
            if (in_str && !esc && (p & 3u) == 0 && p + 4 <= b && !fj_has_quote_or_backslash(fj_word(t, p))) { p += 3; continue; }
            const uint32_t ch = t[p];
            if (in_str) {
                if (esc) esc = false;
                else if (ch == '\\') esc = true;
                else if (ch == '"') { in_str = false; if (WRITE) s->tok_end[ntok - 1] = (uint16_t)p; }
                continue;
            }
            // a backslash outside a string is never valid JSON; fph_quotes skipped the byte after it, so do the same and give up
            if (esc) { esc = false; s->irregular = 1; continue; }
            if (ch == '\\') { esc = true; s->irregular = 1; continue; }
            uint8_t type;
            if (ch == '"') { type = TK_STR; in_str = true; }
            else if (ch == '{') type = TK_OBJ_OPEN; else if (ch == '}') type = TK_OBJ_CLOSE;
            else if (ch == '[') type = TK_ARR_OPEN; else if (ch == ']') type = TK_ARR_CLOSE;
            else if (ch == ',') type = TK_COMMA; else if (ch == ':') type = TK_COLON;
            else if (fj_ws(ch)) continue;
            else { if (p > 0 && fj_scalar_byte(t[p - 1])) continue; type = TK_SCALAR; }
            if (WRITE) {
                s->tok_pos[ntok] = (uint16_t)p; s->tok_type[ntok] = type; s->tok_br[ntok] = (uint16_t)nbr; s->tok_flags[ntok] = 0; s->tok_aux[ntok] = 0;
                if (type <= TK_ARR_CLOSE) s->br_tok[nbr] = (uint16_t)ntok;
            }
            ++ntok;
            if (type <= TK_ARR_CLOSE) ++nbr;
        }
        if (b == c.n && in_str) s->irregular = 1;          // unterminated string
    }
    if (!WRITE) s->scan[tid] = ntok | (nbr << 16);
}
LGW_HD void fph_count(FastCtx& c, int tid) { fph_tokens_impl<false>(c, tid); }
LGW_HD void fph_tokens(FastCtx& c, int tid) { fph_tokens_impl<true>(c, tid); }

LGW_HD void fph_brackets(FastCtx& c, int tid) {
    if (tid != 0) return;
    FastShared* s = c.s;
    if (s->n_tok > LGW_FAST_MAXT || s->n_br > LGW_FAST_MAXBR || s->n_tok == 0) { s->irregular = 1; return; }
    uint8_t st_type[LGW_BODY_MAXD]; uint16_t st_br[LGW_BODY_MAXD];
    uint32_t depth = 0;
    for (uint32_t b = 0; b < s->n_br; ++b) {
        const uint8_t t = s->tok_type[s->br_tok[b]];
        if (t == TK_OBJ_OPEN || t == TK_ARR_OPEN) {
            if (depth >= LGW_BODY_MAXD) { s->irregular = 1; return; }
            st_type[depth] = t; st_br[depth] = (uint16_t)b; ++depth;
        } else {
            if (depth == 0 || st_type[depth - 1] != t - 1) { s->irregular = 1; return; }
            --depth;
        }
        s->br_depth_after[b] = (uint8_t)depth;
        s->br_ctx_after[b] = depth ? (st_type[depth - 1] == TK_OBJ_OPEN ? 1 : 2) : 0;
        s->br_open_after[b] = depth ? st_br[depth - 1] : (uint16_t)0xffff;
        s->br_nkeys[b] = 0;
    }
    if (depth != 0) s->irregular = 1;
    // the root must be an object that closes with the last token
    if (s->tok_type[0] != TK_OBJ_OPEN || s->tok_type[s->n_tok - 1] != TK_OBJ_CLOSE) s->irregular = 1;
}

LGW_HD void fph_tok_ctx(FastCtx& c, int tid) {
    FastShared* s = c.s;
    if (s->irregular) return;
    for (uint32_t i = tid; i < s->n_tok; i += LGW_FAST_THREADS) {
        const uint32_t nb = s->tok_br[i];
        const uint8_t ctx = nb ? s->br_ctx_after[nb - 1] : 0, depth = nb ? s->br_depth_after[nb - 1] : 0;
        s->tok_ctx[i] = ctx; s->tok_depth[i] = depth;
        if (s->tok_type[i] == TK_STR && ctx == 1 && i > 0 && (s->tok_type[i - 1] == TK_OBJ_OPEN || s->tok_type[i - 1] == TK_COMMA)) s->tok_flags[i] = TF_KEY;
    }
}

LGW_HD bool fj_reserved(const uint8_t* k, uint32_t n) { return BodyRewriter::is_reserved(k, n); }

// scalar token at p: returns the end (first delimiter) and classifies; kind 0 = invalid
LGW_HD uint32_t fj_scalar(const uint8_t* t, uint32_t p, int* kind) {
    uint32_t e = p;
    while (fj_scalar_byte(t[e])) ++e;              // the text is padded with spaces
    *kind = 0;
    uint32_t q = p;
    if (t[q] == '-') ++q;
    if (fj_digit(t[q])) {
        if (t[q] == '0') ++q; else while (fj_digit(t[q])) ++q;
        bool flt = false;
        if (t[q] == '.') { ++q; if (!(fj_digit(t[q]))) return e; while (fj_digit(t[q])) ++q; flt = true; }
        if ((t[q] | 0x20) == 'e') { ++q; if (t[q] == '+' || t[q] == '-') ++q; if (!(fj_digit(t[q]))) return e; while (fj_digit(t[q])) ++q; flt = true; }
        if (q == e) *kind = flt ? 2 : 1;
        return e;
    }
    auto is = [&](const char* w, uint32_t from) { uint32_t i = 0; while (w[i] && t[from + i] == (uint8_t)w[i]) ++i; return w[i] == 0 && from + i == e; };
    if (is("true", p) || is("false", p) || is("null", p)) *kind = 3;
    else if (is("NaN", p) || is("Infinity", p) || is("-Infinity", p)) *kind = 4;
    return e;
}

LGW_HD void fph_tok_check(FastCtx& c, int tid) {
    FastShared* s = c.s;
    if (s->irregular) return;
    const uint8_t* t = s->text;
    enum { E_VALUE, E_VALUE_OR_CLOSE, E_KEY, E_KEY_OR_CLOSE, E_COLON, E_AFTER };
    for (uint32_t i = tid; i < s->n_tok; i += LGW_FAST_THREADS) {
        const uint8_t type = s->tok_type[i];
        int exp;
        if (i == 0) exp = E_VALUE;
        else {
            if (s->tok_depth[i] == 0) { s->irregular = 1; return; }      // something after the root value
            const uint8_t pt = s->tok_type[i - 1];
            if (pt == TK_OBJ_OPEN) exp = E_KEY_OR_CLOSE;
            else if (pt == TK_ARR_OPEN) exp = E_VALUE_OR_CLOSE;
            else if (pt == TK_COMMA) exp = s->tok_ctx[i - 1] == 1 ? E_KEY : E_VALUE;
            else if (pt == TK_COLON) exp = E_VALUE;
            else if (pt == TK_STR) exp = (s->tok_flags[i - 1] & TF_KEY) ? E_COLON : E_AFTER;
            else exp = E_AFTER;
        }
        bool ok;
        switch (type) {
        case TK_STR: ok = exp == E_VALUE || exp == E_VALUE_OR_CLOSE || exp == E_KEY || exp == E_KEY_OR_CLOSE; break;
        case TK_SCALAR: case TK_OBJ_OPEN: case TK_ARR_OPEN: ok = exp == E_VALUE || exp == E_VALUE_OR_CLOSE; break;
        case TK_COLON: ok = exp == E_COLON; break;
        case TK_COMMA: ok = exp == E_AFTER; break;
        case TK_OBJ_CLOSE: ok = exp == E_KEY_OR_CLOSE || exp == E_AFTER; break;
        default: ok = exp == E_VALUE_OR_CLOSE || exp == E_AFTER; break;      // TK_ARR_CLOSE
        }
        if (!ok) { s->irregular = 1; return; }
        if (type == TK_SCALAR) {
            int kind;
            const uint32_t p = s->tok_pos[i], e = fj_scalar(t, p, &kind);
            s->tok_end[i] = (uint16_t)e;
            if (kind == 0 || e - p > LGW_BODY_NUMCAP || (kind == 4 && c.mode == RM_HTTPX028)) { s->irregular = 1; return; }
            uint8_t f = 0;
            if (kind == 2) f = TF_FLOAT;
            else if (kind == 1 && e - p == 2 && t[p] == '-' && t[p + 1] == '0') f = TF_NEGZERO;
            s->tok_flags[i] = f;
        } else if (type == TK_STR && (s->tok_flags[i] & TF_KEY)) {
            const uint32_t p = s->tok_pos[i] + 1, e = s->tok_end[i], len = e - p;
            if (len > LGW_BODY_KEYCAP) { s->irregular = 1; return; }
            uint32_t h = 2166136261u; bool ident = len > 0;
            for (uint32_t q = p; q < e; ++q) {
                const uint32_t ch = t[q];
                if (ch == '\\' || (ch >= 0x80 && c.mode == RM_JSON5)) { s->irregular = 1; return; }
                h ^= ch; h *= 16777619u;
                const bool alpha = (ch | 0x20) - 'a' < 26u || ch == '_' || ch == '$';
                if (!(alpha || (q > p && ch - '0' < 10u))) ident = false;
            }
            const uint32_t owner = s->br_open_after[s->tok_br[i] - 1];
            uint32_t v = (h ^ (len * 0x9E3779B1u)) ^ (owner * 0x85EBCA6Bu);
            if (v == 0) v = 1;
            if (fast_atomic_add(&s->n_keys, 1) >= LGW_FAST_HT / 2) { s->irregular = 1; return; }
            if (fast_atomic_add(&s->br_nkeys[owner], 1) >= LGW_BODY_MAXKEYS) { s->irregular = 1; return; }
            for (uint32_t slot = v & (LGW_FAST_HT - 1);; slot = (slot + 1) & (LGW_FAST_HT - 1)) {
                const uint32_t old = fast_atomic_cas(&s->ht[slot], 0, v);
                if (old == 0) break;
                if (old == v) { s->irregular = 1; return; }          // same hash in the same object: let the exact machine decide
            }
            uint8_t f = TF_KEY;
            if (c.mode == RM_JSON5 && ident && !fj_reserved(t + p, len)) f |= TF_UNQUOTED;
            if (s->tok_depth[i] == 1) {
                f |= TF_TOPKEY;
                fast_atomic_add(&s->top_members, 1);
                for (uint32_t k = 0; k < c.n_ops; ++k) {
                    const BodyOp& op = c.ops[k];
                    if (op.key_len != len) continue;
                    uint32_t j = 0;
                    while (j < len && c.blob[op.key_off + j] == t[p + j]) ++j;
                    if (j == len) { s->tok_aux[i] = (uint8_t)(k + 1); fast_atomic_or(&s->matched, 1u << k); break; }
                }
            }
            s->tok_flags[i] = f;
        }
    }
}

// members: which top-level key does a token belong to?  per-thread token segments + block max-scan
LGW_HD uint32_t fj_tok_per_thread(const FastShared* s) { return (s->n_tok + LGW_FAST_THREADS - 1) / LGW_FAST_THREADS; }
LGW_HD void fph_member_a(FastCtx& c, int tid) {
    FastShared* s = c.s;
    uint32_t last = 0;
    if (!s->irregular) {
        const uint32_t tpt = fj_tok_per_thread(s), a = tid * tpt, b = a + tpt < s->n_tok ? a + tpt : s->n_tok;
        for (uint32_t i = a; i < b; ++i) if (s->tok_flags[i] & TF_TOPKEY) last = i + 1;
    }
    s->scan[tid] = last;
}
LGW_HD void fph_member_b(FastCtx& c, int tid) {
    FastShared* s = c.s;
    if (s->irregular) return;
    const uint32_t tpt = fj_tok_per_thread(s), a = tid * tpt, b = a + tpt < s->n_tok ? a + tpt : s->n_tok;
    uint32_t cur = s->scan[tid];                       // 1 + index of the nearest top-level key before this segment (0: none)
    for (uint32_t i = a; i < b; ++i) {
        if (s->tok_flags[i] & TF_TOPKEY) { cur = i + 1; continue; }
        if (cur == 0) continue;
        const uint32_t key = cur - 1, op1 = s->tok_aux[key];
        if (op1 == 0 || (c.ops[op1 - 1].flags & 3u)) continue;                     // not assigned, assigned only if absent, or a presence probe
        if (i < key + 2 || i == s->n_tok - 1) continue;                            // the colon; the root's closing brace
        if (s->tok_depth[i] == 1 && s->tok_type[i] == TK_COMMA) continue;          // the comma that ends the member
        uint8_t f = s->tok_flags[i] | TF_SKIP;
        if (i == key + 2) { f |= TF_REPLACE; s->tok_aux[i] = (uint8_t)op1; }
        s->tok_flags[i] = f;
    }
}

// the rendering walk over one thread's chunk: WRITE=false counts, WRITE=true stores at out+base
template <bool WRITE>
LGW_HD uint32_t fj_walk(FastCtx& c, int tid, uint32_t base) {
    FastShared* s = c.s;
    const uint32_t a = tid * c.chunk, b = a + c.chunk < c.n ? a + c.chunk : c.n;
    if (a >= c.n) return 0;
    const uint8_t* t = s->text;
    const int mode = c.mode;
    uint32_t o = 0;
    uint8_t* const out = c.out;
#define FJ_PUT(x) do { if (WRITE) out[base + o] = (uint8_t)(x); ++o; } while (0)
    int ti = (int)s->th_tok0[tid] - 1;
    uint32_t next_pos = (uint32_t)(ti + 1) < s->n_tok ? s->tok_pos[ti + 1] : 0xffffffffu;
    uint8_t fl = ti >= 0 ? s->tok_flags[ti] : 0;
    bool in_str = s->th_instr[tid] != 0;
    uint32_t p = a;
    if (in_str) p += fj_swallowed(t, a);
    uint8_t buf[12];
    while (p < b) {
        if (in_str && (p & 3u) == 0 && p + 4 <= b) {                 // four plain characters at once
            const uint32_t v = fj_word(t, p);
            uint32_t special = fj_has_quote_or_backslash(v) | fj_has_less(v, 0x20);
            if (mode != RM_HTTPX028) special |= (v | (v + 0x01010101u)) & 0x80808080u;      // >= 0x7f needs \\uXXXX there
            if (!special) {
                if (!(fl & TF_SKIP)) {
                    if (WRITE) { out[base + o] = (uint8_t)v; out[base + o + 1] = (uint8_t)(v >> 8); out[base + o + 2] = (uint8_t)(v >> 16); out[base + o + 3] = (uint8_t)(v >> 24); }
                    o += 4;
                }
                p += 4;
                continue;
            }
        }
        const uint32_t ch = t[p];
        if (in_str) {
            const bool quiet = (fl & TF_SKIP) != 0;
            if (ch == '"') { in_str = false; if (!quiet && !(fl & TF_UNQUOTED)) FJ_PUT('"'); ++p; continue; }
            if (ch == '\\') {
                const uint32_t e = t[p + 1];
                uint32_t cp; bool lone = false, silent = false; uint32_t adv = 2;
                switch (e) {
                case '"': cp = '"'; break;   case '\\': cp = '\\'; break;   case '/': cp = '/'; break;
                case 'b': cp = 8; break;     case 'f': cp = 12; break;     case 'n': cp = 10; break;
                case 'r': cp = 13; break;    case 't': cp = 9; break;
                case 'u': {
                    const int v = fj_hex4(t + p + 2);
                    if (v < 0) { s->irregular = 1; return o; }
                    cp = (uint32_t)v; adv = 6;
                    if (cp >= 0xD800 && cp <= 0xDBFF) {
                        int lo = -1;
                        if (t[p + 6] == '\\' && t[p + 7] == 'u') lo = fj_hex4(t + p + 8);
                        if (lo >= 0xDC00 && lo <= 0xDFFF) cp = 0x10000 + ((cp - 0xD800) << 10) + ((uint32_t)lo - 0xDC00);
                        else lone = true;
                    } else if (cp >= 0xDC00 && cp <= 0xDFFF) {
                        int hi = -1;
                        if (p >= 6 && t[p - 5] == 'u' && fj_active_backslash(t, p - 6)) hi = fj_hex4(t + p - 4);
                        if (hi >= 0xD800 && hi <= 0xDBFF) silent = true;          // rendered with its high half
                        else lone = true;
                    }
                    break; }
                default: s->irregular = 1; return o;
                }
                if (!quiet && !silent) {
                    const int k = fj_render_cp(cp, lone, mode, buf);
                    if (k < 0) { s->irregular = 1; return o; }                     // the encoder raises: exact machine reports it
                    for (int j = 0; j < k; ++j) FJ_PUT(buf[j]);
                }
                p += adv;
                continue;
            }
            if (ch < 0x20) { s->irregular = 1; return o; }
            if (!quiet) {
                if (ch < 0x7f || mode == RM_HTTPX028) FJ_PUT(ch);
                else if (ch == 0x7f) { const int k = fj_render_cp(0x7f, false, mode, buf); for (int j = 0; j < k; ++j) FJ_PUT(buf[j]); }
                else if ((ch & 0xC0u) != 0x80u) {                                  // lead byte: decode (validated by fph_utf8)
                    const uint32_t len = fj_seq_len(ch);
                    uint32_t cp = len == 2 ? (ch & 31u) : (len == 3 ? (ch & 15u) : (ch & 7u));
                    for (uint32_t k = 1; k < len; ++k) cp = (cp << 6) | (t[p + k] & 63u);
                    const int k = fj_render_cp(cp, false, mode, buf);
                    for (int j = 0; j < k; ++j) FJ_PUT(buf[j]);
                }
            }
            ++p;
            continue;
        }
        // outside strings only token starts produce output
        if (p != next_pos) { ++p; continue; }
        ++ti; fl = s->tok_flags[ti];
        next_pos = (uint32_t)(ti + 1) < s->n_tok ? s->tok_pos[ti + 1] : 0xffffffffu;
        const uint8_t type = s->tok_type[ti];
        if (fl & TF_REPLACE) { const BodyOp& op = c.ops[s->tok_aux[ti] - 1]; for (uint32_t j = 0; j < op.rval_len; ++j) FJ_PUT(c.blob[op.rval_off + j]); }
        if (type == TK_STR) { in_str = true; if (!(fl & (TF_SKIP | TF_UNQUOTED))) FJ_PUT('"'); ++p; continue; }
        if (!(fl & TF_SKIP)) {
            switch (type) {
            case TK_OBJ_OPEN: FJ_PUT('{'); break;
            case TK_ARR_OPEN: FJ_PUT('['); break;
            case TK_ARR_CLOSE: FJ_PUT(']'); break;
            case TK_COMMA: FJ_PUT(','); if (mode != RM_HTTPX028) FJ_PUT(' '); break;
            case TK_COLON: FJ_PUT(':'); if (mode != RM_HTTPX028) FJ_PUT(' '); break;
            case TK_OBJ_CLOSE:
                if ((uint32_t)ti == s->n_tok - 1) {                                // root: append the assigned keys the client did not send
                    bool any = s->top_members > 0;
                    for (uint32_t k = 0; k < c.n_ops; ++k) {
                        const BodyOp& op = c.ops[k];
                        if ((s->matched & (1u << k)) || (op.flags & 2u)) continue;
                        if (any) { FJ_PUT(','); if (mode != RM_HTTPX028) FJ_PUT(' '); }
                        for (uint32_t j = 0; j < op.rkey_len; ++j) FJ_PUT(c.blob[op.rkey_off + j]);
                        FJ_PUT(':'); if (mode != RM_HTTPX028) FJ_PUT(' ');
                        for (uint32_t j = 0; j < op.rval_len; ++j) FJ_PUT(c.blob[op.rval_off + j]);
                        any = true;
                    }
                }
                FJ_PUT('}');
                break;
            default: {                                                             // TK_SCALAR: the whole token at its first byte
                const uint32_t e = s->tok_end[ti];
                if (fl & TF_FLOAT) {
                    char fb[32];
                    const int k = format_float((const char*)t + p, e - p, fb);
                    if (k < 0) { s->irregular = 1; return o; }
                    for (int j = 0; j < k; ++j) FJ_PUT(fb[j]);
                } else if (fl & TF_NEGZERO) FJ_PUT('0');
                else for (uint32_t q = p; q < e; ++q) FJ_PUT(t[q]);
                break; }
            }
        }
        ++p;
    }
#undef FJ_PUT
    return o;
}
LGW_HD void fph_size(FastCtx& c, int tid) { c.s->scan[tid] = c.s->irregular ? 0u : fj_walk<false>(c, tid, 0); }
LGW_HD void fph_write(FastCtx& c, int tid) {
    FastShared* s = c.s;
    if (s->irregular || s->scan[LGW_FAST_THREADS] > c.cap) return;
    fj_walk<true>(c, tid, s->scan[tid]);
}

// ---- block-wide scans over s->scan[0..THREADS): exclusive, total in scan[THREADS] --------------------
template <bool MAX>
LGW_HD void fast_scan(FastShared* s) {
#ifdef __CUDA_ARCH__
    __syncthreads();
    if (threadIdx.x < 32) {
        const uint32_t lane = threadIdx.x;
        uint32_t v[LGW_FAST_THREADS / 32];
        uint32_t acc = 0;
        #pragma unroll
        for (int k = 0; k < LGW_FAST_THREADS / 32; ++k) { v[k] = s->scan[lane * (LGW_FAST_THREADS / 32) + k]; acc = MAX ? (acc > v[k] ? acc : v[k]) : acc + v[k]; }
        uint32_t inc = acc;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (uint32_t)d) inc = MAX ? (inc > o ? inc : o) : inc + o; }
        uint32_t run = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) run = 0;
        #pragma unroll
        for (int k = 0; k < LGW_FAST_THREADS / 32; ++k) { s->scan[lane * (LGW_FAST_THREADS / 32) + k] = run; run = MAX ? (run > v[k] ? run : v[k]) : run + v[k]; }
        if (lane == 31) s->scan[LGW_FAST_THREADS] = run;
    }
    __syncthreads();
#else
    uint32_t run = 0;
    for (int i = 0; i < LGW_FAST_THREADS; ++i) { const uint32_t v = s->scan[i]; s->scan[i] = run; run = MAX ? (run > v ? run : v) : run + v; }
    s->scan[LGW_FAST_THREADS] = run;
#endif
}

#ifdef __CUDA_ARCH__
#define LGW_FPHASE(fn) do { fn(c, (int)threadIdx.x); __syncthreads(); } while (0)
#else
#define LGW_FPHASE(fn) do { for (int _t = 0; _t < LGW_FAST_THREADS; ++_t) fn(c, _t); } while (0)
#endif

// One body.  Returns BS_OK / BS_OVERFLOW (with *out_len) or LGW_FAST_IRREGULAR.  On the device every thread
// of the block calls this with identical arguments and gets the same answer.
LGW_HD uint32_t fast_rewrite(FastShared* sh, const uint8_t* in, uint32_t n, int mode, const BodyOp* ops, uint32_t n_ops,
                             const uint8_t* blob, uint8_t* out, uint32_t cap, uint32_t* out_len, uint32_t* matched_out) {
    *out_len = 0; *matched_out = 0;
    if (n == 0 || n > LGW_FAST_MAXB) return LGW_FAST_IRREGULAR;
    FastCtx c{sh, in, n, mode & 0xff, ops, n_ops, blob, out, cap, 0};      // (bit 8 = response mode: only matters for non-object roots, which are irregular here)
    c.chunk = (((n + LGW_FAST_THREADS - 1) / LGW_FAST_THREADS) + 7u) & ~7u;
    LGW_FPHASE(fph_load);
    LGW_FPHASE(fph_utf8);
    LGW_FPHASE(fph_quotes);
    fast_scan<false>(sh);
    LGW_FPHASE(fph_count);
    fast_scan<false>(sh);
#ifdef __CUDA_ARCH__
    if (threadIdx.x == 0)
#endif
    { sh->n_tok = sh->scan[LGW_FAST_THREADS] & 0xffffu; sh->n_br = sh->scan[LGW_FAST_THREADS] >> 16; }
#ifdef __CUDA_ARCH__
    __syncthreads();
#endif
    LGW_FPHASE(fph_tokens);
    LGW_FPHASE(fph_brackets);
    LGW_FPHASE(fph_tok_ctx);
    LGW_FPHASE(fph_tok_check);
    LGW_FPHASE(fph_member_a);
    fast_scan<true>(sh);
    LGW_FPHASE(fph_member_b);
    LGW_FPHASE(fph_size);
    fast_scan<false>(sh);
    if (sh->irregular) return LGW_FAST_IRREGULAR;
    *out_len = sh->scan[LGW_FAST_THREADS];
    *matched_out = sh->matched;
    if (*out_len > cap) return BS_OVERFLOW;
    LGW_FPHASE(fph_write);
    return BS_OK;
}

}  // namespace lgw
