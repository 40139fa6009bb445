// Lean JSON recogniser for the bulk kernel: validity (handler view / tap view) and the four
// top-level keys the reference's loops test on every event -- "error", "detail", "code", "usage"
// (request_handler.py:50,86,123,133; chat_logging.py:134,137).  Everything else the tap reads
// (choices walk, usage numbers) matters only for events that carry "usage" or "error"; those are
// rare and are re-parsed with the full machine (json_machine.cuh).
//
// Same grammar as json_machine.cuh (CPython json.loads, strict).  Table driven outside strings:
// class = CLS[byte], entry = TRANS[state][class] (next state + action), so a structural byte costs
// two shared-memory loads and a handful of ALU ops; bytes inside strings take a three-compare
// fast path.  tests/test_machine_cpu.py fuzzes it against the full machine and CPython.
#pragma once
#include <stdint.h>
#include <initializer_list>
#include "json_machine.cuh"

namespace lgw {

enum LeanSt : uint8_t {
    L_VALUE = 0, L_VALUE_OR_END, L_KEY_OR_END, L_KEY, L_COLON, L_AFTER,
    L_NUM_MINUS, L_NUM_ZERO, L_NUM_INT, L_NUM_DOT, L_NUM_FRAC, L_NUM_E, L_NUM_ESIGN, L_NUM_EXP,
    L_DONE, L_TRAIL_B, L_ERR,                 // 0..16: table rows
    L_STR, L_STR_ESC, L_STR_U, L_LIT          // arithmetic states
};
#define LGW_LEAN_ROWS 17

enum LeanCls : uint8_t {
    C_OTHER = 0, C_WS, C_QUOTE, C_LBRACE, C_RBRACE, C_LBRACK, C_RBRACK, C_COLON, C_COMMA,
    C_MINUS, C_PLUS, C_ZERO, C_DIGIT, C_DOT, C_EXP, C_LIT, C_PYWS, C_NCLS
};

enum LeanAct : uint8_t {
    A_NONE = 0, A_OPEN_OBJ = 1, A_OPEN_ARR = 2, A_CLOSE_OBJ = 3, A_CLOSE_ARR = 4, A_COMMA = 5,
    A_BEGIN_STR = 6, A_SPECIAL = 7
};

struct LeanTables {
    uint8_t cls[256];
    uint8_t trans[LGW_LEAN_ROWS * 32];       // entry = next | (action << 5)
};

constexpr uint8_t lean_entry(uint8_t next, uint8_t act) { return (uint8_t)(next | (act << 5)); }

constexpr LeanTables make_lean_tables() {
    LeanTables t{};
    for (int i = 0; i < 256; ++i) t.cls[i] = C_OTHER;
    t.cls[' '] = t.cls['\t'] = t.cls['\n'] = t.cls['\r'] = C_WS;
    t.cls['"'] = C_QUOTE; t.cls['{'] = C_LBRACE; t.cls['}'] = C_RBRACE; t.cls['['] = C_LBRACK; t.cls[']'] = C_RBRACK;
    t.cls[':'] = C_COLON; t.cls[','] = C_COMMA; t.cls['-'] = C_MINUS; t.cls['+'] = C_PLUS; t.cls['0'] = C_ZERO;
    for (int c = '1'; c <= '9'; ++c) t.cls[c] = C_DIGIT;
    t.cls['.'] = C_DOT; t.cls['e'] = t.cls['E'] = C_EXP;
    t.cls['t'] = t.cls['f'] = t.cls['n'] = t.cls['N'] = t.cls['I'] = C_LIT;
    t.cls[0x0b] = t.cls[0x0c] = t.cls[0x1c] = t.cls[0x1d] = t.cls[0x1e] = t.cls[0x1f] = C_PYWS;

    for (int i = 0; i < LGW_LEAN_ROWS * 32; ++i) t.trans[i] = lean_entry(L_ERR, A_NONE);
    auto set = [&](int st, int cl, uint8_t next, uint8_t act) { t.trans[st * 32 + cl] = lean_entry(next, act); };
    // value starts
    for (int st : {(int)L_VALUE, (int)L_VALUE_OR_END}) {
        set(st, C_WS, (uint8_t)st, A_NONE);
        set(st, C_QUOTE, L_STR, A_BEGIN_STR);
        set(st, C_LBRACE, L_KEY_OR_END, A_OPEN_OBJ);
        set(st, C_LBRACK, L_VALUE_OR_END, A_OPEN_ARR);
        set(st, C_MINUS, L_NUM_MINUS, A_NONE);
        set(st, C_ZERO, L_NUM_ZERO, A_NONE);
        set(st, C_DIGIT, L_NUM_INT, A_NONE);
        set(st, C_LIT, L_ERR, A_SPECIAL);
    }
    set(L_VALUE_OR_END, C_RBRACK, L_AFTER, A_CLOSE_ARR);
    set(L_KEY_OR_END, C_WS, L_KEY_OR_END, A_NONE); set(L_KEY_OR_END, C_QUOTE, L_STR, A_BEGIN_STR); set(L_KEY_OR_END, C_RBRACE, L_AFTER, A_CLOSE_OBJ);
    set(L_KEY, C_WS, L_KEY, A_NONE); set(L_KEY, C_QUOTE, L_STR, A_BEGIN_STR);
    set(L_COLON, C_WS, L_COLON, A_NONE); set(L_COLON, C_COLON, L_VALUE, A_NONE);
    // after a value (also: the character that terminates a number)
    for (int st : {(int)L_AFTER, (int)L_NUM_ZERO, (int)L_NUM_INT, (int)L_NUM_FRAC, (int)L_NUM_EXP}) {
        set(st, C_WS, L_AFTER, A_NONE);
        set(st, C_COMMA, L_ERR, A_COMMA);
        set(st, C_RBRACE, L_AFTER, A_CLOSE_OBJ);
        set(st, C_RBRACK, L_AFTER, A_CLOSE_ARR);
    }
    set(L_NUM_MINUS, C_ZERO, L_NUM_ZERO, A_NONE); set(L_NUM_MINUS, C_DIGIT, L_NUM_INT, A_NONE); set(L_NUM_MINUS, C_LIT, L_ERR, A_SPECIAL);
    set(L_NUM_ZERO, C_DOT, L_NUM_DOT, A_NONE); set(L_NUM_ZERO, C_EXP, L_NUM_E, A_NONE);
    set(L_NUM_INT, C_ZERO, L_NUM_INT, A_NONE); set(L_NUM_INT, C_DIGIT, L_NUM_INT, A_NONE); set(L_NUM_INT, C_DOT, L_NUM_DOT, A_NONE); set(L_NUM_INT, C_EXP, L_NUM_E, A_NONE);
    set(L_NUM_DOT, C_ZERO, L_NUM_FRAC, A_NONE); set(L_NUM_DOT, C_DIGIT, L_NUM_FRAC, A_NONE);
    set(L_NUM_FRAC, C_ZERO, L_NUM_FRAC, A_NONE); set(L_NUM_FRAC, C_DIGIT, L_NUM_FRAC, A_NONE); set(L_NUM_FRAC, C_EXP, L_NUM_E, A_NONE);
    set(L_NUM_E, C_PLUS, L_NUM_ESIGN, A_NONE); set(L_NUM_E, C_MINUS, L_NUM_ESIGN, A_NONE); set(L_NUM_E, C_ZERO, L_NUM_EXP, A_NONE); set(L_NUM_E, C_DIGIT, L_NUM_EXP, A_NONE);
    set(L_NUM_ESIGN, C_ZERO, L_NUM_EXP, A_NONE); set(L_NUM_ESIGN, C_DIGIT, L_NUM_EXP, A_NONE);
    set(L_NUM_EXP, C_ZERO, L_NUM_EXP, A_NONE); set(L_NUM_EXP, C_DIGIT, L_NUM_EXP, A_NONE);
    set(L_DONE, C_WS, L_DONE, A_NONE); set(L_DONE, C_PYWS, L_ERR, A_SPECIAL);
    set(L_TRAIL_B, C_WS, L_TRAIL_B, A_NONE); set(L_TRAIL_B, C_PYWS, L_TRAIL_B, A_NONE);
    return t;
}

#if defined(__CUDACC__)
__device__ const LeanTables g_lean_tables_dev = make_lean_tables();
#endif
static const LeanTables g_lean_tables_host = make_lean_tables();

constexpr uint64_t lit_word(const char* s) {
    uint64_t v = 0;
    for (int i = 0; s[i]; ++i) v |= (uint64_t)(uint8_t)s[i] << (8 * i);
    return v;
}

struct LeanMachine {
    uint32_t st, depth, flags, in_key, key_esc, kstart, ucount, strip;
    uint64_t stack, litw;

    LGW_HD void reset(bool strip_mode) {
        st = L_VALUE; depth = 0; flags = 0; in_key = 0; key_esc = 0; kstart = 0; ucount = 0; strip = strip_mode ? 1u : 0u;
        stack = 0; litw = 0;
    }

    // key text [kstart, kend) at depth 1 without escapes: compare with the four tracked names
    template <class R>
    LGW_HD void match_key(const R& rd, uint32_t kend) {
        const uint32_t n = kend - kstart;
        if (n < 4 || n > 6) return;
        uint64_t k = 0;
        for (uint32_t i = 0; i < n; ++i) k |= (uint64_t)rd.at(kstart + i) << (8 * i);
        if (n == 5) { if (k == lit_word("error")) flags |= TK_ERROR; else if (k == lit_word("usage")) flags |= TK_USAGE; }
        else if (n == 6) { if (k == lit_word("detail")) flags |= TK_DETAIL; }
        else if (k == lit_word("code")) flags |= TK_CODE;
    }
    // same with escapes decoded (rare)
    template <class R>
    LGW_HD void match_key_escaped(const R& rd, uint32_t kend) {
        uint64_t k = 0; uint32_t n = 0; bool bad = false;
        for (uint32_t i = kstart; i < kend; ++i) {
            uint32_t c = rd.at(i);
            if (c == '\\') {
                const uint32_t e = rd.at(++i);
                if (e == 'u') {
                    c = 0;
                    for (int d = 0; d < 4; ++d) { const uint32_t h = rd.at(++i); c = (c << 4) | (h - '0' < 10u ? h - '0' : (h | 0x20) - 'a' + 10); }
                } else c = e == 'b' ? 8 : e == 'f' ? 12 : e == 'n' ? 10 : e == 'r' ? 13 : e == 't' ? 9 : e;
            }
            if (c >= 0x80) bad = true;
            if (n < 8) k |= (uint64_t)(c & 0xff) << (8 * n);
            ++n;
        }
        if (bad) return;
        if (n == 5) { if (k == lit_word("error")) flags |= TK_ERROR; else if (k == lit_word("usage")) flags |= TK_USAGE; }
        else if (n == 6) { if (k == lit_word("detail")) flags |= TK_DETAIL; }
        else if (n == 4 && k == lit_word("code")) flags |= TK_CODE;
    }

    // R provides at(pos) (event bytes, for key matching), cls(byte) and trans(index) (the tables)
    template <class R>
    LGW_HD void step(uint32_t c, uint32_t pos, const R& rd) {
        if (st == L_STR) {
            if (c == '"') {
                if (in_key) {
                    st = L_COLON;
                    if (depth == 1) { if (key_esc) match_key_escaped(rd, pos); else match_key(rd, pos); }
                } else st = L_AFTER;
            } else if (c == '\\') { st = L_STR_ESC; key_esc = 1; }
            else if (c < 0x20) st = L_ERR;
            return;
        }
        if (st < LGW_LEAN_ROWS) {
            const uint32_t cl = rd.cls(c);
            const uint32_t e = rd.trans(st * 32 + cl);
            const uint32_t act = e >> 5;
            const uint32_t prev = st;
            st = e & 31u;
            if (act == A_NONE) return;
            switch (act) {
            case A_BEGIN_STR:
                in_key = (prev == L_KEY || prev == L_KEY_OR_END) ? 1u : 0u;
                kstart = pos + 1; key_esc = 0;
                return;
            case A_OPEN_OBJ: case A_OPEN_ARR:
                if (depth >= LGW_MAX_DEPTH) { flags |= PF_TOO_DEEP; st = L_ERR; return; }
                if (act == A_OPEN_OBJ) stack |= (1ull << depth); else stack &= ~(1ull << depth);
                ++depth;
                return;
            case A_CLOSE_OBJ: case A_CLOSE_ARR: {
                if (depth == 0) { st = L_ERR; return; }
                const bool top_obj = (stack >> (depth - 1)) & 1ull;
                if (top_obj != (act == A_CLOSE_OBJ)) { st = L_ERR; return; }
                --depth;
                st = depth == 0 ? L_DONE : L_AFTER;
                return; }
            case A_COMMA:
                if (depth == 0) { st = L_ERR; return; }
                st = ((stack >> (depth - 1)) & 1ull) ? L_KEY : L_VALUE;
                return;
            default:   // A_SPECIAL
                if (prev == L_DONE) { st = strip ? L_TRAIL_B : L_ERR; return; }
                if (prev == L_NUM_MINUS) { if (c == 'I') { litw = lit_word("nfinity"); st = L_LIT; } else st = L_ERR; return; }
                litw = c == 't' ? lit_word("rue") : c == 'f' ? lit_word("alse") : c == 'n' ? lit_word("ull")
                     : c == 'N' ? lit_word("aN") : lit_word("nfinity");
                st = L_LIT;
                return;
            }
        }
        if (st == L_STR_ESC) {
            if (c == 'u') { st = L_STR_U; ucount = 0; }
            else if (c == '"' || c == '\\' || c == '/' || c == 'b' || c == 'f' || c == 'n' || c == 'r' || c == 't') st = L_STR;
            else st = L_ERR;
            return;
        }
        if (st == L_STR_U) {
            if (c - '0' < 10u || (c | 0x20) - 'a' < 6u) { if (++ucount == 4) st = L_STR; }
            else st = L_ERR;
            return;
        }
        // L_LIT
        if ((uint32_t)(litw & 0xff) != c) { st = L_ERR; return; }
        litw >>= 8;
        if (litw == 0) st = depth == 0 ? L_DONE : L_AFTER;
    }

    LGW_HD uint32_t finish() const {
        uint32_t f = flags;
        if (st == L_DONE) f |= PF_VALID_A | PF_VALID_B;
        else if (st == L_TRAIL_B) f |= PF_VALID_B;
        return f;
    }
};

// byte source + tables for host code and for device code that reads plain memory
struct PlainEnv {
    const uint8_t* a; uint32_t na;          // two-piece rope (carry, chunk)
    const uint8_t* b; uint32_t nb;
    const uint8_t* cls_tab; const uint8_t* trans_tab;
    LGW_HD uint32_t at(uint32_t i) const { return i < na ? a[i] : b[i - na]; }
    LGW_HD uint32_t cls(uint32_t c) const { return cls_tab[c]; }
    LGW_HD uint32_t trans(uint32_t i) const { return trans_tab[i]; }
};

LGW_HD const LeanTables& lean_tables() {
#if defined(__CUDA_ARCH__)
    return g_lean_tables_dev;
#else
    return g_lean_tables_host;
#endif
}

// reference byte loop over [s, e) of an env (the bulk kernel has its own word-wise loop)
template <class R>
LGW_HD uint32_t lean_parse(const R& rd, uint32_t s, uint32_t e, bool strip) {
    LeanMachine m;
    m.reset(strip);
    for (uint32_t i = s; i < e; ++i) m.step(rd.at(i), i, rd);
    return m.finish();
}

}  // namespace lgw
