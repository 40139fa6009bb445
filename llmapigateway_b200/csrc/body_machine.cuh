// Request-body rewrite: parse the client's JSON body, apply one attempt's
// key assignments at the top level, and re-serialise it the way the reference's encoder would.
//
//   llm_gateway_core/api/v1/chat.py:31-45      parse (400 conditions)               -> scan_body()
//   chat.py:112-119, :135-139, :150, :164-168   deepcopy + payload[key] = value      -> ops (BodyOp list)
//   services/request_handler.py:23              httpx `json=` encoder (streaming)    -> RM_HTTPX028 / RM_HTTPX027
//   services/request_handler.py:153             json5.dumps (non-streaming)          -> RM_JSON5 (unpinned)
//
// A Python dict keeps insertion order, `payload[k] = v` overwrites in place or appends, and
// json.dumps walks the dict in order -- so for a body without duplicate keys the output is a
// token-by-token re-rendering of the input with (a) the values of assigned top-level keys replaced in
// place and (b) the not-yet-present assigned keys appended before the closing brace.  Bodies with
// duplicate keys (any level) are reported as LGW_BODY_EXOTIC instead of being guessed.
//
// Host/device portable: tests fuzz it against CPython through the g++ build (test aid).
#pragma once
#include <stdint.h>
#include "json_machine.cuh"

namespace lgw {

enum RenderMode : int { RM_HTTPX028 = 0, RM_HTTPX027 = 1, RM_JSON5 = 2 };
enum BodyStatus : uint32_t {
    BS_OK = 0,
    BS_PARSE_ERROR = 1,     // chat.py:37-39 -> HTTP 400 (not UTF-8, not JSON, not an object, no "model" key)
    BS_NO_MODEL = 2,        // chat.py:44-45 -> HTTP 400 (model falsy)
    BS_OVERFLOW = 3,        // output slot too small
    BS_EXOTIC = 4,          // duplicate keys, float with > 15 significant digits, key/number too long, nesting too deep ...
    BS_ENCODE_ERROR = 5     // the reference's encoder raises: NaN/Infinity with allow_nan=False, lone surrogate with ensure_ascii=False
};

struct BodyOp {             // == lgw_body_op
    uint32_t key_off, key_len;      // decoded key text (UTF-8) in the blob, for matching
    uint32_t rkey_off, rkey_len;    // rendered key token (quotes included when the mode quotes it)
    uint32_t rval_off, rval_len;    // rendered value
    uint32_t flags;                 // bit0: only if the key is absent from the client's body (chat.py:114)
                                    // bit1: presence probe -- nothing is assigned or appended, the key only sets its bit in `matched`
                                    //       (request_handler.py:167 `"error" in response_json`)
    uint32_t _pad;
};

#define LGW_BODY_KEYCAP 128
#define LGW_BODY_NUMCAP 48
#define LGW_BODY_MAXD 16
#define LGW_BODY_MAXKEYS 24

struct BodyScan {           // == lgw_body_scan: what chat.py:41-45 reads
    uint32_t status;        // BS_OK / BS_PARSE_ERROR / BS_NO_MODEL
    uint32_t model_len;     // decoded model string stored in the caller's buffer (when a string)
    uint8_t model_kind, model_truthy, stream_kind, stream_truthy;
    uint32_t _pad;
};

// Number token with a fraction or exponent -> the text Python's repr(float(token)) gives, when the token has
// <= 15 significant digits (then the shortest round-trip digits ARE the token's digits, David Gay's
// guarantee for doubles) and its magnitude is inside 1e-290..1e290.  Returns the length (<= 24) or -1
// ("exotic": the caller reports LGW_BODY_EXOTIC).  Format rule = CPython float_repr_style 'short', repr:
// exponent form iff decpt <= -4 or decpt > 16, at least two exponent digits.
LGW_HD int format_float(const char* t, uint32_t n, char* out) {
    uint32_t i = 0; int o = 0;
    if (n && t[0] == '-') { out[o++] = '-'; i = 1; }
    char dig[20]; int nd = 0; int decpt = 0; bool seen_nz = false, extra_nz = false, in_frac = false; int exp10 = 0;
    for (; i < n; ++i) {
        const char c = t[i];
        if (c == '.') { in_frac = true; continue; }
        if (c == 'e' || c == 'E') {
            ++i; bool eneg = false;
            if (i < n && (t[i] == '+' || t[i] == '-')) { eneg = t[i] == '-'; ++i; }
            int e = 0; for (; i < n; ++i) if (e < 100000) e = e * 10 + (t[i] - '0');
            exp10 = eneg ? -e : e;
            break;
        }
        if (c != '0' || seen_nz) {
            seen_nz = true;
            if (nd < 17) dig[nd++] = c; else if (c != '0') extra_nz = true;
            if (!in_frac) ++decpt;
        } else if (in_frac) --decpt;                 // zeros between the point and the first significant digit
    }
    while (nd > 0 && dig[nd - 1] == '0') --nd;
    if (nd == 0) { out[o++] = '0'; out[o++] = '.'; out[o++] = '0'; return o; }
    if (nd > 15 || extra_nz) return -1;
    decpt += exp10;
    if (decpt < -290 || decpt > 290) return -1;
    if (decpt > -4 && decpt <= 16) {
        if (decpt <= 0) { out[o++] = '0'; out[o++] = '.'; for (int k = 0; k < -decpt; ++k) out[o++] = '0'; for (int k = 0; k < nd; ++k) out[o++] = dig[k]; }
        else if (decpt < nd) { for (int k = 0; k < decpt; ++k) out[o++] = dig[k]; out[o++] = '.'; for (int k = decpt; k < nd; ++k) out[o++] = dig[k]; }
        else { for (int k = 0; k < nd; ++k) out[o++] = dig[k]; for (int k = nd; k < decpt; ++k) out[o++] = '0'; out[o++] = '.'; out[o++] = '0'; }
    } else {
        out[o++] = dig[0];
        if (nd > 1) { out[o++] = '.'; for (int k = 1; k < nd; ++k) out[o++] = dig[k]; }
        out[o++] = 'e';
        int e = decpt - 1;
        if (e < 0) { out[o++] = '-'; e = -e; } else out[o++] = '+';
        if (e >= 100) { out[o++] = (char)('0' + e / 100); e %= 100; }
        out[o++] = (char)('0' + e / 10); out[o++] = (char)('0' + e % 10);
    }
    return o;
}

struct BodyRewriter {
    // configuration
    int mode;
    const BodyOp* ops; uint32_t n_ops; const uint8_t* blob;
    uint8_t* out; uint32_t cap, len;
    // parser
    uint8_t st, depth, in_key, lit_id, lit_pos, ucount, neg_lit, skipping;
    uint8_t skip_depth, pending_replace;     // pending_replace: 1 + op index whose value replaces the next value
    uint32_t status;
    uint64_t stack;
    uint32_t ucode, pending_high;
    uint32_t matched;                        // bit i: op i's key is present in the client's body
    uint32_t top_members;
    // key buffer (decoded UTF-8)
    uint8_t kbuf[LGW_BODY_KEYCAP]; uint32_t klen; uint8_t k_has_surrogate;
    // number buffer (raw text)
    char nbuf[LGW_BODY_NUMCAP]; uint32_t nlen; uint8_t n_float;
    // duplicate-key detection
    uint32_t seen[LGW_BODY_MAXD][LGW_BODY_MAXKEYS]; uint8_t nseen[LGW_BODY_MAXD];
    // scan outputs (scan_body only)
    uint8_t want_scan; uint8_t cur_top_key;  // 1 model, 2 stream
    BodyScan* scan; uint8_t* model_buf; uint32_t model_cap;
    uint8_t str_nonempty, root_obj, response_mode, root_kind, probe_str;     // probe_str: this value string takes part in the `in` test of request_handler.py:167

    LGW_HD void fail(uint32_t s) { if (status == BS_OK) status = s; st = S_ERR; }
    LGW_HD void soft(uint32_t s) { if (status == BS_OK) status = s; }          // keep parsing, remember the verdict
    LGW_HD bool top_is_obj() const { return (stack >> (depth - 1)) & 1ull; }

    LGW_HD void emit(uint32_t b) {
        if (skipping) return;
#ifdef __CUDA_ARCH__
        if (len < cap && (threadIdx.x & 31u) == 0) out[len] = (uint8_t)b;      // the warp runs the machine redundantly; lane 0 writes
#else
        if (len < cap) out[len] = (uint8_t)b;
#endif
        ++len;
    }
    LGW_HD void emit_blob(uint32_t off, uint32_t n) { for (uint32_t i = 0; i < n; ++i) emit(blob[off + i]); }
    LGW_HD void emit_hex4(uint32_t cu) {
        emit('\\'); emit('u');
        for (int s = 12; s >= 0; s -= 4) { const uint32_t h = (cu >> s) & 15u; emit(h < 10 ? '0' + h : 'a' + h - 10); }
    }
    LGW_HD void emit_comma() { emit(','); if (mode != RM_HTTPX028) emit(' '); }
    LGW_HD void emit_colon() { emit(':'); if (mode != RM_HTTPX028) emit(' '); }

    // one decoded character of a string (key text goes to kbuf first and is rendered at the closing quote)
    LGW_HD void emit_char(uint32_t cp, bool lone_surrogate) {
        const bool ascii_only = mode != RM_HTTPX028;
        if (cp == '"') { emit('\\'); emit('"'); return; }
        if (cp == '\\') { emit('\\'); emit('\\'); return; }
        if (cp == '\n') { emit('\\'); emit('n'); return; }
        if (cp == '\r') { emit('\\'); emit('r'); return; }
        if (cp == '\t') { emit('\\'); emit('t'); return; }
        if (cp == 8) { emit('\\'); emit('b'); return; }
        if (cp == 12) { emit('\\'); emit('f'); return; }
        if (mode == RM_JSON5 && cp == 11) { emit('\\'); emit('v'); return; }
        if (mode == RM_JSON5 && cp == 0) { emit('\\'); emit('0'); return; }
        if (cp < 0x20) { emit_hex4(cp); return; }
        if (cp < 0x7f) { emit(cp); return; }
        if (cp == 0x7f) { if (ascii_only) emit_hex4(cp); else emit(cp); return; }
        if (lone_surrogate) { if (ascii_only) emit_hex4(cp); else if (!skipping) soft(BS_ENCODE_ERROR); return; }
        if (ascii_only) {
            if (cp >= 0x10000) { const uint32_t v = cp - 0x10000; emit_hex4(0xD800 + (v >> 10)); emit_hex4(0xDC00 + (v & 0x3FF)); }
            else emit_hex4(cp);
            return;
        }
        if (cp < 0x800) { emit(0xC0 | (cp >> 6)); emit(0x80 | (cp & 63)); }
        else if (cp < 0x10000) { emit(0xE0 | (cp >> 12)); emit(0x80 | ((cp >> 6) & 63)); emit(0x80 | (cp & 63)); }
        else { emit(0xF0 | (cp >> 18)); emit(0x80 | ((cp >> 12) & 63)); emit(0x80 | ((cp >> 6) & 63)); emit(0x80 | (cp & 63)); }
    }

    // ---- decoded characters arrive here -------------------------------------------------------
    LGW_HD void key_byte(uint32_t b) { if (klen < LGW_BODY_KEYCAP) kbuf[klen] = (uint8_t)b; ++klen; }
    LGW_HD void str_cp(uint32_t cp, bool lone) {
        str_nonempty = 1;
        if (in_key) {
            if (lone) { k_has_surrogate = 1; key_byte(0xED); key_byte(0xA0 | ((cp >> 6) & 31)); key_byte(0x80 | (cp & 63)); return; }   // CESU-style marker
            if (cp < 0x80) key_byte(cp);
            else if (cp < 0x800) { key_byte(0xC0 | (cp >> 6)); key_byte(0x80 | (cp & 63)); }
            else if (cp < 0x10000) { key_byte(0xE0 | (cp >> 12)); key_byte(0x80 | ((cp >> 6) & 63)); key_byte(0x80 | (cp & 63)); }
            else { key_byte(0xF0 | (cp >> 18)); key_byte(0x80 | ((cp >> 12) & 63)); key_byte(0x80 | ((cp >> 6) & 63)); key_byte(0x80 | (cp & 63)); }
            return;
        }
        if (want_scan && depth == 1 && cur_top_key == 1 && model_buf) {       // model text for the rule lookup (chat.py:41,48)
            uint8_t tmp[4]; uint32_t n = 0;
            if (cp < 0x80) tmp[n++] = (uint8_t)cp;
            else if (cp < 0x800) { tmp[n++] = 0xC0 | (cp >> 6); tmp[n++] = 0x80 | (cp & 63); }
            else if (cp < 0x10000) { tmp[n++] = 0xE0 | (cp >> 12); tmp[n++] = 0x80 | ((cp >> 6) & 63); tmp[n++] = 0x80 | (cp & 63); }
            else { tmp[n++] = 0xF0 | (cp >> 18); tmp[n++] = 0x80 | ((cp >> 12) & 63); tmp[n++] = 0x80 | ((cp >> 6) & 63); tmp[n++] = 0x80 | (cp & 63); }
            for (uint32_t i = 0; i < n; ++i) { if (scan->model_len < model_cap) model_buf[scan->model_len] = tmp[i]; ++scan->model_len; }
            if (lone) soft(BS_EXOTIC);
        }
        if (probe_str) {
            if (lone) { probe_byte(0xED); probe_byte(0xA0 | ((cp >> 6) & 31)); probe_byte(0x80 | (cp & 63)); }     // never equal to a probe's UTF-8
            else if (cp < 0x80) probe_byte(cp);
            else if (cp < 0x800) { probe_byte(0xC0 | (cp >> 6)); probe_byte(0x80 | (cp & 63)); }
            else if (cp < 0x10000) { probe_byte(0xE0 | (cp >> 12)); probe_byte(0x80 | ((cp >> 6) & 63)); probe_byte(0x80 | (cp & 63)); }
            else { probe_byte(0xF0 | (cp >> 18)); probe_byte(0x80 | ((cp >> 12) & 63)); probe_byte(0x80 | ((cp >> 6) & 63)); probe_byte(0x80 | (cp & 63)); }
        }
        emit_char(cp, lone);
    }
    LGW_HD void flush_high() { if (pending_high) { const uint32_t h = pending_high; pending_high = 0; str_cp(h, true); } }
    LGW_HD void escape_cp(uint32_t cp) {
        if (cp >= 0xD800 && cp <= 0xDBFF) { flush_high(); pending_high = cp; return; }
        if (cp >= 0xDC00 && cp <= 0xDFFF) {
            if (pending_high) { const uint32_t h = pending_high; pending_high = 0; str_cp(0x10000 + ((h - 0xD800) << 10) + (cp - 0xDC00), false); }
            else str_cp(cp, true);
            return;
        }
        flush_high(); str_cp(cp, false);
    }

    // ---- values -------------------------------------------------------------------------------
    LGW_HD void value_begins() {        // called once when any value token starts
        if (pending_replace && !skipping) {
            const BodyOp& op = ops[pending_replace - 1];
            emit_blob(op.rval_off, op.rval_len);
            skipping = 1; skip_depth = depth;
        }
        pending_replace = 0;
    }
    LGW_HD void value_ends(uint8_t kind, bool truthy) {        // depth already back at the value's level
        if (skipping && depth == skip_depth) skipping = 0;
        if (want_scan && depth == 1 && cur_top_key) {
            if (cur_top_key == 1) { scan->model_kind = kind; scan->model_truthy = truthy; }
            else { scan->stream_kind = kind; scan->stream_truthy = truthy; }
            cur_top_key = 0;
        }
        if (depth == 0) root_kind = kind;
        st = depth == 0 ? S_DONE : S_AFTER;
    }

    // request_handler.py:167 on a root that is not a dict: `"error" in <list>` asks whether an ELEMENT equals the
    // string, `"error" in <str>` whether it occurs as a substring.  The probed string's text is collected in kbuf
    // (elements) or kept as a sliding window of its last bytes (root string).
    LGW_HD void probe_byte(uint32_t b) {
        if (probe_str == 1) { key_byte(b); return; }                 // element of a root list: whole text (bounded; probes are short)
        // root string: does the text so far END with a probe?
        if (klen < 16) kbuf[klen++] = (uint8_t)b;
        else { for (uint32_t i = 0; i < 15; ++i) kbuf[i] = kbuf[i + 1]; kbuf[15] = (uint8_t)b; }
        for (uint32_t k = 0; k < n_ops; ++k) {
            const BodyOp& op = ops[k];
            if (!(op.flags & 2u) || op.key_len > klen || op.key_len == 0) continue;
            uint32_t j = 0;
            while (j < op.key_len && blob[op.key_off + j] == kbuf[klen - op.key_len + j]) ++j;
            if (j == op.key_len) matched |= 1u << k;
        }
    }
    LGW_HD void probe_end() {                                        // end of a list element: equal to a probe?
        if (probe_str == 1 && klen <= LGW_BODY_KEYCAP) {
            for (uint32_t k = 0; k < n_ops; ++k) {
                const BodyOp& op = ops[k];
                if (!(op.flags & 2u) || op.key_len != klen) continue;
                uint32_t j = 0;
                while (j < klen && blob[op.key_off + j] == kbuf[j]) ++j;
                if (j == klen) matched |= 1u << k;
            }
        }
        probe_str = 0;
    }

    LGW_HD uint32_t key_hash() const {
        uint32_t h = 2166136261u;
        const uint32_t n = klen < LGW_BODY_KEYCAP ? klen : LGW_BODY_KEYCAP;
        for (uint32_t i = 0; i < n; ++i) { h ^= kbuf[i]; h *= 16777619u; }
        return h ^ (klen * 0x9E3779B1u);
    }

    static LGW_HD bool is_reserved(const uint8_t* k, uint32_t n) {
        const char* words[] = {"break", "case", "catch", "continue", "debugger", "default", "delete", "do", "else", "finally", "for", "function",
                               "if", "in", "instanceof", "new", "return", "switch", "this", "throw", "try", "typeof", "var", "void", "while", "with",
                               "class", "const", "enum", "export", "extends", "import", "super", "null", "true", "false",
                               "implements", "interface", "let", "package", "private", "protected", "public", "static", "yield"};
        for (int w = 0; w < 45; ++w) {
            const char* s = words[w]; uint32_t i = 0;
            while (i < n && s[i] && (uint8_t)s[i] == k[i]) ++i;
            if (i == n && s[i] == 0) return true;
        }
        return false;
    }

    LGW_HD void end_key() {
        if (klen > LGW_BODY_KEYCAP) { soft(BS_EXOTIC); klen = LGW_BODY_KEYCAP; }
        // duplicate detection for this object (objects inside a value that is being replaced are dropped by the reference too)
        if (skipping) { /* nothing to check */ }
        else if (depth > LGW_BODY_MAXD) soft(BS_EXOTIC);
        else {
            const uint32_t h = key_hash();
            uint8_t& n = nseen[depth - 1];
            for (uint32_t i = 0; i < n; ++i) if (seen[depth - 1][i] == h) soft(BS_EXOTIC);
            if (n < LGW_BODY_MAXKEYS) seen[depth - 1][n++] = h; else soft(BS_EXOTIC);
        }
        uint32_t hit = 0;
        if (depth == 1) {
            ++top_members;
            for (uint32_t i = 0; i < n_ops; ++i) {
                const BodyOp& op = ops[i];
                if (op.key_len != klen) continue;
                uint32_t j = 0;
                while (j < klen && blob[op.key_off + j] == kbuf[j]) ++j;
                if (j == klen) { hit = i + 1; break; }
            }
            if (want_scan) {
                cur_top_key = 0;
                if (klen == 5 && kbuf[0] == 'm' && kbuf[1] == 'o' && kbuf[2] == 'd' && kbuf[3] == 'e' && kbuf[4] == 'l') { cur_top_key = 1; scan->model_len = 0; }
                else if (klen == 6 && kbuf[0] == 's' && kbuf[1] == 't' && kbuf[2] == 'r' && kbuf[3] == 'e' && kbuf[4] == 'a' && kbuf[5] == 'm') cur_top_key = 2;
            }
        }
        if (hit) {
            matched |= 1u << (hit - 1);
            if (!(ops[hit - 1].flags & 3u)) {            // assigned: rendered key, value replaced in place
                emit_blob(ops[hit - 1].rkey_off, ops[hit - 1].rkey_len);
                pending_replace = (uint8_t)hit;
                st = S_COLON;
                return;
            }
        }
        // render the client's own key
        bool unquoted = false;
        if (mode == RM_JSON5 && klen > 0 && !k_has_surrogate) {
            bool ident = true;
            for (uint32_t i = 0; i < klen; ++i) {
                const uint32_t c = kbuf[i];
                const bool alpha = (c | 0x20) - 'a' < 26u || c == '_' || c == '$';
                if (!(alpha || (i > 0 && c - '0' < 10u))) { ident = false; if (c >= 0x80 && !skipping) soft(BS_EXOTIC); break; }   // non-ASCII identifiers: unpinned
            }
            unquoted = ident && !is_reserved(kbuf, klen);
        }
        if (unquoted) { for (uint32_t i = 0; i < klen; ++i) emit(kbuf[i]); }
        else {
            emit('"');
            uint32_t i = 0;
            while (i < klen) {                       // decode kbuf (valid UTF-8 + surrogate markers) back to code points
                uint32_t c = kbuf[i], cp, n;
                if (c < 0x80) { cp = c; n = 1; }
                else if (c < 0xE0) { cp = ((c & 31) << 6) | (kbuf[i + 1] & 63); n = 2; }
                else if (c < 0xF0) { cp = ((c & 15) << 12) | ((kbuf[i + 1] & 63) << 6) | (kbuf[i + 2] & 63); n = 3; }
                else { cp = ((c & 7) << 18) | ((kbuf[i + 1] & 63) << 12) | ((kbuf[i + 2] & 63) << 6) | (kbuf[i + 3] & 63); n = 4; }
                emit_char(cp, cp >= 0xD800 && cp <= 0xDFFF);
                i += n;
            }
            emit('"');
        }
        st = S_COLON;
    }

    // float text -> repr(float(text)): see format_float()
    LGW_HD void emit_float() {
        char buf[32];
        const int k = format_float(nbuf, nlen, buf);
        if (k < 0) { if (nbuf[0] == '-') emit('-'); soft(BS_EXOTIC); return; }
        for (int j = 0; j < k; ++j) emit((uint8_t)buf[j]);
    }

    LGW_HD void end_number() {
        bool nonzero = false;
        for (uint32_t i = 0; i < nlen && i < LGW_BODY_NUMCAP; ++i) { const char c = nbuf[i]; if (c == 'e' || c == 'E') break; if (c >= '1' && c <= '9') nonzero = true; }
        if (nlen > LGW_BODY_NUMCAP) {
            if (!skipping) soft(BS_EXOTIC);
        } else if (!n_float) {
            if (nlen == 2 && nbuf[0] == '-' && nbuf[1] == '0') emit('0');        // int("-0") == 0
            else for (uint32_t i = 0; i < nlen; ++i) emit(nbuf[i]);
        } else if (!skipping) emit_float();       // a value that is being replaced is parsed but never rendered
        value_ends(n_float ? KD_FLT : KD_INT, nonzero);
    }
    LGW_HD void num_char(uint32_t c) { if (nlen < LGW_BODY_NUMCAP) nbuf[nlen] = (char)c; ++nlen; }

    LGW_HD void open_container(bool is_obj) {
        if (depth >= LGW_BODY_MAXD) { soft(BS_EXOTIC); if (depth >= 63) { fail(BS_EXOTIC); return; } }
        value_begins();
        if (depth == 0) root_obj = is_obj;
        emit(is_obj ? '{' : '[');
        if (is_obj) stack |= (1ull << depth); else stack &= ~(1ull << depth);
        ++depth;
        if (is_obj && depth <= LGW_BODY_MAXD) nseen[depth - 1] = 0;
        st = is_obj ? S_KEY_OR_END : S_VALUE_OR_END;
    }
    LGW_HD void close_container(bool is_obj, bool empty) {
        if (depth == 0 || top_is_obj() != is_obj) { fail(BS_PARSE_ERROR); return; }
        if (depth == 1 && is_obj) {                   // append the assigned keys the client did not send (chat.py dict order)
            bool any = top_members > 0;
            for (uint32_t i = 0; i < n_ops; ++i) {
                if ((matched & (1u << i)) || (ops[i].flags & 2u)) continue;
                if (any) emit_comma();
                emit_blob(ops[i].rkey_off, ops[i].rkey_len); emit_colon(); emit_blob(ops[i].rval_off, ops[i].rval_len);
                any = true;
            }
        }
        --depth;
        emit(is_obj ? '}' : ']');
        value_ends(is_obj ? KD_OBJ : KD_ARR, !empty);
    }

    LGW_HD void begin_value(uint32_t c) {
        if (c == '"') {
            value_begins(); emit('"'); in_key = 0; st = S_STR; pending_high = 0; str_nonempty = 0;
            probe_str = 0;
            if (response_mode && (depth == 0 || (depth == 1 && !root_obj))) { probe_str = depth == 0 ? 2 : 1; klen = 0; }
        }
        else if (c == '{') open_container(true);
        else if (c == '[') open_container(false);
        else if (c == '-') { value_begins(); nlen = 0; n_float = 0; num_char(c); st = S_NUM_MINUS; }
        else if (c == '0') { value_begins(); nlen = 0; n_float = 0; num_char(c); st = S_NUM_ZERO; }
        else if (c - '1' < 9u) { value_begins(); nlen = 0; n_float = 0; num_char(c); st = S_NUM_INT; }
        else if (c == 't') { value_begins(); st = S_LIT; lit_id = 0; lit_pos = 1; neg_lit = 0; }
        else if (c == 'f') { value_begins(); st = S_LIT; lit_id = 1; lit_pos = 1; neg_lit = 0; }
        else if (c == 'n') { value_begins(); st = S_LIT; lit_id = 2; lit_pos = 1; neg_lit = 0; }
        else if (c == 'N') { value_begins(); st = S_LIT; lit_id = 3; lit_pos = 1; neg_lit = 0; }
        else if (c == 'I') { value_begins(); st = S_LIT; lit_id = 4; lit_pos = 1; neg_lit = 0; }
        else fail(BS_PARSE_ERROR);
    }
    static LGW_HD const char* lit_text(int id) { return id == 0 ? "true" : id == 1 ? "false" : id == 2 ? "null" : id == 3 ? "NaN" : "Infinity"; }
    static LGW_HD bool is_ws(uint32_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }

    LGW_HD void feed(uint32_t c) {
        for (;;) {
            switch (st) {
            case S_STR:
                if (c == '"') {
                    flush_high();
                    if (in_key) { end_key(); return; }
                    if (probe_str) probe_end();
                    emit('"'); value_ends(KD_STR, str_nonempty != 0); return;
                }
                if (c == '\\') { st = S_STR_ESC; return; }
                if (c < 0x20) { fail(BS_PARSE_ERROR); return; }
                if (c < 0x80) { flush_high(); str_cp(c, false); return; }
                // raw UTF-8 lead byte (the body was validated as UTF-8): collect continuation bytes
                flush_high();
                if (c >= 0xF0) { ucode = c & 7; ucount = 3; } else if (c >= 0xE0) { ucode = c & 15; ucount = 2; } else { ucode = c & 31; ucount = 1; }
                st = S_TRAIL_B;          // UTF-8 continuation bytes of a raw character inside a string
                return;
            case S_TRAIL_B:
                ucode = (ucode << 6) | (c & 63);
                if (--ucount == 0) { st = S_STR; str_cp(ucode, false); }
                return;
            case S_NUM_E:
                n_float = 1;
                if (c == '+' || c == '-') { num_char(c); st = S_NUM_ESIGN; return; }
                if (c - '0' < 10u) { num_char(c); st = S_NUM_EXP; return; }
                fail(BS_PARSE_ERROR); return;
            case S_STR_ESC:
                st = S_STR;
                switch (c) {
                case '"': flush_high(); str_cp('"', false); break;   case '\\': flush_high(); str_cp('\\', false); break;
                case '/': flush_high(); str_cp('/', false); break;   case 'b': flush_high(); str_cp(8, false); break;
                case 'f': flush_high(); str_cp(12, false); break;    case 'n': flush_high(); str_cp(10, false); break;
                case 'r': flush_high(); str_cp(13, false); break;    case 't': flush_high(); str_cp(9, false); break;
                case 'u': st = S_STR_U; ucount = 0; ucode = 0; break;
                default: fail(BS_PARSE_ERROR);
                }
                return;
            case S_STR_U: {
                uint32_t d;
                if (c - '0' < 10u) d = c - '0';
                else if ((c | 0x20) - 'a' < 6u) d = (c | 0x20) - 'a' + 10;
                else { fail(BS_PARSE_ERROR); return; }
                ucode = (ucode << 4) | d;
                if (++ucount == 4) { st = S_STR; escape_cp(ucode); }
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
                if (c == '"') { in_key = 1; st = S_STR; klen = 0; k_has_surrogate = 0; pending_high = 0; return; }
                fail(BS_PARSE_ERROR); return;
            case S_KEY:
                if (is_ws(c)) return;
                if (c == '"') { in_key = 1; st = S_STR; klen = 0; k_has_surrogate = 0; pending_high = 0; return; }
                fail(BS_PARSE_ERROR); return;
            case S_COLON:
                if (is_ws(c)) return;
                if (c == ':') { emit_colon(); st = S_VALUE; return; }
                fail(BS_PARSE_ERROR); return;
            case S_AFTER:
                if (is_ws(c)) return;
                if (c == ',') { emit_comma(); st = top_is_obj() ? S_KEY : S_VALUE; return; }
                if (c == '}') { close_container(true, false); return; }
                if (c == ']') { close_container(false, false); return; }
                fail(BS_PARSE_ERROR); return;
            case S_NUM_MINUS:
                if (c == '0') { num_char(c); st = S_NUM_ZERO; return; }
                if (c - '1' < 9u) { num_char(c); st = S_NUM_INT; return; }
                if (c == 'I') { st = S_LIT; lit_id = 4; lit_pos = 1; neg_lit = 1; return; }
                fail(BS_PARSE_ERROR); return;
            case S_NUM_ZERO:
                if (c == '.') { num_char(c); st = S_NUM_DOT; return; }
                if ((c | 0x20) == 'e') { num_char(c); st = S_NUM_E; return; }
                end_number(); continue;
            case S_NUM_INT:
                if (c - '0' < 10u) { num_char(c); return; }
                if (c == '.') { num_char(c); st = S_NUM_DOT; return; }
                if ((c | 0x20) == 'e') { num_char(c); st = S_NUM_E; return; }
                end_number(); continue;
            case S_NUM_DOT:
                if (c - '0' < 10u) { n_float = 1; num_char(c); st = S_NUM_FRAC; return; }
                fail(BS_PARSE_ERROR); return;
            case S_NUM_FRAC:
                if (c - '0' < 10u) { num_char(c); return; }
                if ((c | 0x20) == 'e') { num_char(c); st = S_NUM_E; return; }
                end_number(); continue;
            case S_NUM_ESIGN:
                if (c - '0' < 10u) { num_char(c); st = S_NUM_EXP; return; }
                fail(BS_PARSE_ERROR); return;
            case S_NUM_EXP:
                if (c - '0' < 10u) { num_char(c); return; }
                end_number(); continue;
            case S_LIT: {
                const char* t = lit_text(lit_id);
                if ((uint32_t)(uint8_t)t[lit_pos] != c) { fail(BS_PARSE_ERROR); return; }
                ++lit_pos;
                if (t[lit_pos] == 0) {
                    if (lit_id >= 3) {
                        if (mode == RM_HTTPX028 && !skipping) soft(BS_ENCODE_ERROR);      // allow_nan=False
                        if (neg_lit) emit('-');
                    }
                    for (const char* p = t; *p; ++p) emit((uint8_t)*p);
                    value_ends(lit_id == 0 ? KD_TRUE : lit_id == 1 ? KD_FALSE : lit_id == 2 ? KD_NULL : KD_FLT, lit_id == 0 || lit_id >= 3);
                }
                return; }
            case S_DONE:
                if (is_ws(c)) return;
                fail(BS_PARSE_ERROR); return;
            default: return;
            }
        }
    }

    LGW_HD void reset(int render_mode, const BodyOp* o, uint32_t n, const uint8_t* b, uint8_t* dst, uint32_t dst_cap) {
        mode = render_mode; ops = o; n_ops = n; blob = b; out = dst; cap = dst_cap; len = 0;
        st = S_VALUE; depth = 0; in_key = 0; lit_id = 0; lit_pos = 0; ucount = 0; neg_lit = 0; skipping = 0; skip_depth = 0; pending_replace = 0;
        status = BS_OK; stack = 0; ucode = 0; pending_high = 0; matched = 0; top_members = 0; klen = 0; k_has_surrogate = 0; nlen = 0; n_float = 0;
        want_scan = 0; cur_top_key = 0; scan = nullptr; model_buf = nullptr; model_cap = 0; str_nonempty = 0; root_obj = 0; root_kind = KD_ABSENT; probe_str = 0; response_mode = (uint8_t)((render_mode >> 8) & 1); mode = render_mode & 0xff;
    }
};

// UTF-8 validity of the whole body (chat.py:32 `.decode('utf-8')`), then the machine
LGW_HD bool body_utf8_valid(const uint8_t* p, uint32_t n);

// Length (<= 32) of the run of bytes at in[i..] that are plain string content in `mode`: not a quote,
// backslash or control character, and -- when the mode escapes non-ASCII -- below 0x7f.  In
// RM_HTTPX028 raw UTF-8 (already validated) and DEL are copied as they are.
// On the device every lane of the warp that owns the body calls this with identical arguments.
LGW_HD uint32_t plain_run(const uint8_t* in, uint32_t i, uint32_t n, int mode) {
#ifdef __CUDA_ARCH__
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t c = i + lane < n ? in[i + lane] : (uint32_t)'"';
    const bool special = c == '"' || c == '\\' || c < 0x20 || (mode != RM_HTTPX028 && c >= 0x7f);
    const uint32_t mask = __ballot_sync(0xffffffffu, special);
    return mask ? (uint32_t)__ffs(mask) - 1u : 32u;
#else
    uint32_t r = 0;
    while (r < 32 && i + r < n) {
        const uint32_t c = in[i + r];
        if (c == '"' || c == '\\' || c < 0x20 || (mode != RM_HTTPX028 && c >= 0x7f)) break;
        ++r;
    }
    return r;
#endif
}
LGW_HD void copy_run(uint8_t* out, uint32_t cap, uint32_t at, const uint8_t* in, uint32_t i, uint32_t r) {
#ifdef __CUDA_ARCH__
    const uint32_t lane = threadIdx.x & 31u;
    if (lane < r && at + lane < cap) out[at + lane] = in[i + lane];
#else
    for (uint32_t k = 0; k < r; ++k) if (at + k < cap) out[at + k] = in[i + k];
#endif
}

// Rewrite one body.  Returns the status; *out_len = bytes the output needs (may exceed cap -> BS_OVERFLOW).
// (m.reset() done by the caller; known_ascii lets the device skip the sequential UTF-8 walk)
LGW_HD_NOINLINE uint32_t rewrite_body_checked(BodyRewriter& m, const uint8_t* in, uint32_t n, bool known_ascii, uint32_t* out_len) {
    const int mode = m.mode;
    uint8_t* const out = m.out;
    const uint32_t cap = m.cap;
    *out_len = 0;
    if (!known_ascii && !body_utf8_valid(in, n)) return BS_PARSE_ERROR;
    uint32_t i = 0;
    while (i < n && m.st != S_ERR) {
        if (m.st == S_STR && !m.in_key && !m.pending_high && !m.probe_str) {
            // bulk path for string content: a run of characters that every mode renders verbatim
            const uint32_t r = plain_run(in, i, n, mode);
            if (r) {
                if (!m.skipping) { copy_run(out, cap, m.len, in, i, r); m.len += r; }
                m.str_nonempty = 1;
                i += r;
                continue;
            }
        }
        m.feed(in[i]); ++i;
    }
    if (m.st != S_ERR) m.feed(' ');                   // a number at the very end of the document (root scalar) ends here
    if (m.st == S_ERR) return m.status ? m.status : BS_PARSE_ERROR;
    if (m.st != S_DONE) return BS_PARSE_ERROR;
    *out_len = m.len;
    if (m.status) return m.status;
    if (m.len > cap) return BS_OVERFLOW;
    return BS_OK;
}

LGW_HD uint32_t rewrite_body(BodyRewriter& m, const uint8_t* in, uint32_t n, int mode, const BodyOp* ops, uint32_t n_ops,
                             const uint8_t* blob, uint8_t* out, uint32_t cap, uint32_t* out_len) {
    m.reset(mode, ops, n_ops, blob, out, cap);
    return rewrite_body_checked(m, in, n, false, out_len);
}

// chat.py:31-45 for one body: validity, model (text when a string) and stream truthiness
LGW_HD_NOINLINE void scan_body(BodyRewriter& m, const uint8_t* in, uint32_t n, BodyScan* sc, uint8_t* model_buf, uint32_t model_cap) {
    m.reset(RM_HTTPX027, nullptr, 0, nullptr, nullptr, 0);
    m.skipping = 1; m.skip_depth = 0xff;             // never emit
    m.want_scan = 1; m.scan = sc; m.model_buf = model_buf; m.model_cap = model_cap;
    sc->status = BS_PARSE_ERROR; sc->_pad = 0; sc->model_len = 0; sc->model_kind = KD_ABSENT; sc->model_truthy = 0; sc->stream_kind = KD_ABSENT; sc->stream_truthy = 0;
    if (!body_utf8_valid(in, n)) return;
    for (uint32_t i = 0; i < n && m.st != S_ERR; ++i) m.feed(in[i]);
    if (m.st != S_DONE) return;
    // a non-object body makes chat.py:35 raise -> 400; so does a missing "model" key (:36)
    if (!m.root_obj || sc->model_kind == KD_ABSENT) return;
    sc->status = sc->model_truthy ? BS_OK : BS_NO_MODEL;
}

LGW_HD bool body_utf8_valid(const uint8_t* p, uint32_t n) {
    uint32_t i = 0;
    while (i < n) {
        const uint32_t c = p[i];
        if (c < 0x80) { ++i; continue; }
        uint32_t need, lo = 0x80, hi = 0xBF;
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

}  // namespace lgw
