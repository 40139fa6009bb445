// Bulk path of the SSE step: k_prime -> k_relay -> k_commit.   (included from sse_kernels.cuh)
//
// Regular streams -- committed (or committing on their first chunk), carries empty and equal at the
// start of the bulk region, every chunk valid UTF-8 and non-empty, no run of three or more LFs, no
// event the tap turns into an extra row ("error"), at most one usage event per step -- are exactly
// the streams for which the reference's chunk-by-chunk split (request_handler.py:111-115,
// chat_logging.py:108-112) equals a split of the CONCATENATED text on every LF LF pair.  For those,
// events are independent: the chunk in which an event completes owns it, recognises it, and posts its
// findings with atomics.  Everything else is flagged irregular and redone by k_commit with the exact
// sequential machine, so the result never depends on which path ran (tests run both and the oracle).
//
// Event templates.  SSE deltas of a stream are the same JSON skeleton over and over with a few values
// changing.  Each persistent block keeps TWO validated events as templates (bytes + for every byte position
// the id of the string or number VALUE span it lies in), shared engine-wide through TemplateCache.  An event
// that is byte-identical to a template outside value spans, holds plain string bytes / valid escapes up to the
// closing quote inside string spans and a valid JSON number inside number spans drives the recogniser through
// the same states as the template: same validity, same top-level keys -- no byte of it needs to be
// re-validated beyond the comparison (match_window for events inside the staged window, match_template
// anywhere).  Anything else takes the byte-wise recogniser.

#ifndef LGW_RELAY_THREADS
#define LGW_RELAY_THREADS 64
#endif
#ifndef LGW_RELAY_BLOCKS_PER_SM
#define LGW_RELAY_BLOCKS_PER_SM 16
#endif
#define LGW_TILE_VECS (LGW_TILE_BYTES / 16)
#define LGW_HALO_BYTES 512u          /* read-only look-ahead into the next tile: events that straddle the tile end */
#define LGW_STAGE_BYTES (LGW_TILE_BYTES + LGW_HALO_BYTES)
#define LGW_TPL_MAX 504u            /* longest event kept as a template */
#define LGW_TPL_IDS 32u

// Tile in shared memory, one pad word after every 64-byte row: 32 lanes reading the same byte position of 32
// consecutive 64-byte rows (the 64-byte-event pattern) then hit 32 different banks, and the address of a byte
// is two instructions away from its offset:  physical offset = d + 4 * (d / 64).
__device__ __forceinline__ uint32_t phys(uint32_t d) { return d + ((d >> 6) << 2); }
#define LGW_TILE_PHYS_BYTES (LGW_STAGE_BYTES + LGW_STAGE_BYTES / 16 + 16)
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }

__shared__ __align__(16) uint8_t sh_tile[LGW_TILE_PHYS_BYTES];
__shared__ __align__(4) uint8_t sh_cls[256];
__shared__ __align__(4) uint8_t sh_trans[LGW_LEAN_ROWS * 32];
__shared__ uint32_t sh_seg_lo, sh_seg_hi;
#define LGW_TC_STAGED 16u
__shared__ uint32_t sh_tile_chunk[LGW_TC_STAGED];
#ifdef LGW_DEBUG_TIMING
__device__ unsigned long long g_dbg[16];
#define DBG_STAMP(i) do { if (blockIdx.x == 7 && threadIdx.x == 0 && tile == tile_first + 3) g_dbg[i] = clock64(); } while (0)
#else
#define DBG_STAMP(i) do {} while (0)
#endif
// the block's event templates (two slots): validated events kept as skeletons
#define LGW_TPL_SLOTS 2u
#define LGW_TPL_STRIDE (LGW_TPL_MAX + 16u)
#define LGW_TPL_MAPSTRIDE (LGW_TPL_MAX + 8u)
__shared__ __align__(4) uint8_t sh_tpl_bytes[LGW_TPL_SLOTS * LGW_TPL_STRIDE];      // zero padded text
__shared__ __align__(4) uint8_t sh_tpl_strid[LGW_TPL_SLOTS * LGW_TPL_MAPSTRIDE];   // value-span id per position (0xff: literal)
__shared__ uint16_t sh_tpl_sstart[LGW_TPL_SLOTS][LGW_TPL_IDS];                     // span start (numbers: first char)
__shared__ uint16_t sh_tpl_send[LGW_TPL_SLOTS][LGW_TPL_IDS];                       // strings: closing quote; numbers: terminator
__shared__ uint8_t sh_tpl_skind[LGW_TPL_SLOTS][LGW_TPL_IDS];                       // 0 string value, 1 number value
__shared__ uint32_t sh_tpl_len[LGW_TPL_SLOTS], sh_tpl_flags[LGW_TPL_SLOTS], sh_tpl_cls[LGW_TPL_SLOTS], sh_tpl_valid[LGW_TPL_SLOTS];
__shared__ uint32_t sh_tpl_cand;                  // slot 0: lowest thread with a candidate event in this tile
__shared__ unsigned long long sh_tpl_cand1;             // slot 1: (position << 32 | segment end) of the first event that missed slot 0 (one word: the pair must belong together)
__shared__ uint32_t sh_tpl0_canon;                          // this block's slot 0 IS the engine-wide cache's slot 0
__shared__ uint32_t sh_tpl_tries1;                        // failed attempts to build slot 1 (give up after a few)
__shared__ uint32_t sh_tpl_miss, sh_tpl_replace1, sh_tpl_cstate[LGW_TPL_SLOTS];   // misses in the current tile; slot 1 was re-learnt; cache states
// block-staged copy of a second-template candidate (its tile is gone; one thread walking global
// memory byte by byte would stall the whole block)
__shared__ __align__(16) uint8_t sh_stage[LGW_TPL_STRIDE + 16];

// shared-memory loads by 32-bit shared-window address held in a register (the compiler otherwise
// rebuilds the CTA's shared-window base -- S2UR SR_CgaCtaId + ULEA -- in front of every access)
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) { uint32_t v; asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
// 16-byte read-only load that does not allocate in L1: the tile is staged in shared memory anyway, and with 225 KB of
// shared memory per SM the L1 that is left is a few KB -- it should keep the chunk offsets and segment plans that were
// prefetched for the walk, not the bytes streaming through
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
__device__ __forceinline__ uint32_t opaque(uint32_t x) { asm volatile("" : "+r"(x)); return x; }

// byte source + tables of the bulk kernel: the staged tile, global memory outside it
struct TileEnv {
    uint32_t tile_s, cls_s, trans_s, tpl_s, strid_s;   // shared-window addresses (opaque registers)
    const uint8_t* __restrict__ g;     // whole packed buffer
    uint32_t t0, n_bytes;
    // aligned 32-bit word containing byte `pos` (little endian); bytes past n_bytes read as 0
    __device__ __forceinline__ uint32_t word(uint32_t pos) const {
        const uint32_t p4 = pos & ~3u, d = p4 - t0;
        if (d < LGW_STAGE_BYTES) return lds_u32(tile_s + phys(d));
        return word_global(p4);
    }
    __device__ __noinline__ uint32_t word_global(uint32_t p4) const {
        if (p4 >= n_bytes) return 0;
        if (p4 + 4 <= n_bytes) return __ldg(reinterpret_cast<const uint32_t*>(g + p4));
        uint32_t w = 0;
        for (uint32_t k = 0; k < 4 && p4 + k < n_bytes; ++k) w |= (uint32_t)__ldg(g + p4 + k) << (8 * k);
        return w;
    }
    // the four bytes at pos .. pos+3 (any alignment)
    __device__ __forceinline__ uint32_t wordu(uint32_t pos) const {
        const uint32_t lo = word(pos);
        if ((pos & 3u) == 0) return lo;
        return __funnelshift_r(lo, word(pos + 4), 8 * (pos & 3u));
    }
    __device__ __forceinline__ uint32_t at(uint32_t pos) const { return (word(pos) >> (8 * (pos & 3u))) & 0xffu; }
    __device__ __forceinline__ uint32_t cls(uint32_t c) const { return lds_u8(cls_s + c); }
    __device__ __forceinline__ uint32_t trans(uint32_t i) const { return lds_u8(trans_s + i); }
    // template text (shared address tb): the four bytes at k .. k+3
    __device__ __forceinline__ uint32_t tplu(uint32_t tb, uint32_t k) const {
        const uint32_t lo = lds_u32(tb + (k & ~3u));
        if ((k & 3u) == 0) return lo;
        return __funnelshift_r(lo, lds_u32(tb + (k & ~3u) + 4), 8 * (k & 3u));
    }
};

// the same reader over global memory only (k_relay_long: chunks that were left out of the tile walk)
struct GlobalEnv {
    uint32_t cls_s, trans_s, tpl_s, strid_s;
    const uint8_t* __restrict__ g;
    uint32_t n_bytes;
    __device__ __forceinline__ uint32_t word(uint32_t pos) const {
        const uint32_t p4 = pos & ~3u;
        if (p4 + 4 <= n_bytes) return __ldg(reinterpret_cast<const uint32_t*>(g + p4));
        uint32_t w = 0;
        for (uint32_t k = 0; k < 4 && p4 + k < n_bytes; ++k) w |= (uint32_t)__ldg(g + p4 + k) << (8 * k);
        return w;
    }
    __device__ __forceinline__ uint32_t wordu(uint32_t pos) const {
        const uint32_t lo = word(pos);
        if ((pos & 3u) == 0) return lo;
        return __funnelshift_r(lo, word(pos + 4), 8 * (pos & 3u));
    }
    __device__ __forceinline__ uint32_t at(uint32_t pos) const { return (word(pos) >> (8 * (pos & 3u))) & 0xffu; }
    __device__ __forceinline__ uint32_t cls(uint32_t c) const { return lds_u8(cls_s + c); }
    __device__ __forceinline__ uint32_t trans(uint32_t i) const { return lds_u8(trans_s + i); }
    __device__ __forceinline__ uint32_t tplu(uint32_t tb, uint32_t k) const {
        const uint32_t lo = lds_u32(tb + (k & ~3u));
        if ((k & 3u) == 0) return lo;
        return __funnelshift_r(lo, lds_u32(tb + (k & ~3u) + 4), 8 * (k & 3u));
    }
};

// first byte of w (little endian) that is '"', '\\' or < 0x20: index 0..3, or 4 when none
__device__ __forceinline__ uint32_t first_special(uint32_t w) {
    const uint32_t q = w ^ 0x22222222u, b = w ^ 0x5c5c5c5cu, c = w & 0xe0e0e0e0u;
    const uint32_t m = (((q - 0x01010101u) & ~q) | ((b - 0x01010101u) & ~b) | ((c - 0x01010101u) & ~c)) & 0x80808080u;
    return m ? (uint32_t)(__ffs(m) - 1) >> 3 : 4u;      // the lowest flagged byte is exact (borrows only travel upward)
}

// length of the valid escape sequence whose backslash sits at `ib` (2, or 6 for \\uXXXX), 0 when the recogniser would
// reject it (lean_json.cuh L_STR_ESC / L_STR_U)
template <class ENV>
__device__ __noinline__ uint32_t escape_length(const ENV& env, uint32_t ib) {
    const uint32_t w = env.wordu(ib + 1);
    const uint32_t c = w & 0xffu;
    if (c == 'u') {
        const uint32_t h = (w >> 8) | (env.at(ib + 5) << 24);     // the four hex digits
        for (uint32_t j = 0; j < 4; ++j) { const uint32_t x = (h >> (8 * j)) & 0xffu; if (!(x - '0' < 10u || (x | 0x20u) - 'a' < 6u)) return 0; }
        return 6;
    }
    return (c == '"' || c == '\\' || c == '/' || c == 'b' || c == 'f' || c == 'n' || c == 'r' || c == 't') ? 2u : 0u;
}

// Does the event starting at ps follow template `slot`?  The event may differ from the template in
// any number of VALUE spans: inside a string value it may hold any plain bytes and valid escapes (no quote, bare backslash
// or control byte) up to its closing quote; a number value may be any valid JSON number (checked
// with the number rows of the recogniser's table).  Everything else must be byte-identical, so the
// recogniser would walk the same states: same validity, same top-level keys.
// On success *end = position of the first LF of the event's LF LF separator (verified to be there).
template <class ENV>
__device__ __forceinline__ bool match_template(const ENV& env, uint32_t slot, uint32_t ps, uint32_t* end) {
    const uint32_t lenA = sh_tpl_len[slot];
    const uint32_t tb = env.tpl_s + slot * LGW_TPL_STRIDE, sb = env.strid_s + slot * LGW_TPL_MAPSTRIDE;
    uint32_t ia = 0, ib = ps;
    bool fresh = false;                      // a span was just skipped: the next bytes must agree
    for (;;) {
        // equal run: both texts advance a word at a time, so their alignments stay fixed; keep the
        // aligned words rolling (one load per side per step) and funnel-shift them into place
        // (a funnel shift by 0 returns the low word)
        if (ia < lenA) {
            const uint32_t shB = 8 * (ib & 3u), shA = 8 * (ia & 3u);
            uint32_t pb = ib & ~3u, pa = tb + (ia & ~3u);
            uint32_t loB = env.word(pb), loA = lds_u32(pa);
            uint32_t left = lenA - ia, x = 0;
            while (left > 4) {
                const uint32_t hiB = env.word(pb + 4), hiA = lds_u32(pa + 4);
                x = __funnelshift_r(loB, hiB, shB) ^ __funnelshift_r(loA, hiA, shA);
                if (x) break;
                pb += 4; pa += 4; loB = hiB; loA = hiA; left -= 4;
            }
            if (!x) {                                   // the last (possibly partial) word
                const uint32_t hiB = env.word(pb + 4), hiA = lds_u32(pa + 4);
                x = __funnelshift_r(loB, hiB, shB) ^ __funnelshift_r(loA, hiA, shA);
                if (left < 4) x &= (1u << (8 * left)) - 1u;
            }
            uint32_t adv = (lenA - ia) - left;          // bytes matched in whole words
            adv += x ? ((uint32_t)(__ffs(x) - 1) >> 3) : left;
            if (adv == 0 && fresh) return false;
            ia += adv; ib += adv;
            fresh = false;
        }
        if (ia >= lenA) break;
        const uint32_t id = lds_u8(sb + ia);
        if (id == 0xffu) return false;       // the texts part at a literal position
        if (sh_tpl_skind[slot][id] == 0) {   // string value: plain bytes and valid escapes up to the closing quote
            for (;;) {
                const uint32_t w4 = env.wordu(ib);
                const uint32_t k = first_special(w4);
                ib += k;
                if (k == 4) { if (ib - ps > 8192u) return false; continue; }
                const uint32_t sp = (w4 >> (8 * k)) & 0xffu;            // the special byte itself: no second load
                if (sp == '"') break;                                   // the closing quote
                if (sp != '\\') return false;                           // a control byte
                const uint32_t adv = escape_length(env, ib);            // (out of line: keeps the scan loop tight)
                if (adv == 0) return false;
                ib += adv;
            }
        } else {                             // number value: re-validate the event's own number
            uint32_t bs = ib - (ia - sh_tpl_sstart[slot][id]);
            uint32_t st = L_VALUE;
            {   // common case first: a plain run of digits ("0" or [1-9][0-9]*)
                uint32_t q = bs, c = env.at(q);
                if (c - '1' < 9u) { do { c = env.at(++q); } while (c - '0' < 10u); }
                else if (c == '0') c = env.at(++q);
                else q = bs;
                if (q != bs && c != '.' && c != 'e' && c != 'E') { ib = q; ia = sh_tpl_send[slot][id]; fresh = true; continue; }
            }
            for (;;) {
                const uint32_t cl = env.cls(env.at(bs));
                if (cl < C_MINUS || cl > C_EXP) break;
                st = env.trans(st * 32 + cl) & 31u;
                if (st == L_ERR || bs - ps > 8192u) return false;
                ++bs;
            }
            if (!(st == L_NUM_ZERO || st == L_NUM_INT || st == L_NUM_FRAC || st == L_NUM_EXP)) return false;
            ib = bs;
        }
        ia = sh_tpl_send[slot][id];
        fresh = true;
    }
    if ((env.wordu(ib) & 0xffffu) != 0x0a0au) return false;     // LF LF must follow
    *end = ib;
    return true;
}

// ---- window matcher ------------------------------------------------------------------------------------
// The same judgement as match_template for an event that lies inside the staged window [t0, t0 + LGW_STAGE_BYTES):
// positions are offsets into the window, every load is a plain shared-memory load (no "is it staged?" test and no
// out-of-line global fallback in front of each word).  Returns 1 match, 0 no match, 2 "left the window" (the caller then
// asks match_template, which reads anywhere).
#ifndef LGW_WINDOW_MATCH
#define LGW_WINDOW_MATCH 1
#endif
#ifndef LGW_EQ_UNROLL
#define LGW_EQ_UNROLL 1
#endif
// Experimental (off): chunks much longer than their neighbours (the 300-byte usage chunk among 64-byte deltas) are queued by
// k_relay and walked by k_relay_long, one warp per chunk, instead of by one lane whose warp and block wait for it.
// Measured on B200: k_relay 146 -> 120 us, but k_relay_long costs 32 us (thread per chunk: 43 us), so the
// step is slower for now.  Parity is the same either way (tests/test_sse_gpu.py passes with -DLGW_DEFER_LONG=1).
#ifndef LGW_DEFER_LONG
#define LGW_DEFER_LONG 0
#endif
// first byte of w that is not an ASCII digit: index 0..3, or 4 when all four are digits
__device__ __forceinline__ uint32_t first_nondigit(uint32_t w) {
    const uint32_t t = w ^ 0x30303030u;                                         // digits become 0x00..0x09
    const uint32_t m = (((t & 0x7f7f7f7fu) + 0x76767676u) | t) & 0x80808080u;   // byte flagged <=> t >= 0x0a (no carries between bytes)
    return m ? (uint32_t)(__ffs(m) - 1) >> 3 : 4u;
}
__device__ __forceinline__ uint32_t match_window(const TileEnv& env, uint32_t slot, uint32_t ps, uint32_t* end) {
    const uint32_t lenA = sh_tpl_len[slot];
    const uint32_t tb = env.tpl_s + slot * LGW_TPL_STRIDE, sb = env.strid_s + slot * LGW_TPL_MAPSTRIDE;
    const uint32_t ts = env.tile_s;
#define LGW_WW(d4) lds_u32(ts + phys(d4))
    uint32_t ia = 0, db = ps - env.t0;                 // db: offset of the event's current byte in the window
    if (db >= LGW_STAGE_BYTES - 16u) return 2;
    bool fresh = false;
    for (;;) {
        if (ia < lenA) {
            uint32_t left = lenA - ia, x = 0;
            if (db + left + 12u > LGW_STAGE_BYTES) return 2;
            const uint32_t shB = 8 * (db & 3u), shA = 8 * (ia & 3u);
            uint32_t qb = db & ~3u, pa = tb + (ia & ~3u);
            uint32_t loB = LGW_WW(qb), loA = lds_u32(pa);
            const uint32_t left0 = left;
#if LGW_EQ_UNROLL
            while (left > 8) {                          // two words per round: four independent loads, one branch
                const uint32_t m1B = LGW_WW(qb + 4), m2B = LGW_WW(qb + 8), m1A = lds_u32(pa + 4), m2A = lds_u32(pa + 8);
                const uint32_t x1 = __funnelshift_r(loB, m1B, shB) ^ __funnelshift_r(loA, m1A, shA);
                const uint32_t x2 = __funnelshift_r(m1B, m2B, shB) ^ __funnelshift_r(m1A, m2A, shA);
                if (x1 | x2) {
                    if (x1) x = x1; else { x = x2; qb += 4; pa += 4; left -= 4; loB = m1B; loA = m1A; }
                    break;
                }
                qb += 8; pa += 8; left -= 8; loB = m2B; loA = m2A;
            }
            while (!x && left > 4) {
#else
            while (left > 4) {
#endif
                const uint32_t hiB = LGW_WW(qb + 4), hiA = lds_u32(pa + 4);
                x = __funnelshift_r(loB, hiB, shB) ^ __funnelshift_r(loA, hiA, shA);
                if (x) break;
                qb += 4; pa += 4; loB = hiB; loA = hiA; left -= 4;
            }
            if (!x) {
                const uint32_t hiB = LGW_WW(qb + 4), hiA = lds_u32(pa + 4);
                x = __funnelshift_r(loB, hiB, shB) ^ __funnelshift_r(loA, hiA, shA);
                if (left < 4) x &= (1u << (8 * left)) - 1u;
            }
            uint32_t adv = left0 - left;
            adv += x ? ((uint32_t)(__ffs(x) - 1) >> 3) : left;
            if (adv == 0 && fresh) return 0;
            ia += adv; db += adv;
            fresh = false;
        }
        if (ia >= lenA) break;
        const uint32_t id = lds_u8(sb + ia);
        if (id == 0xffu) return 0;
        if (sh_tpl_skind[slot][id] == 0) {   // string value
            for (;;) {
                uint32_t q = db & ~3u;
                const uint32_t sh = 8 * (db & 3u);
                if (q + 12u > LGW_STAGE_BYTES) return 2;
                uint32_t lo = LGW_WW(q), hi = LGW_WW(q + 4), w4, k;
                for (;;) {
                    w4 = __funnelshift_r(lo, hi, sh);
                    k = first_special(w4);
                    if (k != 4) break;
                    q += 4;
                    if (q + 12u > LGW_STAGE_BYTES) return 2;
                    lo = hi; hi = LGW_WW(q + 4);
                }
                db = q + (sh >> 3) + k;
                const uint32_t sp = (w4 >> (8 * k)) & 0xffu;
                if (sp == '"') break;
                if (sp != '\\') return 0;
                const uint32_t adv = escape_length(env, env.t0 + db);
                if (adv == 0) return 0;
                db += adv;
            }
        } else {                             // number value: re-validate the event's own number
            const uint32_t bs = db - (ia - sh_tpl_sstart[slot][id]);
            uint32_t q = bs & ~3u;
            const uint32_t sh = 8 * (bs & 3u);
            if (q + 12u > LGW_STAGE_BYTES) return 2;
            uint32_t lo = LGW_WW(q), hi = LGW_WW(q + 4);
            uint32_t w4 = __funnelshift_r(lo, hi, sh);
            const uint32_t c0 = w4 & 0xffu;
            uint32_t pe = bs;                                     // first byte after the number
            bool plain = false;                                   // common case: "0" or [1-9][0-9]*, four digits per step
            if (c0 == '0') { pe = bs + 1; plain = true; }
            else if (c0 - '1' < 9u) {
                for (;;) {
                    const uint32_t k = first_nondigit(w4);
                    pe += k;
                    if (k != 4) break;
                    q += 4;
                    if (q + 12u > LGW_STAGE_BYTES) return 2;
                    lo = hi; hi = LGW_WW(q + 4);
                    w4 = __funnelshift_r(lo, hi, sh);
                }
                plain = true;
            }
            if (plain) {
                const uint32_t c = (LGW_WW(pe & ~3u) >> (8 * (pe & 3u))) & 0xffu;
                if (c == '.' || c == 'e' || c == 'E') plain = false;
            }
            if (!plain) {                                         // sign, fraction, exponent: the recogniser's number rows
                uint32_t st = L_VALUE;
                pe = bs;
                for (;;) {
                    if (pe + 12u > LGW_STAGE_BYTES) return 2;
                    const uint32_t cl = env.cls((LGW_WW(pe & ~3u) >> (8 * (pe & 3u))) & 0xffu);
                    if (cl < C_MINUS || cl > C_EXP) break;
                    st = env.trans(st * 32 + cl) & 31u;
                    if (st == L_ERR) return 0;
                    ++pe;
                }
                if (!(st == L_NUM_ZERO || st == L_NUM_INT || st == L_NUM_FRAC || st == L_NUM_EXP)) return 0;
            }
            db = pe;
        }
        ia = sh_tpl_send[slot][id];
        fresh = true;
    }
    {   // LF LF must follow -- and no third LF (a run of three is the sequential path's business; the caller need not look again)
        const uint32_t q = db & ~3u;
        if (q + 12u > LGW_STAGE_BYTES) return 2;
        const uint32_t w = __funnelshift_r(LGW_WW(q), LGW_WW(q + 4), 8 * (db & 3u));
        if ((w & 0xffffu) != 0x0a0au || ((w >> 16) & 0xffu) == '\n') return 0;
    }
#undef LGW_WW
    *end = env.t0 + db;
    return 1;
}
__device__ __noinline__ bool match_template_far(const TileEnv& env, uint32_t slot, uint32_t ps, uint32_t* end) { return match_template(env, slot, ps, end); }

// reader over the block-staged copy [base, base + LGW_TPL_STRIDE)
struct StageEnv {
    uint32_t base, cls_s, trans_s;
    __device__ __forceinline__ uint32_t at(uint32_t pos) const { const uint32_t d = pos - base; return d < LGW_TPL_STRIDE ? (uint32_t)sh_stage[d] : 0u; }
    __device__ __forceinline__ uint32_t cls(uint32_t c) const { return lds_u8(cls_s + c); }
    __device__ __forceinline__ uint32_t trans(uint32_t i) const { return lds_u8(trans_s + i); }
};

// One thread validates the event at ps and installs it as template `slot` (or leaves the slot
// invalid).  limit = end of the segment.
template <class ENV>
__device__ __noinline__ void build_template(const ENV* env, uint32_t slot, uint32_t ps, uint32_t limit) {
    uint32_t cls = PC_NONE;
    const uint32_t c0 = env->at(ps);
    if (c0 == '{') cls = PC_BRACE;
    else if (c0 == 'd' && env->at(ps + 1) == 'a' && env->at(ps + 2) == 't' && env->at(ps + 3) == 'a' && env->at(ps + 4) == ':' && env->at(ps + 5) == ' ' && env->at(ps + 6) == '{') cls = PC_DATA;
    if (cls == PC_NONE) return;
    uint8_t* map = sh_tpl_strid + slot * LGW_TPL_MAPSTRIDE;
    LeanMachine lm;
    lm.reset(cls == PC_DATA);
    const uint32_t skip = cls == PC_DATA ? 6u : 0u;
    for (uint32_t k = 0; k < skip; ++k) map[k] = 0xff;
    uint32_t pos = ps + skip, n_ids = 0, cur_s = 0xff, cur_n = 0xff;
    for (;;) {
        if (pos + 1 >= limit || pos - ps >= LGW_TPL_MAX) return;
        const uint32_t c = env->at(pos);
        if (c == '\n' && env->at(pos + 1) == '\n') break;
        const uint32_t prev = lm.st;
        const bool in_val = prev == L_STR && !lm.in_key;
        const bool in_num = prev >= L_NUM_MINUS && prev <= L_NUM_EXP;
        map[pos - ps] = in_val ? (uint8_t)cur_s : in_num ? (uint8_t)cur_n : (uint8_t)0xff;
        if (in_val && c == '"' && cur_s != 0xff) sh_tpl_send[slot][cur_s] = (uint16_t)(pos - ps);
        lm.step(c, pos, *env);
        const bool now_num = lm.st >= L_NUM_MINUS && lm.st <= L_NUM_EXP;
        if (in_num && !now_num && cur_n != 0xff) sh_tpl_send[slot][cur_n] = (uint16_t)(pos - ps);      // the terminator
        if (prev < LGW_LEAN_ROWS && lm.st == L_STR && !lm.in_key) {           // a string VALUE opens (returning from an escape keeps the id)
            cur_s = n_ids < LGW_TPL_IDS ? n_ids++ : 0xffu;
            if (cur_s != 0xff) { sh_tpl_skind[slot][cur_s] = 0; sh_tpl_sstart[slot][cur_s] = (uint16_t)(pos + 1 - ps); sh_tpl_send[slot][cur_s] = 0xffff; }
        }
        if (!in_num && now_num) {                                             // a number opens
            cur_n = n_ids < LGW_TPL_IDS ? n_ids++ : 0xffu;
            if (cur_n != 0xff) { sh_tpl_skind[slot][cur_n] = 1; sh_tpl_sstart[slot][cur_n] = (uint16_t)(pos - ps); sh_tpl_send[slot][cur_n] = 0xffff; }
            map[pos - ps] = (uint8_t)cur_n;                                    // the number's first character belongs to the span
        }
        ++pos;
    }
    const uint32_t f = lm.finish();
    if (!(f & PF_VALID_A) || (f & (TK_ERROR | TK_DETAIL | TK_CODE))) return;
    if (pos + 2 < limit && env->at(pos + 2) == '\n') return;
    const uint32_t len = pos - ps;
    for (uint32_t k = 0; k < n_ids; ++k) if (sh_tpl_send[slot][k] >= len) return;      // every span must be closed
    uint8_t* text = sh_tpl_bytes + slot * LGW_TPL_STRIDE;
    for (uint32_t k = 0; k < LGW_TPL_STRIDE; ++k) text[k] = k < len ? (uint8_t)env->at(ps + k) : (uint8_t)0;
    sh_tpl_len[slot] = len; sh_tpl_flags[slot] = f; sh_tpl_cls[slot] = cls;
    __threadfence_block();
    sh_tpl_valid[slot] = 1;
}

// publish a freshly built local template to the engine-wide cache (first writer wins per slot;
// `replace` lets a block that re-learnt a slot overwrite a stale entry)
// Slot 1 is only meaningful next to the slot 0 it was learnt against ("first event that missed slot 0"): a block
// whose own slot 0 is not the cache's (it lost the publication race in a cold step) must not define the cache's
// slot 1 -- it could publish the very skeleton that is already in slot 0 and leave the real second shape (the
// usage event) without a template in every later step.
__device__ __noinline__ bool publish_template(TemplateCache* tc, uint32_t slot, bool replace) {
    const uint32_t expect = replace ? 2u : 0u;
    if (slot == 1 && !sh_tpl0_canon) return false;
    if (atomicCAS(&tc->state[slot], expect, 1u) != expect) return false;
    tc->len[slot] = sh_tpl_len[slot]; tc->flags[slot] = sh_tpl_flags[slot]; tc->cls[slot] = sh_tpl_cls[slot];
    for (uint32_t k = 0; k < LGW_TPL_IDS; ++k) { tc->sstart[slot][k] = sh_tpl_sstart[slot][k]; tc->send[slot][k] = sh_tpl_send[slot][k]; tc->skind[slot][k] = sh_tpl_skind[slot][k]; }
    for (uint32_t k = 0; k < LGW_TPLC_TEXT; ++k) tc->text[slot][k] = sh_tpl_bytes[slot * LGW_TPL_STRIDE + k];
    for (uint32_t k = 0; k < LGW_TPLC_MAP; ++k) tc->map[slot][k] = sh_tpl_strid[slot * LGW_TPL_MAPSTRIDE + k];
    __threadfence();
    atomicExch(&tc->state[slot], 2u);
    if (slot == 0) sh_tpl0_canon = 1;
    return true;
}

// rare path: chunk-level UTF-8 validation (request_handler.py:111 decodes each chunk on its own)
template <class ENV>
__device__ __noinline__ bool chunk_utf8_ok(const ENV* rd, uint32_t o, uint32_t e) {
    uint32_t p = o;
    while (p < e) {
        const uint32_t ch = rd->at(p);
        if (ch < 0x80) { ++p; continue; }
        uint32_t need, l = 0x80, h = 0xBF;
        if (ch >= 0xC2 && ch <= 0xDF) need = 1;
        else if (ch == 0xE0) { need = 2; l = 0xA0; }
        else if (ch >= 0xE1 && ch <= 0xEC) need = 2;
        else if (ch == 0xED) { need = 2; h = 0x9F; }
        else if (ch >= 0xEE && ch <= 0xEF) need = 2;
        else if (ch == 0xF0) { need = 3; l = 0x90; }
        else if (ch >= 0xF1 && ch <= 0xF3) need = 3;
        else if (ch == 0xF4) { need = 3; h = 0x8F; }
        else return false;
        if (p + need >= e) return false;
        uint32_t b = rd->at(p + 1);
        if (b < l || b > h) return false;
        for (uint32_t k = 2; k <= need; ++k) { b = rd->at(p + k); if (b < 0x80 || b > 0xBF) return false; }
        p += need + 1;
    }
    return true;
}

// rare path: where does the event that is open at byte o begin?  (o is not right after a separator)
// returns false when the stream must go to the sequential path
template <class ENV>
__device__ __noinline__ bool find_open_event_start(const ENV* rd, uint32_t o, uint32_t relay_begin, uint32_t seg_end, uint32_t carry_cap, uint32_t* out_b) {
    uint32_t k = o;
    bool found = false;
    const uint32_t limit = (o - relay_begin > carry_cap + 2) ? o - carry_cap - 2 : relay_begin;
    while (k >= limit + 2) {
        if (rd->at(k - 1) == '\n' && rd->at(k - 2) == '\n') { found = true; break; }
        --k;
    }
    uint32_t b;
    if (found) {
        b = k;
        if (k >= relay_begin + 3 && rd->at(k - 3) == '\n') return false;          // LF run >= 3
        if (b < seg_end && rd->at(b) == '\n') return false;                        // LF run >= 3
    } else {
        b = relay_begin;
        if (limit != relay_begin) return false;                                    // open event longer than the carry capacity
    }
    if (o - b > carry_cap) return false;
    *out_b = b;
    return true;
}

#if LGW_DEFER_LONG
// ---- warp matcher (k_relay_long) -------------------------------------------------------------------------
// The judgement of match_template made by a whole warp for ONE event: every lane holds the same arguments, lane l
// compares the l-th word of the current 128 bytes, a ballot finds the first difference / the first special byte of a
// string value.  Number values are short and are re-validated by every lane alike.
template <class ENV>
__device__ __forceinline__ bool match_template_warp(const ENV& env, uint32_t slot, uint32_t ps, uint32_t* end) {
    const uint32_t full = 0xFFFFFFFFu, lane = threadIdx.x & 31u;
    const uint32_t lenA = sh_tpl_len[slot];
    const uint32_t tb = env.tpl_s + slot * LGW_TPL_STRIDE, sb = env.strid_s + slot * LGW_TPL_MAPSTRIDE;
    uint32_t ia = 0, ib = ps;
    bool fresh = false;
    for (;;) {
        if (ia < lenA) {
            uint32_t run = 0;
            for (;;) {                                   // equal run, 128 bytes per round
                const uint32_t left = lenA - ia, k = 4 * lane;
                uint32_t x = 0;
                if (k < left) {
                    x = env.tplu(tb, ia + k) ^ env.wordu(ib + k);
                    if (left - k < 4) x &= (1u << (8 * (left - k))) - 1u;
                }
                const uint32_t m = __ballot_sync(full, x != 0);
                if (m == 0) {
                    const uint32_t adv = left < 128u ? left : 128u;
                    ia += adv; ib += adv; run += adv;
                    if (ia >= lenA) break;
                    continue;
                }
                const uint32_t j = (uint32_t)__ffs(m) - 1u;
                const uint32_t xj = __shfl_sync(full, x, j);
                const uint32_t adv = 4 * j + ((uint32_t)(__ffs(xj) - 1) >> 3);
                ia += adv; ib += adv; run += adv;
                break;
            }
            if (run == 0 && fresh) return false;
            fresh = false;
        }
        if (ia >= lenA) break;
        const uint32_t id = lds_u8(sb + ia);
        if (id == 0xffu) return false;
        if (sh_tpl_skind[slot][id] == 0) {   // string value: plain bytes and valid escapes up to the closing quote
            for (;;) {
                const uint32_t w4 = env.wordu(ib + 4 * lane);
                const uint32_t k = first_special(w4);
                const uint32_t m = __ballot_sync(full, k != 4u);
                if (m == 0) { ib += 128u; if (ib - ps > 8192u) return false; continue; }
                const uint32_t j = (uint32_t)__ffs(m) - 1u;
                const uint32_t kj = __shfl_sync(full, k, j), wj = __shfl_sync(full, w4, j);
                ib += 4 * j + kj;
                const uint32_t sp = (wj >> (8 * kj)) & 0xffu;
                if (sp == '"') break;
                if (sp != '\\') return false;
                const uint32_t adv = escape_length(env, ib);
                if (adv == 0) return false;
                ib += adv;
            }
        } else {                             // number value: re-validate the event's own number
            uint32_t bs = ib - (ia - sh_tpl_sstart[slot][id]);
            uint32_t st = L_VALUE;
            {   // common case first: a plain run of digits ("0" or [1-9][0-9]*)
                uint32_t q = bs, c = env.at(q);
                if (c - '1' < 9u) { do { c = env.at(++q); } while (c - '0' < 10u); }
                else if (c == '0') c = env.at(++q);
                else q = bs;
                if (q != bs && c != '.' && c != 'e' && c != 'E') { ib = q; ia = sh_tpl_send[slot][id]; fresh = true; continue; }
            }
            for (;;) {
                const uint32_t cl = env.cls(env.at(bs));
                if (cl < C_MINUS || cl > C_EXP) break;
                st = env.trans(st * 32 + cl) & 31u;
                if (st == L_ERR || bs - ps > 8192u) return false;
                ++bs;
            }
            if (!(st == L_NUM_ZERO || st == L_NUM_INT || st == L_NUM_FRAC || st == L_NUM_EXP)) return false;
            ib = bs;
        }
        ia = sh_tpl_send[slot][id];
        fresh = true;
    }
    if ((env.wordu(ib) & 0xffffu) != 0x0a0au) return false;     // LF LF must follow
    *end = ib;
    return true;
}

#endif  // LGW_DEFER_LONG

// ---- one chunk: the events that complete in it ---------------------------------------------------------
// BULK: called by k_relay for a chunk that starts in the staged tile (ENV = TileEnv).  Otherwise: k_relay_long, reading
// global memory (ENV = GlobalEnv).  Findings are posted with atomics, so which kernel walked a chunk does not matter.
template <class ENV, bool BULK>
__device__ __forceinline__ void walk_chunk(const StepArgs& a, const ENV& env, const uint32_t c, const uint32_t seg_lo, const uint32_t seg_hi,
                                           const int tile_high, const uint32_t utf8_from, const bool have_tpl0, const bool have_tpl1,
                                           uint32_t& acc_seg, uint32_t& ev_a, uint32_t& ev_b, const uint32_t defer_min, const bool lead) {
    const uint32_t n_bytes = a.n_bytes;      // (!BULK: a whole warp walks the chunk, every lane alike; `lead` posts the sums)
    uint32_t seg = seg_lo;
    if (seg_lo != seg_hi) {                                // segment of chunk c: last seg with seg_chunk[seg] <= c
        uint32_t lo = seg_lo, hi = seg_hi;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(a.seg_chunk + mid + 1) <= c) lo = mid + 1; else hi = mid; }
        seg = lo;
    }
    SegPlan* pl = a.s.plan + seg;
    const uint32_t seg_end = pl->seg_end, kept_chunk = pl->kept_chunk;
    const bool is_kept = c == kept_chunk;
    // text visible to both loops starts after the kept chunk; the kept chunk itself is tap-only
    const uint32_t relay_begin = (kept_chunk != 0xFFFFFFFFu && !is_kept) ? pl->kept_end : pl->relay_begin;
    const uint32_t o = __ldg(a.chunk_off + c), e = __ldg(a.chunk_off + c + 1);
    if (o < relay_begin || pl->irregular) return;
    if (e == o) { pl->irregular = 1; return; }
    if (BULK && e - o > defer_min) {                       // a chunk much longer than its neighbours: one lane would walk it while its warp and,
        const uint32_t qi = atomicAdd(a.s.long_count, 1u);   // at the next barrier, its block wait -> k_relay_long takes it, thread per long chunk
        if (qi < a.s.long_cap) { a.s.long_q[2 * qi] = c; a.s.long_q[2 * qi + 1] = seg; return; }
    }
    if (seg != acc_seg) {
        if (acc_seg != 0xFFFFFFFFu) { if (ev_a) atomicAdd(&a.s.plan[acc_seg].n_events_a, ev_a); if (ev_b) atomicAdd(&a.s.plan[acc_seg].n_events_b, ev_b); }
        acc_seg = seg; ev_a = ev_b = 0;
    }

    // chunk-level UTF-8: ASCII tiles need no check; the part of a chunk beyond the tile is
    // scanned for bytes >= 0x80 with 16-byte loads first
    bool need_utf8 = tile_high != 0;
    const uint32_t scan_from = BULK ? utf8_from : (o & ~15u);     // (bytes of a neighbouring chunk in the first or last vector can only cause a spurious check)
    if (!need_utf8 && e > scan_from) {
        uint32_t hi_bits = 0;
        for (uint32_t v = scan_from >> 4; (v << 4) < e; ++v) {
            if ((v << 4) + 16 <= n_bytes) { const uint4 x = __ldg(reinterpret_cast<const uint4*>(a.data) + v); hi_bits |= x.x | x.y | x.z | x.w; }
            else hi_bits = 0x80;                      // ragged end: take the exact path
        }
        need_utf8 = (hi_bits & 0x80808080u) != 0;     // bytes after e in the last vector can only cause a spurious check
    }
    if (need_utf8 && !chunk_utf8_ok(&env, o, e)) { pl->irregular = 1; return; }

    // where does the event that is open at the start of this chunk begin?
    uint32_t b = o;
    if (o >= relay_begin + 3) {                               // common case: one read of the bytes o-3 .. o
        const uint32_t w = env.wordu(o - 3);
        if ((w & 0x00ffff00u) == 0x000a0a00u) {               // a separator ends right before the chunk
            if ((w & 0xffu) == '\n' || (w >> 24) == '\n') { pl->irregular = 1; return; }   // LF run >= 3
        } else if (!find_open_event_start(&env, o, relay_begin, seg_end, a.t.carry_cap, &b)) { pl->irregular = 1; return; }
    } else if (o != relay_begin) {
        const bool sep_before = o >= relay_begin + 2 && (env.wordu(o - 2) & 0xffffu) == 0x0a0au;
        if (sep_before) {
            if ((o >= relay_begin + 3 && env.at(o - 3) == '\n') || env.at(o) == '\n') { pl->irregular = 1; return; }   // LF run >= 3
        } else if (!find_open_event_start(&env, o, relay_begin, seg_end, a.t.carry_cap, &b)) { pl->irregular = 1; return; }
    }

    // walk the events that complete inside this chunk (second LF of the separator in [o, e))
    bool irregular = false, primed = false;
    uint32_t us_b = 0, a_usage = 0; unsigned long long last_usage = 0;
    uint32_t ps = b;
    while (ps < e) {
        uint32_t cls = PC_NONE, f = 0, pos = 0;
        bool ended = false;
        uint32_t hit = 2;
        bool lf3_clear = false;                                  // the window matcher saw that no third LF follows
#pragma unroll 1
        for (uint32_t sl = 0; sl < LGW_TPL_SLOTS && hit == 2u; ++sl) {   // (one copy of the matcher in the loop: fewer registers, faster)
            if (!(sl ? have_tpl1 : have_tpl0)) continue;
#if LGW_WINDOW_MATCH
            if constexpr (BULK) {
                uint32_t r = match_window(env, sl, ps, &pos);
                if (r == 2u) r = match_template_far(env, sl, ps, &pos) ? 1u : 0u;
                else if (r) lf3_clear = true;
                if (r) hit = sl;
            } else
#endif
#if LGW_DEFER_LONG
            if (match_template_warp(env, sl, ps, &pos)) hit = sl;
#else
            if (match_template(env, sl, ps, &pos)) hit = sl;
#endif
        }
        if (hit < 2) {
            if (pos + 1 >= e) break;                         // the separator completes in a later chunk
            ended = true; cls = sh_tpl_cls[hit]; f = sh_tpl_flags[hit];
        } else {
            // classify the event prefix: "data: {" (handler + tap), "{" (tap only), anything else is skipped
            const uint32_t w0 = env.word(ps) >> (8 * (ps & 3u));
            if ((w0 & 0xffu) == '{') cls = PC_BRACE;
            else if ((w0 & 0xffu) == 'd') {
                if ((ps & 3u) == 0) cls = (w0 == 0x61746164u && (env.word(ps + 4) & 0xFFFFFFu) == 0x7b203au) ? PC_DATA : PC_NONE;
                else cls = (env.at(ps + 1) == 'a' && env.at(ps + 2) == 't' && env.at(ps + 3) == 'a' && env.at(ps + 4) == ':' && env.at(ps + 5) == ' ' && env.at(ps + 6) == '{') ? PC_DATA : PC_NONE;
            }
            // nominate the event for the second template (any valid event is a sound template, wherever it came from)
            if (BULK && cls != PC_NONE && have_tpl0) {
                if (!have_tpl1 && sh_tpl_tries1 < 3) atomicMin(&sh_tpl_cand1, ((unsigned long long)ps << 32) | seg_end);
                else if (have_tpl1) atomicAdd(&sh_tpl_miss, 1u);
            }
            LeanMachine lm;
            lm.reset(cls == PC_DATA);
            pos = ps + (cls == PC_DATA ? 6u : 0u);          // "data: " holds no LF
            while (pos < e && !ended) {                      // word-wise byte loop
                uint32_t w = env.word(pos) >> (8 * (pos & 3u));
                uint32_t nb = 4 - (pos & 3u);
                if (nb > e - pos) nb = e - pos;
#pragma unroll 1
                for (; nb; --nb, w >>= 8, ++pos) {
                    const uint32_t ch = w & 0xffu;
                    // hot path: a plain byte inside a string changes nothing
                    if (lm.st == L_STR && ch >= 0x20u && ch != '"' && ch != '\\') continue;
                    if (ch == '\n') {
                        if (pos + 1 >= e) { pos = e; break; }     // a separator starting on the last byte completes later
                        if (env.at(pos + 1) == '\n') { ended = true; break; }
                    }
                    if (cls != PC_NONE) lm.step(ch, pos, env);
                }
            }
            if (!ended) break;                               // the open event completes in a later chunk
            f = lm.finish();
        }
        // ---- one complete event [ps, pos) ----
        if (!lf3_clear && pos + 2 < seg_end && env.at(pos + 2) == '\n') { irregular = true; break; }   // LF run >= 3
        if (cls != PC_NONE) {
            if (cls == PC_DATA) {
                if (is_kept) {                               // priming loop on the kept chunk, request_handler.py:82-91
                    if (!primed) {
                        if (!(f & PF_VALID_A) || (f & (TK_ERROR | TK_DETAIL))) { irregular = true; break; }   // attempt fails: exact path
                        primed = true;
                    }
                } else {                                     // handler loop, request_handler.py:122-134
                    ++ev_a;
                    if ((f & PF_VALID_A) && !(f & TK_CODE) && (f & TK_USAGE)) a_usage = 1;
                }
            }
            if (f & PF_VALID_B) {                            // tap loop, chat_logging.py:123-141
                ++ev_b;
                // events with "error" (extra DB row) go to the sequential path; a "usage" event is
                // a candidate whose values k_commit stashes for extraction on demand
                if (f & TK_ERROR) { irregular = true; break; }
                if (f & TK_USAGE) { ++us_b; last_usage = ((unsigned long long)(ps + 1) << 32) | (pos - ps); }
            }
        }
        ps = pos + 2;
    }
    if (is_kept && !irregular) {
        // the speculation holds when a real event was accepted and the chunk ends on a separator
        if (primed && ps == e) pl->prime_ok = 1; else irregular = true;
    }
    if (irregular) { pl->irregular = 1; return; }
    if (us_b && lead) { atomicAdd(&pl->n_usage_b, us_b); atomicMax(&pl->last_usage, last_usage); }
    if (a_usage) pl->a_usage = 1;
}

// ---- k_prime ---------------------------------------------------------------------------------------
// thread i: (a) tile table entry i = first chunk that starts at or after byte i*TILE;
//           (b) segment i: the plan the bulk kernel works to.  Fresh streams are SPECULATED to commit on
//               their first non-empty chunk; the bulk kernel verifies, k_commit applies or falls back.
__global__ void __launch_bounds__(128) k_prime(StepArgs a, uint32_t n_tiles) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    // tile table: tile_chunk[t] = first c in [chunk_lo, chunk_hi] with chunk_off[c] >= start of tile t.  One thread per
    // chunk writes the tiles whose start falls in (chunk_off[c-1], chunk_off[c]] -- one coalesced pass over the offsets
    // instead of a binary search of ~20 dependent loads per tile.
    if (i == 0) *a.s.long_count = 0;                       // queue of the chunks k_relay leaves to k_relay_long
    const uint32_t n_off = a.chunk_hi - a.chunk_lo + 1;
    const uint32_t j0 = i * 4u;                              // four consecutive chunks per thread, their five offsets loaded up front
    if (j0 < n_off) {
        uint32_t offs[5];                                    // offs[k] = offset of chunk chunk_lo + j0 + k - 1
#pragma unroll
        for (uint32_t k = 0; k < 5; ++k) { const uint32_t j = j0 + k; offs[k] = (j >= 1u && j - 1u < n_off) ? __ldg(a.chunk_off + a.chunk_lo + j - 1u) : 0u; }
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k) {
            const uint32_t j = j0 + k;
            if (j < n_off) {
                const uint32_t c = a.chunk_lo + j, off = offs[k + 1];
                uint32_t t_first = 0;
                if (j > 0) { const uint32_t prev = offs[k]; t_first = prev < a.tile_base ? 0u : (prev - a.tile_base) / LGW_TILE_BYTES + 1u; }
                uint32_t t_last = off < a.tile_base ? 0u : (off - a.tile_base) / LGW_TILE_BYTES;      // off >= tile_base always holds for c >= chunk_lo
                if (off < a.tile_base) t_first = 1;                                                  // (defensive: nothing to write)
                if (c == a.chunk_hi) t_last = n_tiles;                                               // tiles past the last offset
                if (t_last > n_tiles) t_last = n_tiles;
                for (uint32_t t = t_first; t <= t_last; ++t) a.s.tile_chunk[t] = c;
            }
        }
    }
    if (i >= a.n_segs) return;
    const uint32_t seg = i, c0 = a.seg_chunk[seg], c1 = a.seg_chunk[seg + 1];
    const uint32_t slot = a.seg_slot[seg];
    const StreamHdr st = a.t.state[slot].h;
    SegPlan p;
    p.seg_end = a.chunk_off[c1]; p.relay_begin = p.seg_end; p.irregular = 0; p.last_usage = 0; p.a_usage = 0;
    p.n_events_a = p.n_events_b = p.n_usage_b = 0; p._pad[0] = p._pad[1] = 0;
    p.kept_chunk = 0xFFFFFFFFu; p.kept_end = 0; p.prime_ok = 0;
    p.resume_chunk = c0; p.emit_chunk_begin = (st.phase == PH_COMMITTED) ? c0 : c1;
    if (st.phase == PH_COMMITTED && c0 < c1) {
        if ((st.flags & SF_SYNCED) && st.carry_a_len == 0) p.relay_begin = a.chunk_off[c0];
        else p.irregular = 1;
    } else if (st.phase == PH_PRIMING && c0 < c1) {
        uint32_t c = c0;
        while (c < c1 && a.chunk_off[c + 1] == a.chunk_off[c]) ++c;          // empty chunks are never yielded (:60-63)
        if (c < c1 && st.carry_a_len == 0) {
            p.kept_chunk = c; p.kept_end = a.chunk_off[c + 1]; p.relay_begin = a.chunk_off[c]; p.emit_chunk_begin = c;
        } else if (c < c1) p.irregular = 1;                                    // an event is open from an earlier step
    }
    a.s.plan[seg] = p;
}

// tables of the recogniser and the engine-wide event templates into shared memory (k_relay, k_relay_long)
__device__ __forceinline__ void relay_prologue(const StepArgs& a, const uint32_t tid) {
    for (uint32_t k = tid; k < 64 + LGW_LEAN_ROWS * 8; k += LGW_RELAY_THREADS) {
        if (k < 64) reinterpret_cast<uint32_t*>(sh_cls)[k] = reinterpret_cast<const uint32_t*>(g_lean_tables_dev.cls)[k];
        else reinterpret_cast<uint32_t*>(sh_trans)[k - 64] = reinterpret_cast<const uint32_t*>(g_lean_tables_dev.trans)[k - 64];
    }
    if (tid == 0) { sh_tpl_valid[0] = sh_tpl_valid[1] = 0; sh_tpl_len[0] = sh_tpl_len[1] = 0; sh_tpl_cand1 = ~0ull; sh_tpl_tries1 = 0; sh_tpl0_canon = 0; }
    __syncthreads();
    {   // start from the engine-wide template cache when it has entries
        const TemplateCache* tc = a.s.tpl_cache;
        if (tid < LGW_TPL_SLOTS) sh_tpl_cstate[tid] = *reinterpret_cast<const volatile uint32_t*>(&tc->state[tid]);
        if (tid == 0) { sh_tpl_miss = 0; sh_tpl_replace1 = 0; }
        __syncthreads();
        for (uint32_t slot = 0; slot < LGW_TPL_SLOTS; ++slot) {
            if (sh_tpl_cstate[slot] != 2u) continue;                                                // one decision for the whole block
            for (uint32_t k = tid; k < LGW_TPLC_TEXT; k += LGW_RELAY_THREADS) sh_tpl_bytes[slot * LGW_TPL_STRIDE + k] = tc->text[slot][k];
            for (uint32_t k = tid; k < LGW_TPLC_MAP; k += LGW_RELAY_THREADS) sh_tpl_strid[slot * LGW_TPL_MAPSTRIDE + k] = tc->map[slot][k];
            if (tid < LGW_TPL_IDS) { sh_tpl_sstart[slot][tid] = tc->sstart[slot][tid]; sh_tpl_send[slot][tid] = tc->send[slot][tid]; sh_tpl_skind[slot][tid] = tc->skind[slot][tid]; }
            if (tid == 0) { sh_tpl_len[slot] = tc->len[slot]; sh_tpl_flags[slot] = tc->flags[slot]; sh_tpl_cls[slot] = tc->cls[slot]; sh_tpl_valid[slot] = 1; if (slot == 0) sh_tpl0_canon = 1; }
        }
    }
}

// ---- k_relay ---------------------------------------------------------------------------------------
// Persistent blocks: block b handles tiles [b*tiles_per_block, ...) so that its template carries over.
__global__ void __launch_bounds__(LGW_RELAY_THREADS, LGW_RELAY_BLOCKS_PER_SM) k_relay(StepArgs a, uint32_t n_tiles, uint32_t tiles_per_block) {
    const uint32_t tid = threadIdx.x;
    const uint32_t n_bytes = a.n_bytes;
    // this block's slice of the tile table (tiles_per_block + 1 entries) once, instead of two L2 round trips at the top of every tile
    const uint32_t tile_first = blockIdx.x * tiles_per_block;
    const uint32_t tile_last = min(n_tiles, tile_first + tiles_per_block);
    if (tid < LGW_TC_STAGED && tile_first + tid <= tile_last) sh_tile_chunk[tid] = a.s.tile_chunk[tile_first + tid];
    relay_prologue(a, tid);

    TileEnv env;
    env.tile_s = opaque((uint32_t)__cvta_generic_to_shared(sh_tile));
    env.cls_s = opaque((uint32_t)__cvta_generic_to_shared(sh_cls));
    env.trans_s = opaque((uint32_t)__cvta_generic_to_shared(sh_trans));
    env.tpl_s = opaque((uint32_t)__cvta_generic_to_shared(sh_tpl_bytes));
    env.strid_s = opaque((uint32_t)__cvta_generic_to_shared(sh_tpl_strid));
    env.g = a.data; env.n_bytes = n_bytes;

    uint32_t seg_hint = 0;                     // lower bound of the segment index of this thread's next search
    if (tid >= LGW_RELAY_THREADS - 32 && tid < LGW_RELAY_THREADS - 30 && tile_first < tile_last) {      // first tile: one binary search
        const uint32_t c = a.s.tile_chunk[tile_first];
        uint32_t lo = 0, hi = a.n_segs;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(a.seg_chunk + mid + 1) <= c) lo = mid + 1; else hi = mid; }
        seg_hint = lo;
    }
    for (uint32_t tile = tile_first; tile < tile_last; ++tile) {
        const uint32_t t0 = a.tile_base + tile * LGW_TILE_BYTES;
        env.t0 = t0;
        const uint32_t ti = tile - tile_first;
        const uint32_t c_lo = ti < LGW_TC_STAGED ? sh_tile_chunk[ti] : a.s.tile_chunk[tile];
        const uint32_t c_hi = ti + 1 < LGW_TC_STAGED ? sh_tile_chunk[ti + 1] : a.s.tile_chunk[tile + 1];
        // the walk of a chunk starts with dependent global loads (its offsets, its segment's plan): pull those lines
        // into L1 now, while the tile is being staged
        for (uint32_t c = c_lo + tid; c <= c_hi; c += LGW_RELAY_THREADS) prefetch_l1(a.chunk_off + c);
        DBG_STAMP(0);
        __syncthreads();                                       // the previous tile's readers are done
        DBG_STAMP(1);
        // slot 1 adapts: when most events of the previous tile matched neither template, forget it and
        // re-learn it from an event that missed (the engine-wide cache entry is then replaced too)
        if (tid == 0) {
            if (sh_tpl_valid[1] && sh_tpl_miss > LGW_RELAY_THREADS / 2) { sh_tpl_valid[1] = 0; sh_tpl_tries1 = 0; sh_tpl_replace1 = 1; }
            sh_tpl_miss = 0;
        }
        __syncthreads();
        // second template: learnt from the first event of the previous tiles that missed slot 0; the
        // whole block stages its bytes from global memory, one thread validates the staged copy
        if (!sh_tpl_valid[1] && sh_tpl_cand1 != ~0ull) {
            const uint32_t cps = (uint32_t)(sh_tpl_cand1 >> 32), climit = (uint32_t)sh_tpl_cand1;
            for (uint32_t k = tid; k < LGW_TPL_STRIDE; k += LGW_RELAY_THREADS) sh_stage[k] = cps + k < n_bytes ? __ldg(a.data + cps + k) : (uint8_t)0;
            __syncthreads();
            if (tid == 0) {
                StageEnv senv{cps, env.cls_s, env.trans_s};
                build_template(&senv, 1, cps, climit);
                sh_tpl_cand1 = ~0ull;
                if (!sh_tpl_valid[1]) ++sh_tpl_tries1; else publish_template(a.s.tpl_cache, 1, sh_tpl_replace1 != 0);
            }
            __syncthreads();
        }

        // (0) segment range of this tile's chunks
        //     (a block walks consecutive tiles, so both ends are at or just after the previous tile's last segment:
        //     a few forward steps instead of a binary search of dependent loads over all segments)
        if (tid >= LGW_RELAY_THREADS - 32 && tid < LGW_RELAY_THREADS - 30 && c_hi > c_lo) {
            const uint32_t c = tid == LGW_RELAY_THREADS - 32 ? c_lo : c_hi - 1;
            uint32_t lo = seg_hint, hi = a.n_segs;
            for (uint32_t k = 0; k < 4 && lo < hi && __ldg(a.seg_chunk + lo + 1) <= c; ++k) ++lo;
            if (lo < hi && __ldg(a.seg_chunk + lo + 1) <= c) {
                while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(a.seg_chunk + mid + 1) <= c) lo = mid + 1; else hi = mid; }
            }
            seg_hint = lo;                      // (the thread looking for the first chunk's segment lags the other by one tile: still a lower bound)
            if (tid == LGW_RELAY_THREADS - 32) sh_seg_lo = lo; else sh_seg_hi = lo;
            prefetch_l1(a.s.plan + (lo < a.n_segs ? lo : 0u));      // the walk starts from these plans: have them in L1 by the time the tile is staged
        }
        if (tid == 0) sh_tpl_cand = 0xFFFFFFFFu;

        // (1) re-emit: position-preserving 16-byte copy of the tile, staged (swizzled) into shared memory
        //     on the way; note whether the tile has any byte >= 0x80 (UTF-8 checks are skipped otherwise)
        uint32_t high = 0;
        if (t0 + LGW_STAGE_BYTES <= n_bytes) {              // every tile but the last: no bounds tests
            const uint4* __restrict__ src = reinterpret_cast<const uint4*>(a.data + t0);
            uint4* __restrict__ dst = reinterpret_cast<uint4*>(a.out + t0);
            uint4 x[LGW_TILE_VECS / LGW_RELAY_THREADS];
#pragma unroll
            for (uint32_t k = 0; k < LGW_TILE_VECS / LGW_RELAY_THREADS; ++k) x[k] = ldg_stream(src + k * LGW_RELAY_THREADS + tid);
#pragma unroll
            for (uint32_t k = 0; k < LGW_TILE_VECS / LGW_RELAY_THREADS; ++k) {
                const uint32_t v = k * LGW_RELAY_THREADS + tid;
                dst[v] = x[k];
                high |= x[k].x | x[k].y | x[k].z | x[k].w;
                const uint32_t pa = env.tile_s + phys(v << 4);
                sts_u32(pa, x[k].x); sts_u32(pa + 4, x[k].y); sts_u32(pa + 8, x[k].z); sts_u32(pa + 12, x[k].w);
            }
            if (tid < LGW_HALO_BYTES / 16) {
                const uint32_t v = LGW_TILE_VECS + tid;
                const uint4 h = ldg_stream(src + v);
                high |= h.x | h.y | h.z | h.w;
                const uint32_t pa = env.tile_s + phys(v << 4);
                sts_u32(pa, h.x); sts_u32(pa + 4, h.y); sts_u32(pa + 8, h.z); sts_u32(pa + 12, h.w);
            }
        } else {
            const uint4* __restrict__ src = reinterpret_cast<const uint4*>(a.data + t0);
            uint4* __restrict__ dst = reinterpret_cast<uint4*>(a.out + t0);
            uint4 x[LGW_TILE_VECS / LGW_RELAY_THREADS];
#pragma unroll
            for (uint32_t k = 0; k < LGW_TILE_VECS / LGW_RELAY_THREADS; ++k) {
                const uint32_t v = k * LGW_RELAY_THREADS + tid;
                x[k] = (t0 + v * 16 + 16 <= n_bytes) ? ldg_stream(src + v) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (uint32_t k = 0; k < LGW_TILE_VECS / LGW_RELAY_THREADS; ++k) {
                const uint32_t v = k * LGW_RELAY_THREADS + tid;
                const uint32_t pos = t0 + v * 16;
                if (pos + 16 <= n_bytes) {
                    dst[v] = x[k];
                } else if (pos < n_bytes) {                        // ragged end of the buffer
                    uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
                    for (uint32_t b = pos; b < n_bytes; ++b) {
                        const uint32_t c = a.data[b]; a.out[b] = (uint8_t)c;
                        const uint32_t sh = c << (8 * ((b - pos) & 3)), wi = (b - pos) >> 2;
                        if (wi == 0) w0 |= sh; else if (wi == 1) w1 |= sh; else if (wi == 2) w2 |= sh; else w3 |= sh;
                    }
                    x[k] = make_uint4(w0, w1, w2, w3);
                }
                high |= x[k].x | x[k].y | x[k].z | x[k].w;
                const uint32_t pa = env.tile_s + phys(v << 4);      // a 16-byte vector never straddles a 64-byte row
                sts_u32(pa, x[k].x); sts_u32(pa + 4, x[k].y); sts_u32(pa + 8, x[k].z); sts_u32(pa + 12, x[k].w);
            }
            // halo: the first bytes of the next tile, staged only (their own tile copies them out)
            if (tid < LGW_HALO_BYTES / 16) {
                const uint32_t v = LGW_TILE_VECS + tid;
                const uint32_t pos = t0 + v * 16;
                uint4 h = make_uint4(0, 0, 0, 0);
                if (pos + 16 <= n_bytes) h = ldg_stream(src + v);
                else if (pos < n_bytes) {
                    uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
                    for (uint32_t b = pos; b < n_bytes; ++b) {
                        const uint32_t c = a.data[b];
                        const uint32_t sh = c << (8 * ((b - pos) & 3)), wi = (b - pos) >> 2;
                        if (wi == 0) w0 |= sh; else if (wi == 1) w1 |= sh; else if (wi == 2) w2 |= sh; else w3 |= sh;
                    }
                    h = make_uint4(w0, w1, w2, w3);
                }
                high |= h.x | h.y | h.z | h.w;                    // (the halo counts: chunks that end inside it need no further scan)
                const uint32_t pa = env.tile_s + phys(v << 4);
                sts_u32(pa, h.x); sts_u32(pa + 4, h.y); sts_u32(pa + 8, h.z); sts_u32(pa + 12, h.w);
            }
        }
        DBG_STAMP(2);
        const int tile_high = __syncthreads_or((high & 0x80808080u) != 0);
        DBG_STAMP(3);
        const uint32_t seg_lo = sh_seg_lo, seg_hi = sh_seg_hi;
        if (tid <= seg_hi - seg_lo && tid < 8) prefetch_l1(a.s.plan + seg_lo + tid);

        // (1b) no template yet: the lowest thread whose chunk starts with an event builds one
        if (!sh_tpl_valid[0]) {
            uint32_t cand_ps = 0, cand_limit = 0;
            if (c_lo + tid < c_hi) {
                const uint32_t c = c_lo + tid;
                uint32_t lo = seg_lo, hi = seg_hi;
                while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(a.seg_chunk + mid + 1) <= c) lo = mid + 1; else hi = mid; }
                const SegPlan* pl = a.s.plan + lo;
                const uint32_t o = __ldg(a.chunk_off + c);
                const uint32_t begin = (pl->kept_chunk != 0xFFFFFFFFu && c > pl->kept_chunk) ? pl->kept_end : pl->relay_begin;
                if (!pl->irregular && o >= begin && o + 8 < pl->seg_end && (o == begin || (o >= begin + 2 && (env.wordu(o - 2) & 0xffffu) == 0x0a0au))) {
                    const uint32_t w = env.wordu(o);
                    if ((w & 0xffu) == '{' || w == 0x61746164u) { cand_ps = o; cand_limit = pl->seg_end; atomicMin(&sh_tpl_cand, tid); }
                }
            }
            __syncthreads();
            if (sh_tpl_cand == tid) { build_template(&env, 0, cand_ps, cand_limit); if (sh_tpl_valid[0]) publish_template(a.s.tpl_cache, 0, false); }
            __syncthreads();
        }
        const bool have_tpl0 = sh_tpl_valid[0] != 0, have_tpl1 = sh_tpl_valid[1] != 0;

        DBG_STAMP(4);
        // (2) events of the chunks that START in this tile
        uint32_t acc_seg = 0xFFFFFFFFu, ev_a = 0, ev_b = 0;       // per-thread counters of the current segment
        // chunks longer than 128 B and than twice this tile's average chunk are left to k_relay_long
        uint32_t defer_min = 0xFFFFFFFFu;
#if LGW_DEFER_LONG
        if (c_hi > c_lo && a.n_bytes < 0xFFFF0000u) { defer_min = 2u * (LGW_TILE_BYTES / (c_hi - c_lo)); if (defer_min < 128u) defer_min = 128u; }
#endif
        for (uint32_t c = c_lo + tid; c < c_hi; c += LGW_RELAY_THREADS)
            walk_chunk<TileEnv, true>(a, env, c, seg_lo, seg_hi, tile_high, t0 + LGW_STAGE_BYTES, have_tpl0, have_tpl1, acc_seg, ev_a, ev_b, defer_min, true);

        DBG_STAMP(5);
        // (3) post the event counters: one atomic per warp when the whole warp worked on one segment
        {
            const uint32_t full = 0xFFFFFFFFu;
            const uint32_t seg0 = __shfl_sync(full, acc_seg, 0);
            const bool uniform = __all_sync(full, acc_seg == seg0);
            if (uniform) {
                if (seg0 != 0xFFFFFFFFu) {
                    const uint32_t sa = __reduce_add_sync(full, ev_a), sb = __reduce_add_sync(full, ev_b);
                    if ((tid & 31) == 0) { if (sa) atomicAdd(&a.s.plan[seg0].n_events_a, sa); if (sb) atomicAdd(&a.s.plan[seg0].n_events_b, sb); }
                }
            } else if (acc_seg != 0xFFFFFFFFu) {
                if (ev_a) atomicAdd(&a.s.plan[acc_seg].n_events_a, ev_a);
                if (ev_b) atomicAdd(&a.s.plan[acc_seg].n_events_b, ev_b);
            }
        }
    }
}

#if LGW_DEFER_LONG
// ---- k_relay_long ----------------------------------------------------------------------------------
// One WARP per chunk that k_relay left out (much longer than its neighbours: in a warp of 64-byte deltas the one lane
// with a 300-byte usage event holds 31 lanes and, at the tile barrier, its whole block).  Here every lane has a long
// chunk.  Same walk, same templates (from the engine-wide cache), bytes read from global memory.
__global__ void __launch_bounds__(LGW_RELAY_THREADS) k_relay_long(StepArgs a) {
    const uint32_t tid = threadIdx.x;
    uint32_t n = *a.s.long_count;
    if (n > a.s.long_cap) n = a.s.long_cap;
    const uint32_t wpb = LGW_RELAY_THREADS / 32u, warp = tid >> 5, lane = tid & 31u;      // one WARP per long chunk
    if (blockIdx.x * wpb >= n) return;                        // (uniform per block)
    relay_prologue(a, tid);
    __syncthreads();
    GlobalEnv env;
    env.cls_s = opaque((uint32_t)__cvta_generic_to_shared(sh_cls));
    env.trans_s = opaque((uint32_t)__cvta_generic_to_shared(sh_trans));
    env.tpl_s = opaque((uint32_t)__cvta_generic_to_shared(sh_tpl_bytes));
    env.strid_s = opaque((uint32_t)__cvta_generic_to_shared(sh_tpl_strid));
    env.g = a.data; env.n_bytes = a.n_bytes;
    // The long chunks are not walked by k_relay any more, so it never meets (and never learns) their skeleton: the second
    // template is learnt here, from the first event of this block's first chunk when it does not follow slot 0.
    if (sh_tpl_valid[0] && !sh_tpl_valid[1]) {                  // (uniform per block)
        if (tid == 0) {
            const uint32_t i0 = blockIdx.x * wpb;
            const uint32_t c = a.s.long_q[2 * i0], seg = a.s.long_q[2 * i0 + 1];
            const SegPlan* pl = a.s.plan + seg;
            const uint32_t o = __ldg(a.chunk_off + c);
            const uint32_t begin = (pl->kept_chunk != 0xFFFFFFFFu && c > pl->kept_chunk) ? pl->kept_end : pl->relay_begin;
            uint32_t pos_unused;
            if (!pl->irregular && o >= begin && o + 8 < pl->seg_end && (o == begin || (o >= begin + 2 && (env.wordu(o - 2) & 0xffffu) == 0x0a0au))
                && !match_template(env, 0, o, &pos_unused)) {
                build_template(&env, 1, o, pl->seg_end);
                if (sh_tpl_valid[1]) publish_template(a.s.tpl_cache, 1, false);
            }
        }
        __syncthreads();
    }
    const bool have_tpl0 = sh_tpl_valid[0] != 0, have_tpl1 = sh_tpl_valid[1] != 0;
    for (uint32_t i = blockIdx.x * wpb + warp; i < n; i += gridDim.x * wpb) {
        const uint32_t c = a.s.long_q[2 * i], lo = a.s.long_q[2 * i + 1];      // (chunk, its segment)
        uint32_t acc_seg = 0xFFFFFFFFu, ev_a = 0, ev_b = 0;
        walk_chunk<GlobalEnv, false>(a, env, c, lo, lo, 0, 0u, have_tpl0, have_tpl1, acc_seg, ev_a, ev_b, 0xFFFFFFFFu, lane == 0);
        if (lane == 0 && acc_seg != 0xFFFFFFFFu) {
            if (ev_a) atomicAdd(&a.s.plan[acc_seg].n_events_a, ev_a);
            if (ev_b) atomicAdd(&a.s.plan[acc_seg].n_events_b, ev_b);
        }
    }
}

#endif  // LGW_DEFER_LONG

// ---- k_commit --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) k_commit(StepArgs a) {
    const uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= a.n_segs) return;
    const uint32_t c1 = a.seg_chunk[seg + 1];
    const uint32_t slot = a.seg_slot[seg];
    StreamHdr st = a.t.state[slot].h;
    const StepIO io = make_io(a, slot, &st);
    const SegPlan p = a.s.plan[seg];
    uint32_t emit_begin = p.emit_chunk_begin;
    const bool speculated = p.kept_chunk != 0xFFFFFFFFu;
    if ((st.phase == PH_COMMITTED || st.phase == PH_PRIMING) && p.resume_chunk < c1) {
        const uint8_t* __restrict__ d = a.data;
        bool sequential = p.irregular || p.n_usage_b > 1;      // several usage candidates: let the exact path count them
        if (st.phase == PH_PRIMING && !(speculated && p.prime_ok)) sequential = true;
        const uint32_t ups = (uint32_t)(p.last_usage >> 32) - 1u, ulen = (uint32_t)p.last_usage;   // the winning usage event
        if (!sequential && p.last_usage && ulen > LGW_PENDING_CAP) sequential = true;
        if (sequential) {
            emit_begin = (st.phase == PH_COMMITTED) ? p.resume_chunk : c1;
            run_chunks(io, a.data, a.chunk_off, p.resume_chunk, c1, emit_begin, false);
        } else {
            uint32_t first = p.resume_chunk;
            if (speculated) {                       // apply the verified commit (request_handler.py:89-90, chat_logging.py:200)
                st.phase = PH_COMMITTED; st.verdict = VD_OK;
                st.flags |= SF_EMITTED_ANY | SF_SYNCED;
                st.carry_a_len = st.carry_b_len = 0;
                ++st.n_events_a;                    // the priming parse of the first real event
                first = p.kept_chunk;
            }
            const uint32_t nch = c1 - first, nby = p.seg_end - p.relay_begin;
            st.n_chunks_in += nch; st.n_chunks_emitted += nch; st.bytes_in += nby; st.bytes_emitted += nby;
            st.n_events_a += p.n_events_a; st.n_events_b += p.n_events_b;
            if (p.a_usage) st.flags |= SF_A_USAGE_BOUND;
            if (p.last_usage) {                     // the last usage-bearing event wins (chat_logging.py:134-135):
                if (st.flags & SF_PENDING) resolve_pending(io);     // an older stash must be settled first
                // stash its text as whole 16-byte vectors; the values are extracted on demand
                const uint32_t off = ups & 15u, nv = (off + ulen + 15u) >> 4;
                const uint4* src = reinterpret_cast<const uint4*>(d + (ups - off));
                uint4* dst = reinterpret_cast<uint4*>(io.pending);
                if (((size_t)(ups - off) + ((size_t)nv << 4)) <= a.n_bytes) {
                    // eight loads in flight per round (dst may alias src for the compiler: a plain loop is one DRAM round trip per vector)
                    for (uint32_t k0 = 0; k0 < nv; k0 += 8) {
                        uint4 r[8];
#pragma unroll
                        for (uint32_t j = 0; j < 8; ++j) if (k0 + j < nv) r[j] = __ldg(src + k0 + j);
#pragma unroll
                        for (uint32_t j = 0; j < 8; ++j) if (k0 + j < nv) dst[k0 + j] = r[j];
                    }
                }
                else { for (uint32_t k = 0; k < ulen; ++k) io.pending[off + k] = __ldg(d + ups + k); }
                st.pending_len = ulen | (off << 16); st.flags |= SF_PENDING; ++st.n_usage_b;
            }
            // new carry = text after the last separator (both loops: SF_SYNCED)
            const uint32_t text_begin = speculated ? p.kept_end : p.relay_begin;
            const uint32_t tby = p.seg_end - text_begin;
            uint32_t tail = text_begin;
            if (tby == 0 || (tby >= 2 && __ldg(d + p.seg_end - 1) == '\n' && __ldg(d + p.seg_end - 2) == '\n')) tail = p.seg_end;
            else {
                const uint32_t limit = tby > a.t.carry_cap + 2 ? p.seg_end - a.t.carry_cap - 2 : text_begin;
                for (uint32_t k = p.seg_end; k >= limit + 2; --k)
                    if (__ldg(d + k - 1) == '\n' && __ldg(d + k - 2) == '\n') { tail = k; break; }
            }
            const uint32_t n = p.seg_end - tail;
            if (n > a.t.carry_cap) { st.carry_a_len = 0; st.flags |= SF_CARRY_OVERFLOW; }
            else { for (uint32_t k = 0; k < n; ++k) io.carry_a[k] = d[tail + k]; st.carry_a_len = n; }
        }
        a.t.state[slot].h = st;
    }
    SegResult res;
    fill_seg_result(st, emit_begin, c1, res);
    a.seg_out[seg] = res;
}

static inline cudaError_t launch_step_fast(const StepArgs& a, int sm_count, cudaStream_t stream, cudaEvent_t* ev, int* launched) {
    cudaError_t r;
    const uint32_t n_tiles = (a.n_bytes - a.tile_base + LGW_TILE_BYTES - 1) / LGW_TILE_BYTES;
    const uint32_t n_off = a.chunk_hi - a.chunk_lo + 1;
    uint32_t prime_blocks = (n_off + 511) / 512;                        // chunks: four per thread
    if (prime_blocks < (a.n_segs + 127) / 128) prime_blocks = (a.n_segs + 127) / 128;   // segments: one thread each
    k_prime<<<prime_blocks, 128, 0, stream>>>(a, n_tiles); ++*launched;
    if ((r = cudaEventRecord(ev[1], stream)) != cudaSuccess) return r;
    if (n_tiles) {
        const uint32_t max_blocks = (uint32_t)sm_count * LGW_RELAY_BLOCKS_PER_SM;
        const uint32_t tpb = (n_tiles + max_blocks - 1) / max_blocks;
        const uint32_t blocks = (n_tiles + tpb - 1) / tpb;
        k_relay<<<blocks, LGW_RELAY_THREADS, 0, stream>>>(a, n_tiles, tpb); ++*launched;
#if LGW_DEFER_LONG
        k_relay_long<<<(uint32_t)sm_count * 16u, LGW_RELAY_THREADS, 0, stream>>>(a); ++*launched;
#endif
    }
    if ((r = cudaEventRecord(ev[2], stream)) != cudaSuccess) return r;
    if (a.n_segs) { k_commit<<<(a.n_segs + 63) / 64, 64, 0, stream>>>(a); ++*launched; }
    if ((r = cudaEventRecord(ev[3], stream)) != cudaSuccess) return r;
    return cudaGetLastError();
}
