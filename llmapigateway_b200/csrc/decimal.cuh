// Decimal text -> int64 / IEEE-754 double with CPython's semantics (int(text) / float(text)),
// for the numbers the tap copies out of a usage object (chat_logging.py:248-261).
//   * integers: exact int64, else KD_BIG
//   * floats: correctly rounded (round-half-even).  Fast paths: Clinger (exact double ops) and
//     Eisel-Lemire with a 128-bit power-of-ten table; the remaining cases (subnormals, overflow
//     edge, EL's ambiguous half-way results) go to an exact big-integer path.  More than 19
//     significant digits are handled by converting both bracketing 19-digit values; if they
//     disagree the result is KD_FLT_INEXACT (reported, never guessed).
#pragma once
#include <stdint.h>
#include <string.h>

namespace lgw {

#ifndef LGW_KIND_CONSTS
#define LGW_KIND_CONSTS
#define LGW_KD_INT 1
#define LGW_KD_FLT 2
#define LGW_KD_BIG 7
#define LGW_KD_FLT_INEXACT 10
#endif

struct U128 { uint64_t lo, hi; };

#if defined(__CUDACC__)
__device__ __constant__ U128 g_pow10_dev[] = {
#include "pow10_table.inc"
};
#endif
static const U128 g_pow10_host[] = {
#include "pow10_table.inc"
};
#if defined(__CUDA_ARCH__)
#define LGW_POW10(i) g_pow10_dev[i]
#else
#define LGW_POW10(i) g_pow10_host[i]
#endif
#define LGW_POW10_QMIN (-348)
#define LGW_POW10_QMAX 347

LGW_HD void mul64(uint64_t a, uint64_t b, uint64_t& hi, uint64_t& lo) {
#if defined(__CUDA_ARCH__)
    lo = a * b; hi = __umul64hi(a, b);
#else
    unsigned __int128 p = (unsigned __int128)a * b; lo = (uint64_t)p; hi = (uint64_t)(p >> 64);
#endif
}
LGW_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}
LGW_HD double bits2dbl(uint64_t b) { double d; memcpy(&d, &b, 8); return d; }
LGW_HD uint64_t dbl2bits(double d) { uint64_t b; memcpy(&b, &d, 8); return b; }

// Eisel-Lemire; returns false when the result is not certain (caller falls back).
LGW_HD bool eisel_lemire(uint64_t man, int exp10, uint64_t& out_bits) {
    if (man == 0) { out_bits = 0; return true; }
    if (exp10 < LGW_POW10_QMIN || exp10 > LGW_POW10_QMAX) return false;
    const int clz = clz64(man);
    man <<= clz;
    int64_t ret_exp2 = (int64_t)((217706 * exp10) >> 16) + 64 + 1023 - clz;
    const U128 p = LGW_POW10(exp10 - LGW_POW10_QMIN);
    uint64_t x_hi, x_lo;
    mul64(man, p.hi, x_hi, x_lo);
    if ((x_hi & 0x1FF) == 0x1FF && x_lo + man < man) {
        uint64_t y_hi, y_lo;
        mul64(man, p.lo, y_hi, y_lo);
        uint64_t m_hi = x_hi, m_lo = x_lo + y_hi;
        if (m_lo < x_lo) ++m_hi;
        if ((m_hi & 0x1FF) == 0x1FF && m_lo + 1 == 0 && y_lo + man < man) return false;
        x_hi = m_hi; x_lo = m_lo;
    }
    const uint64_t msb = x_hi >> 63;
    uint64_t ret_man = x_hi >> (msb + 9);
    ret_exp2 -= (int64_t)(1 ^ msb);
    if (x_lo == 0 && (x_hi & 0x1FF) == 0 && (ret_man & 3) == 1) return false;
    ret_man += ret_man & 1;
    ret_man >>= 1;
    if (ret_man >> 53) { ret_man >>= 1; ret_exp2 += 1; }
    if (ret_exp2 <= 0 || ret_exp2 >= 0x7FF) return false;       // subnormal / overflow: exact path
    out_bits = ((uint64_t)ret_exp2 << 52) | (ret_man & 0x000FFFFFFFFFFFFFull);
    return true;
}

// ---- exact big-integer path (rare) -----------------------------------------------------------
// value = man * 10^exp10, man < 2^64.  Little-endian 32-bit limbs.
#define LGW_BIG_LIMBS 44            /* 1408 bits: 64 + log2(10^400) */
struct Big {
    uint32_t w[LGW_BIG_LIMBS]; int n;
    LGW_HD void set64(uint64_t v) { n = 0; for (int i = 0; i < LGW_BIG_LIMBS; ++i) w[i] = 0; w[0] = (uint32_t)v; w[1] = (uint32_t)(v >> 32); n = w[1] ? 2 : (w[0] ? 1 : 0); }
    LGW_HD bool mul_small(uint32_t m) {
        uint64_t c = 0;
        for (int i = 0; i < n; ++i) { uint64_t t = (uint64_t)w[i] * m + c; w[i] = (uint32_t)t; c = t >> 32; }
        if (c) { if (n >= LGW_BIG_LIMBS) return false; w[n++] = (uint32_t)c; }
        return true;
    }
    LGW_HD int bitlen() const { return n == 0 ? 0 : 32 * (n - 1) + (64 - clz64((uint64_t)w[n - 1])); }
    LGW_HD int bit(int i) const { return (i >> 5) < n ? (int)((w[i >> 5] >> (i & 31)) & 1u) : 0; }
    LGW_HD bool any_below(int i) const {          // any set bit strictly below position i
        for (int k = 0; k < (i >> 5) && k < n; ++k) if (w[k]) return true;
        if ((i >> 5) < n && (i & 31)) return (w[i >> 5] & ((1u << (i & 31)) - 1)) != 0;
        return false;
    }
    LGW_HD bool shl(int s) {                      // *= 2^s
        if (n == 0 || s == 0) return true;
        const int ws = s >> 5, bs = s & 31;
        const int nn = n + ws + (bs ? 1 : 0);
        if (nn > LGW_BIG_LIMBS) return false;
        for (int i = nn - 1; i >= 0; --i) {
            const int src = i - ws;
            const uint32_t hi = (src >= 0 && src < n) ? w[src] : 0u;
            const uint32_t lo = (src - 1 >= 0 && src - 1 < n) ? w[src - 1] : 0u;
            w[i] = bs ? ((hi << bs) | (lo >> (32 - bs))) : hi;
        }
        n = nn;
        while (n > 0 && w[n - 1] == 0) --n;
        return true;
    }
    LGW_HD int cmp(const Big& o) const {
        int a = n, b = o.n; if (a != b) return a < b ? -1 : 1;
        for (int i = a - 1; i >= 0; --i) if (w[i] != o.w[i]) return w[i] < o.w[i] ? -1 : 1;
        return 0;
    }
    LGW_HD void sub(const Big& o) {               // this -= o (this >= o)
        int64_t br = 0;
        for (int i = 0; i < n; ++i) { int64_t t = (int64_t)w[i] - (i < o.n ? o.w[i] : 0) - br; br = t < 0; w[i] = (uint32_t)t; }
        while (n > 0 && w[n - 1] == 0) --n;
    }
    LGW_HD void shl1() {                          // *= 2 (caller guarantees room)
        uint32_t c = 0;
        for (int i = 0; i < n; ++i) { uint32_t t = w[i]; w[i] = (t << 1) | c; c = t >> 31; }
        if (c && n < LGW_BIG_LIMBS) w[n++] = c;
    }
};

// round (sig * 2^e2, plus sticky) to double; sig holds >= 55 significant bits in a u64 with the
// top bit set at position 63.
LGW_HD uint64_t round_pack(uint64_t sig, int e2 /* value = sig * 2^e2 */, bool sticky) {
    // target: 53-bit mantissa.  unbiased exponent of the top bit = e2 + 63
    int top = e2 + 63;                       // floor(log2(value))
    int shift = 11;                          // drop 11 bits to keep 53
    if (top < -1022) shift += (-1022 - top); // subnormal: keep fewer bits
    if (shift > 64) return 0;
    uint64_t kept, half, rest;
    if (shift == 64) { kept = 0; half = sig >> 63; rest = sig & 0x7FFFFFFFFFFFFFFFull; }
    else { kept = sig >> shift; half = (sig >> (shift - 1)) & 1; rest = sig & ((1ull << (shift - 1)) - 1); }
    if (half && (rest || sticky || (kept & 1))) ++kept;
    if (top < -1022) {
        // subnormal (or rounds up to the smallest normal: kept == 2^52 encodes exactly that)
        return kept;
    }
    if (kept >> 53) { kept >>= 1; ++top; }
    if (top > 1023) return 0x7FF0000000000000ull;
    return ((uint64_t)(top + 1023) << 52) | (kept & 0x000FFFFFFFFFFFFFull);
}

LGW_HD_NOINLINE bool exact_convert(uint64_t man, int exp10, uint64_t& out_bits) {
    if (man == 0) { out_bits = 0; return true; }
    if (exp10 > 330) { out_bits = 0x7FF0000000000000ull; return true; }
    if (exp10 < -400) { out_bits = 0; return true; }
    Big num; num.set64(man);
    if (exp10 >= 0) {
        for (int i = 0; i < exp10; ++i) if (!num.mul_small(10)) { out_bits = 0x7FF0000000000000ull; return true; }
        const int bl = num.bitlen();
        // take the top 64 bits
        uint64_t sig = 0;
        for (int i = 0; i < 64; ++i) sig = (sig << 1) | (uint64_t)(bl - 1 - i >= 0 ? num.bit(bl - 1 - i) : 0);
        const bool sticky = bl > 64 ? num.any_below(bl - 64) : false;
        out_bits = round_pack(sig, bl - 64, sticky);
        return true;
    }
    // value = man / 10^k: binary long division producing 64 quotient bits after alignment
    const int k = -exp10;
    Big den; den.set64(1);
    for (int i = 0; i < k; ++i) if (!den.mul_small(10)) return false;
    // align so that num >= den and num < 2*den  (track the binary exponent)
    const int bn = num.bitlen(), bd = den.bitlen();
    int e2 = 0;
    if (bd > bn) { if (!num.shl(bd - bn)) return false; e2 = -(bd - bn); }
    else if (bn > bd) { if (!den.shl(bn - bd)) return false; e2 = bn - bd; }
    if (num.cmp(den) < 0) { num.shl1(); --e2; }
    // now 1 <= num/den < 2 ; quotient bit 63 is 1
    uint64_t sig = 0;
    for (int i = 0; i < 64; ++i) {
        sig <<= 1;
        if (num.cmp(den) >= 0) { num.sub(den); sig |= 1; }
        num.shl1();
    }
    const bool sticky = num.n != 0;
    out_bits = round_pack(sig, e2 - 63, sticky);
    return true;
}

#define LGW_P10_LIST {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22}
#if defined(__CUDACC__)
__device__ __constant__ double g_p10_dev[23] = LGW_P10_LIST;      // (a local array would be rebuilt on the stack by every call)
#endif
static const double g_p10_host[23] = LGW_P10_LIST;

LGW_HD bool dec_to_double(uint64_t man, int exp10, uint64_t& bits) {
    // Clinger: both operands exact doubles
    if (man < (1ull << 53) && exp10 >= -22 && exp10 <= 22) {
#if defined(__CUDA_ARCH__)
        const double* p10 = g_p10_dev;
#else
        const double* p10 = g_p10_host;
#endif
        double d = (double)man;
        d = exp10 < 0 ? d / p10[-exp10] : d * p10[exp10];
        bits = dbl2bits(d);
        return true;
    }
    if (eisel_lemire(man, exp10, bits)) return true;
    return exact_convert(man, exp10, bits);
}

// ---- digit accumulator -----------------------------------------------------------------------
struct DecAcc {
    uint64_t mant;       // first <= 19 significant digits
    int32_t nd;          // how many digits are in mant
    int32_t e10;         // decimal exponent to apply to mant (from position of the point / dropped digits)
    int32_t expv;        // explicit exponent magnitude (clamped)
    uint8_t neg, exp_neg, is_float, nonzero, trunc;

    LGW_HD void reset() { mant = 0; nd = 0; e10 = 0; expv = 0; neg = exp_neg = is_float = nonzero = trunc = 0; }
    LGW_HD void digit(uint32_t d, bool frac) {
        nonzero |= (d != 0);
        if (nd < 19) {
            mant = mant * 10 + d;
            if (mant) ++nd;
            if (frac) --e10;
        } else {
            if (!frac) ++e10;
            if (d) trunc = 1;
        }
    }
    LGW_HD void exp_digit(uint32_t d) { if (expv < 100000) expv = expv * 10 + (int32_t)d; }

    // kind / bits / truthiness as CPython sees the literal
    LGW_HD void finish(uint8_t& kind, int64_t& bits, bool& truthy) const {
        if (!is_float) {
            truthy = nonzero;
            if (e10 > 0) { kind = LGW_KD_BIG; bits = neg ? -1 : 1; return; }           // > 19 digits
            if (neg) {
                if (mant <= (1ull << 63)) { kind = LGW_KD_INT; bits = (int64_t)(0 - mant); }
                else { kind = LGW_KD_BIG; bits = -1; }
            } else {
                if (mant <= 0x7FFFFFFFFFFFFFFFull) { kind = LGW_KD_INT; bits = (int64_t)mant; }
                else { kind = LGW_KD_BIG; bits = 1; }
            }
            return;
        }
        int64_t e = (int64_t)e10 + (exp_neg ? -(int64_t)expv : (int64_t)expv);
        if (e > 100000) e = 100000;
        if (e < -100000) e = -100000;
        uint64_t b = 0; bool ok = dec_to_double(mant, (int)e, b);
        if (ok && trunc) {
            uint64_t b2 = 0;
            ok = dec_to_double(mant + 1, (int)e, b2) && b2 == b;
        }
        kind = ok ? LGW_KD_FLT : LGW_KD_FLT_INEXACT;
        if (neg) b |= 0x8000000000000000ull;
        bits = (int64_t)b;
        truthy = (b & 0x7FFFFFFFFFFFFFFFull) != 0;
    }
};

}  // namespace lgw
