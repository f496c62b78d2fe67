// numparse.cuh -- decimal text -> binary64, correctly rounded (the value Python's float() returns), usable on
// the host and on the device.  Used by the HITRAN CSV ingestion kernel (k5_ingest.cuh): the reference parses
// every field with float(cell) (pyradUtilities.py:421-448), so bit-exact parity needs exact conversion.
//
// Algorithm: Eisel-Lemire (D. Lemire, "Number parsing at a gigabyte per second", 2021): the decimal
// significand w (up to 19 digits, exact in 64 bits) times a 128-bit approximation of 5^q decides all but the
// provably-safe cases; the table is generated exactly by scripts/gen_pow5_table.py.  Inputs outside the grammar
//   [ws] [+-] digits [. digits] [(e|E) [+-] digits] [ws]      (at least one digit, at most 19 significant digits)
// are reported as unparsable -- the caller fails loudly, as float() would raise in the reference.
#pragma once
#include <stdint.h>
#include "pow5_table.h"

namespace prb {

static const uint64_t POW5_HOST[] = PRB_POW5_TABLE_INIT;
#ifdef __CUDACC__
__device__ const uint64_t POW5_DEV[] = PRB_POW5_TABLE_INIT;
#define PRB_HD __host__ __device__ __forceinline__
#else
#define PRB_HD inline
#endif

PRB_HD uint64_t pow5_word(int i) {
#ifdef __CUDA_ARCH__
    return POW5_DEV[i];
#else
    return POW5_HOST[i];
#endif
}

struct U128 { uint64_t lo, hi; };

PRB_HD U128 mul64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    return U128{a * b, __umul64hi(a, b)};
#else
    const unsigned __int128 p = (unsigned __int128)a * b;
    return U128{(uint64_t)p, (uint64_t)(p >> 64)};
#endif
}

PRB_HD int clz64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}

// w * 10^q -> the IEEE-754 binary64 bit pattern (sign excluded), round to nearest even.  w != 0 handled by caller.
PRB_HD uint64_t decimal_to_bits(uint64_t w, int64_t q) {
    if (w == 0 || q < PRB_POW5_MIN_Q) return 0;
    if (q > PRB_POW5_MAX_Q) return 0x7FF0000000000000ULL;
    const int lz = clz64(w);
    w <<= lz;
    const int index = 2 * (int)(q - PRB_POW5_MIN_Q);
    U128 prod = mul64(w, pow5_word(index));
    const uint64_t precision_mask = 0xFFFFFFFFFFFFFFFFULL >> 55;                 // 52 + 3 bits of precision wanted
    if ((prod.hi & precision_mask) == precision_mask) {
        const U128 second = mul64(w, pow5_word(index + 1));
        prod.lo += second.hi;
        if (second.hi > prod.lo) prod.hi++;
    }
    const int upperbit = (int)(prod.hi >> 63);
    const int shift = upperbit + 64 - 52 - 3;
    uint64_t mantissa = prod.hi >> shift;
    int64_t power2 = (((152170 + 65536) * q) >> 16) + 63 + upperbit - lz + 1023;
    if (power2 <= 0) {                                                            // subnormal
        if (-power2 + 1 >= 64) return 0;
        mantissa >>= -power2 + 1;
        mantissa += (mantissa & 1);
        mantissa >>= 1;
        power2 = (mantissa < (1ULL << 52)) ? 0 : 1;
        return (mantissa & ~(1ULL << 52)) | ((uint64_t)power2 << 52);
    }
    if (prod.lo <= 1 && q >= -4 && q <= 23 && (mantissa & 3) == 1) {              // exactly halfway: ties to even
        if ((mantissa << shift) == prod.hi) mantissa &= ~1ULL;
    }
    mantissa += (mantissa & 1);
    mantissa >>= 1;
    if (mantissa >= (2ULL << 52)) {
        mantissa = 1ULL << 52;
        power2++;
    }
    mantissa &= ~(1ULL << 52);
    if (power2 >= 0x7FF) return 0x7FF0000000000000ULL;
    return mantissa | ((uint64_t)power2 << 52);
}

PRB_HD bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\f' || c == '\v'; }

// Parse [p, end) as one decimal number.  Returns false when the text is not a plain decimal literal.  More than 19
// significant digits (what fits a 64-bit mantissa): the first 19 are kept and the rest only recorded as "zero" or "not
// zero"; the true value then lies in [w, w + 1) x 10^e, so when both ends round to the same double that double is the
// correctly rounded result (Python's float()) -- always the case for trailing zeros, and for all but ~1 in 10^3 random
// long literals.  The rare undecided literal is reported as unparsable rather than rounded by guess.
PRB_HD bool parse_double(const char *p, const char *end, double *out) {
    while (p < end && is_ws(*p)) ++p;
    while (end > p && is_ws(end[-1])) --end;
    if (p >= end) return false;
    bool neg = false;
    if (*p == '+' || *p == '-') { neg = *p == '-'; ++p; }
    uint64_t w = 0;
    int sig = 0;                 // significant digits accumulated into w
    int64_t exp10 = 0;
    int digits = 0;
    bool dropped_nonzero = false;   // a digit beyond the 19th was not '0'
    for (; p < end && *p >= '0' && *p <= '9'; ++p, ++digits) {
        if (w != 0 || *p != '0') {
            if (sig >= 19) {        // integer digit beyond the mantissa: scales the value by ten
                dropped_nonzero |= *p != '0';
                ++exp10;
            } else {
                w = w * 10 + (uint64_t)(*p - '0');
                ++sig;
            }
        }
    }
    if (p < end && *p == '.') {
        ++p;
        for (; p < end && *p >= '0' && *p <= '9'; ++p, ++digits) {
            if (w != 0 || *p != '0') {
                if (sig >= 19) {    // fraction digit beyond the mantissa: no change of scale
                    dropped_nonzero |= *p != '0';
                    continue;
                }
                w = w * 10 + (uint64_t)(*p - '0');
                ++sig;
            }
            --exp10;
        }
    }
    if (digits == 0) return false;
    if (p < end && (*p == 'e' || *p == 'E')) {
        ++p;
        bool eneg = false;
        if (p < end && (*p == '+' || *p == '-')) { eneg = *p == '-'; ++p; }
        if (p >= end || *p < '0' || *p > '9') return false;
        int64_t ev = 0;
        for (; p < end && *p >= '0' && *p <= '9'; ++p)
            if (ev < 100000) ev = ev * 10 + (*p - '0');
        exp10 += eneg ? -ev : ev;
    }
    if (p != end) return false;
    uint64_t bits = decimal_to_bits(w, exp10);
    if (dropped_nonzero && decimal_to_bits(w + 1, exp10) != bits) return false;   // the dropped digits would decide the rounding
    if (neg) bits |= 0x8000000000000000ULL;
#ifdef __CUDA_ARCH__
    *out = __longlong_as_double((long long)bits);
#else
    union { uint64_t u; double d; } cv;
    cv.u = bits;
    *out = cv.d;
#endif
    return true;
}

}  // namespace prb
