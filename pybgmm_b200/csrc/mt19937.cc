// mt19937.cc -- the stream of random.random() values that utils.draw consumes (pybgmm/utils/utils.py:15),
// produced in bulk so a whole sweep's uniforms can be handed to the device in one copy.
//
// CPython's random.random() is MT19937 (Matsumoto & Nishimura 1998) with the 53-bit assembly
//     a = genrand_uint32() >> 5;  b = genrand_uint32() >> 6;  (a * 2^26 + b) / 2^53
// and random.getstate()[1] is the 624 state words followed by the position.  This file restates that
// published algorithm; `state` round-trips with random.getstate()/setstate().
#include <stdint.h>

#include "../../include/bgmm_b200.h"

namespace {
constexpr int MT_N = 624, MT_M = 397;
constexpr uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX_A = 0x9908b0dfu;

inline void regenerate(uint32_t *mt) {
    int kk = 0;
    for (; kk < MT_N - MT_M; ++kk) {
        const uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
        mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
    }
    for (; kk < MT_N - 1; ++kk) {
        const uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
        mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
    }
    const uint32_t y = (mt[MT_N - 1] & UPPER) | (mt[0] & LOWER);
    mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
}

inline uint32_t next_u32(uint32_t *mt, uint32_t &pos) {
    if (pos >= (uint32_t)MT_N) {
        regenerate(mt);
        pos = 0;
    }
    uint32_t y = mt[pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
}  // namespace

extern "C" int bgmm_mt19937_fill(uint32_t *state, double *out, int64_t n) {
    if (!state || (n > 0 && !out) || n < 0) return BGMM_EINVAL;
    uint32_t pos = state[MT_N];
    if (pos > (uint32_t)MT_N) return BGMM_EINVAL;
    for (int64_t t = 0; t < n; ++t) {
        const uint32_t a = next_u32(state, pos) >> 5;
        const uint32_t b = next_u32(state, pos) >> 6;
        out[t] = (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
    }
    state[MT_N] = pos;
    return BGMM_OK;
}
