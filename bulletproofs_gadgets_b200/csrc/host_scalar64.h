// host_scalar64.h -- host-only fast path for the handful of scalar products / inversions the protocol drivers do between
// device steps (challenge inverses, x powers, ...): 4 x 64-bit limbs with unsigned __int128 products.
// Reduction mod l = 2^252 + c uses x = lo - c * (x >> 252), applied three times (values shrink 512 -> 385 -> 258 -> 131 bits).
#pragma once
#include "fe25519.cuh"

namespace bpgh {
typedef unsigned __int128 u128;
static const u64 L64[4] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0, 0x1000000000000000ULL};
static const u64 C64[2] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL};

template <int NA, int NB>
static inline void mul64(u64 *r, const u64 *a, const u64 *b) {
    for (int i = 0; i < NA + NB; i++) r[i] = 0;
    for (int i = 0; i < NA; i++) {
        u64 c = 0;
        for (int j = 0; j < NB; j++) { u128 t = (u128)a[i] * b[j] + r[i + j] + c; r[i + j] = (u64)t; c = (u64)(t >> 64); }
        r[i + NB] = c;
    }
}
// x: 8 limbs (512 bits) -> out: 4 limbs, fully reduced
static inline void reduce512_64(u64 out[4], const u64 x[8]) {
    const u64 M60 = 0x0FFFFFFFFFFFFFFFULL;
    u64 lo[4] = {x[0], x[1], x[2], x[3] & M60}, hi[5];
    for (int i = 0; i < 5; i++) hi[i] = (x[i + 3] >> 60) | ((i + 4 < 8 ? x[i + 4] : 0) << 4);
    u64 y[7];
    mul64<5, 2>(y, hi, C64);
    u64 ylo[4] = {y[0], y[1], y[2], y[3] & M60}, yhi[3];
    for (int i = 0; i < 3; i++) yhi[i] = (y[i + 3] >> 60) | (y[i + 4] << 4);
    u64 z[5];
    mul64<3, 2>(z, yhi, C64);
    u64 zlo[4] = {z[0], z[1], z[2], z[3] & M60};
    u64 zhi = (z[3] >> 60) | (z[4] << 4);
    u64 w[3];
    mul64<1, 2>(w, &zhi, C64);
    // acc = 2l + lo + zlo - ylo - w   (always positive, < 5l)
    u64 acc[5];
    u128 t = 0;
    const u64 twoL[5] = {L64[0] << 1, (L64[1] << 1) | (L64[0] >> 63), (L64[2] << 1) | (L64[1] >> 63), (L64[3] << 1) | (L64[2] >> 63), L64[3] >> 63};
    for (int i = 0; i < 5; i++) { t += (u128)twoL[i] + (i < 4 ? lo[i] : 0) + (i < 4 ? zlo[i] : 0); acc[i] = (u64)t; t >>= 64; }
    u64 br = 0;
    for (int i = 0; i < 5; i++) { u128 d = (u128)acc[i] - (i < 4 ? ylo[i] : 0) - br; acc[i] = (u64)d; br = (u64)(d >> 64) & 1; }
    br = 0;
    for (int i = 0; i < 5; i++) { u128 d = (u128)acc[i] - (i < 3 ? w[i] : 0) - br; acc[i] = (u64)d; br = (u64)(d >> 64) & 1; }
    for (int k = 0; k < 5; k++) {
        bool ge = acc[4] != 0;
        if (!ge) {
            ge = true;
            for (int i = 3; i >= 0; i--) { if (acc[i] > L64[i]) break; if (acc[i] < L64[i]) { ge = false; break; } }
        }
        if (!ge) break;
        u64 b2 = 0;
        for (int i = 0; i < 5; i++) { u128 d = (u128)acc[i] - (i < 4 ? L64[i] : 0) - b2; acc[i] = (u64)d; b2 = (u64)(d >> 64) & 1; }
    }
    for (int i = 0; i < 4; i++) out[i] = acc[i];
}
static inline void to64(u64 o[4], const sc &a) { for (int i = 0; i < 4; i++) o[i] = (u64)a.v[2 * i] | ((u64)a.v[2 * i + 1] << 32); }
static inline sc from64(const u64 a[4]) { sc r; for (int i = 0; i < 4; i++) { r.v[2 * i] = (u32)a[i]; r.v[2 * i + 1] = (u32)(a[i] >> 32); } return r; }
static inline sc sc_mul64(const sc &a, const sc &b) {
    u64 x[4], y[4], p[8], r[4];
    to64(x, a); to64(y, b);
    mul64<4, 4>(p, x, y);
    reduce512_64(r, p);
    return from64(r);
}
static inline sc sc_wide64(const uint8_t b[64]) {
    u64 x[8], r[4];
    for (int i = 0; i < 8; i++) { u64 v = 0; for (int j = 7; j >= 0; j--) v = (v << 8) | b[8 * i + j]; x[i] = v; }
    reduce512_64(r, x);
    return from64(r);
}
static inline sc sc_invert64(const sc &a) { // a^(l-2)
    const u64 e[4] = {L64[0] - 2, L64[1], L64[2], L64[3]};
    sc one; for (int i = 0; i < 8; i++) one.v[i] = 0; one.v[0] = 1;
    sc base = sc_mul64(a, one), acc = one;
    for (int i = 252; i >= 0; i--) {
        acc = sc_mul64(acc, acc);
        if ((e[i >> 6] >> (i & 63)) & 1) acc = sc_mul64(acc, base);
    }
    return acc;
}
} // namespace bpgh
