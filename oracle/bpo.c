/* oracle/bpo.c -- C CPU oracle for the Bulletproofs R1CS hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * See oracle/bpo.h for the scope statement.  This file restates, on the CPU, the algorithms the
 * reference reaches through un-vendored crates (curve25519-dalek 1.x, merlin 1.x, the
 * bulletproofs `develop` fork: /root/reference/Cargo.toml:8,10,17-20) plus the in-tree MiMC
 * (/root/reference/src/mimc_hash/mimc.rs:7-97).  The published algorithms followed are
 * RFC 8032 / RFC 9496 (field, Edwards, ristretto255), FIPS 202 (Keccak), the STROBE-128 /
 * Merlin specification, and the Bulletproofs R1CS protocol as restated in SURVEY.md App. A.
 * It is validated against oracle/pyref.py (big-int Python), RFC 9496 vectors, the Merlin test
 * vector and the reference's MiMC/Merkle known answers (tests/test_oracle.py).
 *
 * Parity status: MiMC/Merkle pinned by the reference's own KATs; MSM / Pedersen / IPP / proof
 * bytes are "parity unpinned" at the reference boundary (the reference holds no golden bytes).
 *
 * Representation: GF(2^255-19) in 5 x 51-bit limbs (unsigned __int128 products); scalars mod l
 * as 4 x 64-bit limbs.  Data-parallel loops use OpenMP when bpo_set_threads(n>1) was called.
 */
#include "bpo.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef uint8_t u8;

static int g_threads = 1;
void bpo_set_threads(int n) { g_threads = n < 1 ? 1 : n; }

/* ============================================================ field GF(2^255-19) */
typedef struct { u64 v[5]; } fe;
#define M51 0x7FFFFFFFFFFFFULL

static const fe FE_ZERO = {{0, 0, 0, 0, 0}};
static const fe FE_ONE = {{1, 0, 0, 0, 0}};

static inline void fe_carry(fe *h) {
    u64 c;
    c = h->v[0] >> 51; h->v[0] &= M51; h->v[1] += c;
    c = h->v[1] >> 51; h->v[1] &= M51; h->v[2] += c;
    c = h->v[2] >> 51; h->v[2] &= M51; h->v[3] += c;
    c = h->v[3] >> 51; h->v[3] &= M51; h->v[4] += c;
    c = h->v[4] >> 51; h->v[4] &= M51; h->v[0] += 19 * c;
    c = h->v[0] >> 51; h->v[0] &= M51; h->v[1] += c;
}
static inline void fe_add(fe *h, const fe *f, const fe *g) {
    for (int i = 0; i < 5; i++) h->v[i] = f->v[i] + g->v[i];
    fe_carry(h);
}
static inline void fe_sub(fe *h, const fe *f, const fe *g) {
    /* + 4p keeps limbs non-negative for g limbs < 2^52 */
    h->v[0] = f->v[0] + 0x1FFFFFFFFFFFB4ULL - g->v[0];
    for (int i = 1; i < 5; i++) h->v[i] = f->v[i] + 0x1FFFFFFFFFFFFCULL - g->v[i];
    fe_carry(h);
}
static inline void fe_neg(fe *h, const fe *f) { fe_sub(h, &FE_ZERO, f); }

static inline void fe_mul(fe *h, const fe *f, const fe *g) {
    u64 f0 = f->v[0], f1 = f->v[1], f2 = f->v[2], f3 = f->v[3], f4 = f->v[4];
    u64 g0 = g->v[0], g1 = g->v[1], g2 = g->v[2], g3 = g->v[3], g4 = g->v[4];
    u64 g1_19 = 19 * g1, g2_19 = 19 * g2, g3_19 = 19 * g3, g4_19 = 19 * g4;
    u128 r0 = (u128)f0 * g0 + (u128)f1 * g4_19 + (u128)f2 * g3_19 + (u128)f3 * g2_19 + (u128)f4 * g1_19;
    u128 r1 = (u128)f0 * g1 + (u128)f1 * g0 + (u128)f2 * g4_19 + (u128)f3 * g3_19 + (u128)f4 * g2_19;
    u128 r2 = (u128)f0 * g2 + (u128)f1 * g1 + (u128)f2 * g0 + (u128)f3 * g4_19 + (u128)f4 * g3_19;
    u128 r3 = (u128)f0 * g3 + (u128)f1 * g2 + (u128)f2 * g1 + (u128)f3 * g0 + (u128)f4 * g4_19;
    u128 r4 = (u128)f0 * g4 + (u128)f1 * g3 + (u128)f2 * g2 + (u128)f3 * g1 + (u128)f4 * g0;
    u64 c;
    r1 += (u64)(r0 >> 51); u64 h0 = (u64)r0 & M51;
    r2 += (u64)(r1 >> 51); u64 h1 = (u64)r1 & M51;
    r3 += (u64)(r2 >> 51); u64 h2 = (u64)r2 & M51;
    r4 += (u64)(r3 >> 51); u64 h3 = (u64)r3 & M51;
    c = (u64)(r4 >> 51);   u64 h4 = (u64)r4 & M51;
    h0 += 19 * c;
    c = h0 >> 51; h0 &= M51; h1 += c;
    h->v[0] = h0; h->v[1] = h1; h->v[2] = h2; h->v[3] = h3; h->v[4] = h4;
}
static inline void fe_sq(fe *h, const fe *f) { fe_mul(h, f, f); }
static inline void fe_mul_small(fe *h, const fe *f, u64 s) {
    u128 r; u64 c = 0;
    for (int i = 0; i < 5; i++) { r = (u128)f->v[i] * s + c; h->v[i] = (u64)r & M51; c = (u64)(r >> 51); }
    h->v[0] += 19 * c;
    fe_carry(h);
}
static void fe_frombytes(fe *h, const u8 s[32]) {
    u64 w[4];
    for (int i = 0; i < 4; i++) { w[i] = 0; for (int j = 7; j >= 0; j--) w[i] = (w[i] << 8) | s[8 * i + j]; }
    h->v[0] = w[0] & M51;
    h->v[1] = ((w[0] >> 51) | (w[1] << 13)) & M51;
    h->v[2] = ((w[1] >> 38) | (w[2] << 26)) & M51;
    h->v[3] = ((w[2] >> 25) | (w[3] << 39)) & M51;
    h->v[4] = (w[3] >> 12) & M51; /* bit 255 dropped */
}
static void fe_tobytes(u8 s[32], const fe *f) {
    fe t = *f;
    fe_carry(&t); fe_carry(&t);
    /* t < 2^255 + small; compute q = floor((t + 19) / 2^255) */
    u64 q = (t.v[0] + 19) >> 51;
    q = (t.v[1] + q) >> 51; q = (t.v[2] + q) >> 51; q = (t.v[3] + q) >> 51; q = (t.v[4] + q) >> 51;
    t.v[0] += 19 * q;
    u64 c;
    c = t.v[0] >> 51; t.v[0] &= M51; t.v[1] += c;
    c = t.v[1] >> 51; t.v[1] &= M51; t.v[2] += c;
    c = t.v[2] >> 51; t.v[2] &= M51; t.v[3] += c;
    c = t.v[3] >> 51; t.v[3] &= M51; t.v[4] += c;
    t.v[4] &= M51;
    u64 w[4];
    w[0] = t.v[0] | (t.v[1] << 51);
    w[1] = (t.v[1] >> 13) | (t.v[2] << 38);
    w[2] = (t.v[2] >> 26) | (t.v[3] << 25);
    w[3] = (t.v[3] >> 39) | (t.v[4] << 12);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) s[8 * i + j] = (u8)(w[i] >> (8 * j));
}
static int fe_iszero(const fe *f) { u8 s[32]; fe_tobytes(s, f); u8 r = 0; for (int i = 0; i < 32; i++) r |= s[i]; return r == 0; }
static int fe_isneg(const fe *f) { u8 s[32]; fe_tobytes(s, f); return s[0] & 1; }
static int fe_eq(const fe *a, const fe *b) { u8 x[32], y[32]; fe_tobytes(x, a); fe_tobytes(y, b); return memcmp(x, y, 32) == 0; }
static void fe_abs(fe *h, const fe *f) { if (fe_isneg(f)) fe_neg(h, f); else *h = *f; }
static void fe_sqn(fe *h, const fe *f, int n) { *h = *f; for (int i = 0; i < n; i++) fe_sq(h, h); }
/* z^(2^250-1) and z^11 helper for invert / pow22523 */
static void fe_pow_2_250_1(fe *out, fe *z11, const fe *z) {
    fe z2, z9, t, z_5_0, z_10_0, z_20_0, z_50_0, z_100_0;
    fe_sq(&z2, z); fe_sqn(&t, &z2, 2); fe_mul(&z9, &t, z); fe_mul(z11, &z9, &z2);
    fe_sq(&t, z11); fe_mul(&z_5_0, &t, &z9);
    fe_sqn(&t, &z_5_0, 5); fe_mul(&z_10_0, &t, &z_5_0);
    fe_sqn(&t, &z_10_0, 10); fe_mul(&z_20_0, &t, &z_10_0);
    fe_sqn(&t, &z_20_0, 20); fe_mul(&t, &t, &z_20_0);
    fe_sqn(&t, &t, 10); fe_mul(&z_50_0, &t, &z_10_0);
    fe_sqn(&t, &z_50_0, 50); fe_mul(&z_100_0, &t, &z_50_0);
    fe_sqn(&t, &z_100_0, 100); fe_mul(&t, &t, &z_100_0);
    fe_sqn(&t, &t, 50); fe_mul(out, &t, &z_50_0);
}
static void fe_invert(fe *out, const fe *z) { fe t, z11; fe_pow_2_250_1(&t, &z11, z); fe_sqn(&t, &t, 5); fe_mul(out, &t, &z11); }
static void fe_pow22523(fe *out, const fe *z) { fe t, z11; fe_pow_2_250_1(&t, &z11, z); fe_sqn(&t, &t, 2); fe_mul(out, &t, z); }

static void fe_from_le_hex(fe *h, const char *hex_be) { /* 64 hex chars, big endian integer */
    u8 b[32];
    for (int i = 0; i < 32; i++) {
        unsigned v = 0;
        for (int k = 0; k < 2; k++) { char ch = hex_be[2 * i + k]; v = v * 16 + (ch <= '9' ? ch - '0' : (ch | 32) - 'a' + 10); }
        b[31 - i] = (u8)v;
    }
    fe_frombytes(h, b);
}

/* curve constants (values cross-checked against pyref.py in tests) */
static fe C_D, C_D2, C_SQRT_M1, C_INVSQRT_A_MINUS_D, C_SQRT_AD_MINUS_ONE, C_ONE_MINUS_D_SQ, C_D_MINUS_ONE_SQ;

/* RFC 9496 4.2 SQRT_RATIO_M1 */
static int fe_sqrt_ratio_m1(fe *r_out, const fe *u, const fe *v) {
    fe v3, v7, r, check, t, neg_u, neg_u_i;
    fe_sq(&t, v); fe_mul(&v3, &t, v);
    fe_sq(&t, &v3); fe_mul(&v7, &t, v);
    fe_mul(&t, u, &v7); fe_pow22523(&t, &t);
    fe_mul(&r, u, &v3); fe_mul(&r, &r, &t);
    fe_sq(&t, &r); fe_mul(&check, v, &t);
    fe_neg(&neg_u, u); fe_mul(&neg_u_i, &neg_u, &C_SQRT_M1);
    int correct = fe_eq(&check, u), flipped = fe_eq(&check, &neg_u), flipped_i = fe_eq(&check, &neg_u_i);
    if (flipped || flipped_i) fe_mul(&r, &r, &C_SQRT_M1);
    fe_abs(r_out, &r);
    return correct || flipped;
}

/* ============================================================ scalars mod l */
typedef struct { u64 v[4]; } sc;
static const u64 SC_L[4] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0, 0x1000000000000000ULL};
static const u64 SC_C[2] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL}; /* l - 2^252 */
static const sc SC_ZERO = {{0, 0, 0, 0}};
static const sc SC_ONE = {{1, 0, 0, 0}};

static void sc_load(sc *s, const u8 b[32]) { for (int i = 0; i < 4; i++) { u64 w = 0; for (int j = 7; j >= 0; j--) w = (w << 8) | b[8 * i + j]; s->v[i] = w; } }
static void sc_store(u8 b[32], const sc *s) { for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) b[8 * i + j] = (u8)(s->v[i] >> (8 * j)); }
static int sc_geq_l(const u64 a[4]) { for (int i = 3; i >= 0; i--) { if (a[i] > SC_L[i]) return 1; if (a[i] < SC_L[i]) return 0; } return 1; }
static void sc_sub_l(u64 a[4]) { u64 br = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)a[i] - SC_L[i] - br; a[i] = (u64)d; br = (u64)(d >> 64) & 1; } }
/* generic little-endian multi-limb helpers */
static void mp_mul(u64 *r, const u64 *a, int na, const u64 *b, int nb) {
    for (int i = 0; i < na + nb; i++) r[i] = 0;
    for (int i = 0; i < na; i++) { u64 c = 0; for (int j = 0; j < nb; j++) { u128 t = (u128)a[i] * b[j] + r[i + j] + c; r[i + j] = (u64)t; c = (u64)(t >> 64); } r[i + nb] = c; }
}
/* reduce an n-limb (n<=8) little-endian integer mod l using l = 2^252 + c */
static void sc_reduce_limbs(sc *out, const u64 *x, int n) {
    u64 lo[4] = {0, 0, 0, 0}, hi[5] = {0, 0, 0, 0, 0};
    u64 xx[9]; for (int i = 0; i < 9; i++) xx[i] = i < n ? x[i] : 0;
    for (int i = 0; i < 4; i++) lo[i] = xx[i]; lo[3] &= 0x0FFFFFFFFFFFFFFFULL;
    for (int i = 0; i < 5; i++) hi[i] = (xx[i + 3] >> 60) | (xx[i + 4] << 4); /* x >> 252, <= 260 bits */
    /* x = lo - c*hi ; y = c*hi (<= 385 bits, 7 limbs) */
    u64 y[7]; mp_mul(y, hi, 5, SC_C, 2);
    u64 ylo[4], yhi[3];
    for (int i = 0; i < 4; i++) ylo[i] = y[i]; ylo[3] &= 0x0FFFFFFFFFFFFFFFULL;
    for (int i = 0; i < 3; i++) yhi[i] = (y[i + 3] >> 60) | ((i + 4 < 7 ? y[i + 4] : 0) << 4); /* <= 133 bits */
    u64 z[5]; mp_mul(z, yhi, 3, SC_C, 2); /* <= 258 bits */
    u64 zlo[4], zhi;
    for (int i = 0; i < 4; i++) zlo[i] = z[i]; zlo[3] &= 0x0FFFFFFFFFFFFFFFULL;
    zhi = (z[3] >> 60) | (z[4] << 4); /* <= 6 bits */
    u64 w[3]; mp_mul(w, &zhi, 1, SC_C, 2); /* c*zhi <= 131 bits */
    /* result = lo - ylo + zlo - w  (mod l); compute in signed 5-limb arithmetic with +2l bias */
    u64 acc[5] = {0, 0, 0, 0, 0};
    u64 twol[5] = {0, 0, 0, 0, 0};
    { u64 c = 0; for (int i = 0; i < 4; i++) { u128 t = (u128)SC_L[i] * 2 + c; twol[i] = (u64)t; c = (u64)(t >> 64); } twol[4] = c; }
    { u64 c = 0; for (int i = 0; i < 5; i++) { u128 t = (u128)twol[i] + (i < 4 ? lo[i] : 0) + c; acc[i] = (u64)t; c = (u64)(t >> 64); } }
    { u64 c = 0; for (int i = 0; i < 5; i++) { u128 t = (u128)acc[i] + (i < 4 ? zlo[i] : 0) + c; acc[i] = (u64)t; c = (u64)(t >> 64); } }
    { u64 b = 0; for (int i = 0; i < 5; i++) { u128 t = (u128)acc[i] - (i < 4 ? ylo[i] : 0) - b; acc[i] = (u64)t; b = (u64)(t >> 64) & 1; } }
    { u64 b = 0; for (int i = 0; i < 5; i++) { u128 t = (u128)acc[i] - (i < 3 ? w[i] : 0) - b; acc[i] = (u64)t; b = (u64)(t >> 64) & 1; } }
    /* 0 < acc < 4l + ...; subtract l while >= l (acc[4] may be nonzero) */
    for (int k = 0; k < 6; k++) {
        if (acc[4] || sc_geq_l(acc)) { u64 b = 0; for (int i = 0; i < 5; i++) { u128 t = (u128)acc[i] - (i < 4 ? SC_L[i] : 0) - b; acc[i] = (u64)t; b = (u64)(t >> 64) & 1; } }
    }
    for (int i = 0; i < 4; i++) out->v[i] = acc[i];
}
static void sc_reduce(sc *r, const sc *a) { sc_reduce_limbs(r, a->v, 4); }
static void sc_mul(sc *r, const sc *a, const sc *b) { u64 t[8]; mp_mul(t, a->v, 4, b->v, 4); sc_reduce_limbs(r, t, 8); }
static void sc_add(sc *r, const sc *a, const sc *b) {
    u64 t[5]; u64 c = 0;
    for (int i = 0; i < 4; i++) { u128 s = (u128)a->v[i] + b->v[i] + c; t[i] = (u64)s; c = (u64)(s >> 64); }
    t[4] = c; sc_reduce_limbs(r, t, 5);
}
static void sc_neg(sc *r, const sc *a) { /* l - (a mod l) */
    sc t; sc_reduce(&t, a);
    if ((t.v[0] | t.v[1] | t.v[2] | t.v[3]) == 0) { *r = t; return; }
    u64 br = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)SC_L[i] - t.v[i] - br; r->v[i] = (u64)d; br = (u64)(d >> 64) & 1; }
}
static void sc_sub(sc *r, const sc *a, const sc *b) { sc nb; sc_neg(&nb, b); sc_add(r, a, &nb); }
static void sc_muladd(sc *r, const sc *a, const sc *b, const sc *c) { sc t; sc_mul(&t, a, b); sc_add(r, &t, c); }
static void sc_from_u64(sc *r, u64 x) { r->v[0] = x; r->v[1] = r->v[2] = r->v[3] = 0; }
static int sc_iszero(const sc *a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }
static void sc_invert(sc *r, const sc *a) { /* a^(l-2), square-and-multiply */
    u64 e[4] = {SC_L[0] - 2, SC_L[1], SC_L[2], SC_L[3]};
    sc base, acc = SC_ONE; sc_reduce(&base, a);
    for (int i = 252; i >= 0; i--) { sc_mul(&acc, &acc, &acc); if ((e[i >> 6] >> (i & 63)) & 1) sc_mul(&acc, &acc, &base); }
    *r = acc;
}
static void sc_wide(sc *r, const u8 b[64]) { u64 t[8]; for (int i = 0; i < 8; i++) { u64 w = 0; for (int j = 7; j >= 0; j--) w = (w << 8) | b[8 * i + j]; t[i] = w; } sc_reduce_limbs(r, t, 8); }

void bpo_sc_reduce(const u8 in[32], u8 out[32]) { sc a, r; sc_load(&a, in); sc_reduce(&r, &a); sc_store(out, &r); }
void bpo_sc_wide(const u8 in[64], u8 out[32]) { sc r; sc_wide(&r, in); sc_store(out, &r); }
void bpo_sc_mul(const u8 a[32], const u8 b[32], u8 out[32]) { sc x, y, r; sc_load(&x, a); sc_load(&y, b); sc_mul(&r, &x, &y); sc_store(out, &r); }
void bpo_sc_add(const u8 a[32], const u8 b[32], u8 out[32]) { sc x, y, r; sc_load(&x, a); sc_load(&y, b); sc_add(&r, &x, &y); sc_store(out, &r); }
void bpo_sc_invert(const u8 a[32], u8 out[32]) { sc x, r; sc_load(&x, a); sc_invert(&r, &x); sc_store(out, &r); }

/* ============================================================ Edwards / ristretto255 */
typedef struct { fe X, Y, Z, T; } ge;          /* extended */
typedef struct { fe YpX, YmX, Z, T2d; } ge_c;  /* "projective Niels" / cached */

static void ge_identity(ge *p) { p->X = FE_ZERO; p->Y = FE_ONE; p->Z = FE_ONE; p->T = FE_ZERO; }
static void ge_to_cached(ge_c *c, const ge *p) { fe_add(&c->YpX, &p->Y, &p->X); fe_sub(&c->YmX, &p->Y, &p->X); c->Z = p->Z; fe_mul(&c->T2d, &p->T, &C_D2); }
static void ge_neg(ge *r, const ge *p) { fe_neg(&r->X, &p->X); r->Y = p->Y; r->Z = p->Z; fe_neg(&r->T, &p->T); }
/* add-2008-hwcd-3 (complete for a=-1); sign=-1 subtracts */
static void ge_add_cached(ge *r, const ge *p, const ge_c *q, int sign) {
    fe a, b, c, d, e, f, g, h, t;
    fe_sub(&t, &p->Y, &p->X); fe_mul(&a, &t, sign > 0 ? &q->YmX : &q->YpX);
    fe_add(&t, &p->Y, &p->X); fe_mul(&b, &t, sign > 0 ? &q->YpX : &q->YmX);
    fe_mul(&c, &p->T, &q->T2d); if (sign < 0) fe_neg(&c, &c);
    fe_mul(&d, &p->Z, &q->Z); fe_add(&d, &d, &d);
    fe_sub(&e, &b, &a); fe_sub(&f, &d, &c); fe_add(&g, &d, &c); fe_add(&h, &b, &a);
    fe_mul(&r->X, &e, &f); fe_mul(&r->Y, &g, &h); fe_mul(&r->Z, &f, &g); fe_mul(&r->T, &e, &h);
}
static void ge_add(ge *r, const ge *p, const ge *q) { ge_c c; ge_to_cached(&c, q); ge_add_cached(r, p, &c, 1); }
static void ge_sub(ge *r, const ge *p, const ge *q) { ge_c c; ge_to_cached(&c, q); ge_add_cached(r, p, &c, -1); }
/* dbl-2008-hwcd, a=-1 */
static void ge_dbl(ge *r, const ge *p) {
    fe a, b, c, d, e, f, g, h, t;
    fe_sq(&a, &p->X); fe_sq(&b, &p->Y); fe_sq(&c, &p->Z); fe_add(&c, &c, &c);
    fe_neg(&d, &a);
    fe_add(&t, &p->X, &p->Y); fe_sq(&t, &t); fe_sub(&e, &t, &a); fe_sub(&e, &e, &b);
    fe_add(&g, &d, &b); fe_sub(&f, &g, &c); fe_sub(&h, &d, &b);
    fe_mul(&r->X, &e, &f); fe_mul(&r->Y, &g, &h); fe_mul(&r->Z, &f, &g); fe_mul(&r->T, &e, &h);
}
static void ge_scalarmul(ge *r, const sc *k, const ge *p) { /* plain double-and-add on reduced k */
    sc kr; sc_reduce(&kr, k);
    ge acc; ge_identity(&acc);
    for (int i = 252; i >= 0; i--) { ge_dbl(&acc, &acc); if ((kr.v[i >> 6] >> (i & 63)) & 1) ge_add(&acc, &acc, p); }
    *r = acc;
}
static int ge_is_identity_coset(const ge *p) { fe t; fe_mul(&t, &p->X, &p->Y); return fe_iszero(&t); }

static int ristretto_decode(ge *p, const u8 b[32]) {
    fe s, ss, u1, u2, u2s, v, t, I, Dx, Dy, x, y;
    u8 chk[32];
    fe_frombytes(&s, b); fe_tobytes(chk, &s);
    if (memcmp(chk, b, 32) != 0 || (b[0] & 1)) return 0; /* non-canonical (>= p or bit 255) or negative */
    fe_sq(&ss, &s); fe_sub(&u1, &FE_ONE, &ss); fe_add(&u2, &FE_ONE, &ss); fe_sq(&u2s, &u2);
    fe_sq(&t, &u1); fe_mul(&t, &t, &C_D); fe_neg(&t, &t); fe_sub(&v, &t, &u2s);
    fe_mul(&t, &v, &u2s);
    int ok = fe_sqrt_ratio_m1(&I, &FE_ONE, &t);
    fe_mul(&Dx, &I, &u2); fe_mul(&Dy, &I, &Dx); fe_mul(&Dy, &Dy, &v);
    fe_mul(&x, &s, &Dx); fe_add(&x, &x, &x); fe_abs(&x, &x);
    fe_mul(&y, &u1, &Dy); fe_mul(&t, &x, &y);
    if (!ok || fe_isneg(&t) || fe_iszero(&y)) return 0;
    p->X = x; p->Y = y; p->Z = FE_ONE; p->T = t;
    return 1;
}
static void ristretto_encode(u8 out[32], const ge *p) {
    fe u1, u2, t, I, d1, d2, zinv, X = p->X, Y = p->Y, den, s;
    fe_add(&u1, &p->Z, &p->Y); fe_sub(&t, &p->Z, &p->Y); fe_mul(&u1, &u1, &t);
    fe_mul(&u2, &p->X, &p->Y);
    fe_sq(&t, &u2); fe_mul(&t, &t, &u1);
    fe_sqrt_ratio_m1(&I, &FE_ONE, &t);
    fe_mul(&d1, &I, &u1); fe_mul(&d2, &I, &u2);
    fe_mul(&zinv, &d1, &d2); fe_mul(&zinv, &zinv, &p->T);
    fe_mul(&t, &p->T, &zinv);
    if (fe_isneg(&t)) { fe ix, iy; fe_mul(&ix, &p->X, &C_SQRT_M1); fe_mul(&iy, &p->Y, &C_SQRT_M1); X = iy; Y = ix; fe_mul(&den, &d1, &C_INVSQRT_A_MINUS_D); }
    else den = d2;
    fe_mul(&t, &X, &zinv);
    if (fe_isneg(&t)) fe_neg(&Y, &Y);
    fe_sub(&t, &p->Z, &Y); fe_mul(&s, &den, &t); fe_abs(&s, &s);
    fe_tobytes(out, &s);
}
static void elligator_map(ge *p, const fe *t0) {
    fe r, u, v, t, s, s_prime, c, N, w0, w1, w2, w3, ss;
    fe_sq(&r, t0); fe_mul(&r, &r, &C_SQRT_M1);
    fe_add(&u, &r, &FE_ONE); fe_mul(&u, &u, &C_ONE_MINUS_D_SQ);
    fe_mul(&t, &r, &C_D); fe_neg(&t, &t); fe_sub(&t, &t, &FE_ONE); /* -1 - r d */
    fe_add(&v, &r, &C_D); fe_mul(&v, &v, &t);
    int sq = fe_sqrt_ratio_m1(&s, &u, &v);
    fe_mul(&s_prime, &s, t0); fe_abs(&s_prime, &s_prime); fe_neg(&s_prime, &s_prime);
    if (!sq) { s = s_prime; c = r; } else fe_neg(&c, &FE_ONE);
    fe_sub(&t, &r, &FE_ONE); fe_mul(&N, &c, &t); fe_mul(&N, &N, &C_D_MINUS_ONE_SQ); fe_sub(&N, &N, &v);
    fe_mul(&w0, &s, &v); fe_add(&w0, &w0, &w0);
    fe_mul(&w1, &N, &C_SQRT_AD_MINUS_ONE);
    fe_sq(&ss, &s); fe_sub(&w2, &FE_ONE, &ss); fe_add(&w3, &FE_ONE, &ss);
    fe_mul(&p->X, &w0, &w3); fe_mul(&p->Y, &w2, &w1); fe_mul(&p->Z, &w1, &w3); fe_mul(&p->T, &w0, &w2);
}
static void from_uniform_bytes(ge *p, const u8 b[64]) {
    fe r1, r2; ge p1, p2;
    fe_frombytes(&r1, b); fe_frombytes(&r2, b + 32);
    elligator_map(&p1, &r1); elligator_map(&p2, &r2);
    ge_add(p, &p1, &p2);
}

/* ============================================================ Keccak-f[1600], SHAKE256, SHA3-512 */
static const u64 KRC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL,
    0x000000000000808BULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
    0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KROT[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
static const int KPIL[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
#define ROL64(x, n) (((x) << (n)) | ((x) >> (64 - (n))))
static void keccak_f(u64 st[25]) {
    u64 bc[5], t;
    for (int r = 0; r < 24; r++) {
        for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
        for (int i = 0; i < 5; i++) { t = bc[(i + 4) % 5] ^ ROL64(bc[(i + 1) % 5], 1); for (int j = 0; j < 25; j += 5) st[j + i] ^= t; }
        t = st[1];
        for (int i = 0; i < 24; i++) { int j = KPIL[i]; u64 b = st[j]; st[j] = ROL64(t, KROT[i]); t = b; }
        for (int j = 0; j < 25; j += 5) { for (int i = 0; i < 5; i++) bc[i] = st[j + i]; for (int i = 0; i < 5; i++) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5]; }
        st[0] ^= KRC[r];
    }
}
/* byte view of the state assumes a little-endian host (x86-64 / aarch64) */
typedef struct { u64 st[25]; size_t pos, rate; } sponge;
static void sponge_init(sponge *s, size_t rate) { memset(s, 0, sizeof *s); s->rate = rate; }
static void sponge_absorb(sponge *s, const u8 *d, size_t n) { u8 *b = (u8 *)s->st; for (size_t i = 0; i < n; i++) { b[s->pos++] ^= d[i]; if (s->pos == s->rate) { keccak_f(s->st); s->pos = 0; } } }
static void sponge_finish(sponge *s, u8 pad) { u8 *b = (u8 *)s->st; b[s->pos] ^= pad; b[s->rate - 1] ^= 0x80; keccak_f(s->st); s->pos = 0; }
static void sponge_squeeze(sponge *s, u8 *out, size_t n) { u8 *b = (u8 *)s->st; for (size_t i = 0; i < n; i++) { if (s->pos == s->rate) { keccak_f(s->st); s->pos = 0; } out[i] = b[s->pos++]; } }
static void sha3_512(u8 out[64], const u8 *d, size_t n) { sponge s; sponge_init(&s, 72); sponge_absorb(&s, d, n); sponge_finish(&s, 0x06); sponge_squeeze(&s, out, 64); }

/* ============================================================ STROBE-128 / Merlin */
#define STROBE_R 166
#define FL_I 1
#define FL_A 2
#define FL_C 4
#define FL_T 8
#define FL_M 16
#define FL_K 32
struct bpo_transcript { u64 st[25]; u8 pos, pos_begin, cur_flags; };
typedef struct bpo_transcript strobe;
static void strobe_run_f(strobe *s) { u8 *b = (u8 *)s->st; b[s->pos] ^= s->pos_begin; b[s->pos + 1] ^= 0x04; b[STROBE_R + 1] ^= 0x80; keccak_f(s->st); s->pos = 0; s->pos_begin = 0; }
static void strobe_absorb(strobe *s, const u8 *d, size_t n) { u8 *b = (u8 *)s->st; for (size_t i = 0; i < n; i++) { b[s->pos++] ^= d[i]; if (s->pos == STROBE_R) strobe_run_f(s); } }
static void strobe_overwrite(strobe *s, const u8 *d, size_t n) { u8 *b = (u8 *)s->st; for (size_t i = 0; i < n; i++) { b[s->pos++] = d[i]; if (s->pos == STROBE_R) strobe_run_f(s); } }
static void strobe_squeeze(strobe *s, u8 *d, size_t n) { u8 *b = (u8 *)s->st; for (size_t i = 0; i < n; i++) { d[i] = b[s->pos]; b[s->pos++] = 0; if (s->pos == STROBE_R) strobe_run_f(s); } }
static void strobe_begin_op(strobe *s, u8 flags, int more) {
    if (more) return;
    u8 old = s->pos_begin; s->pos_begin = s->pos + 1; s->cur_flags = flags;
    u8 d[2] = {old, flags}; strobe_absorb(s, d, 2);
    if ((flags & (FL_C | FL_K)) && s->pos != 0) strobe_run_f(s);
}
static void strobe_meta_ad(strobe *s, const u8 *d, size_t n, int more) { strobe_begin_op(s, FL_M | FL_A, more); strobe_absorb(s, d, n); }
static void strobe_ad(strobe *s, const u8 *d, size_t n, int more) { strobe_begin_op(s, FL_A, more); strobe_absorb(s, d, n); }
static void strobe_prf(strobe *s, u8 *d, size_t n, int more) { strobe_begin_op(s, FL_I | FL_A | FL_C, more); strobe_squeeze(s, d, n); }
static void strobe_key(strobe *s, const u8 *d, size_t n, int more) { strobe_begin_op(s, FL_A | FL_C, more); strobe_overwrite(s, d, n); }
static void strobe_init(strobe *s, const u8 *label, size_t n) {
    memset(s, 0, sizeof *s);
    u8 *b = (u8 *)s->st;
    const u8 hdr[6] = {1, STROBE_R + 2, 1, 0, 1, 96};
    memcpy(b, hdr, 6); memcpy(b + 6, "STROBEv1.0.2", 12);
    keccak_f(s->st);
    strobe_meta_ad(s, label, n, 0);
}
static void u32le(u8 o[4], size_t n) { o[0] = (u8)n; o[1] = (u8)(n >> 8); o[2] = (u8)(n >> 16); o[3] = (u8)(n >> 24); }
static void tr_append(strobe *s, const char *label, const u8 *msg, size_t n) { u8 l4[4]; u32le(l4, n); strobe_meta_ad(s, (const u8 *)label, strlen(label), 0); strobe_meta_ad(s, l4, 4, 1); strobe_ad(s, msg, n, 0); }
static void tr_append_u64(strobe *s, const char *label, u64 x) { u8 b[8]; for (int i = 0; i < 8; i++) b[i] = (u8)(x >> (8 * i)); tr_append(s, label, b, 8); }
static void tr_challenge(strobe *s, const char *label, u8 *out, size_t n) { u8 l4[4]; u32le(l4, n); strobe_meta_ad(s, (const u8 *)label, strlen(label), 0); strobe_meta_ad(s, l4, 4, 1); strobe_prf(s, out, n, 0); }
static void tr_init(strobe *s, const u8 *label, size_t n) { strobe_init(s, (const u8 *)"Merlin v1.0", 11); tr_append(s, "dom-sep", label, n); }
static void tr_append_scalar(strobe *s, const char *label, const sc *x) { u8 b[32]; sc_store(b, x); tr_append(s, label, b, 32); }
static void tr_challenge_scalar(strobe *s, const char *label, sc *out) { u8 b[64]; tr_challenge(s, label, b, 64); sc_wide(out, b); }
static int tr_validate_and_append_point(strobe *s, const char *label, const u8 p[32]) { u8 z = 0; for (int i = 0; i < 32; i++) z |= p[i]; if (!z) return 0; tr_append(s, label, p, 32); return 1; }
/* TranscriptRng */
static void rng_begin(strobe *rng, const strobe *t) { *rng = *t; }
static void rng_rekey(strobe *rng, const char *label, const u8 *w, size_t n) { u8 l4[4]; u32le(l4, n); strobe_meta_ad(rng, (const u8 *)label, strlen(label), 0); strobe_meta_ad(rng, l4, 4, 1); strobe_key(rng, w, n, 0); }
static void rng_finalize(strobe *rng, const u8 ext32[32]) { strobe_meta_ad(rng, (const u8 *)"rng", 3, 0); strobe_key(rng, ext32, 32, 0); }
static void rng_fill(strobe *rng, u8 *out, size_t n) { u8 l4[4]; u32le(l4, n); strobe_meta_ad(rng, l4, 4, 0); strobe_prf(rng, out, n, 0); }
static void rng_scalar(strobe *rng, sc *out) { u8 b[64]; rng_fill(rng, b, 64); sc_wide(out, b); }

bpo_transcript *bpo_transcript_new(const u8 *label, size_t len) { strobe *s = (strobe *)malloc(sizeof *s); tr_init(s, label, len); return s; }
void bpo_transcript_free(bpo_transcript *t) { free(t); }
void bpo_transcript_append(bpo_transcript *t, const u8 *label, size_t ll, const u8 *msg, size_t ml) { u8 l4[4]; u32le(l4, ml); strobe_meta_ad(t, label, ll, 0); strobe_meta_ad(t, l4, 4, 1); strobe_ad(t, msg, ml, 0); }
void bpo_transcript_challenge(bpo_transcript *t, const u8 *label, size_t ll, u8 *out, size_t n) { u8 l4[4]; u32le(l4, n); strobe_meta_ad(t, label, ll, 0); strobe_meta_ad(t, l4, 4, 1); strobe_prf(t, out, n, 0); }

/* ============================================================ constants, generators */
static ge G_B, G_BBLIND;
static int g_init_done = 0;
static void bpo_init(void) {
    if (g_init_done) return;
    /* d = -121665/121666 */
    fe a, b;
    fe_from_le_hex(&C_D, "52036cee2b6ffe738cc740797779e89800700a4d4141d8ab75eb4dca135978a3");
    fe_add(&C_D2, &C_D, &C_D);
    fe_from_le_hex(&C_SQRT_M1, "2b8324804fc1df0b2b4d00993dfbd7a72f431806ad2fe478c4ee1b274a0ea0b0");
    /* derived constants, computed rather than transcribed */
    fe_neg(&a, &FE_ONE); fe_sub(&a, &a, &C_D);            /* a - d = -1 - d */
    fe_sqrt_ratio_m1(&C_INVSQRT_A_MINUS_D, &FE_ONE, &a);
    fe_sqrt_ratio_m1(&b, &a, &FE_ONE);                     /* sqrt(ad - 1) = sqrt(-d - 1); take the ODD root */
    if (!fe_isneg(&b)) fe_neg(&b, &b);
    C_SQRT_AD_MINUS_ONE = b;
    fe_sq(&a, &C_D); fe_sub(&C_ONE_MINUS_D_SQ, &FE_ONE, &a);
    fe_sub(&a, &C_D, &FE_ONE); fe_sq(&C_D_MINUS_ONE_SQ, &a);
    static const u8 BP[32] = {0xe2, 0xf2, 0xae, 0x0a, 0x6a, 0xbc, 0x4e, 0x71, 0xa8, 0x84, 0xa9, 0x61, 0xc5, 0x00, 0x51, 0x5f,
                              0x58, 0xe3, 0x0b, 0x6a, 0xa5, 0x82, 0xdd, 0x8d, 0xb6, 0xa6, 0x59, 0x45, 0xe0, 0x8d, 0x2d, 0x76};
    ristretto_decode(&G_B, BP);
    u8 h[64]; sha3_512(h, BP, 32);
    from_uniform_bytes(&G_BBLIND, h);
    g_init_done = 1;
}
void bpo_pedersen_gens(u8 B32[32], u8 Bb32[32]) { bpo_init(); ristretto_encode(B32, &G_B); ristretto_encode(Bb32, &G_BBLIND); }
void bpo_from_uniform_bytes(const u8 in[64], u8 out[32]) { bpo_init(); ge p; from_uniform_bytes(&p, in); ristretto_encode(out, &p); }
int bpo_point_decode_ok(const u8 p32[32]) { bpo_init(); ge p; return ristretto_decode(&p, p32); }
int bpo_point_add(const u8 a[32], const u8 b[32], u8 out[32]) { bpo_init(); ge p, q, r; if (!ristretto_decode(&p, a) || !ristretto_decode(&q, b)) return -1; ge_add(&r, &p, &q); ristretto_encode(out, &r); return 0; }
int bpo_point_mul(const u8 s[32], const u8 p32[32], u8 out[32]) { bpo_init(); ge p, r; sc k; if (!ristretto_decode(&p, p32)) return -1; sc_load(&k, s); ge_scalarmul(&r, &k, &p); ristretto_encode(out, &r); return 0; }

/* BulletproofGens chains, cached as extended points (party 0).  SHAKE256("GeneratorsChain"||c||u32le(0)) */
static ge *g_G = NULL, *g_H = NULL; static size_t g_gens_n = 0;
static void gens_chain(ge *out, char c, size_t n) {
    sponge s; sponge_init(&s, 136);
    u8 lab[20]; memcpy(lab, "GeneratorsChain", 15); lab[15] = (u8)c; lab[16] = lab[17] = lab[18] = lab[19] = 0;
    sponge_absorb(&s, lab, 20); sponge_finish(&s, 0x1F);
    u8 *stream = (u8 *)malloc(64 * n);
    sponge_squeeze(&s, stream, 64 * n);
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (long i = 0; i < (long)n; i++) from_uniform_bytes(&out[i], stream + 64 * i);
    free(stream);
}
static void gens_ensure(size_t n) {
    bpo_init();
    if (n <= g_gens_n) return;
    size_t cap = 64; while (cap < n) cap *= 2;
    free(g_G); free(g_H);
    g_G = (ge *)malloc(cap * sizeof(ge)); g_H = (ge *)malloc(cap * sizeof(ge));
    gens_chain(g_G, 'G', cap); gens_chain(g_H, 'H', cap);
    g_gens_n = cap;
}
void bpo_gens(size_t i0, size_t n, u8 *G32, u8 *H32) {
    gens_ensure(i0 + n);
    for (size_t i = 0; i < n; i++) { if (G32) ristretto_encode(G32 + 32 * i, &g_G[i0 + i]); if (H32) ristretto_encode(H32 + 32 * i, &g_H[i0 + i]); }
}

/* ============================================================ MSM algorithms */
/* signed radix-2^w digits of a reduced scalar (dalek Scalar::to_radix_2w); returns digit count */
static int sc_radix_2w(int16_t *digits, const sc *k, int w) {
    int nd = (256 + w - 1) / w; /* dalek: digits_count = (256 + w - 1)/w */
    u64 carry = 0; u64 radix = 1ULL << w, mask = radix - 1;
    for (int i = 0; i < nd; i++) {
        int bit = i * w, limb = bit >> 6, off = bit & 63;
        u64 b;
        if (limb >= 4) b = 0;
        else if (off + w <= 64 || limb == 3) b = (k->v[limb] >> off) & mask;
        else b = ((k->v[limb] >> off) | (k->v[limb + 1] << (64 - off))) & mask;
        u64 coef = carry + b;
        carry = (coef + radix / 2) >> w;
        digits[i] = (int16_t)((int64_t)coef - (int64_t)(carry << w));
    }
    /* for a reduced scalar (< 2^253) and w in {4..8,16} the final carry is zero, except w=8 where dalek
       folds it into the last digit; we never hit it because 253 < 256 - 1 */
    return nd;
}
static void ge_table8(ge_c t[8], const ge *p) { /* [P,2P,..,8P] cached */
    ge cur = *p; ge_to_cached(&t[0], &cur);
    for (int i = 1; i < 8; i++) { ge_add_cached(&cur, &cur, &t[0], 1); ge_to_cached(&t[i], &cur); }
}
/* dalek `multiscalar_mul` (Straus, radix 16, no zero skipping): range [0,n) of one thread */
static void msm_straus_range(ge *out, const sc *scalars, const ge *points, size_t n) {
    ge_c(*tab)[8] = (ge_c(*)[8])malloc(n * sizeof(ge_c[8]));
    int8_t(*dig)[64] = (int8_t(*)[64])malloc(n * 64);
    for (size_t i = 0; i < n; i++) {
        ge_table8(tab[i], &points[i]);
        sc k; sc_reduce(&k, &scalars[i]);
        int16_t d[64]; sc_radix_2w(d, &k, 4);
        for (int j = 0; j < 64; j++) dig[i][j] = (int8_t)d[j];
    }
    ge q; ge_identity(&q);
    for (int j = 63; j >= 0; j--) {
        ge_dbl(&q, &q); ge_dbl(&q, &q); ge_dbl(&q, &q); ge_dbl(&q, &q);
        for (size_t i = 0; i < n; i++) {
            int d = dig[i][j];
            if (d > 0) ge_add_cached(&q, &q, &tab[i][d - 1], 1);
            else if (d < 0) ge_add_cached(&q, &q, &tab[i][-d - 1], -1);
            else { ge_c idc; ge id; ge_identity(&id); ge_to_cached(&idc, &id); ge_add_cached(&q, &q, &idc, 1); } /* const-time: add identity */
        }
    }
    free(tab); free(dig);
    *out = q;
}
/* dalek vartime Pippenger: w = 6 (n<500), 7 (n<800), 8 otherwise; signed digits, 2^(w-1) buckets */
static void msm_pippenger_range(ge *out, const sc *scalars, const ge *points, size_t n) {
    int w = n < 500 ? 6 : (n < 800 ? 7 : 8);
    int nd = (256 + w - 1) / w, nb = 1 << (w - 1);
    int16_t *dig = (int16_t *)malloc(n * nd * sizeof(int16_t));
    ge_c *pc = (ge_c *)malloc(n * sizeof(ge_c));
    ge *buckets = (ge *)malloc(nb * sizeof(ge));
    for (size_t i = 0; i < n; i++) { sc k; sc_reduce(&k, &scalars[i]); sc_radix_2w(dig + i * nd, &k, w); ge_to_cached(&pc[i], &points[i]); }
    ge total; ge_identity(&total);
    for (int col = nd - 1; col >= 0; col--) {
        for (int b = 0; b < nb; b++) ge_identity(&buckets[b]);
        for (size_t i = 0; i < n; i++) {
            int d = dig[i * nd + col];
            if (d > 0) ge_add_cached(&buckets[d - 1], &buckets[d - 1], &pc[i], 1);
            else if (d < 0) ge_add_cached(&buckets[-d - 1], &buckets[-d - 1], &pc[i], -1);
        }
        ge run = buckets[nb - 1], sum = buckets[nb - 1];
        for (int b = nb - 2; b >= 0; b--) { ge_add(&run, &run, &buckets[b]); ge_add(&sum, &sum, &run); }
        for (int k = 0; k < w; k++) ge_dbl(&total, &total);
        ge_add(&total, &total, &sum);
    }
    free(dig); free(pc); free(buckets);
    *out = total;
}
/* width-5 NAF (dalek Scalar::non_adjacent_form) */
static void sc_naf5(int8_t naf[256], const sc *k) {
    memset(naf, 0, 256);
    u64 x[5] = {k->v[0], k->v[1], k->v[2], k->v[3], 0};
    int pos = 0; u64 carry = 0;
    while (pos < 256) {
        int idx = pos >> 6, off = pos & 63;
        u64 bits = off < 59 ? (x[idx] >> off) : ((x[idx] >> off) | (x[idx + 1] << (64 - off)));
        u64 window = carry + (bits & 31);
        if ((window & 1) == 0) { pos += 1; continue; }
        if (window < 16) { carry = 0; naf[pos] = (int8_t)window; }
        else { carry = 1; naf[pos] = (int8_t)((int)window - 32); }
        pos += 5;
    }
}
static void ge_table_odd8(ge_c t[8], const ge *p) { /* [P,3P,5P,..,15P] */
    ge p2, cur = *p; ge_dbl(&p2, p); ge_c p2c; ge_to_cached(&p2c, &p2);
    ge_to_cached(&t[0], &cur);
    for (int i = 1; i < 8; i++) { ge_add_cached(&cur, &cur, &p2c, 1); ge_to_cached(&t[i], &cur); }
}
/* dalek vartime Straus (n < 190): width-5 NAF, interleaved */
static void msm_straus_naf_range(ge *out, const sc *scalars, const ge *points, size_t n) {
    ge_c(*tab)[8] = (ge_c(*)[8])malloc((n ? n : 1) * sizeof(ge_c[8]));
    int8_t(*naf)[256] = (int8_t(*)[256])malloc((n ? n : 1) * 256);
    for (size_t i = 0; i < n; i++) { sc k; sc_reduce(&k, &scalars[i]); sc_naf5(naf[i], &k); ge_table_odd8(tab[i], &points[i]); }
    ge q; ge_identity(&q);
    for (int j = 255; j >= 0; j--) {
        ge_dbl(&q, &q);
        for (size_t i = 0; i < n; i++) {
            int d = naf[i][j];
            if (d > 0) ge_add_cached(&q, &q, &tab[i][d / 2], 1);
            else if (d < 0) ge_add_cached(&q, &q, &tab[i][(-d) / 2], -1);
        }
    }
    free(tab); free(naf);
    *out = q;
}
static void msm_naive_range(ge *out, const sc *scalars, const ge *points, size_t n) {
    ge acc; ge_identity(&acc);
    for (size_t i = 0; i < n; i++) { ge t; ge_scalarmul(&t, &scalars[i], &points[i]); ge_add(&acc, &acc, &t); }
    *out = acc;
}
/* top level: split the point range over OpenMP threads, sum the partials */
static void msm_run(ge *out, const sc *scalars, const ge *points, size_t n, int algo) {
    int T = g_threads; if ((size_t)T > n / 64 + 1) T = (int)(n / 64 + 1);
    ge *part = (ge *)malloc(T * sizeof(ge));
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int t = 0; t < T; t++) {
        size_t lo = n * t / T, hi = n * (t + 1) / T, m = hi - lo;
        if (algo == BPO_MSM_NAIVE) msm_naive_range(&part[t], scalars + lo, points + lo, m);
        else if (algo == BPO_MSM_STRAUS_CT) msm_straus_range(&part[t], scalars + lo, points + lo, m);
        else if (m < 190) msm_straus_naf_range(&part[t], scalars + lo, points + lo, m);
        else msm_pippenger_range(&part[t], scalars + lo, points + lo, m);
    }
    ge acc = part[0];
    for (int t = 1; t < T; t++) ge_add(&acc, &acc, &part[t]);
    free(part);
    *out = acc;
}
int bpo_msm(const u8 *scalars, const u8 *points32, size_t n, u8 out32[32], int algo) {
    bpo_init();
    sc *s = (sc *)malloc((n ? n : 1) * sizeof(sc)); ge *p = (ge *)malloc((n ? n : 1) * sizeof(ge));
    int ok = 1;
    for (size_t i = 0; i < n; i++) { sc_load(&s[i], scalars + 32 * i); if (!ristretto_decode(&p[i], points32 + 32 * i)) ok = 0; }
    if (ok) { ge r; msm_run(&r, s, p, n, algo); ristretto_encode(out32, &r); }
    free(s); free(p);
    return ok ? 0 : -1;
}
int bpo_msm_gens(const u8 *sG, const u8 *sH, size_t n, size_t offset, const u8 *es, const u8 *ep, size_t k, u8 out32[32], int algo) {
    gens_ensure(offset + n);
    size_t tot = (sG ? n : 0) + (sH ? n : 0) + k;
    sc *s = (sc *)malloc((tot ? tot : 1) * sizeof(sc)); ge *p = (ge *)malloc((tot ? tot : 1) * sizeof(ge));
    size_t c = 0; int ok = 1;
    if (sG) for (size_t i = 0; i < n; i++) { sc_load(&s[c], sG + 32 * i); p[c++] = g_G[offset + i]; }
    if (sH) for (size_t i = 0; i < n; i++) { sc_load(&s[c], sH + 32 * i); p[c++] = g_H[offset + i]; }
    for (size_t i = 0; i < k; i++) { sc_load(&s[c], es + 32 * i); if (!ristretto_decode(&p[c++], ep + 32 * i)) ok = 0; }
    if (ok) { ge r; msm_run(&r, s, p, tot, algo); ristretto_encode(out32, &r); }
    free(s); free(p);
    return ok ? 0 : -1;
}
void bpo_pedersen_commit(const u8 *v, const u8 *r, size_t n, u8 *out32) {
    bpo_init();
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (long i = 0; i < (long)n; i++) {
        sc s[2]; ge p[2] = {G_B, G_BBLIND}, q;
        sc_load(&s[0], v + 32 * i); sc_load(&s[1], r + 32 * i);
        msm_straus_range(&q, s, p, 2); /* dalek: const-time multiscalar_mul of 2 points */
        ristretto_encode(out32 + 32 * i, &q);
    }
}
/* one IPP generator fold (dalek: vartime_multiscalar_mul of 2 points per element = NAF Straus) */
static void fold_points(ge *out, const sc *sl, const sc *sr, const ge *PL, const ge *PR, size_t n, int per_elem_scalars) {
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (long i = 0; i < (long)n; i++) {
        sc s[2] = {per_elem_scalars ? sl[i] : sl[0], per_elem_scalars ? sr[i] : sr[0]};
        ge p[2] = {PL[i], PR[i]}, q;
        msm_straus_naf_range(&q, s, p, 2);
        out[i] = q;
    }
}
int bpo_fold_points(const u8 sl[32], const u8 sr[32], const u8 *PL32, const u8 *PR32, size_t n, u8 *out32) {
    bpo_init();
    ge *L = (ge *)malloc(n * sizeof(ge)), *R = (ge *)malloc(n * sizeof(ge)), *O = (ge *)malloc(n * sizeof(ge));
    int ok = 1; sc a, b; sc_load(&a, sl); sc_load(&b, sr);
    for (size_t i = 0; i < n; i++) if (!ristretto_decode(&L[i], PL32 + 32 * i) || !ristretto_decode(&R[i], PR32 + 32 * i)) ok = 0;
    if (ok) { fold_points(O, &a, &b, L, R, n, 0); for (size_t i = 0; i < n; i++) ristretto_encode(out32 + 32 * i, &O[i]); }
    free(L); free(R); free(O);
    return ok ? 0 : -1;
}

/* ============================================================ MiMC (src/mimc_hash/mimc.rs) */
static sc g_mimc_c[486]; static int g_mimc_set = 0;
void bpo_set_mimc_constants(const u8 *c) { for (int i = 0; i < 486; i++) { sc t; sc_load(&t, c + 32 * i); t.v[3] &= 0x7FFFFFFFFFFFFFFFULL; sc_reduce(&g_mimc_c[i], &t); } g_mimc_set = 1; }
/* mimc.rs:7-23 with k = 0; optional witness trace in gadget order (mimc_hash_gadget.rs:133-144) */
static void mimc_encrypt(sc *state, u8 *trace) {
    for (int i = 0; i < 486; i++) {
        sc t, t2, t3;
        sc_add(&t, state, &g_mimc_c[i]); sc_mul(&t2, &t, &t); sc_mul(&t3, &t2, &t);
        if (trace) { /* multiplier 2i: (t,t,t^2); multiplier 2i+1: (t^2,t,t^3) */
            u8 *o = trace + (size_t)i * 192;
            sc_store(o, &t); sc_store(o + 32, &t); sc_store(o + 64, &t2);
            sc_store(o + 96, &t2); sc_store(o + 128, &t); sc_store(o + 160, &t3);
        }
        *state = t3;
    }
}
void bpo_mimc_sponge(const u8 *blocks, size_t nblocks, u8 out32[32], u8 *trace) {
    sc st = SC_ZERO;
    for (size_t b = 0; b < nblocks; b++) { sc x; sc_load(&x, blocks + 32 * b); x.v[3] &= 0x7FFFFFFFFFFFFFFFULL; sc_add(&st, &st, &x); mimc_encrypt(&st, trace ? trace + b * 486 * 192 : NULL); }
    sc_store(out32, &st);
}
/* mimc.rs:61-97: be_to_scalars, pad, sponge */
void bpo_mimc_hash(const u8 *data, size_t len, u8 out32[32]) {
    size_t nb = (len + 31) / 32; if (nb == 0) nb = 0;
    u8 *le = (u8 *)calloc(32 * (nb + 1) + 32, 1);
    for (size_t i = 0; i < len; i++) le[i] = data[len - 1 - i];
    if (nb == 0) { /* be_to_scalars of empty input yields no blocks; reference would panic on last().unwrap() */ nb = 1; }
    u8 *last = le + 32 * (nb - 1);
    last[31] &= 0x7F; /* from_bits */
    int l = 32; while (l > 0 && last[l - 1] == 0) l--;
    if (l < 32) { u8 pad = (u8)(32 - l); for (int i = l; i < 32; i++) last[i] = pad; }
    else { memset(le + 32 * nb, 32, 32); nb++; }
    bpo_mimc_sponge(le, nb, out32, NULL);
    free(le);
}

/* ============================================================ R1CS prove / verify */
static size_t next_pow2(size_t n) { size_t p = 1; while (p < n) p <<= 1; return p; }
static void sc_inner(sc *out, const sc *a, const sc *b, size_t n) {
    int T = g_threads; sc part[64]; if (T > 64) T = 64;
#pragma omp parallel for num_threads(T) schedule(static, 1)
    for (int t = 0; t < T; t++) { sc acc = SC_ZERO; for (size_t i = n * t / T; i < n * (t + 1) / T; i++) sc_muladd(&acc, &a[i], &b[i], &acc); part[t] = acc; }
    sc acc = SC_ZERO; for (int t = 0; t < T; t++) sc_add(&acc, &acc, &part[t]);
    *out = acc;
}
/* flattened constraints; coefficients are arbitrary 32-byte scalars */
static void flatten(const sc *z, size_t n, size_t m, size_t q, const uint32_t *row_ptr, const uint32_t *term_var, const u8 *term_coeff,
                    sc *wL, sc *wR, sc *wO, sc *wV, sc *wc) {
    for (size_t i = 0; i < n; i++) wL[i] = wR[i] = wO[i] = SC_ZERO;
    for (size_t i = 0; i < m; i++) wV[i] = SC_ZERO;
    *wc = SC_ZERO;
    sc e = *z;
    for (size_t r = 0; r < q; r++) {
        for (uint32_t k = row_ptr[r]; k < row_ptr[r + 1]; k++) {
            sc c, ec; sc_load(&c, term_coeff + 32 * (size_t)k); sc_mul(&ec, &e, &c);
            uint32_t kind = term_var[k] >> 29, idx = term_var[k] & 0x1FFFFFFF;
            if (kind == BPO_VAR_L) sc_add(&wL[idx], &wL[idx], &ec);
            else if (kind == BPO_VAR_R) sc_add(&wR[idx], &wR[idx], &ec);
            else if (kind == BPO_VAR_O) sc_add(&wO[idx], &wO[idx], &ec);
            else if (kind == BPO_VAR_V) sc_sub(&wV[idx], &wV[idx], &ec);
            else sc_sub(wc, wc, &ec);
        }
        sc_mul(&e, &e, z);
    }
}
typedef struct { u8 L[32][32], R[32][32]; int lg; sc a, b; } ipp_proof;
/* InnerProductProof::create; G,H are consumed (folded in place) */
static void ipp_create(strobe *t, ipp_proof *pr, const ge *Q, const sc *Gf, const sc *Hf, ge *G, ge *H, sc *a, sc *b, size_t n) {
    tr_append(t, "dom-sep", (const u8 *)"ipp v1", 6);
    tr_append_u64(t, "n", n);
    pr->lg = 0;
    int first = 1;
    sc *ts = (sc *)malloc((2 * n + 1) * sizeof(sc)); ge *tp = (ge *)malloc((2 * n + 1) * sizeof(ge));
    sc *fs1 = (sc *)malloc((n / 2 + 1) * sizeof(sc)), *fs2 = (sc *)malloc((n / 2 + 1) * sizeof(sc));
    while (n != 1) {
        n /= 2;
        sc cL, cR; sc_inner(&cL, a, b + n, n); sc_inner(&cR, a + n, b, n);
        ge Lp, Rp;
        for (int side = 0; side < 2; side++) {
            /* L: a_L o G_R, b_R o H_L, c_L Q ;  R: a_R o G_L, b_L o H_R, c_R Q */
            const sc *av = side == 0 ? a : a + n, *bv = side == 0 ? b + n : b;
            const ge *gp = side == 0 ? G + n : G, *hp = side == 0 ? H : H + n;
            const sc *gf = side == 0 ? Gf + n : Gf, *hf = side == 0 ? Hf : Hf + n;
            for (size_t i = 0; i < n; i++) {
                if (first) { sc_mul(&ts[i], &av[i], &gf[i]); sc_mul(&ts[n + i], &bv[i], &hf[i]); }
                else { ts[i] = av[i]; ts[n + i] = bv[i]; }
                tp[i] = gp[i]; tp[n + i] = hp[i];
            }
            ts[2 * n] = side == 0 ? cL : cR; tp[2 * n] = *Q;
            msm_run(side == 0 ? &Lp : &Rp, ts, tp, 2 * n + 1, BPO_MSM_VARTIME);
        }
        ristretto_encode(pr->L[pr->lg], &Lp); ristretto_encode(pr->R[pr->lg], &Rp);
        tr_append(t, "L", pr->L[pr->lg], 32); tr_append(t, "R", pr->R[pr->lg], 32);
        pr->lg++;
        sc u, ui; tr_challenge_scalar(t, "u", &u); sc_invert(&ui, &u);
        for (size_t i = 0; i < n; i++) {
            sc x, y;
            sc_mul(&x, &a[i], &u); sc_mul(&y, &ui, &a[n + i]); sc_add(&a[i], &x, &y);
            sc_mul(&x, &b[i], &ui); sc_mul(&y, &u, &b[n + i]); sc_add(&b[i], &x, &y);
        }
        if (first) {
            for (size_t i = 0; i < n; i++) { sc_mul(&fs1[i], &ui, &Gf[i]); sc_mul(&fs2[i], &u, &Gf[n + i]); }
            fold_points(G, fs1, fs2, G, G + n, n, 1);
            for (size_t i = 0; i < n; i++) { sc_mul(&fs1[i], &u, &Hf[i]); sc_mul(&fs2[i], &ui, &Hf[n + i]); }
            fold_points(H, fs1, fs2, H, H + n, n, 1);
        } else {
            fold_points(G, &ui, &u, G, G + n, n, 0);
            fold_points(H, &u, &ui, H, H + n, n, 0);
        }
        first = 0;
    }
    pr->a = a[0]; pr->b = b[0];
    free(ts); free(tp); free(fs1); free(fs2);
}
static void ge_commit(ge *out, const sc *v, const sc *r) { sc s[2] = {*v, *r}; ge p[2] = {G_B, G_BBLIND}; msm_straus_range(out, s, p, 2); }

long bpo_r1cs_prove(const u8 *label, size_t label_len, size_t gens_capacity, size_t n, const u8 *aL8, const u8 *aR8, const u8 *aO8,
                    size_t m, const u8 *v8, const u8 *vb8, size_t q, const uint32_t *row_ptr, const uint32_t *term_var,
                    const u8 *term_coeff, const u8 ext_rng32[32], int flags, u8 *V_out, u8 *proof, size_t proof_cap) {
    bpo_init();
    size_t N = next_pow2(n);
    if (gens_capacity < N) return -2;
    int lgN = 0; while (((size_t)1 << lgN) < N) lgN++;
    size_t need = ((flags & 1) ? 0 : 1) + 32 * (size_t)((flags & 1 ? 14 : 11) + 2 * lgN + 2);
    if (proof_cap < need) return -3;
    gens_ensure(N);
    strobe t; tr_init(&t, label, label_len);
    tr_append(&t, "dom-sep", (const u8 *)"r1cs v1", 7);
    /* commit high-level variables */
    u8 *Venc = (u8 *)malloc(32 * (m ? m : 1));
    bpo_pedersen_commit(v8, vb8, m, Venc);
    for (size_t i = 0; i < m; i++) tr_append(&t, "V", Venc + 32 * i, 32);
    if (V_out) memcpy(V_out, Venc, 32 * m);
    free(Venc);
    tr_append_u64(&t, "m", m);
    strobe rng; rng_begin(&rng, &t);
    for (size_t i = 0; i < m; i++) rng_rekey(&rng, "v_blinding", vb8 + 32 * i, 32);
    rng_finalize(&rng, ext_rng32);
    sc *aL = (sc *)malloc(N * sizeof(sc)), *aR = (sc *)malloc(N * sizeof(sc)), *aO = (sc *)malloc(N * sizeof(sc));
    sc *sL = (sc *)malloc(N * sizeof(sc)), *sR = (sc *)malloc(N * sizeof(sc));
    for (size_t i = 0; i < n; i++) { sc_load(&aL[i], aL8 + 32 * i); sc_load(&aR[i], aR8 + 32 * i); sc_load(&aO[i], aO8 + 32 * i); }
    sc ib, ob, sb;
    rng_scalar(&rng, &ib); rng_scalar(&rng, &ob); rng_scalar(&rng, &sb);
    for (size_t i = 0; i < n; i++) rng_scalar(&rng, &sL[i]);
    for (size_t i = 0; i < n; i++) rng_scalar(&rng, &sR[i]);
    /* A_I1, A_O1, S1: const-time Straus in dalek */
    sc *ts = (sc *)malloc((2 * N + 2) * sizeof(sc)); ge *tp = (ge *)malloc((2 * N + 2) * sizeof(ge));
    ge AI, AO, S; u8 AIe[32], AOe[32], Se[32];
    ts[0] = ib; tp[0] = G_BBLIND; for (size_t i = 0; i < n; i++) { ts[1 + i] = aL[i]; tp[1 + i] = g_G[i]; ts[1 + n + i] = aR[i]; tp[1 + n + i] = g_H[i]; }
    msm_run(&AI, ts, tp, 2 * n + 1, BPO_MSM_STRAUS_CT);
    ts[0] = ob; for (size_t i = 0; i < n; i++) ts[1 + i] = aO[i];
    msm_run(&AO, ts, tp, n + 1, BPO_MSM_STRAUS_CT);
    ts[0] = sb; for (size_t i = 0; i < n; i++) { ts[1 + i] = sL[i]; ts[1 + n + i] = sR[i]; }
    msm_run(&S, ts, tp, 2 * n + 1, BPO_MSM_STRAUS_CT);
    ristretto_encode(AIe, &AI); ristretto_encode(AOe, &AO); ristretto_encode(Se, &S);
    tr_append(&t, "A_I1", AIe, 32); tr_append(&t, "A_O1", AOe, 32); tr_append(&t, "S1", Se, 32);
    tr_append(&t, "dom-sep", (const u8 *)"r1cs-1phase", 11);
    u8 Z32[32]; memset(Z32, 0, 32);
    tr_append(&t, "A_I2", Z32, 32); tr_append(&t, "A_O2", Z32, 32); tr_append(&t, "S2", Z32, 32);
    sc y, z; tr_challenge_scalar(&t, "y", &y); tr_challenge_scalar(&t, "z", &z);
    sc *wL = (sc *)malloc((n + 1) * sizeof(sc)), *wR = (sc *)malloc((n + 1) * sizeof(sc)), *wO = (sc *)malloc((n + 1) * sizeof(sc)), *wV = (sc *)malloc((m + 1) * sizeof(sc)), wc;
    flatten(&z, n, m, q, row_ptr, term_var, term_coeff, wL, wR, wO, wV, &wc);
    sc yinv; sc_invert(&yinv, &y);
    sc *yp = (sc *)malloc(N * sizeof(sc)), *yip = (sc *)malloc(N * sizeof(sc));
    yp[0] = SC_ONE; yip[0] = SC_ONE;
    for (size_t i = 1; i < N; i++) { sc_mul(&yp[i], &yp[i - 1], &y); sc_mul(&yip[i], &yip[i - 1], &yinv); }
    sc *l1 = (sc *)malloc(N * sizeof(sc)), *r0 = (sc *)malloc(N * sizeof(sc)), *r1 = (sc *)malloc(N * sizeof(sc)), *r3 = (sc *)malloc(N * sizeof(sc));
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (long i = 0; i < (long)n; i++) {
        sc_muladd(&l1[i], &yip[i], &wR[i], &aL[i]);
        sc_sub(&r0[i], &wO[i], &yp[i]);
        sc_muladd(&r1[i], &yp[i], &aR[i], &wL[i]);
        sc_mul(&r3[i], &yp[i], &sR[i]);
    }
    const sc *l2 = aO, *l3 = sL;
    sc t1, t2, t3, t4, t5, t6, tmp;
    sc_inner(&t1, l1, r0, n);
    sc_inner(&t2, l1, r1, n); sc_inner(&tmp, l2, r0, n); sc_add(&t2, &t2, &tmp);
    sc_inner(&t3, l2, r1, n); sc_inner(&tmp, l3, r0, n); sc_add(&t3, &t3, &tmp);
    sc_inner(&t4, l1, r3, n); sc_inner(&tmp, l3, r1, n); sc_add(&t4, &t4, &tmp);
    sc_inner(&t5, l2, r3, n);
    sc_inner(&t6, l3, r3, n);
    sc tb1, tb3, tb4, tb5, tb6;
    rng_scalar(&rng, &tb1); rng_scalar(&rng, &tb3); rng_scalar(&rng, &tb4); rng_scalar(&rng, &tb5); rng_scalar(&rng, &tb6);
    const sc *tv[5] = {&t1, &t3, &t4, &t5, &t6}, *tbv[5] = {&tb1, &tb3, &tb4, &tb5, &tb6};
    static const char *TL[5] = {"T_1", "T_3", "T_4", "T_5", "T_6"};
    u8 Te[5][32];
    for (int k = 0; k < 5; k++) { ge T; ge_commit(&T, tv[k], tbv[k]); ristretto_encode(Te[k], &T); tr_append(&t, TL[k], Te[k], 32); }
    sc u, x; tr_challenge_scalar(&t, "u", &u); tr_challenge_scalar(&t, "x", &x);
    sc tb2 = SC_ZERO;
    for (size_t i = 0; i < m; i++) { sc vb; sc_load(&vb, vb8 + 32 * i); sc_muladd(&tb2, &wV[i], &vb, &tb2); }
    sc xp[7]; xp[0] = SC_ONE; for (int k = 1; k <= 6; k++) sc_mul(&xp[k], &xp[k - 1], &x);
    const sc *tc[6] = {&t1, &t2, &t3, &t4, &t5, &t6}, *tbc[6] = {&tb1, &tb2, &tb3, &tb4, &tb5, &tb6};
    sc tx = SC_ZERO, txb = SC_ZERO;
    for (int k = 0; k < 6; k++) { sc_muladd(&tx, tc[k], &xp[k + 1], &tx); sc_muladd(&txb, tbc[k], &xp[k + 1], &txb); }
    sc *lv = (sc *)malloc(N * sizeof(sc)), *rv = (sc *)malloc(N * sizeof(sc));
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (long i = 0; i < (long)N; i++) {
        if ((size_t)i < n) {
            sc a, b;
            sc_mul(&a, &xp[1], &l1[i]); sc_muladd(&a, &xp[2], &l2[i], &a); sc_muladd(&lv[i], &xp[3], &l3[i], &a);
            sc_muladd(&b, &xp[1], &r1[i], &r0[i]); sc_muladd(&rv[i], &xp[3], &r3[i], &b);
        } else { lv[i] = SC_ZERO; sc_neg(&rv[i], &yp[i]); }
    }
    sc eb; sc_muladd(&eb, &x, &sb, &ob); sc_muladd(&eb, &x, &eb, &ib); sc_mul(&eb, &x, &eb);
    tr_append_scalar(&t, "t_x", &tx); tr_append_scalar(&t, "t_x_blinding", &txb); tr_append_scalar(&t, "e_blinding", &eb);
    sc w; tr_challenge_scalar(&t, "w", &w);
    ge Q; ge_scalarmul(&Q, &w, &G_B);
    sc *Gf = (sc *)malloc(N * sizeof(sc)), *Hf = (sc *)malloc(N * sizeof(sc));
    for (size_t i = 0; i < N; i++) { Gf[i] = i < n ? SC_ONE : u; sc_mul(&Hf[i], &yip[i], &Gf[i]); }
    ge *Gc = (ge *)malloc(N * sizeof(ge)), *Hc = (ge *)malloc(N * sizeof(ge));
    memcpy(Gc, g_G, N * sizeof(ge)); memcpy(Hc, g_H, N * sizeof(ge));
    ipp_proof ipp;
    ipp_create(&t, &ipp, &Q, Gf, Hf, Gc, Hc, lv, rv, N);
    /* serialise */
    u8 *o = proof;
    if (flags & 1) { memcpy(o, AIe, 32); memcpy(o + 32, AOe, 32); memcpy(o + 64, Se, 32); memset(o + 96, 0, 96); o += 192; }
    else { *o++ = 0; memcpy(o, AIe, 32); memcpy(o + 32, AOe, 32); memcpy(o + 64, Se, 32); o += 96; }
    for (int k = 0; k < 5; k++) { memcpy(o, Te[k], 32); o += 32; }
    sc_store(o, &tx); sc_store(o + 32, &txb); sc_store(o + 64, &eb); o += 96;
    for (int j = 0; j < ipp.lg; j++) { memcpy(o, ipp.L[j], 32); memcpy(o + 32, ipp.R[j], 32); o += 64; }
    sc_store(o, &ipp.a); sc_store(o + 32, &ipp.b); o += 64;
    free(aL); free(aR); free(aO); free(sL); free(sR); free(ts); free(tp); free(wL); free(wR); free(wO); free(wV);
    free(yp); free(yip); free(l1); free(r0); free(r1); free(r3); free(lv); free(rv); free(Gf); free(Hf); free(Gc); free(Hc);
    return (long)(o - proof);
}

static int sc_load_canonical(sc *s, const u8 b[32]) { sc_load(s, b); return !sc_geq_l(s->v); }

int bpo_r1cs_verify(const u8 *label, size_t label_len, size_t gens_capacity, size_t n, size_t m, const u8 *V32, size_t q,
                    const uint32_t *row_ptr, const uint32_t *term_var, const u8 *term_coeff, const u8 *proof, size_t proof_len,
                    const u8 ext_rng32[32], int flags) {
    bpo_init();
    /* R1CSProof::from_bytes */
    const u8 *A[6]; u8 Z32[32]; memset(Z32, 0, 32);
    const u8 *f; size_t nf;
    if (flags & 1) {
        if (proof_len % 32 || proof_len < 14 * 32) return 0;
        for (int i = 0; i < 6; i++) A[i] = proof + 32 * i;
        f = proof + 192; nf = proof_len / 32 - 6;
    } else {
        if (proof_len == 0) return 0;
        int ver = proof[0]; size_t body = proof_len - 1;
        if (body % 32 || (ver != 0 && ver != 1)) return 0;
        if (body < (size_t)(ver == 0 ? 11 : 14) * 32) return 0;
        if (ver == 0) { for (int i = 0; i < 3; i++) { A[i] = proof + 1 + 32 * i; A[3 + i] = Z32; } f = proof + 1 + 96; nf = body / 32 - 3; }
        else { for (int i = 0; i < 6; i++) A[i] = proof + 1 + 32 * i; f = proof + 1 + 192; nf = body / 32 - 6; }
    }
    if (nf < 10 || (nf - 10) % 2) return 0;
    const u8 *Tp[5]; for (int k = 0; k < 5; k++) Tp[k] = f + 32 * k;
    sc tx, txb, eb, ia, ibb;
    if (!sc_load_canonical(&tx, f + 160) || !sc_load_canonical(&txb, f + 192) || !sc_load_canonical(&eb, f + 224)) return 0;
    size_t lg = (nf - 10) / 2;
    if (lg >= 32) return 0;
    const u8 *LR = f + 256;
    if (!sc_load_canonical(&ia, LR + 64 * lg) || !sc_load_canonical(&ibb, LR + 64 * lg + 32)) return 0;

    strobe t; tr_init(&t, label, label_len);
    tr_append(&t, "dom-sep", (const u8 *)"r1cs v1", 7);
    for (size_t i = 0; i < m; i++) tr_append(&t, "V", V32 + 32 * i, 32);
    tr_append_u64(&t, "m", m);
    static const char *AL[6] = {"A_I1", "A_O1", "S1", "A_I2", "A_O2", "S2"};
    for (int i = 0; i < 3; i++) if (!tr_validate_and_append_point(&t, AL[i], A[i])) return 0;
    tr_append(&t, "dom-sep", (const u8 *)"r1cs-1phase", 11);
    size_t N = next_pow2(n);
    if (gens_capacity < N) return 0;
    for (int i = 3; i < 6; i++) tr_append(&t, AL[i], A[i], 32);
    sc y, z, u, x, w;
    tr_challenge_scalar(&t, "y", &y); tr_challenge_scalar(&t, "z", &z);
    static const char *TL[5] = {"T_1", "T_3", "T_4", "T_5", "T_6"};
    for (int k = 0; k < 5; k++) if (!tr_validate_and_append_point(&t, TL[k], Tp[k])) return 0;
    tr_challenge_scalar(&t, "u", &u); tr_challenge_scalar(&t, "x", &x);
    tr_append_scalar(&t, "t_x", &tx); tr_append_scalar(&t, "t_x_blinding", &txb); tr_append_scalar(&t, "e_blinding", &eb);
    tr_challenge_scalar(&t, "w", &w);
    /* ipp verification scalars */
    if (N != ((size_t)1 << lg)) return 0;
    tr_append(&t, "dom-sep", (const u8 *)"ipp v1", 6); tr_append_u64(&t, "n", N);
    sc ch[32], chi[32], usq[32], uisq[32], allinv = SC_ONE;
    for (size_t j = 0; j < lg; j++) {
        if (!tr_validate_and_append_point(&t, "L", LR + 64 * j)) return 0;
        if (!tr_validate_and_append_point(&t, "R", LR + 64 * j + 32)) return 0;
        tr_challenge_scalar(&t, "u", &ch[j]);
        sc_invert(&chi[j], &ch[j]); sc_mul(&allinv, &allinv, &chi[j]);
        sc_mul(&usq[j], &ch[j], &ch[j]); sc_mul(&uisq[j], &chi[j], &chi[j]);
    }
    gens_ensure(N);
    sc *wL = (sc *)malloc((N + 1) * sizeof(sc)), *wR = (sc *)malloc((N + 1) * sizeof(sc)), *wO = (sc *)malloc((N + 1) * sizeof(sc)), *wV = (sc *)malloc((m + 1) * sizeof(sc)), wc;
    flatten(&z, n, m, q, row_ptr, term_var, term_coeff, wL, wR, wO, wV, &wc);
    for (size_t i = n; i < N; i++) wL[i] = wR[i] = wO[i] = SC_ZERO;
    sc *s = (sc *)malloc(N * sizeof(sc)), *yi = (sc *)malloc(N * sizeof(sc));
    s[0] = allinv;
    for (size_t i = 1; i < N; i++) { int lgi = 63 - __builtin_clzll((unsigned long long)i); sc_mul(&s[i], &s[i - ((size_t)1 << lgi)], &usq[lg - 1 - lgi]); }
    sc yinv; sc_invert(&yinv, &y);
    yi[0] = SC_ONE; for (size_t i = 1; i < N; i++) sc_mul(&yi[i], &yi[i - 1], &yinv);
    size_t tot = 6 + m + 5 + 2 + 2 * N + 2 * lg;
    sc *ms = (sc *)malloc(tot * sizeof(sc)); ge *mp = (ge *)malloc(tot * sizeof(ge));
    sc *ynwR = (sc *)malloc(N * sizeof(sc));
    for (size_t i = 0; i < N; i++) sc_mul(&ynwR[i], &wR[i], &yi[i]);
    sc delta; sc_inner(&delta, ynwR, wL, n);
    strobe rng; rng_begin(&rng, &t); rng_finalize(&rng, ext_rng32);
    sc r; rng_scalar(&rng, &r);
    sc xx, xxx, rxx, tmp, tmp2;
    sc_mul(&xx, &x, &x); sc_mul(&xxx, &xx, &x); sc_mul(&rxx, &r, &xx);
    size_t c = 0; int ok = 1;
    ms[c] = x; ms[c + 1] = xx; ms[c + 2] = xxx; sc_mul(&ms[c + 3], &u, &x); sc_mul(&ms[c + 4], &u, &xx); sc_mul(&ms[c + 5], &u, &xxx);
    for (int i = 0; i < 6; i++) if (!ristretto_decode(&mp[c + i], A[i])) ok = 0;
    c += 6;
    for (size_t j = 0; j < m; j++) { sc_mul(&ms[c], &wV[j], &rxx); if (!ristretto_decode(&mp[c], V32 + 32 * j)) ok = 0; c++; }
    sc_mul(&ms[c], &r, &x); sc_mul(&ms[c + 1], &rxx, &x); sc_mul(&ms[c + 2], &rxx, &xx); sc_mul(&ms[c + 3], &rxx, &xxx); sc_mul(&tmp, &rxx, &xx); sc_mul(&ms[c + 4], &tmp, &xx);
    for (int k = 0; k < 5; k++) if (!ristretto_decode(&mp[c + k], Tp[k])) ok = 0;
    c += 5;
    /* B: w(t_x - ab) + r(x^2(wc + delta) - t_x) */
    sc ab; sc_mul(&ab, &ia, &ibb); sc_sub(&tmp, &tx, &ab); sc_mul(&tmp, &w, &tmp);
    sc_add(&tmp2, &wc, &delta); sc_mul(&tmp2, &xx, &tmp2); sc_sub(&tmp2, &tmp2, &tx); sc_muladd(&ms[c], &r, &tmp2, &tmp); mp[c] = G_B; c++;
    sc_mul(&tmp, &r, &txb); sc_add(&tmp, &tmp, &eb); sc_neg(&ms[c], &tmp); mp[c] = G_BBLIND; c++;
#pragma omp parallel for num_threads(g_threads) schedule(static)
    for (long i = 0; i < (long)N; i++) {
        sc uf = (size_t)i < n ? SC_ONE : u, g, h, a1, b1;
        sc_mul(&g, &x, &ynwR[i]); sc_mul(&a1, &ia, &s[i]); sc_sub(&g, &g, &a1); sc_mul(&ms[c + i], &uf, &g);
        sc_muladd(&h, &x, &wL[i], &wO[i]); sc_mul(&b1, &ibb, &s[N - 1 - i]); sc_sub(&h, &h, &b1); sc_mul(&h, &yi[i], &h); sc_sub(&h, &h, &SC_ONE);
        sc_mul(&ms[c + N + i], &uf, &h);
        mp[c + i] = g_G[i]; mp[c + N + i] = g_H[i];
    }
    c += 2 * N;
    for (size_t j = 0; j < lg; j++) { ms[c] = usq[j]; if (!ristretto_decode(&mp[c], LR + 64 * j)) ok = 0; c++; }
    for (size_t j = 0; j < lg; j++) { ms[c] = uisq[j]; if (!ristretto_decode(&mp[c], LR + 64 * j + 32)) ok = 0; c++; }
    int accept = 0;
    if (ok) { ge res; msm_run(&res, ms, mp, tot, BPO_MSM_VARTIME); accept = ge_is_identity_coset(&res); }
    free(wL); free(wR); free(wO); free(wV); free(s); free(yi); free(ms); free(mp); free(ynwR);
    return accept;
}
