// ge25519.cuh -- extended twisted-Edwards points over GF(2^255-19), ristretto255 encode/decode,
// Elligator map.  Device equivalents of curve25519-dalek 1.x EdwardsPoint / RistrettoPoint /
// CompressedRistretto (reference dependency, /root/reference/Cargo.toml:8; restated per RFC 9496
// and SURVEY.md App. A.2).  Host bodies exist for CPU unit tests and for the host-side glue.
#pragma once
#include "fe25519.cuh"

struct ge { fe X, Y, Z, T; };        // extended: x=X/Z, y=Y/Z, T=XY/Z          (128 B)
struct ge_pn { fe YpX, YmX, Z, T2d; }; // projective Niels (for variable points)   (128 B)
struct ge_an { fe ypx, ymx, t2d; };  // affine Niels (Z=1; resident generators)   ( 96 B)

// Curve constants live in one struct: h_K on the host (filled by bpg_init_constants(), which derives
// everything from d and sqrt(-1) and self-checks) and c_K in __constant__ memory on the device.
struct bpg_consts { fe d, d2, sqrtm1, invsqrt_a_minus_d, sqrt_ad_minus_one, one_minus_d_sq, d_minus_one_sq; };
#ifdef __CUDACC__
static __constant__ bpg_consts c_K; // libbpg is a single translation unit (bpg.cu)
#endif
extern bpg_consts h_K;
#ifdef __CUDA_ARCH__
#define KK c_K
#else
#define KK h_K
#endif

BPG_HD void ge_identity(ge &p) { fe_set0(p.X); fe_set1(p.Y); fe_set1(p.Z); fe_set0(p.T); }
BPG_HD void ge_neg(ge &r, const ge &p) { fe_neg(r.X, p.X); r.Y = p.Y; r.Z = p.Z; fe_neg(r.T, p.T); }
BPG_HD void ge_to_pn(ge_pn &c, const ge &p) { fe_add(c.YpX, p.Y, p.X); fe_sub(c.YmX, p.Y, p.X); c.Z = p.Z; fe_mul(c.T2d, p.T, KK.d2); }
// requires Z == 1
BPG_HD void ge_affine_to_an(ge_an &c, const fe &x, const fe &y) { fe t; fe_add(c.ypx, y, x); fe_sub(c.ymx, y, x); fe_mul(t, x, y); fe_mul(c.t2d, t, KK.d2); }
BPG_HD void ge_an_neg(ge_an &r, const ge_an &p) { r.ypx = p.ymx; r.ymx = p.ypx; fe_neg(r.t2d, p.t2d); }
BPG_HD void ge_pn_neg(ge_pn &r, const ge_pn &p) { r.YpX = p.YmX; r.YmX = p.YpX; r.Z = p.Z; fe_neg(r.T2d, p.T2d); }

// r = p + q  (add-2008-hwcd-3, complete for a = -1): 8M with projective Niels
BPG_HD void ge_add_pn(ge &r, const ge &p, const ge_pn &q) {
    fe a, b, c, d, e, f, g, h, t;
    fe_sub(t, p.Y, p.X); fe_mul(a, t, q.YmX);
    fe_add(t, p.Y, p.X); fe_mul(b, t, q.YpX);
    fe_mul(c, p.T, q.T2d);
    fe_mul(d, p.Z, q.Z); fe_dbl(d, d);
    fe_sub(e, b, a); fe_sub(f, d, c); fe_add(g, d, c); fe_add(h, b, a);
    fe_mul(r.X, e, f); fe_mul(r.Y, g, h); fe_mul(r.Z, f, g); fe_mul(r.T, e, h);
}
// r = p + q with q affine Niels: 7M
BPG_HD void ge_add_an(ge &r, const ge &p, const ge_an &q) {
    fe a, b, c, d, e, f, g, h, t;
    fe_sub(t, p.Y, p.X); fe_mul(a, t, q.ymx);
    fe_add(t, p.Y, p.X); fe_mul(b, t, q.ypx);
    fe_mul(c, p.T, q.t2d);
    fe_dbl(d, p.Z);
    fe_sub(e, b, a); fe_sub(f, d, c); fe_add(g, d, c); fe_add(h, b, a);
    fe_mul(r.X, e, f); fe_mul(r.Y, g, h); fe_mul(r.Z, f, g); fe_mul(r.T, e, h);
}
// same addition with the two groups of four independent multiplications interleaved (see fe_mul4)
BPG_HD void ge_add_pn_ilp(ge &r, const ge &p, const ge_pn &q) {
    fe a, b, c, d, e, f, g, h, t0, t1;
    fe_sub(t0, p.Y, p.X); fe_add(t1, p.Y, p.X);
    fe_mul4(a, t0, q.YmX, b, t1, q.YpX, c, p.T, q.T2d, d, p.Z, q.Z);
    fe_dbl(d, d);
    fe_sub(e, b, a); fe_sub(f, d, c); fe_add(g, d, c); fe_add(h, b, a);
    fe_mul4(r.X, e, f, r.Y, g, h, r.Z, f, g, r.T, e, h);
}
// mixed addition with an affine-Niels operand: 3 + 4 interleaved multiplications
BPG_HD void ge_add_an_ilp(ge &r, const ge &p, const ge_an &q) {
    fe a, b, c, d, e, f, g, h, t0, t1;
    fe_sub(t0, p.Y, p.X); fe_add(t1, p.Y, p.X);
    fe_mul3(a, t0, q.ymx, b, t1, q.ypx, c, p.T, q.t2d);
    fe_dbl(d, p.Z);
    fe_sub(e, b, a); fe_sub(f, d, c); fe_add(g, d, c); fe_add(h, b, a);
    fe_mul4(r.X, e, f, r.Y, g, h, r.Z, f, g, r.T, e, h);
}
BPG_HD void ge_add_ilp(ge &r, const ge &p, const ge &q) { ge_pn c; ge_to_pn(c, q); ge_add_pn_ilp(r, p, c); }
BPG_HD void ge_dbl_ilp(ge &r, const ge &p) {
    fe a, b, c, d, e, f, g, h, t;
    fe_add(t, p.X, p.Y);
    fe_mul4(a, p.X, p.X, b, p.Y, p.Y, c, p.Z, p.Z, t, t, t);
    fe_dbl(c, c);
    fe_neg(d, a);
    fe_sub(e, t, a); fe_sub(e, e, b);
    fe_add(g, d, b); fe_sub(f, g, c); fe_sub(h, d, b);
    fe_mul4(r.X, e, f, r.Y, g, h, r.Z, f, g, r.T, e, h);
}
BPG_HD void ge_add(ge &r, const ge &p, const ge &q) { ge_pn c; ge_to_pn(c, q); ge_add_pn(r, p, c); }
BPG_HD void ge_sub(ge &r, const ge &p, const ge &q) { ge_pn c, n; ge_to_pn(c, q); ge_pn_neg(n, c); ge_add_pn(r, p, n); }
// dbl-2008-hwcd, a = -1: 4M + 4S
BPG_HD void ge_dbl(ge &r, const ge &p) {
    fe a, b, c, d, e, f, g, h, t;
    fe_sqr(a, p.X); fe_sqr(b, p.Y); fe_sqr(c, p.Z); fe_dbl(c, c);
    fe_neg(d, a);
    fe_add(t, p.X, p.Y); fe_sqr(t, t); fe_sub(e, t, a); fe_sub(e, e, b);
    fe_add(g, d, b); fe_sub(f, g, c); fe_sub(h, d, b);
    fe_mul(r.X, e, f); fe_mul(r.Y, g, h); fe_mul(r.Z, f, g); fe_mul(r.T, e, h);
}
// identity of the ristretto group: the whole coset E[4]  <=>  X*Y == 0
BPG_HD int ge_is_identity_coset(const ge &p) { fe t; fe_mul(t, p.X, p.Y); return fe_iszero(t); }

// RFC 9496 4.2 SQRT_RATIO_M1
BPG_HD int fe_sqrt_ratio_m1(fe &r_out, const fe &u, const fe &v) {
    fe v3, v7, r, check, t, neg_u, neg_u_i;
    fe_sqr(t, v); fe_mul(v3, t, v);
    fe_sqr(t, v3); fe_mul(v7, t, v);
    fe_mul(t, u, v7); fe_pow22523(t, t);
    fe_mul(r, u, v3); fe_mul(r, r, t);
    fe_sqr(t, r); fe_mul(check, v, t);
    fe_neg(neg_u, u); fe_mul(neg_u_i, neg_u, KK.sqrtm1);
    int correct = fe_eq(check, u), flipped = fe_eq(check, neg_u), flipped_i = fe_eq(check, neg_u_i);
    fe ri; fe_mul(ri, r, KK.sqrtm1);
    fe_cmov(r, ri, flipped | flipped_i);
    fe_abs(r_out, r);
    return correct | flipped;
}
BPG_HD int fe_bytes_canonical_nonneg(const u8 *b) { // s < p and s even
    fe s, c; fe_frombytes(s, b); fe_canon(c, s);
    u8 chk[32]; fe_tobytes(chk, c);
    int same = 1;
    for (int i = 0; i < 32; i++) same &= (chk[i] == b[i]);
    return same && !(b[0] & 1);
}
// RFC 9496 4.3.1; returns 1 on success
BPG_HD int ristretto_decode(ge &p, const u8 *b) {
    fe s, ss, u1, u2, u2s, v, t, I, Dx, Dy, x, y, one;
    fe_set1(one);
    int canon = fe_bytes_canonical_nonneg(b);
    fe_frombytes(s, b);
    fe_sqr(ss, s); fe_sub(u1, one, ss); fe_add(u2, one, ss); fe_sqr(u2s, u2);
    fe_sqr(t, u1); fe_mul(t, t, KK.d); fe_neg(t, t); fe_sub(v, t, u2s);
    fe_mul(t, v, u2s);
    int ok = fe_sqrt_ratio_m1(I, one, t);
    fe_mul(Dx, I, u2); fe_mul(Dy, I, Dx); fe_mul(Dy, Dy, v);
    fe_mul(x, s, Dx); fe_dbl(x, x); fe_abs(x, x);
    fe_mul(y, u1, Dy); fe_mul(t, x, y);
    p.X = x; p.Y = y; p.Z = one; p.T = t;
    return canon && ok && !fe_isneg(t) && !fe_iszero(y);
}
// RFC 9496 4.3.2
BPG_HD void ristretto_encode(u8 *out, const ge &p) {
    fe u1, u2, t, I, d1, d2, zinv, X, Y, den, s, ix, iy, dalt, ny, one;
    fe_set1(one);
    fe_add(u1, p.Z, p.Y); fe_sub(t, p.Z, p.Y); fe_mul(u1, u1, t);
    fe_mul(u2, p.X, p.Y);
    fe_sqr(t, u2); fe_mul(t, t, u1);
    (void)fe_sqrt_ratio_m1(I, one, t);
    fe_mul(d1, I, u1); fe_mul(d2, I, u2);
    fe_mul(zinv, d1, d2); fe_mul(zinv, zinv, p.T);
    fe_mul(t, p.T, zinv);
    int rot = fe_isneg(t);
    fe_mul(ix, p.X, KK.sqrtm1); fe_mul(iy, p.Y, KK.sqrtm1); fe_mul(dalt, d1, KK.invsqrt_a_minus_d);
    X = p.X; Y = p.Y; den = d2;
    fe_cmov(X, iy, rot); fe_cmov(Y, ix, rot); fe_cmov(den, dalt, rot);
    fe_mul(t, X, zinv);
    fe_neg(ny, Y); fe_cmov(Y, ny, fe_isneg(t));
    fe_sub(t, p.Z, Y); fe_mul(s, den, t); fe_abs(s, s);
    fe_tobytes(out, s);
}
// RFC 9496 4.3.4 MAP
BPG_HD void elligator_map(ge &p, const fe &t0) {
    fe r, u, v, t, s, s_prime, c, N, w0, w1, w2, w3, ss, one, mone;
    fe_set1(one); fe_neg(mone, one);
    fe_sqr(r, t0); fe_mul(r, r, KK.sqrtm1);
    fe_add(u, r, one); fe_mul(u, u, KK.one_minus_d_sq);
    fe_mul(t, r, KK.d); fe_neg(t, t); fe_sub(t, t, one);
    fe_add(v, r, KK.d); fe_mul(v, v, t);
    int sq = fe_sqrt_ratio_m1(s, u, v);
    fe_mul(s_prime, s, t0); fe_abs(s_prime, s_prime); fe_neg(s_prime, s_prime);
    c = mone;
    fe_cmov(s, s_prime, !sq); fe_cmov(c, r, !sq);
    fe_sub(t, r, one); fe_mul(N, c, t); fe_mul(N, N, KK.d_minus_one_sq); fe_sub(N, N, v);
    fe_mul(w0, s, v); fe_dbl(w0, w0);
    fe_mul(w1, N, KK.sqrt_ad_minus_one);
    fe_sqr(ss, s); fe_sub(w2, one, ss); fe_add(w3, one, ss);
    fe_mul(p.X, w0, w3); fe_mul(p.Y, w2, w1); fe_mul(p.Z, w1, w3); fe_mul(p.T, w0, w2);
}
// RistrettoPoint::from_uniform_bytes
BPG_HD void ge_from_uniform_bytes(ge &p, const u8 *b64) {
    fe r1, r2; ge p1, p2;
    fe_frombytes(r1, b64); fe_frombytes(r2, b64 + 32);
    elligator_map(p1, r1); elligator_map(p2, r2);
    ge_add(p, p1, p2);
}
// plain double-and-add (host glue / tests; k must be reduced)
BPG_HD void ge_scalarmul(ge &r, const sc &k, const ge &p) {
    ge acc; ge_identity(acc);
    ge_pn pc; ge_to_pn(pc, p);
#pragma unroll 1
    for (int i = 252; i >= 0; i--) {
        ge_dbl(acc, acc);
        if ((k.v[i >> 5] >> (i & 31)) & 1) ge_add_pn(acc, acc, pc);
    }
    r = acc;
}
