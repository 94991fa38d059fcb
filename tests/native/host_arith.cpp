// Host-only build of the product's limb arithmetic headers (the portable bodies of fe25519.cuh /
// ge25519.cuh) exposed through a C ABI so tests/test_host_arith.py can compare them with the big-int
// oracle on CPU.  The device bodies (PTX carry chains) are checked on the GPU by tests/test_gpu_*.py.
#include "../../bulletproofs_gadgets_b200/csrc/consts.h"
#include "../../bulletproofs_gadgets_b200/csrc/host_scalar64.h"
bpg_consts h_K;
extern "C" {
int ht_init() { return bpg_init_constants_host(); }
void ht_consts(uint8_t *out7x32) {
    const fe *k[7] = {&h_K.d, &h_K.d2, &h_K.sqrtm1, &h_K.invsqrt_a_minus_d, &h_K.sqrt_ad_minus_one, &h_K.one_minus_d_sq, &h_K.d_minus_one_sq};
    for (int i = 0; i < 7; i++) fe_tobytes(out7x32 + 32 * i, *k[i]);
}
static void ldfe(fe &r, const uint8_t *b) { for (int i = 0; i < 8; i++) r.v[i] = (u32)b[4*i] | ((u32)b[4*i+1] << 8) | ((u32)b[4*i+2] << 16) | ((u32)b[4*i+3] << 24); }
// raw 256-bit in, canonical out
void ht_fe_mul(const uint8_t *a, const uint8_t *b, uint8_t *o) { fe x, y, r; ldfe(x, a); ldfe(y, b); fe_mul(r, x, y); fe_tobytes(o, r); }
void ht_fe_sqr(const uint8_t *a, uint8_t *o) { fe x, r; ldfe(x, a); fe_sqr(r, x); fe_tobytes(o, r); }
void ht_fe_add(const uint8_t *a, const uint8_t *b, uint8_t *o) { fe x, y, r; ldfe(x, a); ldfe(y, b); fe_add(r, x, y); fe_tobytes(o, r); }
void ht_fe_sub(const uint8_t *a, const uint8_t *b, uint8_t *o) { fe x, y, r; ldfe(x, a); ldfe(y, b); fe_sub(r, x, y); fe_tobytes(o, r); }
void ht_fe_inv(const uint8_t *a, uint8_t *o) { fe x, r; ldfe(x, a); fe_invert(r, x); fe_tobytes(o, r); }
void ht_sc_mul(const uint8_t *a, const uint8_t *b, uint8_t *o) { sc x, y, r; sc_frombytes(x, a); sc_frombytes(y, b); sc_mul(r, x, y); sc_tobytes(o, r); }
void ht_sc_reduce(const uint8_t *a, uint8_t *o) { sc x, r; sc_frombytes(x, a); sc_reduce(r, x); sc_tobytes(o, r); }
void ht_sc_addsub(const uint8_t *a, const uint8_t *b, uint8_t *oadd, uint8_t *osub) { sc x, y, r; sc_frombytes(x, a); sc_frombytes(y, b); sc_add_r(r, x, y); sc_tobytes(oadd, r); sc_sub_r(r, x, y); sc_tobytes(osub, r); }
void ht_sc_invert(const uint8_t *a, uint8_t *o) { sc x, r; sc_frombytes(x, a); sc_invert(r, x); sc_tobytes(o, r); }
void ht_sc_wide(const uint8_t *a64, uint8_t *o) { u32 R[16]; for (int i = 0; i < 16; i++) R[i] = (u32)a64[4*i] | ((u32)a64[4*i+1] << 8) | ((u32)a64[4*i+2] << 16) | ((u32)a64[4*i+3] << 24); sc r; sc_reduce512(r, R); sc_tobytes(o, r); }
int ht_decode_encode(const uint8_t *in, uint8_t *out) { ge p; if (!ristretto_decode(p, in)) return 0; ristretto_encode(out, p); return 1; }
int ht_point_mul_add(const uint8_t *k, const uint8_t *p32, const uint8_t *q32, uint8_t *out) { // k*P + Q, then doubled once more via dbl path: out = k*P + Q ; out2 = 2*out
    ge p, q, r; sc s; if (!ristretto_decode(p, p32) || !ristretto_decode(q, q32)) return 0;
    sc_frombytes(s, k); sc_reduce(s, s); ge_scalarmul(r, s, p); ge_add(r, r, q); ristretto_encode(out, r);
    ge d; ge_dbl(d, r); ristretto_encode(out + 32, d);
    // affine-Niels path: normalise q, add as affine Niels
    fe zi, x, y; fe_invert(zi, q.Z); fe_mul(x, q.X, zi); fe_mul(y, q.Y, zi);
    ge_an an; ge_affine_to_an(an, x, y); ge r2; ge_scalarmul(r2, s, p); ge_add_an_ilp(r2, r2, an); ristretto_encode(out + 64, r2);
    ge_an nn; ge_an_neg(nn, an); ge_add_an(r2, r2, nn); ge r3; ge_scalarmul(r3, s, p); ge_sub(r3, r3, r2); out[96] = (uint8_t)ge_is_identity_coset(r3);
    return 1;
}
// ILP variants must agree with the plain ones: out = (P + Q) via ge_add_ilp, out+32 = 2P via ge_dbl_ilp
int ht_ilp(const uint8_t *p32, const uint8_t *q32, uint8_t *out) {
    ge p, q, r; if (!ristretto_decode(p, p32) || !ristretto_decode(q, q32)) return 0;
    ge_add_ilp(r, p, q); ristretto_encode(out, r);
    ge_dbl_ilp(r, p); ristretto_encode(out + 32, r);
    return 1;
}
void ht_sc64(const uint8_t *a, const uint8_t *b, const uint8_t *w64, uint8_t *omul, uint8_t *oinv, uint8_t *owide) {
    sc x, y; sc_frombytes(x, a); sc_frombytes(y, b);
    sc_tobytes(omul, bpgh::sc_mul64(x, y)); sc_tobytes(oinv, bpgh::sc_invert64(x)); sc_tobytes(owide, bpgh::sc_wide64(w64));
}
void ht_from_uniform(const uint8_t *b64, uint8_t *out) { ge p; ge_from_uniform_bytes(p, b64); ristretto_encode(out, p); }
}
