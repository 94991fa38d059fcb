// consts.h -- host-side initialisation of the curve constants (h_K).  Everything is derived from
// d = -121665/121666 and sqrt(-1) = 2^((p-1)/4) and self-checked, rather than transcribed.
#pragma once
#include "ge25519.cuh"

inline int bpg_init_constants_host() {
    static int done = 0;
    if (done) return 0;
    fe one, t, a, b;
    fe_set1(one);
    // d = -121665 / 121666
    fe n121665, n121666, inv;
    fe_set0(n121665); n121665.v[0] = 121665u;
    fe_set0(n121666); n121666.v[0] = 121666u;
    fe_invert(inv, n121666);
    fe_mul(t, n121665, inv);
    fe_neg(h_K.d, t);
    fe_add(h_K.d2, h_K.d, h_K.d);
    // sqrt(-1) = 2^((p-1)/4) ; (p-1)/4 = 2^253 - 5 = 4*(2^251 - 2) + 3  => 2^((p-1)/4) = ((2^(2^251-2)))^4 * 8
    // compute via pow22523: x^((p-5)/8) = x^(2^252-3).  2^((p-1)/4) = (2^(2^252-3))^2 * 2^1 ... (2*(2^252-3) + 1 = 2^253 - 5)
    fe two; fe_add(two, one, one);
    fe_pow22523(t, two);
    fe_sqr(t, t);
    fe_mul(h_K.sqrtm1, t, two);
    fe_sqr(t, h_K.sqrtm1);
    fe_add(t, t, one);
    if (!fe_iszero(t)) return -1; // sqrt(-1)^2 != -1
    // self-check d: d*121666 + 121665 == 0
    fe_mul(t, h_K.d, n121666); fe_add(t, t, n121665);
    if (!fe_iszero(t)) return -2;
    fe_neg(a, one); fe_sub(a, a, h_K.d); // a - d = -1 - d
    (void)fe_sqrt_ratio_m1(h_K.invsqrt_a_minus_d, one, a);
    (void)fe_sqrt_ratio_m1(b, a, one); // sqrt(ad - 1) = sqrt(-d - 1): the ODD root
    if (!fe_isneg(b)) fe_neg(b, b);
    h_K.sqrt_ad_minus_one = b;
    fe_sqr(t, b);
    if (!fe_eq(t, a)) return -3;
    fe_sqr(t, h_K.d); fe_sub(h_K.one_minus_d_sq, one, t);
    fe_sub(t, h_K.d, one); fe_sqr(h_K.d_minus_one_sq, t);
    done = 1;
    return 0;
}
