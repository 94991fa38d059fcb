// fe25519.cuh -- GF(2^255-19) and Z_l arithmetic in 8 x 32-bit saturated limbs for sm_100a.
//
// Replaces (device side) the field/scalar backends of curve25519-dalek 1.x that the reference
// reaches through /root/reference/Cargo.toml:8 (Scalar used directly at src/mimc_hash/mimc.rs:7-23).
//
// Multiplication is an even/odd split of the 8x8 schoolbook product into carry chains of
// mad.lo.cc / madc.hi.cc pairs; ptxas fuses each pair into one IMAD.WIDE.U32.X, so a field
// multiply is ~72 IMAD.WIDE (64 products + 8 for the 2^256 = 38 fold) plus ~45 integer adds.
// Field elements are kept "weakly reduced": any value in [0, 2^256) congruent mod p.
// Every function also has a portable host body so the exact limb logic is unit-tested on CPU.
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define BPG_HD __host__ __device__ __forceinline__
#define BPG_D __device__ __forceinline__
#else
#define BPG_HD inline
#define BPG_D inline
#endif

typedef uint32_t u32;
typedef uint64_t u64;
typedef uint8_t u8;

struct fe { u32 v[8]; };
struct sc { u32 v[8]; };

// ---------------------------------------------------------------- carry-chain primitives
// acc[0..7] += {x0,x1,x2,x3} * b placed at 64-bit slots; acc[8] += carry-out
BPG_HD void mac4(u32 *acc, u32 x0, u32 x1, u32 x2, u32 x3, u32 b) {
#ifdef __CUDA_ARCH__
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7]), "+r"(acc[8])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
#else
    u32 xs[4] = {x0, x1, x2, x3};
    u64 c = 0;
    for (int k = 0; k < 4; k++) {
        u64 p = (u64)xs[k] * b;
        u64 s = (u64)acc[2 * k] + (u32)p + c;
        acc[2 * k] = (u32)s;
        c = s >> 32;
        s = (u64)acc[2 * k + 1] + (u32)(p >> 32) + c;
        acc[2 * k + 1] = (u32)s;
        c = s >> 32;
    }
    acc[8] += (u32)c;
#endif
}

// r[0..n-1] = a[0..n-1] + b[0..n-1], returns carry (n = 8)
BPG_HD u32 add8(u32 *r, const u32 *a, const u32 *b) {
    u32 c;
#ifdef __CUDA_ARCH__
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]), "r"(b[1]),
          "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    u64 s = 0;
    for (int i = 0; i < 8; i++) { s += (u64)a[i] + b[i]; r[i] = (u32)s; s >>= 32; }
    c = (u32)s;
#endif
    return c;
}
// r = a - b, returns borrow (0/1)
BPG_HD u32 sub8(u32 *r, const u32 *a, const u32 *b) {
    u32 c;
#ifdef __CUDA_ARCH__
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]), "r"(b[1]),
          "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    c &= 1u;
#else
    u64 br = 0;
    for (int i = 0; i < 8; i++) { u64 d = (u64)a[i] - b[i] - br; r[i] = (u32)d; br = (d >> 32) & 1; }
    c = (u32)br;
#endif
    return c;
}
// r[0..7] += small (32-bit), returns carry
BPG_HD u32 addsmall8(u32 *r, u32 x) {
    u32 c;
#ifdef __CUDA_ARCH__
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.cc.u32 %7, %7, 0;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "=r"(c)
        : "r"(x));
#else
    u64 s = x;
    for (int i = 0; i < 8; i++) { s += r[i]; r[i] = (u32)s; s >>= 32; }
    c = (u32)s;
#endif
    return c;
}
BPG_HD u32 subsmall8(u32 *r, u32 x) {
    u32 c;
#ifdef __CUDA_ARCH__
    asm("sub.cc.u32 %0, %0, %9;\n\t"
        "subc.cc.u32 %1, %1, 0;\n\t"
        "subc.cc.u32 %2, %2, 0;\n\t"
        "subc.cc.u32 %3, %3, 0;\n\t"
        "subc.cc.u32 %4, %4, 0;\n\t"
        "subc.cc.u32 %5, %5, 0;\n\t"
        "subc.cc.u32 %6, %6, 0;\n\t"
        "subc.cc.u32 %7, %7, 0;\n\t"
        "subc.u32 %8, 0, 0;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "=r"(c)
        : "r"(x));
    c &= 1u;
#else
    u64 br = x;
    u32 cc = 0;
    for (int i = 0; i < 8; i++) { u64 d = (u64)r[i] - (i == 0 ? br : 0) - cc; r[i] = (u32)d; cc = (u32)((d >> 32) & 1); }
    c = cc;
#endif
    return c;
}

// 512-bit product R[0..15] = a * b (8 x 8 limbs)
BPG_HD void mul512(u32 *R, const u32 *a, const u32 *b) {
    u32 E[18], O[18];
#pragma unroll
    for (int i = 0; i < 18; i++) { E[i] = 0; O[i] = 0; }
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if ((j & 1) == 0) {
            mac4(E + j, a[0], a[2], a[4], a[6], b[j]);
            mac4(O + j, a[1], a[3], a[5], a[7], b[j]);
        } else {
            mac4(O + j - 1, a[0], a[2], a[4], a[6], b[j]);
            mac4(E + j + 1, a[1], a[3], a[5], a[7], b[j]);
        }
    }
    // R = E + (O << 32)
    R[0] = E[0];
    u32 c = add8(R + 1, E + 1, O);
    u32 t[8];
    u32 c2 = add8(t, E + 9, O + 8); // words 9..16 (E[16], O[15] are zero for a valid product)
    (void)c2;
    c = addsmall8(t, c);
    (void)c;
#pragma unroll
    for (int i = 0; i < 7; i++) R[9 + i] = t[i];
}

// same as mac4 but never reordered: used by the 4-way interleaved multiply below to force instruction-level parallelism
BPG_HD void mac4v(u32 *acc, u32 x0, u32 x1, u32 x2, u32 x3, u32 b) {
#ifdef __CUDA_ARCH__
    asm volatile("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7]), "+r"(acc[8])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
#else
    mac4(acc, x0, x1, x2, x3, b);
#endif
}
// NW independent 512-bit products with their carry chains interleaved row by row (2*NW independent chains in flight):
// a single warp then runs close to the IMAD.WIDE issue rate instead of waiting on one carry chain at a time.
template <int NW>
BPG_HD void mul512_ilp(u32 (*R)[16], const u32 *const *a, const u32 *const *b) {
    u32 E[NW][18], O[NW][18];
#pragma unroll
    for (int k = 0; k < NW; k++)
#pragma unroll
        for (int i = 0; i < 18; i++) { E[k][i] = 0; O[k][i] = 0; }
#pragma unroll
    for (int j = 0; j < 8; j++) {
#pragma unroll
        for (int k = 0; k < NW; k++) {
            if ((j & 1) == 0) {
                mac4v(E[k] + j, a[k][0], a[k][2], a[k][4], a[k][6], b[k][j]);
                mac4v(O[k] + j, a[k][1], a[k][3], a[k][5], a[k][7], b[k][j]);
            } else {
                mac4v(O[k] + j - 1, a[k][0], a[k][2], a[k][4], a[k][6], b[k][j]);
                mac4v(E[k] + j + 1, a[k][1], a[k][3], a[k][5], a[k][7], b[k][j]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NW; k++) {
        R[k][0] = E[k][0];
        u32 c = add8(R[k] + 1, E[k] + 1, O[k]);
        u32 t[8];
        (void)add8(t, E[k] + 9, O[k] + 8);
        (void)addsmall8(t, c);
#pragma unroll
        for (int i = 0; i < 7; i++) R[k][9 + i] = t[i];
    }
}

// ---- squaring: 28 cross products (doubled by a 1-bit shift of the 512-bit partial) + 8 diagonal squares = 36 wide
// multiplies instead of 64.  Cross products a_i a_j (i < j) are accumulated in the same even/odd carry-chain style as
// mul512, with chains of 1..4 products.
template <int K>
BPG_HD void macK(u32 *acc, const u32 *x, u32 b) { // acc[0..2K-1] += {x[0], x[2], ..} * b at 64-bit slots ; acc[2K] += carry
#ifdef __CUDA_ARCH__
    if (K == 1) {
        asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]) : "r"(x[0]), "r"(b));
    } else if (K == 2) {
        asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\tmadc.hi.cc.u32 %1, %5, %7, %1;\n\tmadc.lo.cc.u32 %2, %6, %7, %2;\n\tmadc.hi.cc.u32 %3, %6, %7, %3;\n\t"
            "addc.u32 %4, %4, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]) : "r"(x[0]), "r"(x[2]), "r"(b));
    } else if (K == 3) {
        asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\tmadc.hi.cc.u32 %1, %7, %10, %1;\n\tmadc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\t"
            "madc.lo.cc.u32 %4, %9, %10, %4;\n\tmadc.hi.cc.u32 %5, %9, %10, %5;\n\taddc.u32 %6, %6, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6])
            : "r"(x[0]), "r"(x[2]), "r"(x[4]), "r"(b));
    } else {
        mac4(acc, x[0], x[2], x[4], x[6], b);
    }
#else
    u64 c = 0;
    for (int k = 0; k < K; k++) {
        u64 p = (u64)x[2 * k] * b;
        u64 t = (u64)acc[2 * k] + (u32)p + c;
        acc[2 * k] = (u32)t; c = t >> 32;
        t = (u64)acc[2 * k + 1] + (u32)(p >> 32) + c;
        acc[2 * k + 1] = (u32)t; c = t >> 32;
    }
    acc[2 * K] += (u32)c;
#endif
}
BPG_HD void sqr512(u32 *R, const u32 *a) {
    // E: products whose word position i + j is even, O: odd positions (stored shifted down by one word)
    u32 E[18], O[18];
#pragma unroll
    for (int i = 0; i < 18; i++) { E[i] = 0; O[i] = 0; }
    // row j multiplies a_j with the lower limbs a_i, i < j, split by the parity of i
    // j = 1: i = 0           -> position 1 (odd)
    macK<1>(O + 0, a + 0, a[1]);
    // j = 2: i = 0 -> pos 2 (even) ; i = 1 -> pos 3 (odd)
    macK<1>(E + 2, a + 0, a[2]);
    macK<1>(O + 2, a + 1, a[2]);
    // j = 3: i = 0, 2 -> pos 3, 5 (odd) ; i = 1 -> pos 4 (even)
    macK<2>(O + 2, a + 0, a[3]);
    macK<1>(E + 4, a + 1, a[3]);
    // j = 4: i = 0, 2 -> pos 4, 6 (even) ; i = 1, 3 -> pos 5, 7 (odd)
    macK<2>(E + 4, a + 0, a[4]);
    macK<2>(O + 4, a + 1, a[4]);
    // j = 5: i = 0, 2, 4 -> pos 5, 7, 9 (odd) ; i = 1, 3 -> pos 6, 8 (even)
    macK<3>(O + 4, a + 0, a[5]);
    macK<2>(E + 6, a + 1, a[5]);
    // j = 6: i = 0, 2, 4 -> pos 6, 8, 10 (even) ; i = 1, 3, 5 -> pos 7, 9, 11 (odd)
    macK<3>(E + 6, a + 0, a[6]);
    macK<3>(O + 6, a + 1, a[6]);
    // j = 7: i = 0, 2, 4, 6 -> pos 7, 9, 11, 13 (odd) ; i = 1, 3, 5 -> pos 8, 10, 12 (even)
    macK<4>(O + 6, a + 0, a[7]);
    macK<3>(E + 8, a + 1, a[7]);
    // C = E + (O << 32)
    u32 C[16];
    C[0] = E[0];
    u32 c = add8(C + 1, E + 1, O);
    u32 t[8];
    (void)add8(t, E + 9, O + 8);
    (void)addsmall8(t, c);
#pragma unroll
    for (int i = 0; i < 7; i++) C[9 + i] = t[i];
    // 2C (C < 2^511 because 2C <= a^2 < 2^512)
    u32 D2[16];
    D2[0] = C[0] << 1;
#pragma unroll
    for (int i = 1; i < 16; i++) D2[i] = (C[i] << 1) | (C[i - 1] >> 31);
    // diagonal squares
    u32 Dg[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { u64 p = (u64)a[i] * a[i]; Dg[2 * i] = (u32)p; Dg[2 * i + 1] = (u32)(p >> 32); }
    u32 c1 = add8(R, D2, Dg);
    u32 hi[8];
    (void)add8(hi, D2 + 8, Dg + 8);
    (void)addsmall8(hi, c1);
#pragma unroll
    for (int i = 0; i < 8; i++) R[8 + i] = hi[i];
}

// ---------------------------------------------------------------- field
// reduce a 512-bit product: 2^256 = 38 (mod p)
BPG_HD void fe_reduce512(fe &r, const u32 *R) {
    u32 F[10], G[10];
#pragma unroll
    for (int i = 0; i < 8; i++) { F[i] = R[i]; G[i] = 0; }
    F[8] = F[9] = 0;
    G[8] = G[9] = 0;
    mac4(F, R[8], R[10], R[12], R[14], 38u);
    mac4(G, R[9], R[11], R[13], R[15], 38u);
    // S = F + (G << 32), 9 words
    u32 S[8];
    S[0] = F[0];
    u32 hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { lo[i] = i < 7 ? F[i + 1] : F[8]; hi[i] = G[i]; }
    u32 t[8];
    (void)add8(t, lo, hi); // words 1..8 ; word 8 < 2^7 so no carry out
#pragma unroll
    for (int i = 0; i < 7; i++) S[i + 1] = t[i];
    u32 top = t[7];
    u32 c = addsmall8(S, top * 38u);
    S[0] += 38u & (0u - c); // wrapped value is tiny, cannot carry again
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = S[i];
}
BPG_HD void fe_mul(fe &r, const fe &a, const fe &b) {
    u32 R[16];
    mul512(R, a.v, b.v);
    fe_reduce512(r, R);
}
BPG_HD void fe_sqr(fe &r, const fe &a) {
    u32 R[16];
    sqr512(R, a.v);
    fe_reduce512(r, R);
}
// four independent field multiplications, interleaved (latency-bound kernels only: ~200 live registers)
BPG_HD void fe_mul4(fe &r0, const fe &a0, const fe &b0, fe &r1, const fe &a1, const fe &b1, fe &r2, const fe &a2, const fe &b2, fe &r3,
                    const fe &a3, const fe &b3) {
    u32 R[4][16];
    const u32 *pa[4] = {a0.v, a1.v, a2.v, a3.v}, *pb[4] = {b0.v, b1.v, b2.v, b3.v};
    mul512_ilp<4>(R, pa, pb);
    fe_reduce512(r0, R[0]); fe_reduce512(r1, R[1]); fe_reduce512(r2, R[2]); fe_reduce512(r3, R[3]);
}
BPG_HD void fe_mul3(fe &r0, const fe &a0, const fe &b0, fe &r1, const fe &a1, const fe &b1, fe &r2, const fe &a2, const fe &b2) {
    u32 R[3][16];
    const u32 *pa[3] = {a0.v, a1.v, a2.v}, *pb[3] = {b0.v, b1.v, b2.v};
    mul512_ilp<3>(R, pa, pb);
    fe_reduce512(r0, R[0]); fe_reduce512(r1, R[1]); fe_reduce512(r2, R[2]);
}
BPG_HD void fe_add(fe &r, const fe &a, const fe &b) {
    u32 c = add8(r.v, a.v, b.v);
    c = addsmall8(r.v, 38u & (0u - c));
    r.v[0] += 38u & (0u - c);
}
BPG_HD void fe_sub(fe &r, const fe &a, const fe &b) {
    u32 br = sub8(r.v, a.v, b.v);
    br = subsmall8(r.v, 38u & (0u - br));
    r.v[0] -= 38u & (0u - br);
}
BPG_HD void fe_set0(fe &r) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
}
BPG_HD void fe_set1(fe &r) { fe_set0(r); r.v[0] = 1; }
BPG_HD void fe_neg(fe &r, const fe &a) { fe z; fe_set0(z); fe_sub(r, z, a); }
BPG_HD void fe_dbl(fe &r, const fe &a) { fe_add(r, a, a); }
// canonical representative in [0, p)
BPG_HD void fe_canon(fe &r, const fe &a) {
    r = a;
    u32 t = r.v[7] >> 31;
    r.v[7] &= 0x7FFFFFFFu;
    (void)addsmall8(r.v, 19u * t); // < 2^255 + 19
    // if r >= p  <=>  r + 19 >= 2^255
    fe q = r;
    (void)addsmall8(q.v, 19u);
    u32 ge = q.v[7] >> 31;
    q.v[7] &= 0x7FFFFFFFu;
    u32 m = 0u - ge;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (q.v[i] & m) | (r.v[i] & ~m);
}
BPG_HD void fe_tobytes(u8 *s, const fe &a) {
    fe c;
    fe_canon(c, a);
    for (int i = 0; i < 8; i++) { s[4 * i] = (u8)c.v[i]; s[4 * i + 1] = (u8)(c.v[i] >> 8); s[4 * i + 2] = (u8)(c.v[i] >> 16); s[4 * i + 3] = (u8)(c.v[i] >> 24); }
}
BPG_HD void fe_frombytes(fe &r, const u8 *s) { // bit 255 dropped
    for (int i = 0; i < 8; i++) r.v[i] = (u32)s[4 * i] | ((u32)s[4 * i + 1] << 8) | ((u32)s[4 * i + 2] << 16) | ((u32)s[4 * i + 3] << 24);
    r.v[7] &= 0x7FFFFFFFu;
}
BPG_HD int fe_iszero(const fe &a) {
    fe c; fe_canon(c, a);
    u32 x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x |= c.v[i];
    return x == 0;
}
BPG_HD int fe_isneg(const fe &a) { fe c; fe_canon(c, a); return c.v[0] & 1; }
BPG_HD int fe_eq(const fe &a, const fe &b) { fe d; fe_sub(d, a, b); return fe_iszero(d); }
BPG_HD void fe_cmov(fe &r, const fe &a, int cond) {
    u32 m = 0u - (u32)(cond != 0);
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (a.v[i] & m) | (r.v[i] & ~m);
}
BPG_HD void fe_abs(fe &r, const fe &a) { fe n; fe_neg(n, a); int neg = fe_isneg(a); r = a; fe_cmov(r, n, neg); }
BPG_HD void fe_sqn(fe &r, const fe &a, int n) {
    r = a;
#pragma unroll 1
    for (int i = 0; i < n; i++) fe_sqr(r, r);
}
// z^(2^250-1), z^11
BPG_HD void fe_pow_2_250_1(fe &out, fe &z11, const fe &z) {
    fe z2, z9, t, z5, z10, z20, z50, z100;
    fe_sqr(z2, z); fe_sqn(t, z2, 2); fe_mul(z9, t, z); fe_mul(z11, z9, z2);
    fe_sqr(t, z11); fe_mul(z5, t, z9);
    fe_sqn(t, z5, 5); fe_mul(z10, t, z5);
    fe_sqn(t, z10, 10); fe_mul(z20, t, z10);
    fe_sqn(t, z20, 20); fe_mul(t, t, z20);
    fe_sqn(t, t, 10); fe_mul(z50, t, z10);
    fe_sqn(t, z50, 50); fe_mul(z100, t, z50);
    fe_sqn(t, z100, 100); fe_mul(t, t, z100);
    fe_sqn(t, t, 50); fe_mul(out, t, z50);
}
BPG_HD void fe_invert(fe &out, const fe &z) { fe t, z11; fe_pow_2_250_1(t, z11, z); fe_sqn(t, t, 5); fe_mul(out, t, z11); }
BPG_HD void fe_pow22523(fe &out, const fe &z) { fe t, z11; fe_pow_2_250_1(t, z11, z); fe_sqn(t, t, 2); fe_mul(out, t, z); }

// ---------------------------------------------------------------- scalars mod l = 2^252 + c
#define BPG_SC_C0 0x5cf5d3edu
#define BPG_SC_C1 0x5812631au
#define BPG_SC_C2 0xa2f79cd6u
#define BPG_SC_C3 0x14def9deu

// generic little-endian multi-limb multiply r[na+nb] = a[na]*b[nb]
template <int NA, int NB>
BPG_HD void mp_mul(u32 *r, const u32 *a, const u32 *b) {
#pragma unroll
    for (int i = 0; i < NA + NB; i++) r[i] = 0;
#pragma unroll
    for (int i = 0; i < NA; i++) {
        u32 c = 0;
#pragma unroll
        for (int j = 0; j < NB; j++) {
            u64 t = (u64)a[i] * b[j] + r[i + j] + c;
            r[i + j] = (u32)t;
            c = (u32)(t >> 32);
        }
        r[i + NB] = c;
    }
}
BPG_HD int sc_geq_l(const u32 *a) { // a[0..7] >= l ?
    const u32 Lw[8] = {BPG_SC_C0, BPG_SC_C1, BPG_SC_C2, BPG_SC_C3, 0, 0, 0, 0x10000000u};
    int ge = 1, decided = 0;
#pragma unroll
    for (int i = 7; i >= 0; i--) {
        int gt = a[i] > Lw[i], lt = a[i] < Lw[i];
        if (!decided && gt) { ge = 1; decided = 1; }
        if (!decided && lt) { ge = 0; decided = 1; }
    }
    return ge;
}
// reduce x[0..15] (512-bit) mod l.  x = lo - c*hi with hi = x >> 252, applied three times (see oracle/bpo.c)
BPG_HD void sc_reduce512(sc &out, const u32 *x) {
    const u32 Cw[4] = {BPG_SC_C0, BPG_SC_C1, BPG_SC_C2, BPG_SC_C3};
    u32 lo[8], hi[9];
#pragma unroll
    for (int i = 0; i < 8; i++) lo[i] = x[i];
    lo[7] &= 0x0FFFFFFFu;
#pragma unroll
    for (int i = 0; i < 9; i++) hi[i] = (x[i + 7] >> 28) | ((i + 8 < 16 ? x[i + 8] : 0u) << 4); // <= 260 bits
    u32 y[13];
    mp_mul<9, 4>(y, hi, Cw); // <= 385 bits
    u32 ylo[8], yhi[5];
#pragma unroll
    for (int i = 0; i < 8; i++) ylo[i] = y[i];
    ylo[7] &= 0x0FFFFFFFu;
#pragma unroll
    for (int i = 0; i < 5; i++) yhi[i] = (y[i + 7] >> 28) | ((i + 8 < 13 ? y[i + 8] : 0u) << 4); // <= 133 bits
    u32 z[9];
    mp_mul<5, 4>(z, yhi, Cw); // <= 258 bits
    u32 zlo[8];
#pragma unroll
    for (int i = 0; i < 8; i++) zlo[i] = z[i];
    zlo[7] &= 0x0FFFFFFFu;
    u32 zhi = (z[7] >> 28) | (z[8] << 4); // <= 6 bits
    u32 w[5];
    mp_mul<1, 4>(w, &zhi, Cw); // <= 131 bits
    // acc = 2l + lo + zlo - ylo - w  (9 words, always positive)
    const u32 twoL[9] = {0xb9eba7dau, 0xb024c634u, 0x45ef39acu, 0x29bdf3bdu, 0, 0, 0, 0x20000000u, 0};
    u32 acc[9];
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) { s += (u64)twoL[i] + (i < 8 ? lo[i] : 0u) + (i < 8 ? zlo[i] : 0u); acc[i] = (u32)s; s >>= 32; }
    u32 br = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        u64 d = (u64)acc[i] - (i < 8 ? ylo[i] : 0u) - br;
        acc[i] = (u32)d; br = (u32)((d >> 32) & 1);
    }
    br = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        u64 d = (u64)acc[i] - (i < 5 ? w[i] : 0u) - br;
        acc[i] = (u32)d; br = (u32)((d >> 32) & 1);
    }
    const u32 Lw[9] = {BPG_SC_C0, BPG_SC_C1, BPG_SC_C2, BPG_SC_C3, 0, 0, 0, 0x10000000u, 0};
#pragma unroll 1
    for (int k = 0; k < 5; k++) {
        int ge = acc[8] != 0 || sc_geq_l(acc);
        u32 m = 0u - (u32)ge;
        u32 b2 = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) {
            u64 d = (u64)acc[i] - (Lw[i] & m) - b2;
            acc[i] = (u32)d; b2 = (u32)((d >> 32) & 1);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) out.v[i] = acc[i];
}
BPG_HD void sc_mul(sc &r, const sc &a, const sc &b) {
    u32 R[16];
    mul512(R, a.v, b.v);
    sc_reduce512(r, R);
}
BPG_HD void sc_reduce(sc &r, const sc &a) {
    u32 R[16];
#pragma unroll
    for (int i = 0; i < 16; i++) R[i] = i < 8 ? a.v[i] : 0u;
    sc_reduce512(r, R);
}
// inputs must be reduced (< l); output reduced
BPG_HD void sc_add_r(sc &r, const sc &a, const sc &b) {
    u32 t[8];
    (void)add8(t, a.v, b.v); // < 2l < 2^254, no carry
    const u32 Lw[8] = {BPG_SC_C0, BPG_SC_C1, BPG_SC_C2, BPG_SC_C3, 0, 0, 0, 0x10000000u};
    u32 d[8];
    u32 br = sub8(d, t, Lw);
    u32 m = 0u - br; // borrow -> keep t
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (t[i] & m) | (d[i] & ~m);
}
BPG_HD void sc_sub_r(sc &r, const sc &a, const sc &b) {
    u32 d[8];
    u32 br = sub8(d, a.v, b.v);
    const u32 Lw[8] = {BPG_SC_C0, BPG_SC_C1, BPG_SC_C2, BPG_SC_C3, 0, 0, 0, 0x10000000u};
    u32 t[8];
    (void)add8(t, d, Lw);
    u32 m = 0u - br;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (t[i] & m) | (d[i] & ~m);
}
BPG_HD void sc_neg_r(sc &r, const sc &a) { sc z; for (int i = 0; i < 8; i++) z.v[i] = 0; sc_sub_r(r, z, a); }
BPG_HD void sc_set_u32(sc &r, u32 x) { for (int i = 0; i < 8; i++) r.v[i] = 0; r.v[0] = x; }
BPG_HD int sc_iszero(const sc &a) { u32 x = 0; for (int i = 0; i < 8; i++) x |= a.v[i]; return x == 0; }
BPG_HD void sc_frombytes(sc &r, const u8 *s) {
    for (int i = 0; i < 8; i++) r.v[i] = (u32)s[4 * i] | ((u32)s[4 * i + 1] << 8) | ((u32)s[4 * i + 2] << 16) | ((u32)s[4 * i + 3] << 24);
}
BPG_HD void sc_tobytes(u8 *s, const sc &a) {
    for (int i = 0; i < 8; i++) { s[4 * i] = (u8)a.v[i]; s[4 * i + 1] = (u8)(a.v[i] >> 8); s[4 * i + 2] = (u8)(a.v[i] >> 16); s[4 * i + 3] = (u8)(a.v[i] >> 24); }
}
// a^(l-2)
BPG_HD void sc_invert(sc &r, const sc &a) {
    const u32 e[8] = {BPG_SC_C0 - 2u, BPG_SC_C1, BPG_SC_C2, BPG_SC_C3, 0, 0, 0, 0x10000000u};
    sc base, acc;
    sc_reduce(base, a);
    sc_set_u32(acc, 1);
#pragma unroll 1
    for (int i = 252; i >= 0; i--) {
        sc_mul(acc, acc, acc);
        if ((e[i >> 5] >> (i & 31)) & 1) sc_mul(acc, acc, base);
    }
    r = acc;
}
// a^e for a small exponent
BPG_HD void sc_pow_u32(sc &r, const sc &a, u32 e) {
    sc acc, base = a;
    sc_set_u32(acc, 1);
#pragma unroll 1
    while (e) {
        if (e & 1) sc_mul(acc, acc, base);
        sc_mul(base, base, base);
        e >>= 1;
    }
    r = acc;
}
