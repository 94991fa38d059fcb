// host_keccak_lanes.cpp -- W independent merlin TranscriptRng streams advanced in lock step, one SIMD lane each.
//
// merlin 1.x TranscriptRng::fill_bytes (reference dependency, /root/reference/Cargo.toml:10; reached from
// Prover::prove at /root/reference/src/bin/prover.rs:93) costs one Keccak-f[1600] per 64-byte draw and a proof draws
// 2n of them in sequence (the blinding vectors s_L, s_R).  One stream cannot be parallelised, but the streams of
// different proofs are independent: this file runs W of them side by side in the 64-bit lanes of AVX2 (W = 4) or
// AVX-512 (W = 8) registers.  It is compiled once per instruction set (-DBPG_LANES=4 -mavx2 / -DBPG_LANES=8 -mavx512f,
// see Makefile) and selected at run time by host_rng_service.h; the scalar path in host_merlin.h stays the definition
// of the byte stream and tests/test_abi_and_host.py compares the two.
//
// Steady state of a stream between two 64-byte draws (STROBE position 64, pos_begin 0; see host_merlin.h):
//   meta-AD(le32(64)) ; PRF(64)  ==  bytes 64..73 ^= {00 12 40 00 00 00 41 07 47 04}, byte 167 ^= 80, Keccak-f,
//   output = bytes 0..63, which are then zeroed.
#include <immintrin.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifndef BPG_LANES
#error "compile with -DBPG_LANES=4 (AVX2) or -DBPG_LANES=8 (AVX-512)"
#endif

#if BPG_LANES == 8
typedef __m512i vec;
#define V_LOAD(p) _mm512_loadu_si512((const void *)(p))
#define V_STORE(p, v) _mm512_storeu_si512((void *)(p), v)
#define V_XOR(a, b) _mm512_xor_si512(a, b)
#define V_XOR3(a, b, c) _mm512_ternarylogic_epi64(a, b, c, 0x96)
#define V_CHI(a, b, c) _mm512_ternarylogic_epi64(a, b, c, 0xD2) /* a ^ (~b & c) */
#define V_ROL(a, n) _mm512_rol_epi64(a, n)
#define V_SET1(x) _mm512_set1_epi64((long long)(x))
#define V_ZERO() _mm512_setzero_si512()
#define FN_NAME bpg_rng_lanes_avx512
#else
typedef __m256i vec;
#define V_LOAD(p) _mm256_loadu_si256((const __m256i *)(p))
#define V_STORE(p, v) _mm256_storeu_si256((__m256i *)(p), v)
#define V_XOR(a, b) _mm256_xor_si256(a, b)
#define V_XOR3(a, b, c) _mm256_xor_si256(a, _mm256_xor_si256(b, c))
#define V_CHI(a, b, c) _mm256_xor_si256(a, _mm256_andnot_si256(b, c))
#define V_ROL(a, n) _mm256_or_si256(_mm256_slli_epi64(a, n), _mm256_srli_epi64(a, 64 - (n)))
#define V_SET1(x) _mm256_set1_epi64x((long long)(x))
#define V_ZERO() _mm256_setzero_si256()
#define FN_NAME bpg_rng_lanes_avx2
#endif

static const uint64_t RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL, 0x0000000080000001ULL,
    0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
    0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
    0x000000000000800AULL, 0x800000008000000AULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

// st: 25 words x W lanes (word-major).  Performs `steps` steady-state draws on every lane; draw i of lane l is written to
// out[l] + i * stride[l] (64 bytes; idle lanes pass a 64-byte dummy with stride 0).
extern "C" void FN_NAME(uint64_t *st, uint8_t *const *out, const size_t *stride, size_t steps) {
    vec a[25];
    for (int i = 0; i < 25; i++) a[i] = V_LOAD(st + (size_t)i * BPG_LANES);
    const vec k8 = V_SET1(0x0741000000401200ULL), k9 = V_SET1(0x0447ULL), k20 = V_SET1(0x8000000000000000ULL);
    uint64_t tmp[8 * BPG_LANES];
    uint8_t *o[BPG_LANES];
    for (int l = 0; l < BPG_LANES; l++) o[l] = out[l];
    for (size_t s = 0; s < steps; s++) {
        a[8] = V_XOR(a[8], k8);
        a[9] = V_XOR(a[9], k9);
        a[20] = V_XOR(a[20], k20);
        for (int rnd = 0; rnd < 24; rnd++) {
            vec c0 = V_XOR3(a[0], a[5], V_XOR3(a[10], a[15], a[20]));
            vec c1 = V_XOR3(a[1], a[6], V_XOR3(a[11], a[16], a[21]));
            vec c2 = V_XOR3(a[2], a[7], V_XOR3(a[12], a[17], a[22]));
            vec c3 = V_XOR3(a[3], a[8], V_XOR3(a[13], a[18], a[23]));
            vec c4 = V_XOR3(a[4], a[9], V_XOR3(a[14], a[19], a[24]));
            vec d0 = V_XOR(c4, V_ROL(c1, 1));
            vec d1 = V_XOR(c0, V_ROL(c2, 1));
            vec d2 = V_XOR(c1, V_ROL(c3, 1));
            vec d3 = V_XOR(c2, V_ROL(c4, 1));
            vec d4 = V_XOR(c3, V_ROL(c0, 1));
            vec b0 = V_XOR(a[0], d0);
            vec b1 = V_ROL(V_XOR(a[6], d1), 44);
            vec b2 = V_ROL(V_XOR(a[12], d2), 43);
            vec b3 = V_ROL(V_XOR(a[18], d3), 21);
            vec b4 = V_ROL(V_XOR(a[24], d4), 14);
            vec b5 = V_ROL(V_XOR(a[3], d3), 28);
            vec b6 = V_ROL(V_XOR(a[9], d4), 20);
            vec b7 = V_ROL(V_XOR(a[10], d0), 3);
            vec b8 = V_ROL(V_XOR(a[16], d1), 45);
            vec b9 = V_ROL(V_XOR(a[22], d2), 61);
            vec b10 = V_ROL(V_XOR(a[1], d1), 1);
            vec b11 = V_ROL(V_XOR(a[7], d2), 6);
            vec b12 = V_ROL(V_XOR(a[13], d3), 25);
            vec b13 = V_ROL(V_XOR(a[19], d4), 8);
            vec b14 = V_ROL(V_XOR(a[20], d0), 18);
            vec b15 = V_ROL(V_XOR(a[4], d4), 27);
            vec b16 = V_ROL(V_XOR(a[5], d0), 36);
            vec b17 = V_ROL(V_XOR(a[11], d1), 10);
            vec b18 = V_ROL(V_XOR(a[17], d2), 15);
            vec b19 = V_ROL(V_XOR(a[23], d3), 56);
            vec b20 = V_ROL(V_XOR(a[2], d2), 62);
            vec b21 = V_ROL(V_XOR(a[8], d3), 55);
            vec b22 = V_ROL(V_XOR(a[14], d4), 39);
            vec b23 = V_ROL(V_XOR(a[15], d0), 41);
            vec b24 = V_ROL(V_XOR(a[21], d1), 2);
            a[0] = V_XOR(V_CHI(b0, b1, b2), V_SET1(RC[rnd]));
            a[1] = V_CHI(b1, b2, b3);
            a[2] = V_CHI(b2, b3, b4);
            a[3] = V_CHI(b3, b4, b0);
            a[4] = V_CHI(b4, b0, b1);
            a[5] = V_CHI(b5, b6, b7);
            a[6] = V_CHI(b6, b7, b8);
            a[7] = V_CHI(b7, b8, b9);
            a[8] = V_CHI(b8, b9, b5);
            a[9] = V_CHI(b9, b5, b6);
            a[10] = V_CHI(b10, b11, b12);
            a[11] = V_CHI(b11, b12, b13);
            a[12] = V_CHI(b12, b13, b14);
            a[13] = V_CHI(b13, b14, b10);
            a[14] = V_CHI(b14, b10, b11);
            a[15] = V_CHI(b15, b16, b17);
            a[16] = V_CHI(b16, b17, b18);
            a[17] = V_CHI(b17, b18, b19);
            a[18] = V_CHI(b18, b19, b15);
            a[19] = V_CHI(b19, b15, b16);
            a[20] = V_CHI(b20, b21, b22);
            a[21] = V_CHI(b21, b22, b23);
            a[22] = V_CHI(b22, b23, b24);
            a[23] = V_CHI(b23, b24, b20);
            a[24] = V_CHI(b24, b20, b21);
        }
        // squeeze 64 bytes per lane: words 0..7, transposed through a small buffer, then zeroed
        for (int w = 0; w < 8; w++) { V_STORE(tmp + (size_t)w * BPG_LANES, a[w]); a[w] = V_ZERO(); }
        for (int l = 0; l < BPG_LANES; l++) {
            uint64_t row[8];
            for (int w = 0; w < 8; w++) row[w] = tmp[(size_t)w * BPG_LANES + l];
            memcpy(o[l], row, 64);
            o[l] += stride[l];
        }
    }
    for (int i = 0; i < 25; i++) V_STORE(st + (size_t)i * BPG_LANES, a[i]);
}
