// host_merlin.h -- host-side Keccak-f[1600], SHAKE256 / SHA3-512 and the STROBE-128 / Merlin transcript.
//
// Replaces merlin 1.x `Transcript` / `TranscriptRng` (reference dependency, /root/reference/Cargo.toml:10;
// used at src/bin/prover.rs:52, src/bin/verifier.rs:51) and the SHAKE256 "GeneratorsChain" of the bulletproofs
// fork.  Fiat-Shamir is inherently sequential, so it stays on the host between device steps.
// Written from FIPS 202 and the STROBE v1.0.2 / Merlin specifications.
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

namespace bpgh {

static inline uint64_t rol64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

inline void keccak_f1600(uint64_t a[25]) {
    static const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL,
        0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL,
        0x0000000080008009ULL, 0x000000008000000AULL, 0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL,
        0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    // rho offsets r[x][y] indexed as lane x + 5y
    static const int RHO[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    for (int rnd = 0; rnd < 24; rnd++) {
        uint64_t c[5], d[5], b[25];
        for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
        for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
        // rho + pi: B[y][2x+3y] = rot(A[x][y])
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) {
                int src = x + 5 * y, dst = y + 5 * ((2 * x + 3 * y) % 5);
                b[dst] = RHO[src] ? rol64(a[src], RHO[src]) : a[src];
            }
        for (int y = 0; y < 5; y++)
            for (int x = 0; x < 5; x++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        a[0] ^= RC[rnd];
    }
}

// generic sponge over a little-endian host
struct Sponge {
    uint64_t st[25];
    size_t pos, rate;
    explicit Sponge(size_t r) : pos(0), rate(r) { memset(st, 0, sizeof st); }
    void absorb(const uint8_t *d, size_t n) {
        uint8_t *b = (uint8_t *)st;
        for (size_t i = 0; i < n; i++) { b[pos++] ^= d[i]; if (pos == rate) { keccak_f1600(st); pos = 0; } }
    }
    void finish(uint8_t pad) { uint8_t *b = (uint8_t *)st; b[pos] ^= pad; b[rate - 1] ^= 0x80; keccak_f1600(st); pos = 0; }
    void squeeze(uint8_t *out, size_t n) {
        uint8_t *b = (uint8_t *)st;
        for (size_t i = 0; i < n; i++) { if (pos == rate) { keccak_f1600(st); pos = 0; } out[i] = b[pos++]; }
    }
};
inline void sha3_512(uint8_t out[64], const uint8_t *d, size_t n) { Sponge s(72); s.absorb(d, n); s.finish(0x06); s.squeeze(out, 64); }
// bulletproofs GeneratorsChain::new(label): SHAKE256("GeneratorsChain" || label), 64 bytes per generator
inline void generators_chain_stream(char which, uint32_t party, uint8_t *out, size_t npoints) {
    Sponge s(136);
    uint8_t lab[20];
    memcpy(lab, "GeneratorsChain", 15);
    lab[15] = (uint8_t)which;
    lab[16] = (uint8_t)party; lab[17] = (uint8_t)(party >> 8); lab[18] = (uint8_t)(party >> 16); lab[19] = (uint8_t)(party >> 24);
    s.absorb(lab, 20);
    s.finish(0x1F);
    s.squeeze(out, 64 * npoints);
}

// ---------------------------------------------------------------- STROBE-128 as used by Merlin
struct Strobe {
    enum { R = 166, F_I = 1, F_A = 2, F_C = 4, F_T = 8, F_M = 16, F_K = 32 };
    uint64_t st[25];
    uint8_t pos, pos_begin, cur_flags;
    uint8_t *bytes() { return (uint8_t *)st; }
    void init(const uint8_t *label, size_t n) {
        memset(st, 0, sizeof st);
        const uint8_t hdr[6] = {1, R + 2, 1, 0, 1, 96};
        memcpy(bytes(), hdr, 6);
        memcpy(bytes() + 6, "STROBEv1.0.2", 12);
        keccak_f1600(st);
        pos = pos_begin = cur_flags = 0;
        meta_ad(label, n, false);
    }
    void run_f() { uint8_t *b = bytes(); b[pos] ^= pos_begin; b[pos + 1] ^= 0x04; b[R + 1] ^= 0x80; keccak_f1600(st); pos = 0; pos_begin = 0; }
    void absorb(const uint8_t *d, size_t n) { uint8_t *b = bytes(); for (size_t i = 0; i < n; i++) { b[pos++] ^= d[i]; if (pos == R) run_f(); } }
    void overwrite(const uint8_t *d, size_t n) { uint8_t *b = bytes(); for (size_t i = 0; i < n; i++) { b[pos++] = d[i]; if (pos == R) run_f(); } }
    void squeeze(uint8_t *d, size_t n) { uint8_t *b = bytes(); for (size_t i = 0; i < n; i++) { d[i] = b[pos]; b[pos++] = 0; if (pos == R) run_f(); } }
    void begin_op(uint8_t flags, bool more) {
        if (more) return;
        uint8_t old = pos_begin;
        pos_begin = pos + 1;
        cur_flags = flags;
        uint8_t d[2] = {old, flags};
        absorb(d, 2);
        if ((flags & (F_C | F_K)) && pos != 0) run_f();
    }
    void meta_ad(const uint8_t *d, size_t n, bool more) { begin_op(F_M | F_A, more); absorb(d, n); }
    void ad(const uint8_t *d, size_t n, bool more) { begin_op(F_A, more); absorb(d, n); }
    void prf(uint8_t *d, size_t n, bool more) { begin_op(F_I | F_A | F_C, more); squeeze(d, n); }
    void key(const uint8_t *d, size_t n, bool more) { begin_op(F_A | F_C, more); overwrite(d, n); }
};
static inline void le32(uint8_t o[4], size_t n) { o[0] = (uint8_t)n; o[1] = (uint8_t)(n >> 8); o[2] = (uint8_t)(n >> 16); o[3] = (uint8_t)(n >> 24); }

struct Transcript {
    Strobe s;
    Transcript() {}
    Transcript(const uint8_t *label, size_t n) { s.init((const uint8_t *)"Merlin v1.0", 11); append("dom-sep", label, n); }
    void append(const char *label, const uint8_t *msg, size_t n) { append_raw((const uint8_t *)label, strlen(label), msg, n); }
    void append_raw(const uint8_t *label, size_t ll, const uint8_t *msg, size_t n) {
        uint8_t l4[4]; le32(l4, n);
        s.meta_ad(label, ll, false); s.meta_ad(l4, 4, true); s.ad(msg, n, false);
    }
    void append_u64(const char *label, uint64_t x) { uint8_t b[8]; for (int i = 0; i < 8; i++) b[i] = (uint8_t)(x >> (8 * i)); append(label, b, 8); }
    void challenge(const char *label, uint8_t *out, size_t n) { challenge_raw((const uint8_t *)label, strlen(label), out, n); }
    void challenge_raw(const uint8_t *label, size_t ll, uint8_t *out, size_t n) {
        uint8_t l4[4]; le32(l4, n);
        s.meta_ad(label, ll, false); s.meta_ad(l4, 4, true); s.prf(out, n, false);
    }
    // bulletproofs TranscriptProtocol::validate_and_append_point: the identity encoding is an error
    bool validate_and_append_point(const char *label, const uint8_t p[32]) {
        uint8_t z = 0; for (int i = 0; i < 32; i++) z |= p[i];
        if (!z) return false;
        append(label, p, 32);
        return true;
    }
};
// merlin TranscriptRngBuilder / TranscriptRng
struct TranscriptRng {
    Strobe s;
    explicit TranscriptRng(const Transcript &t) : s(t.s) {}
    void rekey_with_witness_bytes(const char *label, const uint8_t *w, size_t n) {
        uint8_t l4[4]; le32(l4, n);
        s.meta_ad((const uint8_t *)label, strlen(label), false); s.meta_ad(l4, 4, true); s.key(w, n, false);
    }
    void finalize(const uint8_t ext32[32]) { s.meta_ad((const uint8_t *)"rng", 3, false); s.key(ext32, 32, false); }
    void fill_bytes(uint8_t *out, size_t n) { uint8_t l4[4]; le32(l4, n); s.meta_ad(l4, 4, false); s.prf(out, n, false); }
};

} // namespace bpgh
