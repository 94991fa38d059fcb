// tools/dfma_bench.cu -- A/B prototype: GF(2^255-19) multiplication on the FP64 pipe (DFMA) vs the IMAD.WIDE carry chains
// of csrc/fe25519.cuh (VERDICT r01 "next" item 3: fmaheavy 70 % busy, fp64 pipe idle in k_msm_accumulate).
//
// Representation: 5 limbs of radix 2^51 held as integer-valued doubles (0 <= limb < 2^52).  A 51x51-bit limb product is
// split exactly with two round-toward-zero FMAs and one subtraction:
//     h = fma_rz(a, b, 2^104)            -> 2^104 + floor(a b / 2^52) 2^52      (ulp of [2^104, 2^105) is 2^52)
//     l = fma_rz(a, b, (2^104 + 2^52) - h) -> 2^52 + (a b mod 2^52)              (exact)
// so the mantissa fields of h and l are the two 52-bit halves of the product, and column sums are taken on the raw bit
// patterns with 64-bit integer additions (the accumulated exponent fields are a known constant per column).
// Prints: mismatches of the DFMA multiply against the integer one on random inputs, and the throughput of dependent
// multiplication chains at full occupancy for both (same harness as bpg_bench_imad).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I bulletproofs_gadgets_b200/csrc -o tools/dfma_bench tools/dfma_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "fe25519.cuh"

struct fd { double v[5]; };

#define C104 20282409603651670423947251286016.0 /* 2^104 */
#define C104_52 20282409603651674927546878656512.0 /* 2^104 + 2^52 */
#define C52 4503599627370496.0                   /* 2^52 */
#define BITS_104 0x4670000000000000ULL
#define BITS_52 0x4330000000000000ULL
#define MASK51 0x7FFFFFFFFFFFFULL

__device__ __forceinline__ void fd_from_fe(fd &r, const fe &a) { // canonical 8x32 -> 5x51
    uint64_t w[4];
    fe c;
    fe_canon(c, a);
    for (int i = 0; i < 4; i++) w[i] = (uint64_t)c.v[2 * i] | ((uint64_t)c.v[2 * i + 1] << 32);
    uint64_t l0 = w[0] & MASK51;
    uint64_t l1 = ((w[0] >> 51) | (w[1] << 13)) & MASK51;
    uint64_t l2 = ((w[1] >> 38) | (w[2] << 26)) & MASK51;
    uint64_t l3 = ((w[2] >> 25) | (w[3] << 39)) & MASK51;
    uint64_t l4 = (w[3] >> 12) & MASK51;
    r.v[0] = (double)l0; r.v[1] = (double)l1; r.v[2] = (double)l2; r.v[3] = (double)l3; r.v[4] = (double)l4;
}
__device__ __forceinline__ void fe_from_fd(fe &r, const fd &a) {
    // limbs may be up to 2^52: accumulate into 4 x 64 with carries, then reduce through the integer code
    unsigned __int128 acc = 0;
    uint64_t l[5];
    for (int i = 0; i < 5; i++) l[i] = (uint64_t)a.v[i];
    // value = sum l_i 2^(51 i) < 2^257: fold the part above 2^255 with 19
    uint64_t w[5] = {0, 0, 0, 0, 0};
    acc = (unsigned __int128)l[0] + ((unsigned __int128)l[1] << 51);
    w[0] = (uint64_t)acc; acc >>= 64;
    acc += ((unsigned __int128)l[2] << 38);
    w[1] = (uint64_t)acc; acc >>= 64;
    acc += ((unsigned __int128)l[3] << 25);
    w[2] = (uint64_t)acc; acc >>= 64;
    acc += ((unsigned __int128)l[4] << 12);
    w[3] = (uint64_t)acc; acc >>= 64;
    w[4] = (uint64_t)acc; // < 2^2
    // r = w[0..3] + 38 * w[4]  (2^256 = 38)
    fe t;
    for (int i = 0; i < 4; i++) { t.v[2 * i] = (uint32_t)w[i]; t.v[2 * i + 1] = (uint32_t)(w[i] >> 32); }
    fe add; fe_set0(add); add.v[0] = (uint32_t)(38 * w[4]);
    fe_add(r, t, add);
}

// ---- variant A: 25 independent (h, l) splits, integer column sums
__device__ __forceinline__ void fd_mul(fd &r, const fd &a, const fd &b) {
    uint64_t LO[9], HI[9];
#pragma unroll
    for (int k = 0; k < 9; k++) { LO[k] = 0; HI[k] = 0; }
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = 0; j < 5; j++) {
            double h = __fma_rz(a.v[i], b.v[j], C104);
            double l = __fma_rz(a.v[i], b.v[j], C104_52 - h);
            HI[i + j] += (uint64_t)__double_as_longlong(h);
            LO[i + j] += (uint64_t)__double_as_longlong(l);
        }
    // strip the exponent fields: column k holds cnt_k = min(k, 8 - k) + 1 products
    uint64_t T[10];
#pragma unroll
    for (int k = 0; k < 9; k++) {
        uint64_t cnt = (uint64_t)((k < 4 ? k : 8 - k) + 1);
        LO[k] -= cnt * BITS_52;
        HI[k] -= cnt * BITS_104;
    }
    // product = sum_k (LO_k + 2^52 HI_k) 2^(51 k)  =>  T_k = LO_k + 2 HI_{k-1}
    T[0] = LO[0];
#pragma unroll
    for (int k = 1; k < 9; k++) T[k] = LO[k] + 2 * HI[k - 1];
    T[9] = 2 * HI[8];
    // fold 2^255 = 19 ; T_k < 15 * 2^52, so 19 T_k < 2^61
    uint64_t U[5];
#pragma unroll
    for (int k = 0; k < 5; k++) U[k] = T[k] + 19 * T[k + 5];
    // carry chain to 51-bit limbs
    uint64_t c;
    c = U[0] >> 51; U[0] &= MASK51; U[1] += c;
    c = U[1] >> 51; U[1] &= MASK51; U[2] += c;
    c = U[2] >> 51; U[2] &= MASK51; U[3] += c;
    c = U[3] >> 51; U[3] &= MASK51; U[4] += c;
    c = U[4] >> 51; U[4] &= MASK51; U[0] += 19 * c; // c < 2^11: U[0] < 2^51 + 2^16
#pragma unroll
    for (int k = 0; k < 5; k++) r.v[k] = __longlong_as_double((long long)(U[k] | BITS_52)) - C52;
}

// ---- variant B: the high halves of a column are accumulated by the FMA itself (a chain of <= 4 products stays below
// 2^104 for limbs < 2^51), which removes 25 integer additions; the chain's running value replaces 2^104 as the addend.
__device__ __forceinline__ void fd_mul_chain(fd &r, const fd &a, const fd &b) {
    uint64_t LO[9], HI[9];
#pragma unroll
    for (int k = 0; k < 9; k++) {
        const int cnt = (k < 4 ? k : 8 - k) + 1;
        const int i0 = k < 4 ? 0 : k - 4;
        uint64_t lo = 0, hi = 0;
        double hprev = C104;
        int inchain = 0;
#pragma unroll
        for (int t = 0; t < cnt; t++) {
            const int i = i0 + t, j = k - i;
            if (inchain == 3) { hi += (uint64_t)__double_as_longlong(hprev) - BITS_104; hprev = C104; inchain = 0; } // 5-product column: 3 + 2
            double h = __fma_rz(a.v[i], b.v[j], hprev);
            double l = __fma_rz(a.v[i], b.v[j], (hprev + C52) - h);
            lo += (uint64_t)__double_as_longlong(l);
            hprev = h;
            inchain++;
        }
        hi += (uint64_t)__double_as_longlong(hprev) - BITS_104;
        LO[k] = lo - (uint64_t)cnt * BITS_52;
        HI[k] = hi;
    }
    uint64_t T[10];
    T[0] = LO[0];
#pragma unroll
    for (int k = 1; k < 9; k++) T[k] = LO[k] + 2 * HI[k - 1];
    T[9] = 2 * HI[8];
    uint64_t U[5];
#pragma unroll
    for (int k = 0; k < 5; k++) U[k] = T[k] + 19 * T[k + 5];
    uint64_t c;
    c = U[0] >> 51; U[0] &= MASK51; U[1] += c;
    c = U[1] >> 51; U[1] &= MASK51; U[2] += c;
    c = U[2] >> 51; U[2] &= MASK51; U[3] += c;
    c = U[3] >> 51; U[3] &= MASK51; U[4] += c;
    c = U[4] >> 51; U[4] &= MASK51; U[0] += 19 * c;
#pragma unroll
    for (int k = 0; k < 5; k++) r.v[k] = __longlong_as_double((long long)(U[k] | BITS_52)) - C52;
}

__device__ __forceinline__ uint32_t rnd32(uint32_t &s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

template <int VAR>
__global__ void k_check(uint32_t *mism, int rounds) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    fe a, b;
    for (int i = 0; i < 8; i++) { a.v[i] = rnd32(s); b.v[i] = rnd32(s); }
    if ((threadIdx.x & 7) == 0) for (int i = 0; i < 8; i++) a.v[i] = 0xFFFFFFFFu; // edge: all-ones limbs
    if ((threadIdx.x & 7) == 1) for (int i = 0; i < 8; i++) b.v[i] = (i == 7) ? 0x7FFFFFFFu : 0xFFFFFFECu; // p - 1
    fd da, db;
    fd_from_fe(da, a); fd_from_fe(db, b);
    uint32_t bad = 0;
    for (int r = 0; r < rounds; r++) {
        fe c; fd dc;
        fe_mul(c, a, b);
        if (VAR == 0) fd_mul(dc, da, db); else fd_mul_chain(dc, da, db);
        fe back;
        fe_from_fd(back, dc);
        if (!fe_eq(back, c)) bad++;
        a = b; b = c; da = db; db = dc; // Fibonacci-style chain: outputs feed the next multiplication un-normalised
    }
    if (bad) atomicAdd(mism, bad);
}

template <int VAR>
__global__ void __launch_bounds__(256) k_bench(uint32_t *out, int iters) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (VAR == 2) {
        fe a, b;
#pragma unroll
        for (int i = 0; i < 8; i++) { a.v[i] = t * 2654435761u + i; b.v[i] = t * 40503u + 77u * i + 1; }
#pragma unroll 1
        for (int it = 0; it < iters; it++) { fe_mul(a, a, b); fe_mul(b, b, a); }
        uint32_t x = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) x ^= a.v[i] ^ b.v[i];
        if (x == 0x12345678u) out[t] = x;
    } else if (VAR == 4) {
        // warp-specialised: even warps run the integer chain, odd warps the DFMA chain (nothing shared but the SM).  With
        // independent pipes this takes max(T_int, T_dfma) / 2, with a shared issue port (T_int + T_dfma) / 2.
        if ((threadIdx.x >> 5) & 1) {
            fd a, b;
#pragma unroll
            for (int i = 0; i < 5; i++) { a.v[i] = (double)((t * 2654435761u + i) & 0xFFFFFu) * 1048576.0 + 3.0; b.v[i] = (double)((t * 40503u + 77u * i + 1) & 0xFFFFFu) * 524288.0 + 5.0; }
#pragma unroll 1
            for (int it = 0; it < iters; it++) { fd_mul(a, a, b); fd_mul(b, b, a); }
            double x = 0;
#pragma unroll
            for (int i = 0; i < 5; i++) x += a.v[i] + b.v[i];
            if (x == 12345.678) out[t] = 1;
        } else {
            fe a, b;
#pragma unroll
            for (int i = 0; i < 8; i++) { a.v[i] = t * 2654435761u + i; b.v[i] = t * 40503u + 77u * i + 1; }
#pragma unroll 1
            for (int it = 0; it < iters; it++) { fe_mul(a, a, b); fe_mul(b, b, a); }
            uint32_t x = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) x ^= a.v[i] ^ b.v[i];
            if (x == 0x12345678u) out[t] = x;
        }
    } else if (VAR == 3) {
        // co-issue: one IMAD.WIDE chain and one DFMA chain per thread, independent of each other (upper bound of what a
        // mixed addition that splits its 7 multiplications over both pipes could reach; counted as 2 + 2 fe_mul per iteration)
        fe a, b;
        fd c, d;
#pragma unroll
        for (int i = 0; i < 8; i++) { a.v[i] = t * 2654435761u + i; b.v[i] = t * 40503u + 77u * i + 1; }
#pragma unroll
        for (int i = 0; i < 5; i++) { c.v[i] = (double)((t * 2654435761u + i) & 0xFFFFFu) * 1048576.0 + 3.0; d.v[i] = (double)((t * 40503u + 77u * i + 1) & 0xFFFFFu) * 524288.0 + 5.0; }
#pragma unroll 1
        for (int it = 0; it < iters; it++) { fe_mul(a, a, b); fd_mul(c, c, d); fe_mul(b, b, a); fd_mul(d, d, c); }
        uint32_t x = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) x ^= a.v[i] ^ b.v[i];
        double y = 0;
#pragma unroll
        for (int i = 0; i < 5; i++) y += c.v[i] + d.v[i];
        if (x == 0x12345678u && y == 12345.678) out[t] = x;
    } else {
        fd a, b;
#pragma unroll
        for (int i = 0; i < 5; i++) { a.v[i] = (double)((t * 2654435761u + i) & 0xFFFFFu) * 1048576.0 + 3.0; b.v[i] = (double)((t * 40503u + 77u * i + 1) & 0xFFFFFu) * 524288.0 + 5.0; }
#pragma unroll 1
        for (int it = 0; it < iters; it++) {
            if (VAR == 0) { fd_mul(a, a, b); fd_mul(b, b, a); } else { fd_mul_chain(a, a, b); fd_mul_chain(b, b, a); }
        }
        double x = 0;
#pragma unroll
        for (int i = 0; i < 5; i++) x += a.v[i] + b.v[i];
        if (x == 12345.678) out[t] = 1;
    }
}

static int g_reps = 5; // 1 under ncu (`dfma_bench ncu`): one launch per variant
template <int VAR>
static void bench(const char *name, int sms, int blocks_per_sm, int threads, int iters) {
    uint32_t *d;
    cudaMalloc(&d, (size_t)sms * blocks_per_sm * threads * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_bench<VAR><<<sms * blocks_per_sm, threads>>>(d, 8);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < g_reps; rep++) {
        cudaEventRecord(e0);
        k_bench<VAR><<<sms * blocks_per_sm, threads>>>(d, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double muls = (double)sms * blocks_per_sm * threads * iters * (VAR == 3 ? 4.0 : 2.0);
    printf("{\"variant\": \"%s\", \"blocks_per_sm\": %d, \"threads\": %d, \"ms\": %.4f, \"fe_mul_per_s\": %.4e, \"clk_per_sm_per_fe_mul_at_1965MHz\": %.3f}\n", name,
           blocks_per_sm, threads, best, muls / (best * 1e-3), 1.965e9 * sms / (muls / (best * 1e-3)));
    cudaFree(d);
}

int main(int argc, char **argv) {
    const bool ncu_mode = argc > 1;
    if (ncu_mode) g_reps = 1;
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d}\n", p.name, sms);
    uint32_t *d_m, h_m[2] = {0, 0};
    cudaMalloc(&d_m, 8);
    cudaMemset(d_m, 0, 8);
    k_check<0><<<64, 128>>>(d_m, 64);
    k_check<1><<<64, 128>>>(d_m + 1, 64);
    cudaMemcpy(h_m, d_m, 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    printf("{\"check\": \"dfma fe_mul vs integer fe_mul, 64 x 128 threads x 64 chained products\", \"mismatch_independent\": %u, \"mismatch_chained_hi\": %u, \"cuda\": \"%s\"}\n",
           h_m[0], h_m[1], cudaGetErrorString(e));
    for (int bps : {2, 4, 8}) {
        if (ncu_mode && bps != 4) continue;
        bench<2>("imad_wide_8x32", sms, bps, 256, 400);
        bench<0>("dfma_5x51_independent", sms, bps, 256, 400);
        bench<1>("dfma_5x51_chained_hi", sms, bps, 256, 400);
        bench<3>("co_issue_imad_chain_plus_dfma_chain", sms, bps, 256, 400);
        bench<4>("warp_specialised_even_imad_odd_dfma", sms, bps, 256, 400);
    }
    return 0;
}
