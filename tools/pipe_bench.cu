// tools/pipe_bench.cu -- raw issue-rate microbenchmarks for the integer / fp64 pipes on sm_100a.
// Used to establish the integer-multiply roofline (`imad_peak`) that MEASURED_PEAKS.json does not carry.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_bench tools/pipe_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__device__ __forceinline__ void body(uint64_t *out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t a0 = t, a1 = t + 1, a2 = t + 2, a3 = t + 3, a4 = t + 4, a5 = t + 5, a6 = t + 6, a7 = t + 7;
    uint32_t x = t * 2654435761u + 1, y = t * 40503u + 3;
    double d0 = t, d1 = t + 1, d2 = t + 2, d3 = t + 3, d4 = 1.0000001, d5 = 0.999999;
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
        if (MODE == 0) { // mad.wide.u32 (IMAD.WIDE.U32): 8 independent chains, multiplicand = low word of the accumulator
            asm volatile("{ .reg .u32 l0,l1,l2,l3,l4,l5,l6,l7,h;\n\t"
                         "mov.b64 {l0,h}, %0; mov.b64 {l1,h}, %1; mov.b64 {l2,h}, %2; mov.b64 {l3,h}, %3;\n\t"
                         "mov.b64 {l4,h}, %4; mov.b64 {l5,h}, %5; mov.b64 {l6,h}, %6; mov.b64 {l7,h}, %7;\n\t"
                         "mad.wide.u32 %0, l0, %8, %0; mad.wide.u32 %1, l1, %9, %1; mad.wide.u32 %2, l2, %8, %2; mad.wide.u32 %3, l3, %9, %3;\n\t"
                         "mad.wide.u32 %4, l4, %8, %4; mad.wide.u32 %5, l5, %9, %5; mad.wide.u32 %6, l6, %8, %6; mad.wide.u32 %7, l7, %9, %7; }"
                         : "+l"(a0), "+l"(a1), "+l"(a2), "+l"(a3), "+l"(a4), "+l"(a5), "+l"(a6), "+l"(a7) : "r"(x), "r"(y));
        } else if (MODE == 1) { // mad.lo.u32 (IMAD), 8 independent
            uint32_t *p = (uint32_t *)&a0; (void)p;
            asm volatile("mad.lo.u32 %0, %0, %8, %9; mad.lo.u32 %1, %1, %8, %9; mad.lo.u32 %2, %2, %8, %9; mad.lo.u32 %3, %3, %8, %9;"
                         "mad.lo.u32 %4, %4, %8, %9; mad.lo.u32 %5, %5, %8, %9; mad.lo.u32 %6, %6, %8, %9; mad.lo.u32 %7, %7, %8, %9;"
                         : "+r"(*(uint32_t *)&a0), "+r"(*(uint32_t *)&a1), "+r"(*(uint32_t *)&a2), "+r"(*(uint32_t *)&a3), "+r"(*(uint32_t *)&a4),
                           "+r"(*(uint32_t *)&a5), "+r"(*(uint32_t *)&a6), "+r"(*(uint32_t *)&a7) : "r"(x), "r"(y));
        } else if (MODE == 2) { // mad.hi.u32 (IMAD.HI)
            asm volatile("mad.hi.u32 %0, %0, %8, %9; mad.hi.u32 %1, %1, %8, %9; mad.hi.u32 %2, %2, %8, %9; mad.hi.u32 %3, %3, %8, %9;"
                         "mad.hi.u32 %4, %4, %8, %9; mad.hi.u32 %5, %5, %8, %9; mad.hi.u32 %6, %6, %8, %9; mad.hi.u32 %7, %7, %8, %9;"
                         : "+r"(*(uint32_t *)&a0), "+r"(*(uint32_t *)&a1), "+r"(*(uint32_t *)&a2), "+r"(*(uint32_t *)&a3), "+r"(*(uint32_t *)&a4),
                           "+r"(*(uint32_t *)&a5), "+r"(*(uint32_t *)&a6), "+r"(*(uint32_t *)&a7) : "r"(x), "r"(y));
        } else if (MODE == 3) { // fma.rn.f64 (DFMA), 4 independent x2
            asm volatile("fma.rn.f64 %0, %0, %4, %5; fma.rn.f64 %1, %1, %4, %5; fma.rn.f64 %2, %2, %4, %5; fma.rn.f64 %3, %3, %4, %5;"
                         "fma.rn.f64 %0, %0, %5, %4; fma.rn.f64 %1, %1, %5, %4; fma.rn.f64 %2, %2, %5, %4; fma.rn.f64 %3, %3, %5, %4;"
                         : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3) : "d"(d4), "d"(d5));
        } else if (MODE == 4) { // add.cc / addc chain (IADD3 with carry), 8 per iteration
            asm volatile("add.cc.u32 %0, %0, %8; addc.cc.u32 %1, %1, %9; addc.cc.u32 %2, %2, %8; addc.cc.u32 %3, %3, %9;"
                         "addc.cc.u32 %4, %4, %8; addc.cc.u32 %5, %5, %9; addc.cc.u32 %6, %6, %8; addc.u32 %7, %7, %9;"
                         : "+r"(*(uint32_t *)&a0), "+r"(*(uint32_t *)&a1), "+r"(*(uint32_t *)&a2), "+r"(*(uint32_t *)&a3), "+r"(*(uint32_t *)&a4),
                           "+r"(*(uint32_t *)&a5), "+r"(*(uint32_t *)&a6), "+r"(*(uint32_t *)&a7) : "r"(x), "r"(y));
        } else if (MODE == 5) { // fused pairs as in fe_mul: mad.lo.cc + madc.hi.cc (IMAD.WIDE.U32.X): two chains of 4 products
            uint32_t *w0 = (uint32_t *)&a0, *w1 = (uint32_t *)&a1, *w2 = (uint32_t *)&a2, *w3 = (uint32_t *)&a3;
            uint32_t *w4 = (uint32_t *)&a4, *w5 = (uint32_t *)&a5, *w6 = (uint32_t *)&a6, *w7 = (uint32_t *)&a7;
            asm volatile("mad.lo.cc.u32 %0, %8, %12, %0; madc.hi.cc.u32 %1, %8, %12, %1; madc.lo.cc.u32 %2, %9, %12, %2; madc.hi.cc.u32 %3, %9, %12, %3;"
                         "madc.lo.cc.u32 %4, %10, %12, %4; madc.hi.cc.u32 %5, %10, %12, %5; madc.lo.cc.u32 %6, %11, %12, %6; madc.hi.u32 %7, %11, %12, %7;"
                         : "+r"(w0[0]), "+r"(w1[0]), "+r"(w2[0]), "+r"(w3[0]), "+r"(w4[0]), "+r"(w5[0]), "+r"(w6[0]), "+r"(w7[0])
                         : "r"(w0[1]), "r"(w2[1]), "r"(w4[1]), "r"(w6[1]), "r"(x));
            asm volatile("mad.lo.cc.u32 %0, %8, %12, %0; madc.hi.cc.u32 %1, %8, %12, %1; madc.lo.cc.u32 %2, %9, %12, %2; madc.hi.cc.u32 %3, %9, %12, %3;"
                         "madc.lo.cc.u32 %4, %10, %12, %4; madc.hi.cc.u32 %5, %10, %12, %5; madc.lo.cc.u32 %6, %11, %12, %6; madc.hi.u32 %7, %11, %12, %7;"
                         : "+r"(w0[1]), "+r"(w1[1]), "+r"(w2[1]), "+r"(w3[1]), "+r"(w4[1]), "+r"(w5[1]), "+r"(w6[1]), "+r"(w7[1])
                         : "r"(w1[0]), "r"(w3[0]), "r"(w5[0]), "r"(w7[0]), "r"(y));
        }
    }
    uint64_t r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7 ^ (uint64_t)(d0 + d1 + d2 + d3);
    if (r == 0x1234567812345678ULL) out[t] = r;
}
// MODE < 16: every warp runs the same loop.  MODE = 16 + 4 A + B (A, B in {0: IMAD.WIDE, 1: IMAD, 3: DFMA, 2 -> 4: IADD3 carry}):
// warp-specialised mix -- even warps run loop A, odd warps loop B, nothing shared between them but the SM.  If the two pipes
// issue independently the mix takes max(T_A, T_B) / 2, if they share an issue port (T_A + T_B) / 2.
template <int MODE>
__global__ void __launch_bounds__(256) k(uint64_t *out) {
    if (MODE < 16) body<MODE>(out);
    else {
        constexpr int A = ((MODE - 16) >> 2) == 2 ? 4 : ((MODE - 16) >> 2), B = ((MODE - 16) & 3) == 2 ? 4 : ((MODE - 16) & 3);
        if ((threadIdx.x >> 5) & 1) body<B>(out); else body<A>(out);
    }
}
template <int MODE>
static void run(const char *name, double ops_per_iter, int sms, int clk_khz) {
    uint64_t *d;
    cudaMalloc(&d, (size_t)sms * 8 * 256 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms * 8, 256>>>(d);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k<MODE><<<sms * 8, 256>>>(d);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double ops = (double)sms * 8 * 256 * ITERS * ops_per_iter;
    double rate = ops / (best * 1e-3);
    printf("{\"pipe\": \"%s\", \"ms\": %.4f, \"ops_per_s\": %.4e, \"ops_per_clk_per_sm_at_max_clock\": %.2f}\n", name, best, rate, rate / sms / (clk_khz * 1e3));
    cudaFree(d);
}
int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"device\": \"%s\", \"sms\": %d, \"max_clock_khz\": %d}\n", p.name, p.multiProcessorCount, clk);
    run<0>("imad_wide_u32", 8, p.multiProcessorCount, clk);
    run<5>("imad_wide_u32_x_carry_chain", 8, p.multiProcessorCount, clk);
    run<1>("imad_lo_u32", 8, p.multiProcessorCount, clk);
    run<2>("imad_hi_u32", 8, p.multiProcessorCount, clk);
    run<3>("dfma_f64", 8, p.multiProcessorCount, clk);
    run<4>("iadd3_carry", 8, p.multiProcessorCount, clk);
    run<16 + 4 * 0 + 3>("mix_even_imad_wide_odd_dfma", 8, p.multiProcessorCount, clk);
    run<16 + 4 * 0 + 2>("mix_even_imad_wide_odd_iadd3", 8, p.multiProcessorCount, clk);
    run<16 + 4 * 3 + 2>("mix_even_dfma_odd_iadd3", 8, p.multiProcessorCount, clk);
    run<16 + 4 * 0 + 1>("mix_even_imad_wide_odd_imad_lo", 8, p.multiProcessorCount, clk);
    run<16 + 4 * 3 + 1>("mix_even_dfma_odd_imad_lo", 8, p.multiProcessorCount, clk);
    return 0;
}
