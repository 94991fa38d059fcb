// kernels_vec.cuh -- scalar-field (mod l) kernels: batched MiMC (K8) and the O(n) vector phases of the R1CS
// prover / verifier and of the inner-product argument (K9).
#pragma once
#include "kernels_core.cuh"

// ---------------------------------------------------------------- MiMC (src/mimc_hash/mimc.rs:7-40)
#define MIMC_ROUNDS 486
__constant__ sc c_mimc[MIMC_ROUNDS];

// One thread per independent sponge.  state += block; 486 x state = (state + c_i)^3  (zero key).
// trace (nullable): per absorbed block 972 multipliers x (a_L, a_R, a_O): (t,t,t^2) then (t^2,t,t^3)
// (mimc_hash_gadget.rs:133-144), block-major in absorption order of the whole batch.
__global__ void __launch_bounds__(128) k_mimc_sponge(const sc *__restrict__ blocks, const uint32_t *__restrict__ block_off, uint32_t n,
                                                      sc *__restrict__ out, sc *__restrict__ trace) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sc st;
    sc_set_u32(st, 0);
    uint32_t b0 = block_off[i], b1 = block_off[i + 1];
#pragma unroll 1
    for (uint32_t b = b0; b < b1; b++) {
        sc x;
        ld_sc(x, &blocks[b]);
        x.v[7] &= 0x7FFFFFFFu; // Scalar::from_bits
        sc_reduce(x, x);
        sc_add_r(st, st, x);
        sc *tr = trace ? trace + (size_t)b * (MIMC_ROUNDS * 6) : nullptr;
#pragma unroll 1
        for (int r = 0; r < MIMC_ROUNDS; r++) {
            sc t, t2, t3;
            sc_add_r(t, st, c_mimc[r]);
            sc_mul(t2, t, t);
            sc_mul(t3, t2, t);
            if (tr) {
                sc *o = tr + 6 * r;
                st_sc(o, t); st_sc(o + 1, t); st_sc(o + 2, t2);
                st_sc(o + 3, t2); st_sc(o + 4, t); st_sc(o + 5, t3);
            }
            st = t3;
        }
    }
    st_sc(&out[i], st);
}
