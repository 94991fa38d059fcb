// bpg_internal.h -- context, device buffers and kernel-launch prototypes shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/bpg.h"
#include "ge25519.cuh"

#define BPG_WBITS 16                  // signed window width of the resident tables
#define BPG_NWIN 16                   // floor(253/16)+1 windows
#define BPG_NBW (1u << (BPG_WBITS - 1)) // |digit| in 1..32768
#define BPG_NBP (BPG_NBW + 32u)       // buckets per group incl. unused bucket 0 and padding (multiple of 4)
#define BPG_MAX_GROUPS 4
#define BPG_MAX_SEGS 12
#define BPG_CHUNK 64                  // sorted pairs summed by one thread of the accumulate kernel
#define BPG_HEAVY_SPAN 48             // buckets spanning more chunks than this go to the block-wide tree kernel

#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { bpg_set_cuda_error(e_, __FILE__, __LINE__); return BPG_E_CUDA; } } while (0)
void bpg_set_cuda_error(cudaError_t e, const char *file, int line);

struct msm_seg {
    const sc *scalars; // device, 32 B each
    uint32_t n;        // terms
    uint32_t p0;       // unified point index of term 0 (G: [0,cap)  H: [cap,2cap)  B: 2cap  B~: 2cap+1)
    uint32_t group;    // output group
    uint32_t reduce;   // 1: scalars may be >= l (host-supplied, Scalar::from_bits semantics)
    uint32_t start;    // filled by msm_run: global term index of term 0
    uint32_t pad;
};
struct msm_plan { msm_seg seg[BPG_MAX_SEGS]; int nseg; int ngroups; uint32_t total; };

struct dev_buf { // grow-only device buffer
    void *p = nullptr; size_t cap = 0;
    int ensure(size_t bytes);
    void release();
};

struct bpg_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, stream2 = nullptr;
    cudaEvent_t ev = nullptr;
    // resident generators
    size_t cap = 0;           // per-chain capacity (power of two); tables cover 2*cap+2 points
    uint32_t ptotal = 0;
    ge_an *tab = nullptr;     // [BPG_NWIN][ptotal]
    ge_an *comb = nullptr;    // [2][32][128] signed 8-bit comb for B and B~ (Pedersen commits)
    // MSM workspace
    dev_buf counts, offsets, cursor, sorted, partial, buckets, lvlP, lvlQ, heavy, results;
    // generic scratch
    dev_buf scratch[16];
    void *h_pinned = nullptr; size_t h_pinned_cap = 0;
    uint64_t launches = 0;    // kernels launched (reported by bench.py as gpu_launches)
    std::string last_error;
};

// ---- kernels.cu
int k_upload_constants();
int k_derive_gens(bpg_ctx *c, const uint8_t *h_stream_G, const uint8_t *h_stream_H, size_t cap);
int k_compress(bpg_ctx *c, cudaStream_t s, const ge *d_pts, size_t n, uint8_t *d_out32);
int k_table_point_export(bpg_ctx *c, cudaStream_t s, uint32_t p0, size_t n, uint8_t *d_out32);
int k_pedersen_commit(bpg_ctx *c, cudaStream_t s, const sc *d_v, const sc *d_r, size_t n, uint8_t *d_out32);
int k_decompress(bpg_ctx *c, cudaStream_t s, const uint8_t *d_in32, size_t n, ge *d_out, uint32_t *d_ok);
int k_varbase_msm(bpg_ctx *c, cudaStream_t s, const sc *d_scalars, const ge *d_pts, size_t n, ge *d_out /*1*/);
int k_fold_points(bpg_ctx *c, cudaStream_t s, const sc *d_sl, const sc *d_sr, const ge *d_PL, const ge *d_PR, size_t n, ge *d_out);
int k_points_sum(bpg_ctx *c, cudaStream_t s, const ge *d_pts, size_t n, ge *d_out);
int k_mimc(bpg_ctx *c, cudaStream_t s, const sc *d_blocks, const uint32_t *d_block_off, size_t n_hashes, sc *d_out, sc *d_trace);
int k_mimc_set_constants(const uint8_t *consts486x32);

// ---- msm.cu
int msm_run(bpg_ctx *c, cudaStream_t s, msm_plan *plan, ge *d_out /* ngroups extended points */);

// ---- vec.cu : scalar-vector kernels of the R1CS prover / verifier / IPP
struct csc_dev { const uint32_t *col_ptr; const uint32_t *row; const sc *coeff; }; // per variable kind
int v_pow_tables(bpg_ctx *c, cudaStream_t s, const sc *d_base /*1*/, uint32_t max_exp, sc *d_lo /*1024*/, sc *d_hi);
int v_flatten(bpg_ctx *c, cudaStream_t s, csc_dev m, uint32_t ncols, const sc *d_zlo, const sc *d_zhi, int negate, sc *d_out);
int v_reduce_sum(bpg_ctx *c, cudaStream_t s, const sc *d_partials, uint32_t nparts, uint32_t nsums, sc *d_out);
