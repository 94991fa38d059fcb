// bpg_internal.h -- context, device buffers and kernel-launch prototypes shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <memory>
#include <string>
#include <vector>

#include "../../include/bpg.h"
#include "ge25519.cuh"

#define BPG_WBITS 16                  // signed window width of the resident tables
#define BPG_NWIN 16                   // floor(253/16)+1 windows
#define BPG_NBW (1u << (BPG_WBITS - 1)) // |digit| in 1..32768
#define BPG_NBP (129u * 256u)          // buckets per group: index = |digit| in [0, 32768], padded to 129 rows x 256 columns
#define BPG_MAX_GROUPS 4
#define BPG_MAX_SEGS 12
#define BPG_CHUNK 64                  // max sorted pairs summed by one thread of the accumulate kernel (8..64, sized per MSM)
#define BPG_SMALL_MSM_TERMS 4096      // MSMs up to this many terms use split 8-bit digits and 2 x 129 buckets per group (msm_run)
#define BPG_PRIV_MSM_TERMS (1u << 19) // single-group MSMs from this many terms use the shared-memory privatised histogram / scatter
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
    uint32_t alt;      // 0, or 1 + k: group ^= bit k of the term's index inside the segment (IPP rounds: left/right halves)
    uint32_t j0;       // index of term 0 inside the caller's full vector (non-zero when a rank holds a slice of the segment)
};
struct msm_plan {
    msm_seg seg[BPG_MAX_SEGS]; int nseg; int ngroups; uint32_t total;
    const ge_an *tab;   // window tables the point indices refer to (nullptr: the resident generators, ctx->tab)
    uint32_t ptotal;    // points per window of `tab`
    int shard;          // 1: with a sharded context (bpg_ctx_set_shard) every vector segment is cut by point range, this rank sums
                        //    its slice and the partial results are all-gathered and added (protocol drivers only)
    int varbase;        // 1: `tab` holds `ptotal` variable points in affine-Niels form (no window tables): bucket method per window +
                        //    Horner recombination (bpg_msm, the verifier's own points)
    int lean;           // 1: throughput sizing (long accumulate chunks, k_msm_rowcol_lean) -- set by the protocol drivers when
                        //    several proofs are in flight in this process (bpg_lean_now)
};

struct dev_buf { // grow-only device buffer
    void *p = nullptr; size_t cap = 0;
    int ensure(size_t bytes);
    void release();
};

// Window tables of the generators of one device: built once, shared by every context of the process on that device
// (BulletproofGens is a deterministic function of the capacity).  One 201 MB table set at 2^16 instead of one per prover
// thread also lets the 126 MB L2 hold most of what the accumulate kernel gathers.
struct gens_tables {
    int device = 0;
    size_t cap = 0;
    uint32_t ptotal = 0;
    ge_an *tab = nullptr, *comb = nullptr;
    ~gens_tables();
};

// bpg_r1cs_prove_prefetch (include/bpg.h): the device-free opening of one future proof, produced by a background host thread
struct prefetch_slot;
void prefetch_slot_free(prefetch_slot *p);

struct bpg_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, stream2 = nullptr;
    // Priority split (bpg_ctx_create): `stream` and `stream2` are created at the device's HIGHEST priority, `acc_stream` at the
    // LOWEST, and only k_msm_accumulate is launched there (fenced by ev_acc[0] / ev_acc[1]).  With many provers sharing a GPU
    // a multi-wave accumulate grid otherwise holds every block slot until its last wave, and the memory- / latency-bound
    // kernels of all other proofs queue behind it; with the split the block scheduler hands freed slots to them first and
    // they overlap with the integer-pipe-bound accumulation.
    cudaStream_t acc_stream = nullptr;
    cudaEvent_t ev_acc[2] = {nullptr, nullptr};
    cudaEvent_t ev = nullptr, ev2 = nullptr;
    // resident generators
    size_t cap = 0;           // per-chain capacity (power of two); tables cover 2*cap+2 points
    uint32_t ptotal = 0;
    std::shared_ptr<gens_tables> gens; // owner of tab / comb below (shared between contexts unless BPG_PRIVATE_TABLES=1)
    ge_an *tab = nullptr;     // [BPG_NWIN][ptotal]
    ge_an *comb = nullptr;    // [2][32][128] signed 8-bit comb for B and B~ (Pedersen commits)
    // MSM workspace
    dev_buf counts, offsets, cursor, sorted, partial, buckets, lvlP, lvlQ, heavy, results;
    // generic scratch
    dev_buf scratch[16];
    dev_buf batch_gh;         // batch verification: every proof's g | h scalars
    dev_buf heavy_part;       // segment sums of heavy buckets (k_msm_heavy)
    dev_buf vb_sums;          // variable-base MSM: the 16 window sums of every group
    dev_buf mat_pts, mat_ext, mat_tab; // late fold: materialised G^(k) | H^(k), their window chain, their affine-Niels tables
    // one proof split over the ranks of a node (bpg_ctx_set_shard): exchange buffers and the caller's all-gather
    int shard_rank = 0, shard_world = 1;
    void *shard_send = nullptr, *shard_recv = nullptr; size_t shard_cap = 0;
    bpg_allgather_fn shard_fn = nullptr; void *shard_user = nullptr;
    void *comm = nullptr;     // ncclComm_t of a context sharded with bpg_comm_init (exchange on the context's stream)
    dev_buf comm_recv;        // gathered partial points of all ranks
    void *h_pinned = nullptr; size_t h_pinned_cap = 0; // two pinned staging buffers for the raw transcript-RNG draws (prover.inl)
    cudaEvent_t ev_stage[2] = {nullptr, nullptr};      // upload of staging buffer b has completed
    prefetch_slot *pre[2] = {nullptr, nullptr};
    cudaEvent_t tev[16] = {nullptr};
    int prof_on = 0;
    std::vector<cudaEvent_t> prof_ev; // pairs (start, stop) around k_msm_accumulate
    uint32_t *prof_pairs = nullptr;   // pinned: pair count per profiled launch
    size_t prof_n = 0;
    uint64_t launches = 0;    // kernels launched (reported by bench.py as gpu_launches)
    std::string last_error;
};

// bpg.cu
int bpg_stream_sync(bpg_ctx *ctx, cudaStream_t s);
int msm_run(bpg_ctx *c, cudaStream_t s, msm_plan *plan, ge *d_out /* ngroups extended points */);
// late fold (kernels_msm.cuh): tables of the 2 n' folded generators (+ B at index 2 n') into ctx->mat_tab, ptotal = 2 n' + 2
int msm_materialise_fold(bpg_ctx *c, cudaStream_t s, uint32_t N, uint32_t nprime, const sc *d_EG, const sc *d_EH, int lean, int shard);
