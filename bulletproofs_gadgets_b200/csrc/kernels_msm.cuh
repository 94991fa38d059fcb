// kernels_msm.cuh -- fixed-base signed-digit Pippenger over the resident window tables (K2/K4/K6).
//
// Replaces RistrettoPoint::multiscalar_mul / vartime_multiscalar_mul of curve25519-dalek 1.x as called by
// the bulletproofs fork inside Prover::prove, InnerProductProof::create and Verifier::verify
// (reference call sites src/bin/prover.rs:93, src/bin/verifier.rs:90).
//
// Every term (scalar k, resident point P) is recoded into 16 signed 16-bit digits d_w.  Because the table holds
// 2^(16w) P for every window, the pair (term, w) contributes d_w * tab[w][P] and ALL windows share one bucket
// set per output group: bucket |d| accumulates sign(d) * tab[w][P].  The result is sum_b b * bucket[b]; there
// are no doublings and no per-window combine.
//   1. k_msm_digits<COUNT>   histogram of bucket populations            (atomics on a 128 KB..400 KB array, L2)
//   2. k_msm_scan            exclusive scan -> bucket offsets
//   3. k_msm_digits<SCATTER> counting-sort scatter of (table index | sign) into bucket order
//   4. k_msm_accumulate      each thread sums one fixed-size chunk of the sorted pairs (uniform work whatever the
//                            scalar distribution); runs that cover a whole bucket are written directly, runs cut
//                            by a chunk boundary go to per-chunk partial slots
//   5. k_msm_finish          per bucket: identity if empty, sum of partial slots if split over few chunks,
//                            else queued for 6
//   6. k_msm_heavy           block-wide tree reduction for buckets split over many chunks (e.g. the bit-valued
//                            a_L/a_R of range proofs put half of all terms in bucket 1)
//   7. k_msm_wsum_level / k_msm_wsum_tail   log-depth evaluation of sum_b b * bucket[b]
#pragma once
#include "kernels_core.cuh"

struct msm_params {
    msm_seg seg[BPG_MAX_SEGS];
    int nseg;
    uint32_t total;
    uint32_t ptotal;
};

// signed 16-bit digits: d in [-32768, 32767], sum d_w 2^(16w) = k  (k < 2^253 so the top digit never carries out)
__device__ __forceinline__ void sc_digits16(int *d, const sc &k) {
    int carry = 0;
#pragma unroll
    for (int w = 0; w < 16; w++) {
        int raw = (int)((k.v[w >> 1] >> (16 * (w & 1))) & 0xFFFF) + carry;
        carry = raw >= 32768;
        d[w] = raw - (carry << 16);
    }
}

template <int SCATTER>
__global__ void __launch_bounds__(256) k_msm_digits(msm_params P, uint32_t *__restrict__ counts_or_cursor, uint32_t *__restrict__ sorted) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.total) return;
    int si = 0;
#pragma unroll
    for (int k = 1; k < BPG_MAX_SEGS; k++)
        if (k < P.nseg && t >= P.seg[k].start) si = k;
    const msm_seg &S = P.seg[si];
    uint32_t j = t - S.start;
    sc k;
    ld_sc(k, &S.scalars[j]);
    if (S.reduce) sc_reduce(k, k);
    int d[16];
    sc_digits16(d, k);
    uint32_t grp = S.group;
    if (S.alt) grp ^= (j >> (S.alt - 1)) & 1u;
    uint32_t base = grp * BPG_NBP;
    uint32_t pidx = S.p0 + j;
#pragma unroll
    for (int w = 0; w < 16; w++) {
        int dw = d[w];
        if (dw == 0) continue;
        uint32_t mag = dw < 0 ? (uint32_t)(-dw) : (uint32_t)dw;
        if (SCATTER) {
            uint32_t pos = atomicAdd(&counts_or_cursor[base + mag], 1u);
            sorted[pos] = ((uint32_t)w * P.ptotal + pidx) | (dw < 0 ? 0x80000000u : 0u);
        } else {
            atomicAdd(&counts_or_cursor[base + mag], 1u);
        }
    }
}

// exclusive scan of counts[0..n) into offsets[0..n] and cursor[0..n) ; single block, n <= 4 * 32800
__global__ void __launch_bounds__(1024) k_msm_scan(const uint32_t *__restrict__ counts, uint32_t n, uint32_t *__restrict__ offsets, uint32_t *__restrict__ cursor) {
    __shared__ uint32_t part[1024];
    uint32_t t = threadIdx.x;
    uint32_t per = (n + 1023u) / 1024u;
    uint32_t lo = t * per, hi = min(n, lo + per);
    uint32_t s = 0;
    for (uint32_t i = lo; i < hi; i++) s += counts[i];
    part[t] = s;
    __syncthreads();
    // Hillis-Steele inclusive scan over 1024 partials
    for (uint32_t off = 1; off < 1024; off <<= 1) {
        uint32_t v = t >= off ? part[t - off] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    uint32_t run = t == 0 ? 0 : part[t - 1];
    for (uint32_t i = lo; i < hi; i++) { offsets[i] = run; cursor[i] = run; run += counts[i]; }
    if (t == 1023) offsets[n] = part[1023];
}

// value -> table entry (negated if the sign bit is set)
__device__ __forceinline__ void msm_load_entry(ge_an &a, const ge_an *__restrict__ tab, uint32_t v) {
    ld_an(a, &tab[v & 0x7FFFFFFFu]);
    if (v & 0x80000000u) {
        fe t = a.ypx; a.ypx = a.ymx; a.ymx = t;
        fe_neg(a.t2d, a.t2d);
    }
}

// partial slot layout: partial[2*chunk + 0] = run touching the chunk start, [2*chunk + 1] = run touching the chunk end only
__global__ void __launch_bounds__(128) k_msm_accumulate(const uint32_t *__restrict__ sorted, const uint32_t *__restrict__ offsets, uint32_t nbuckets,
                                                         const ge_an *__restrict__ tab, ge *__restrict__ buckets, ge *__restrict__ partial) {
    uint32_t M = offsets[nbuckets];
    uint32_t chunk = blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t start64 = (uint64_t)chunk * BPG_CHUNK;
    if (start64 >= M) return;
    uint32_t start = (uint32_t)start64, end = min(M, start + BPG_CHUNK);
    // bucket containing `start`: largest b with offsets[b] <= start
    uint32_t lo = 0, hi = nbuckets;
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (offsets[mid] <= start) lo = mid; else hi = mid; }
    uint32_t b = lo;
    uint32_t bend = offsets[b + 1];
    while (bend <= start) { b++; bend = offsets[b + 1]; } // skip empty buckets that share the offset
    uint32_t bstart = offsets[b];
    uint32_t run_start = start;
    ge acc;
    ge_identity(acc);
    uint32_t pos = start;
#pragma unroll 1
    while (pos < end) {
        uint32_t stop = min(end, bend);
        // prefetch the sorted values of this run piece in registers one ahead
        uint32_t v = sorted[pos];
#pragma unroll 1
        while (pos < stop) {
            uint32_t vn = (pos + 1 < stop) ? sorted[pos + 1] : 0u;
            ge_an a;
            msm_load_entry(a, tab, v);
            ge_add_an(acc, acc, a);
            v = vn;
            pos++;
        }
        // run [run_start, pos) of bucket b ends here (bucket boundary or chunk end)
        bool complete = (run_start == bstart) && (pos == bend);
        if (complete) st_ge(&buckets[b], acc);
        else st_ge(&partial[2ull * chunk + (run_start == start ? 0 : 1)], acc);
        if (pos < end) { // advance to the next non-empty bucket
            do { b++; bstart = bend; bend = offsets[b + 1]; } while (bend <= pos);
            run_start = pos;
            ge_identity(acc);
        }
    }
}

__global__ void __launch_bounds__(128) k_msm_finish(const uint32_t *__restrict__ offsets, uint32_t nbuckets, ge *__restrict__ buckets,
                                                     const ge *__restrict__ partial, uint32_t *__restrict__ heavy_list, uint32_t *__restrict__ heavy_count) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbuckets) return;
    uint32_t s = offsets[b], e = offsets[b + 1];
    if (s == e) { ge id; ge_identity(id); st_ge(&buckets[b], id); return; }
    uint32_t c0 = s / BPG_CHUNK, c1 = (e - 1) / BPG_CHUNK;
    if (c0 == c1) return; // written by the accumulate kernel
    if (c1 - c0 + 1 > BPG_HEAVY_SPAN) { heavy_list[atomicAdd(heavy_count, 1u)] = b; return; }
    ge acc;
    ld_ge(acc, &partial[2ull * c0 + ((s % BPG_CHUNK) == 0 ? 0 : 1)]);
#pragma unroll 1
    for (uint32_t c = c0 + 1; c <= c1; c++) {
        ge q;
        ld_ge(q, &partial[2ull * c]);
        ge_add(acc, acc, q);
    }
    st_ge(&buckets[b], acc);
}
// one block per heavy bucket: threads stride over its partial slots, then tree-reduce in shared memory
__global__ void __launch_bounds__(128) k_msm_heavy(const uint32_t *__restrict__ offsets, ge *__restrict__ buckets, const ge *__restrict__ partial,
                                                    const uint32_t *__restrict__ heavy_list, const uint32_t *__restrict__ heavy_count) {
    __shared__ ge smem[128];
    uint32_t nh = *heavy_count;
    for (uint32_t h = blockIdx.x; h < nh; h += gridDim.x) {
        uint32_t b = heavy_list[h];
        uint32_t s = offsets[b], e = offsets[b + 1];
        uint32_t c0 = s / BPG_CHUNK, c1 = (e - 1) / BPG_CHUNK;
        ge acc;
        ge_identity(acc);
        for (uint32_t c = c0 + threadIdx.x; c <= c1; c += blockDim.x) {
            ge q;
            uint32_t slot = (c == c0 && (s % BPG_CHUNK) != 0) ? 1u : 0u;
            ld_ge(q, &partial[2ull * c + slot]);
            ge_add(acc, acc, q);
        }
        block_sum_points(acc, smem);
        if (threadIdx.x == 0) st_ge(&buckets[b], acc);
        __syncthreads();
    }
}

// ---- weighted bucket sum.  Items i in [0,n): contribution i*P_i + Q_i.  One thread folds 4 items:
//      P'_s = 4 * sum_k P_{4s+k},   Q'_s = sum_k (k*P_{4s+k} + Q_{4s+k}).   grid.y = group.
__device__ __forceinline__ void wsum_fold4(ge &Pn, ge &Qn, const ge *Pin, const ge *Qin, uint32_t n, uint32_t s) {
    ge run, wacc, t;
    uint32_t i3 = 4 * s + 3, i2 = 4 * s + 2, i1 = 4 * s + 1, i0 = 4 * s;
    if (i3 < n) ld_ge(run, &Pin[i3]); else ge_identity(run);
    wacc = run;
    if (i2 < n) { ld_ge(t, &Pin[i2]); ge_add(run, run, t); }
    ge_add(wacc, wacc, run);
    if (i1 < n) { ld_ge(t, &Pin[i1]); ge_add(run, run, t); }
    ge_add(wacc, wacc, run);
    if (i0 < n) { ld_ge(t, &Pin[i0]); ge_add(run, run, t); }
    if (Qin) {
#pragma unroll 1
        for (uint32_t k = 0; k < 4; k++)
            if (4 * s + k < n) { ld_ge(t, &Qin[4 * s + k]); ge_add(wacc, wacc, t); }
    }
    ge_dbl(run, run);
    ge_dbl(run, run);
    Pn = run;
    Qn = wacc;
}
__global__ void __launch_bounds__(128) k_msm_wsum_level(const ge *__restrict__ Pin, const ge *__restrict__ Qin, uint32_t n, uint32_t in_stride,
                                                         ge *__restrict__ Pout, ge *__restrict__ Qout, uint32_t out_stride) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t nseg = (n + 3) / 4;
    if (s >= nseg) return;
    uint32_t g = blockIdx.y;
    ge Pn, Qn;
    wsum_fold4(Pn, Qn, Pin + (size_t)g * in_stride, Qin ? Qin + (size_t)g * in_stride : nullptr, n, s);
    st_ge(&Pout[(size_t)g * out_stride + s], Pn);
    st_ge(&Qout[(size_t)g * out_stride + s], Qn);
}
// finishes the recursion inside one block per group (n <= 512 items), result -> out[g]
__global__ void __launch_bounds__(128) k_msm_wsum_tail(ge *__restrict__ P, ge *__restrict__ Q, uint32_t n, uint32_t stride, ge *__restrict__ P2,
                                                        ge *__restrict__ Q2, ge *__restrict__ out) {
    uint32_t g = blockIdx.x;
    ge *Pin = P + (size_t)g * stride, *Qin = Q + (size_t)g * stride, *Pout = P2 + (size_t)g * stride, *Qout = Q2 + (size_t)g * stride;
    while (true) {
        uint32_t nseg = (n + 3) / 4;
        for (uint32_t s = threadIdx.x; s < nseg; s += blockDim.x) {
            ge Pn, Qn;
            wsum_fold4(Pn, Qn, Pin, Qin, n, s);
            st_ge(&Pout[s], Pn);
            st_ge(&Qout[s], Qn);
        }
        __syncthreads();
        if (nseg == 1) break;
        ge *tp = Pin; Pin = Pout; Pout = tp;
        tp = Qin; Qin = Qout; Qout = tp;
        n = nseg;
    }
    if (threadIdx.x == 0) { ge r; ld_ge(r, &Qout[0]); st_ge(&out[g], r); }
}
