// prover.inl -- MiMC entry points, R1CS circuit upload, Prover::prove / Verifier::verify drivers (host side).

// ================================================================ MiMC
extern "C" int bpg_mimc_set_constants(bpg_ctx *ctx, const uint8_t *consts486x32) {
    if (!ctx || !consts486x32) return BPG_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    std::vector<sc> c(MIMC_ROUNDS);
    for (int i = 0; i < MIMC_ROUNDS; i++) {
        sc t;
        sc_frombytes(t, consts486x32 + 32 * i);
        t.v[7] &= 0x7FFFFFFFu; // Scalar::from_bits (mimc.rs:69)
        sc_reduce(c[i], t);
    }
    {   // per-device symbol: uploaded under the constants lock and complete before any kernel can read it
        std::lock_guard<std::mutex> lk(g_const_mu);
        CUDA_TRY(cudaDeviceSynchronize());
        CUDA_TRY(cudaMemcpyToSymbol(c_mimc, c.data(), sizeof(sc) * MIMC_ROUNDS));
        CUDA_TRY(cudaDeviceSynchronize());
    }
    return BPG_OK;
}
extern "C" int bpg_mimc_sponge_batch(bpg_ctx *ctx, const uint8_t *blocks, const uint32_t *block_off, size_t n, uint8_t *out32, uint8_t *trace) {
    if (!ctx || (n && (!blocks || !block_off || !out32))) return BPG_E_ARG;
    if (!n) return BPG_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    size_t nblocks = block_off[n];
    size_t tbytes = trace ? nblocks * MIMC_ROUNDS * 6 * 32 : 0;
    CTX_TRY(ctx->scratch[6].ensure(32 * nblocks + 4 * (n + 1) + 64));
    CTX_TRY(ctx->scratch[7].ensure(32 * n + tbytes + 64));
    uint8_t *din = (uint8_t *)ctx->scratch[6].p, *dout = (uint8_t *)ctx->scratch[7].p;
    CUDA_TRY(cudaMemcpyAsync(din, blocks, 32 * nblocks, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(din + 32 * nblocks, block_off, 4 * (n + 1), cudaMemcpyHostToDevice, s));
    // event slots 14 / 15 bracket the kernel alone (bpg_event_elapsed_ms(ctx, 14, 15, &ms) after the call: MiMC roofline in bench.py)
    for (int e = 14; e < 16; e++) if (!ctx->tev[e]) CUDA_TRY(cudaEventCreate(&ctx->tev[e]));
    CUDA_TRY(cudaEventRecord(ctx->tev[14], s));
    k_mimc_sponge<<<LAUNCH_1D(n, 128), 0, s>>>((const sc *)din, (const uint32_t *)(din + 32 * nblocks), (uint32_t)n, (sc *)dout,
                                                trace ? (sc *)(dout + 32 * n) : nullptr);
    KCHECK();
    CUDA_TRY(cudaEventRecord(ctx->tev[15], s));
    D2H_TRY(ctx, out32, dout, 32 * n, s);
    if (trace) D2H_TRY(ctx, trace, dout + 32 * n, tbytes, s);
    SYNC_TRY(ctx, s);
    return BPG_OK;
}
// mimc_hash: be_to_scalars + pad (mimc.rs:61-97, conversions.rs:26-30) on the host, sponge on the device
extern "C" int bpg_mimc_hash_batch(bpg_ctx *ctx, const uint8_t *data, const uint64_t *offsets, size_t n, uint8_t *out32) {
    if (!ctx || (n && (!offsets || !out32))) return BPG_E_ARG;
    std::vector<uint8_t> blocks;
    std::vector<uint32_t> boff(n + 1, 0);
    for (size_t i = 0; i < n; i++) {
        size_t len = (size_t)(offsets[i + 1] - offsets[i]);
        if (len == 0) return BPG_E_ARG; // the reference panics on an empty preimage (mimc.rs:78 last().unwrap())
        const uint8_t *p = data + offsets[i];
        size_t nb = (len + 31) / 32;
        size_t base = blocks.size();
        blocks.resize(base + 32 * (nb + 1), 0);
        uint8_t *le = blocks.data() + base;
        for (size_t k = 0; k < len; k++) le[k] = p[len - 1 - k];
        uint8_t *last = le + 32 * (nb - 1);
        last[31] &= 0x7F;
        int l = 32;
        while (l > 0 && last[l - 1] == 0) l--;
        if (l < 32) { for (int k = l; k < 32; k++) last[k] = (uint8_t)(32 - l); blocks.resize(base + 32 * nb); }
        else { memset(le + 32 * nb, 32, 32); nb++; }
        boff[i + 1] = boff[i] + (uint32_t)nb;
    }
    return bpg_mimc_sponge_batch(ctx, blocks.data(), boff.data(), n, out32, nullptr);
}

// ================================================================ R1CS circuit (flattened-constraint matrix, column-major on the device)
#define BPG_VCOL 32u
struct bpg_circuit {
    bpg_ctx *ctx;
    size_t n, m, q, T;
    uint32_t ncols;              // 3n + m + 1 : wL | wR | wO | wV | wc
    uint32_t nv, nsplit, npartial;
    uint32_t *d_vcol_ptr = nullptr, *d_vcol_dst = nullptr, *d_row = nullptr, *d_split = nullptr;
    sc *d_coeff = nullptr;
};
extern "C" int bpg_circuit_create(bpg_ctx *ctx, size_t n, size_t m, size_t q, const uint32_t *row_ptr, const uint32_t *term_var,
                                  const uint8_t *term_coeff, bpg_circuit **out) {
    if (!ctx || !out || !row_ptr || (q && row_ptr[q] && (!term_var || !term_coeff))) return BPG_E_ARG;
    *out = nullptr;
    if (n >= (1u << 22) || m >= (1u << 22)) return BPG_E_SIZE;
    CUDA_TRY(cudaSetDevice(ctx->device));
    size_t T = row_ptr[q];
    uint32_t ncols = (uint32_t)(3 * n + m + 1);
    // column of each term, counting sort by column (stable => rows ascending inside a column)
    std::vector<uint32_t> col_of(T), cnt(ncols + 1, 0);
    for (size_t k = 0; k < T; k++) {
        uint32_t kind = term_var[k] >> 29, idx = term_var[k] & 0x1FFFFFFFu, c;
        if (kind <= 2) { if (idx >= n) return BPG_E_ARG; c = (uint32_t)(kind * n + idx); }
        else if (kind == 3) { if (idx >= m) return BPG_E_ARG; c = (uint32_t)(3 * n + idx); }
        else if (kind == 4) c = (uint32_t)(3 * n + m);
        else return BPG_E_ARG;
        col_of[k] = c;
        cnt[c + 1]++;
    }
    for (uint32_t c = 0; c < ncols; c++) cnt[c + 1] += cnt[c];
    std::vector<uint32_t> e_row(T ? T : 1), pos(cnt.begin(), cnt.end() - 1);
    std::vector<sc> e_coeff(T ? T : 1);
    for (size_t r = 0; r < q; r++)
        for (uint32_t k = row_ptr[r]; k < row_ptr[r + 1]; k++) {
            uint32_t c = col_of[k], p = pos[c]++;
            e_row[p] = (uint32_t)r;
            sc cf, red;
            sc_frombytes(cf, term_coeff + 32 * (size_t)k);
            sc_reduce(red, cf);
            if (c >= 3 * n) sc_neg_r(red, red); // wV and wc accumulate -z^k * coeff
            e_coeff[p] = red;
        }
    // virtual columns of <= BPG_VCOL entries
    std::vector<uint32_t> vptr, vdst, split;
    vptr.push_back(0);
    uint32_t npartial = 0;
    for (uint32_t c = 0; c < ncols; c++) {
        uint32_t k0 = cnt[c], k1 = cnt[c + 1], len = k1 - k0;
        if (len <= BPG_VCOL) { vptr.push_back(k1); vdst.push_back(c); continue; } // also writes zero for empty columns
        uint32_t parts = (len + BPG_VCOL - 1) / BPG_VCOL;
        split.push_back(c); split.push_back(npartial); split.push_back(parts);
        for (uint32_t p = 0; p < parts; p++) { vptr.push_back(std::min(k1, k0 + (p + 1) * BPG_VCOL)); vdst.push_back(0x80000000u | npartial++); }
    }
    bpg_circuit *c = new bpg_circuit();
    c->ctx = ctx; c->n = n; c->m = m; c->q = q; c->T = T; c->ncols = ncols;
    c->nv = (uint32_t)vdst.size(); c->nsplit = (uint32_t)(split.size() / 3); c->npartial = npartial;
    if (split.empty()) split.push_back(0);
    cudaError_t e = cudaSuccess;
    auto up = [&](void **d, const void *h, size_t bytes) {
        if (e != cudaSuccess) return;
        e = cudaMalloc(d, bytes ? bytes : 4);
        if (e == cudaSuccess && bytes) e = cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice);
    };
    up((void **)&c->d_vcol_ptr, vptr.data(), vptr.size() * 4);
    up((void **)&c->d_vcol_dst, vdst.data(), vdst.size() * 4);
    up((void **)&c->d_row, e_row.data(), e_row.size() * 4);
    up((void **)&c->d_coeff, e_coeff.data(), e_coeff.size() * sizeof(sc));
    up((void **)&c->d_split, split.data(), split.size() * 4);
    if (e != cudaSuccess) { bpg_set_cuda_error(e, __FILE__, __LINE__); ctx->last_error = g_cuda_err; bpg_circuit_destroy(c); return BPG_E_CUDA; }
    *out = c;
    return BPG_OK;
}
extern "C" void bpg_circuit_destroy(bpg_circuit *c) {
    if (!c) return;
    cudaFree(c->d_vcol_ptr); cudaFree(c->d_vcol_dst); cudaFree(c->d_row); cudaFree(c->d_coeff); cudaFree(c->d_split);
    delete c;
}

// ================================================================ witness evaluation on the device
extern "C" int bpg_witness_eval(bpg_ctx *ctx, size_t n, size_t m, const uint32_t *lc_ptr, const uint32_t *term_var, const uint8_t *term_coeff,
                                const uint8_t *v, uint8_t *aL, uint8_t *aR, uint8_t *aO) {
    if (!ctx || (n && (!lc_ptr || !aL || !aR || !aO)) || (m && !v)) return BPG_E_ARG;
    if (!n) return BPG_OK;
    if (n >= (1u << 22) || m >= (1u << 22)) return BPG_E_SIZE;
    size_t T = lc_ptr[2 * n];
    if (T && (!term_var || !term_coeff)) return BPG_E_ARG;
    CUDA_TRY(cudaSetDevice(ctx->device));
    // dependency levels: assigned multipliers (empty pair of linear combinations) are level 0
    std::vector<uint32_t> level(n, 0);
    uint32_t depth = 0;
    for (size_t i = 0; i < n; i++) {
        uint32_t t0 = lc_ptr[2 * i], tm = lc_ptr[2 * i + 1], t1 = lc_ptr[2 * i + 2];
        if (tm < t0 || t1 < tm || t1 > T) return BPG_E_ARG; // the two term ranges of a multiplier must be nested in [0, T]
        if (t0 == t1) continue;
        uint32_t lv = 1;
        for (uint32_t t = t0; t < t1; t++) {
            uint32_t kind = term_var[t] >> 29, idx = term_var[t] & 0x1FFFFFFFu;
            if (kind <= 2) { if (idx >= i) return BPG_E_ARG; lv = std::max(lv, level[idx] + 1); } // only earlier multipliers
            else if (kind == 3) { if (idx >= m) return BPG_E_ARG; }
            else if (kind != 4) return BPG_E_ARG;
        }
        level[i] = lv;
        depth = std::max(depth, lv);
    }
    std::vector<uint32_t> lptr(depth + 2, 0), order(n);
    for (size_t i = 0; i < n; i++) lptr[level[i] + 1]++;
    for (uint32_t l = 0; l <= depth; l++) lptr[l + 1] += lptr[l];
    { std::vector<uint32_t> pos(lptr.begin(), lptr.end() - 1); for (size_t i = 0; i < n; i++) order[pos[level[i]]++] = (uint32_t)i; }
    cudaStream_t s = ctx->stream;
    size_t b_w = 3 * 32 * n, b_v = 32 * m, b_ord = 4 * n, b_ptr = 4 * (2 * n + 1), b_tv = 4 * T, b_tc = 32 * T, b_lp = 4 * ((size_t)depth + 2);
    CTX_TRY(ctx->scratch[8].ensure(b_w + b_v + 64));
    CTX_TRY(ctx->scratch[11].ensure(b_ord + b_ptr + b_tv + b_tc + b_lp + 256));
    sc *d_aL = (sc *)ctx->scratch[8].p, *d_aR = d_aL + n, *d_aO = d_aR + n, *d_v = d_aO + n;
    uint8_t *d = (uint8_t *)ctx->scratch[11].p;
    sc *d_tc = (sc *)d;                                  // 32-byte aligned first
    uint32_t *d_order = (uint32_t *)(d + b_tc), *d_ptr = d_order + n, *d_tv = d_ptr + (2 * n + 1), *d_lptr = d_tv + T;
    CUDA_TRY(cudaMemcpyAsync(d_aL, aL, 32 * n, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_aR, aR, 32 * n, cudaMemcpyHostToDevice, s));
    if (m) CUDA_TRY(cudaMemcpyAsync(d_v, v, b_v, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_order, order.data(), b_ord, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_ptr, lc_ptr, b_ptr, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_lptr, lptr.data(), b_lp, cudaMemcpyHostToDevice, s));
    if (T) {
        CUDA_TRY(cudaMemcpyAsync(d_tv, term_var, b_tv, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(d_tc, term_coeff, b_tc, cudaMemcpyHostToDevice, s));
        k_sc_reduce_inplace<<<LAUNCH_1D(T, 256), 0, s>>>(d_tc, (uint32_t)T);
        KCHECK();
    }
    k_sc_reduce_inplace<<<LAUNCH_1D(2 * n + m, 256), 0, s>>>(d_aL, (uint32_t)(2 * n)); // caller-assigned a_L, a_R (from_bits semantics)
    KCHECK();
    if (m) { k_sc_reduce_inplace<<<LAUNCH_1D(m, 256), 0, s>>>(d_v, (uint32_t)m); KCHECK(); }
    if (lptr[1]) { k_witness_assigned<<<LAUNCH_1D(lptr[1], 128), 0, s>>>(d_order, lptr[1], d_aL, d_aR, d_aO); KCHECK(); }
    // wide levels: one grid each; runs of consecutive narrow levels (<= 256 multipliers): one single-block launch per run
    const uint32_t NARROW = 256;
    for (uint32_t l = 1; l <= depth;) {
        uint32_t k0 = lptr[l], k1 = lptr[l + 1];
        if (k1 - k0 > NARROW) {
            k_witness_level<<<LAUNCH_1D(k1 - k0, 128), 0, s>>>(d_order, k0, k1, d_ptr, d_tv, d_tc, d_aL, d_aR, d_aO, d_v);
            KCHECK();
            l++;
            continue;
        }
        uint32_t l1 = l + 1;
        while (l1 <= depth && lptr[l1 + 1] - lptr[l1] <= NARROW) l1++;
        k_witness_levels_block<<<1, 256, 0, s>>>(d_order, d_lptr, l, l1, d_ptr, d_tv, d_tc, d_aL, d_aR, d_aO, d_v);
        KCHECK();
        l = l1;
    }
    D2H_TRY(ctx, aL, d_aL, 32 * n, s);
    D2H_TRY(ctx, aR, d_aR, 32 * n, s);
    D2H_TRY(ctx, aO, d_aO, 32 * n, s);
    SYNC_TRY(ctx, s);
    CUDA_TRY(cudaMemsetAsync(ctx->scratch[8].p, 0, b_w + b_v, s)); // the witness is a prover secret
    return BPG_OK;
}

// ================================================================ host scalar helpers
namespace {
const sc SC_ONE_H = {{1, 0, 0, 0, 0, 0, 0, 0}};
inline sc h_mul(const sc &a, const sc &b) { return bpgh::sc_mul64(a, b); }
inline sc h_add(const sc &a, const sc &b) { sc r; sc_add_r(r, a, b); return r; }
inline sc h_sub(const sc &a, const sc &b) { sc r; sc_sub_r(r, a, b); return r; }
inline sc h_inv(const sc &a) { return bpgh::sc_invert64(a); }
inline sc h_wide(const uint8_t b[64]) { return bpgh::sc_wide64(b); }
inline sc challenge_scalar(bpgh::Transcript &t, const char *label) { uint8_t b[64]; t.challenge(label, b, 64); return h_wide(b); }
inline sc rng_scalar(bpgh::TranscriptRng &rng) { uint8_t b[64]; rng.fill_bytes(b, 64); return h_wide(b); }
inline void append_scalar(bpgh::Transcript &t, const char *label, const sc &x) { uint8_t b[32]; sc_tobytes(b, x); t.append(label, b, 32); }
inline size_t next_pow2(size_t n) { size_t p = 1; while (p < n) p <<= 1; return p; }
inline bool sc_canonical_bytes(sc &out, const uint8_t *b) { sc_frombytes(out, b); return !sc_geq_l(out.v); }

// device-side power tables of one base
struct pow_tab { sc *lo, *hi; };
int make_pow_tables(bpg_ctx *ctx, cudaStream_t s, const sc *d_base, uint32_t max_exp, sc *d_store, pow_tab &t) {
    uint32_t nhi = (max_exp >> 10) + 2;
    t.lo = d_store; t.hi = d_store + 1024;
    k_pow_tables<<<LAUNCH_1D(1024 + nhi, 128), 0, s>>>(d_base, t.lo, t.hi, nhi);
    KCHECK();
    return BPG_OK;
}
inline size_t pow_tab_size(uint32_t max_exp) { return 1024 + (max_exp >> 10) + 2; }

int run_flatten(bpg_ctx *ctx, cudaStream_t s, bpg_circuit *c, const pow_tab &z, sc *d_w, dev_buf &partial_buf) {
    CTX_TRY(partial_buf.ensure(((size_t)c->npartial + 1) * sizeof(sc)));
    if (c->nv) {
        k_flatten_vcols<<<LAUNCH_1D(c->nv, 128), 0, s>>>(c->d_vcol_ptr, c->d_vcol_dst, c->nv, c->d_row, c->d_coeff, z.lo, z.hi, d_w, (sc *)partial_buf.p);
        KCHECK();
    }
    if (c->nsplit) {
        k_flatten_split<<<c->nsplit, 128, 0, s>>>(c->d_split, c->nsplit, (const sc *)partial_buf.p, d_w);
        KCHECK();
    }
    return BPG_OK;
}
} // namespace

// optional phase trace (BPG_TRACE=1): wall-clock per phase of bpg_r1cs_prove, printed to stderr
#include <chrono>
struct phase_trace {
    bool on; std::chrono::steady_clock::time_point t0; std::string out;
    phase_trace() : on(getenv("BPG_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char *name) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        char b[96]; snprintf(b, sizeof b, " %s=%.3f", name, std::chrono::duration<double, std::milli>(t1 - t0).count());
        out += b; t0 = t1;
    }
    ~phase_trace() { if (on) fprintf(stderr, "[bpg prove ms]%s\n", out.c_str()); }
};

// ================================================================ prefetch of a proof's transcript-RNG stream
// A byte-exact proof needs 2n SEQUENTIAL TranscriptRng draws (one Keccak-f each: ~0.7 s at n = 2^20, whatever the hardware) before its
// S commitment can start; nothing else of the proof can overlap them.  A prover that knows its next job (a queue of proofs) hides
// that latency behind the device work of the CURRENT proof: bpg_r1cs_prove_prefetch commits the openings (one small kernel),
// then a background host thread replays the transcript opening, draws the stream through the lane-batched RNG service and stages
// the raw draws in HBM over the context's second stream.  The matching bpg_r1cs_prove finds everything ready.
struct prefetch_slot {
    std::thread worker;
    bool active = false;
    int rc = BPG_OK;
    // key
    bpg_circuit *circ = nullptr;
    std::vector<uint8_t> label, vbl;
    uint8_t ext[32];
    // products
    std::vector<uint8_t> Venc;
    bpgh::Transcript t;       // after dom-sep, V*, m
    bpgh::Strobe rng_state;   // TranscriptRng after ib, ob, sb and the 2n bulk draws
    sc ib, ob, sb;
    dev_buf d_raw;            // 2n x 64 raw bytes, device
    void *h_stage = nullptr;  // 2 pinned chunk buffers of this slot
    cudaEvent_t ev[2] = {nullptr, nullptr}, done = nullptr;
};
static const size_t BPG_RNG_CHUNK = (size_t)1 << 17; // draws per staging chunk (8 MB)
void prefetch_slot_free(prefetch_slot *p) {
    if (!p) return;
    if (p->worker.joinable()) p->worker.join();
    if (p->h_stage) { explicit_bzero(p->h_stage, 2 * BPG_RNG_CHUNK * 64); cudaFreeHost(p->h_stage); }
    if (p->d_raw.p) cudaMemset(p->d_raw.p, 0, p->d_raw.cap);
    p->d_raw.release();
    for (int i = 0; i < 2; i++) if (p->ev[i]) cudaEventDestroy(p->ev[i]);
    if (p->done) cudaEventDestroy(p->done);
    explicit_bzero(&p->rng_state, sizeof p->rng_state);
    delete p;
}
static void prefetch_discard(prefetch_slot *p) { // keep the buffers, drop (and wipe) the contents
    if (p->worker.joinable()) p->worker.join();
    p->active = false;
    explicit_bzero(&p->rng_state, sizeof p->rng_state);
    explicit_bzero(&p->ib, sizeof p->ib); explicit_bzero(&p->ob, sizeof p->ob); explicit_bzero(&p->sb, sizeof p->sb);
}
static bool prefetch_matches(const prefetch_slot *p, bpg_circuit *c, const uint8_t *label, size_t label_len, const uint8_t *v_blinding, const uint8_t *ext) {
    return p && p->active && p->circ == c && p->label.size() == label_len && memcmp(p->label.data(), label, label_len) == 0 &&
           p->vbl.size() == 32 * c->m && (c->m == 0 || memcmp(p->vbl.data(), v_blinding, 32 * c->m) == 0) && memcmp(p->ext, ext, 32) == 0;
}
extern "C" int bpg_r1cs_prove_prefetch(bpg_ctx *ctx, bpg_circuit *c, const uint8_t *label, size_t label_len, const uint8_t *v, const uint8_t *v_blinding,
                                       const uint8_t ext_rng32[32], unsigned flags) {
    if (!ctx || !c || !label || !ext_rng32 || (c->m && (!v || !v_blinding))) return BPG_E_ARG;
    if (flags & BPG_FLAG_FAST_BLINDING) return BPG_OK; // nothing sequential to hide
    CUDA_TRY(cudaSetDevice(ctx->device));
    int si = -1;
    for (int i = 0; i < 2; i++) if (!ctx->pre[i] || !ctx->pre[i]->active) { si = i; break; }
    if (si < 0) return BPG_OK; // both slots hold pending proofs: the hint is dropped, the proof will draw its stream itself
    if (!ctx->pre[si]) ctx->pre[si] = new prefetch_slot();
    prefetch_slot *p = ctx->pre[si];
    if (p->worker.joinable()) p->worker.join();
    size_t n = c->n, m = c->m;
    if (!p->h_stage) {
        CUDA_TRY(cudaHostAlloc(&p->h_stage, 2 * BPG_RNG_CHUNK * 64, cudaHostAllocDefault));
        for (int i = 0; i < 2; i++) CUDA_TRY(cudaEventCreateWithFlags(&p->ev[i], cudaEventDisableTiming | cudaEventBlockingSync));
        CUDA_TRY(cudaEventCreateWithFlags(&p->done, cudaEventDisableTiming));
    }
    CTX_TRY(p->d_raw.ensure(128 * n + 64));
    p->circ = c; p->label.assign(label, label + label_len); p->vbl.assign(v_blinding, v_blinding + 32 * m); memcpy(p->ext, ext_rng32, 32);
    p->Venc.assign(32 * (m ? m : 1), 0);
    if (m) CTX_TRY(bpg_pedersen_commit(ctx, v, v_blinding, m, p->Venc.data())); // the only device work of the opening (this thread, main stream)
    else SYNC_TRY(ctx, ctx->stream); // (the commit call synchronises too: nothing of an earlier proof still touches this slot's buffers)
    p->rc = BPG_OK;
    p->active = true;
    int device = ctx->device;
    cudaStream_t s2 = ctx->stream2;
    p->worker = std::thread([p, n, m, device, s2] {
        cudaSetDevice(device);
        bpgh::Transcript t(p->label.data(), p->label.size());
        t.append("dom-sep", (const uint8_t *)"r1cs v1", 7);
        for (size_t i = 0; i < m; i++) t.append("V", p->Venc.data() + 32 * i, 32);
        t.append_u64("m", m);
        p->t = t;
        bpgh::TranscriptRng rng(t);
        for (size_t i = 0; i < m; i++) rng.rekey_with_witness_bytes("v_blinding", p->vbl.data() + 32 * i, 32);
        rng.finalize(p->ext);
        p->ib = rng_scalar(rng); p->ob = rng_scalar(rng); p->sb = rng_scalar(rng);
        size_t total = 2 * n, done = 0;
        int used[2] = {0, 0};
        for (size_t ck = 0; done < total; ck++) {
            int b = (int)(ck & 1);
            size_t cnt = std::min(BPG_RNG_CHUNK, total - done);
            uint8_t *hb = (uint8_t *)p->h_stage + (size_t)b * BPG_RNG_CHUNK * 64;
            if (used[b] && cudaEventSynchronize(p->ev[b]) != cudaSuccess) { p->rc = BPG_E_CUDA; break; }
            bpgh::RngService::get().draw64(rng, hb, cnt);
            if (cudaMemcpyAsync((uint8_t *)p->d_raw.p + 64 * done, hb, 64 * cnt, cudaMemcpyHostToDevice, s2) != cudaSuccess ||
                cudaEventRecord(p->ev[b], s2) != cudaSuccess) { p->rc = BPG_E_CUDA; break; }
            used[b] = 1;
            done += cnt;
        }
        if (cudaEventRecord(p->done, s2) != cudaSuccess) p->rc = BPG_E_CUDA;
        if (cudaEventSynchronize(p->done) != cudaSuccess) p->rc = BPG_E_CUDA; // uploads complete: the staging buffers can be wiped
        explicit_bzero(p->h_stage, 2 * BPG_RNG_CHUNK * 64);
        p->rng_state = rng.s;
        explicit_bzero(&rng, sizeof rng);
    });
    return BPG_OK;
}

// ================================================================ Prover::prove
// Buffer map (ctx->scratch):  8: witness aL|aR|aO|sL|sR (5N)   9: small scalars   10: power tables   11: w (3n+m+1)
//                            12: l1|r0|r1|r3 (4n) -> later sG|sH (2N)   13: a|b|EG|EH (4N)   14: partial sums   15: flatten partials
extern "C" long bpg_r1cs_prove(bpg_ctx *ctx, bpg_circuit *c, const uint8_t *label, size_t label_len, const uint8_t *aL, const uint8_t *aR,
                               const uint8_t *aO, const uint8_t *v, const uint8_t *v_blinding, const uint8_t ext_rng32[32], unsigned flags,
                               uint8_t *V_out, uint8_t *proof, size_t proof_cap) {
    bpg_inflight_guard inflight_;
    if (!ctx || !c || !label || !ext_rng32 || !proof) return BPG_E_ARG;
    size_t n = c->n, m = c->m;
    if ((n && (!aL || !aR || !aO)) || (m && (!v || !v_blinding))) return BPG_E_ARG;
    size_t N = next_pow2(n);
    int lgN = 0;
    while (((size_t)1 << lgN) < N) lgN++;
    bool legacy = flags & BPG_FLAG_LEGACY_FRAMING;
    size_t need = (legacy ? 0 : 1) + 32 * (size_t)((legacy ? 14 : 11) + 2 * lgN + 2);
    if (proof_cap < need) return BPG_E_SIZE;
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (ctx->cap < N) return BPG_E_SIZE; // R1CSError::InvalidGeneratorsLength
    cudaStream_t s = ctx->stream;
    const uint32_t pB = (uint32_t)(2 * ctx->cap), pBb = pB + 1;

    const int shard_on = ctx->shard_world > 1; // one proof over several ranks: MSMs cut by point range (bpg_ctx_set_shard)
    phase_trace tr;
    // a prefetched opening for exactly this proof (bpg_r1cs_prove_prefetch)?  Anything else that is pending stays for its own proof.
    prefetch_slot *pre = nullptr;
    if (!(flags & BPG_FLAG_FAST_BLINDING))
        for (int i = 0; i < 2 && !pre; i++)
            if (prefetch_matches(ctx->pre[i], c, label, label_len, v_blinding, ext_rng32)) pre = ctx->pre[i];
    if (pre) {
        if (pre->worker.joinable()) pre->worker.join();
        if (pre->rc != BPG_OK) { prefetch_discard(pre); pre = nullptr; }
    }
    bpgh::Transcript t(label, label_len);
    std::vector<uint8_t> Venc(32 * (m ? m : 1));
    bpgh::TranscriptRng rng(t);
    sc ib, ob, sb;
    if (pre) {
        t = pre->t;
        Venc = pre->Venc;
        rng.s = pre->rng_state;
        ib = pre->ib; ob = pre->ob; sb = pre->sb;
    } else {
        t.append("dom-sep", (const uint8_t *)"r1cs v1", 7);
        // ---- V_i = v_i B + blinding_i B~  (batched; Prover::commit appends each to the transcript)
        if (m) CTX_TRY(bpg_pedersen_commit(ctx, v, v_blinding, m, Venc.data()));
        for (size_t i = 0; i < m; i++) t.append("V", Venc.data() + 32 * i, 32);
        t.append_u64("m", m);
        rng = bpgh::TranscriptRng(t);
        for (size_t i = 0; i < m; i++) rng.rekey_with_witness_bytes("v_blinding", v_blinding + 32 * i, 32);
        rng.finalize(ext_rng32);
        ib = rng_scalar(rng); ob = rng_scalar(rng); sb = rng_scalar(rng);
    }
    if (V_out && m) memcpy(V_out, Venc.data(), 32 * m);
    tr.mark("commitV");

    // ---- upload the witness, start A_I1 / A_O1 while the host draws s_L, s_R
    CTX_TRY(ctx->scratch[8].ensure((5 * N + 8) * sizeof(sc)));
    CTX_TRY(ctx->scratch[9].ensure(256 * sizeof(sc)));
    sc *d_aL = (sc *)ctx->scratch[8].p, *d_aR = d_aL + N, *d_aO = d_aR + N, *d_sL = d_aO + N, *d_sR = d_sL + N;
    sc *d_small = (sc *)ctx->scratch[9].p; // [0..2] blindings, [3] y, [4] z, [5] yinv, [6..9] x,x2,x3,u, [10] w, [11..12] u,uinv, [13..14] cL*w,cR*w, [16..] T scalars
    if (n) {
        cudaMemcpyKind wk = (flags & BPG_FLAG_WITNESS_ON_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        CUDA_TRY(cudaMemcpyAsync(d_aL, aL, 32 * n, wk, s));
        CUDA_TRY(cudaMemcpyAsync(d_aR, aR, 32 * n, wk, s));
        CUDA_TRY(cudaMemcpyAsync(d_aO, aO, 32 * n, wk, s));
        k_sc_reduce_inplace<<<LAUNCH_1D(3 * N, 256), 0, s>>>(d_aL, (uint32_t)(2 * N + n)); // aL|aR|aO regions (N-strided; padding is never read)
        KCHECK();
    }
    sc h_small[3] = {ib, ob, sb};
    CUDA_TRY(cudaMemcpyAsync(d_small, h_small, sizeof h_small, cudaMemcpyHostToDevice, s));
    CTX_TRY(ctx->results.ensure(8 * sizeof(ge) + 64));
    CTX_TRY(ctx->scratch[14].ensure(4096 * 6 * sizeof(sc)));
    ge *res = (ge *)ctx->results.p;
    uint8_t *d_enc = (uint8_t *)(res + 4);
    msm_plan plan;
    memset(&plan, 0, sizeof plan); plan.lean = bpg_lean_now(); plan.shard = shard_on;
    plan.ngroups = 2;
    auto add_seg = [&](const sc *sp, size_t cnt, uint32_t p0, uint32_t g) {
        if (!cnt) return;
        msm_seg &sg = plan.seg[plan.nseg++];
        sg.scalars = sp; sg.n = (uint32_t)cnt; sg.p0 = p0; sg.group = g; sg.reduce = 0;
    };
    // (Evaluating <a_L, G> and <a_R, H> as output groups of their own -- each then within the privatised scatter's reach -- was
    // measured: 58.0 -> 57.8 ms per 2^20 proof, not worth a third bucket set; profiles/r02_sort_variants.jsonl.)
    add_seg(d_aL, n, 0, 0); add_seg(d_aR, n, (uint32_t)ctx->cap, 0); add_seg(d_small + 0, 1, pBb, 0);
    add_seg(d_aO, n, 0, 1); add_seg(d_small + 1, 1, pBb, 1);
    CTX_TRY(msm_run(ctx, s, &plan, res));
    tr.mark("launchAIAO");
    // s_L, s_R: 2n sequential TranscriptRng draws on the host (byte-exact with the reference) ...
    if (flags & BPG_FLAG_FAST_BLINDING) {
        // ... or a transcript-seeded counter-mode expansion on the device (one Keccak-f per element): valid proofs, other bytes
        uint8_t seed[64]; // two independent 32-byte seeds: s_L, s_R
        rng.fill_bytes(seed, 64);
        CUDA_TRY(cudaMemcpyAsync(d_small + 60, seed, 64, cudaMemcpyHostToDevice, s));
        if (n) {
            k_expand_blinding<<<LAUNCH_1D(n, 128), 0, s>>>((const uint64_t *)(d_small + 60), (uint32_t)n, d_sL);
            KCHECK();
            k_expand_blinding<<<LAUNCH_1D(n, 128), 0, s>>>((const uint64_t *)(d_small + 61), (uint32_t)n, d_sR);
            KCHECK();
        }
        tr.mark("rng(device)");
    } else {
        // The raw 64-byte draws come from the lane-batched RNG service (one Keccak-f per draw, shared SIMD registers with the
        // other provers of this process); Scalar::from_bytes_mod_order_wide happens on the device.  The stream is drawn in
        // chunks into two pinned staging buffers and each chunk is uploaded while the next one is drawn: no pageable-memory
        // staging copy by the driver (measured at n = 2^20: 128 MB per proof), and only 2 x 8 MB of secrets to wipe.
        if (n && pre) {
            // the stream was drawn and staged in HBM while the previous proof was on the device
            CUDA_TRY(cudaStreamWaitEvent(s, pre->done, 0));
            tr.mark("rng(prefetched)");
            k_sc_reduce_wide<<<LAUNCH_1D(2 * n, 128), 0, s>>>((const uint32_t *)pre->d_raw.p, (uint32_t)n, d_sL, d_sR);
            KCHECK();
            CUDA_TRY(cudaMemsetAsync(pre->d_raw.p, 0, 128 * n, s));
            prefetch_discard(pre);
        } else if (n) {
            const size_t CHUNK = BPG_RNG_CHUNK; // draws per chunk (8 MB)
            if (!ctx->h_pinned) {
                CUDA_TRY(cudaHostAlloc(&ctx->h_pinned, 2 * CHUNK * 64, cudaHostAllocDefault));
                ctx->h_pinned_cap = 2 * CHUNK * 64;
                CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_stage[0], cudaEventDisableTiming | cudaEventBlockingSync));
                CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_stage[1], cudaEventDisableTiming | cudaEventBlockingSync));
            }
            CTX_TRY(ctx->scratch[12].ensure((4 * N + 8) * sizeof(sc))); // l1|r0|r1|r3 later; free until the commit MSMs are done
            uint8_t *d_raw = (uint8_t *)ctx->scratch[12].p;
            size_t total = 2 * n, done = 0;
            int used[2] = {0, 0};
            for (size_t c = 0; done < total; c++) {
                int b = (int)(c & 1);
                size_t cnt = std::min(CHUNK, total - done);
                uint8_t *hb = (uint8_t *)ctx->h_pinned + (size_t)b * CHUNK * 64;
                if (used[b]) CUDA_TRY(cudaEventSynchronize(ctx->ev_stage[b])); // the previous upload from this buffer has finished
                bpgh::RngService::get().draw64(rng, hb, cnt);
                CUDA_TRY(cudaMemcpyAsync(d_raw + 64 * done, hb, 64 * cnt, cudaMemcpyHostToDevice, s));
                CUDA_TRY(cudaEventRecord(ctx->ev_stage[b], s));
                used[b] = 1;
                done += cnt;
            }
            tr.mark("rng");
            k_sc_reduce_wide<<<LAUNCH_1D(2 * n, 128), 0, s>>>((const uint32_t *)d_raw, (uint32_t)n, d_sL, d_sR);
            KCHECK();
            SYNC_TRY(ctx, s); // the staging buffers are reused by this context's next proof
            explicit_bzero(ctx->h_pinned, std::min(ctx->h_pinned_cap, 64 * total)); // the raw draws are prover secrets (they determine s_L, s_R)
        }
    }
    memset(&plan, 0, sizeof plan); plan.lean = bpg_lean_now(); plan.shard = shard_on;
    plan.ngroups = 1;
    add_seg(d_sL, n, 0, 0); add_seg(d_sR, n, (uint32_t)ctx->cap, 0); add_seg(d_small + 2, 1, pBb, 0);
    CTX_TRY(msm_run(ctx, s, &plan, res + 2));
    CTX_TRY(run_compress(ctx, s, res, 3, d_enc));
    uint8_t AIe[32], AOe[32], Se[32], h_enc[96];
    D2H_TRY(ctx, h_enc, d_enc, 96, s);
    SYNC_TRY(ctx, s);
    tr.mark("commitMSMs");
    memcpy(AIe, h_enc, 32); memcpy(AOe, h_enc + 32, 32); memcpy(Se, h_enc + 64, 32);
    t.append("A_I1", AIe, 32); t.append("A_O1", AOe, 32); t.append("S1", Se, 32);
    t.append("dom-sep", (const uint8_t *)"r1cs-1phase", 11);
    uint8_t Z32[32] = {0};
    t.append("A_I2", Z32, 32); t.append("A_O2", Z32, 32); t.append("S2", Z32, 32);
    sc y = challenge_scalar(t, "y"), z = challenge_scalar(t, "z");
    sc yinv = h_inv(y);

    // ---- scalar-vector phase on the device
    size_t ptz = pow_tab_size((uint32_t)c->q + 1), pty = pow_tab_size((uint32_t)N);
    CTX_TRY(ctx->scratch[10].ensure((ptz + 2 * pty) * sizeof(sc)));
    CTX_TRY(ctx->scratch[11].ensure(((size_t)c->ncols + 1) * sizeof(sc)));
    CTX_TRY(ctx->scratch[12].ensure((4 * N + 8) * sizeof(sc)));
    CTX_TRY(ctx->scratch[13].ensure((4 * N + 8) * sizeof(sc)));
    sc h3[3] = {y, z, yinv};
    CUDA_TRY(cudaMemcpyAsync(d_small + 3, h3, sizeof h3, cudaMemcpyHostToDevice, s));
    pow_tab tz, ty, tyi;
    sc *d_pt = (sc *)ctx->scratch[10].p;
    CTX_TRY(make_pow_tables(ctx, s, d_small + 4, (uint32_t)c->q + 1, d_pt, tz));
    CTX_TRY(make_pow_tables(ctx, s, d_small + 3, (uint32_t)N, d_pt + ptz, ty));
    CTX_TRY(make_pow_tables(ctx, s, d_small + 5, (uint32_t)N, d_pt + ptz + pty, tyi));
    sc *d_w = (sc *)ctx->scratch[11].p;
    CTX_TRY(run_flatten(ctx, s, c, tz, d_w, ctx->scratch[15]));
    sc *d_l1 = (sc *)ctx->scratch[12].p, *d_r0 = d_l1 + N, *d_r1 = d_r0 + N, *d_r3 = d_r1 + N;
    sc *d_parts = (sc *)ctx->scratch[14].p;
    unsigned pb = (unsigned)std::min<size_t>((n + 127) / 128, 1184);
    sc h_t[6];
    memset(h_t, 0, sizeof h_t);
    std::vector<sc> h_wV(m ? m : 1);
    if (n) {
        k_poly_phase1<<<pb, 128, 0, s>>>((uint32_t)n, d_aL, d_aR, d_aO, d_sL, d_sR, d_w, ty.lo, ty.hi, tyi.lo, tyi.hi, d_l1, d_r0, d_r1, d_r3, d_parts);
        KCHECK();
        k_sum_partials<6><<<1, 128, 0, s>>>(d_parts, pb, d_small + 16);
        KCHECK();
        D2H_TRY(ctx, h_t, d_small + 16, sizeof h_t, s);
    }
    if (m) D2H_TRY(ctx, h_wV.data(), d_w + 3 * n, 32 * m, s);
    SYNC_TRY(ctx, s);
    tr.mark("flatten+poly1");
    sc tb[7];
    tb[1] = rng_scalar(rng); tb[3] = rng_scalar(rng); tb[4] = rng_scalar(rng); tb[5] = rng_scalar(rng); tb[6] = rng_scalar(rng);
    // T_1, T_3, T_4, T_5, T_6
    uint8_t tv8[5 * 32], tbv8[5 * 32], Te[5 * 32];
    const int TK[5] = {1, 3, 4, 5, 6};
    for (int k = 0; k < 5; k++) { sc_tobytes(tv8 + 32 * k, h_t[TK[k] - 1]); sc_tobytes(tbv8 + 32 * k, tb[TK[k]]); }
    CTX_TRY(bpg_pedersen_commit(ctx, tv8, tbv8, 5, Te));
    static const char *TL[5] = {"T_1", "T_3", "T_4", "T_5", "T_6"};
    for (int k = 0; k < 5; k++) t.append(TL[k], Te + 32 * k, 32);
    sc u = challenge_scalar(t, "u"), x = challenge_scalar(t, "x");
    sc tb2; sc_set_u32(tb2, 0);
    for (size_t i = 0; i < m; i++) { sc vb, r; sc_frombytes(vb, v_blinding + 32 * i); sc_reduce(r, vb); tb2 = h_add(tb2, h_mul(h_wV[i], r)); }
    tb[2] = tb2;
    sc xp[7];
    xp[0] = SC_ONE_H;
    for (int k = 1; k <= 6; k++) xp[k] = h_mul(xp[k - 1], x);
    sc tx, txb;
    sc_set_u32(tx, 0); sc_set_u32(txb, 0);
    for (int k = 1; k <= 6; k++) { tx = h_add(tx, h_mul(h_t[k - 1], xp[k])); txb = h_add(txb, h_mul(tb[k], xp[k])); }
    sc eb = h_mul(x, h_add(ib, h_mul(x, h_add(ob, h_mul(x, sb)))));
    append_scalar(t, "t_x", tx); append_scalar(t, "t_x_blinding", txb); append_scalar(t, "e_blinding", eb);
    sc w = challenge_scalar(t, "w");

    // ---- l(x), r(x) and the inner-product argument
    sc h4[5] = {xp[1], xp[2], xp[3], u, w};
    CUDA_TRY(cudaMemcpyAsync(d_small + 6, h4, sizeof h4, cudaMemcpyHostToDevice, s));
    sc *d_a = (sc *)ctx->scratch[13].p, *d_b = d_a + N, *d_EG = d_b + N, *d_EH = d_EG + N;
    k_poly_phase2<<<LAUNCH_1D(N, 128), 0, s>>>((uint32_t)n, (uint32_t)N, d_small + 6, d_l1, d_aO, d_sL, d_r0, d_r1, d_r3, ty.lo, ty.hi, tyi.lo, tyi.hi, d_a,
                                                d_b, d_EG, d_EH);
    KCHECK();
    tr.mark("T+poly2");
    t.append("dom-sep", (const uint8_t *)"ipp v1", 6);
    t.append_u64("n", N);
    sc *d_sG = (sc *)ctx->scratch[12].p, *d_sH = d_sG + N; // l1.. are dead now
    std::vector<uint8_t> LR(64 * (lgN ? lgN : 1));
    // Late fold (kernels_msm.cuh): rounds j < k0 run over the N original generators with expanded scalars; at round k0 the
    // n' = N >> k0 folded generators are materialised and the remaining rounds run over them (EG = EH = 1 again).
    int k0 = lgN; // no late fold
    bool late = (flags & BPG_FLAG_FORCE_LATE_FOLD) ? lgN >= 2 : (!(flags & BPG_FLAG_NO_LATE_FOLD) && lgN >= 15);
    // folded generators kept per vector: 512, or N / 128 for large circuits (round 1: N / 256) (measured at N = 2^16: 256 / 512 / 1024 / 2048 ->
    // 309 / 315 / 303 / 242 proofs/s; single proof at N = 2^20: 512 / 2048 / 4096 / 8192 -> 71.7 / 65.9 / 64.3 / 63.4 ms,
    // N = 2^19: 48.9 / 46.3 / 45.8 / 46.5 ms).  Round 2, after the full-warp k_mat_reduce and the block-per-output sort of the
    // materialisation: N / 128 (single proof at N = 2^20: 4096 / 8192 / 16384 -> 62.8 / 61.0 / 61.7 ms; 24 provers: 17.7 / 18.9 /
    // 18.6 proofs/s).  BPG_LATE_FOLD_LG overrides the log2 for experiments.
    static int late_lg_env = -1;
    if (late_lg_env < 0) { const char *e = getenv("BPG_LATE_FOLD_LG"); late_lg_env = e ? atoi(e) : 0; if (late_lg_env < 1 || late_lg_env > 14) late_lg_env = 0; }
    // (a proof sharded over several ranks replicates the reduction and the table build of the folded generators on every rank:
    // one halving earlier there -- 8 GPUs, N = 2^20: 19.9 ms with N / 256 against 21.4 ms with N / 128)
    int late_lg = late_lg_env ? late_lg_env : std::max(9, lgN - 7 - (shard_on ? 1 : 0));
    if (shard_on) while (late_lg > 1 && ((size_t)2 << late_lg) * sizeof(ge) > ctx->shard_cap) late_lg--; // partial outputs must fit one exchange
    if (late) k0 = std::max(1, lgN - late_lg);
    size_t Ncur = N;                       // generators of the current basis
    uint32_t pG = 0, pH = (uint32_t)ctx->cap, pQ = pB; // point indices of G_0, H_0 and B in the current tables
    const ge_an *tabcur = nullptr;
    uint32_t ptcur = 0;
    for (int j = 0; j < lgN; j++) {
        uint32_t nj = (uint32_t)(N >> j), h = nj >> 1;
        if (j == k0) {
            CTX_TRY(msm_materialise_fold(ctx, s, (uint32_t)N, nj, d_EG, d_EH, bpg_lean_now(), shard_on));
            Ncur = nj;
            tabcur = (const ge_an *)ctx->mat_tab.p; ptcur = 2 * nj + 2;
            pG = 0; pH = nj; pQ = 2 * nj;
            k_sc_fill_one<<<LAUNCH_1D(Ncur, 128), 0, s>>>(d_EG, (uint32_t)Ncur);
            KCHECK();
            k_sc_fill_one<<<LAUNCH_1D(Ncur, 128), 0, s>>>(d_EH, (uint32_t)Ncur);
            KCHECK();
        }
        unsigned cb = (unsigned)std::min<size_t>((h + 127) / 128, 1184);
        k_ipp_cross<<<cb, 128, 0, s>>>(h, d_a, d_b, d_parts);
        KCHECK();
        k_sum_partials<2><<<1, 128, 0, s>>>(d_parts, cb, d_small + 24);
        KCHECK();
        k_ipp_cw<<<1, 32, 0, s>>>(d_small + 24, d_small + 10, d_small + 13);
        KCHECK();
        // a rank of a sharded proof only needs the expanded scalars (and the per-generator factors behind them) of its own
        // point range of the original generators; after the late fold the vectors are tiny and every rank keeps all of them
        uint32_t e0 = 0, e1 = (uint32_t)Ncur;
        if (shard_on && j < k0) shard_slice((uint32_t)Ncur, ctx->shard_rank, ctx->shard_world, e0, e1);
        if (e1 > e0) {
            k_ipp_expand<<<LAUNCH_1D(e1 - e0, 128), 0, s>>>((uint32_t)Ncur, nj, d_a, d_b, d_EG, d_EH, d_sG, d_sH, e0, e1);
            KCHECK();
        }
        memset(&plan, 0, sizeof plan); plan.lean = bpg_lean_now(); plan.shard = shard_on;
        plan.ngroups = 2;
        plan.tab = tabcur; plan.ptotal = ptcur;
        add_seg(d_sG, Ncur, pG, 1); plan.seg[plan.nseg - 1].alt = 1 + (uint32_t)__builtin_ctz(h); // G_i: right half -> L (group 0)
        add_seg(d_sH, Ncur, pH, 0); plan.seg[plan.nseg - 1].alt = 1 + (uint32_t)__builtin_ctz(h); // H_i: right half -> R (group 1)
        add_seg(d_small + 13, 1, pQ, 0);
        add_seg(d_small + 14, 1, pQ, 1);
        CTX_TRY(msm_run(ctx, s, &plan, res));
        CTX_TRY(run_compress(ctx, s, res, 2, d_enc));
        D2H_TRY(ctx, LR.data() + 64 * j, d_enc, 64, s);
        SYNC_TRY(ctx, s);
        t.append("L", LR.data() + 64 * j, 32);
        t.append("R", LR.data() + 64 * j + 32, 32);
        sc uj = challenge_scalar(t, "u");
        sc uu[2] = {uj, h_inv(uj)};
        CUDA_TRY(cudaMemcpyAsync(d_small + 11, uu, sizeof uu, cudaMemcpyHostToDevice, s));
        // (a, b fold in full on every rank; grid = what is needed: the rank's factor range and the lower half of a, b)
        k_ipp_fold<<<LAUNCH_1D(std::max<size_t>(e1, h), 128), 0, s>>>((uint32_t)Ncur, nj, d_small + 11, d_a, d_b, d_EG, d_EH, e0, e1);
        KCHECK();
    }
    tr.mark("ipp");
    sc fab[2];
    D2H_TRY(ctx, &fab[0], d_a, 32, s);
    D2H_TRY(ctx, &fab[1], d_b, 32, s);
    SYNC_TRY(ctx, s);

    // ---- R1CSProof::to_bytes
    uint8_t *o = proof;
    if (legacy) { memcpy(o, AIe, 32); memcpy(o + 32, AOe, 32); memcpy(o + 64, Se, 32); memset(o + 96, 0, 96); o += 192; }
    else { *o++ = 0; memcpy(o, AIe, 32); memcpy(o + 32, AOe, 32); memcpy(o + 64, Se, 32); o += 96; }
    memcpy(o, Te, 160); o += 160;
    sc_tobytes(o, tx); sc_tobytes(o + 32, txb); sc_tobytes(o + 64, eb); o += 96;
    memcpy(o, LR.data(), 64 * (size_t)lgN); o += 64 * (size_t)lgN;
    sc_tobytes(o, fab[0]); sc_tobytes(o + 32, fab[1]); o += 64;
    // Prover secrets do not outlive the call: witness, s_L / s_R, l(x) / r(x) vectors, blinding scalars and seeds on the device
    // (asynchronous, ordered before this context's next use of the buffers), blindings and the RNG state on the host.
    CUDA_TRY(cudaMemsetAsync(ctx->scratch[8].p, 0, (5 * N + 8) * sizeof(sc), s));
    CUDA_TRY(cudaMemsetAsync(ctx->scratch[12].p, 0, (4 * N + 8) * sizeof(sc), s));
    CUDA_TRY(cudaMemsetAsync(ctx->scratch[13].p, 0, (4 * N + 8) * sizeof(sc), s));
    CUDA_TRY(cudaMemsetAsync(d_small, 0, 256 * sizeof(sc), s));
    explicit_bzero(&rng, sizeof rng); explicit_bzero(tb, sizeof tb); explicit_bzero(h_small, sizeof h_small);
    explicit_bzero(&ib, sizeof ib); explicit_bzero(&ob, sizeof ob); explicit_bzero(&sb, sizeof sb);
    return (long)(o - proof);
}

// ================================================================ Verifier::verify
// Stage 1 (per proof): R1CSProof::from_bytes, transcript replay, device-side scalar preparation.  Leaves the G/H scalars
// g_i, h_i of the verification equation on the device (d_gh if given, else a scratch buffer) and the scalars / encodings
// of the proof's own points, B and B~ on the host.  status = 0: the proof is rejected already (format / transcript).
struct vprep {
    int status = 0;
    size_t N = 0, k = 0;
    sc sB, sBb;
    std::vector<uint8_t> es, ep; // k scalars, k compressed points
    sc *d_g = nullptr, *d_h = nullptr;
    // state carried from the enqueue half to the completion half (after the device results have been read back)
    size_t n = 0, m = 0, lg = 0;
    sc x, u, r, w, tx, txb, eb, ia, ibb, usq[32], uisq[32];
    const uint8_t *A[6] = {nullptr}, *Tp = nullptr, *LR = nullptr, *V32 = nullptr;
    sc *d_slot = nullptr; // device: delta | wV[0..m) | wc
};
static const uint8_t BPG_ZERO32[32] = {0};
static int verify_prepare(bpg_ctx *ctx, bpg_circuit *c, const uint8_t *label, size_t label_len, const uint8_t *V32, const uint8_t *proof,
                          size_t proof_len, const uint8_t ext_rng32[32], unsigned flags, vprep &out, sc *d_gh, sc *d_slot) {
    if (!ctx || !c || !label || !proof || !ext_rng32) return BPG_E_ARG;
    out.status = 0;
    size_t n = c->n, m = c->m;
    if (m && !V32) return BPG_E_ARG;
    // ---- R1CSProof::from_bytes (FormatError -> reject)
    const uint8_t *A[6];
    const uint8_t *Z32 = BPG_ZERO32;
    const uint8_t *f;
    size_t nf;
    if (flags & BPG_FLAG_LEGACY_FRAMING) {
        if (proof_len % 32 || proof_len < 14 * 32) return BPG_OK;
        for (int i = 0; i < 6; i++) A[i] = proof + 32 * i;
        f = proof + 192; nf = proof_len / 32 - 6;
    } else {
        if (proof_len == 0) return BPG_OK;
        int ver = proof[0];
        size_t body = proof_len - 1;
        if (body % 32 || (ver != 0 && ver != 1)) return BPG_OK;
        if (body < (size_t)(ver == 0 ? 11 : 14) * 32) return BPG_OK;
        if (ver == 0) { for (int i = 0; i < 3; i++) { A[i] = proof + 1 + 32 * i; A[3 + i] = Z32; } f = proof + 97; nf = body / 32 - 3; }
        else { for (int i = 0; i < 6; i++) A[i] = proof + 1 + 32 * i; f = proof + 193; nf = body / 32 - 6; }
    }
    if (nf < 10 || (nf - 10) % 2) return BPG_OK;
    const uint8_t *Tp = f;
    sc tx, txb, eb, ia, ibb;
    if (!sc_canonical_bytes(tx, f + 160) || !sc_canonical_bytes(txb, f + 192) || !sc_canonical_bytes(eb, f + 224)) return BPG_OK;
    size_t lg = (nf - 10) / 2;
    if (lg >= 32) return BPG_OK;
    const uint8_t *LR = f + 256;
    if (!sc_canonical_bytes(ia, LR + 64 * lg) || !sc_canonical_bytes(ibb, LR + 64 * lg + 32)) return BPG_OK;

    bpgh::Transcript t(label, label_len);
    t.append("dom-sep", (const uint8_t *)"r1cs v1", 7);
    for (size_t i = 0; i < m; i++) t.append("V", V32 + 32 * i, 32);
    t.append_u64("m", m);
    static const char *AL[6] = {"A_I1", "A_O1", "S1", "A_I2", "A_O2", "S2"};
    for (int i = 0; i < 3; i++) if (!t.validate_and_append_point(AL[i], A[i])) return BPG_OK;
    t.append("dom-sep", (const uint8_t *)"r1cs-1phase", 11);
    size_t N = next_pow2(n);
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (ctx->cap < N) return BPG_OK; // InvalidGeneratorsLength -> Err
    for (int i = 3; i < 6; i++) t.append(AL[i], A[i], 32);
    sc y = challenge_scalar(t, "y"), z = challenge_scalar(t, "z");
    static const char *TL[5] = {"T_1", "T_3", "T_4", "T_5", "T_6"};
    for (int k = 0; k < 5; k++) if (!t.validate_and_append_point(TL[k], Tp + 32 * k)) return BPG_OK;
    sc u = challenge_scalar(t, "u"), x = challenge_scalar(t, "x");
    append_scalar(t, "t_x", tx); append_scalar(t, "t_x_blinding", txb); append_scalar(t, "e_blinding", eb);
    sc w = challenge_scalar(t, "w");
    if (N != ((size_t)1 << lg)) return BPG_OK;
    t.append("dom-sep", (const uint8_t *)"ipp v1", 6);
    t.append_u64("n", N);
    sc ch[32], usq[32], uisq[32], allinv = SC_ONE_H;
    for (size_t j = 0; j < lg; j++) {
        if (!t.validate_and_append_point("L", LR + 64 * j)) return BPG_OK;
        if (!t.validate_and_append_point("R", LR + 64 * j + 32)) return BPG_OK;
        ch[j] = challenge_scalar(t, "u");
    }
    { // batch inversion of the lg challenges
        sc prod = SC_ONE_H, pre[33];
        for (size_t j = 0; j < lg; j++) { pre[j] = prod; prod = h_mul(prod, ch[j]); }
        sc inv = h_inv(prod);
        allinv = inv;
        for (size_t j = lg; j-- > 0;) { sc cj = h_mul(inv, pre[j]); inv = h_mul(inv, ch[j]); usq[j] = h_mul(ch[j], ch[j]); uisq[j] = h_mul(cj, cj); }
    }
    sc yinv = h_inv(y);
    bpgh::TranscriptRng rng(t);
    rng.finalize(ext_rng32);
    sc r = rng_scalar(rng);

    cudaStream_t s = ctx->stream;
    CTX_TRY(ctx->scratch[9].ensure(256 * sizeof(sc)));
    sc *d_small = (sc *)ctx->scratch[9].p;
    sc hs[8 + 32];
    hs[0] = z; hs[1] = yinv; hs[2] = allinv; hs[3] = x; hs[4] = ia; hs[5] = ibb; hs[6] = u;
    for (size_t j = 0; j < lg; j++) hs[8 + j] = usq[j];
    CUDA_TRY(cudaMemcpyAsync(d_small, hs, sizeof(sc) * (8 + lg), cudaMemcpyHostToDevice, s));
    size_t ptz = pow_tab_size((uint32_t)c->q + 1), pty = pow_tab_size((uint32_t)N);
    CTX_TRY(ctx->scratch[10].ensure((ptz + 2 * pty) * sizeof(sc)));
    CTX_TRY(ctx->scratch[11].ensure(((size_t)c->ncols + 1) * sizeof(sc)));
    if (!d_gh) { CTX_TRY(ctx->scratch[12].ensure((2 * N + 8) * sizeof(sc))); d_gh = (sc *)ctx->scratch[12].p; }
    CTX_TRY(ctx->scratch[14].ensure(4096 * 6 * sizeof(sc)));
    pow_tab tz, tyi, ts;
    sc *d_pt = (sc *)ctx->scratch[10].p;
    CTX_TRY(make_pow_tables(ctx, s, d_small + 0, (uint32_t)c->q + 1, d_pt, tz));
    CTX_TRY(make_pow_tables(ctx, s, d_small + 1, (uint32_t)N, d_pt + ptz, tyi));
    ts.lo = d_pt + ptz + pty; ts.hi = ts.lo + 1024;
    uint32_t nshi = (uint32_t)(N >> 10) + 2;
    k_s_tables<<<LAUNCH_1D(1024 + nshi, 128), 0, s>>>(d_small + 8, d_small + 2, (uint32_t)lg, ts.lo, ts.hi, nshi);
    KCHECK();
    sc *d_w = (sc *)ctx->scratch[11].p;
    CTX_TRY(run_flatten(ctx, s, c, tz, d_w, ctx->scratch[15]));
    sc *d_g = d_gh, *d_h = d_g + N, *d_parts = (sc *)ctx->scratch[14].p;
    unsigned vb = (unsigned)std::min<size_t>((N + 127) / 128, 1184);
    // vs = [x, a, b, u] at d_small + 3
    k_verify_scalars<<<vb, 128, 0, s>>>((uint32_t)n, (uint32_t)N, d_small + 3, d_w, tyi.lo, tyi.hi, ts.lo, ts.hi, d_g, d_h, d_parts);
    KCHECK();
    k_sum_partials<1><<<1, 128, 0, s>>>(d_parts, vb, d_slot);
    KCHECK();
    CUDA_TRY(cudaMemcpyAsync(d_slot + 1, d_w + 3 * n, 32 * (m + 1), cudaMemcpyDeviceToDevice, s)); // wV[0..m) | wc
    // the completion half needs these once delta, wV, wc have been read back
    out.n = n; out.m = m; out.lg = lg; out.x = x; out.u = u; out.r = r; out.w = w; out.tx = tx; out.txb = txb; out.eb = eb;
    out.ia = ia; out.ibb = ibb;
    for (size_t j = 0; j < lg; j++) { out.usq[j] = usq[j]; out.uisq[j] = uisq[j]; }
    for (int i = 0; i < 6; i++) out.A[i] = A[i];
    out.Tp = Tp; out.LR = LR; out.V32 = V32; out.d_slot = d_slot;
    out.N = N; out.k = 6 + m + 5 + 2 * lg; out.d_g = d_g; out.d_h = d_h;
    out.status = 1;
    return BPG_OK;
}
// completion half: scalars / points of the proof's own terms from the read-back slot delta | wV | wc
static void verify_complete(vprep &out, const sc *h_slot) {
    const sc &x = out.x, &u = out.u, &r = out.r, &w = out.w;
    size_t m = out.m, lg = out.lg, k = out.k;
    const sc &delta = h_slot[0], &wc = h_slot[1 + m];
    const sc *h_wV = h_slot + 1;
    const uint8_t *const *A = out.A;
    const uint8_t *Tp = out.Tp, *LR = out.LR, *V32 = out.V32;
    const sc &tx = out.tx, &txb = out.txb, &eb = out.eb, &ia = out.ia, &ibb = out.ibb;
    const sc *usq = out.usq, *uisq = out.uisq;
    sc xx = h_mul(x, x), xxx = h_mul(xx, x), rxx = h_mul(r, xx);
    std::vector<uint8_t> es(32 * k), ep(32 * k);
    size_t ci = 0;
    auto put = [&](const sc &sv, const uint8_t *pt) { sc_tobytes(es.data() + 32 * ci, sv); memcpy(ep.data() + 32 * ci, pt, 32); ci++; };
    put(x, A[0]); put(xx, A[1]); put(xxx, A[2]); put(h_mul(u, x), A[3]); put(h_mul(u, xx), A[4]); put(h_mul(u, xxx), A[5]);
    for (size_t j = 0; j < m; j++) put(h_mul(h_wV[j], rxx), V32 + 32 * j);
    put(h_mul(r, x), Tp); put(h_mul(rxx, x), Tp + 32); put(h_mul(rxx, xx), Tp + 64); put(h_mul(rxx, xxx), Tp + 96); put(h_mul(h_mul(rxx, xx), xx), Tp + 128);
    for (size_t j = 0; j < lg; j++) put(usq[j], LR + 64 * j);
    for (size_t j = 0; j < lg; j++) put(uisq[j], LR + 64 * j + 32);
    sc sB = h_add(h_mul(w, h_sub(tx, h_mul(ia, ibb))), h_mul(r, h_sub(h_mul(xx, h_add(wc, delta)), tx)));
    sc sBb; sc_neg_r(sBb, h_add(eb, h_mul(r, txb)));
    out.sB = sB; out.sBb = sBb;
    out.es.swap(es); out.ep.swap(ep);
}

// Stage 2: one check  sum_i rho_i * (verification equation of proof i) == identity  over a set of prepared proofs.
// One proof with rho = 1 is exactly Verifier::verify's single multiscalar multiplication.  The fixed-base part runs on the
// resident tables (scalars of all proofs are accumulated per generator first), the proofs' own points are decompressed
// and multiplied on the second stream.
static int verify_finish(bpg_ctx *ctx, const std::vector<vprep *> &S, const std::vector<sc> &rho, int *pass) {
    *pass = 0;
    cudaStream_t s = ctx->stream, s2 = ctx->stream2;
    size_t Nmax = 0, ktot = 0;
    for (vprep *p : S) { Nmax = std::max(Nmax, p->N); ktot += p->k; }
    bool single = S.size() == 1 && rho[0].v[0] == 1 && !(rho[0].v[1] | rho[0].v[2] | rho[0].v[3] | rho[0].v[4] | rho[0].v[5] | rho[0].v[6] | rho[0].v[7]);
    CTX_TRY(ctx->scratch[9].ensure(256 * sizeof(sc)));
    sc *d_small = (sc *)ctx->scratch[9].p;
    const sc *d_g, *d_h;
    sc sB, sBb;
    sc_set_u32(sB, 0); sc_set_u32(sBb, 0);
    std::vector<uint8_t> es(32 * ktot), ep(32 * ktot);
    size_t off = 0;
    if (single) {
        d_g = S[0]->d_g; d_h = S[0]->d_h;
        sB = S[0]->sB; sBb = S[0]->sBb;
        memcpy(es.data(), S[0]->es.data(), 32 * ktot); memcpy(ep.data(), S[0]->ep.data(), 32 * ktot);
    } else {
        CTX_TRY(ctx->scratch[13].ensure((2 * Nmax + 8) * sizeof(sc)));
        sc *acc = (sc *)ctx->scratch[13].p;
        CUDA_TRY(cudaMemsetAsync(acc, 0, 2 * Nmax * sizeof(sc), s));
        for (size_t i = 0; i < S.size(); i++) {
            vprep *p = S[i];
            k_axpy_gh<<<LAUNCH_1D(p->N, 128), 0, s>>>(rho[i], p->d_g, p->d_h, (uint32_t)p->N, acc, acc + Nmax);
            KCHECK();
            sB = h_add(sB, h_mul(rho[i], p->sB));
            sBb = h_add(sBb, h_mul(rho[i], p->sBb));
            for (size_t j = 0; j < p->k; j++) {
                sc v; sc_frombytes(v, p->es.data() + 32 * j);
                sc_tobytes(es.data() + 32 * (off + j), h_mul(rho[i], v));
            }
            memcpy(ep.data() + 32 * off, p->ep.data(), 32 * p->k);
            off += p->k;
        }
        d_g = acc; d_h = acc + Nmax;
    }
    size_t k = ktot, N = Nmax;
    sc hb[2] = {sB, sBb};
    CUDA_TRY(cudaMemcpyAsync(d_small + 50, hb, sizeof hb, cudaMemcpyHostToDevice, s));
    CTX_TRY(ctx->scratch[2].ensure(64 * k + 64));
    uint8_t *d_e = (uint8_t *)ctx->scratch[2].p;
    CTX_TRY(ctx->results.ensure(8 * sizeof(ge) + 64));
    ge *res = (ge *)ctx->results.p;
    uint32_t *d_ok = (uint32_t *)(res + 8);
    uint32_t one = 1, ok = 1;
    // few own points: per-term scalar multiplications on the second stream, concurrent with the fixed-base MSM;
    // many (thousands of commitments, batches): the bucket engine, which shares this context's MSM workspace -> same stream
    cudaStream_t vs = k >= BPG_VARBASE_BUCKET_TERMS ? s : s2;
    CUDA_TRY(cudaMemcpyAsync(d_e, es.data(), 32 * k, cudaMemcpyHostToDevice, vs));
    CUDA_TRY(cudaMemcpyAsync(d_e + 32 * k, ep.data(), 32 * k, cudaMemcpyHostToDevice, vs));
    CUDA_TRY(cudaMemcpyAsync(d_ok, &one, 4, cudaMemcpyHostToDevice, vs));
    CTX_TRY(varbase_msm_dev(ctx, vs, d_e, d_e + 32 * k, k, res + 1, d_ok, ctx->scratch[3], ctx->scratch[4]));
    CUDA_TRY(cudaEventRecord(ctx->ev2, vs));
    const uint32_t pB = (uint32_t)(2 * ctx->cap);
    msm_plan plan;
    memset(&plan, 0, sizeof plan); plan.lean = bpg_lean_now(); plan.shard = 0;
    plan.ngroups = 1;
    auto add_seg = [&](const sc *sp, size_t cnt, uint32_t p0) {
        msm_seg &sg = plan.seg[plan.nseg++];
        sg.scalars = sp; sg.n = (uint32_t)cnt; sg.p0 = p0; sg.group = 0; sg.reduce = 0;
    };
    add_seg(d_g, N, 0); add_seg(d_h, N, (uint32_t)ctx->cap); add_seg(d_small + 50, 1, pB); add_seg(d_small + 51, 1, pB + 1);
    CTX_TRY(msm_run(ctx, s, &plan, res));
    CUDA_TRY(cudaStreamWaitEvent(s, ctx->ev2, 0));
    k_points_sum_kernel<<<1, 64, 0, s>>>(res, 2, res + 2);
    KCHECK();
    uint8_t *d_enc = (uint8_t *)(res + 4), enc[32];
    CTX_TRY(run_compress(ctx, s, res + 2, 1, d_enc));
    D2H_TRY(ctx, enc, d_enc, 32, s);
    D2H_TRY(ctx, &ok, d_ok, 4, s);
    SYNC_TRY(ctx, s);
    uint8_t nz = 0;
    for (int i = 0; i < 32; i++) nz |= enc[i];
    *pass = (ok && nz == 0) ? 1 : 0; // identity coset <=> all-zero encoding
    return BPG_OK;
}

extern "C" int bpg_r1cs_verify(bpg_ctx *ctx, bpg_circuit *c, const uint8_t *label, size_t label_len, const uint8_t *V32, const uint8_t *proof,
                               size_t proof_len, const uint8_t ext_rng32[32], unsigned flags, int *accept) {
    bpg_inflight_guard inflight_;
    if (!accept) return BPG_E_ARG;
    *accept = 0;
    if (!ctx || !c || !label || !proof || !ext_rng32) return BPG_E_ARG;
    if (c->m && !V32) return BPG_E_ARG;
    vprep p;
    CUDA_TRY(cudaSetDevice(ctx->device));
    CTX_TRY(ctx->scratch[7].ensure((c->m + 4) * sizeof(sc)));
    CTX_TRY(verify_prepare(ctx, c, label, label_len, V32, proof, proof_len, ext_rng32, flags, p, nullptr, (sc *)ctx->scratch[7].p));
    if (!p.status) return BPG_OK;
    std::vector<sc> h_slot(c->m + 2);
    D2H_TRY(ctx, h_slot.data(), p.d_slot, 32 * (c->m + 2), ctx->stream);
    SYNC_TRY(ctx, ctx->stream);
    verify_complete(p, h_slot.data());
    std::vector<vprep *> S = {&p};
    std::vector<sc> rho = {SC_ONE_H};
    return verify_finish(ctx, S, rho, accept);
}

// Batch verification: accept[i] is what bpg_r1cs_verify would return for proof i.  All proofs are prepared,
// then ONE combined check with random weights rho_i (derived from every ext_rng32) decides the whole batch; if it fails,
// the set is bisected until the failing proofs are isolated (a single proof is checked with rho = 1, i.e. exactly as by
// bpg_r1cs_verify), so verdicts are identical to one-by-one verification up to the 2^-250 soundness error of the weights.
static int verify_bisect(bpg_ctx *ctx, std::vector<vprep> &preps, const std::vector<sc> &rhos, const std::vector<size_t> &idx, size_t lo, size_t hi,
                         bool known_bad, int *accept) {
    if (lo >= hi) return BPG_OK;
    int pass = 0;
    if (!known_bad || hi - lo == 1) {
        std::vector<vprep *> S;
        std::vector<sc> rho;
        for (size_t t = lo; t < hi; t++) { S.push_back(&preps[idx[t]]); rho.push_back(hi - lo == 1 ? SC_ONE_H : rhos[idx[t]]); }
        CTX_TRY(verify_finish(ctx, S, rho, &pass));
        if (pass) { for (size_t t = lo; t < hi; t++) accept[idx[t]] = 1; return BPG_OK; }
        if (hi - lo == 1) { accept[idx[lo]] = 0; return BPG_OK; }
    }
    size_t mid = lo + (hi - lo) / 2;
    // the set [lo, hi) is known to fail: check the left half; if it passes, the right half must contain the failure
    std::vector<vprep *> S;
    std::vector<sc> rho;
    for (size_t t = lo; t < mid; t++) { S.push_back(&preps[idx[t]]); rho.push_back(mid - lo == 1 ? SC_ONE_H : rhos[idx[t]]); }
    CTX_TRY(verify_finish(ctx, S, rho, &pass));
    if (pass) {
        for (size_t t = lo; t < mid; t++) accept[idx[t]] = 1;
        return verify_bisect(ctx, preps, rhos, idx, mid, hi, true, accept);
    }
    if (mid - lo == 1) accept[idx[lo]] = 0;
    else CTX_TRY(verify_bisect(ctx, preps, rhos, idx, lo, mid, true, accept));
    return verify_bisect(ctx, preps, rhos, idx, mid, hi, false, accept);
}

extern "C" int bpg_r1cs_verify_batch(bpg_ctx *ctx, size_t count, bpg_circuit *const *circuits, const uint8_t *const *labels, const size_t *label_lens,
                                     const uint8_t *const *V32, const uint8_t *const *proofs, const size_t *proof_lens, const uint8_t *ext_rng32,
                                     unsigned flags, int *accept) {
    bpg_inflight_guard inflight_;
    if (!ctx || (count && (!circuits || !labels || !label_lens || !V32 || !proofs || !proof_lens || !ext_rng32 || !accept))) return BPG_E_ARG;
    if (!count) return BPG_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    for (size_t i = 0; i < count; i++) {
        accept[i] = 0;
        if (!circuits[i] || !labels[i] || !proofs[i] || (circuits[i]->m && !V32[i])) return BPG_E_ARG;
    }
    // device storage for every proof's g_i | h_i
    size_t total = 0, slots = 0;
    std::vector<size_t> goff(count), soff(count);
    for (size_t i = 0; i < count; i++) {
        goff[i] = total; total += 2 * next_pow2(circuits[i]->n) + 8;
        soff[i] = slots; slots += circuits[i]->m + 2;
    }
    CTX_TRY(ctx->batch_gh.ensure((total + slots + 8) * sizeof(sc)));
    sc *d_all = (sc *)ctx->batch_gh.p, *d_slots = d_all + total;
    std::vector<vprep> preps(count);
    std::vector<size_t> idx;
    // every proof's transcript replay (host) and scalar preparation (device) is enqueued back to back; one read-back follows
    for (size_t i = 0; i < count; i++) {
        CTX_TRY(verify_prepare(ctx, circuits[i], labels[i], label_lens[i], V32[i], proofs[i], proof_lens[i], ext_rng32 + 32 * i, flags, preps[i], d_all + goff[i],
                               d_slots + soff[i]));
        if (preps[i].status) idx.push_back(i);
    }
    std::vector<sc> h_slots(slots + 1);
    D2H_TRY(ctx, h_slots.data(), d_slots, 32 * slots, ctx->stream);
    SYNC_TRY(ctx, ctx->stream);
    for (size_t i : idx) verify_complete(preps[i], h_slots.data() + soff[i]);
    // Weights rho_i = wide_reduce(64-byte block i of SHAKE256("bpg batch v2" || count || for every proof: label, V, proof bytes
    // (length-prefixed) || all ext_rng32)).  They are bound to the STATEMENTS and PROOFS of the batch, not only to the caller's
    // randomness: the final IPP scalars a, b of a proof are in no transcript, so weights that ignore the proof bytes would let a
    // prover who can predict ext_rng32 submit two tampered copies whose errors cancel (rho_1 d_1 + rho_2 d_2 = 0).  With the
    // proofs absorbed the weights are a random-oracle output the prover cannot choose its proofs against; ext_rng32 (fresh
    // secret randomness, the reference's thread_rng) additionally hides them from an offline search.
    std::vector<sc> rhos(count);
    {
        bpgh::Sponge sp(136);
        auto absorb_u64 = [&](uint64_t v) { uint8_t b8[8]; for (int b = 0; b < 8; b++) b8[b] = (uint8_t)(v >> (8 * b)); sp.absorb(b8, 8); };
        sp.absorb((const uint8_t *)"bpg batch v2", 12);
        absorb_u64(count);
        for (size_t i = 0; i < count; i++) {
            absorb_u64(label_lens[i]); sp.absorb(labels[i], label_lens[i]);
            absorb_u64(circuits[i]->m); if (circuits[i]->m) sp.absorb(V32[i], 32 * circuits[i]->m);
            absorb_u64(proof_lens[i]); sp.absorb(proofs[i], proof_lens[i]);
        }
        sp.absorb(ext_rng32, 32 * count);
        sp.finish(0x1F);
        for (size_t i = 0; i < count; i++) {
            uint8_t w[64];
            sp.squeeze(w, 64);
            rhos[i] = h_wide(w);
        }
    }
    return verify_bisect(ctx, preps, rhos, idx, 0, idx.size(), false, accept);
}
