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
    CUDA_TRY(cudaMemcpyToSymbol(c_mimc, c.data(), sizeof(sc) * MIMC_ROUNDS));
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
    k_mimc_sponge<<<LAUNCH_1D(n, 128), 0, s>>>((const sc *)din, (const uint32_t *)(din + 32 * nblocks), (uint32_t)n, (sc *)dout,
                                                trace ? (sc *)(dout + 32 * n) : nullptr);
    KCHECK();
    CUDA_TRY(cudaMemcpyAsync(out32, dout, 32 * n, cudaMemcpyDeviceToHost, s));
    if (trace) CUDA_TRY(cudaMemcpyAsync(trace, dout + 32 * n, tbytes, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
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

// ================================================================ R1CS (filled in below)
struct bpg_circuit { int dummy; };
extern "C" int bpg_circuit_create(bpg_ctx *, size_t, size_t, size_t, const uint32_t *, const uint32_t *, const uint8_t *, bpg_circuit **) { return BPG_E_ARG; }
extern "C" void bpg_circuit_destroy(bpg_circuit *) {}
extern "C" long bpg_r1cs_prove(bpg_ctx *, bpg_circuit *, const uint8_t *, size_t, const uint8_t *, const uint8_t *, const uint8_t *, const uint8_t *,
                               const uint8_t *, const uint8_t *, unsigned, uint8_t *, uint8_t *, size_t) { return BPG_E_ARG; }
extern "C" int bpg_r1cs_verify(bpg_ctx *, bpg_circuit *, const uint8_t *, size_t, const uint8_t *, const uint8_t *, size_t, const uint8_t *, unsigned, int *) { return BPG_E_ARG; }
