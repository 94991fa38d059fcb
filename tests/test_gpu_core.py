"""GPU parity tests (through the C ABI of libbpg.so) against the CPU oracle: generators, Pedersen commitments,
fixed-base MSM over the resident tables, variable-base MSM, generator fold."""
import random

import pytest

import oracle_lib as ol
from oracle import pyref as pr

pytestmark = pytest.mark.gpu
L = pr.L


@pytest.fixture(scope="module")
def ctx():
    import bulletproofs_gadgets_b200 as bpg
    c = bpg.Context(0)
    c.gens_ensure(4096)
    yield c
    c.close()


def rs(rnd, bits=255):
    return rnd.randrange(2 ** bits).to_bytes(32, "little")


def test_imad_microbench_runs(ctx):
    ms, mac = ctx.bench_imad(200)
    assert ms > 0 and mac > 0
    print("fe_mul chain: %.3f ms, %.2f T MAC32/s" % (ms, mac / ms / 1e9))


def test_generators_match_oracle(ctx):
    B, Bb = ctx.pedersen_gens()
    assert (B, Bb) == ol.pedersen_gens()
    for i0, n in ((0, 64), (1000, 7), (4096 - 5, 5)):
        assert ctx.gens_export(i0, n) == ol.gens(i0, n)


def test_pedersen_commit_matches_oracle(ctx):
    rnd = random.Random(1)
    edge = [0, 1, L - 1, L, L + 1, 2 ** 255 - 1, 2 ** 252, 128, 127, 2 ** 128]
    vals = [e.to_bytes(32, "little") for e in edge] + [rs(rnd) for _ in range(300)]
    v = b"".join(vals)
    r = b"".join(reversed(vals))
    assert ctx.pedersen_commit(v, r) == ol.pedersen_commit(v, r)
    z = bytes(32)
    assert ctx.pedersen_commit(z, z) == bytes(32)
    # > 512 commitments: the one-thread-per-commitment kernel (<= 512 use one warp each)
    big = [rs(rnd) for _ in range(700)] + [e.to_bytes(32, "little") for e in edge]
    v, r = b"".join(big), b"".join(reversed(big))
    assert ctx.pedersen_commit(v, r) == ol.pedersen_commit(v, r)


@pytest.mark.parametrize("n,dist", [(1, "uniform"), (2, "uniform"), (63, "uniform"), (64, "uniform"), (65, "uniform"), (1000, "uniform"),
                                    (4096, "uniform"), (4096, "bits"), (4096, "small"), (3000, "same"), (100, "zero"), (777, "unreduced")])
def test_msm_gens_matches_oracle(ctx, n, dist):
    rnd = random.Random(n * 7 + len(dist))

    def draw():
        if dist == "uniform":
            return (rnd.randrange(L)).to_bytes(32, "little")
        if dist == "bits":
            return rnd.randrange(2).to_bytes(32, "little")
        if dist == "small":
            return rnd.randrange(2 ** 64).to_bytes(32, "little")
        if dist == "same":
            return (L - 5).to_bytes(32, "little")
        if dist == "zero":
            return bytes(32)
        return rs(rnd, 256)

    sG = b"".join(draw() for _ in range(n))
    sH = b"".join(draw() for _ in range(n))
    off = 4096 - n if n < 4096 else 0
    assert ctx.msm_gens(sG, sH, n, off) == ol.msm_gens(sG, sH, n, off)
    assert ctx.msm_gens(sG, None, n, off) == ol.msm_gens(sG, None, n, off)
    assert ctx.msm_gens(None, sH, n, 0) == ol.msm_gens(None, sH, n, 0)


def test_msm_gens_with_extra_points(ctx):
    rnd = random.Random(5)
    n, k = 200, 9
    sG, sH = b"".join(rs(rnd) for _ in range(n)), b"".join(rs(rnd) for _ in range(n))
    es = b"".join(rs(rnd) for _ in range(k))
    G, _ = ol.gens(0, 4)
    B, Bb = ol.pedersen_gens()
    ep = B + Bb + G + ol.point_mul(rs(rnd), B) + ol.point_mul(rs(rnd), Bb) + bytes(32)
    assert ctx.msm_gens(sG, sH, n, 3, es, ep) == ol.msm_gens(sG, sH, n, 3, es, ep)
    import bulletproofs_gadgets_b200 as bpg
    bad = ep[:64] + b"\x01" + bytes(31) + ep[96:]
    with pytest.raises(bpg.BpgError) as e:
        ctx.msm_gens(sG, sH, n, 3, es, bad)
    assert e.value.code == -3
    with pytest.raises(bpg.BpgError) as e:
        ctx.msm_gens(sG, sH, n, 4096 - 10)
    assert e.value.code == -2


def test_variable_base_msm_and_fold(ctx):
    rnd = random.Random(6)
    n = 150
    sc = b"".join(rs(rnd) for _ in range(n))
    G, H = ol.gens(0, n)
    assert ctx.msm(sc, G) == ol.msm(sc, G, ol.VARTIME)
    assert ctx.msm(b"", b"") == bytes(32)
    u = rnd.randrange(L)
    ui = pr.sc_inv(u)
    out = ctx.fold_points(ui, u, G[:32 * 64], G[32 * 64:32 * 128])
    assert out == ol.fold_points(pr.sc_bytes(ui), pr.sc_bytes(u), G[:32 * 64], G[32 * 64:32 * 128])


def test_fold_points_shared_scalar_straus_edges(ctx):
    """the literal IPP generator fold G'_i = u^-1 G_i + u G_{i+h} (dalek InnerProductProof::create, behind prover.rs:93): ragged
    sizes around the 64-thread block, scalars 0, 1, l - 1, unreduced 2^255 - 1, digits that carry into every window"""
    rnd = random.Random(61)
    G, H = ol.gens(0, 331)
    for n in (1, 63, 64, 65, 331):
        for sl, sr in ((0, 0), (1, 0), (0, 1), (L - 1, 1), (2 ** 255 - 1, 2 ** 255 - 1), (int("8" * 63, 16) % L, int("f" * 63, 16) % L),
                       (rnd.randrange(L), rnd.randrange(L))):
            a, b = sl.to_bytes(32, "little"), sr.to_bytes(32, "little")  # the device reduces mod l itself (a12: Scalar semantics)
            want = ol.fold_points(pr.sc_bytes(sl % L), pr.sc_bytes(sr % L), G[:32 * n], H[:32 * n])
            assert ctx.fold_points(a, b, G[:32 * n], H[:32 * n]) == want, (n, sl, sr)


def test_device_resident_and_partial_sum(ctx):
    rnd = random.Random(7)
    n = 512
    sG, sH = b"".join(rs(rnd, 252) for _ in range(n)), b"".join(rs(rnd, 252) for _ in range(n))
    dG, dH = ctx.dev_alloc(32 * n), ctx.dev_alloc(32 * n)
    ctx.dev_upload(dG, sG); ctx.dev_upload(dH, sH)
    want = ol.msm_gens(sG, sH, n, 0)
    assert ctx.msm_gens_dev(dG, dH, n, 0) == want
    # point-range split (the multi-GPU MSM shape): two partial sums combined
    h = n // 2
    import ctypes as C
    p0 = ctx.msm_gens_partial_dev(dG, dH, h, 0)
    p1 = ctx.msm_gens_partial_dev(C.c_void_p(dG.value + 32 * h), C.c_void_p(dH.value + 32 * h), h, h)
    assert ctx.points_sum_compress(p0 + p1) == want
    ctx.dev_free(dG); ctx.dev_free(dH)


def test_mimc_matches_reference_kats(ctx):
    import json, os
    kats = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "mimc_kats.json")))
    pre = [bytes.fromhex(p) for p, _, _ in kats["hash"]]
    got = ctx.mimc_hash_batch(pre)
    for (p, dig, src), g in zip(kats["hash"], got):
        assert g[::-1].hex() == dig, src
    pairs = [[bytes.fromhex(l)[::-1].ljust(32, b"\0"), bytes.fromhex(r)[::-1].ljust(32, b"\0")] for l, r, _, _ in kats["node"]]
    got, _ = ctx.mimc_sponge_batch(pairs)
    for (_, _, dig, src), g in zip(kats["node"], got):
        assert g[::-1].hex() == dig, src


def test_mimc_batch_and_trace_match_oracle(ctx):
    rnd = random.Random(9)
    lists = [[rs(rnd, 256) for _ in range(rnd.randrange(1, 4))] for _ in range(300)]
    got, tr = ctx.mimc_sponge_batch(lists, trace=True)
    off = 0
    for bl, g in zip(lists, got):
        d, t = ol.mimc_sponge(b"".join(bl), trace=True)
        assert g == d
        assert tr[off:off + len(t)] == t
        off += len(t)
    pre = [rnd.randbytes(rnd.randrange(1, 100)) for _ in range(200)] + [b"\x00", b"\x00" * 40, b"\xff" * 32, b"\x01" + b"\x00" * 31]
    assert ctx.mimc_hash_batch(pre) == [ol.mimc_hash(p) for p in pre]


@pytest.mark.parametrize("dist,lg", [("uniform", 19), ("bits", 19), ("uniform", 20)])
def test_large_msm_split_consistency(dist, lg):
    """2^20 / 2^21-term MSM over the resident generators (shared-memory privatised histogram / scatter, >= 2^19 terms) equals the sum of
    its four quarter ranges (2^18 terms each: plain global-atomic kernels) -- a size-independent property at BASELINE scale that
    cross-checks the two sort paths; `bits` puts half of all pairs into bucket 1 (block-tree path for heavy buckets)."""
    import ctypes as C
    import numpy as np
    import bulletproofs_gadgets_b200 as bpg
    c = bpg.Context(0)
    h = 1 << lg  # 2^(lg+1) terms; lg = 20: privatised histogram + plain scatter
    c.gens_ensure(h)
    rng = np.random.default_rng(5)
    if dist == "uniform":
        raw = rng.integers(0, 256, size=(2 * h, 32), dtype=np.uint8)
        raw[:, 31] &= 0x0F
    else:
        raw = np.zeros((2 * h, 32), dtype=np.uint8)
        raw[:, 0] = rng.integers(0, 2, size=2 * h, dtype=np.uint8)
    d = c.dev_alloc(64 * h)
    c.dev_upload(d, raw.tobytes())
    dG, dH = d, C.c_void_p(d.value + 32 * h)
    full = c.msm_gens_dev(dG, dH, h, 0)
    parts = []
    q = h // 4
    for k in range(4):
        parts.append(c.msm_gens_partial_dev(C.c_void_p(dG.value + 32 * q * k), C.c_void_p(dH.value + 32 * q * k), q, q * k))
    assert c.points_sum_compress(b"".join(parts)) == full
    assert full != bytes(32)
    c.dev_free(d)
    c.close()


@pytest.mark.parametrize("n", [8192, 10000, 40000])
def test_variable_base_bucket_msm_matches_oracle(ctx, n):
    """bpg_msm from 8192 terms on: decompress to affine Niels + the bucket engine over the points themselves (16 windows = 16
    bucket groups, Horner recombination).  Unreduced scalars, repeated points, the identity encoding, zero scalars."""
    import bulletproofs_gadgets_b200 as bpg
    rnd = random.Random(n)
    G, H = ol.gens(0, 2048)
    base = [G[32 * i:32 * i + 32] for i in range(2048)] + [H[32 * i:32 * i + 32] for i in range(2048)] + [bytes(32)]
    pts = b"".join(base[rnd.randrange(len(base))] for _ in range(n))
    scs = [rs(rnd, 256) for _ in range(n)]
    for i in range(0, n, 97):
        scs[i] = bytes(32)
    scs[1], scs[2], scs[3] = (L - 1).to_bytes(32, "little"), L.to_bytes(32, "little"), (1).to_bytes(32, "little")
    sc = b"".join(scs)
    assert ctx.msm(sc, pts) == ol.msm(sc, pts, ol.VARTIME)
    if n == 8192:
        bad = pts[:32 * 5000] + b"\x01" + bytes(31) + pts[32 * 5001:]  # s = 1 is not a valid ristretto encoding
        with pytest.raises(bpg.BpgError) as e:
            ctx.msm(sc, bad)
        assert e.value.code == -3
        # all terms cancel: k P - k P
        half = n // 2
        neg = b"".join(((L - int.from_bytes(ol.sc_reduce(s), "little")) % L).to_bytes(32, "little") for s in scs[:half])
        assert ctx.msm(b"".join(scs[:half]) + neg, pts[:32 * half] * 2) == bytes(32)


def test_verify_with_many_commitments_uses_bucket_path(ctx):
    """m = 8400 committed values (>= 8192: the verifier's own points go through the variable-base bucket engine, on the main
    stream): proof bytes equal the oracle's, both verifiers accept, and a tampered commitment is rejected"""
    from bulletproofs_gadgets_b200 import gadgets
    ctx.gens_ensure(1 << 16)
    inst = gadgets.bounds_check_batch_instance(2800, 1, seed=77)
    assert inst["m"] == 8400
    circ = gadgets.Circuit(ctx, inst["n"], inst["m"], inst["csr"])
    ext = b"\x61" * 32
    proof, V = circ.prove(inst, ext)
    rp, tv, tc = inst["csr"]
    ol.lib().bpo_set_threads(8)
    try:
        assert (proof, V) == ol.r1cs_prove(inst["label"], 1 << 16, inst["aL"], inst["aR"], inst["aO"], inst["vals"], inst["blinds"], rp, tv, tc.tobytes(), ext)
        assert circ.verify(inst["label"], V, proof, b"\x62" * 32)
        assert ol.r1cs_verify(inst["label"], 1 << 16, inst["n"], V, rp, tv, tc.tobytes(), proof, b"\x63" * 32)
        Vbad = bytearray(V)
        Vbad[32 * 4000:32 * 4001] = V[32 * 4001:32 * 4002]
        assert not circ.verify(inst["label"], bytes(Vbad), proof, b"\x62" * 32)
    finally:
        ol.lib().bpo_set_threads(1)
        circ.close()
