"""GPU parity tests for the R1CS prover / verifier path (a3-a9): proof bytes must equal the CPU oracle's under the
same transcript label and seeded randomness; verdicts must match on honest, tampered and malformed proofs."""
import random

import pytest

import circuits
import oracle_lib as ol
from oracle import pyref as pr

pytestmark = pytest.mark.gpu
L = pr.L


@pytest.fixture(scope="module")
def ctx():
    import bulletproofs_gadgets_b200 as bpg
    c = bpg.Context(0)
    c.gens_ensure(2048)
    yield c
    c.close()


def gpu_prove(ctx, inst, ext, flags=0):
    import ctypes as C
    rp, tv, tc = inst["csr"]
    n, m, q = len(inst["aL"]) // 32, len(inst["vals"]) // 32, len(rp) - 1
    h = C.c_void_p()
    ctx.check(ctx.lib.bpg_circuit_create(ctx.h, n, m, q, (C.c_uint32 * (q + 1))(*rp), (C.c_uint32 * max(1, len(tv)))(*tv), tc, C.byref(h)))
    cap = 1 + 32 * (14 + 64 + 2)
    proof, V = C.create_string_buffer(cap), C.create_string_buffer(32 * max(1, m))
    rc = ctx.lib.bpg_r1cs_prove(ctx.h, h, inst["label"], len(inst["label"]), inst["aL"], inst["aR"], inst["aO"], inst["vals"], inst["blinds"], ext,
                                flags, V, proof, cap)
    ctx.lib.bpg_circuit_destroy(h)
    if rc < 0:
        ctx.check(rc)
    return proof.raw[:rc], V.raw[:32 * m]


def gpu_verify(ctx, inst, V, proof, ext=b"\x07" * 32, flags=0, label=None, n=None):
    import ctypes as C
    rp, tv, tc = inst["csr"]
    n = len(inst["aL"]) // 32 if n is None else n
    m, q = len(V) // 32, len(rp) - 1
    h = C.c_void_p()
    ctx.check(ctx.lib.bpg_circuit_create(ctx.h, n, m, q, (C.c_uint32 * (q + 1))(*rp), (C.c_uint32 * max(1, len(tv)))(*tv), tc, C.byref(h)))
    acc = C.c_int(-1)
    lab = inst["label"] if label is None else label
    rc = ctx.lib.bpg_r1cs_verify(ctx.h, h, lab, len(lab), V, proof, len(proof), ext, flags, C.byref(acc))
    ctx.lib.bpg_circuit_destroy(h)
    ctx.check(rc)
    return bool(acc.value)


def oracle_prove(inst, cap, ext, flags=0):
    rp, tv, tc = inst["csr"]
    return ol.r1cs_prove(inst["label"], cap, inst["aL"], inst["aR"], inst["aO"], inst["vals"], inst["blinds"], rp, tv, tc, ext, flags)


@pytest.mark.parametrize("nmul,seed", [(0, 1), (1, 2), (2, 3), (3, 4), (5, 5), (8, 6), (13, 7), (64, 8), (100, 9), (300, 10)])
def test_chain_proof_bytes_match_oracle(ctx, nmul, seed):
    inst = circuits.chain_instance(nmul, seed)
    ext = bytes(range(32))
    want, Vw = oracle_prove(inst, 2048, ext)
    got, V = gpu_prove(ctx, inst, ext)
    assert V == Vw
    assert got == want
    rp, tv, tc = inst["csr"]
    assert gpu_verify(ctx, inst, V, got)
    assert ol.r1cs_verify(inst["label"], 2048, nmul, V, rp, tv, tc, got, b"\x01" * 32)
    # tampering anywhere -> reject, exactly like the oracle
    rnd = random.Random(seed)
    for _ in range(6):
        bad = bytearray(got)
        pos = rnd.randrange(1, len(bad))
        bad[pos] ^= 1 << rnd.randrange(8)
        assert gpu_verify(ctx, inst, V, bytes(bad)) == ol.r1cs_verify(inst["label"], 2048, nmul, V, rp, tv, tc, bytes(bad), b"\x01" * 32) == False
    assert not gpu_verify(ctx, inst, V, got, label=b"other label")
    assert not gpu_verify(ctx, inst, V, got[:-32])
    assert not gpu_verify(ctx, inst, V, got[:-1])
    assert not gpu_verify(ctx, inst, V, b"")
    assert not gpu_verify(ctx, inst, V, b"\x02" + got[1:])
    if len(V) >= 64:
        assert not gpu_verify(ctx, inst, V[32:64] + V[:32] + V[64:], got)
        assert not gpu_verify(ctx, inst, b"\x01" + V[1:], got)  # commitment that fails to decompress


@pytest.mark.parametrize("n,seed", [(1, 30), (37, 31), (256, 32), (257, 33), (1000, 34)])
def test_dense_proof_bytes_match_oracle(ctx, n, seed):
    inst = circuits.random_dense_instance(n, seed, m=3)
    ext = bytes([seed]) * 32
    want, Vw = oracle_prove(inst, 2048, ext)
    got, V = gpu_prove(ctx, inst, ext)
    assert (got, V) == (want, Vw)
    assert gpu_verify(ctx, inst, V, got)


def test_legacy_framing_and_fast_blinding(ctx):
    inst = circuits.chain_instance(20, 77)
    ext = b"\x42" * 32
    want, _ = oracle_prove(inst, 2048, ext, flags=1)
    got, V = gpu_prove(ctx, inst, ext, flags=1)
    assert got == want and len(got) == 32 * (14 + 2 * 5 + 2)
    assert gpu_verify(ctx, inst, V, got, flags=1)
    assert not gpu_verify(ctx, inst, V, got, flags=0)
    fast, V2 = gpu_prove(ctx, inst, ext, flags=2)
    rp, tv, tc = inst["csr"]
    assert V2 == V and fast != oracle_prove(inst, 2048, ext)[0]
    assert gpu_verify(ctx, inst, V, fast)
    assert ol.r1cs_verify(inst["label"], 2048, 20, V, rp, tv, tc, fast, bytes(32))


def test_unsatisfied_witness_and_capacity(ctx):
    import bulletproofs_gadgets_b200 as bpg
    inst = circuits.chain_instance(9, 55, wrong=True)
    proof, V = gpu_prove(ctx, inst, bytes(32))
    assert proof == oracle_prove(inst, 2048, bytes(32))[0]
    assert not gpu_verify(ctx, inst, V, proof)
    big = circuits.random_dense_instance(2049, 3, m=1)
    with pytest.raises(bpg.BpgError) as e:
        gpu_prove(ctx, big, bytes(32))
    assert e.value.code == -2  # InvalidGeneratorsLength


def test_heavy_column_constraints(ctx):
    """a committed variable and the constant appear in thousands of constraints (split flatten columns)."""
    inst = circuits.chain_instance(700, 91)
    want, _ = oracle_prove(inst, 2048, b"\x09" * 32)
    got, V = gpu_prove(ctx, inst, b"\x09" * 32)
    assert got == want
    assert gpu_verify(ctx, inst, V, got)


def test_reference_style_api_roundtrip(ctx):
    """Prover/Verifier mirror used like the reference's unit tests (e.g. bounds_check_gadget.rs:80-98):
    range_proof of utils.rs:5-35 on a 16-bit value: is_ok for an in-range value, is_err otherwise."""
    import bulletproofs_gadgets_b200 as bpg

    def range_proof(cs, v_lc, v_assign, nbits):
        exp2 = 1
        acc = list(v_lc)
        for i in range(nbits):
            bit = None if v_assign is None else ((v_assign >> i) & 1)
            a, b, o = cs.allocate_multiplier(None if bit is None else (1 - bit, bit))
            cs.constrain([(o, 1)])
            cs.constrain([(a, 1), (b, 1), (bpg.api.ONE, L - 1)])
            acc = acc + [(b, (-exp2) % L)]
            exp2 *= 2
        cs.constrain(acc)

    bp = bpg.BulletproofGens.new(64, 1, ctx=ctx)
    for value, ok in ((513, True), (65535, True), (65536, False)):
        prover = bpg.Prover.new(b"RangeProofTest", ctx=ctx)
        _, var = prover.commit(value, 123456789)
        range_proof(prover, [(var, 1)], value, 16)
        proof, V = prover.prove(bp, ext_rng32=b"\x11" * 32)
        verifier = bpg.Verifier.new(b"RangeProofTest", ctx=ctx)
        var = verifier.commit(V[0])
        range_proof(verifier, [(var, 1)], None, 16)
        if ok:
            verifier.verify(proof, None, bp)
        else:
            with pytest.raises(bpg.R1CSError):
                verifier.verify(proof, None, bp)


@pytest.mark.parametrize("nmul,seed", [(3, 21), (4, 22), (5, 23), (8, 24), (13, 25), (64, 26), (100, 27), (300, 28), (700, 29), (1500, 30)])
def test_late_fold_proof_bytes_match_oracle(ctx, nmul, seed):
    """Late fold (materialised G^(k), H^(k) after the first IPP rounds, kernels_msm.cuh): same L_j / R_j, hence the same proof
    bytes as the oracle (which folds the generators in every round like dalek) and as the path without it."""
    import bulletproofs_gadgets_b200._lib as lb
    inst = circuits.chain_instance(nmul, seed)
    ext = bytes(range(1, 33))
    want, Vw = oracle_prove(inst, 2048, ext)
    forced, V1 = gpu_prove(ctx, inst, ext, lb.FLAG_FORCE_LATE_FOLD)
    plain, V2 = gpu_prove(ctx, inst, ext, lb.FLAG_NO_LATE_FOLD)
    assert V1 == V2 == Vw
    assert forced == want
    assert plain == want
    assert gpu_verify(ctx, inst, V1, forced)


@pytest.mark.parametrize("mode", [0, 1])
def test_sizing_modes_give_identical_proofs(ctx, mode):
    """latency vs throughput kernel sizing (bpg_set_sizing_mode: accumulate chunk length, row/column reduction variant) must not
    change a single byte; sizes chosen so that the commitment MSMs and the first IPP rounds take the 2^15-bucket path"""
    import bulletproofs_gadgets_b200._lib as lb
    inst = circuits.chain_instance(1500, 77)
    ext = bytes(range(3, 35))
    want, Vw = oracle_prove(inst, 2048, ext)
    ctx.lib.bpg_set_sizing_mode(mode)
    try:
        got, V = gpu_prove(ctx, inst, ext)
        forced, _ = gpu_prove(ctx, inst, ext, lb.FLAG_FORCE_LATE_FOLD)
        assert gpu_verify(ctx, inst, V, got)
    finally:
        ctx.lib.bpg_set_sizing_mode(-1)
    assert V == Vw and got == want and forced == want


def test_concurrent_provers_match_oracle():
    """eight provers (own context each) in eight host threads: shared generator tables, shared transcript-RNG lanes, automatic
    throughput sizing -- every proof byte-identical to the oracle's"""
    import threading
    import bulletproofs_gadgets_b200 as bpg
    inst = circuits.chain_instance(700, 31)
    want, Vw = oracle_prove(inst, 2048, bytes(range(32)))
    ctxs = [bpg.Context(0) for _ in range(8)]
    for c in ctxs:
        c.gens_ensure(2048)
    out, bar = {}, threading.Barrier(8)

    def work(k):
        bar.wait()
        for rep in range(3):
            out[(k, rep)] = gpu_prove(ctxs[k], inst, bytes(range(32)))

    ths = [threading.Thread(target=work, args=(k,)) for k in range(8)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for c in ctxs:
        c.close()
    assert len(out) == 24 and all(v == (want, Vw) for v in out.values())


def test_prefetched_opening_gives_identical_proofs(ctx):
    """bpg_r1cs_prove_prefetch: the transcript-RNG stream of a future proof drawn in the background.  Proof bytes must equal the
    oracle's whether the hint was given, given for another proof (ignored / kept pending), or given for two proofs at once."""
    import circuits
    inst = circuits.chain_instance(300, 77)
    import ctypes as C
    rp, tv, tc = inst["csr"]
    h = C.c_void_p()
    ctx.check(ctx.lib.bpg_circuit_create(ctx.h, inst["n"], 3, len(rp) - 1, (C.c_uint32 * len(rp))(*rp), (C.c_uint32 * max(1, len(tv)))(*tv), tc, C.byref(h)))

    def prove(ext):
        cap = 1 + 32 * (14 + 64 + 2)
        proof, V = C.create_string_buffer(cap), C.create_string_buffer(96)
        rc = ctx.lib.bpg_r1cs_prove(ctx.h, h, inst["label"], len(inst["label"]), inst["aL"], inst["aR"], inst["aO"], inst["vals"], inst["blinds"], ext, 0, V, proof, cap)
        assert rc > 0
        return proof.raw[:rc], V.raw

    def prefetch(ext, blinds=None):
        ctx.check(ctx.lib.bpg_r1cs_prove_prefetch(ctx.h, h, inst["label"], len(inst["label"]), inst["vals"], blinds or inst["blinds"], ext, 0))

    exts = [bytes([k]) * 32 for k in range(1, 6)]
    want = [oracle_prove(inst, 4096, e) for e in exts]
    assert prove(exts[0]) == want[0]                      # no hint
    prefetch(exts[1])
    assert prove(exts[1]) == want[1]                      # hinted
    prefetch(exts[2]); prefetch(exts[3])                  # two pending
    prefetch(exts[4])                                     # third hint: dropped
    assert prove(exts[3]) == want[3]                      # consumed out of order
    assert prove(exts[4]) == want[4]                      # never prefetched
    assert prove(exts[2]) == want[2]                      # still pending from before
    prefetch(exts[0], blinds=inst["blinds"][32:] + inst["blinds"][:32])  # hint for other blindings: must not be used
    assert prove(exts[0]) == want[0]
    ctx.lib.bpg_circuit_destroy(h)


def test_page_locked_host_buffers_give_identical_proofs(ctx):
    """bpg_host_alloc: a witness handed over in page-locked host memory (one asynchronous DMA) must give the proof bytes of the
    same witness in ordinary memory, which equal the oracle's; freeing and re-allocating leaves the library usable"""
    import circuits
    from bulletproofs_gadgets_b200 import gadgets
    inst = circuits.chain_instance(500, 91)
    rp, tv, tc = inst["csr"]
    import numpy as np
    circ = gadgets.Circuit(ctx, inst["n"], 3, (np.array(rp, dtype=np.uint32), np.array(tv, dtype=np.uint32), tc))
    ext = b"\x5a" * 32
    want = circ.prove(inst, ext)
    assert want == oracle_prove(inst, 4096, ext)
    for _ in range(2):
        pinned, handles = dict(inst), []
        for k in ("aL", "aR", "aO"):
            ptr, hnd = ctx.host_alloc(inst[k])
            pinned[k] = ptr
            handles.append(hnd)
        assert circ.prove(pinned, ext) == want
        for hnd in handles:
            ctx.host_free(hnd)
    circ.close()
