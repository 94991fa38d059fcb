"""GPU parity tests on the BASELINE.json workload families at sizes the oracle finishes in seconds, and at full size
through size-independent properties (prove -> verify accepts, any tampering rejects, the oracle's verifier accepts the
GPU's proof)."""
import random

import numpy as np
import pytest

import oracle_lib as ol
from oracle import pyref as pr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import bulletproofs_gadgets_b200 as bpg
    c = bpg.Context(0)
    yield c
    c.close()


def oracle_prove(inst, cap, ext):
    rp, tv, tc = inst["csr"]
    return ol.r1cs_prove(inst["label"], cap, inst["aL"], inst["aR"], inst["aO"], inst["vals"], inst["blinds"], rp, tv, tc.tobytes(), ext)


def oracle_verify(inst, cap, V, proof):
    rp, tv, tc = inst["csr"]
    return ol.r1cs_verify(inst["label"], cap, inst["n"], V, rp, tv, tc.tobytes(), proof, b"\x05" * 32)


def test_merkle_path_depth3_bytes_match_oracle(ctx):
    from bulletproofs_gadgets_b200 import gadgets
    ctx.gens_ensure(8192)
    inst = gadgets.merkle_path_instance(3, ctx=ctx)
    host = gadgets.merkle_path_instance(3, trace_on_device=False)
    assert (inst["aL"], inst["aR"], inst["aO"]) == (host["aL"], host["aR"], host["aO"])  # device MiMC trace == host big-int trace
    circ = gadgets.Circuit(ctx, inst["n"], inst["m"], inst["csr"])
    ext = b"\x21" * 32
    proof, V = circ.prove(inst, ext)
    assert (proof, V) == oracle_prove(inst, 8192, ext)
    assert circ.verify(inst["label"], V, proof)
    circ.close()


@pytest.mark.parametrize("count,nbytes", [(1, 1), (3, 2), (16, 8)])
def test_bounds_check_batch_bytes_match_oracle(ctx, count, nbytes):
    """bit-valued a_L / a_R (range proofs): half of all MSM terms land in bucket 1 (block-wide heavy-bucket path)"""
    from bulletproofs_gadgets_b200 import gadgets
    ctx.gens_ensure(4096)
    inst = gadgets.bounds_check_batch_instance(count, nbytes, seed=count)
    circ = gadgets.Circuit(ctx, inst["n"], inst["m"], inst["csr"])
    ext = bytes([count]) * 32
    proof, V = circ.prove(inst, ext)
    assert (proof, V) == oracle_prove(inst, 4096, ext)
    assert circ.verify(inst["label"], V, proof)
    circ.close()
    # a value outside the range: bits no longer recompose -> both verifiers reject
    hi = (1 << (8 * nbytes)) - 1
    bad = gadgets.bounds_check_batch_instance(count, nbytes, seed=count, values=[hi + 5] + inst["values"][1:])
    circ = gadgets.Circuit(ctx, bad["n"], bad["m"], bad["csr"])
    proof, V = circ.prove(bad, ext)
    assert not circ.verify(bad["label"], V, proof)
    assert not oracle_verify(bad, 4096, V, proof)
    circ.close()


def test_large_bounds_batch_properties(ctx):
    """1024 x 64-bit bounds checks = 2^17 multipliers, 3072 commitments (1/4 of BASELINE config 3): the oracle's VERIFIER
    accepts the GPU proof; tampering is rejected by both; commitments equal the oracle's batch"""
    from bulletproofs_gadgets_b200 import gadgets
    ctx.gens_ensure(1 << 17)
    inst = gadgets.bounds_check_batch_instance(1024, 8, seed=99)
    assert inst["n"] == 1 << 17 and inst["m"] == 3072
    circ = gadgets.Circuit(ctx, inst["n"], inst["m"], inst["csr"])
    proof, V = circ.prove(inst, b"\x31" * 32)
    assert V == ol.pedersen_commit(inst["vals"], inst["blinds"])
    assert circ.verify(inst["label"], V, proof)
    ol.lib().bpo_set_threads(8)
    try:
        assert oracle_verify(inst, 1 << 17, V, proof)
        bad = bytearray(proof)
        bad[200] ^= 4
        assert not circ.verify(inst["label"], V, bytes(bad))
        assert not oracle_verify(inst, 1 << 17, V, bytes(bad))
    finally:
        ol.lib().bpo_set_threads(1)
    circ.close()


def test_large_mimc_chain_properties(ctx):
    """270 absorbed MiMC blocks = 262 440 multipliers -> N = 2^19 with 261 848 padded positions (exercises the u-factor
    padding of the inner-product argument at scale); oracle verifier must accept the GPU proof"""
    from bulletproofs_gadgets_b200 import gadgets
    ctx.gens_ensure(1 << 19)
    inst = gadgets.mimc_chain_instance(270, ctx=ctx)
    assert inst["n"] == 270 * 972
    circ = gadgets.Circuit(ctx, inst["n"], inst["m"], inst["csr"])
    proof, V = circ.prove(inst, b"\x41" * 32)
    assert len(proof) == 1 + 32 * (11 + 2 * 19 + 2)
    assert circ.verify(inst["label"], V, proof)
    ol.lib().bpo_set_threads(8)
    try:
        assert oracle_verify(inst, 1 << 19, V, proof)
    finally:
        ol.lib().bpo_set_threads(1)
    bad = bytearray(proof)
    bad[-33] ^= 1
    assert not circ.verify(inst["label"], V, bytes(bad))
    circ.close()


def test_batch_verification_verdicts_match_oracle(ctx):
    """BASELINE config 5 in miniature: many small independent proofs (own transcript each), ~10 % deliberately invalid;
    GPU verdicts must equal the oracle's one by one."""
    import circuits
    import ctypes as C
    ctx.gens_ensure(512)
    rnd = random.Random(3)
    items = []
    for k in range(24):
        inst = circuits.chain_instance(rnd.randrange(1, 40), 1000 + k, wrong=(k % 9 == 4))
        rp, tv, tc = inst["csr"]
        proof, V = ol.r1cs_prove(inst["label"], 512, inst["aL"], inst["aR"], inst["aO"], inst["vals"], inst["blinds"], rp, tv, tc, bytes([k]) * 32)
        if k % 7 == 3:
            b = bytearray(proof)
            b[rnd.randrange(1, len(b))] ^= 1 << rnd.randrange(8)
            proof = bytes(b)
        items.append((inst, V, proof))
    n_bad = 0
    for inst, V, proof in items:
        rp, tv, tc = inst["csr"]
        want = ol.r1cs_verify(inst["label"], 512, inst["n"], V, rp, tv, tc, proof, bytes(32))
        h = C.c_void_p()
        ctx.check(ctx.lib.bpg_circuit_create(ctx.h, inst["n"], 3, len(rp) - 1, (C.c_uint32 * len(rp))(*rp), (C.c_uint32 * max(1, len(tv)))(*tv), tc, C.byref(h)))
        acc = C.c_int(-1)
        ctx.check(ctx.lib.bpg_r1cs_verify(ctx.h, h, inst["label"], len(inst["label"]), V, proof, len(proof), bytes(32), 0, C.byref(acc)))
        ctx.lib.bpg_circuit_destroy(h)
        assert bool(acc.value) == want
        n_bad += (not want)
    assert 3 <= n_bad <= 12
