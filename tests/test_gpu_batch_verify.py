"""Batch verification (BASELINE config 5 shape): verdicts of bpg_r1cs_verify_batch must equal one-by-one
verification (GPU) and the CPU oracle, proof by proof, for batches that mix valid proofs with every kind of failure."""
import ctypes as C
import random

import pytest

import circuits
import oracle_lib as ol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import bulletproofs_gadgets_b200 as bpg
    c = bpg.Context(0)
    c.gens_ensure(512)
    yield c
    c.close()


def make_items(ctx, count, seed, bad_every):
    rnd = random.Random(seed)
    items, expect = [], []
    for k in range(count):
        nm = rnd.choice([0, 1, 3, 8, 17, 40, 100, 130, 260])
        kind = "ok"
        if bad_every and k % bad_every == bad_every - 1:
            kind = rnd.choice(["wrong_witness", "flip_proof", "bad_V", "truncated", "label", "flip_scalar"])
        inst = circuits.chain_instance(nm, 9000 + k, wrong=(kind == "wrong_witness"))
        rp, tv, tc = inst["csr"]
        proof, V = ol.r1cs_prove(inst["label"], 512, inst["aL"], inst["aR"], inst["aO"], inst["vals"], inst["blinds"], rp, tv, tc, bytes([k % 256]) * 32)
        label = inst["label"]
        if kind == "flip_proof":
            b = bytearray(proof); b[rnd.randrange(1, 1 + 32 * 8)] ^= 1 << rnd.randrange(8); proof = bytes(b)
        elif kind == "flip_scalar":
            b = bytearray(proof); b[-5] ^= 2; proof = bytes(b)
        elif kind == "bad_V":
            V = b"\x01" + V[1:]
        elif kind == "truncated":
            proof = proof[:-32]
        elif kind == "label":
            label = b"other"
        h = C.c_void_p()
        ctx.check(ctx.lib.bpg_circuit_create(ctx.h, inst["n"], 3, len(rp) - 1, (C.c_uint32 * len(rp))(*rp), (C.c_uint32 * max(1, len(tv)))(*tv), tc, C.byref(h)))
        want = ol.r1cs_verify(label, 512, inst["n"], V, rp, tv, tc, proof, bytes(32))
        items.append((h, label, V, proof, bytes([7 + k % 200]) * 32))
        expect.append(want)
        assert want == (kind == "ok"), kind
    return items, expect


def single(ctx, it):
    acc = C.c_int(-1)
    ctx.check(ctx.lib.bpg_r1cs_verify(ctx.h, it[0], it[1], len(it[1]), it[2], it[3], len(it[3]), it[4], 0, C.byref(acc)))
    return bool(acc.value)


@pytest.mark.parametrize("count,bad_every", [(1, 0), (2, 2), (33, 0), (48, 5), (40, 1), (64, 31)])
def test_batch_verdicts_equal_individual_and_oracle(ctx, count, bad_every):
    items, expect = make_items(ctx, count, 100 * count + bad_every, bad_every)
    got = ctx.verify_batch(items)
    assert got == expect
    assert [single(ctx, it) for it in items] == expect
    for it in items:
        ctx.lib.bpg_circuit_destroy(it[0])


def test_batch_uses_far_fewer_launches_than_one_by_one(ctx):
    items, expect = make_items(ctx, 64, 77, 0)
    l0 = ctx.launch_count()
    assert ctx.verify_batch(items) == expect
    l1 = ctx.launch_count()
    assert [single(ctx, it) for it in items] == expect
    l2 = ctx.launch_count()
    assert (l1 - l0) < 0.6 * (l2 - l1)
    assert ctx.verify_batch([]) == []
    for it in items:
        ctx.lib.bpg_circuit_destroy(it[0])
