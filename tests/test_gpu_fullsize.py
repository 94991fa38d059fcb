"""GPU parity at the FULL sizes BASELINE.json names: the GPU prover's proof bytes (and commitments) must equal the CPU
oracle prover's, byte for byte, under the same transcript label, blindings and ext_rng32 -- not merely verify.

  configs[1]  merkle_tree membership with mimc_hash, depth 32          n = 63 180     N = 2^16   m = 4
  configs[2]  4096 bounds_check 64-bit range gadgets, one proof        n = 524 288    N = 2^19   m = 12 288  (bit-valued witness)
  configs[3]  2^20-multiplier circuit (the reference's ignored 512-leaf test size, merkle_tree_gadget.rs:473-545)
                                                                       n = 993 384    N = 2^20   m = 512     20 IPP rounds

The oracle runs with all host cores (OpenMP); these three tests take about a minute on a 16-core GPU box."""
import os

import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import bulletproofs_gadgets_b200 as bpg
    c = bpg.Context(0)
    yield c
    c.close()


@pytest.fixture(autouse=True)
def oracle_threads():
    ol.lib().bpo_set_threads(os.cpu_count() or 1)
    yield
    ol.lib().bpo_set_threads(1)


def _check_full(ctx, inst, cap, ext):
    from bulletproofs_gadgets_b200 import gadgets
    ctx.gens_ensure(cap)
    circ = gadgets.Circuit(ctx, inst["n"], inst["m"], inst["csr"])
    try:
        proof, V = circ.prove(inst, ext)
        rp, tv, tc = inst["csr"]
        want, Vw = ol.r1cs_prove(inst["label"], cap, inst["aL"], inst["aR"], inst["aO"], inst["vals"], inst["blinds"], rp, tv, tc.tobytes(), ext)
        assert V == Vw
        assert proof == want
        assert circ.verify(inst["label"], V, proof)
        assert ol.r1cs_verify(inst["label"], cap, inst["n"], V, rp, tv, tc.tobytes(), proof, b"\x09" * 32)
        bad = bytearray(proof)
        bad[len(bad) // 2] ^= 0x10
        assert not circ.verify(inst["label"], V, bytes(bad))
    finally:
        circ.close()


def test_config1_merkle_depth32_proof_bytes_equal_oracle(ctx):
    from bulletproofs_gadgets_b200 import gadgets
    inst = gadgets.merkle_path_instance(32, ctx=ctx)
    assert inst["n"] == 63180 and inst["m"] == 4
    _check_full(ctx, inst, 1 << 16, b"\x51" * 32)


def test_config2_4096_bounds_checks_proof_bytes_equal_oracle(ctx):
    from bulletproofs_gadgets_b200 import gadgets
    inst = gadgets.bounds_check_batch_instance(4096, 8, seed=5)
    assert inst["n"] == 1 << 19 and inst["m"] == 12288
    _check_full(ctx, inst, 1 << 19, b"\x52" * 32)


def test_config3_2p20_multipliers_proof_bytes_equal_oracle(ctx):
    from bulletproofs_gadgets_b200 import gadgets
    """the reference's own largest test circuit (test_merkle_tree_gadget_512, #[ignore]d there as too slow): 512 committed leaves
    = W1, 511 MiMC nodes; the root the device computes level by level is the one pinned at merkle_tree_gadget.rs:476"""
    inst = gadgets.merkle_tree_instances(512, [None], ctx=ctx)[0]
    assert inst["n"] == 993384 and inst["m"] == 512
    assert inst["root"].to_bytes(32, "big") == gadgets.REF_ROOT_512
    _check_full(ctx, inst, 1 << 20, b"\x53" * 32)
