"""Front-end driver: the reference's own fixtures (tests/golden/fixtures = /root/reference/tests/resources
+ example.*, copied by make_golden.py) through the `.gadgets` language, gadgets, buffer replay and OR conjunction.

CPU part: constraint systems are assembled on the host; sizes must equal the counts SURVEY section 4 derived from the Rust
sources, every constraint must be satisfied by the witness, and the verifier must rebuild the very same rows.
GPU part: `prover <stem>` -> .coms / .proof -> `verifier <stem>` prints true for all 13 fixtures (the reference's CI,
.github/workflows/integration_tests.yml:19-58) and false for falsified statements (the reference's is_err unit cases)."""
import os
import shutil

import pytest

import oracle_lib as ol
from circuits import host_witness
from oracle import pyref as pr

L = pr.L
FX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fixtures")
# name -> (multipliers, commitments) : SURVEY.md section 4 "Fixture circuit sizes"
SIZES = {"example": (14988, 33), "bounds_check": (1440, 9), "equality": (0, 12), "inequality": (24, 36), "less_than": (1137, 12),
         "merkle_tree": (27216, 31), "mimc_hash": (18468, 32), "set_membership": (15600, 72), "or": (9452, None), "or2": (17236, None),
         "or3": (3, 3), "or4": (22561, None), "or5": (4721, None)}


def rd(name):
    with open(os.path.join(FX, name)) as f:
        return f.read()


def unsatisfied_rows(p):
    aL, aR, aO = host_witness(p)
    vals = {0: aL, 1: aR, 2: aO, 3: p.v}
    rp, tv, tc = p.csr()
    bad = 0
    for r in range(len(rp) - 1):
        acc = 0
        for k in range(rp[r], rp[r + 1]):
            kind, idx = tv[k] >> 29, tv[k] & 0x1FFFFFFF
            acc += int.from_bytes(tc[32 * k:32 * k + 32], "little") * (1 if kind == 4 else vals[kind][idx])
        bad += (acc % L != 0)
    return bad


@pytest.mark.parametrize("name", sorted(SIZES))
def test_fixture_assembles_like_the_reference(name):
    from bulletproofs_gadgets_b200 import frontend as fe
    run = fe.ProverRun(name.encode(), rd(name + ".gadgets"), rd(name + ".inst"), rd(name + ".wtns"), test_seed=1)
    p = run.prover
    n, m = SIZES[name]
    assert p.get_num_multiplications() == n
    if m is not None:
        assert len(p.v) == m
    assert unsatisfied_rows(p) == 0
    coms = "".join("%s = 0x%s\n" % (nm, "00" * 32) for nm in run.coms_names)
    vr = fe.VerifierRun(name.encode(), rd(name + ".gadgets"), rd(name + ".inst"), coms)
    assert vr.verifier.csr() == p.csr()
    assert len(set(run.coms_names)) == len(run.coms_names)
    assert all(nm[0] in "CD" for nm in run.coms_names)


def test_falsified_statements_leave_constraints_unsatisfied():
    from bulletproofs_gadgets_b200 import frontend as fe
    cases = [("EQUALS W0 W1\n", "", "W0 = 0x43\nW1 = 0x44\n"),
             ("UNEQUAL W0 W1\n", "", "W0 = 0x43\nW1 = 0x43\n"),
             ("LESS_THAN W0 W1\n", "", "W0 = 0x44\nW1 = 0x43\n"),
             ("LESS_THAN W0 W1\n", "", "W0 = 0x43\nW1 = 0x43\n"),
             ("BOUND W0 I0 I1\n", "I0 = 0x11\nI1 = 0x64\n", "W0 = 0x65\n"),
             ("SET_MEMBER W0 I0 I1\n", "I0 = 0x11\nI1 = 0x64\n", "W0 = 0x65\n"),
             ("HASH W1 W0\n", "", "W0 = 0x43\nW1 = 0x0cfb0c17618211c607febf703ac3f3078f7d96798fae9d4a1682bc592f7cb127\n"),
             ("OR\n[\n{\nEQUALS W0 W1\n}\n{\nLESS_THAN W1 W0\n}\n]\n", "", "W0 = 0x43\nW1 = 0x44\n")]
    for gadgets, inst, wtns in cases:
        run = fe.ProverRun(b"neg", gadgets, inst, wtns, test_seed=2)
        assert unsatisfied_rows(run.prover) > 0, gadgets
    # ... and the true versions are satisfied (incl. an OR whose first clause is false)
    for gadgets, inst, wtns in [("OR\n[\n{\nEQUALS W0 W1\n}\n{\nLESS_THAN W0 W1\n}\n]\n", "", "W0 = 0x43\nW1 = 0x44\n"),
                                ("HASH W1 W0\n", "", "W0 = 0x43\nW1 = 0x0cfb0c17618211c607febf703ac3f3078f7d96798fae9d4a1682bc592f7cb126\n"),
                                ("SET_MEMBER W0 I0 I1\n", "I0 = 0x11\nI1 = 0x64\n", "W0 = 0x64\n")]:
        run = fe.ProverRun(b"pos", gadgets, inst, wtns, test_seed=2)
        assert unsatisfied_rows(run.prover) == 0, gadgets


def test_merkle_pattern_parser_matches_grammar_ordering():
    from bulletproofs_gadgets_b200 import frontend as fe
    import re
    inst, wtns, pat = fe.parse_merkle_tree(re.findall(r"[()]|[WI]\d+", "((W1 I3) (I6 W4))"))
    assert (inst, wtns, pat) == (["I3", "I6"], ["W1", "W4"], (("W", "I"), ("I", "W")))
    inst, wtns, pat = fe.parse_merkle_tree(re.findall(r"[()]|[WI]\d+", "(W0 ((I1 W2) I3))"))
    assert (inst, wtns, pat) == (["I1", "I3"], ["W0", "W2"], ("W", (("I", "W"), "I")))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SIZES))
def test_cli_fixture_prover_then_verifier(name, tmp_path):
    """the reference's CI: prover <stem>; verifier <stem> -> true, on the GPU path, through the file formats"""
    import bulletproofs_gadgets_b200 as bpg
    from bulletproofs_gadgets_b200 import frontend as fe
    ctx = bpg.Context.default()
    for ext in (".gadgets", ".inst", ".wtns"):
        shutil.copy(os.path.join(FX, name + ext), str(tmp_path / (name + ext)))
    stem = str(tmp_path / name)
    nc = fe.prover_main(stem, test_seed=7, ext_rng32=b"\x17" * 32, ctx=ctx, label="fixtures/" + name)
    assert nc > 0
    assert os.path.getsize(stem + ".proof") % 32 == 1
    assert fe.verifier_main(stem, ctx=ctx, label="fixtures/" + name) is True
    # the verifier must be given the same label string as the prover (prover.rs:52 / verifier.rs:51)
    assert fe.verifier_main(stem, ctx=ctx, label="fixtures/other") is False
    # a flipped proof byte -> false
    with open(stem + ".proof", "rb") as f:
        proof = bytearray(f.read())
    proof[len(proof) // 2] ^= 1
    with open(stem + ".proof", "wb") as f:
        f.write(bytes(proof))
    assert fe.verifier_main(stem, ctx=ctx, label="fixtures/" + name) is False


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SIZES))
def test_device_witness_evaluation_matches_big_integers(name):
    """Prover::eval of every multiply() (cs_buffer.rs:94-97 replayed by prover.rs:102-117) runs on the device,
    level by level; a_L, a_R, a_O must equal the host big-integer evaluation for every reference fixture"""
    import bulletproofs_gadgets_b200 as bpg
    from bulletproofs_gadgets_b200 import frontend as fe
    run = fe.ProverRun(name.encode(), rd(name + ".gadgets"), rd(name + ".inst"), rd(name + ".wtns"), test_seed=1, ctx=bpg.Context.default())
    assert run.prover.witness() == host_witness(run.prover)


@pytest.mark.gpu
def test_device_witness_evaluation_edge_cases():
    """no multipliers; only assigned multipliers; a pending half-allocated multiplier; unreduced inputs; a deep chain; a
    combination referencing a LATER multiplier is rejected (BPG_E_ARG), never evaluated"""
    import ctypes as C
    import random

    import bulletproofs_gadgets_b200 as bpg
    ctx = bpg.Context.default()
    rnd = random.Random(5)
    p = bpg.Prover.new(b"w", ctx=ctx)
    assert p.witness() == ([], [], [])
    _, v0 = p.commit(rnd.randrange(L), 1)
    p.allocate_multiplier((L - 1, L - 1))
    x = p.allocate(7)
    assert p.witness() == host_witness(p) and p.witness()[2] == [1, 0]
    p.allocate(9)
    cur = [(x, 3), (v0, L - 2), (bpg.api.ONE, rnd.randrange(L))]
    for _ in range(300):
        _, r, o = p.multiply(cur, cur + [(bpg.api.ONE, 1)])
        cur = [(o, rnd.randrange(L)), (r, 5), (v0, 1)]
    w = p.witness()
    assert w == host_witness(p) and w[2][1] == 63
    n = 3
    ptr = (C.c_uint32 * (2 * n + 1))(0, 0, 0, 1, 2, 2, 2)
    tv = (C.c_uint32 * 2)((0 << 29) | 2, 4 << 29)  # multiplier 1 references a_L[2]
    bufs = [C.create_string_buffer(32 * n) for _ in range(3)]
    assert ctx.lib.bpg_witness_eval(ctx.h, n, 0, ptr, tv, bytes(64), None, *bufs) == -4  # BPG_E_ARG
    bad = (C.c_uint32 * (2 * n + 1))(0, 0, 0, 2, 1, 2, 2)  # a middle pointer outside its multiplier's range
    assert ctx.lib.bpg_witness_eval(ctx.h, n, 0, bad, (C.c_uint32 * 2)(4 << 29, 4 << 29), bytes(64), None, *bufs) == -4


@pytest.mark.gpu
def test_cli_falsified_statements_are_rejected_and_oracle_agrees(tmp_path):
    """verdict parity on statements that are false (the reference's is_err unit cases): the GPU verifier and the CPU oracle
    verifier both reject; the GPU proof bytes of a true statement equal the oracle prover's bytes"""
    import bulletproofs_gadgets_b200 as bpg
    from bulletproofs_gadgets_b200 import frontend as fe
    ctx = bpg.Context.default()
    cases = [("EQUALS W0 W1\n", "", "W0 = 0x43\nW1 = 0x44\n", False),
             ("EQUALS W0 W1\n", "", "W0 = 0x43\nW1 = 0x43\n", True),
             ("UNEQUAL W0 W1\n", "", "W0 = 0x43\nW1 = 0x43\n", False),
             ("LESS_THAN W0 W1\n", "", "W0 = 0x44\nW1 = 0x43\n", False),
             ("LESS_THAN W0 W1\n", "", "W0 = 0x42\nW1 = 0x43\n", True),
             ("BOUND W0 I0 I1\n", "I0 = 0x11\nI1 = 0x64\n", "W0 = 0x65\n", False),
             ("SET_MEMBER W0 I0 I1\n", "I0 = 0x11\nI1 = 0x64\n", "W0 = 0x65\n", False),
             ("OR\n[\n{\nEQUALS W0 W1\n}\n{\nLESS_THAN W1 W0\n}\n]\n", "", "W0 = 0x43\nW1 = 0x44\n", False),
             ("OR\n[\n{\nEQUALS W0 W1\n}\n{\nLESS_THAN W0 W1\n}\n]\n", "", "W0 = 0x43\nW1 = 0x44\n", True)]
    for k, (gadgets, inst, wtns, want) in enumerate(cases):
        run = fe.ProverRun(b"case", gadgets, inst, wtns, test_seed=k, ctx=ctx)
        coms, proof, _ = run.finish(ext_rng32=bytes([k]) * 32)
        vr = fe.VerifierRun(b"case", gadgets, inst, coms, ctx=ctx)
        assert vr.finish(proof) is want, gadgets
        # the oracle on the same constraint system / witness / randomness
        p = run.prover
        rp, tv, tc = p.csr()
        enc = lambda xs: b"".join(int(x).to_bytes(32, "little") for x in xs)
        aL, aR, aO = p.witness_bytes()  # evaluated on the device (bpg_witness_eval) ...
        assert (aL, aR, aO) == tuple(enc(x) for x in host_witness(p))  # ... equal to the big-integer evaluation
        oproof, oV = ol.r1cs_prove(b"case", 1 << 12, aL, aR, aO, enc(p.v), enc(p.v_blinding), rp, tv, tc, bytes([k]) * 32)
        assert oproof == proof
        assert ol.r1cs_verify(b"case", 1 << 12, p.get_num_multiplications(), oV, rp, tv, tc, proof, bytes(32)) is want
