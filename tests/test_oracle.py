"""CPU tests: pin the oracle (oracle/pyref.py big-int + oracle/bpo.c) against every known answer the
reference holds for the hot path and against public vectors (RFC 9496, Merlin),
then check the C oracle against the Python one byte-for-byte (scalars, points, MSM algorithms, proofs)."""
import hashlib
import json
import os
import random

import pytest

import circuits
import oracle_lib as ol
from oracle import pyref as pr

L = pr.L
GOLD = os.path.join(os.path.dirname(__file__), "golden")
with open(os.path.join(GOLD, "mimc_consts.json")) as f:
    pr.set_mimc_constants([int.from_bytes(bytes.fromhex(h), "little") & ((1 << 255) - 1) for h in json.load(f)])
with open(os.path.join(GOLD, "mimc_kats.json")) as f:
    KATS = json.load(f)
with open(os.path.join(GOLD, "merkle_fixtures.json")) as f:
    MERKLE = json.load(f)

RISTRETTO_MULTIPLES = [
    "00" * 32,
    "e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76",
    "6a493210f7499cd17fecb510ae0cea23a110e8d5b901f8acadd3095c73a3b919",
    "94741f5d5d52755ece4f23f044ee27d5d1ea1e2bd196b462166b16152a9d0259",
    "da80862773358b466ffadfe0b3293ab3d9fd53c5ea6c955358f568322daf6a57",
    "e882b131016b52c1d3337080187cf768423efccbb517bb495ab812c4160ff44e",
]
# RFC 9496 A.3: encodings that must be rejected
BAD_ENCODINGS = [
    "00ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff",
    "ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f",
    "f3ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f",
    "edffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f",
    "0100000000000000000000000000000000000000000000000000000000000000",
    "01ffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f",
    "ed57ffd8c914fb201471d1c3d245ce3c746fcbe63a3679d51b6a516ebebe0e20",
    "c34c4e1826e5d403b78e246e88aa051c36ccf0aafebffe137d148a2bf9104562",
    "c940e5a4404157cfb1628b108db051a8d439e1a421394ec4ebccb9ec92a8ac78",
    "47cfc5497c53dc8e61c91d17fd626ffb1c49e2bca94eed052281b510b1117a24",
    "f1c6165d33367351b0da8f6e4511010c68174a03b6581212c71c0e1d026c3c72",
    "87260f7a2f12495118360f02c26a470f450dadf34a413d21042b43b9d93e1309",
    "26948d35ca62e643e26a83177332e6b6afeb9d08e4268b650f1f5bbd8d81d371",
    "4eac077a713c57b4f4397629a4145982c661f48044dd3f96427d40b147d9742f",
]


def rs(rnd):
    return rnd.randrange(2 ** 255).to_bytes(32, "little")


# ----------------------------------------------------------------------------- MiMC / Merkle (reference KATs)
def test_mimc_hash_kats_python():
    assert len(KATS["hash"]) >= 10
    for pre, dig, src in KATS["hash"]:
        assert pr.sc_bytes(pr.mimc_hash(bytes.fromhex(pre)))[::-1].hex() == dig, src


def test_mimc_hash_kats_c():
    for pre, dig, src in KATS["hash"]:
        assert ol.mimc_hash(bytes.fromhex(pre))[::-1].hex() == dig, src


def test_merkle_node_kats_both():
    assert len(KATS["node"]) >= 15
    for l, r, dig, src in KATS["node"]:
        lb, rb = bytes.fromhex(l)[::-1].ljust(32, b"\0"), bytes.fromhex(r)[::-1].ljust(32, b"\0")
        assert ol.mimc_sponge(lb + rb)[::-1].hex() == dig, src
        assert pr.sc_bytes(pr.merkle_node(pr.sc_from_bits(lb), pr.sc_from_bits(rb)))[::-1].hex() == dig, src


def _merkle_root(pattern, values):
    """CLI path: leaves are MiMC-hashed first (prover.rs:324-334), inner nodes = unpadded 2-block sponge."""
    toks = pattern.replace("(", " ( ").replace(")", " ) ").split()

    def parse(i):
        if toks[i] == "(":
            l, i = parse(i + 1)
            r, i = parse(i)
            assert toks[i] == ")"
            return ol.mimc_sponge(l + r), i + 1
        return ol.mimc_hash(bytes.fromhex(values[toks[i]])), i + 1

    root, _ = parse(0)
    return root


def test_merkle_fixture_roots():
    for fx in MERKLE:
        if fx.get("leaves_hashed", True) and all(k in fx["values"] for k in fx["pattern"].replace("(", " ").replace(")", " ").split()):
            assert _merkle_root(fx["pattern"], fx["values"])[::-1].hex() == fx["root"], fx["source"]


def test_mimc_trace_matches_python():
    rnd = random.Random(5)
    blocks = [rnd.randrange(L) for _ in range(2)]
    dig, tr = ol.mimc_sponge(b"".join(pr.sc_bytes(b) for b in blocks), trace=True)
    aL, aR, aO, d = pr.mimc_trace(blocks)
    assert dig == pr.sc_bytes(d)
    flat = b"".join(pr.sc_bytes(aL[i]) + pr.sc_bytes(aR[i]) + pr.sc_bytes(aO[i]) for i in range(len(aL)))
    assert tr == flat and len(aL) == 2 * 972


# ----------------------------------------------------------------------------- scalars
def test_scalar_ops_c_vs_bigint():
    rnd = random.Random(1)
    edge = [0, 1, L - 1, L, L + 1, 2 * L, 2 ** 255 - 1, 2 ** 252, 2 ** 252 - 1]
    vals = edge + [rnd.randrange(2 ** 255) for _ in range(300)]
    for a in vals:
        for b in (vals[rnd.randrange(len(vals))], edge[rnd.randrange(len(edge))]):
            ab, bb = a.to_bytes(32, "little"), b.to_bytes(32, "little")
            assert ol.sc_mul(ab, bb) == pr.sc_bytes(a * b % L)
            assert ol.sc_add(ab, bb) == pr.sc_bytes((a + b) % L)
        assert ol.sc_reduce(a.to_bytes(32, "little")) == pr.sc_bytes(a % L)
    for a in vals[1:40]:
        if a % L:
            assert ol.sc_invert(a.to_bytes(32, "little")) == pr.sc_bytes(pr.sc_inv(a))
    for w in (b"\xff" * 64, b"\0" * 64, rnd.randbytes(64), rnd.randbytes(64)):
        assert ol.sc_wide(w) == pr.sc_bytes(pr.sc_wide(w))


# ----------------------------------------------------------------------------- ristretto255 (RFC 9496) + generators
def test_ristretto_basepoint_multiples():
    B, Bb = ol.pedersen_gens()
    assert B.hex() == RISTRETTO_MULTIPLES[1]
    assert Bb.hex() == "8c9240b456a9e6dc65c377a1048d745f94a08cdb7f44cbcd7b46f34048871134"  # dalek's published B_blinding
    assert pr.ristretto_encode(pr.B_BLINDING) == Bb
    for k, h in enumerate(RISTRETTO_MULTIPLES):
        assert ol.point_mul(k.to_bytes(32, "little"), B).hex() == h
        assert pr.ristretto_encode(pr.pt_mul(k, pr.BASEPOINT)).hex() == h


def test_ristretto_bad_encodings_rejected():
    for h in BAD_ENCODINGS:
        b = bytes.fromhex(h)
        assert pr.ristretto_decode(b) is None, h
        assert ol.lib().bpo_point_decode_ok(b) == 0, h


def test_elligator_vector_and_random():
    v = hashlib.sha512(b"Ristretto is traditionally a short shot of espresso coffee").digest()
    want = "3066f82a1a747d45120d1740f14358531a8f04bbffe6a819f86dfe50f44a0a46"
    assert ol.from_uniform_bytes(v).hex() == want
    assert pr.ristretto_encode(pr.from_uniform_bytes(v)).hex() == want
    rnd = random.Random(2)
    for _ in range(20):
        b = rnd.randbytes(64)
        assert ol.from_uniform_bytes(b) == pr.ristretto_encode(pr.from_uniform_bytes(b))


def test_generators_chain():
    G, H = ol.gens(0, 6)
    assert G[:32].hex() == "fc3b25801422672a6a8d3adb5d8457d4301fe92324b4fc56ae934c8713ddfe2d"
    assert G[32:64].hex() == "ae817fdef62f713dd169dc8a26406f68be0bd3cd53652614636b0801567c4264"
    assert H[:32].hex() == "ba698f6dd08c501e32b55d2ee7259f6019d629fa2ba4d7039c5de157cba4df73"
    Gp, Hp = pr.bulletproof_gens(6)
    assert b"".join(map(pr.ristretto_encode, Gp)) == G and b"".join(map(pr.ristretto_encode, Hp)) == H
    G2, _ = ol.gens(3, 2)
    assert G2 == G[96:160]


def test_group_law_vs_libsodium():
    nacl = pytest.importorskip("nacl.bindings")
    rnd = random.Random(3)
    # Ed25519 standard basepoint (y = 4/5, x even): libsodium only accepts prime-order-subgroup points, and
    # ristretto decode() of the basepoint encoding is merely a coset representative.
    y = 4 * pr.inv(5) % pr.P
    x = pr.sqrt_ratio_m1((y * y - 1) % pr.P, (pr.D * y * y + 1) % pr.P)[1]
    B = (x, y, 1, x * y % pr.P)
    assert pr.pt_eq(B, pr.BASEPOINT)

    def ed_bytes(p):
        X, Y, Z, _ = p
        zi = pr.inv(Z)
        x, y = X * zi % pr.P, Y * zi % pr.P
        return (y | ((x & 1) << 255)).to_bytes(32, "little")

    for _ in range(5):
        a, b = rnd.randrange(1, L), rnd.randrange(1, L)
        pa, pb = pr.pt_mul(a, B), pr.pt_mul(b, B)
        assert nacl.crypto_core_ed25519_add(ed_bytes(pa), ed_bytes(pb)) == ed_bytes(pr.pt_add(pa, pb))
        assert nacl.crypto_scalarmult_ed25519_noclamp(pr.sc_bytes(b), ed_bytes(pa)) == ed_bytes(pr.pt_mul(b, pa))
        # C oracle agrees with the big-int model on the same operations (ristretto encodings)
        ea, eb = pr.ristretto_encode(pa), pr.ristretto_encode(pb)
        assert ol.point_add(ea, eb) == pr.ristretto_encode(pr.pt_add(pa, pb))
        assert ol.point_mul(pr.sc_bytes(b), ea) == pr.ristretto_encode(pr.pt_mul(b, pa))


# ----------------------------------------------------------------------------- Merlin
def test_merlin_vector():
    want = "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"
    t = ol.Transcript(b"test protocol")
    t.append(b"some label", b"some data")
    assert t.challenge(b"challenge", 32).hex() == want
    t = pr.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == want


def test_merlin_long_messages_c_vs_python():
    rnd = random.Random(4)
    tc, tp = ol.Transcript(b"x"), pr.Transcript(b"x")
    for n in (0, 1, 165, 166, 167, 400):
        m = rnd.randbytes(n)
        tc.append(b"lab", m)
        tp.append_message(b"lab", m)
        assert tc.challenge(b"c", 200) == tp.challenge_bytes(b"c", 200)


# ----------------------------------------------------------------------------- MSM / Pedersen / fold
def test_msm_algorithms_agree():
    rnd = random.Random(6)
    n = 600
    sc = b"".join(rs(rnd) for _ in range(n))
    G, _ = ol.gens(0, n)
    ref20 = pr.ristretto_encode(pr.msm([int.from_bytes(sc[32 * i:32 * i + 32], "little") for i in range(20)], pr.bulletproof_gens(20)[0]))
    for algo in (ol.NAIVE, ol.STRAUS_CT, ol.VARTIME):
        assert ol.msm(sc[:640], G[:640], algo) == ref20
    for m in (1, 2, 189, 190, 499, 500, 600):  # dalek's algorithm switch points
        r = ol.msm(sc[:32 * m], G[:32 * m], ol.NAIVE)
        assert r == ol.msm(sc[:32 * m], G[:32 * m], ol.VARTIME) == ol.msm(sc[:32 * m], G[:32 * m], ol.STRAUS_CT)
    ol.lib().bpo_set_threads(4)
    try:
        assert ol.msm(sc, G, ol.VARTIME) == ol.msm(sc, G, ol.NAIVE)
    finally:
        ol.lib().bpo_set_threads(1)
    assert ol.msm(b"", b"", ol.VARTIME) == bytes(32)
    assert ol.msm(sc[:32], bytes.fromhex(BAD_ENCODINGS[4]), ol.VARTIME) is None


def test_msm_gens_and_extras():
    rnd = random.Random(7)
    n = 40
    sG, sH = b"".join(rs(rnd) for _ in range(n)), b"".join(rs(rnd) for _ in range(n))
    G, H = ol.gens(8, n)
    B, Bb = ol.pedersen_gens()
    es = rs(rnd)
    assert ol.msm_gens(sG, sH, n, 8, es, Bb) == ol.msm(sG + sH + es, G + H + Bb, ol.NAIVE)
    assert ol.msm_gens(sG, None, n, 8) == ol.msm(sG, G, ol.NAIVE)


def test_pedersen_and_fold():
    rnd = random.Random(8)
    v, r = [rnd.randrange(2 ** 255) for _ in range(5)], [rnd.randrange(2 ** 255) for _ in range(5)]
    got = ol.pedersen_commit(b"".join(x.to_bytes(32, "little") for x in v), b"".join(x.to_bytes(32, "little") for x in r))
    assert got == b"".join(pr.ristretto_encode(pr.pedersen_commit(a, b)) for a, b in zip(v, r))
    G, H = ol.gens(0, 8)
    u = rnd.randrange(L)
    ui = pr.sc_inv(u)
    out = ol.fold_points(pr.sc_bytes(ui), pr.sc_bytes(u), G[:128], G[128:])
    Gp = pr.bulletproof_gens(8)[0]
    assert out == b"".join(pr.ristretto_encode(pr.msm([ui, u], [Gp[i], Gp[4 + i]])) for i in range(4))


# ----------------------------------------------------------------------------- R1CS proofs: C == Python, byte for byte
@pytest.mark.parametrize("nmul,seed", [(0, 1), (1, 2), (2, 3), (3, 4), (5, 5), (8, 6), (13, 7)])
def test_r1cs_proof_bytes_c_vs_python(nmul, seed):
    inst = circuits.chain_instance(nmul, seed)
    t = pr.Transcript(inst["label"])
    cs = pr.ConstraintSystem(t, True)
    Vv = [cs.commit(v, b)[1] for v, b in zip(inst["ivals"], inst["iblinds"])]
    circuits.chain_wire(cs, Vv, seed + 1000, nmul)
    ext = bytes(range(32))
    proof, _ = pr.r1cs_prove(cs, 64, ext)
    rp, tv, tc = inst["csr"]
    proof_c, V = ol.r1cs_prove(inst["label"], 64, inst["aL"], inst["aR"], inst["aO"], inst["vals"], inst["blinds"], rp, tv, tc, ext)
    assert V == b"".join(cs.V)
    assert proof_c == proof
    # verifier: both accept; tampering / wrong label / non-canonical scalar rejected
    tvr = pr.Transcript(inst["label"])
    csv = pr.ConstraintSystem(tvr, False)
    Vv = [csv.commit_verifier(V[32 * i:32 * i + 32]) for i in range(3)]
    circuits.chain_wire(csv, Vv, seed + 1000, nmul)
    assert pr.r1cs_verify(csv, proof, 64, b"\x07" * 32)
    args = (64, nmul, V, rp, tv, tc)
    assert ol.r1cs_verify(inst["label"], *args, proof, b"\x07" * 32)
    for pos in (1 + 3, 1 + 32 * 8 + 3, len(proof) - 1, len(proof) - 40):
        bad = bytearray(proof)
        bad[pos] ^= 0x10
        assert not ol.r1cs_verify(inst["label"], *args, bytes(bad), b"\x07" * 32)
    assert not ol.r1cs_verify(b"other", *args, proof, b"\x07" * 32)
    assert not ol.r1cs_verify(inst["label"], *args, proof[:-32], b"\x07" * 32)
    p2, _ = ol.r1cs_prove(inst["label"], 64, inst["aL"], inst["aR"], inst["aO"], inst["vals"], inst["blinds"], rp, tv, tc, ext, flags=1)
    assert ol.r1cs_verify(inst["label"], *args, p2, b"\x07" * 32, flags=1)


def test_r1cs_unsatisfied_witness_rejected():
    inst = circuits.chain_instance(6, 11, wrong=True)
    rp, tv, tc = inst["csr"]
    proof, V = ol.r1cs_prove(inst["label"], 8, inst["aL"], inst["aR"], inst["aO"], inst["vals"], inst["blinds"], rp, tv, tc, bytes(32))
    assert not ol.r1cs_verify(inst["label"], 8, 6, V, rp, tv, tc, proof, bytes(32))
    with pytest.raises(ValueError):
        ol.r1cs_prove(inst["label"], 4, inst["aL"], inst["aR"], inst["aO"], inst["vals"], inst["blinds"], rp, tv, tc, bytes(32))


def test_r1cs_dense_instance_multithreaded():
    inst = circuits.random_dense_instance(300, 21)
    rp, tv, tc = inst["csr"]
    p1, V = ol.r1cs_prove(inst["label"], 512, inst["aL"], inst["aR"], inst["aO"], inst["vals"], inst["blinds"], rp, tv, tc, bytes(32))
    ol.lib().bpo_set_threads(4)
    try:
        p2, _ = ol.r1cs_prove(inst["label"], 512, inst["aL"], inst["aR"], inst["aO"], inst["vals"], inst["blinds"], rp, tv, tc, bytes(32))
        assert ol.r1cs_verify(inst["label"], 512, 300, V, rp, tv, tc, p1, bytes(32))
    finally:
        ol.lib().bpo_set_threads(1)
    assert p1 == p2
