"""CPU test of the product's limb arithmetic (portable bodies of csrc/fe25519.cuh, ge25519.cuh) against
the big-int oracle.  The PTX carry-chain bodies of the same functions are exercised by the -m gpu tests."""
import ctypes as C
import os
import random
import subprocess

import pytest

from oracle import pyref as pr

HERE = os.path.dirname(os.path.abspath(__file__))
P, Lo = pr.P, pr.L


@pytest.fixture(scope="module")
def ht():
    src = os.path.join(HERE, "native", "host_arith.cpp")
    so = os.path.join(HERE, "native", "libhost_arith.so")
    hdrs = [os.path.join(HERE, "..", "bulletproofs_gadgets_b200", "csrc", h) for h in ("fe25519.cuh", "ge25519.cuh", "consts.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(p) for p in [src] + hdrs):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-shared", "-fPIC", "-o", so, src])
    lib = C.CDLL(so)
    assert lib.ht_init() == 0
    return lib


def b(x):
    return x.to_bytes(32, "little")


def val(buf):
    return int.from_bytes(buf.raw[:32], "little")


def test_constants_derived_match_bigint(ht):
    buf = C.create_string_buffer(7 * 32)
    ht.ht_consts(buf)
    want = [pr.D, pr.D2, pr.SQRT_M1, pr.INVSQRT_A_MINUS_D, pr.SQRT_AD_MINUS_ONE, pr.ONE_MINUS_D_SQ, pr.D_MINUS_ONE_SQ]
    for i, w in enumerate(want):
        assert int.from_bytes(buf.raw[32 * i:32 * i + 32], "little") == w, i
    # decimal constants recorded in SURVEY.md App. A.2
    assert want[2] == 19681161376707505956807079304988542015446066515923890162744021073123829784752
    assert want[4] == 25063068953384623474111414158702152701244531502492656460079210482610430750235


def test_field_ops_weakly_reduced_inputs(ht):
    rnd = random.Random(1)
    edge = [0, 1, 2, 19, 38, P - 1, P, P + 1, 2 * P, 2 * P + 1, 2 ** 256 - 1, 2 ** 256 - 38, 2 ** 256 - 39, 2 ** 255,
            2 ** 255 - 1, 2 ** 255 - 19, 2 ** 255 + 18, 2 ** 255 + 19, 2 ** 256 - 2 ** 32, 2 ** 224 - 1]
    vals = edge + [rnd.randrange(2 ** 256) for _ in range(400)]
    o = C.create_string_buffer(32)
    for x in vals:
        for y in (rnd.choice(vals), rnd.choice(edge), x):
            ht.ht_fe_mul(b(x), b(y), o)
            assert val(o) == x * y % P, (x, y)
            ht.ht_fe_sqr(b(x), o)
            assert val(o) == x * x % P, x
            ht.ht_fe_add(b(x), b(y), o)
            assert val(o) == (x + y) % P, (x, y)
            ht.ht_fe_sub(b(x), b(y), o)
            assert val(o) == (x - y) % P, (x, y)
    for x in (1, 2, 12345, P - 1, rnd.randrange(P)):
        ht.ht_fe_inv(b(x), o)
        assert val(o) == pr.inv(x)


def test_scalar_ops(ht):
    rnd = random.Random(2)
    edge = [0, 1, Lo - 1, Lo, Lo + 1, 2 * Lo, 2 ** 252, 2 ** 252 - 1, 2 ** 255 - 1, 2 ** 256 - 1]
    vals = edge + [rnd.randrange(2 ** 256) for _ in range(400)]
    o, o2 = C.create_string_buffer(32), C.create_string_buffer(32)
    for x in vals:
        for y in (rnd.choice(vals), rnd.choice(edge)):
            ht.ht_sc_mul(b(x), b(y), o)
            assert val(o) == x * y % Lo, (x, y)
        ht.ht_sc_reduce(b(x), o)
        assert val(o) == x % Lo
        xr, yr = x % Lo, rnd.choice(vals) % Lo
        ht.ht_sc_addsub(b(xr), b(yr), o, o2)
        assert val(o) == (xr + yr) % Lo and val(o2) == (xr - yr) % Lo
    for w in [b"\xff" * 64, bytes(64)] + [rnd.randbytes(64) for _ in range(20)]:
        ht.ht_sc_wide(w, o)
        assert val(o) == int.from_bytes(w, "little") % Lo
    for x in (1, 2, 987654321, Lo - 1, rnd.randrange(Lo)):
        ht.ht_sc_invert(b(x), o)
        assert val(o) == pr.sc_inv(x)


def test_host_scalar64_fast_path(ht):
    """csrc/host_scalar64.h (4 x 64-bit limb products used by the protocol drivers on the host)"""
    rnd = random.Random(4)
    edge = [0, 1, Lo - 1, Lo, Lo + 1, 2 ** 252, 2 ** 255 - 1, 2 ** 256 - 1]
    vals = edge + [rnd.randrange(2 ** 256) for _ in range(300)]
    om, oi, ow = C.create_string_buffer(32), C.create_string_buffer(32), C.create_string_buffer(32)
    for x in vals:
        y = rnd.choice(vals)
        w = rnd.choice([b"\xff" * 64, bytes(64), rnd.randbytes(64)])
        ht.ht_sc64(b(x), b(y), w, om, oi, ow)
        assert val(om) == x * y % Lo
        assert val(oi) == pow(x % Lo, Lo - 2, Lo)
        assert val(ow) == int.from_bytes(w, "little") % Lo


def test_group_ops_and_ristretto(ht):
    rnd = random.Random(3)
    o = C.create_string_buffer(32)
    o128 = C.create_string_buffer(128)
    for _ in range(8):
        k, k2 = rnd.randrange(Lo), rnd.randrange(Lo)
        Pp, Q = pr.pt_mul(k2, pr.BASEPOINT), pr.B_BLINDING
        assert ht.ht_point_mul_add(b(k), pr.ristretto_encode(Pp), pr.ristretto_encode(Q), o128) == 1
        R = pr.pt_add(pr.pt_mul(k, Pp), Q)
        assert o128.raw[:32] == pr.ristretto_encode(R)            # projective-Niels add
        assert o128.raw[32:64] == pr.ristretto_encode(pr.pt_dbl(R))  # doubling
        assert o128.raw[64:96] == pr.ristretto_encode(R)          # affine-Niels add
        assert o128.raw[96] == 1                                   # negation + coset identity test
        o64 = C.create_string_buffer(64)
        assert ht.ht_ilp(pr.ristretto_encode(Pp), pr.ristretto_encode(Q), o64) == 1
        assert o64.raw[:32] == pr.ristretto_encode(pr.pt_add(Pp, Q)) and o64.raw[32:] == pr.ristretto_encode(pr.pt_dbl(Pp))
        w = rnd.randbytes(64)
        ht.ht_from_uniform(w, o)
        assert o.raw == pr.ristretto_encode(pr.from_uniform_bytes(w))
        e = pr.ristretto_encode(R)
        assert ht.ht_decode_encode(e, o) == 1 and o.raw == e
    assert ht.ht_decode_encode(bytes(32), o) == 1 and o.raw == bytes(32)
    for bad in (b"\x01" + bytes(31), b"\xff" * 32, bytes.fromhex("edffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f")):
        assert ht.ht_decode_encode(bad, o) == 0
