"""ctypes loader for the C CPU oracle (oracle/libbpo.so).  TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None


def build():
    subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(ORACLE_DIR, "libbpo.so")
    src = os.path.join(ORACLE_DIR, "bpo.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        build()
    L = C.CDLL(so)
    u8p, sz, i32 = C.c_char_p, C.c_size_t, C.c_int
    u32p = C.POINTER(C.c_uint32)
    sig = {
        "bpo_set_mimc_constants": (None, [u8p]),
        "bpo_set_threads": (None, [i32]),
        "bpo_sc_reduce": (None, [u8p, u8p]), "bpo_sc_wide": (None, [u8p, u8p]),
        "bpo_sc_mul": (None, [u8p, u8p, u8p]), "bpo_sc_add": (None, [u8p, u8p, u8p]),
        "bpo_sc_invert": (None, [u8p, u8p]),
        "bpo_point_decode_ok": (i32, [u8p]), "bpo_point_add": (i32, [u8p, u8p, u8p]),
        "bpo_point_mul": (i32, [u8p, u8p, u8p]), "bpo_from_uniform_bytes": (None, [u8p, u8p]),
        "bpo_pedersen_gens": (None, [u8p, u8p]), "bpo_gens": (None, [sz, sz, u8p, u8p]),
        "bpo_pedersen_commit": (None, [u8p, u8p, sz, u8p]),
        "bpo_msm": (i32, [u8p, u8p, sz, u8p, i32]),
        "bpo_msm_gens": (i32, [u8p, u8p, sz, sz, u8p, u8p, sz, u8p, i32]),
        "bpo_fold_points": (i32, [u8p, u8p, u8p, u8p, sz, u8p]),
        "bpo_transcript_new": (C.c_void_p, [u8p, sz]), "bpo_transcript_free": (None, [C.c_void_p]),
        "bpo_transcript_append": (None, [C.c_void_p, u8p, sz, u8p, sz]),
        "bpo_transcript_challenge": (None, [C.c_void_p, u8p, sz, u8p, sz]),
        "bpo_mimc_hash": (None, [u8p, sz, u8p]), "bpo_mimc_sponge": (None, [u8p, sz, u8p, u8p]),
        "bpo_r1cs_prove": (C.c_long, [u8p, sz, sz, sz, u8p, u8p, u8p, sz, u8p, u8p, sz, u32p, u32p, u8p, u8p, i32, u8p, u8p, sz]),
        "bpo_r1cs_verify": (i32, [u8p, sz, sz, sz, sz, u8p, sz, u32p, u32p, u8p, u8p, sz, u8p, i32]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    with open(os.path.join(ROOT, "tests", "golden", "mimc_consts.json")) as fh:
        consts = b"".join(bytes.fromhex(h) for h in json.load(fh))
    L.bpo_set_mimc_constants(consts)
    _LIB = L
    return L


NAIVE, STRAUS_CT, VARTIME = 0, 1, 2


def _buf(n):
    return C.create_string_buffer(n)


def sc_mul(a, b):
    o = _buf(32); lib().bpo_sc_mul(a, b, o); return o.raw


def sc_add(a, b):
    o = _buf(32); lib().bpo_sc_add(a, b, o); return o.raw


def sc_invert(a):
    o = _buf(32); lib().bpo_sc_invert(a, o); return o.raw


def sc_reduce(a):
    o = _buf(32); lib().bpo_sc_reduce(a, o); return o.raw


def sc_wide(a):
    o = _buf(32); lib().bpo_sc_wide(a, o); return o.raw


def point_add(a, b):
    o = _buf(32)
    return o.raw if lib().bpo_point_add(a, b, o) == 0 else None


def point_mul(s, p):
    o = _buf(32)
    return o.raw if lib().bpo_point_mul(s, p, o) == 0 else None


def from_uniform_bytes(b64):
    o = _buf(32); lib().bpo_from_uniform_bytes(b64, o); return o.raw


def pedersen_gens():
    a, b = _buf(32), _buf(32); lib().bpo_pedersen_gens(a, b); return a.raw, b.raw


def gens(i0, n):
    g, h = _buf(32 * n), _buf(32 * n); lib().bpo_gens(i0, n, g, h); return g.raw, h.raw


def pedersen_commit(v, r):
    n = len(v) // 32
    o = _buf(32 * n); lib().bpo_pedersen_commit(v, r, n, o); return o.raw


def msm(scalars, points, algo=VARTIME):
    n = len(scalars) // 32
    o = _buf(32)
    return o.raw if lib().bpo_msm(scalars, points, n, o, algo) == 0 else None


def msm_gens(sG, sH, n, offset=0, extra_scalars=b"", extra_points=b"", algo=VARTIME):
    o = _buf(32)
    k = len(extra_scalars) // 32
    rc = lib().bpo_msm_gens(sG, sH, n, offset, extra_scalars or None, extra_points or None, k, o, algo)
    return o.raw if rc == 0 else None


def fold_points(sl, sr, PL, PR):
    n = len(PL) // 32
    o = _buf(32 * n)
    return o.raw if lib().bpo_fold_points(sl, sr, PL, PR, n, o) == 0 else None


def mimc_hash(data):
    o = _buf(32); lib().bpo_mimc_hash(data, len(data), o); return o.raw


def mimc_sponge(blocks, trace=False):
    nb = len(blocks) // 32
    o = _buf(32)
    tr = _buf(nb * 486 * 192) if trace else None
    lib().bpo_mimc_sponge(blocks, nb, o, tr)
    return (o.raw, tr.raw) if trace else o.raw


class Transcript:
    def __init__(self, label):
        self.h = lib().bpo_transcript_new(label, len(label))

    def append(self, label, msg):
        lib().bpo_transcript_append(self.h, label, len(label), msg, len(msg))

    def challenge(self, label, n):
        o = _buf(n); lib().bpo_transcript_challenge(self.h, label, len(label), o, n); return o.raw

    def __del__(self):
        lib().bpo_transcript_free(self.h)


def _u32(a):
    """ctypes uint32 array from a list or (zero-copy) from a numpy uint32 array"""
    try:
        import numpy as np
        if isinstance(a, np.ndarray):
            a = np.ascontiguousarray(a, dtype=np.uint32)
            _u32.keep = a
            return a.ctypes.data_as(C.POINTER(C.c_uint32))
    except ImportError:
        pass
    return (C.c_uint32 * max(1, len(a)))(*a)


def r1cs_prove(label, gens_capacity, aL, aR, aO, v, v_blinding, row_ptr, term_var, term_coeff, ext_rng32, flags=0):
    """-> (proof bytes, V commitments) ; raises on error."""
    n, m, q = len(aL) // 32, len(v) // 32, len(row_ptr) - 1
    cap = 1 + 32 * (14 + 64 + 2)
    proof, V = _buf(cap), _buf(32 * max(1, m))
    rc = lib().bpo_r1cs_prove(label, len(label), gens_capacity, n, aL, aR, aO, m, v, v_blinding, q, _u32(row_ptr), _u32(term_var),
                              term_coeff, ext_rng32, flags, V, proof, cap)
    if rc < 0:
        raise ValueError("bpo_r1cs_prove rc=%d" % rc)
    return proof.raw[:rc], V.raw[:32 * m]


def r1cs_verify(label, gens_capacity, n, V, row_ptr, term_var, term_coeff, proof, ext_rng32, flags=0):
    m, q = len(V) // 32, len(row_ptr) - 1
    return bool(lib().bpo_r1cs_verify(label, len(label), gens_capacity, n, m, V, q, _u32(row_ptr), _u32(term_var), term_coeff,
                                      proof, len(proof), ext_rng32, flags))
