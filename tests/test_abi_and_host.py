"""CPU tests: libbpg.so loads and exports every symbol include/bpg.h declares (no compute without a GPU), fails loudly
when asked to compute without a device, and the host-side logic (Merlin transcript, gadget wiring, CSR builders)
matches the oracle."""
import os
import re

import pytest

import circuits
import oracle_lib as ol
from oracle import pyref as pr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = pr.L


def test_library_exports_every_declared_symbol():
    import bulletproofs_gadgets_b200 as bpg
    lib = bpg.load()
    hdr = open(os.path.join(ROOT, "include", "bpg.h")).read()
    declared = set(re.findall(r"\b(bpg_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(bpg._lib.SYMBOLS), declared ^ set(bpg._lib.SYMBOLS)


def _c_prototypes():
    hdr = open(os.path.join(ROOT, "include", "bpg.h")).read()
    h = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = re.findall(r"\n\s*[A-Za-z_][A-Za-z0-9_ \*]*?\b(bpg_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", h)
    out = []
    for name, args in protos:
        args = " ".join(args.split())
        out.append((name, 0 if args in ("", "void") else args.count(",") + 1))
    return out


def test_rust_ffi_matches_header():
    """shim/src/ffi.rs (the reference-side binding) declares exactly the prototypes of include/bpg.h: same names,
    same order, same arity, pointer-ness of every parameter, and the same constants.  No Rust toolchain exists in this image, so
    the check is textual; shim/build.rs must build the same sources the Makefile builds."""
    ffi = open(os.path.join(ROOT, "shim", "src", "ffi.rs")).read()
    rust = re.findall(r"pub fn (bpg_[a-z0-9_]+)\((.*?)\)(?: -> [^;]+)?;", ffi)
    c = _c_prototypes()
    assert [n for n, _ in rust] == [n for n, _ in c]
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "bpg.h")).read(), flags=re.S)
    for (name, args), (_, arity) in zip(rust, c):
        params = [a for a in args.split(", ") if a]
        assert len(params) == arity, name
        cargs = re.search(name + r"\s*\(([^;]*?)\)\s*;", hdr).group(1)
        cparams = [a.strip() for a in " ".join(cargs.split()).split(",")] if arity else []
        for rp, cp in zip(params, cparams):
            is_ptr_c = "*" in cp or "[" in cp or "bpg_allgather_fn" in cp
            is_ptr_r = "*" in rp or "Option<" in rp
            assert is_ptr_c == is_ptr_r, (name, rp, cp)
    for cname, val in re.findall(r"#define (BPG_[A-Z_]+) \(?(-?\d+)u?\)?", open(os.path.join(ROOT, "include", "bpg.h")).read()):
        m = re.search(r"pub const %s: [iu]32 = (-?\d+);" % cname, ffi)
        assert m and int(m.group(1)) == int(val), cname
    build_rs = open(os.path.join(ROOT, "shim", "build.rs")).read()
    mk = open(os.path.join(ROOT, "bulletproofs_gadgets_b200", "csrc", "Makefile")).read()
    for token in ("arch=compute_100a,code=sm_100a", "bpg.cu", "host_keccak_lanes.cpp", "-ldl"):
        assert token in build_rs and token in mk, token
    lib_rs = open(os.path.join(ROOT, "shim", "src", "lib.rs")).read()
    for used in set(re.findall(r"ffi::(bpg_[a-z0-9_]+)", lib_rs)):
        assert used in [n for n, _ in c], used


def test_no_cpu_fallback_without_device():
    import torch

    import bulletproofs_gadgets_b200 as bpg
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(bpg.BpgError) as e:
        bpg.Context(0)
    assert e.value.code == -1  # BPG_E_CUDA


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "bulletproofs_gadgets_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".inl", ".cpp")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "oracle_lib" not in src and "libbpo" not in src and "pyref" not in src and "bpo.h" not in src, fn


def test_host_merlin_transcript_matches_vector_and_oracle():
    import random

    import bulletproofs_gadgets_b200 as bpg
    t = bpg.Transcript(b"test protocol")
    t.append_message(b"some label", b"some data")
    assert t.challenge_bytes(b"challenge", 32).hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"
    rnd = random.Random(3)
    a, b = bpg.Transcript(b"x"), ol.Transcript(b"x")
    for n in (0, 1, 31, 165, 166, 167, 500):
        m = rnd.randbytes(n)
        a.append_message(b"lab", m)
        b.append(b"lab", m)
        assert a.challenge_bytes(b"c", 300) == b.challenge(b"c", 300)


def test_gadget_wiring_fast_builder_equals_reference_shaped_wiring():
    """merkle_path_instance (numpy CSR) must produce exactly the rows the generic MimcHash256 + MerkleTree256 wiring
    records through the ConstraintSystem mirror, and a witness the C oracle proves and verifies."""
    import numpy as np

    import bulletproofs_gadgets_b200 as bpg
    from bulletproofs_gadgets_b200 import gadgets
    depth = 2
    inst = gadgets.merkle_path_instance(depth, trace_on_device=False)
    vals = [int.from_bytes(inst["vals"][32 * i:32 * i + 32], "little") for i in range(4)]
    blinds = [int.from_bytes(inst["blinds"][32 * i:32 * i + 32], "little") for i in range(4)]
    prover = bpg.Prover.new(inst["label"])
    Vv = [prover.commit(v, b)[1] for v, b in zip(vals, blinds)]
    gadgets.mimc_hash_gadget_wire(prover, [Vv[0]], [Vv[2], Vv[3]], [(Vv[1], 1)])
    # instance leaves: the sibling constants of the fast builder
    rp, tv, tc = inst["csr"]
    tcb = tc.tobytes()
    sib = []
    for k in range(depth):
        row = 1 + 1944 + 1 + 3888 * k + 1944  # first row of the second block of level k
        terms = range(rp[row], rp[row + 1])
        one_terms = [int.from_bytes(tcb[32 * t:32 * t + 32], "little") for t in terms if tv[t] >> 29 == 4]
        sib.append(one_terms[0])
    pattern = "W"
    for _ in range(depth):
        pattern = (pattern, "I")
    pattern = (("W", "I"), "I")
    gadgets.merkle_wire(prover, [(bpg.api.ONE, inst["root"])], pattern, [[(Vv[1], 1)]], [[(bpg.api.ONE, s)] for s in sib])
    rp2, tv2, tc2 = prover.csr()
    assert prover.get_num_multiplications() == inst["n"] == 972 + 1944 * depth
    assert list(rp) == rp2 and list(tv) == tv2 and tcb == tc2
    enc = lambda xs: b"".join(int(x).to_bytes(32, "little") for x in xs)
    from circuits import host_witness
    aL, aR, aO = host_witness(prover)
    assert enc(aL) == inst["aL"] and enc(aR) == inst["aR"] and enc(aO) == inst["aO"]
    proof, V = ol.r1cs_prove(inst["label"], 8192, inst["aL"], inst["aR"], inst["aO"], inst["vals"], inst["blinds"], rp, tv, tcb, bytes(32))
    assert ol.r1cs_verify(inst["label"], 8192, inst["n"], V, rp, tv, tcb, proof, bytes(32))
    # the image really is the reference's mimc_hash of the leaf (CLI semantics, prover.rs:171)
    assert vals[1] == int.from_bytes(ol.mimc_hash(b"\x43"), "little")


def test_mimc_preprocess_matches_reference_padding_rules():
    from bulletproofs_gadgets_b200 import gadgets
    for data in (b"\x43", b"John", bytes(range(1, 32)), bytes(range(1, 33)), b"\x00\x01", b"\xff" * 40):
        sc = gadgets.be_to_scalars(data)
        d = gadgets.mimc_preprocess(sc)
        blocks = sc[:-1] + [d[0]] if len(d) == 2 else sc + [d[0]]
        assert gadgets.mimc_sponge_int(blocks) == int.from_bytes(ol.mimc_hash(data), "little"), data
        if len(d) == 2:
            assert (sc[-1] + d[1]) % L == d[0] % L
