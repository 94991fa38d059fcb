#!/usr/bin/env python3
"""BASELINE config 5 shape: many small independent proofs (own transcript each, ~1 % deliberately invalid) verified by
K concurrent contexts on one GPU; verdicts are checked against the expected ones.  Proof generation (setup) uses the GPU
prover through the same C ABI."""
import ctypes as C
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import bulletproofs_gadgets_b200 as bpg
    import circuits
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    total = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    ctxs = [bpg.Context(0) for _ in range(K)]
    for c in ctxs:
        c.gens_ensure(512)
    shapes = {}
    items = []
    for k in range(total):
        nm = 100 + (k % 5) * 90          # 100 .. 460 multipliers (the set_membership / less_than range of SURVEY 8a config 5)
        wrong = (k % 101 == 7)
        inst = circuits.chain_instance(nm, 7000 + (k % 40), wrong=wrong) if (nm, wrong, k % 40) not in shapes else shapes[(nm, wrong, k % 40)]
        shapes[(nm, wrong, k % 40)] = inst
        items.append((inst, not wrong))
    # one circuit handle per (context, shape); proofs made once per shape on context 0
    proofs = {}
    handles = [dict() for _ in range(K)]
    for key, inst in shapes.items():
        rp, tv, tc = inst["csr"]
        for ci, c in enumerate(ctxs):
            h = C.c_void_p()
            c.check(c.lib.bpg_circuit_create(c.h, inst["n"], 3, len(rp) - 1, (C.c_uint32 * len(rp))(*rp), (C.c_uint32 * max(1, len(tv)))(*tv), tc, C.byref(h)))
            handles[ci][key] = h
        cap = 1 + 32 * 80
        proof, V = C.create_string_buffer(cap), C.create_string_buffer(96)
        c0 = ctxs[0]
        rc = c0.lib.bpg_r1cs_prove(c0.h, handles[0][key], inst["label"], len(inst["label"]), inst["aL"], inst["aR"], inst["aO"], inst["vals"], inst["blinds"],
                                   bytes(32), 0, V, proof, cap)
        assert rc > 0
        proofs[key] = (proof.raw[:rc], V.raw)
    keys = [(100 + (k % 5) * 90, (k % 101 == 7), k % 40) for k in range(total)]

    def lane(ci):
        c = ctxs[ci]
        out = []
        for k in range(ci, total, K):
            key = keys[k]
            inst = shapes[key]
            proof, V = proofs[key]
            acc = C.c_int(-1)
            c.check(c.lib.bpg_r1cs_verify(c.h, handles[ci][key], inst["label"], len(inst["label"]), V, proof, len(proof), bytes(32), 0, C.byref(acc)))
            out.append((k, bool(acc.value)))
        return out

    def lane_batched(ci, B=64):
        c = ctxs[ci]
        out = []
        mine = list(range(ci, total, K))
        for b0 in range(0, len(mine), B):
            ks = mine[b0:b0 + B]
            batch = [(handles[ci][keys[k]], shapes[keys[k]]["label"], proofs[keys[k]][1], proofs[keys[k]][0], bytes([k % 251]) * 32) for k in ks]
            out += list(zip(ks, c.verify_batch(batch)))
        return out

    pool = ThreadPoolExecutor(max_workers=K)
    list(pool.map(lane, range(K)))  # warm-up
    t0 = time.perf_counter()
    res = list(pool.map(lane, range(K)))
    dt = time.perf_counter() - t0
    verdicts = dict(x for r in res for x in r)
    ok = all(verdicts[k] == items[k][1] for k in range(total))
    list(pool.map(lane_batched, range(K)))
    t0 = time.perf_counter()
    resb = list(pool.map(lane_batched, range(K)))
    dtb = time.perf_counter() - t0
    vb = dict(x for r in resb for x in r)
    okb = all(vb[k] == items[k][1] for k in range(total))
    print(json.dumps({"contexts": K, "proofs": total, "verifications_per_sec": total / dt, "batched_verifications_per_sec": total / dtb,
                      "batched_verdicts_match_expected": okb, "batch_size": 64, "invalid": sum(1 for it in items if not it[1]),
                      "verdicts_match_expected": ok}))


if __name__ == "__main__":
    main()
