#!/usr/bin/env python3
"""One proof of a BASELINE config on one context, for profiling under ncu / BPG_TRACE=1.
usage: python tools/prove_once.py {merkle32|bounds4096|chain1022|chainK:<blocks>} [reps] [fast|exact]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bulletproofs_gadgets_b200 as bpg
    from bulletproofs_gadgets_b200 import gadgets
    which = sys.argv[1] if len(sys.argv) > 1 else "chain1022"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    fast = (sys.argv[3] if len(sys.argv) > 3 else "fast") == "fast"
    ctx = bpg.Context(0)
    if which == "merkle32":
        inst, cap = gadgets.merkle_path_instance(32, ctx=ctx), 1 << 16
    elif which == "bounds4096":
        inst, cap = gadgets.bounds_check_batch_instance(4096, 8, seed=5), 1 << 19
    else:
        blocks = 1022 if which == "chain1022" else int(which.split(":")[1])
        inst = gadgets.mimc_chain_instance(blocks, ctx=ctx)
        cap = 1
        while cap < inst["n"]:
            cap *= 2
    ctx.gens_ensure(cap)
    circ = gadgets.Circuit(ctx, inst["n"], inst["m"], inst["csr"])
    flags = bpg._lib.FLAG_FAST_BLINDING if fast else 0
    ext = b"\x33" * 32
    proof, V = circ.prove(inst, ext, flags)
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    for _ in range(reps):
        proof, V = circ.prove(inst, ext, flags)
    dt = (time.perf_counter() - t0) / reps
    print("n=%d cap=%d prove %.2f ms (%s), %d launches per proof" % (inst["n"], cap, dt * 1e3, "fast blinding" if fast else "byte-exact", (ctx.launch_count() - l0) // reps))
    t0 = time.perf_counter()
    ok = circ.verify(inst["label"], V, proof)
    print("verify %.2f ms -> %s" % ((time.perf_counter() - t0) * 1e3, ok))
    circ.close()
    ctx.close()


if __name__ == "__main__":
    main()
