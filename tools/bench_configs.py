#!/usr/bin/env python3
"""Times the BASELINE.json configs that are not the bench.py headline (run on a B200):
  config 3: 4096 x 64-bit BoundsCheck in one proof (n = 2^19 multipliers, m = 12 288, bit-valued a_L/a_R)
  config 4: 2^20-multiplier class circuit (1022 absorbed MiMC blocks = 993 384 multipliers, N = 2^20, 20 IPP rounds)
  config 5: batch verification of many small proofs (one context; rank-sharding is bench.py / parallel.py territory)
Prints one JSON object; each proof is verified by the GPU verifier, and tampered copies must be rejected."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def timed(f, reps=1):
    best = None
    out = None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = f()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best * 1e3, out


def main():
    import bulletproofs_gadgets_b200 as bpg
    from bulletproofs_gadgets_b200 import gadgets
    FAST, RES = bpg._lib.FLAG_FAST_BLINDING, bpg._lib.FLAG_WITNESS_ON_DEVICE
    ctx = bpg.Context(0)
    res = {}
    which = sys.argv[1:] or ["0", "3", "4", "5"]

    def run(name, inst, cap):
        t_g, _ = timed(lambda: ctx.gens_ensure(cap))
        t_c, circ = timed(lambda: gadgets.Circuit(ctx, inst["n"], inst["m"], inst["csr"]))
        ext = b"\x33" * 32
        circ.prove(inst, ext, FAST)  # warm-up (buffers)
        t_exact, (proof, V) = timed(lambda: circ.prove(inst, ext), 2)
        t_fast, (proof_f, _) = timed(lambda: circ.prove(inst, ext, FAST), 2)
        t_ver, ok = timed(lambda: circ.verify(inst["label"], V, proof), 2)
        ok_f = circ.verify(inst["label"], V, proof_f)
        bad = bytearray(proof)
        bad[77] ^= 2
        rej = not circ.verify(inst["label"], V, bytes(bad))
        res[name] = {"n_multipliers": inst["n"], "commitments": inst["m"], "constraints": int(len(inst["csr"][0]) - 1), "gens_capacity": cap,
                     "gens_tables_ms": t_g, "circuit_upload_ms": t_c, "prove_ms_byte_exact": t_exact, "prove_ms_fast_blinding": t_fast,
                     "verify_ms": t_ver, "proof_bytes": len(proof), "verifier_accepts": bool(ok and ok_f), "tamper_rejected": rej}
        circ.close()
        if os.environ.get("BPG_CONFIGS_CPU", "1") == "1":
            # the CPU oracle beside it (all host cores; test infrastructure, timed as the baseline only)
            import bench
            cores = os.cpu_count() or 1
            bench.oracle_prove(inst, cap, ext, cores)  # first call derives the generators
            t_cpu, proof_cpu, V_cpu = bench.oracle_prove(inst, cap, ext, cores)
            res[name].update({"cpu_prove_ms": 1e3 * t_cpu, "cpu_cores": cores, "proof_bytes_equal_oracle": (proof_cpu, V_cpu) == (proof, V)})

    if "0" in which:
        # config 0: the reference's own example.gadgets/.inst/.wtns through the front-end driver (prover + verifier binaries)
        import tempfile, shutil
        from bulletproofs_gadgets_b200 import frontend as fe
        fx = os.path.join(ROOT, "tests", "golden", "fixtures")
        d = tempfile.mkdtemp()
        for ext in (".gadgets", ".inst", ".wtns"):
            shutil.copy(os.path.join(fx, "example" + ext), os.path.join(d, "example" + ext))
        stem = os.path.join(d, "example")
        ctx.gens_ensure(1 << 14)
        fe.prover_main(stem, test_seed=1, ctx=ctx, label="example")
        t_p, nc = timed(lambda: fe.prover_main(stem, test_seed=1, ctx=ctx, label="example"), 2)
        t_v, ok = timed(lambda: fe.verifier_main(stem, ctx=ctx, label="example"), 2)
        prun = fe.ProverRun(b"example", open(stem + ".gadgets").read(), open(stem + ".inst").read(), open(stem + ".wtns").read(), test_seed=1, ctx=ctx)
        t_dev, _ = timed(lambda: prun.prover.prove(bpg.BulletproofGens.new(1 << 14, 1, ctx=ctx), ext_rng32=bytes(32)), 2)
        res["config0_example_cli"] = {"constraints": nc, "multipliers": prun.prover.get_num_multiplications(), "prover_cli_ms_incl_python_frontend": t_p,
                                      "verifier_cli_ms_incl_python_frontend": t_v, "prove_call_ms": t_dev, "verifier_prints": "true" if ok else "false"}
        if os.environ.get("BPG_CONFIGS_CPU", "1") == "1":
            import oracle_lib as ol
            p = prun.prover
            rp, tv, tc = p.csr()
            enc = lambda xs: b"".join(int(x).to_bytes(32, "little") for x in xs)
            aL, aR, aO = p.witness_bytes()
            ol.lib().bpo_set_threads(os.cpu_count() or 1)
            ol.gens(0, 1 << 14)
            t_cpu, (proof_cpu, _) = timed(lambda: ol.r1cs_prove(b"example", 1 << 14, aL, aR, aO, enc(p.v), enc(p.v_blinding), rp, tv, tc, bytes(32)), 2)
            proof_gpu, _ = p.prove(bpg.BulletproofGens.new(1 << 14, 1, ctx=ctx), ext_rng32=bytes(32))
            res["config0_example_cli"].update({"cpu_prove_ms": t_cpu, "cpu_cores": os.cpu_count(), "proof_bytes_equal_oracle": proof_cpu == proof_gpu})
    if "3" in which:
        t0 = time.perf_counter()
        inst = gadgets.bounds_check_batch_instance(4096, 8, seed=5)
        res["config3_build_s"] = time.perf_counter() - t0
        run("config3_4096_bounds_checks", inst, 1 << 19)
    if "4" in which:
        t0 = time.perf_counter()
        inst = gadgets.mimc_chain_instance(1022, ctx=ctx)
        res["config4_build_s"] = time.perf_counter() - t0
        run("config4_2p20_multipliers", inst, 1 << 20)
    if "5" in which:
        import circuits
        import ctypes as C
        import oracle_lib as ol  # only to PRODUCE the small proofs quickly on the CPU; the verifications timed are the GPU's
        items = []
        for k in range(256):
            inst = circuits.chain_instance(8 + (k % 56), 5000 + k, wrong=(k % 97 == 5))
            rp, tv, tc = inst["csr"]
            proof, V = ol.r1cs_prove(inst["label"], 64, inst["aL"], inst["aR"], inst["aO"], inst["vals"], inst["blinds"], rp, tv, tc, bytes([k % 251]) * 32)
            h = C.c_void_p()
            ctx.check(ctx.lib.bpg_circuit_create(ctx.h, inst["n"], 3, len(rp) - 1, (C.c_uint32 * len(rp))(*rp), (C.c_uint32 * max(1, len(tv)))(*tv), tc, C.byref(h)))
            items.append((inst, V, proof, h, k % 97 != 5))
        ctx.gens_ensure(64)

        def verify_all():
            verdicts = []
            for inst, V, proof, h, _ in items:
                acc = C.c_int(-1)
                ctx.check(ctx.lib.bpg_r1cs_verify(ctx.h, h, inst["label"], len(inst["label"]), V, proof, len(proof), bytes(32), 0, C.byref(acc)))
                verdicts.append(bool(acc.value))
            return verdicts
        verify_all()
        t, verdicts = timed(verify_all, 2)
        res["config5_batch_verify"] = {"proofs": len(items), "ms_total": t, "verifications_per_sec_one_context": len(items) / t * 1e3,
                                       "verdicts_match_expected": verdicts == [it[4] for it in items]}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
