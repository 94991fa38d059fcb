"""small end-to-end case for compute-sanitizer (closed on this pool in round 2: `gpurun` answers that the tool stays closed; the case still runs plain as a smoke of every path): generators, Pedersen, MSM (uniform + bit-valued), MiMC, prove, verify, witness evaluation, variable-base MSM, page-locked buffers, prefetch"""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bulletproofs_gadgets_b200 as bpg
from bulletproofs_gadgets_b200 import gadgets
import circuits
ctx = bpg.Context(0)
ctx.gens_ensure(256)
rnd = random.Random(1)
rs = lambda: rnd.randrange(2 ** 255).to_bytes(32, "little")
n = 200
sG, sH = b"".join(rs() for _ in range(n)), b"".join(rs() for _ in range(n))
ctx.msm_gens(sG, sH, n, 3)
bits = b"".join(rnd.randrange(2).to_bytes(32, "little") for _ in range(n))
ctx.msm_gens(bits, bits, n, 0)
ctx.pedersen_commit(sG[:320], sH[:320])
G, H = ctx.gens_export(0, 8)
ctx.msm(sG[:256], G)
ctx.fold_points(5, 7, G[:128], G[128:])
ctx.mimc_hash_batch([b"abc", b"\x01" * 32])
inst = gadgets.bounds_check_batch_instance(2, 1, seed=2)
circ = gadgets.Circuit(ctx, inst["n"], inst["m"], inst["csr"])
for flags in (0, 2, 16, 8):  # byte-exact, fast blinding, forced late fold, no late fold
    proof, V = circ.prove(inst, b"\x01" * 32, flags)
    assert circ.verify(inst["label"], V, proof)
circ.close()
# a circuit large enough for the 2^15-bucket MSM path (> 4096 terms) in both kernel sizings, with the late fold forced
ctx.gens_ensure(4096)
inst = circuits.chain_instance(2100, 3)
circ = gadgets.Circuit(ctx, inst["n"], len(inst["vals"]) // 32, inst["csr"])
for mode in (0, 1):
    ctx.lib.bpg_set_sizing_mode(mode)
    proof, V = circ.prove(inst, b"\x02" * 32, 16)
    assert circ.verify(inst["label"], V, proof)
ctx.lib.bpg_set_sizing_mode(-1)
circ.close()
# round 2: device-side witness evaluation (wide level + runs of narrow levels), variable-base bucket MSM (>= 8192 points),
# page-locked host buffers, prefetched opening
L = circuits.L
p = bpg.Prover.new(b"san", ctx=ctx)
_, v0 = p.commit(rnd.randrange(L), 1)
wide = [p.multiply([(v0, k + 1)], [(bpg.api.ONE, k + 2)])[2] for k in range(600)]   # one wide level
cur = [(wide[0], 3), (v0, 1)]
for _ in range(40):                                                               # a narrow chain
    _, _, o = p.multiply(cur, cur + [(bpg.api.ONE, 1)])
    cur = [(o, 1), (wide[5], 2)]
assert p.witness() == circuits.host_witness(p)
proof, Vs = p.prove(bpg.BulletproofGens.new(4096, 1, ctx=ctx), ext_rng32=bytes(32))
ctx.gens_ensure(8192)
G, H = ctx.gens_export(0, 4200)
sc = b"".join(rs() for _ in range(8400))
ctx.msm(sc, G + H)
inst = circuits.chain_instance(300, 9)
circ = gadgets.Circuit(ctx, inst["n"], 3, inst["csr"])
pinned, handles = dict(inst), []
for k in ("aL", "aR", "aO"):
    ptr, hnd = ctx.host_alloc(inst[k])
    pinned[k] = ptr
    handles.append(hnd)
circ.prefetch(inst, b"\x07" * 32)
assert circ.prove(pinned, b"\x07" * 32) == circ.prove(inst, b"\x07" * 32)
for hnd in handles:
    ctx.host_free(hnd)
circ.close()
ctx.close()
print("sanitizer case ok")
