"""MSM sweep + accumulate-kernel timing for one library build (development aid; LABEL=... tags the output line).
usage: python tools/bench_msm.py [log2 sizes...]  -> one JSON line"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
import bulletproofs_gadgets_b200 as bpg  # noqa: E402

sizes = [1 << int(a) for a in sys.argv[1:]] or [1 << 17, 1 << 20]
ctx = bpg.Context(0)
ctx.gens_ensure(max(sizes) // 2)
out = {"variant": os.environ.get("LABEL", os.environ.get("BPG_ACC_VARIANT", "0"))}
for n in sizes:
    ctx.prof_enable(True)
    r = bench.msm_sweep(ctx, [n], reps=8, dist=os.environ.get("DIST", "uniform"))
    nl, kms, pairs = ctx.prof_read()
    ctx.prof_enable(False)
    r[str(n)]["accumulate_ms"] = kms / max(nl, 1)
    r[str(n)]["gadds_per_s"] = pairs / max(kms, 1e-9) / 1e6
    out.update(r)
print(json.dumps(out))
ctx.close()
