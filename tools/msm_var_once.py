#!/usr/bin/env python3
"""bpg_msm (variable-base MSM, host buffers) at one size, a few repetitions, host-timed -- for launch lists under ncu.
usage: python tools/msm_var_once.py [log2 n] [reps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bulletproofs_gadgets_b200 as bpg  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = 1 << lg
ctx = bpg.Context(0)
ctx.gens_ensure(max(64, n // 2))
G, H = ctx.gens_export(0, n // 2)
pts = G + H
raw = np.random.default_rng(3).integers(0, 256, size=(n, 32), dtype=np.uint8)
raw[:, 31] &= 0x0F
sc = raw.tobytes()
ctx.msm(sc, pts)
for _ in range(reps):
    t0 = time.perf_counter()
    ctx.msm(sc, pts)
    print("n=2^%d bpg_msm %.3f ms" % (lg, 1e3 * (time.perf_counter() - t0)), flush=True)
ctx.close()
