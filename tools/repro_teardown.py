import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import faulthandler; faulthandler.enable()
import numpy as np
import bulletproofs_gadgets_b200 as bpg
mode = sys.argv[1] if len(sys.argv) > 1 else "all"
ctx = bpg.Context(0)
ctx.gens_ensure(1 << 16)
n = 1 << 16
raw = np.random.default_rng(1).integers(0, 256, size=(1 << 22, 32), dtype=np.uint8); raw[:, 31] &= 0x0F
d = ctx.dev_alloc(32 * (1 << 22)); ctx.dev_upload(d, raw.tobytes())
if mode in ("all", "prof"):
    ctx.prof_enable(True)
    for _ in range(5):
        ctx.msm_gens_dev(d, C.c_void_p(d.value + 32 * n), n, 0)
    print("prof", ctx.prof_read()); ctx.prof_enable(False)
if mode in ("all", "big"):
    ctx.gens_ensure(1 << 21)
    h = 1 << 21
    print(ctx.msm_gens_dev(d, C.c_void_p(d.value + 32 * h), h, 0).hex()[:16])
ctx.dev_free(d)
print("closing", flush=True)
ctx.close()
print("closed ok", flush=True)
