#!/bin/bash
O=gpurun_out
python -m pytest tests/test_gpu_core.py -m gpu -x -q 2>&1 | tail -3
python - <<'PY' 2>&1 | tail -5
import sys, json
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import bench, bulletproofs_gadgets_b200 as bpg
ctx = bpg.Context(0); ctx.gens_ensure(1 << 18)
ms, mac = ctx.bench_imad(400)
print(json.dumps(bench.fold_sweep(ctx, [1 << 12, 1 << 16, 1 << 18], mac / ms * 1e3)))
PY
