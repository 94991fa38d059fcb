#!/bin/bash
# commitment MSMs split into per-vector groups: parity at full size, timing
O=gpurun_out
python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_r1cs.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -2
python tools/prove_once.py chain1022 3 fast 2>&1 | head -1
python bench.py --steps 8 --warmup 3 --no-extras --no-cpu > $O/r02_bench_splitcommit.json 2> $O/r02_bench_splitcommit.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_splitcommit.json'):
    if l.startswith('{'):
        d=json.loads(l); print('value %.2f e2e %.2f'%(d['value'], d['e2e']['value']))
PY
