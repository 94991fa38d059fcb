#!/bin/bash
O=gpurun_out
python -m pytest tests/test_gpu_r1cs.py tests/test_gpu_fullsize.py tests/test_gpu_core.py -m gpu -x -q 2>&1 | tail -2
for m in 0 1; do echo "== BPG_SCATTER_PG=$m"; BPG_SCATTER_PG=$m python tools/prove_once.py chain1022 3 fast 2>&1 | head -1; done
for m in 0 1; do BPG_SCATTER_PG=$m python bench.py --steps 6 --warmup 3 --no-extras --no-cpu > $O/r02_bench_spg$m.json 2> $O/r02_bench_spg$m.err; done
python - <<'PY'
import json
for m in (0,1):
    for l in open('gpurun_out/r02_bench_spg%d.json'%m):
        if l.startswith('{'):
            d=json.loads(l); print('BPG_SCATTER_PG=%d'%m, 'value %.2f e2e %.2f'%(d['value'], d['e2e']['value']))
PY
