#!/bin/bash
# block size of the privatised histogram / scatter kernels: solo and under load
O=gpurun_out
for t in 1024 512 256; do echo "== BPG_PRIV_THREADS=$t"; BPG_PRIV_THREADS=$t python tools/prove_once.py chain1022 3 fast 2>&1 | head -1; done
for t in 512 1024; do BPG_PRIV_THREADS=$t python bench.py --steps 8 --warmup 3 --no-extras --no-cpu > $O/r02_bench_pth$t.json 2> $O/r02_bench_pth$t.err; done
python - <<'PY'
import json
for t in (1024,512):
    for l in open('gpurun_out/r02_bench_pth%d.json'%t):
        if l.startswith('{'):
            d=json.loads(l); print('BPG_PRIV_THREADS=%d'%t, 'value %.2f e2e %.2f'%(d['value'], d['e2e']['value']))
PY
