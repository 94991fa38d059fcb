#!/bin/bash
O=gpurun_out
for m in 0 1; do echo "== BPG_ACC_PRIO=$m" >> $O/r02_accprio_proof.log; BPG_ACC_PRIO=$m python tools/prove_once.py chain1022 3 fast >> $O/r02_accprio_proof.log 2>&1; BPG_ACC_PRIO=$m python tools/prove_once.py merkle32 5 fast >> $O/r02_accprio_proof.log 2>&1; done
cat $O/r02_accprio_proof.log
for m in 1 0; do BPG_ACC_PRIO=$m python bench.py --steps 8 --warmup 3 --no-extras --no-cpu > $O/r02_bench_accprio$m.json 2> $O/r02_bench_accprio$m.err; done
python - <<'PY'
import json
for m in (0,1):
    for l in open('gpurun_out/r02_bench_accprio%d.json'%m):
        if l.startswith('{'):
            d=json.loads(l); print('BPG_ACC_PRIO=%d'%m, 'value %.2f e2e %.2f cpu_ms %.0f'%(d['value'], d['e2e']['value'], d['host_cpu_ms_per_proof']))
PY
BPG_ACC_PRIO=1 python -m pytest tests/test_gpu_r1cs.py tests/test_gpu_batch_verify.py -m gpu -x -q 2>&1 | tail -2
