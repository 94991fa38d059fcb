#!/bin/bash
O=gpurun_out
python -m pytest tests/test_gpu_r1cs.py tests/test_gpu_fullsize.py tests/test_frontend.py -m gpu -x -q 2>&1 | tail -3 > $O/r02_sg_tests.log; cat $O/r02_sg_tests.log
for lg in 12 13 14; do echo "== BPG_LATE_FOLD_LG=$lg" >> $O/r02_latefold_2p20.log; BPG_LATE_FOLD_LG=$lg python tools/prove_once.py chain1022 3 fast >> $O/r02_latefold_2p20.log 2>&1; done
cat $O/r02_latefold_2p20.log
for lg in 12 13 14; do BPG_LATE_FOLD_LG=$lg python bench.py --steps 3 --warmup 3 --no-extras --no-cpu --provers 24 > $O/r02_bench_lf$lg.json 2> $O/r02_bench_lf$lg.err; done
python - <<'PY'
import json
for lg in (12,13,14):
    for l in open('gpurun_out/r02_bench_lf%d.json'%lg):
        if l.startswith('{'):
            d=json.loads(l); print('LATE_FOLD_LG=%d'%lg, 'value %.2f e2e %.2f cpu_ms %.0f'%(d['value'], d['e2e']['value'], d['host_cpu_ms_per_proof']))
PY
