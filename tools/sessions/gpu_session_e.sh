#!/bin/bash
# two-pass bin sort: parity (whole GPU suite with the new path forced for every privatised MSM), then A/B timings
O=gpurun_out
BPG_SORT2=2 python -m pytest tests/test_gpu_core.py tests/test_gpu_r1cs.py tests/test_gpu_fullsize.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -5 > $O/r02_se_tests.log
cat $O/r02_se_tests.log
: > $O/r02_sort2_ab.jsonl
for m in 0 1 2; do
  BPG_SORT2=$m LABEL="BPG_SORT2=$m uniform" python tools/bench_msm.py 19 20 21 22 2>/dev/null | tail -1 >> $O/r02_sort2_ab.jsonl
  BPG_SORT2=$m DIST=bits LABEL="BPG_SORT2=$m bits" python tools/bench_msm.py 20 22 2>/dev/null | tail -1 >> $O/r02_sort2_ab.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_sort2_ab.jsonl'):
    d=json.loads(l); print(d['variant'], {k:round(v['ms'],3) for k,v in d.items() if k!='variant'})
PY
for m in 0 1; do echo "== BPG_SORT2=$m" >> $O/r02_sort2_proof.log; BPG_SORT2=$m python tools/prove_once.py chain1022 3 fast >> $O/r02_sort2_proof.log 2>&1; done
cat $O/r02_sort2_proof.log
for m in 0 1; do BPG_SORT2=$m python bench.py --steps 3 --warmup 3 --no-extras --no-cpu > $O/r02_bench_sort2_$m.json 2> $O/r02_bench_sort2_$m.err; done
python - <<'PY'
import json
for m in (0,1):
    for l in open('gpurun_out/r02_bench_sort2_%d.json'%m):
        if l.startswith('{'):
            d=json.loads(l); print('BPG_SORT2=%d'%m, 'value %.2f e2e %.2f cpu_ms %.0f'%(d['value'], d['e2e']['value'], d['host_cpu_ms_per_proof']), d['run']['setup_s'])
PY
