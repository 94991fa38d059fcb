#!/bin/bash
O=gpurun_out
python bench.py --provers 96 --steps 5 --warmup 2 --no-extras --no-cpu > $O/r02_bench_P96.json 2> $O/r02_bench_P96.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_P96.json'):
    if l.startswith('{'):
        d=json.loads(l); print('P=96 value %.2f e2e %.2f cpu_ms %.0f'%(d['value'], d['e2e']['value'], d['host_cpu_ms_per_proof']), d['run']['setup_s'])
PY
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r02_gputests_final.log; cat $O/r02_gputests_final.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_2p20.csv python tools/prove_once.py chain1022 1 fast > $O/r02_launches_2p20.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -s 6 -c 1 -o $O/r02_acc_full python tools/prove_once.py chain1022 1 fast > $O/r02_acc_full.log 2>&1
python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_bench_n1_reference_arm.json 2> $O/r02_bench_n1_reference_arm.err
python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err
tail -2 $O/r02_bench_n1.err
