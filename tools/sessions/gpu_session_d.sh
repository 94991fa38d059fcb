#!/bin/bash
O=gpurun_out
for P in 24 32 48; do
  python bench.py --steps 4 --warmup 3 --no-extras --no-cpu --provers $P > $O/r02_bench_P${P}c.json 2> $O/r02_bench_P${P}c.err
done
BPG_SIZING_MODE=0 python bench.py --steps 4 --warmup 3 --no-extras --no-cpu --provers 24 > $O/r02_bench_P24_latsizing.json 2> $O/r02_bench_P24_latsizing.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_msmvar_2p20.csv python tools/msm_var_once.py 20 2 > $O/r02_msmvar_2p20.log 2>&1
python tools/msm_var_once.py 16 3 >> $O/r02_msmvar_2p20.log 2>&1
python tools/msm_var_once.py 18 3 >> $O/r02_msmvar_2p20.log 2>&1
python tools/msm_var_once.py 22 3 >> $O/r02_msmvar_2p20.log 2>&1
grep -h "bpg_msm" $O/r02_msmvar_2p20.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_P*c.json'))+['gpurun_out/r02_bench_P24_latsizing.json']:
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, 'value %.2f e2e %.2f cpu_ms %.0f'%(d['value'], d['e2e']['value'], d['host_cpu_ms_per_proof']))
PY
