#!/bin/bash
O=gpurun_out
: > $O/r02_acc_minblocks.jsonl
for lib in libbpg.so libbpg_mb5.so libbpg_mb6.so; do
  BPG_LIB=$lib LABEL="$lib" python tools/bench_msm.py 18 20 21 2>/dev/null | tail -1 >> $O/r02_acc_minblocks.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_acc_minblocks.jsonl'):
    d=json.loads(l); print(d['variant'], {k:(round(v['ms'],3), round(v['gadds_per_s'],2)) for k,v in d.items() if k!='variant'})
PY
for lib in libbpg.so libbpg_mb5.so libbpg_mb6.so; do echo "== $lib"; BPG_LIB=$lib python tools/prove_once.py chain1022 3 fast 2>&1 | head -1; done
