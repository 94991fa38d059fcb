#!/bin/bash
O=gpurun_out
python -m pytest tests/test_gpu_r1cs.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3 > $O/r02_sf_tests.log; cat $O/r02_sf_tests.log
: > $O/r02_sort2_ab2.jsonl
for m in 0 1 2; do
  BPG_SORT2=$m LABEL="BPG_SORT2=$m uniform" python tools/bench_msm.py 19 20 21 2>/dev/null | tail -1 >> $O/r02_sort2_ab2.jsonl
  BPG_SORT2=$m DIST=bits LABEL="BPG_SORT2=$m bits" python tools/bench_msm.py 20 21 2>/dev/null | tail -1 >> $O/r02_sort2_ab2.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_sort2_ab2.jsonl'):
    d=json.loads(l); print(d['variant'], {k:round(v['ms'],3) for k,v in d.items() if k!='variant'})
PY
for mb in 0 1; do echo "== BPG_MAT_BLOCK=$mb" >> $O/r02_matblock_proof.log; BPG_MAT_BLOCK=$mb python tools/prove_once.py chain1022 3 fast >> $O/r02_matblock_proof.log 2>&1; done
cat $O/r02_matblock_proof.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_2p20_sort2.csv python tools/prove_once.py chain1022 1 fast > $O/r02_launches_2p20_sort2.log 2>&1
python tools/summarise_launches.py $O/r02_launches_2p20_sort2.csv | head -30
