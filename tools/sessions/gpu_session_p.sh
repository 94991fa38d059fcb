#!/bin/bash
# other BASELINE configs with the CPU oracle beside them; full ncu capture of the two scatter kernels of a 2^20 proof
O=gpurun_out
python tools/bench_configs.py 0 3 > $O/r02_configs_0_3.json 2> $O/r02_configs_0_3.err; tail -3 $O/r02_configs_0_3.err; cat $O/r02_configs_0_3.json | head -60
ncu --set full --clock-control none -k regex:'k_msm_scatter_smem|k_msm_digits' -c 4 -o $O/r02_sort_full python tools/prove_once.py chain1022 1 fast > $O/r02_sort_full.log 2>&1
tail -2 $O/r02_sort_full.log
