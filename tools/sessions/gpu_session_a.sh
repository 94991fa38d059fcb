#!/bin/bash
# one-GPU measurement session (round 2): parity after the k_mat_reduce change, pipe microbenchmarks, short bench, ncu launch list + full capture
set -x
O=gpurun_out
python -m pytest tests/test_gpu_r1cs.py tests/test_gpu_configs.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -5 > $O/r02_sa_tests.log
tools/pipe_bench > $O/r02_pipe_bench_mix.jsonl 2>&1
tools/dfma_bench > $O/r02_dfma_bench_ws.jsonl 2>&1
ncu --set full --clock-control none -k regex:k_bench -c 5 -o $O/r02_dfma_ncu tools/dfma_bench ncu > $O/r02_dfma_ncu.log 2>&1
python bench.py --steps 4 --warmup 3 --no-extras --no-cpu > $O/r02_bench_P24b.json 2> $O/r02_bench_P24b.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_2p20.csv python tools/prove_once.py chain1022 1 fast > $O/r02_launches_2p20.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -s 6 -c 1 -o $O/r02_acc_full python tools/prove_once.py chain1022 1 fast > $O/r02_acc_full.log 2>&1
ls -la $O | tail -12
