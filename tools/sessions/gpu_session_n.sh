#!/bin/bash
# final state of round 2: smoke, full GPU test suite, launch list, both bench arms
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; tail -1 $O/r02_smoke.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r02_gputests_final.log; cat $O/r02_gputests_final.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_2p20.csv python tools/prove_once.py chain1022 1 fast > $O/r02_launches_2p20.log 2>&1
python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_bench_n1_reference_arm.json 2> $O/r02_bench_n1_reference_arm.err
python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err
tail -2 $O/r02_bench_n1.err
