#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r02_gputests_d.log; cat $O/r02_gputests_d.log
python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_bench_n1_reference_arm.json 2> $O/r02_bench_n1_reference_arm.err
( time python bench.py --steps 20 --warmup 5 ) > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err
tail -4 $O/r02_bench_n1.err
