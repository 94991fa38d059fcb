#!/bin/bash
# two-GPU session: multi-GPU parity tests, sharded-proof trace, bench.py --gpus 2 with every extra
O=gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5 > $O/r02_gpu_multi_2gpu_pytest.log
BPG_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/prove_sharded.py 1022 3 > $O/r02_sharded_trace_n2.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 4 --warmup 3 > $O/r02_bench_n2.json 2> $O/r02_bench_n2.err
echo rc=$? >> $O/r02_bench_n2.err
cat $O/r02_gpu_multi_2gpu_pytest.log; grep -v "^\[bpg" $O/r02_sharded_trace_n2.log | tail -5; grep "bpg prove" $O/r02_sharded_trace_n2.log | tail -2; tail -3 $O/r02_bench_n2.err; cut -c1-600 $O/r02_bench_n2.json
