#!/bin/bash
# A/B: padded (128 B) table entries x entry prefetch (none / L2 / L1) in k_msm_accumulate
O=gpurun_out
: > $O/r02_acc_pf_pad.jsonl
for lib in libbpg.so libbpg_pad.so; do for pf in 0 1 2; do
  BPG_LIB=$lib BPG_ACC_PF=$pf LABEL="$lib pf=$pf" python tools/bench_msm.py 18 20 21 2>/dev/null | tail -1 >> $O/r02_acc_pf_pad.jsonl
done; done
for lib in libbpg.so libbpg_pad.so; do for pf in 0 2; do
  echo "== $lib pf=$pf" >> $O/r02_acc_pf_pad_proof.log
  BPG_LIB=$lib BPG_ACC_PF=$pf python tools/prove_once.py chain1022 3 fast >> $O/r02_acc_pf_pad_proof.log 2>&1
done; done
BPG_LIB=libbpg_pad.so BPG_ACC_PF=2 python -m pytest tests/test_gpu_r1cs.py tests/test_gpu_core.py -m gpu -x -q 2>&1 | tail -3 > $O/r02_sb_tests.log
cat $O/r02_acc_pf_pad.jsonl | cut -c1-400; cat $O/r02_acc_pf_pad_proof.log $O/r02_sb_tests.log
