#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_paircount.py -m gpu -x -q > gpurun_out/r2e_pc_tests.log 2>&1; rc=$?; echo "pc tests rc=$rc"; tail -4 gpurun_out/r2e_pc_tests.log
for v in "" g1 racy; do
  if [ -n "$v" ]; then export HYPTOK_B200_LIB=$PWD/hyptokenizer_b200/lib/libhyptok_$v.so; fi
  echo "== build ${v:-default}"; HYP_PAIR_COUNT=v3 timeout 300 python tools/bench_pair.py 1024 2>&1 | tail -2
done | tee gpurun_out/r2e_pair_ab.txt
