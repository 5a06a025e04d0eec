#!/bin/bash
# round-2 GPU call D: pair-count v3 (restructured) parity, A/B, one ncu --set full capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_paircount.py tests/test_gpu_freq.py -m gpu -x -q > gpurun_out/r2d_pc_tests.log 2>&1; rc=$?; echo "pc tests rc=$rc"; tail -4 gpurun_out/r2d_pc_tests.log
[ $rc -ne 0 ] && exit 1
for v in v2 v3; do HYP_PAIR_COUNT=$v timeout 300 python tools/bench_pair.py 1024 2>&1 | tail -2; done | tee gpurun_out/r2d_pair_ab.txt
HYP_PAIR_COUNT=v3 python tools/bench_pair.py 1024 > gpurun_out/r2d_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pair_count_v3 -s 2 -c 1 -o gpurun_out/r02_pc_v3 python tools/bench_pair.py 1024 > gpurun_out/r2d_ncu.log 2>&1; echo "ncu rc=$?"
