#!/bin/bash
# round-2 GPU call C: pair-count v3 parity + A/B against v2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_paircount.py tests/test_gpu_freq.py -m gpu -x -q > gpurun_out/r2c_pc_tests.log 2>&1; echo "pc tests rc=$?"; tail -15 gpurun_out/r2c_pc_tests.log
for v in v2 v3; do HYP_PAIR_COUNT=$v timeout 300 python tools/bench_pair.py 1024 2>&1 | tail -2; done | tee gpurun_out/r2c_pair_ab.txt
