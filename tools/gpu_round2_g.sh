#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"; tail -22 gpurun_out/r2g_tests.log
for v in v3 auto; do
  if [ $v = auto ]; then unset HYP_PAIR_COUNT; else export HYP_PAIR_COUNT=$v; fi
  timeout 300 python tools/bench_pair.py 1024 2>&1 | tail -2
done | tee gpurun_out/r2g_pair_ab.txt
unset HYP_PAIR_COUNT
timeout 400 python bench.py --workload c5 --c5-steps 150 > gpurun_out/r2g_c5.json 2> gpurun_out/r2g_c5.err; echo "c5 rc=$?"; cut -c1-1500 gpurun_out/r2g_c5.json; tail -3 gpurun_out/r2g_c5.err
