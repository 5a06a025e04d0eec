#!/bin/bash
# round-2 GPU call K: validation of the templated select + the round's ncu captures (one ncu use per call: all in one chain)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gram_tc.py tests/test_gpu_topk.py -m gpu -q -x > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2k_tests.log
HYP_TC_TIMING=1 timeout 300 python bench.py --workload c3 --steps 3 --warmup 2 --c3-min-steps 3 --no-tf32-peak > gpurun_out/r2k_c3.json 2> gpurun_out/r2k_c3.err; echo "c3 rc=$?"; grep hyp_gram_topk gpurun_out/r2k_c3.err | tail -2
C2="python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
C3="python bench.py --workload c3 --steps 2 --warmup 1 --c3-min-steps 2 --no-tf32-peak"
C4="python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline"
$C2 > gpurun_out/r2k_plain_c2.log 2>&1 && $C3 > gpurun_out/r2k_plain_c3.log 2>&1 && $C4 > gpurun_out/r2k_plain_c4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_c2_launches.csv $C2 > gpurun_out/r2k_ncu1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_c3_launches.csv $C3 > gpurun_out/r2k_ncu2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_c4_launches.csv $C4 > gpurun_out/r2k_ncu3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gram_tc_kernel -s 2 -c 2 -o gpurun_out/r02_gram $C3 > gpurun_out/r2k_ncu4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:merge_loop_resident -s 1 -c 1 -o gpurun_out/r02_merge $C2 > gpurun_out/r2k_ncu5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc_finish|kth_select" -s 2 -c 2 -o gpurun_out/r02_finish $C3 > gpurun_out/r2k_ncu6.log 2>&1
echo "ncu chain rc=$?"
