#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gram_tc.py tests/test_gpu_topk.py tests/test_gpu_merge.py -m gpu -q -x > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2i_tests.log
HYP_TC_TIMING=1 timeout 300 python bench.py --workload c3 --steps 3 --warmup 2 --no-tf32-peak > gpurun_out/r2i_c3.json 2> gpurun_out/r2i_c3.err; echo "c3 rc=$?"; grep hyp_gram_topk gpurun_out/r2i_c3.err | tail -2
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2i_c3.json").read().strip().splitlines()[-1]); print(l["ms_per_step"], l["value"], l["recall"]["bit_identical_to_exact_kernel"])
PY
timeout 400 python bench.py --workload c5 > gpurun_out/r2i_c5.json 2> gpurun_out/r2i_c5.err; echo "c5 rc=$?"; cut -c1-1800 gpurun_out/r2i_c5.json; tail -3 gpurun_out/r2i_c5.err
