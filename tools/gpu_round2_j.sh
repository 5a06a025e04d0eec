#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gram_tc.py tests/test_gpu_topk.py tests/test_gpu_enhanced.py "tests/test_gpu_merge.py::test_device_topk_select_equals_sorted_list" -m gpu -q -x > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2j_tests.log
HYP_TC_TIMING=1 timeout 300 python bench.py --workload c3 --steps 3 --warmup 2 --no-tf32-peak > gpurun_out/r2j_c3.json 2> gpurun_out/r2j_c3.err; echo "c3 rc=$?"; grep hyp_gram_topk gpurun_out/r2j_c3.err | tail -2
timeout 400 python bench.py --workload c5 > gpurun_out/r2j_c5.json 2> gpurun_out/r2j_c5.err; echo "c5 rc=$?"; python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2j_c5.json").read().strip().splitlines()[-1]); c=l["config"]
print(l["value"], c["merges_done"], c["stopped_by"], c["device_s"], c["host_s"], c["candidates_last_refill"], c["device_top"][:3])
PY
timeout 400 python bench.py --workload c5 --c5-v0 10000 --c5-steps 120 > gpurun_out/r2j_c5_10k.json 2> gpurun_out/r2j_c5_10k.err; echo "c5 10k rc=$?"; python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2j_c5_10k.json").read().strip().splitlines()[-1]); c=l["config"]
print(l["value"], c["merges_done"], c["stopped_by"], c["device_s"], c["host_s"], c["candidates_last_refill"], c["device_top"][:3])
PY
