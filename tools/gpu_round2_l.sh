#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gram_tc.py -m gpu -q -x > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2l_tests.log
for cfg in "ss 2" "ts 4" "ts 2" "ts 3" "ts 6"; do
  set -- $cfg
  HYP_TC_ENGINE=$1 HYP_TC_SUB=$2 HYP_TC_TIMING=1 timeout 200 python bench.py --workload c3 --steps 3 --warmup 2 --c3-min-steps 3 --no-tf32-peak > gpurun_out/r2l_c3_$1_$2.json 2> gpurun_out/r2l_c3_$1_$2.err; echo "c3 $1 sub=$2 rc=$?"; grep hyp_gram_topk gpurun_out/r2l_c3_$1_$2.err | tail -1
  python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/r2l_c3_$1_$2.json").read().strip().splitlines()[-1]); print("   ms", l["ms_per_step"], "bit-identical", l["recall"]["bit_identical_to_exact_kernel"], "redone", l["config"]["rows_redone_exactly"])
except Exception as e: print("   parse failed", e)
PY
done
