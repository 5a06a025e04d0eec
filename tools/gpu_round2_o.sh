#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gram_tc.py -m gpu -q -x 2>&1 | tail -3
for v in "" fin4; do
  if [ -n "$v" ]; then export HYPTOK_B200_LIB=$PWD/hyptokenizer_b200/lib/libhyptok_$v.so; fi
  HYP_TC_TIMING=1 timeout 200 python bench.py --workload c3 --steps 3 --warmup 2 --c3-min-steps 3 --no-tf32-peak > gpurun_out/r2o_c3_$v.json 2> gpurun_out/r2o_c3_$v.err; echo "c3 build=${v:-default} rc=$?"; grep hyp_gram_topk gpurun_out/r2o_c3_$v.err | tail -1
  python - <<PY
import json
l=json.loads(open("gpurun_out/r2o_c3_$v.json").read().strip().splitlines()[-1]); print("   ms", l["ms_per_step"], l["recall"]["bit_identical_to_exact_kernel"])
PY
done
