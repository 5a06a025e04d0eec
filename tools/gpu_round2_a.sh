#!/bin/bash
# round-2 GPU call A: tests, smoke, c3 breakdown, default bench, reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" 
tail -5 gpurun_out/r2a_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2a_smoke.log
for sub in 2 3 4; do
  HYP_TC_SUB=$sub HYP_TC_TIMING=1 timeout 300 python bench.py --workload c3 --steps 3 --warmup 2 --no-tf32-peak > gpurun_out/r2a_c3_sub$sub.json 2> gpurun_out/r2a_c3_sub$sub.err; echo "c3 sub=$sub rc=$?"
  grep hyp_gram_topk gpurun_out/r2a_c3_sub$sub.err | tail -2
done
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    l=json.loads(open("gpurun_out/r2a_bench.json").read().strip().splitlines()[-1])
    print({k:l[k] for k in ("value","ms_per_step","clocks")}, l["roofline"]["frac"], l.get("e2e"))
    print({k:v for k,v in l["config"].items() if k.startswith("c3_") or k.startswith("c4_")})
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2a_ref.json
