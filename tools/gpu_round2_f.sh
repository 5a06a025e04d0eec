#!/bin/bash
# round-2 GPU call F: full GPU test suite (new: strict-parity proofs, bench-config oracle trace, radix select, wraps), A/B pair count, c2 bench (merge-loop protocol change)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"; tail -18 gpurun_out/r2f_tests.log
for v in v2 v3 auto; do
  if [ $v = auto ]; then unset HYP_PAIR_COUNT; else export HYP_PAIR_COUNT=$v; fi
  timeout 300 python tools/bench_pair.py 1024 2>&1 | tail -2
done | tee gpurun_out/r2f_pair_ab.txt
unset HYP_PAIR_COUNT
timeout 600 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/r2f_c2.json 2> gpurun_out/r2f_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2f_c2.json").read().strip().splitlines()[-1])
print("c2", l["value"], l["ms_per_step"], l["roofline"]["frac"], l["roofline"]["smem_frac"], l["e2e"]["value"], l["clocks"])
PY
