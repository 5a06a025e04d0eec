#!/bin/bash
# round-2 GPU call R (1 GPU): pair counting after the stream-wide rank table -- parity, the three kernels on the three
# streams, the c4 bench line with the English-letter figure
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_paircount.py tests/test_gpu_freq.py -x -q -m gpu 2>&1 | tail -2
for v in default v1 v2 v3; do
  if [ $v = default ]; then unset HYP_PAIR_COUNT; else export HYP_PAIR_COUNT=$v; fi
  python tools/bench_pair.py 1024 2>&1 | tail -3
done
unset HYP_PAIR_COUNT
timeout 600 python bench.py --workload c4 > gpurun_out/r2r_bench_c4.json 2> gpurun_out/r2r_c4.err; echo "c4 rc=$?"; tail -2 gpurun_out/r2r_c4.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2r_bench_c4.json").read().strip().splitlines()[-1])
print(l["value"], l["roofline"]["frac"], {k:v for k,v in l["config"].items() if k!="workload"})
PY
