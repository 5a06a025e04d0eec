#!/bin/bash
# round-2 GPU call Q (1 GPU): A/B of the pair-count wrap handling (w0 = positions per group, w1 = per-step detour,
# w2 = per-step vote), then the default bench line with the snapshot-mode measurement
mkdir -p gpurun_out
MB=1024 bash tools/ab_pair.sh w0 w1 w2 > gpurun_out/r2q_ab.log 2>&1; cat gpurun_out/r2q_ab.log
unset HYPTOK_B200_LIB
timeout 900 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/r2q_bench_c2.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2q_bench.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r2q_bench_c2.json").read().strip().splitlines()[-1])
print(l["value"], l["e2e"]["value"], json.dumps(l["config"].get("snapshot_cache_semantics")))
PY
