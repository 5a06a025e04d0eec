#!/bin/bash
# round-2 GPU call S (1 GPU, final after the pair-count rank table): full suite, smoke, default bench, reference arm,
# then the pair-count launch list and full capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r2s_tests.log 2>&1; echo "tests rc=$?"; tail -10 gpurun_out/r2s_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2s_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_default_n1.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r2s_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r02_bench_default_n1.json").read().strip().splitlines()[-1])
print({k:l[k] for k in ("value","ms_per_step","clocks","gpu_launches")}, l["roofline"]["frac"], l["roofline"]["smem_frac"], l["e2e"]["value"])
print({k:v for k,v in l["config"].items() if k.startswith("c3_") or k.startswith("c4_")})
print(json.dumps(l["config"].get("snapshot_cache_semantics")))
r=json.loads(open("gpurun_out/r02_bench_reference.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["ms_per_step"], r["cpu_baseline"]["cores"])
PY
C4="python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline"
$C4 > gpurun_out/r2s_plain_c4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_c4_launches.csv $C4 > gpurun_out/r2s_ncu3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pair_count_v3 -s 1 -c 1 -o gpurun_out/r02_pc_v3 $C4 > gpurun_out/r2s_ncu5.log 2>&1
echo "ncu chain rc=$?"
