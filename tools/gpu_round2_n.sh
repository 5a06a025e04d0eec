#!/bin/bash
# round-2 GPU call N (1 GPU, final): full suite, smoke, default bench, reference arm, c5, then the ncu captures of the round
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2n_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2n_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_default_n1.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r2n_ref.err; echo "ref rc=$?"
timeout 400 python bench.py --workload c5 > gpurun_out/r02_bench_c5.json 2> gpurun_out/r2n_c5.err; echo "c5 rc=$?"
timeout 400 python bench.py --workload c5 --c5-v0 10000 --c5-steps 120 > gpurun_out/r02_bench_c5_v10k.json 2> gpurun_out/r2n_c5b.err; echo "c5 10k rc=$?"
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r02_bench_default_n1.json").read().strip().splitlines()[-1])
print({k:l[k] for k in ("value","ms_per_step","clocks","gpu_launches")}, l["roofline"]["frac"], l["roofline"]["smem_frac"], l["e2e"]["value"])
print({k:v for k,v in l["config"].items() if k.startswith("c3_") or k.startswith("c4_")})
r=json.loads(open("gpurun_out/r02_bench_reference.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["ms_per_step"], r["cpu_baseline"]["cores"], r["config"]["as_shipped_merges_per_s"])
PY
C2="python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
C3="python bench.py --workload c3 --steps 2 --warmup 1 --c3-min-steps 2 --no-tf32-peak"
C4="python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline"
$C2 > gpurun_out/r2n_plain_c2.log 2>&1 && $C3 > gpurun_out/r2n_plain_c3.log 2>&1 && $C4 > gpurun_out/r2n_plain_c4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_c2_launches.csv $C2 > gpurun_out/r2n_ncu1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_c3_launches.csv $C3 > gpurun_out/r2n_ncu2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_c4_launches.csv $C4 > gpurun_out/r2n_ncu3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gram_tc_kernel -s 2 -c 2 -o gpurun_out/r02_gram $C3 > gpurun_out/r2n_ncu4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pair_count_v3 -s 1 -c 1 -o gpurun_out/r02_pc_v3 $C4 > gpurun_out/r2n_ncu5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc_finish|kth_select" -s 2 -c 2 -o gpurun_out/r02_finish $C3 > gpurun_out/r2n_ncu6.log 2>&1
echo "ncu chain rc=$?"
