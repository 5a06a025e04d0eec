#!/bin/bash
# round-2 GPU call B (N >= 2 GPUs): multi-GPU tests (NCCL + peer-memory all-gather, second device), smoke, default bench under torchrun
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_topk.py tests/test_gpu_multidevice.py tests/test_gpu_paircount.py -m gpu -x -q -rs > gpurun_out/r2b_tests_n$N.log 2>&1; echo "multi-gpu tests rc=$?"
tail -8 gpurun_out/r2b_tests_n$N.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2b_smoke.log
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/r2b_bench_n$N.json 2> gpurun_out/r2b_bench_n$N.err; echo "bench N=$N rc=$?"
tail -3 gpurun_out/r2b_bench_n$N.err
python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/r2b_bench_n$N.json").read().strip().splitlines()[-1])
    print({k:l[k] for k in ("value","ms_per_step","n_gpus")}, l["roofline"]["frac"], l.get("e2e"))
    print({k:v for k,v in l["config"].items() if k.startswith("c3_") or k.startswith("c4_")})
    print(l["secondary"]["c3"]["config"]["exchange"], l["secondary"]["c3"]["config"]["barrier_status"])
except Exception as e:
    print("bench parse failed", e)
PY
