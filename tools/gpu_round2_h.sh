#!/bin/bash
# round-2 GPU call H (8 GPUs): multi-GPU tests, c3 at N=4 and N=8, the default line at N=8
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_topk.py tests/test_gpu_multidevice.py tests/test_gpu_paircount.py -m gpu -q -k "multi_gpu or nccl or second_device" > gpurun_out/r2h_tests_n8.log 2>&1; echo "multi-gpu tests rc=$?"; tail -5 gpurun_out/r2h_tests_n8.log
for N in 4 8; do
  NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --workload c3 > gpurun_out/r2h_c3_n$N.json 2> gpurun_out/r2h_c3_n$N.err; echo "c3 N=$N rc=$?"
  python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/r2h_c3_n$N.json").read().strip().splitlines()[-1])
    c=l["config"]; print("N=$N", "ms", l["ms_per_step"], "TFLOP/s", l["value"], "local", c["ms_local_only"], "nccl", c["ms_with_nccl_allgather"], "allgather", c["allgather_ms"], l["recall"]["bit_identical_to_exact_kernel"], c["exchange"][:20], c["barrier_status"])
except Exception as e:
    print("parse failed", e)
PY
done
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 > gpurun_out/r2h_bench_n8.json 2> gpurun_out/r2h_bench_n8.err; echo "bench N=8 rc=$?"
python - <<'PY'
import json
try:
    l=json.loads(open("gpurun_out/r2h_bench_n8.json").read().strip().splitlines()[-1])
    print({k:l[k] for k in ("value","ms_per_step","n_gpus")}, l.get("e2e"))
    print({k:v for k,v in l["config"].items() if k.startswith("c3_") or k.startswith("c4_")})
except Exception as e:
    print("bench parse failed", e)
PY
HYP_TC_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --workload c3 --steps 2 --warmup 1 --no-tf32-peak > /dev/null 2> gpurun_out/r2h_c3_n8_timing.err; grep "hyp_gram_topk" gpurun_out/r2h_c3_n8_timing.err | tail -3
