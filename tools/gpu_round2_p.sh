#!/bin/bash
# round-2 GPU call P (8 GPUs, final): multi-GPU tests, then the default line at N = 2, 4, 8
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_topk.py tests/test_gpu_multidevice.py tests/test_gpu_paircount.py -m gpu -q -k "multi_gpu or nccl or second_device" > gpurun_out/r2p_tests_n8.log 2>&1; echo "multi-gpu tests rc=$?"; tail -3 gpurun_out/r2p_tests_n8.log
for N in 2 4 8; do
  NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N > gpurun_out/r02_bench_default_n$N.json 2> gpurun_out/r2p_bench_n$N.err; echo "bench N=$N rc=$?"
  python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/r02_bench_default_n$N.json").read().strip().splitlines()[-1])
    print("N=$N", {k:l[k] for k in ("value","ms_per_step")}, "e2e", l["e2e"]["value"], l["clocks"])
    print("   ", {k:(round(v,4) if isinstance(v,float) else v) for k,v in l["config"].items() if k.startswith("c3_") or k.startswith("c4_")})
except Exception as e:
    print("bench parse failed", e)
PY
done
HYP_TC_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 8 --workload c3 --steps 2 --warmup 1 --c3-min-steps 2 --no-tf32-peak > /dev/null 2> gpurun_out/r2p_c3_n8_timing.err; grep "hyp_gram_topk" gpurun_out/r2p_c3_n8_timing.err | tail -2
