#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 8 2 4 6; do
  HYP_TC_ENGINE=ss HYP_TC_DEBUG=$dbg HYP_TC_TIMING=1 timeout 200 python bench.py --workload c3 --steps 2 --warmup 1 --c3-min-steps 2 --no-tf32-peak > /dev/null 2> gpurun_out/r2m_dbg$dbg.err; echo "ss debug=$dbg rc=$?"; grep hyp_gram_topk gpurun_out/r2m_dbg$dbg.err | tail -1
done
