#!/bin/bash
# A/B of the Gram top-k kernels: tools/ab_gram.sh <variant>...  runs `bench.py --workload c3` once per
# hyptokenizer_b200/lib/libhyptok_<variant>.so on the same box, then the Gram / top-k parity tests on the last one.
# Build a variant with:  make -C hyptokenizer_b200/csrc LIB=../lib/libhyptok_x.so OBJDIR=../lib/obj/x [EXTRA=-D...]
for v in "$@"; do
  export HYPTOK_B200_LIB=$PWD/hyptokenizer_b200/lib/libhyptok_$v.so
  echo "== $v"
  python bench.py --workload c3 --steps 40 --warmup 5 --no-tf32-peak 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['ms_per_step'], l['recall']['bit_identical_to_exact_kernel'])"
done
python -m pytest tests/test_gpu_gram_tc.py tests/test_gpu_topk.py -x -q -m gpu 2>&1 | tail -2
