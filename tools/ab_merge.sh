#!/bin/bash
# A/B the merge-loop kernel: tools/ab_merge.sh <variant>...  runs bench.py (c2, device-timed) once per
# hyptokenizer_b200/lib/libhyptok_<variant>.so on the same GPU box; numbers across boxes differ by a few percent.
# Variants are built with: make -C hyptokenizer_b200/csrc LIB=../lib/libhyptok_X.so OBJDIR=../lib/obj/X EXTRA=-D...
for rep in 1 2; do
for v in "$@"; do
  echo -n "$v: "
  HYPTOK_B200_LIB=$PWD/hyptokenizer_b200/lib/libhyptok_$v.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/tmp/ab_err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['roofline']['kernel_ms'])"
  grep -E "marks|phases" /tmp/ab_err.txt || true
done; done
