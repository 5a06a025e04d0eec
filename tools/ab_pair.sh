#!/bin/bash
# A/B the pair-count kernel: tools/ab_pair.sh <variant>...  runs the parity tests and tools/bench_pair.py once per
# hyptokenizer_b200/lib/libhyptok_<variant>.so on the same GPU box (numbers across boxes differ by a few percent).
# Variants are built with, e.g.:  make -C hyptokenizer_b200/csrc LIB=../lib/libhyptok_s32.so OBJDIR=../lib/obj/s32 EXTRA=-DHYP_PC_SYMS=32
# (HYP_PC_SYMS = dimension of the private table: 28 -> 16 warps per SM, 32 -> 12).  HYP_PAIR_COUNT=v1 selects the
# atomics-only kernel inside any build.
MB=${MB:-1024}
for v in "$@"; do
  export HYPTOK_B200_LIB=$PWD/hyptokenizer_b200/lib/libhyptok_$v.so
  echo "== $v"
  python -m pytest tests/test_gpu_paircount.py -x -q -m gpu 2>&1 | tail -1
  python tools/bench_pair.py "$MB" 2>&1 | tail -3
done
