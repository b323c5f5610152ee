#!/bin/bash
set -u
out=gpurun_out/r02u
mkdir -p "$out"
for p in 6 7 8 9 10 12 16 "8,8" "10,10"; do
  HFB_GJK_PASSES=$p timeout 120 python tests/tools/bench_pairs.py config2 1000000 2>&1 | tail -1 | cut -c1-250 | sed "s/^/PASSES=$p /"
done
HFB_GJK_PASSES=8 timeout 120 python tests/tools/bench_pairs.py config2 1000000 2 2>&1 | tail -1 | cut -c1-250 | sed "s/^/NESTEROV PASSES=8 /"
HFB_GJK_PASSES=6 timeout 120 python tests/tools/bench_pairs.py config2 1000000 2 2>&1 | tail -1 | cut -c1-250 | sed "s/^/NESTEROV PASSES=6 /"
timeout 120 python tests/tools/bench_pairs.py config3 1000000 2>&1 | tail -1 | cut -c1-330 | sed "s/^/C3 unroll8 /"
