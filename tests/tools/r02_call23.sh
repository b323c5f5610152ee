#!/bin/bash
set -u
out=gpurun_out/r02w
mkdir -p "$out"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fuzz_reference.py -m gpu -x -q 2>&1 | tail -3 > "$out/pytest.txt"; cat "$out/pytest.txt"
for p in 10 8 12 10 "10,10"; do
  HFB_GJK_PASSES=$p timeout 120 python tests/tools/bench_pairs.py config2 1000000 2>&1 | tail -1 | cut -c1-250 | sed "s/^/PASSES=$p /"
done
HFB_EPA_OVERLAP=1 timeout 120 python tests/tools/bench_pairs.py config2 1000000 2>&1 | tail -1 | cut -c1-250 | sed "s/^/EPA_OVERLAP=1 /"
HFB_GJK_PASSES=10 timeout 120 python tests/tools/bench_pairs.py config2 1000000 2 2>&1 | tail -1 | cut -c1-250 | sed "s/^/NESTEROV PASSES=10 /"
