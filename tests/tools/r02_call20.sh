#!/bin/bash
set -u
out=gpurun_out/r02t
mkdir -p "$out"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 > "$out/pytest.txt"; cat "$out/pytest.txt"
for o in 1 0 1 0; do
  HFB_GJK_ORDERED=$o timeout 120 python tests/tools/bench_pairs.py config2 1000000 > "$out/pairs_c2_o$o.json" 2>&1; echo "ORDERED=$o $(tail -1 "$out/pairs_c2_o$o.json" | cut -c1-330)"
done
HFB_GJK_ORDERED=1 HFB_GJK_PASSES=4,6 timeout 120 python tests/tools/bench_pairs.py config2 1000000 2>&1 | tail -1 | cut -c1-330
HFB_GJK_ORDERED=1 HFB_GJK_PASSES=5 timeout 120 python tests/tools/bench_pairs.py config2 1000000 2>&1 | tail -1 | cut -c1-330
HFB_GJK_ORDERED=1 HFB_GJK_PASSES=8 timeout 120 python tests/tools/bench_pairs.py config2 1000000 2>&1 | tail -1 | cut -c1-330
