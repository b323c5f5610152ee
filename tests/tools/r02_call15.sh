#!/bin/bash
set -u
out=gpurun_out/r02o
mkdir -p "$out"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > "$out/pytest.txt"; cat "$out/pytest.txt"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_pairs -s 2 -c 1 -o "$out/k_pairs_c3" python tests/tools/bench_pairs.py config3 1000000 > "$out/ncu_c3a.log" 2>&1; tail -2 "$out/ncu_c3a.log"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_epa -s 4 -c 2 -o "$out/k_epa_c3" python tests/tools/bench_pairs.py config3 1000000 > "$out/ncu_c3b.log" 2>&1; tail -2 "$out/ncu_c3b.log"
ls -la "$out"
