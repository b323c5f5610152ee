#!/bin/bash
set -u
out=gpurun_out/r02x
mkdir -p "$out"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > "$out/pytest.txt"; cat "$out/pytest.txt"
for k in 1 2; do timeout 120 python tests/tools/bench_pairs.py config2 1000000 2>&1 | tail -1 | cut -c1-250 | sed "s/^/C2 /"; done
timeout 120 python tests/tools/bench_pairs.py config3 1000000 2>&1 | tail -1 | cut -c1-300 | sed "s/^/C3 /"
timeout 600 python bench.py --workload config5 --steps 5 2>/dev/null | tail -c 700
