#!/bin/bash
set -u
out=gpurun_out/r02s
mkdir -p "$out"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > "$out/pytest.txt"; cat "$out/pytest.txt"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_pairs<.int.2,' -s 1 -c 1 -o /tmp/k_pairs_c3 python tests/tools/bench_pairs.py config3 1000000 > "$out/ncu_c3a.log" 2>&1; tail -2 "$out/ncu_c3a.log"
timeout 300 python tests/tools/ncu_digest.py /tmp/k_pairs_c3.ncu-rep "$out/digest_k_pairs_c3.txt"; head -24 "$out/digest_k_pairs_c3.txt"
