#!/bin/bash
set -u
out=gpurun_out/r02p
mkdir -p "$out"
timeout 900 python -m pytest tests/test_plane_halfspace.py tests/test_cpp_host_api.py -m gpu -x -q 2>&1 | tail -8 > "$out/pytest.txt"; cat "$out/pytest.txt"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_pairs -s 2 -c 1 -o /tmp/k_pairs_c3 python tests/tools/bench_pairs.py config3 1000000 > "$out/ncu_c3a.log" 2>&1; tail -2 "$out/ncu_c3a.log"
timeout 300 python tests/tools/ncu_digest.py /tmp/k_pairs_c3.ncu-rep "$out/digest_k_pairs_c3.txt"; head -20 "$out/digest_k_pairs_c3.txt"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_epa -s 4 -c 2 -o /tmp/k_epa_c3 python tests/tools/bench_pairs.py config3 1000000 > "$out/ncu_c3b.log" 2>&1; tail -2 "$out/ncu_c3b.log"
timeout 300 python tests/tools/ncu_digest.py /tmp/k_epa_c3.ncu-rep "$out/digest_k_epa_c3.txt"; head -36 "$out/digest_k_epa_c3.txt"
