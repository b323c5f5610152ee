#!/bin/bash
# final confirmation of the round on one B200: tests, smoke, the bench lines DESIGN.md quotes, the launch list and
# the DRAM traffic of the dominant kernels
set -u
out=gpurun_out/r02_final
mkdir -p "$out"
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,clocks.sm --format=csv,noheader
timeout 120 python tests/tools/probe_pcie_numa.py | cut -c1-400
for hs in 1 0; do HFB_HULL_SORT=$hs timeout 120 python tests/tools/bench_pairs.py config3 1000000 2>&1 | tail -1 | cut -c1-330 | sed "s/^/HULL_SORT=$hs /"; done
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > "$out/pytest.txt"; cat "$out/pytest.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > "$out/bench.json" 2> "$out/bench.err"; tail -c 600 "$out/bench.json"; echo
timeout 600 python bench.py --workload config5 > "$out/bench_config5.json" 2> "$out/bench_config5.err"; tail -c 400 "$out/bench_config5.json"; echo
timeout 600 python bench.py --impl reference > "$out/bench_reference.json" 2> "$out/bench_reference.err"; tail -c 400 "$out/bench_reference.json"; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$out/launches_bench.csv" python bench.py --steps 2 --warmup 1 > "$out/bench_under_ncu.log" 2>&1; tail -1 "$out/launches_bench.csv" | cut -c1-200
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_gjk --csv --log-file "$out/traffic_c2.csv" python tests/tools/bench_pairs.py config2 1000000 > "$out/traffic_c2.log" 2>&1
python tests/tools/ncu_traffic.py "$out/traffic_c2.csv" k_gjk k_gjk_first | tee "$out/traffic_c2.json"
timeout 120 python tests/tools/bench_pairs.py config2 1000000 2>&1 | tail -1 | cut -c1-330
timeout 120 python tests/tools/bench_pairs.py config3 1000000 2>&1 | tail -1 | cut -c1-330
timeout 200 python tests/tools/bench_bvh.py 100000 2>/dev/null | tail -1 | cut -c1-400
