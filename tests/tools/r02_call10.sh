#!/bin/bash
set -u
out=gpurun_out/r02j
mkdir -p "$out"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > "$out/pytest_gpu.txt"
tail -3 "$out/pytest_gpu.txt"
( time timeout 1500 python bench.py --steps 10 --warmup 3 ) > "$out/bench.json" 2> "$out/bench.err"
tail -3 "$out/bench.err"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02j/bench.json') if l.startswith('{')][-1])
    print('config2 value %.3g e2e %.3g rows %.3g min %.3g cpu %.3g' % (d['value'], d['e2e']['value'], d['e2e_pair_rows']['value'], d['e2e_min_distance_only']['value'], d['cpu_baseline']['value']))
    for k,v in d['workloads'].items():
        print(k, {a: (v[a] if not isinstance(v[a], dict) else v[a].get('value', v[a].get('frac'))) for a in v if a in ('value','e2e','roofline','cpu_baseline','error')})
    print('support', d['convex_support_kernel']['frac'])
except Exception as e:
    print('bench parse failed', e)
PY
# launch list of the bench command (shares only; cold caches)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file "$out/launches_bench.csv" python bench.py --steps 2 --warmup 3 > "$out/ncu_bench.log" 2>&1
tail -1 "$out/ncu_bench.log" | cut -c1-200
# DRAM traffic of the dominant kernels
HFB_GJK_PASSES=6 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_gjk_|k_bvhq$|k_pairs|k_epa" -s 30 -c 40 --csv --log-file "$out/traffic_c2.csv" python tests/tools/bench_pairs.py config2 > /dev/null 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_pairs|k_epa" -s 6 -c 12 --csv --log-file "$out/traffic_c3.csv" python tests/tools/bench_pairs.py config3 > /dev/null 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_bvhq" -s 4 -c 4 --csv --log-file "$out/traffic_c4.csv" python tests/tools/bench_bvh.py 100000 > /dev/null 2>&1
ls -la "$out"
