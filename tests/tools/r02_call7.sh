#!/bin/bash
set -u
out=gpurun_out/r02g
mkdir -p "$out"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > "$out/pytest_gpu.txt"
tail -3 "$out/pytest_gpu.txt"
( time timeout 1500 python bench.py --steps 10 --warmup 3 ) > "$out/bench.json" 2> "$out/bench.err"
tail -3 "$out/bench.err"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02g/bench.json') if l.startswith('{')][-1])
    print('config2 value %.3g e2e %.3g rows %.3g min %.3g cpu %.3g' % (d['value'], d['e2e']['value'], d['e2e_pair_rows']['value'], d['e2e_min_distance_only']['value'], d['cpu_baseline']['value']))
    print('kernels', d['kernels'])
    for k,v in d['workloads'].items():
        print(k, {a: (v[a] if not isinstance(v[a], dict) else v[a].get('value', v[a].get('frac'))) for a in v if a in ('value','e2e','roofline','cpu_baseline','error')})
    print('support', d['convex_support_kernel']['frac'])
except Exception as e:
    print('bench parse failed', e)
PY
