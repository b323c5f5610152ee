#!/bin/bash
set -u
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.4g e2e %.4g frac %.4f launches %d' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches']))
for k,v in d['workloads'].items():
    print(k, '%.4g'%v['value'] if 'value' in v else v)
"
