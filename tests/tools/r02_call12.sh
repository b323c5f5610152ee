#!/bin/bash
# k_bvhq spec rework + GJK passes with overlapped EPA + the full suite / bench / launch lists of call 10
set -u
bash tests/tools/r02_call11.sh
out=gpurun_out/r02l
mkdir -p "$out"
for e in 1 0; do
  HFB_EPA_OVERLAP=$e timeout 200 python tests/tools/bench_pairs.py config2 > "$out/pairs_overlap$e.json" 2> "$out/pairs_overlap$e.err"
  echo "epa_overlap=$e $(python -c "import json; d=json.loads(open('$out/pairs_overlap$e.json').read().strip().splitlines()[-1]); print(round(d['pairs_per_s']/1e6,1), 'Mpairs/s', d['ms_per_step'], d['kernels_ms'], d['checksum'])" 2>/dev/null)"
done
HFB_GJK_PASSES=4,4 timeout 200 python tests/tools/bench_pairs.py config2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('passes 4,4 overlap', round(d['pairs_per_s']/1e6,1), d['ms_per_step'])"
bash tests/tools/r02_call10.sh
