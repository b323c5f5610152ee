#!/bin/bash
set -u
out=gpurun_out/r02q
mkdir -p "$out"
timeout 600 python -m pytest tests/test_bvh_parity.py -m gpu -x -q 2>&1 | tail -3 > "$out/pytest.txt"; cat "$out/pytest.txt"
run() {
  n=$1; shift
  tag=$(echo "n${n}_$*" | tr ' =' '__')
  env "$@" timeout 200 python tests/tools/bench_bvh.py $n > "$out/bvh_$tag.json" 2> "$out/bvh_$tag.err"
  echo "$tag $(python -c "import json,sys; d=json.loads(open('$out/bvh_$tag.json').read().strip().splitlines()[-1]); print({k: (d[k] if not isinstance(d[k], dict) else {a: round(b, 1) for a, b in d[k].items()}) for k in d if 'ms_per' in k or 'queries_per_s' == k or 'identical' in k or 'watchdog' in k or 'phase' in k})" 2>/dev/null)"
}
run 100000 HFB_BVH_ADAPT=0
run 100000 HFB_BVH_ADAPT=1
run 100000 HFB_BVH_ADAPT=1 HFB_BVH_GENS=3
run 400000 HFB_BVH_ADAPT=1
for ge in 8 4 16 32; do
  HFB_GE=$ge timeout 120 python tests/tools/bench_pairs.py config2 1000000 > "$out/pairs_c2_ge$ge.json" 2>&1; echo "GE=$ge $(tail -1 "$out/pairs_c2_ge$ge.json" | cut -c1-420)"
done
HFB_EPA_OVERLAP=1 timeout 120 python tests/tools/bench_pairs.py config2 1000000 > "$out/pairs_c2_ov.json" 2>&1; echo "OVERLAP $(tail -1 "$out/pairs_c2_ov.json" | cut -c1-420)"
