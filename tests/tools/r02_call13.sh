#!/bin/bash
set -u
out=gpurun_out/r02m
mkdir -p "$out"
timeout 900 python -m pytest tests/test_bvh_parity.py tests/test_gpu_unconfirmed.py -m gpu -x -q 2>&1 | tail -5 > "$out/pytest.txt"
tail -2 "$out/pytest.txt"
run() {
  n=$1; shift
  tag=$(echo "n${n}_$*" | tr ' =' '__')
  env "$@" timeout 200 python tests/tools/bench_bvh.py $n > "$out/bvh_$tag.json" 2> "$out/bvh_$tag.err"
  echo "$tag $(python -c "import json,sys; d=json.loads(open('$out/bvh_$tag.json').read().strip().splitlines()[-1]); print({k: (d[k] if not isinstance(d[k], dict) else {a: round(b, 1) for a, b in d[k].items()}) for k in d if 'ms_per' in k or 'queries_per_s' == k or 'identical' in k or 'watchdog' in k or 'phase' in k})" 2>/dev/null)"
}
run 100000 HFB_BVH_SPEC=-1
run 100000 HFB_BVH_SPEC=200
run 100000 HFB_BVH_SPEC=200 HFB_BVH_GENS=1
run 100000 HFB_BVH_SPEC=200 HFB_BVH_GENS=3
run 100000 HFB_BVH_SPEC=300 HFB_BVH_SPEC_BIG=300
run 400000 HFB_BVH_SPEC=200
HFB_BVH_SPEC=200 timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_bvhq -s 3 -c 1 -o "$out/k_bvhq" python tests/tools/bench_bvh.py 100000 > "$out/ncu.log" 2>&1
tail -2 "$out/ncu.log"
