#!/bin/bash
set -u
out=gpurun_out/r02r
mkdir -p "$out"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fuzz_reference.py -m gpu -x -q 2>&1 | tail -3 > "$out/pytest.txt"; cat "$out/pytest.txt"
for r in 8192 0; do
  HFB_EPA_RESUME=$r timeout 120 python tests/tools/bench_pairs.py config2 1000000 > "$out/pairs_c2_r$r.json" 2>&1; echo "RESUME=$r $(tail -1 "$out/pairs_c2_r$r.json" | cut -c1-330)"
  HFB_EPA_RESUME=$r timeout 120 python tests/tools/bench_pairs.py config3 1000000 > "$out/pairs_c3_r$r.json" 2>&1; echo "RESUME=$r $(tail -1 "$out/pairs_c3_r$r.json" | cut -c1-330)"
done
run() {
  n=$1; shift
  tag=$(echo "n${n}_$*" | tr ' =' '__')
  env "$@" timeout 200 python tests/tools/bench_bvh.py $n > "$out/bvh_$tag.json" 2> "$out/bvh_$tag.err"
  echo "$tag $(python -c "import json,sys; d=json.loads(open('$out/bvh_$tag.json').read().strip().splitlines()[-1]); print({k: (d[k] if not isinstance(d[k], dict) else {a: round(b, 1) for a, b in d[k].items()}) for k in d if 'ms_per' in k or 'queries_per_s' == k or 'identical' in k or 'watchdog' in k})" 2>/dev/null)"
}
run 100000 HFB_BVH_WARPS=8
run 100000 HFB_BVH_WARPS=12
