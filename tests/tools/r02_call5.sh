#!/bin/bash
# Round 2, GPU call 5: GJK passes (config 2), k_bvhq v5 (warp-wide votes, 8 / 16 warps).
set -u
out=gpurun_out/r02e
mkdir -p "$out"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_bvh_parity.py tests/test_gpu_unconfirmed.py tests/test_oracle_golden.py -m gpu -x -q 2>&1 | tail -5 > "$out/pytest.txt"
tail -2 "$out/pytest.txt"
for p in "0" "3,3,4" "4,4" "2,2,3,4" "3,3,4,6" "6" "4,6"; do
  HFB_GJK_PASSES=$p timeout 200 python tests/tools/bench_pairs.py config2 > "$out/pairs_$p.json" 2> "$out/pairs_$p.err"
  echo "passes=$p $(python -c "import json; d=json.loads(open('$out/pairs_$p.json').read().strip().splitlines()[-1]); print(round(d['pairs_per_s']/1e6,1), 'Mpairs/s', d['ms_per_step'], d['kernels_ms'], d['checksum'])" 2>/dev/null)"
done
run() {
  n=$1; shift
  tag=$(echo "n${n}_$*" | tr ' =' '__')
  env "$@" timeout 200 python tests/tools/bench_bvh.py $n > "$out/bvh_$tag.json" 2> "$out/bvh_$tag.err"
  echo "$tag $(python -c "import json,sys; d=json.loads(open('$out/bvh_$tag.json').read().strip().splitlines()[-1]); print({k: (d[k] if not isinstance(d[k], dict) else {a: round(b, 1) for a, b in d[k].items()}) for k in d if 'ms_per' in k or 'queries_per_s' == k or 'identical' in k or 'watchdog' in k or 'phase' in k})" 2>/dev/null)"
}
run 100000 HFB_BVH_SPEC=-1
run 100000 HFB_BVH_SPEC=-1 HFB_BVH_WARPS=16
run 100000 HFB_BVH_SPEC=200
run 100000 HFB_BVH_SPEC=200 HFB_BVH_WARPS=16
run 100000 HFB_BVH_SPEC=100 HFB_BVH_SPEC_BIG=200 HFB_BVH_ORDER=1 HFB_BVH_WARPS=16
run 100000 HFB_BVH_SPEC=200 HFB_BVH_WARPS=16 HFB_BVH_GENS=3
run 400000 HFB_BVH_SPEC=200 HFB_BVH_WARPS=16
HFB_BVH_SPEC=200 HFB_BVH_WARPS=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_bvhq -s 2 -c 1 -o "$out/k_bvhq" python tests/tools/bench_bvh.py 100000 > "$out/ncu.log" 2>&1
tail -2 "$out/ncu.log"
HFB_GJK_PASSES=3,3,4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gjk_ -s 15 -c 5 -o "$out/k_gjk" python tests/tools/bench_pairs.py config2 > "$out/ncu2.log" 2>&1
tail -2 "$out/ncu2.log"
