#!/bin/bash
# First GPU call of the next round:
#   1. the whole -m gpu suite on the final code of round 1 (the unconfirmed tests should XPASS),
#   2. the fuzz on the real kernels,
#   3. config 4 under the scheduling knobs the offline model proposes (tests/tools/bvh_sched_model.py).
# Usage:  gpurun --timeout 900 -- 'bash tests/tools/round2_first_call.sh'
# Everything lands in gpurun_out/round2_first/.
set -u
out=gpurun_out/round2_first
mkdir -p "$out"
timeout 300 python -m pytest tests -m gpu -q -rxX 2>&1 | tail -40 > "$out/pytest_gpu.txt"
tail -3 "$out/pytest_gpu.txt"
timeout 200 python tests/tools/fuzz_ref.py --gpu --minutes 2.5 --n 4000 --seed 11 --keep-going > "$out/fuzz_gpu.log" 2>&1
tail -1 "$out/fuzz_gpu.log"
for n in 100000 400000; do
  for knobs in "" "HFB_BVH_BPS=2" "HFB_BVH_BPS=1" "HFB_BVH_QUORUM=1" "HFB_BVH_ORDER=1" \
               "HFB_BVH_QUORUM=1 HFB_BVH_BPS=2" "HFB_BVH_ORDER=1 HFB_BVH_QUORUM=1 HFB_BVH_BPS=2" \
               "HFB_BVH_ORDER=1 HFB_BVH_QUORUM=1 HFB_BVH_BPS=1"; do
    tag=$(echo "n${n}_${knobs:-default}" | tr ' =' '__')
    env $knobs timeout 120 python tests/tools/bench_bvh.py $n > "$out/bvh_$tag.json" 2> "$out/bvh_$tag.err"
    echo "$tag $(python -c "import json,sys; d=json.loads(open('$out/bvh_$tag.json').read().strip().splitlines()[-1]); print({k: d[k] for k in d if 'ms' in k or 'per_s' in k})" 2>/dev/null)"
  done
done
