#!/bin/bash
# Round 2, GPU call 1: the task-system mesh-shape walk (k_bvhq) -- parity, config 4 under its knobs next to the
# lane-per-query kernel, sanitizer, the whole -m gpu suite, one ncu capture.
set -u
out=gpurun_out/r02a
mkdir -p "$out"
nvidia-smi -L | head -2
timeout 900 python -m pytest tests/test_bvh_parity.py tests/test_gpu_unconfirmed.py -m gpu -x -q 2>&1 | tail -15 > "$out/pytest_bvh.txt"
tail -3 "$out/pytest_bvh.txt"
for n in 100000; do
  for knobs in "HFB_BVHQ=0" "HFB_BVH_SPEC=-1" "HFB_BVH_SPEC=-1 HFB_BVH_ORDER=1" "HFB_BVH_SPEC=0" "HFB_BVH_SPEC=0 HFB_BVH_ORDER=1" \
               "HFB_BVH_SPEC=200" "HFB_BVH_SPEC=200 HFB_BVH_ORDER=1" "HFB_BVH_SPEC=1000 HFB_BVH_ORDER=1" "HFB_BVH_SPEC=50 HFB_BVH_ORDER=1"; do
    tag=$(echo "n${n}_${knobs}" | tr ' =' '__')
    env $knobs timeout 200 python tests/tools/bench_bvh.py $n > "$out/bvh_$tag.json" 2> "$out/bvh_$tag.err"
    echo "$tag $(python -c "import json,sys; d=json.loads(open('$out/bvh_$tag.json').read().strip().splitlines()[-1]); print({k: d[k] for k in d if 'ms' in k or 'queries_per_s' == k or 'identical' in k or 'watchdog' in k})" 2>/dev/null)"
  done
done
for knobs in "HFB_BVH_SPEC=200 HFB_BVH_ORDER=1" "HFB_BVH_SPEC=0"; do
  tag=$(echo "n400000_${knobs}" | tr ' =' '__')
  env $knobs timeout 200 python tests/tools/bench_bvh.py 400000 > "$out/bvh_$tag.json" 2> "$out/bvh_$tag.err"
  echo "$tag $(python -c "import json,sys; d=json.loads(open('$out/bvh_$tag.json').read().strip().splitlines()[-1]); print({k: d[k] for k in d if 'ms' in k or 'queries_per_s' == k or 'identical' in k or 'watchdog' in k})" 2>/dev/null)"
done
timeout 400 compute-sanitizer --tool racecheck --print-limit 20 python tests/tools/sanitize_run.py 1500 > "$out/racecheck.txt" 2>&1
tail -4 "$out/racecheck.txt"
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python tests/tools/sanitize_run.py 1500 > "$out/memcheck.txt" 2>&1
tail -3 "$out/memcheck.txt"
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > "$out/pytest_gpu.txt"
tail -3 "$out/pytest_gpu.txt"
HFB_BVH_SPEC=200 HFB_BVH_ORDER=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_bvhq$ -s 2 -c 1 -o "$out/k_bvhq" python tests/tools/bench_bvh.py 100000 > "$out/ncu.log" 2>&1
tail -2 "$out/ncu.log"
