#!/bin/bash
set -u
out=gpurun_out/r02n
mkdir -p "$out"
timeout 120 python tests/tools/probe_pcie_numa.py > "$out/pcie_numa.json" 2> "$out/pcie_numa.err"; cat "$out/pcie_numa.json"; tail -2 "$out/pcie_numa.err"
nvidia-smi topo -m > "$out/topo.txt" 2>&1; head -6 "$out/topo.txt"
lscpu | grep -i "numa\|model name\|socket\|^cpu(s)" > "$out/lscpu.txt"; cat "$out/lscpu.txt"
run() {
  n=$1; shift
  tag=$(echo "n${n}_$*" | tr ' =' '__')
  env "$@" timeout 200 python tests/tools/bench_bvh.py $n > "$out/bvh_$tag.json" 2> "$out/bvh_$tag.err"
  echo "$tag $(python -c "import json,sys; d=json.loads(open('$out/bvh_$tag.json').read().strip().splitlines()[-1]); print({k: (d[k] if not isinstance(d[k], dict) else {a: round(b, 1) for a, b in d[k].items()}) for k in d if 'ms_per' in k or 'queries_per_s' == k or 'identical' in k or 'watchdog' in k or 'phase' in k})" 2>/dev/null)"
}
run 100000 HFB_BVH_WARPS=8
run 100000 HFB_BVH_WARPS=12
run 100000 HFB_BVH_WARPS=12 HFB_BVH_GENS=3
timeout 120 python tests/tools/bench_pairs.py config3 1000000 > "$out/pairs_c3.json" 2>&1; tail -1 "$out/pairs_c3.json" | cut -c1-600
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_pairs -s 2 -c 1 -o "$out/k_pairs_c3" python tests/tools/bench_pairs.py config3 1000000 > "$out/ncu_c3a.log" 2>&1; tail -2 "$out/ncu_c3a.log"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_epa -s 4 -c 2 -o "$out/k_epa_c3" python tests/tools/bench_pairs.py config3 1000000 > "$out/ncu_c3b.log" 2>&1; tail -2 "$out/ncu_c3b.log"
