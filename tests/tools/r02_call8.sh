#!/bin/bash
set -u
out=gpurun_out/r02h
mkdir -p "$out"
timeout 900 python -m pytest tests/test_broadphase.py tests/test_cpp_host_api.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -30 > "$out/pytest.txt"
tail -4 "$out/pytest.txt"
( time timeout 900 python bench.py --workload config5 --steps 10 --warmup 3 ) > "$out/bench5.json" 2> "$out/bench5.err"
tail -4 "$out/bench5.err"; cut -c1-1500 "$out/bench5.json"
for c in 131072 262144 524288 1048576; do
  HFB_CHUNK=$c timeout 300 python - <<PY > "$out/e2e_chunk_$c.txt" 2>&1
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import hppfcl_b200 as hf
from hppfcl_b200 import _pod as P, workloads as W
s = W.config2_scene()
eng = hf.Engine(0)
hs = eng.register_shapes(s["shapes"]); eng.commit()
def pin(a):
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).pin_memory()
    return t, t.numpy().view(a.dtype).reshape(a.shape)
keep = [pin(a) for a in (hs[s["obj_h"]].astype(np.uint32), s["obj_tf"], s["first"], s["second"])]
out = torch.empty(1_000_000 * 96, dtype=torch.uint8).pin_memory().numpy().view(P.distance_result_dtype)
dmin = torch.empty(1_000_000, dtype=torch.float64).pin_memory().numpy()
req = P.DistanceRequestPOD()
for mode in ("full", "min"):
    f = (lambda: eng.batch_distance_objects(keep[0][1], keep[1][1], keep[2][1], keep[3][1], req, out=out)) if mode == "full" else \
        (lambda: eng.batch_distance_objects(keep[0][1], keep[1][1], keep[2][1], keep[3][1], req, out=dmin, min_only=True))
    for _ in range(3): f()
    t0 = time.perf_counter()
    for _ in range(10): f()
    dt = (time.perf_counter() - t0) / 10
    print("chunk $c", mode, "%.3f ms  %.3g pairs/s" % (dt * 1e3, 1e6 / dt))
PY
  cat "$out/e2e_chunk_$c.txt" | tail -2
done
