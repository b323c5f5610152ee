"""e2e rate of hfb_batch_distance_objects on the config-2 scene (pinned host buffers, full result rows back), without
the rest of bench.py: python tests/tools/bench_e2e.py [pairs]   (knobs through the environment, e.g. HFB_CHUNK)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hppfcl_b200 as hf  # noqa: E402
from hppfcl_b200 import _pod as P, workloads as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
w = W.config2_scene(n_objects=100_000, n_pairs=n, seed=0xFC1 + 2)
eng = hf.Engine(0)
hs = eng.register_shapes(w["shapes"])
eng.commit()


def pinned(a):
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).pin_memory()
    return t, t.numpy().view(a.dtype).reshape(a.shape)


keep, ph = [], []
for a in (hs[w["obj_h"] % len(hs)].astype(np.uint32), w["obj_tf"], w["first"], w["second"]):
    t, v = pinned(a)
    keep.append(t)
    ph.append(v)
t_out = torch.empty(n * P.distance_result_dtype.itemsize, dtype=torch.uint8).pin_memory()
host_out = t_out.numpy().view(P.distance_result_dtype)
t_min = torch.empty(n, dtype=torch.float64).pin_memory()
req = P.DistanceRequestPOD()
out = {}
for name, call in (("rows96", lambda: eng.batch_distance_objects(ph[0], ph[1], ph[2], ph[3], req, out=host_out)),
                   ("min_only", lambda: eng.batch_distance_objects(ph[0], ph[1], ph[2], ph[3], req, out=t_min.numpy(), min_only=True))):
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    k = 10
    for _ in range(k):
        call()
    torch.cuda.synchronize()
    out[name] = n * k / (time.perf_counter() - t0)
out["env"] = {k: v for k, v in os.environ.items() if k.startswith("HFB_")}
print(json.dumps(out))
