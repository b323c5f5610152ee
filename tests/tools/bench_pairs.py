"""Config 2 / config 3 device-resident timing without the CPU arm (test tooling): one JSON line with pairs/s and the
per-kernel-family times.  python tests/tools/bench_pairs.py [config2|config3] [pairs] [variant]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hppfcl_b200 as hf  # noqa: E402
from hppfcl_b200 import _pod as P, workloads as W  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "config2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
variant = int(sys.argv[3]) if len(sys.argv) > 3 else (2 if cfg == "config3" else 0)
eng = hf.Engine(0)
if cfg == "config2":
    w = W.config2_mixed_primitives(n, seed=0xFC1 + 2)
    hs = eng.register_shapes(w["shapes"])
else:
    w = W.config3_convex_pairs(n, seed=0xFC1 + 3)
    cids = [eng.register_convex(pts) for pts, _ in w["hulls"]]
    hs = eng.register_shapes(P.make_shapes([P.GEOM_CONVEX] * len(cids), np.zeros((len(cids), 3)), data=cids))
eng.commit()
h1, h2 = hs[w["h1"] % len(hs)].astype(np.uint32), hs[w["h2"] % len(hs)].astype(np.uint32)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).cuda()


d = [dev(h1), dev(w["tf1"]), dev(h2), dev(w["tf2"])]
d_out = torch.empty(n * P.distance_result_dtype.itemsize, dtype=torch.uint8, device="cuda")
req = P.DistanceRequestPOD(gjk_variant=variant)
stream = torch.cuda.current_stream().cuda_stream


def step():
    eng.batch_distance_device(n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), d_out.data_ptr(), req,
                              stream=stream)


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 10
e0.record()
for _ in range(steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
eng.set_profiling(True)
eng.kernel_times(reset=True)
for _ in range(3):
    step()
kt = eng.kernel_times(reset=True)
eng.set_profiling(False)
res = d_out.cpu().numpy().view(P.distance_result_dtype)
chk = float(np.nansum(res["min_distance"][:: max(1, n // 65536)]))
print(json.dumps({"workload": cfg, "pairs": n, "variant": variant, "pairs_per_s": n / (ms * 1e-3), "ms_per_step": ms,
                  "kernels_ms": {k: round(v / 3, 4) for k, v in kt.items() if k.endswith("_ms")}, "checksum": chk,
                  "env": {k: v for k, v in os.environ.items() if k.startswith("HFB_")}}))
