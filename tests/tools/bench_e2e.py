"""e2e-only timing of the host entry point (pinned buffers, H2D + kernels + D2H), for tuning the
chunking of hfb_batch_distance.  Usage: HFB_CHUNK=<pairs> python tests/tools/bench_e2e.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hppfcl_b200 as hf  # noqa: E402
from hppfcl_b200 import _pod as P, workloads as W  # noqa: E402

n = 1_000_000
w = W.config2_mixed_primitives(n, seed=0xFC1 + 2)
eng = hf.Engine(0)
hs = eng.register_shapes(w["shapes"])
eng.commit()
h1, h2 = hs[w["h1"]].astype(np.uint32), hs[w["h2"]].astype(np.uint32)


def pinned(a):
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).pin_memory()
    return t, t.numpy().view(a.dtype).reshape(a.shape)


keep, ph = [], []
for a in (h1, w["tf1"], h2, w["tf2"]):
    t, v = pinned(a)
    keep.append(t)
    ph.append(v)
t_out = torch.empty(n * P.distance_result_dtype.itemsize, dtype=torch.uint8).pin_memory()
out = t_out.numpy().view(P.distance_result_dtype)
req = P.DistanceRequestPOD()
for _ in range(3):
    eng.batch_distance(ph[0], ph[1], ph[2], ph[3], req, out=out)
ts = []
for _ in range(10):
    t0 = time.perf_counter()
    eng.batch_distance(ph[0], ph[1], ph[2], ph[3], req, out=out)
    ts.append(time.perf_counter() - t0)
print("chunk", os.environ.get("HFB_CHUNK", "default"), "e2e pairs/s %.4g (median %.3f ms, min %.3f ms)" % (
    n / np.median(ts), 1e3 * np.median(ts), 1e3 * min(ts)))
