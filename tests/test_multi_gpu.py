"""Multi-GPU inside the boundary (hfb_comm_*): two contexts on two GPUs in ONE process (two threads, one per rank):
communicator, geometry broadcast from rank 0, sharded batches whose records every rank receives, and -- because both
contexts live in one process -- the per-device launch configuration (the EPA kernel's shared-memory opt-in) of each."""
import threading

import numpy as np
import pytest

from tests.common import P, hf
from hppfcl_b200 import workloads as W

pytestmark = pytest.mark.gpu


def _copy_from_device(ptr, nbytes, dev):
    """bytes at a raw device pointer -> numpy (cudaMemcpy through the CUDA runtime torch has loaded)"""
    import ctypes
    import torch
    rt = ctypes.CDLL(None)
    if not hasattr(rt, "cudaMemcpy"):
        import glob, os
        cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + \
            glob.glob("/usr/local/cuda/lib64/libcudart.so*")
        rt = ctypes.CDLL(cands[0])
    out = np.empty(nbytes, dtype=np.uint8)
    torch.cuda.set_device(dev)
    rc = rt.cudaMemcpy(ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(ptr), ctypes.c_size_t(nbytes), ctypes.c_int(2))
    assert rc == 0, "cudaMemcpy failed: %d" % rc
    return out


def _two_gpus():
    import torch
    return torch.cuda.is_available() and torch.cuda.device_count() >= 2


@pytest.mark.skipif(not _two_gpus(), reason="needs two GPUs")
def test_two_ranks_in_one_process():
    import torch
    n = 60_000
    w = W.config2_mixed_primitives(2 * n, pool=2048, types=(P.GEOM_SPHERE, P.GEOM_CAPSULE, P.GEOM_BOX, P.GEOM_CYLINDER,
                                                             P.GEOM_CONE, P.GEOM_ELLIPSOID), seed=21)
    cid = hf.Engine.comm_unique_id()
    engs = [hf.Engine(0), hf.Engine(1)]
    # rank 0 alone registers the geometry
    h0 = engs[0].register_shapes(w["shapes"])
    errors, out = [], [None, None]

    def rank_main(r):
        try:
            torch.cuda.set_device(r)
            eng = engs[r]
            eng.comm_init(cid, r, 2)
            eng.geom_broadcast(0)

            def dev(a):
                return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).cuda(r)

            sl = slice(r * n, (r + 1) * n)
            d = [dev(h0[w["h1"][sl]]), dev(w["tf1"][sl]), dev(h0[w["h2"][sl]]), dev(w["tf2"][sl])]
            stream = torch.cuda.current_stream(r).cuda_stream
            res = []
            for k in range(3):  # three calls: both buffers and the reuse of the first
                ptr = eng.batch_distance_sharded_device(n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(),
                                                        stream=stream)
                eng.comm_wait(stream)
                torch.cuda.synchronize(r)
                nbytes = 2 * n * P.distance_result_dtype.itemsize
                # the gathered buffer is the context's: copy it out through a torch view of the same memory
                host = _copy_from_device(ptr, nbytes, r).view(P.distance_result_dtype)
                res.append(host)
            assert res[0].tobytes() == res[1].tobytes() == res[2].tobytes()
            out[r] = res[0]
            eng.comm_destroy()
        except Exception as ex:  # noqa: BLE001
            errors.append((r, repr(ex)))

    ts = [threading.Thread(target=rank_main, args=(r,)) for r in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=300)
    assert not errors, errors
    assert out[0] is not None and out[0].tobytes() == out[1].tobytes()  # every rank holds all the records
    # ... and they are the single-GPU records of the same pairs
    ref = hf.Engine(0)
    hr = ref.register_shapes(w["shapes"])
    ref.commit()
    want = ref.batch_distance(hr[w["h1"]], w["tf1"], hr[w["h2"]], w["tf2"])
    assert want.tobytes() == out[0].tobytes()
    assert (want["min_distance"] < 0).sum() > 100  # EPA ran on both devices
