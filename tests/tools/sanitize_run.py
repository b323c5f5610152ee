"""Small batches of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool memcheck python tests/tools/sanitize_run.py
Checks results against the oracle as well, so a sanitizer-clean run is also a correct one."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import hppfcl_b200 as hf  # noqa: E402
from hppfcl_b200 import _pod as P, workloads as W  # noqa: E402
from oracle import oracle_lib  # noqa: E402

ALL = (P.GEOM_SPHERE, P.GEOM_CAPSULE, P.GEOM_BOX, P.GEOM_CYLINDER, P.GEOM_CONE, P.GEOM_ELLIPSOID)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
eng, orc = hf.Engine(0), oracle_lib.OracleScene(P)
w = W.config2_mixed_primitives(n, pool=256, types=ALL, seed=5)
hp = eng.register_shapes(w["shapes"])
orc.register_shapes(w["shapes"])
c3 = W.config3_convex_pairs(n, pool=8, nv=64, seed=6)
cids = []
for pts, _ in c3["hulls"]:
    cid = eng.register_convex(pts)
    assert cid == orc.register_convex(pts, None)
    cids.append(cid)
recs = P.make_shapes([P.GEOM_CONVEX] * len(cids), np.zeros((len(cids), 3)), data=cids)
hc = eng.register_shapes(recs)
orc.register_shapes(recs)
rng = np.random.default_rng(3)
verts, tris = W.sphere_mesh(1.0, 10, 6, noise=0.02, rng=rng)
bid = eng.register_bvh_obbrss(None, verts, tris)
obid, _ = orc.register_bvh(verts, tris)
rec = P.make_shapes([P.BV_OBBRSS], [[0, 0, 0]], data=[bid])
hm = eng.register_shapes(rec)
orc.register_shapes(P.make_shapes([P.BV_OBBRSS], [[0, 0, 0]], data=[obid]))
eng.commit()
allh = np.concatenate([hp, hc, hm])
h1 = allh[rng.integers(0, len(allh), n)]
h2 = allh[rng.integers(0, len(allh), n)]
for req in (P.DistanceRequestPOD(), P.DistanceRequestPOD(gjk_variant=P.NesterovAcceleration, epa_tolerance=1e-10)):
    got = eng.batch_distance(h1, w["tf1"], h2, w["tf2"], req)
    ref = orc.batch_distance(h1, w["tf1"], h2, w["tf2"], req, nthreads=0)
    assert np.array_equal(got["status"], ref["status"]) and np.array_equal(got["iterations"], ref["iterations"])
    m = ~np.isnan(ref["min_distance"])
    assert np.array_equal(got["min_distance"][m], ref["min_distance"][m])
cg = eng.batch_collide(h1, w["tf1"], h2, w["tf2"])
cr = orc.batch_collide(h1, w["tf1"], h2, w["tf2"], nthreads=0)
assert np.array_equal(cg["num_contacts"], cr["num_contacts"])
ids = rng.integers(0, len(cids), 2000).astype(np.uint32)
dirs = rng.normal(size=(2000, 3))
gi, _ = eng.batch_convex_support(ids, dirs)
oi, _ = orc.batch_convex_support(ids, dirs)
assert np.array_equal(gi, oi)
print("sanitize_run ok: %d pairs (%d epa), knobs GC=%s GE=%s STAGE=%s REFILL=%s" % (
    n, eng.stats()["epa_pairs"], os.environ.get("HFB_GC"), os.environ.get("HFB_GE"), os.environ.get("HFB_STAGE"),
    os.environ.get("HFB_REFILL")))
