"""Generates the committed golden fixtures tests/golden/*.npz.

Each fixture is a small seeded scene (geometry + pairs + poses + request fields) together with the
result records of the ORACLE (oracle/, the CPU restatement of hpp-fcl that tests/test_oracle_golden.py
pins against the reference's own known-answer tests).  The reference itself cannot be built or
imported in this image (Eigen/Boost are absent), so these are the vectors that travel:
  * tests/test_golden_fixtures.py (no GPU): the oracle and the CPU emulation of the device code must
    reproduce every byte -- guards both against drift;
  * the same file, `-m gpu`: the CUDA path through the C ABI must reproduce every byte.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.common import P, make_scenes  # noqa: E402
from hppfcl_b200 import workloads as W  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ALL_PRIMS = (P.GEOM_SPHERE, P.GEOM_CAPSULE, P.GEOM_BOX, P.GEOM_CYLINDER, P.GEOM_CONE, P.GEOM_ELLIPSOID)


def scene_primitives(sc, n=1500):
    w = W.config2_mixed_primitives(n, pool=256, types=ALL_PRIMS, seed=101)
    hs = sc.register_shapes(w["shapes"])
    return dict(shapes=w["shapes"]), hs[w["h1"]], w["tf1"], hs[w["h2"]], w["tf2"]


def scene_convex(sc, n=600):
    w = W.config3_convex_pairs(n, pool=12, nv=64, seed=102)
    small = [W.icosahedron_from_ellipsoid(r) for r in ((0.2, 0.3, 0.25), (0.5, 0.1, 0.3))]
    hulls = [np.asarray(p) for p, _ in w["hulls"]] + [np.asarray(p) for p, _ in small]
    # registered without neighbours in the oracle: the kernels' exhaustive argmax == its linear scan
    cids = [sc.register_convex(p, None) for p in hulls]
    hs = sc.register_shapes(P.make_shapes([P.GEOM_CONVEX] * len(cids), np.zeros((len(cids), 3)), data=cids))
    rng = np.random.default_rng(7)
    h1, h2 = hs[rng.integers(0, len(hs), n)], hs[rng.integers(0, len(hs), n)]
    geo = {"hull_%02d" % k: h for k, h in enumerate(hulls)}
    return geo, h1, w["tf1"], h2, w["tf2"]


def scene_mesh(sc, n=400):
    rng = np.random.default_rng(103)
    va, ta = W.sphere_mesh(1.0, 14, 8, noise=0.03, rng=rng)
    vb, tb = W.sphere_mesh(0.6, 10, 6, noise=0.05, rng=rng)
    ia, _ = sc.register_bvh(va, ta)
    ib, _ = sc.register_bvh(vb, tb)
    hm = sc.register_shapes(P.make_shapes([P.BV_OBBRSS] * 2, [[0, 0, 0]] * 2, data=[ia, ib]))
    prims = W.random_primitive_shapes(rng, 32, ALL_PRIMS)
    prims["p"] *= 0.3
    hp = sc.register_shapes(prims)
    allh = np.concatenate([hm, hm, hp])  # mesh-mesh, mesh-shape and shape-mesh pairs
    h1 = hm[rng.integers(0, 2, n)]
    h2 = allh[rng.integers(0, len(allh), n)]
    swap = rng.random(n) < 0.25
    h1s, h2s = np.where(swap, h2, h1), np.where(swap, h1, h2)
    tf1 = W.random_transforms(rng, n, (-0.2, -0.2, -0.2), (0.2, 0.2, 0.2))
    tf2 = W.random_transforms(rng, n, (-1.8, -1.8, -1.8), (1.8, 1.8, 1.8))
    geo = dict(va=va, ta=ta, vb=vb, tb=tb, prims=prims)
    return geo, h1s.astype(np.uint32), tf1, h2s.astype(np.uint32), tf2


SCENES = {"primitives": scene_primitives, "convex": scene_convex, "mesh": scene_mesh}
REQUESTS = {
    "distance_default": ("distance", dict()),
    "distance_nesterov": ("distance", dict(gjk_variant=P.NesterovAcceleration)),
    "collide_default": ("collide", dict()),
    "collide_margin": ("collide", dict(security_margin=0.05, num_max_contacts=2)),
}


def run(backend, kind, kw, h1, tf1, h2, tf2):
    if kind == "distance":
        return backend.batch_distance(h1, tf1, h2, tf2, P.DistanceRequestPOD(**kw))
    return backend.batch_collide(h1, tf1, h2, tf2, P.CollisionRequestPOD(**kw))


def main():
    for name, fn in SCENES.items():
        sc = make_scenes(gpu=False, emu=False)
        geo, h1, tf1, h2, tf2 = fn(sc)
        out = dict(h1=h1, tf1=tf1, h2=h2, tf2=tf2)
        out.update({"geo_" + k: v for k, v in geo.items()})
        for rname, (kind, kw) in REQUESTS.items():
            out["res_" + rname] = run(sc.b["oracle"], kind, kw, h1, tf1, h2, tf2)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path) >> 10, "KiB")


if __name__ == "__main__":
    main()
