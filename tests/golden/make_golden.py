"""Generates the committed golden fixtures tests/golden/*.npz.

Each fixture is a small seeded scene (geometry + pairs + poses + request fields) together with result
records.  The records are written by the oracle and, IN THIS SCRIPT, checked field by field against the
reference itself (oracle/_ref/libhppfcl_ref.so: /root/reference compiled in place, `make -C oracle ref`) --
every field the reference defines must be bit-identical, so the vectors that travel to the GPU box are the
reference's own outputs; the oracle adds what the reference has no field for (status word, iteration / node
counters, the signed distance of a collide() without contact).  `fcl_meshes` holds the reference's own test
meshes (test/fcl_resources/env.obj, rob.obj, as used by test/distance.cpp and test/collision.cpp).
  * tests/test_golden_fixtures.py (no GPU): the oracle and the CPU emulation of the device code must
    reproduce every byte -- guards both against drift;
  * the same file, `-m gpu`: the CUDA path through the C ABI must reproduce every byte.
Run from the repo root (needs /root/reference):  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
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


def load_obj(path):
    v, f = [], []
    for line in open(path):
        t = line.split()
        if t and t[0] == "v":
            v.append([float(x) for x in t[1:4]])
        elif t and t[0] == "f":
            f.append([int(x.split("/")[0]) - 1 for x in t[1:4]])
    return np.array(v, dtype=np.float64), np.array(f, dtype=np.uint32)


def scene_fcl_meshes(sc, n=240):
    """env.obj (2 180 triangles, extents of 3 000) and rob.obj (216 triangles): mesh-mesh as in
    test/distance.cpp:89-… (random poses over the environment's extents), plus mesh-shape pairs"""
    res = "/root/reference/test/fcl_resources/"
    va, ta = load_obj(res + "env.obj")
    vb, tb = load_obj(res + "rob.obj")
    ia, _ = sc.register_bvh(va, ta)
    ib, _ = sc.register_bvh(vb, tb)
    hm = sc.register_shapes(P.make_shapes([P.BV_OBBRSS] * 2, [[0, 0, 0]] * 2, data=[ia, ib]))
    rng = np.random.default_rng(104)
    prims = W.random_primitive_shapes(rng, 32, ALL_PRIMS)
    prims["p"] *= 300
    hp = sc.register_shapes(prims)
    third = n // 3
    h1 = np.concatenate([np.full(third, hm[0]), np.full(third, hm[0]), hp[rng.integers(0, len(hp), n - 2 * third)]])
    h2 = np.concatenate([np.full(third, hm[1]), hp[rng.integers(0, len(hp), third)], np.full(n - 2 * third, hm[1])])
    tf1 = W.identity_transforms(n)
    tf2 = W.random_transforms(rng, n, (-3000, -3000, 0), (3000, 3000, 3000))
    tf1[2 * third:] = W.random_transforms(rng, n - 2 * third, (-600, -600, -600), (600, 600, 600))
    tf2[2 * third:]["T"] *= 0.2
    geo = dict(va=va, ta=ta, vb=vb, tb=tb, prims=prims)
    return geo, h1.astype(np.uint32), tf1, h2.astype(np.uint32), tf2


SCENES = {"primitives": scene_primitives, "convex": scene_convex, "mesh": scene_mesh, "fcl_meshes": scene_fcl_meshes}
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


def rebuild(sc, name, z):
    """register a fixture's geometry; handles come out in the order the scene functions above used"""
    if name == "primitives":
        sc.register_shapes(z["geo_shapes"])
    elif name == "convex":
        keys = sorted(k for k in z if k.startswith("geo_hull_"))
        cids = [sc.register_convex(z[k], None) for k in keys]
        sc.register_shapes(P.make_shapes([P.GEOM_CONVEX] * len(cids), np.zeros((len(cids), 3)), data=cids))
    else:
        ia, _ = sc.register_bvh(z["geo_va"], z["geo_ta"])
        ib, _ = sc.register_bvh(z["geo_vb"], z["geo_tb"])
        sc.register_shapes(P.make_shapes([P.BV_OBBRSS] * 2, [[0, 0, 0]] * 2, data=[ia, ib]))
        sc.register_shapes(z["geo_prims"])
    if hasattr(sc, "commit"):
        sc.commit()


def check_against_the_reference(name, out):
    """every field the reference defines, bit for bit (the convex fixture: hulls of 64 vertices go through the
    reference's hill-climb, see DESIGN.md section 4 -- flags exact, distances to 1e-6)"""
    from oracle import oracle_lib
    from tests.common import compare_hill_climb
    oracle_lib.build_ref()
    ref = oracle_lib.RefScene(P)
    if name == "convex":
        import fuzz_ref
        keys = sorted(k for k in out if k.startswith("geo_hull_"))
        cids = [ref.register_convex(out[k], fuzz_ref.hull_tris(out[k])) for k in keys]
        ref.register_shapes(P.make_shapes([P.GEOM_CONVEX] * len(cids), np.zeros((len(cids), 3)), data=cids))
    else:
        rebuild(ref, name, out)
    for rname, (kind, kw) in REQUESTS.items():
        r = run(ref, kind, kw, out["h1"], out["tf1"], out["h2"], out["tf2"])
        o = out["res_" + rname]
        if name == "convex":
            if kind == "collide":
                r, o = r.copy(), o.copy()
                r["distance"][r["num_contacts"] == 0] = 0  # the reference keeps no distance without a contact
                o["distance"][o["num_contacts"] == 0] = 0
                compare_hill_climb(r, o)
            else:
                m = ~np.isnan(r["min_distance"])
                assert np.all(np.abs(r["min_distance"][m] - o["min_distance"][m]) <= 1e-6 * np.maximum(1, np.abs(r["min_distance"][m])))
            continue
        fields = ["p1", "p2", "normal", "b1", "b2"]
        fields += ["min_distance"] if kind == "distance" else ["pos", "distance_lower_bound", "num_contacts"]
        mesh_mesh = (out["h1"] < 2) & (out["h2"] < 2) if name != "primitives" else np.zeros(len(r), dtype=bool)
        for f in fields:
            x, y = r[f], o[f]
            ok = (x == y) | (np.isnan(x) & np.isnan(y)) if x.dtype.kind == "f" else x == y
            if f == "normal" and kind == "distance":  # never written by the reference on the mesh-mesh path
                ok[mesh_mesh] = True
            assert np.all(ok), "%s/%s: %s differs from the reference" % (name, rname, f)
        if kind == "collide":
            c = r["num_contacts"] > 0
            assert np.array_equal(r["distance"][c], o["distance"][c])
    print("  %s: %s the reference build" % (name, "within the hill-climb bar of" if name == "convex" else "identical to"))


def main():
    for name, fn in SCENES.items():
        sc = make_scenes(gpu=False, emu=False)
        geo, h1, tf1, h2, tf2 = fn(sc)
        out = dict(h1=h1, tf1=tf1, h2=h2, tf2=tf2)
        out.update({"geo_" + k: v for k, v in geo.items()})
        for rname, (kind, kw) in REQUESTS.items():
            out["res_" + rname] = run(sc.b["oracle"], kind, kw, h1, tf1, h2, tf2)
        check_against_the_reference(name, out)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path) >> 10, "KiB")


if __name__ == "__main__":
    main()
