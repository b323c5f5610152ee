"""collide()'s distance lower bound against distance(): the reference's suite
(/root/reference/test/distance_lower_bound.cpp:63-267).  For random poses: collide() with a default request and
collide() without contact / lower-bound computation agree on the flag, distance() <= 0 exactly on the colliding
pairs (shape-shape and mesh-shape), and on the free pairs the lower bound never exceeds the distance.  The reference
loads env.obj / rob.obj; the meshes here are the tessellated spheres of the benchmark workloads (with radial noise),
scaled to the same proportions.

Checked on the oracle; the host build of the device code and, where oracle/_ref exists, the reference build must
return the same bits.  CPU only.
"""
import numpy as np

from tests.common import P, compare_distance, make_scenes, ref_agrees
from hppfcl_b200 import workloads as W


MESH_COLLIDE = ("p1", "p2", "normal", "pos", "distance_lower_bound", "b1", "b2", "num_contacts")
MESH_DISTANCE = ("min_distance", "p1", "p2", "b1", "b2")


def both(sc, fn, *a, fields=None):
    ro = getattr(sc.b["oracle"], fn)(*a)
    re = getattr(sc.b["emu"], fn)(*a)
    compare_distance(ro, re, what=fn)
    ref_agrees(sc, fn, ro, a, fn, fields=fields)
    return ro


def lower_bound_suite(sc, h1, h2, t1, check_distance_flag=True, mesh=False):
    n = len(t1)
    h1 = np.full(n, h1, dtype=np.uint32)
    h2 = np.full(n, h2, dtype=np.uint32)
    t2 = W.identity_transforms(n)
    fc, fd = (MESH_COLLIDE, MESH_DISTANCE) if mesh else (None, None)
    c1 = both(sc, "batch_collide", h1, t1, h2, t2, P.CollisionRequestPOD(), fields=fc)  # testDistanceLowerBound :63-79
    c3 = both(sc, "batch_collide", h1, t1, h2, t2, P.CollisionRequestPOD(enable_contact=0), fields=fc)  # testCollide :81-94
    d = both(sc, "batch_distance", h1, t1, h2, t2, P.DistanceRequestPOD(), fields=fd)  # testDistance :96-113
    col1, col3, col2 = c1["num_contacts"] > 0, c3["num_contacts"] > 0, d["min_distance"] <= 0
    assert np.array_equal(col1, col3)
    if check_distance_flag:
        assert np.array_equal(col1, col2)
    free = ~col1
    assert 0 < free.sum() < n
    assert np.all(c1["distance_lower_bound"][free] <= d["min_distance"][free])
    return col1


def test_box_sphere_and_sphere_sphere():  # :160-232
    sc = make_scenes(ref=True)
    sph, box, sph2 = (int(h) for h in sc.register_shapes(P.make_shapes(
        [P.GEOM_SPHERE, P.GEOM_BOX, P.GEOM_SPHERE], [[0.5, 0, 0], [0.5, 0.5, 0.5], [1.0, 0, 0]])))
    sc.commit()
    rng = np.random.default_rng(21)
    t1 = W.random_transforms(rng, 1001, (-2, -2, -2), (2, 2, 2))
    ident = W.identity_transforms(1)
    t1["R"][0], t1["T"][0] = ident["R"][0], ident["T"][0]  # the identity pose first, as there
    lower_bound_suite(sc, sph, box, t1)
    lower_bound_suite(sc, sph, sph2, t1)


def test_mesh_mesh_and_box_mesh():  # :115-158, :234-267
    sc = make_scenes(ref=True)
    rng = np.random.default_rng(22)
    v1, f1 = W.sphere_mesh(1500.0, 24, 12, noise=0.05, rng=rng)  # "environment"
    v2, f2 = W.sphere_mesh(400.0, 16, 8, noise=0.05, rng=rng)    # "robot"
    b1, _ = sc.register_bvh(v1, f1)
    b2, _ = sc.register_bvh(v2, f2)
    m1, m2 = (int(h) for h in sc.register_shapes(P.make_shapes([P.BV_OBBRSS] * 2, np.zeros((2, 3)), data=[b1, b2])))
    box = int(sc.register_shapes(P.make_shapes([P.GEOM_BOX], [[250, 100, 75]]))[0])
    sc.commit()
    t1 = W.random_transforms(rng, 100, (-3000, -3000, 0), (3000, 3000, 3000))
    # mesh-mesh: the reference compares the two collide() flags only (distance() between meshes stops at 0)
    lower_bound_suite(sc, m1, m2, t1, check_distance_flag=False, mesh=True)
    lower_bound_suite(sc, m1, box, t1, mesh=True)
