"""collide() known answers of the primitive pairs, transcribed from the reference's own test
(/root/reference/test/geometric_shapes.cpp:238-1031 -- SURVEY 8c): collision flag with and without contact
computation, and where the reference gives them the normal (to 1e-9, 1e-8 or tol_gjk = 0.01 as there) -- every case
in the frame of the test and under a common rigid transform (random in the reference, fixed here).

Checked on the oracle's records; the host build of the device code (tests/emu) and, where oracle/_ref exists, the
reference build must return the same bits.  CPU only:
tests/test_gpu_parity.py compares the CUDA kernels with the oracle on all of these pair types.
"""
import numpy as np
import pytest

from tests.common import P
from tests.test_plane_known_answers import World, tf, compose, rot, apply, GLOBAL
from hppfcl_b200 import workloads as W

TOL_GJK = 0.01  # geometric_shapes.cpp:55


def test_collide_sphere_sphere():  # :238-329
    w = World()
    s1, s2 = w.shape("sphere", 20), w.shape("sphere", 10)
    w.check(s1, tf(), s2, tf((40, 0, 0)), False)
    w.check(s1, tf(), s2, tf((30, 0, 0)), True, None, None, (1, 0, 0), both_frames=False)
    w.check(s1, tf(), s2, tf((30.01, 0, 0)), False)
    w.check(s1, tf(), s2, tf((29.9, 0, 0)), True, None, None, (1, 0, 0))
    # coinciding centres: the normal is (1, 0, 0), also under the transform (:296-309)
    w.check(s1, tf(), s2, tf(), True, None, None, (1, 0, 0), both_frames=False)
    w.check(s1, w.g, s2, w.g, True, None, None, (1, 0, 0), both_frames=False)
    w.check(s1, tf(), s2, tf((-29.9, 0, 0)), True, None, None, (-1, 0, 0))
    w.check(s1, tf(), s2, tf((-30.0, 0, 0)), True, None, None, (-1, 0, 0), both_frames=False)
    w.check(s1, tf(), s2, tf((-30.01, 0, 0)), False)


def test_collide_box_box():  # :385-453
    w = World()
    s1, s2 = w.shape("box", 20, 40, 50), w.shape("box", 10, 10, 10)
    w.check(s1, tf(), s2, tf(), True)
    w.check(s1, tf(), s2, tf((15, 0, 0)), True, None, None, (1, 0, 0), tol=1e-8, both_frames=False)
    w.check(s1, tf(), s2, tf((15.01, 0, 0)), False)
    half = 3.140 / 6 / 2  # AngleAxis(3.140 / 6, UnitZ)
    w.check(s1, tf(), s2, tf(quat=(np.cos(half), 0, 0, np.sin(half))), True)


def test_box_box_contact_points():  # testBoxBoxContactPoints :334-383, 100 random rotations (:447-452)
    w = World()
    s1, s2 = w.shape("box", 100, 100, 100), w.shape("box", 10, 20, 30)
    rng = np.random.default_rng(5)
    R = W.random_rotations(rng, 100)
    corners = np.array([[x, y, z] for x in (1, -1) for y in (1, -1) for z in (1, -1)], dtype=float) * [5, 10, 15]
    req = P.DistanceRequestPOD(gjk_tolerance=1e-5, epa_tolerance=1e-5)
    w.sc.commit()
    t1 = P.make_transforms(np.eye(3)[None].repeat(100, 0), np.array([[0, 0, -50.0]]).repeat(100, 0))
    t2 = P.make_transforms(R, np.zeros((100, 3)))
    h1, h2 = [s1] * 100, [s2] * 100
    ro = w.sc.b["oracle"].batch_distance(h1, t1, h2, t2, req)
    re = w.sc.b["emu"].batch_distance(h1, t1, h2, t2, req)
    from tests.common import compare_distance
    compare_distance(ro, re, what="box-box contact points")
    for i in range(100):
        r = ro[i]
        assert r["min_distance"] <= 0
        v = corners @ R[i].T  # world vertices of the small box
        low = v[np.argmin(v[:, 2])]
        assert np.allclose(r["normal"], [0, 0, 1], rtol=0, atol=1e-6 * 1.0)  # isApprox(…, 1e-6)
        point = 0.5 * (r["p1"] + r["p2"])
        assert np.linalg.norm(low[:2] - point[:2]) <= 1e-6 * max(1e-300, min(np.linalg.norm(low[:2]), np.linalg.norm(point[:2]))) + 1e-12
        assert low[2] <= point[2] < 0


def test_collide_sphere_box():  # :455-505
    w = World()
    s1, s2 = w.shape("sphere", 20), w.shape("box", 5, 5, 5)
    w.check(s1, tf(), s2, tf(), True)
    w.check(s1, tf(), s2, tf((22.50001, 0, 0)), False, both_frames=False)
    w.check(s1, tf(), s2, tf((22.501, 0, 0)), False)
    w.check(s1, tf(), s2, tf((22.4, 0, 0)), True, None, None, (1, 0, 0), tol=TOL_GJK)


def test_distance_sphere_box():  # :507-528, to machine epsilon (BOOST_CHECK_CLOSE in percent, isApprox relative)
    w = World()
    s, b = w.shape("sphere", 1), w.shape("box", 10, 2, 10)
    w.sc.commit()
    req = P.DistanceRequestPOD()
    r = w.sc.b["oracle"].batch_distance([s], tf(), [b], tf((0, 5, 3)), req)[0]
    e = w.sc.b["emu"].batch_distance([s], tf(), [b], tf((0, 5, 3)), req)[0]
    assert r.tobytes() == e.tobytes()
    eps = np.finfo(float).eps
    assert abs(r["min_distance"] - 3.0) <= eps / 100 * 3.0
    assert np.linalg.norm(r["p1"] - [0, 1, 0]) <= eps
    assert np.linalg.norm(r["p2"] - [0, 4, 0]) <= eps * 4
    assert np.linalg.norm(r["normal"] - [0, 1, 0]) <= eps


def test_collide_sphere_capsule():  # :530-593
    w = World()
    s1, s2 = w.shape("sphere", 20), w.shape("capsule", 5, 10)
    w.check(s1, tf(), s2, tf(), True)
    w.check(s1, tf(), s2, tf((24.9, 0, 0)), True, None, None, (1, 0, 0))
    w.check(s1, tf(), s2, tf((25, 0, 0)), True, None, None, (1, 0, 0), both_frames=False)
    w.check(s1, w.g, s2, compose(w.g, tf((24.999999, 0, 0))), True, None, None, rot(w.g) @ np.array([1.0, 0, 0]),
            both_frames=False)
    w.check(s1, tf(), s2, tf((25.1, 0, 0)), False)


def test_collide_cylinder_cylinder():  # :595-656
    w = World()
    s1, s2 = w.shape("cylinder", 5, 15), w.shape("cylinder", 5, 15)
    w.check(s1, tf(), s2, tf(), True)
    w.check(s1, tf(), s2, tf((9.9, 0, 0)), True, None, None, (1, 0, 0), tol=TOL_GJK)
    w.check(s1, tf(), s2, tf((0, 9.9, 0)), True, None, None, (0, 1, 0), tol=TOL_GJK, both_frames=False)
    w.check(s1, tf(), s2, tf((10.01, 0, 0)), False)


def test_collide_cone_cone():  # :658-734
    w = World()
    s1, s2 = w.shape("cone", 5, 10), w.shape("cone", 5, 10)
    w.check(s1, tf(), s2, tf(), True)
    n = np.array([2 * (5 + 5), 0, 5 + 5], dtype=float)
    n /= np.linalg.norm(n)
    w.check(s1, tf(), s2, tf((9.9, 0, 0.00001)), True, None, None, n, tol=TOL_GJK)
    w.check(s1, tf(), s2, tf((10.1, 0, 0)), False, both_frames=False)
    w.check(s1, tf(), s2, tf((10.001, 0, 0)), False)
    w.check(s1, tf(), s2, tf((0, 0, 9.9)), True, None, None, (0, 0, 1))


def test_collide_cone_cylinder():  # :736-842
    w = World()
    s1, s2 = w.shape("cylinder", 5, 10), w.shape("cone", 5, 10)
    w.check(s1, tf(), s2, tf(), True)
    n = np.array([2 * (5 + 5), 0, -(5 + 5)], dtype=float)
    n /= np.linalg.norm(n)
    w.check(s1, tf(), s2, tf((9.9, 0, 0)), True, None, None, n, tol=TOL_GJK)
    w.check(s1, tf(), s2, tf((9.9, 0, 0.1)), True, None, None, (1, 0, 0), tol=TOL_GJK)
    w.check(s1, tf(), s2, tf((10.01, 0, 0)), False)
    w.check(s1, tf(), s2, tf((10, 0, 0)), True)
    w.check(s1, tf(), s2, tf((0, 0, 9.9)), True, None, None, (0, 0, 1), both_frames=False)
    # (under the common transform: test_flat_faces_overlap_without_contact_computation below)
    w.check(s1, tf(), s2, tf((0, 0, 10.01)), False)
    w.check(s1, tf(), s2, tf((0, 0, 10)), True, None, None, (0, 0, 1), tol=TOL_GJK, both_frames=False)
    w.check(s1, w.g, s2, compose(w.g, tf((0, 0, 10.1))), False, both_frames=False)


def test_flat_faces_overlap_without_contact_computation():
    """What transcribing :820-829 found.  The common transform of this file is a quaternion written with 12 digits:
    its rotation matrix is orthonormal to 8e-14, not to 1e-16.  Cylinder top face against cone base, 0.1 deep, both
    under that transform: GJK's third simplex holds the origin up to |ray| = 1.4e-12 and stops with `Collision`
    (gjk.cpp:228-243, |ray| < tolerance, distance = |ray|).  With enable_contact = false no penetration is computed
    and the solver reports distance = |ray| (narrowphase.h:638-656), which collide() compares with
    collision_distance_threshold = 1e-12 (shape_shape_func.h:148-152): 1.4e-12 > 1e-12, *no collision* -- for shapes
    0.1 inside each other.  With the quaternion normalised |ray| is 8e-15 and the pair collides, which is why the
    reference's own run of this case (Eigen's random unit quaternion) passes; with enable_contact = true EPA runs and
    the contact is reported either way.  The reference compiled in place (oracle/_ref) returns the same bits
    (World.collide asserts it), the oracle restates it, and so do the kernels: a drop-in answers like the library.
    test_flat_faces_under_common_rotations below sweeps the class."""
    w = World()
    s1, s2 = w.shape("cylinder", 5, 10), w.shape("cone", 5, 10)
    t1, t2 = w.g, compose(w.g, tf((0, 0, 9.9)))
    r = w.collide(s1, t1, s2, t2, enable_contact=0)
    assert r["num_contacts"] == 0 and 1e-12 < r["distance_lower_bound"] < 1e-6
    r = w.collide(s1, t1, s2, t2, enable_contact=1)
    assert r["num_contacts"] == 1 and abs(r["distance"] + 0.1) < 1e-9
    assert np.linalg.norm(r["normal"] - rot(w.g) @ np.array([0, 0, 1.0])) < 1e-9
    q = np.array(GLOBAL[1])
    g = tf(GLOBAL[0], quat=tuple(q / np.linalg.norm(q)))  # the same pose, orthonormal to 1e-16
    r = w.collide(s1, g, s2, compose(g, tf((0, 0, 9.9))), enable_contact=0)
    assert r["num_contacts"] == 1 and r["distance_lower_bound"] < 1e-12


@pytest.mark.parametrize("tri,T,normal", [
    (((20, 0, 0), (-20, 0, 0), (0, 20, 0)), (0, 0, 0.001), (0, 0, 1)),
    (((20, 0, 0), (-20, 0, 0), (0, 20, 0)), (0, 0, -0.001), (0, 0, -1)),
    (((30, 0, 0), (9.9, -20, 0), (9.9, 20, 0)), (0, 0, 0.001), (9.9, 0, 0.001)),
    (((30, 0, 0), (9.9, -20, 0), (9.9, 20, 0)), (0, 0, -0.001), (9.9, 0, -0.001)),
    (((30, 0, 0), (-20, 0, 0), (0, 0, 20)), (0, 0.001, 0), (0, 1, 0)),
    (((30, 0, 0), (-20, 0, 0), (0, 0, 20)), (0, -0.001, 0), (0, -1, 0)),
    (((0, 30, 0), (0, -10, 0), (0, 0, 20)), (0.001, 0, 0), (1, 0, 0)),
    (((0, 30, 0), (0, -10, 0), (0, 0, 20)), (-0.001, 0, 0), (-1, 0, 0)),
])
def test_collide_sphere_triangle_touching(tri, T, normal):  # :844-957
    w = World()
    s, t = w.shape("sphere", 10), w.triangle(*tri)
    n = np.array(normal, dtype=float)
    w.check(s, tf(), t, tf(T), True, None, None, n / np.linalg.norm(n))


@pytest.mark.parametrize("tri,T", [
    (((20, 0, 0), (-20, 0, 0), (0, 20, 0)), (0, 0, 10.1)),
    (((20, 0, 0), (-20, 0, 0), (0, 20, 0)), (0, 0, -10.1)),
    (((20, 0, 0), (-20, 0, 0), (0, 0, 20)), (0, 10.1, 0)),
    (((20, 0, 0), (-20, 0, 0), (0, 0, 20)), (0, -10.1, 0)),
    (((0, 20, 0), (0, -20, 0), (0, 0, 20)), (10.1, 0, 0)),
    (((0, 20, 0), (0, -20, 0), (0, 0, 20)), (-10.1, 0, 0)),
])
def test_collide_sphere_triangle_apart(tri, T):  # :959-1030
    w = World()
    s, t = w.shape("sphere", 10), w.triangle(*tri)
    w.check(s, tf(), t, tf(T), False)


def test_box_and_its_hull_agree():
    """test/convex.cpp:89-174 (compare_convex_box): a box against itself and the convex hull of its corners against
    itself, 1 002 poses -- the same collision flag and number of contacts, the same distance lower bound on the free
    pairs (1e-4 percent); contact position, normal and the distance() results are BOOST_WARN there ("there are still
    some bugs") and are reported here as a fraction."""
    sc_w = World()
    sc = sc_w.sc
    box = sc_w.shape("box", 2, 2, 2)
    corners = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], dtype=float)
    cid = sc.register_convex(corners, None)
    hull = int(sc.register_shapes(P.make_shapes([P.GEOM_CONVEX], [[0, 0, 0]], data=[cid]))[0])
    sc.commit()
    rng = np.random.default_rng(17)
    n = 1002
    t1 = W.identity_transforms(n)
    t2 = W.random_transforms(rng, n, (0, 0, 0), (10, 10, 10))
    ident = W.identity_transforms(2)
    t2["R"][:2] = ident["R"]
    t2["T"][0], t2["T"][1] = (3, 0, 0), (0, 0, 0)
    from tests.common import compare_distance
    res = {}
    for name, h in (("box", box), ("hull", hull)):
        hh = np.full(n, h, dtype=np.uint32)
        co = sc.b["oracle"].batch_collide(hh, t1, hh, t2, P.CollisionRequestPOD())
        compare_distance(co, sc.b["emu"].batch_collide(hh, t1, hh, t2, P.CollisionRequestPOD()), what=name)
        di = sc.b["oracle"].batch_distance(hh, t1, hh, t2, P.DistanceRequestPOD())
        compare_distance(di, sc.b["emu"].batch_distance(hh, t1, hh, t2, P.DistanceRequestPOD()), what=name)
        res[name] = (co, di)
    (ca, da), (cb, db) = res["box"], res["hull"]
    assert np.array_equal(ca["num_contacts"], cb["num_contacts"])
    free = ca["num_contacts"] == 0
    assert 0 < free.sum() < n
    a, b = ca["distance_lower_bound"][free], cb["distance_lower_bound"][free]
    assert np.all(np.abs(a - b) <= 1e-4 / 100 * np.minimum(np.abs(a), np.abs(b)))
    # the warn-level items: how often they hold (they need not: flat faces have many closest pairs)
    hit = ~free
    same_pos = np.sum((ca["pos"][hit] - cb["pos"][hit]) ** 2, axis=1) < 1e-4
    same_d = np.abs(da["min_distance"] - db["min_distance"]) <= 1e-4 / 100 * np.abs(da["min_distance"]) + 1e-12
    assert same_d.mean() > 0.99 and same_pos.mean() > 0.5


def test_flat_faces_under_common_rotations():
    """the class of test_flat_faces_overlap_without_contact_computation, swept: flat faces overlapping by 0.1 /
    touching / 0.1 apart along the common axis, both shapes under the same rotation and translation -- where GJK's
    last simplex holds the origin up to rounding and the answers sit next to the thresholds.  Half of the rotations
    come from unit quaternions, half from quaternions rounded to 12 digits (a pose read from a file).  Oracle, host
    build of the device code and reference build must agree bit for bit with and without contact computation.  What
    the library then answers without contact computation for the 0.1-deep pairs: with unit quaternions always
    "colliding"; with the rounded ones "free" for most cylinder-on-cone poses (measured here: 86 % of them)."""
    from tests.common import compare_distance, ref_agrees
    w = World()
    sc = w.sc
    hs = {k: w.shape(*v) for k, v in dict(cyl=("cylinder", 5, 10), cone=("cone", 5, 10), box=("box", 10, 10, 10),
                                          cyl2=("cylinder", 3, 4)).items()}
    sc.commit()
    pairs = [("cyl", "cone", 10.0), ("cyl", "cyl", 10.0), ("box", "box", 10.0), ("cyl", "box", 10.0), ("cyl2", "cone", 7.0),
             ("box", "cone", 10.0)]
    rng = np.random.default_rng(31)
    n = 600
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1)[:, None]
    rounded = np.arange(n) >= n // 2
    q[rounded] = np.round(q[rounded], 12)
    R = W.quat_to_rot(q[:, 0], q[:, 1], q[:, 2], q[:, 3])
    T = rng.uniform(-5, 5, (n, 3))
    missed = {}
    for a, b, reach in pairs:
        for gap in (-0.1, 0.0, 0.1):
            t1 = P.make_transforms(R, T)
            t2 = P.make_transforms(R, T + R[:, :, 2] * (reach + gap))  # shape 2 moved along the common z axis
            h1 = np.full(n, hs[a], dtype=np.uint32)
            h2 = np.full(n, hs[b], dtype=np.uint32)
            out = []
            for ec in (0, 1):
                req = P.CollisionRequestPOD(enable_contact=ec)
                ro = sc.b["oracle"].batch_collide(h1, t1, h2, t2, req)
                compare_distance(ro, sc.b["emu"].batch_collide(h1, t1, h2, t2, req), what="%s-%s %g" % (a, b, gap))
                ref_agrees(sc, "batch_collide", ro, (h1, t1, h2, t2, req), "%s-%s %g" % (a, b, gap))
                out.append(ro)
            ro = sc.b["oracle"].batch_distance(h1, t1, h2, t2, P.DistanceRequestPOD())
            compare_distance(ro, sc.b["emu"].batch_distance(h1, t1, h2, t2, P.DistanceRequestPOD()), what="distance")
            ref_agrees(sc, "batch_distance", ro, (h1, t1, h2, t2, P.DistanceRequestPOD()), "distance")
            if gap < 0:
                assert np.all(out[1]["num_contacts"] == 1) and np.all(np.abs(out[1]["distance"] + 0.1) < 1e-6)
                free = out[0]["num_contacts"] == 0
                assert not free[~rounded].any(), (a, b)
                missed[(a, b)] = float(free[rounded].mean())
            elif gap > 0:
                assert np.all(out[0]["num_contacts"] == 0) and np.all(out[1]["num_contacts"] == 0)
    assert missed[("cyl", "cone")] > 0.5, missed
