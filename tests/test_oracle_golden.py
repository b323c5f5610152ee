"""Pins the oracle (CPU restatement of the hpp-fcl hot path) against the reference's own
known-answer tests -- literal numbers transcribed from /root/reference/test/*.cpp, each case
citing its source.  The same cases are also run through the CPU emulation of the device code
(tests/emu), which must agree with the oracle bit for bit.
"""
import numpy as np
import pytest

from tests.common import P, compare_distance, make_scenes, ref_agrees
from hppfcl_b200 import workloads as W

SQ2 = np.sqrt(2.0)


def tf(T=(0, 0, 0), quat=None, R=None):
    if quat is not None:
        R = W.quat_to_rot(*quat)
    if R is None:
        R = np.eye(3)
    return P.make_transforms(np.asarray(R)[None], np.asarray(T, dtype=float)[None])


def compose(a, b):
    """Transform3f operator* (transform.h:186-188) on POD transforms."""
    Ra = a["R"][0].reshape(3, 3).T
    Rb = b["R"][0].reshape(3, 3).T
    return P.make_transforms((Ra @ Rb)[None], (Ra @ b["T"][0] + a["T"][0])[None])


def apply(t, v):
    return t["R"][0].reshape(3, 3).T @ np.asarray(v, dtype=float) + t["T"][0]


def box(x, y, z):
    return P.make_shapes([P.GEOM_BOX], [[x / 2, y / 2, z / 2]])


def sphere(r, ssr=0.0):
    return P.make_shapes([P.GEOM_SPHERE], [[r, 0, 0]], ssr=ssr)


def capsule(r, lz):
    return P.make_shapes([P.GEOM_CAPSULE], [[r, lz / 2, 0]])


def cylinder(r, lz):
    return P.make_shapes([P.GEOM_CYLINDER], [[r, lz / 2, 0]])


def cone(r, lz):
    return P.make_shapes([P.GEOM_CONE], [[r, lz / 2, 0]])


class Q:
    """single-pair distance()/collide() through oracle AND emulated device code.  With
    HFB_GOLDEN_BACKEND=gpu (tests/test_gpu_parity.py::test_reference_known_answers_on_gpu re-runs this
    file that way) the second backend is the CUDA path through the C ABI, and the record the literal
    numbers are checked on is the GPU's."""

    def __init__(self):
        import os
        self.gpu = os.environ.get("HFB_GOLDEN_BACKEND") == "gpu"
        self.other = "gpu" if self.gpu else "emu"
        # (on the CPU the reference build, where it exists, is a third backend that must return the same bits)
        self.sc = make_scenes(gpu=self.gpu, emu=not self.gpu, ref=not self.gpu)
        self.tris = set()  # TriangleP operands are not in the reference's public dispatch tables

    def add(self, rec):
        return int(self.sc.register_shapes(rec)[0])

    def add_tri(self, a, b, c):
        cid = self.sc.register_convex(np.array([a, b, c], dtype=float), None)
        h = int(self.sc.register_shapes(P.make_shapes([P.GEOM_TRIANGLE], [[0, 0, 0]], data=[cid]))[0])
        self.tris.add(h)
        return h

    def _ref(self, fn, ro, h1, t1, h2, t2, req):
        if "ref" in self.sc.b and h1 not in self.tris and h2 not in self.tris:
            ref_agrees(self.sc, fn, ro, ([h1], t1, [h2], t2, req), "golden")

    def distance(self, h1, t1, h2, t2, **kw):
        req = P.DistanceRequestPOD(**kw)
        self.sc.commit()
        ro = self.sc.b["oracle"].batch_distance([h1], t1, [h2], t2, req)
        re = self.sc.b[self.other].batch_distance([h1], t1, [h2], t2, req)
        compare_distance(ro, re, what="golden")
        self._ref("batch_distance", ro, h1, t1, h2, t2, req)
        return re[0] if self.gpu else ro[0]

    def collide(self, h1, t1, h2, t2, **kw):
        req = P.CollisionRequestPOD(**kw)
        self.sc.commit()
        ro = self.sc.b["oracle"].batch_collide([h1], t1, [h2], t2, req)
        re = self.sc.b[self.other].batch_collide([h1], t1, [h2], t2, req)
        compare_distance(ro, re, what="golden")
        self._ref("batch_collide", ro, h1, t1, h2, t2, req)
        return re[0] if self.gpu else ro[0]


def close_pct(a, b, pct):  # BOOST_CHECK_CLOSE semantics (percent)
    return abs(a - b) <= pct / 100.0 * min(abs(a), abs(b)) + 0.0


# ------------------------------------------------------------------ box_box_distance.cpp
def test_distance_box_box_1():  # test/box_box_distance.cpp:62-101
    q = Q()
    r = q.distance(q.add(box(6, 10, 2)), tf(), q.add(box(2, 2, 2)), tf((25, 20, 5)))
    dx, dy, dz = 25 - 3 - 1, 20 - 5 - 1, 5 - 1 - 1
    assert close_pct(r["min_distance"], np.sqrt(dx * dx + dy * dy + dz * dz), 1e-4)
    assert np.allclose(r["p1"], [3, 5, 1], rtol=1e-8)
    assert np.allclose(r["p2"], [24, 19, 4], rtol=1e-8)


def test_distance_box_box_2():  # :103-142
    q = Q()
    s8 = np.sin(np.pi / 8) / np.sqrt(3)
    r = q.distance(q.add(box(6, 10, 2)), tf(), q.add(box(2, 2, 2)),
                   tf((0, 0, 10), quat=(np.cos(np.pi / 8), s8, s8, s8)))
    assert close_pct(r["min_distance"], -1.62123444 + 10 - 1, 1e-4)
    assert close_pct(r["p1"][0], 0.60947571, 1e-4) and close_pct(r["p1"][1], 0.01175873, 1e-4)
    assert close_pct(r["p1"][2], 1, 1e-6)
    assert close_pct(r["p2"][0], 0.60947571, 1e-4) and close_pct(r["p2"][1], 0.01175873, 1e-4)
    assert close_pct(r["p2"][2], -1.62123444 + 10, 1e-4)


def test_distance_box_box_3():  # :144-216 incl. invariance under a global transform
    q = Q()
    b1, b2 = q.add(box(1, 1, 1)), q.add(box(1, 1, 1))
    t1 = tf((-2, 1, .5), quat=(np.cos(np.pi / 8), 0, 0, np.sin(np.pi / 8)))
    t2 = tf((2, .5, .5), quat=(np.cos(np.pi / 8), 0, np.sin(np.pi / 8), 0))
    r = q.distance(b1, t1, b2, t2)
    assert close_pct(r["min_distance"], 4 - SQ2, 1e-4)
    p1ref, p2ref = np.array([SQ2 / 2 - 2, 1, .5]), np.array([2 - SQ2 / 2, 1, .5])
    assert np.allclose(r["p1"], p1ref, rtol=1e-6, atol=1e-9) and np.allclose(r["p2"], p2ref, rtol=1e-6, atol=1e-9)
    t3 = tf((4, 5, 6), quat=(0.435952844074, -0.718287018243, 0.310622451066, 0.444435113443))
    r = q.distance(b1, compose(t3, t1), b2, compose(t3, t2))
    assert close_pct(r["min_distance"], 4 - SQ2, 1e-4)
    assert np.allclose(r["p1"], apply(t3, p1ref), rtol=1e-6) and np.allclose(r["p2"], apply(t3, p2ref), rtol=1e-6)


def test_distance_box_box_4():  # :218-254 separated / touching / penetrating
    q = Q()
    b1, b2 = q.add(box(1, 1, 1)), q.add(box(1, 1, 1))
    assert close_pct(q.distance(b1, tf((2, 0, 0)), b2, tf())["min_distance"], 1.0, 1e-4)
    assert close_pct(q.distance(b1, tf((1.01, 0, 0)), b2, tf())["min_distance"], 0.01, 2e-3)
    assert close_pct(q.distance(b1, tf((0.99, 0, 0)), b2, tf())["min_distance"], -0.01, 2e-3)
    assert close_pct(q.distance(b1, tf((0, 0, 0)), b2, tf())["min_distance"], -1.0, 2e-3)


# ----------------------------------------------------------------------------- gjk.cpp
@pytest.mark.parametrize("nesterov", [False, True])
@pytest.mark.parametrize("ssr", [0., 0.1, 1., 10., 100.])
def test_gjk_unit_sphere(nesterov, ssr):  # test/gjk.cpp:337-414
    rng = np.random.default_rng(int(ssr * 10) + nesterov)
    for cd in (3, 2.01, 2.0, 1.0):
        for ray in (np.array([1., 0, 0]), None):
            if ray is None:
                ray = rng.normal(size=3)
                ray /= np.linalg.norm(ray)
            sc = make_scenes(emu=False)
            h = int(sc.register_shapes(sphere(1.0, ssr))[0])
            R0, R1 = W.random_rotations(rng, 2)
            t0 = P.make_transforms(R0[None], np.zeros((1, 3)))
            t1 = P.make_transforms(R1[None], (cd * ray)[None])
            g = sc.b["oracle"].gjk_lowlevel(h, t0, h, t1, gjk_max_it=2, gjk_tol=1e-6,
                                            variant=P.NesterovAcceleration if nesterov else P.DefaultGJK)
            expect_collision = cd <= 2 * (1.0 + ssr)
            if expect_collision:
                assert g["gjk_status"] == P.GJK_CollisionWithPenetrationInformation
            else:
                assert g["gjk_status"] == P.GJK_NoCollision
            w0e = R0.T @ ray + ssr * g["normal"]           # tf0.inverse().transform(tf0.T + ray)
            w1e = R0.T @ (cd * ray - ray) - ssr * g["normal"]
            assert np.allclose(g["w0"], w0e, atol=1e-10) and np.allclose(g["w1"], w1e, atol=1e-10)


@pytest.mark.parametrize("T,collide,nesterov,w0e,w1e", [
    ((1.01, 0, 0), False, False, (1., 0, 0), (0., 0, 0)),
    ((1.01, 0, 0), False, True, (1., 0, 0), (0., 0, 0)),
    ((0.5, 0, 0), True, False, (1., 0, 0), (0., 0, 0)),
    ((0.5, 0, 0), True, True, (1., 0, 0), (0., 0, 0)),
    ((-0.5, -0.01, 0), True, False, (0, 1, 0), (0.5, 0, 0)),
    ((-0.5, -0.01, 0), True, True, (0, 1, 0), (0.5, 0, 0)),
])
def test_gjk_triangle_capsule(T, collide, nesterov, w0e, w1e):  # test/gjk.cpp:416-490
    sc = make_scenes(emu=False)
    hc = int(sc.register_shapes(capsule(1., 2.))[0])
    cid = sc.register_convex(np.array([[0., 0, 0], [1., 0, 0], [1., 1, 0]]), None)
    ht = int(sc.register_shapes(P.make_shapes([P.GEOM_TRIANGLE], [[0, 0, 0]], data=[cid]))[0])
    var = P.NesterovAcceleration if nesterov else P.DefaultGJK
    g = sc.b["oracle"].gjk_lowlevel(hc, tf(), ht, tf(T), gjk_max_it=10, gjk_tol=1e-6, variant=var,
                                    run_epa=True, epa_max_it=64, epa_tol=1e-6, epa_guess=(1, 0, 0))
    if collide:
        assert g["gjk_status"] in (P.GJK_Collision, P.GJK_CollisionWithPenetrationInformation)
    else:
        assert g["gjk_status"] == P.GJK_NoCollision
        g2 = sc.b["oracle"].gjk_lowlevel(hc, tf(), ht, tf(T), gjk_max_it=3, gjk_tol=1e-6, guess=g["ray"])
        assert g2["gjk_status"] == P.GJK_NoCollision
    if g["gjk_status"] == P.GJK_Collision:
        assert g["epa_status"] == P.EPA_AccuracyReached
    assert np.allclose(g["w0"], w0e, atol=1e-10)
    assert np.allclose(g["w1"] - np.array(T), w1e, atol=1e-10)


# ------------------------------------------------------------------- capsule_box_{1,2}.cpp
def test_distance_capsule_box_1():  # test/capsule_box_1.cpp:51-115
    q = Q()
    c, b = q.add(capsule(2., 4.)), q.add(box(1., 2., 4.))
    r = q.distance(c, tf((3., 0, 0)), b, tf())
    assert close_pct(r["min_distance"], 0.5, 1e-1)
    assert close_pct(r["p1"][0], 1.0, 1e-1) and abs(r["p1"][1]) < 1e-1
    assert close_pct(r["p2"][0], 0.5, 1e-1) and abs(r["p2"][1]) < 1e-1
    r = q.distance(c, tf((0., 0., 8.)), b, tf())
    assert close_pct(r["min_distance"], 2.0, 1e-1)
    assert abs(r["p1"][0]) < 1e-1 and abs(r["p1"][1]) < 1e-1 and close_pct(r["p1"][2], 4.0, 1e-1)
    assert abs(r["p2"][0]) < 1e-1 and abs(r["p2"][1]) < 1e-1 and close_pct(r["p2"][2], 2.0, 1e-1)
    r = q.distance(c, tf((-10., 0, 0), quat=(SQ2 / 2, 0, SQ2 / 2, 0)), b, tf())
    assert close_pct(r["min_distance"], 5.5, 1e-1)
    assert close_pct(r["p1"][0], -6, 1e-2) and abs(r["p1"][1]) < 1e-1 and abs(r["p1"][2]) < 1e-1
    assert close_pct(r["p2"][0], -0.5, 1e-2) and abs(r["p2"][1]) < 1e-1 and abs(r["p2"][2]) < 1e-1


def test_distance_capsule_box_2():  # test/capsule_box_2.cpp:51-85
    q = Q()
    c, b = q.add(capsule(2., 4.)), q.add(box(1., 2., 4.))
    r = q.distance(c, tf((-10., 0.8, 1.5), quat=(SQ2 / 2, 0, SQ2 / 2, 0)), b, tf())
    assert close_pct(r["min_distance"], 5.5, 1e-2)
    assert close_pct(r["p1"][0], -6, 1e-2) and close_pct(r["p1"][1], 0.8, 1e-1) and close_pct(r["p1"][2], 1.5, 1e-2)
    assert close_pct(r["p2"][0], -0.5, 1e-2) and close_pct(r["p2"][1], 0.8, 1e-1) and close_pct(r["p2"][2], 1.5, 1e-2)


# --------------------------------------------------------------------- capsule_capsule.cpp
def test_distance_capsule_capsule():  # test/capsule_capsule.cpp:218-356
    q = Q()
    c = q.add(capsule(5, 10))
    assert close_pct(q.distance(c, tf(), c, tf((20.1, 0, 0)))["min_distance"], 10.1, 1e-6)
    assert close_pct(q.distance(c, tf(), c, tf((20, 20, 0)))["min_distance"], np.sqrt(800) - 10, 1e-6)
    assert close_pct(q.distance(c, tf(), c, tf((0, 0, 20.1)))["min_distance"], 0.1, 1e-6)
    r = q.distance(c, tf(), c, tf((0, 0, 25.1), quat=(SQ2 / 2, 0, SQ2 / 2, 0)))
    assert close_pct(r["min_distance"], 10.1, 1e-6)
    assert abs(r["p1"][0]) < 1e-4 and abs(r["p1"][1]) < 1e-4 and close_pct(r["p1"][2], 10, 1e-4)
    assert abs(r["p2"][0]) < 1e-4 and abs(r["p2"][1]) < 1e-4 and close_pct(r["p2"][2], 20.1, 1e-4)


def test_collision_capsule_capsule_trivial_and_aligned():  # :56-216 (random property tests, seeded here)
    rng = np.random.default_rng(0)
    n = 20000
    sc = make_scenes()
    hs = sc.register_shapes(np.concatenate([capsule(1., 0.), sphere(1.)]))
    t1 = P.make_transforms(W.random_rotations(rng, n), (rng.random((n, 3)) * 2 - 1) * 2)
    t2 = P.make_transforms(W.random_rotations(rng, n), (rng.random((n, 3)) * 2 - 1) * 2)
    rc = sc.b["oracle"].batch_collide(np.full(n, hs[0]), t1, np.full(n, hs[0]), t2)
    rs = sc.b["oracle"].batch_collide(np.full(n, hs[1]), t1, np.full(n, hs[1]), t2)
    assert np.array_equal(rc["num_contacts"], rs["num_contacts"])
    m = rc["num_contacts"] == 0
    assert np.allclose(rc["distance_lower_bound"][m], rs["distance_lower_bound"][m], rtol=1e-8)
    # aligned capsules (radius .01, length .2)
    radius, length = 0.01, 0.2
    hc = int(sc.register_shapes(capsule(radius, length))[0])
    R = W.random_rotations(rng, n)
    z = np.zeros((n, 3))
    for p2, expect in (((0, 0, 2 * (length / 2 + radius) + 1e-3), 0),
                       ((0, 0, min(length / 2, radius) * (1 - 1e-2)), 1)):
        r = sc.b["oracle"].batch_collide(np.full(n, hc), P.make_transforms(R, z), np.full(n, hc),
                                         P.make_transforms(R, np.tile(p2, (n, 1))))
        assert np.all(r["num_contacts"] == expect)
    for p2, expect in (((0, 0, 2 * (length / 2 + radius) + 1e-3), 0), ((0, 0, 0.01), 1)):
        Tr = rng.random((n, 3)) * 2 - 1
        t1 = P.make_transforms(R, Tr)
        t2 = P.make_transforms(R, np.einsum("nij,j->ni", R, np.array(p2)) + Tr)
        r = sc.b["oracle"].batch_collide(np.full(n, hc), t1, np.full(n, hc), t2)
        e = sc.b["emu"].batch_collide(np.full(n, hc), t1, np.full(n, hc), t2)
        assert np.all(r["num_contacts"] == expect)
        compare_distance(r, e, what="capsule aligned")


# ---------------------------------------------------------- geometric_shapes.cpp :3568-4074
def _rand_tf(rng):
    return P.make_transforms(W.random_rotations(rng, 1), (rng.random((1, 3)) * 2 - 1) * 10)


def test_shape_distance_spheresphere():  # :3568-3638
    q = Q()
    rng = np.random.default_rng(1)
    s1, s2 = q.add(sphere(20)), q.add(sphere(10))
    tr = _rand_tf(rng)
    for base in (tf(), tr):
        for x, want in ((40, 10), (30.1, 0.1), (29.9, None)):
            for swap in (False, True):
                a, b = (compose(base, tf((x, 0, 0))), base) if swap else (base, compose(base, tf((x, 0, 0))))
                d = q.distance(s1, a, s2, b)["min_distance"]
                if want is None:
                    assert d < 0
                else:
                    assert abs(d - want) < 0.001


def test_shape_distance_boxbox():  # :3640-3715
    q = Q()
    rng = np.random.default_rng(2)
    s1, s2 = q.add(box(20, 40, 50)), q.add(box(10, 10, 10))
    tr = _rand_tf(rng)
    assert q.distance(s1, tf(), s2, tf())["min_distance"] <= 0
    assert q.distance(s1, tr, s2, tr)["min_distance"] <= 0
    for T, want in (((10.1, 0, 0), 0.1), ((20.1, 0, 0), 10.1), ((0, 20.2, 0), 10.2), ((10.1, 10.1, 0), 0.1 * 1.414)):
        assert abs(q.distance(s2, tf(), s2, tf(T))["min_distance"] - want) < 0.001
    assert abs(q.distance(s1, tr, s2, compose(tr, tf((15.1, 0, 0))))["min_distance"] - 0.1) < 0.001
    assert abs(q.distance(s1, tf(), s2, tf((20, 0, 0)))["min_distance"] - 5) < 0.001
    assert abs(q.distance(s1, tr, s2, compose(tr, tf((20, 0, 0))))["min_distance"] - 5) < 0.001


def test_shape_distance_cylinderbox():  # :3717-3766 (witness points consistent: both in or both out)
    q = Q()
    cyl, bx = q.add(cylinder(0.029, 0.1)), q.add(box(1.6, 0.6, 0.025))
    t1 = tf((0.041218354748013122, 1.2022554710435607, 0.77338855025700015),
            quat=(0.5279170511703305, -0.50981118132505521, -0.67596178682051911, 0.0668715876735793))
    t2 = tf((-0.29936284351096382, 0.80023864435868775, 0.71750000000000003),
            quat=(0.70738826916719977, 0, 0, 0.70682518110536596))

    def inv(t, p):
        return t["R"][0].reshape(3, 3).T.T @ (p - t["T"][0])

    for (a, ta, b, tb, flip) in ((cyl, t1, bx, t2, False), (bx, t2, cyl, t1, True)):
        r = q.distance(a, ta, b, tb)
        p_on_cyl, p_on_box = (r["p2"], r["p1"]) if flip else (r["p1"], r["p2"])
        # reference naming: p2 is tested against the cylinder, p1 against the box
        p2loc = inv(t1, p_on_box)
        in_cyl = abs(p2loc[2]) <= 0.05 and p2loc[0] ** 2 + p2loc[1] ** 2 <= 0.029
        p1loc = inv(t2, p_on_cyl)
        in_box = bool(np.all(np.abs(p1loc) <= np.array([0.8, 0.3, 0.0125])))
        assert (not in_cyl and not in_box) or (in_cyl and in_box)
    cyl2 = q.add(cylinder(0.06, 0.1))
    t1b = tf((-0.66734052046473924, 0.22219183277457269, 0.76825248755616293),
             quat=(0.52613359459338371, 0.32189408354839893, 0.70415587451837913, -0.35175580165512249))
    r = q.distance(cyl2, t1b, bx, t2)
    assert np.isfinite(r["min_distance"])


def test_shape_distance_boxsphere():  # :3768-3826
    q = Q()
    rng = np.random.default_rng(3)
    s1, s2 = q.add(sphere(20)), q.add(box(5, 5, 5))
    tr = _rand_tf(rng)
    Rt = tr["R"][0].reshape(3, 3).T
    N = 10
    for i in range(N + 1):
        dbox = 0.0001 + (20 + 2.5) * i * 4 / (3 * N)
        r = q.distance(s1, tf((dbox, 0., 0.)), s2, tf())
        assert close_pct(r["min_distance"], dbox - 20 - 2.5, 1e-6)
        assert np.allclose(r["normal"], [-1, 0, 0], atol=1e-6)
        r = q.distance(s1, compose(tr, tf((dbox, 0., 0.))), s2, tr)
        assert close_pct(r["min_distance"], dbox - 20 - 2.5, 1e-6)
        assert np.allclose(r["normal"], -Rt[:, 0], atol=1e-6)
    assert q.distance(s1, tf(), s2, tf())["min_distance"] <= 0
    assert abs(q.distance(s1, tf(), s2, tf((22.6, 0, 0)))["min_distance"] - 0.1) < 0.001
    assert abs(q.distance(s1, tr, s2, compose(tr, tf((22.6, 0, 0))))["min_distance"] - 0.1) < 0.01
    assert abs(q.distance(s1, tf(), s2, tf((40, 0, 0)))["min_distance"] - 17.5) < 0.001
    assert abs(q.distance(s1, tr, s2, compose(tr, tf((40, 0, 0))))["min_distance"] - 17.5) < 0.001


@pytest.mark.parametrize("mk1,mk2,tolA,tolB", [
    (lambda: cylinder(5, 10), lambda: cylinder(5, 10), 0.001, 0.001),   # :3828-3894
    (lambda: cylinder(5, 10), lambda: cone(5, 10), 0.01, 0.02),         # :3964-4030
])
def test_shape_distance_curved(mk1, mk2, tolA, tolB):
    q = Q()
    rng = np.random.default_rng(4)
    s1, s2 = q.add(mk1()), q.add(mk2())
    tr = _rand_tf(rng)
    assert q.distance(s1, tf(), s2, tf())["min_distance"] <= 0       # exactly superposed: EPA worst case
    assert q.distance(s1, tr, s2, tr)["min_distance"] <= 0
    assert abs(q.distance(s1, tf(), s2, tf((10.1, 0, 0)))["min_distance"] - 0.1) < tolA
    assert abs(q.distance(s1, tr, s2, compose(tr, tf((10.1, 0, 0))))["min_distance"] - 0.1) < tolB
    assert abs(q.distance(s1, tf(), s2, tf((40, 0, 0)))["min_distance"] - 30) < 0.01
    assert abs(q.distance(s1, tr, s2, compose(tr, tf((40, 0, 0))))["min_distance"] - 30) < 0.1


def test_shape_distance_conecone():  # :3896-3962
    q = Q()
    rng = np.random.default_rng(5)
    s1 = q.add(cone(5, 10))
    tr = _rand_tf(rng)
    assert q.distance(s1, tf(), s1, tf())["min_distance"] <= 0
    assert abs(q.distance(s1, tf(), s1, tf((10.1, 0, 0)))["min_distance"] - 0.1) < 0.001
    assert abs(q.distance(s1, tr, s1, compose(tr, tf((10.1, 0, 0))))["min_distance"] - 0.1) < 0.001
    assert abs(q.distance(s1, tf(), s1, tf((0, 0, 40)))["min_distance"] - 30) < 1
    assert abs(q.distance(s1, tr, s1, compose(tr, tf((0, 0, 40))))["min_distance"] - 30) < 1


# --------------------------------------------------------------------------- simple.cpp
# Project::project{Line,Triangle,Tetrahedra}(..., p) known answers (test/simple.cpp:16-246).
# The hot path uses the *Origin forms (src/intersect.cpp:570-705); by translation invariance
# project(v..., p) == projectOrigin(v - p ...), which is what is checked here.
_LINE = [((1, 0, 0), 3, 0, (0.5, 0.5)), ((-1, 0, 0), 1, 1, (1, 0)), ((3, 0, 0), 2, 1, (0, 1))]
_TRI = [((1, 1, 1), 7, 4 / 3., (1 / 3., 1 / 3., 1 / 3.)), ((0, 0, 1.5), 1, 0.25, (1, 0, 0)),
        ((1.5, 0, 0), 4, 0.25, (0, 0, 1)), ((0, 1.5, 0), 2, 0.25, (0, 1, 0)),
        ((1, 1, 0), 6, 0.5, (0, 0.5, 0.5)), ((1, 0, 1), 5, 0.5, (0.5, 0, 0.5)), ((0, 1, 1), 3, 0.5, (0.5, 0.5, 0))]
_TET = [((0.5, 0.5, 0.5), 15, 0, (.25, .25, .25, .25)), ((0, 0, 0), 7, 1 / 3., (1 / 3., 1 / 3., 1 / 3., 0)),
        ((0, 1, 1), 11, 1 / 3., (1 / 3., 1 / 3., 0, 1 / 3.)), ((1, 1, 0), 14, 1 / 3., (0, 1 / 3., 1 / 3., 1 / 3.)),
        ((1, 0, 1), 13, 1 / 3., (1 / 3., 0, 1 / 3., 1 / 3.)), ((1.5, 1.5, 1.5), 8, 0.75, (0, 0, 0, 1)),
        ((1.5, -0.5, -0.5), 4, 0.75, (0, 0, 1, 0)), ((-0.5, -0.5, 1.5), 1, 0.75, (1, 0, 0, 0)),
        ((-0.5, 1.5, -0.5), 2, 0.75, (0, 1, 0, 0)), ((0.5, -0.5, 0.5), 5, 0.25, (0.5, 0, 0.5, 0)),
        ((0.5, 1.5, 0.5), 10, 0.25, (0, 0.5, 0, 0.5)), ((1.5, 0.5, 0.5), 12, 0.25, (0, 0, 0.5, 0.5)),
        ((-0.5, 0.5, 0.5), 3, 0.25, (0.5, 0.5, 0, 0)), ((0.5, 0.5, 1.5), 9, 0.25, (0.5, 0, 0, 0.5)),
        ((0.5, 0.5, -0.5), 6, 0.25, (0, 0.5, 0.5, 0))]


@pytest.mark.parametrize("verts,cases", [
    (((0, 0, 0), (2, 0, 0)), _LINE),                       # projection_test_line :16-42
    (((0, 0, 1), (0, 1, 0), (1, 0, 0)), _TRI),             # projection_test_triangle :44-102
    (((0, 0, 1), (0, 1, 0), (1, 0, 0), (1, 1, 1)), _TET),  # projection_test_tetrahedron :104-246
])
def test_projection_known_answers(verts, cases):
    o = make_scenes(emu=False).b["oracle"]
    v = [np.array(x, dtype=float) for x in verts]
    for p, enc, sq, param in cases:
        got_param, got_sq, got_enc = o.project(*[x - np.array(p, dtype=float) for x in v])
        assert got_enc == enc, (p, got_enc, enc)
        assert abs(got_sq - sq) < 1e-6
        assert np.allclose(got_param[:len(param)], param, atol=1e-6)


# -------------------------------------------------------------------- security_margin.cpp
def test_security_margin_sphere_sphere_and_box_box():  # test/security_margin.cpp (margin semantics)
    q = Q()
    s = q.add(sphere(1))
    d = 2.1
    r = q.collide(s, tf(), s, tf((d, 0, 0)))
    assert r["num_contacts"] == 0 and np.isclose(r["distance_lower_bound"], 0.1, rtol=1e-9)
    r = q.collide(s, tf(), s, tf((d, 0, 0)), security_margin=0.1 + 1e-9)
    assert r["num_contacts"] == 1 and np.isclose(r["distance"], 0.1, rtol=1e-9)
    assert np.isclose(r["distance_lower_bound"], -1e-9, atol=1e-12)
    r = q.collide(s, tf(), s, tf((d, 0, 0)), security_margin=-0.1)
    assert r["num_contacts"] == 0
    b = q.add(box(2, 2, 2))
    r = q.collide(b, tf(), b, tf((2.1, 0, 0)), security_margin=0.11)
    assert r["num_contacts"] == 1 and abs(r["distance"] - 0.1) < 1e-6
    r = q.collide(b, tf(), b, tf((2.1, 0, 0)), security_margin=0.09)
    assert r["num_contacts"] == 0 and abs(r["distance_lower_bound"] - 0.01) < 1e-6
    # penetrating boxes with a negative margin smaller than the depth still collide
    r = q.collide(b, tf(), b, tf((1.9, 0, 0)), security_margin=-0.05)
    assert r["num_contacts"] == 1 and abs(r["distance"] + 0.1) < 1e-6


# ------------------------------------------------- accelerated_gjk.cpp / gjk_convergence_criterion.cpp
def test_accelerated_gjk_and_criteria_agree():  # test/accelerated_gjk.cpp:107-190, gjk_convergence_criterion.cpp:92-164
    sc = make_scenes()
    rng = np.random.default_rng(9)
    pts, tris = W.icosahedron_from_ellipsoid((0.7, 1.0, 0.8))
    cid = sc.register_convex(pts, tris)
    recs = np.concatenate([P.make_shapes([P.GEOM_ELLIPSOID], [[0.7, 1.0, 0.8]]), capsule(0.5, 1.0), box(1, .8, .6),
                           P.make_shapes([P.GEOM_CONVEX], [[0, 0, 0]], data=[cid])])
    hs = sc.register_shapes(recs)
    n = 4000
    h1 = hs[rng.integers(0, len(hs), n)]
    h2 = hs[rng.integers(0, len(hs), n)]
    t1 = W.identity_transforms(n)
    T = np.stack([rng.uniform(-3, 3, n), rng.uniform(-3, 3, n), rng.uniform(0, 3, n)], axis=1)  # :122
    t2 = P.make_transforms(W.random_rotations(rng, n), T)
    base = None
    for variant in (P.DefaultGJK, P.NesterovAcceleration, P.PolyakAcceleration):
        for crit in (P.Default, P.DualityGap, P.Hybrid):
            req = P.DistanceRequestPOD(gjk_variant=variant, gjk_convergence_criterion=crit,
                                       gjk_convergence_criterion_type=P.Absolute, enable_signed_distance=0)
            r = sc.b["oracle"].batch_distance(h1, t1, h2, t2, req)
            e = sc.b["emu"].batch_distance(h1, t1, h2, t2, req)
            compare_distance(r, e, what="variant %d crit %d" % (variant, crit))
            assert np.all((r["iterations"] & 0xffff) < 128)
            if base is None:
                base = r
            else:
                g0, g1 = P.status_gjk(base["status"]), P.status_gjk(r["status"])
                col0, col1 = g0 >= P.GJK_CollisionWithPenetrationInformation, g1 >= P.GJK_CollisionWithPenetrationInformation
                # same collision verdict except within tolerance of touching
                disagree = col0 != col1
                assert np.all(np.abs(base["min_distance"][disagree]) < 1e-3)
                m = ~col0 & ~col1
                assert np.allclose(base["min_distance"][m], r["min_distance"][m], atol=1e-4)
