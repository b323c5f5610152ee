"""Known answers of the Plane / Halfspace family, transcribed from the reference's own test
(/root/reference/test/geometric_shapes.cpp, testShapeCollide :167-209 with compareContact :130-163, tolerance
1e-9 unless the case says otherwise): every case once with the poses of the test and once under a common rigid
transform, as the reference does (its transform is random; a fixed one is used here).

Checked on the oracle's records; the host build of the device code (tests/emu) and, where oracle/_ref exists, the
reference build must return the same bits.  CPU only:
the CUDA kernels are compared with the oracle on the plane family by tests/test_plane_halfspace.py.
"""
import numpy as np
import pytest

from tests.common import P, compare_distance, make_scenes
from hppfcl_b200 import workloads as W

GLOBAL = ((4.0, 5.0, 6.0), (0.435952844074, -0.718287018243, 0.310622451066, 0.444435113443))


def tf(T=(0, 0, 0), quat=None):
    R = np.eye(3) if quat is None else W.quat_to_rot(*quat)
    return P.make_transforms(np.asarray(R)[None], np.asarray(T, dtype=float)[None])


def compose(a, b):  # Transform3f operator* (transform.h:186-188)
    Ra = a["R"][0].reshape(3, 3).T
    Rb = b["R"][0].reshape(3, 3).T
    return P.make_transforms((Ra @ Rb)[None], (Ra @ b["T"][0] + a["T"][0])[None])


def rot(t):
    return t["R"][0].reshape(3, 3).T


def apply(t, v):
    return rot(t) @ np.asarray(v, dtype=float) + t["T"][0]


class World:
    def __init__(self):
        self.sc = make_scenes(ref=True)
        self.g = tf(*GLOBAL)

    def shape(self, kind, *p):
        half = {"box": lambda x, y, z: (P.GEOM_BOX, [x / 2, y / 2, z / 2]),
                "sphere": lambda r: (P.GEOM_SPHERE, [r, 0, 0]),
                "capsule": lambda r, lz: (P.GEOM_CAPSULE, [r, lz / 2, 0]),
                "cylinder": lambda r, lz: (P.GEOM_CYLINDER, [r, lz / 2, 0]),
                "cone": lambda r, lz: (P.GEOM_CONE, [r, lz / 2, 0])}[kind](*p)
        return int(self.sc.register_shapes(P.make_shapes([half[0]], [half[1]]))[0])

    def triangle(self, a, b, c):
        cid = self.sc.register_convex(np.array([a, b, c], dtype=float), None)
        return int(self.sc.register_shapes(P.make_shapes([P.GEOM_TRIANGLE], [[0, 0, 0]], data=[cid]))[0])

    def halfspace(self, n, d):
        return int(self.sc.register_halfspaces(P.GEOM_HALFSPACE, np.array([[*n, d]], dtype=float))[0])

    def plane(self, n, d):
        return int(self.sc.register_halfspaces(P.GEOM_PLANE, np.array([[*n, d]], dtype=float))[0])

    def collide(self, h1, t1, h2, t2, enable_contact=1):
        req = P.CollisionRequestPOD(enable_contact=enable_contact)
        self.sc.commit()
        ro = self.sc.b["oracle"].batch_collide([h1], t1, [h2], t2, req)
        re = self.sc.b["emu"].batch_collide([h1], t1, [h2], t2, req)
        compare_distance(ro, re, what="known answers, host build of the device code")
        if "ref" in self.sc.b:  # the reference build: the same bits (it leaves `distance` unset without a contact)
            rr = self.sc.b["ref"].batch_collide([h1], t1, [h2], t2, req)
            a, b = rr.copy(), ro.copy()
            nc = a["num_contacts"] == 0
            a["distance"][nc] = 0
            b["distance"][nc] = 0
            compare_distance(a, b, what="known answers, reference build")
        return ro[0]

    def check(self, h1, t1, h2, t2, expect, point=None, depth=None, normal=None, opposite=False, tol=1e-9,
              both_frames=True):
        """testShapeCollide :167-209, in the frame of the test and under the common transform"""
        for g in ((None, self.g) if both_frames else (None,)):
            a1, a2 = (t1, t2) if g is None else (compose(g, t1), compose(g, t2))
            for ec in (0, 1):
                r = self.collide(h1, a1, h2, a2, enable_contact=ec)
                assert (r["num_contacts"] > 0) == expect, (h1, h2, t1["T"], t2["T"], ec)
            if not expect:
                continue
            assert r["num_contacts"] == 1
            if point is not None:
                e = np.asarray(point, dtype=float) if g is None else apply(g, point)
                assert np.linalg.norm(r["pos"] - e) < tol, (r["pos"], e)
            if depth is not None:
                assert abs(r["distance"] - depth) < tol, (r["distance"], depth)
            if normal is not None:
                e = np.asarray(normal, dtype=float) if g is None else rot(g) @ np.asarray(normal, dtype=float)
                ok = np.linalg.norm(r["normal"] - e) < tol
                if not ok and opposite:
                    ok = np.linalg.norm(r["normal"] + e) < tol
                assert ok, (r["normal"], e)


def axis(k, v=1.0):
    e = np.zeros(3)
    e[k] = v
    return e


# ---------------------------------------------------------------------------------- halfspace x solid
# one row per reference block: (shape, extent of the shape along each axis, contact offset along the other axes)
#   the halfspace {x : n.x <= 0} moved by t along its normal n = e_k covers the shape's points with x_k <= t:
#   deepest point -extent, contact = midpoint of deepest point and its projection, depth = -(extent + t)
@pytest.mark.parametrize("kind,params,extent,base", [
    ("sphere", (10,), (10, 10, 10), (0, 0, 0)),          # collide_halfspacesphere :1275-1362 (x only there)
    ("box", (5, 10, 20), (2.5, 5, 10), (0, 0, 0)),       # collide_halfspacebox :1473-1557 (x only there)
    ("capsule", (5, 10), (5, 5, 10), (0, 0, 0)),         # collide_halfspacecapsule :1624-1860
    ("cylinder", (5, 10), (5, 5, 5), (0, 0, 0)),         # collide_halfspacecylinder :2128-2370
    ("cone", (5, 10), (5, 5, 5), (0, 0, -5)),            # collide_halfspacecone :2673-2915 (deepest point on the base circle)
])
def test_halfspace_solid(kind, params, extent, base):
    w = World()
    s = w.shape(kind, *params)
    for k in range(3):
        hs = w.halfspace(axis(k), 0)
        e = extent[k]
        side = np.array(base, dtype=float)
        side[k] = 0
        if kind == "cone" and k == 2:
            side[:] = 0
        # the reference has the box (and the sphere) against the x normal only; along y and z the box support's
        # `inflate` factor 1 + 1e-10 (support_functions.cpp:146, restated in oracle/narrowphase.cpp) times the
        # half extent reaches the 1e-9 of the test, so those added rows allow for it
        tol = 1e-9 if (kind != "box" or k == 0) else 1e-9 + 1.01e-10 * e
        for t in (0.0, e / 2, -e / 2, e + 0.1):
            point = side + axis(k, (t - e) / 2)
            w.check(s, tf(), hs, tf(axis(k, t)), True, point, -(e + t), -axis(k), tol=tol)
        w.check(s, tf(), hs, tf(axis(k, -(e + 0.1))), False)


def test_halfspace_box_rotated():  # :1549-1556: any rotation of a box that contains the origin still collides
    w = World()
    s, hs = w.shape("box", 5, 10, 20), w.halfspace((1, 0, 0), 0)
    w.check(s, tf(quat=GLOBAL[1]), hs, tf(), True, both_frames=False)


# -------------------------------------------------------------------------------------- plane x solid
def test_plane_sphere():  # collide_planesphere :1364-1471
    w = World()
    s, pl = w.shape("sphere", 10), w.plane((1, 0, 0), 0)
    for eps, n in ((1e-6, (-1, 0, 0)), (-1e-6, (1, 0, 0))):
        p1 = np.array([-10 + eps, 0, 0]) if eps > 0 else np.array([10 + eps, 0, 0])
        w.check(s, tf((eps, 0, 0)), pl, tf(), True, p1 / 2, -10 + abs(eps), n, both_frames=False)
        g = w.g  # :1392-1397: under the transform the sign of the normal is free
        w.check(s, compose(g, tf((eps, 0, 0))), pl, g, True, apply(g, p1 / 2), -10 + abs(eps), rot(g) @ np.array(n, dtype=float),
                opposite=True, both_frames=False)
    w.check(s, tf(), pl, tf((5, 0, 0)), True, (7.5, 0, 0), -5, (1, 0, 0))
    w.check(s, tf(), pl, tf((-5, 0, 0)), True, (-7.5, 0, 0), -5, (-1, 0, 0))
    w.check(s, tf(), pl, tf((-10.1, 0, 0)), False)
    w.check(s, tf(), pl, tf((10.1, 0, 0)), False)


def test_plane_box():  # collide_planebox :1559-1622
    w = World()
    s, pl = w.shape("box", 5, 10, 20), w.plane((1, 0, 0), 0)
    w.check(s, tf(), pl, tf(), True, (1.25, 0, 0), -2.5, (1, 0, 0))
    w.check(s, tf(), pl, tf((1.25, 0, 0)), True, (1.875, 0, 0), -1.25, (1, 0, 0))
    w.check(s, tf(), pl, tf((-1.25, 0, 0)), True, (-1.875, 0, 0), -1.25, (-1, 0, 0))
    w.check(s, tf(), pl, tf((2.51, 0, 0)), False)
    w.check(s, tf(), pl, tf((-2.51, 0, 0)), False)
    w.check(s, tf(quat=GLOBAL[1]), pl, tf(), True, both_frames=False)


def test_plane_capsule():  # collide_planecapsule :1862-2126 (the reference checks depth and, off-centre, the normal)
    w = World()
    s = w.shape("capsule", 5, 10)
    for k, e in ((0, 5.0), (1, 5.0), (2, 10.0)):
        pl = w.plane(axis(k), 0)
        w.check(s, tf(), pl, tf(), True, None, -e, axis(k), opposite=True)
        w.check(s, tf(), pl, tf(axis(k, 2.5)), True, None, -(e - 2.5), axis(k))
        w.check(s, tf(), pl, tf(axis(k, -2.5)), True, None, -(e - 2.5), -axis(k))
        w.check(s, tf(), pl, tf(axis(k, e + 0.1)), False)
        w.check(s, tf(), pl, tf(axis(k, -(e + 0.1))), False)


@pytest.mark.parametrize("kind,base_z", [("cylinder", 0.0),   # collide_planecylinder :2372-2671
                                         ("cone", -5.0)])     # collide_planecone :2917-3212
def test_plane_cylinder_cone(kind, base_z):
    w = World()
    s = w.shape(kind, 5, 10)
    for k in range(3):
        pl = w.plane(axis(k), 0)
        side = np.array([0, 0, base_z if k < 2 else 0.0])
        for eps in (1e-6, -1e-6):  # just off-centre: the deeper side decides; the sign of the normal is free in the reference
            p1 = side + axis(k, (-5 if eps > 0 else 5) + eps)
            p2 = side.copy()
            w.check(s, tf(axis(k, eps)), pl, tf(), True, (p1 + p2) / 2, -5 + abs(eps), axis(k), opposite=True)
        for t, sgn in ((2.5, 1.0), (-2.5, -1.0)):
            p1, p2 = side + axis(k, 5 * sgn), side + axis(k, t)
            w.check(s, tf(), pl, tf(axis(k, t)), True, (p1 + p2) / 2, -2.5, axis(k, sgn))
        far = 5.1 if k < 2 else 10.1
        w.check(s, tf(), pl, tf(axis(k, far)), False)
        w.check(s, tf(), pl, tf(axis(k, -far)), False)


# --------------------------------------------------------------------------------- triangles
TRI_HS = [((20, 0, 0), (-20, 0, 0), (0, 20, 0)), ((30, 0, 0), (-20, 0, 0), (0, 0, 20)), ((0, 30, 0), (0, -10, 0), (0, 0, 20))]


@pytest.mark.parametrize("tri", TRI_HS)
def test_halfspace_triangle(tri):  # collide_halfspacetriangle :1032-1146
    w = World()
    hs, t = w.halfspace((0, 0, 1), 0), w.triangle(*tri)
    w.check(hs, tf(), t, tf((0, 0, -0.001)), True, None, None, (0, 0, 1))
    for T in ((0, 0, 0.001), (1, 1, 0.001), (-1, -1, 0.001)):
        w.check(hs, tf(), t, tf(T), False)


@pytest.mark.parametrize("k,tri", [
    (2, ((20, 0, 0.05), (-20, 0, 0.05), (0, 20, -0.1))),
    (1, ((30, 0.05, 0), (-20, 0.05, 0), (0, -0.1, 20))),
    (0, ((0.05, 30, 0), (0.05, -10, 0), (-0.1, 0, 20))),
])
def test_plane_triangle(k, tri):  # collide_planetriangle :1148-1273
    w = World()
    pl, t = w.plane(axis(k), 0), w.triangle(*tri)
    w.check(pl, tf(), t, tf(), True, None, None, -axis(k))
    w.check(pl, tf(), t, tf(axis(k, 0.05)), True, None, None, axis(k))
    w.check(pl, tf(), t, tf(axis(k, -0.06)), False)
    w.check(pl, tf(), t, tf(axis(k, 0.11)), False)


# ------------------------------------------------------------------- the family among itself
N_RANDOM = np.array([0.680375, -0.211234, 0.566198])  # stands for Vec3f::Random()
N_RANDOM = N_RANDOM / np.linalg.norm(N_RANDOM)
S3 = 1 / np.sqrt(3.0)


def test_plane_plane():  # collide_planeplane :3214-3332
    w = World()
    n, off = N_RANDOM, 3.14
    p1, p2 = w.plane(n, off), w.plane(n, off)
    w.check(p1, tf(), p2, tf(), True, n * off, 0.0, n, both_frames=False)
    g = w.g  # :3240-3249
    ng = rot(g) @ n
    w.check(p1, g, p2, g, True, ng * (off + ng @ g["T"][0]), 0.0, ng, both_frames=False)
    w.check(p1, tf(), w.plane(n, off + 1.19841), tf(), False)
    w.check(p1, tf(), w.plane(n, off - 1.19841), tf(), False)
    a, b = w.plane((1, 0, 0), 3.14), w.plane((0, 0, 1), -2.13)
    w.check(a, tf(), b, tf(), True, (3.14, 0, -2.13), None, (0, -1, 0), both_frames=False)
    w.check(a, w.g, b, w.g, True, None, None, rot(w.g) @ np.array([0, -1.0, 0]), both_frames=False)
    c = w.plane((1, 1, 1), -2.13)
    w.check(a, tf(), c, tf(), True, None, None, (0, -0.5774, 0.5774), tol=1e-3)


def test_halfspace_halfspace():  # collide_halfspacehalfspace :3334-3448
    w = World()
    n, off = N_RANDOM, 3.14
    h1 = w.halfspace(n, off)
    w.check(h1, tf(), w.halfspace(n, off), tf(), True, None, None, n)
    w.check(h1, tf(), w.halfspace(n, off + 1.19841), tf(), True, None, None, n)
    off2 = off - 1.19841
    w.check(h1, tf(), w.halfspace(-n, -off2), tf(), True, None, off2 - off, n)
    a = w.halfspace((1, 0, 0), 3.14)
    w.check(a, tf(), w.halfspace((0, 0, 1), -2.13), tf(), True, None, None, (0, -1, 0))
    w.check(a, tf(), w.halfspace((1, 1, 1), -2.13), tf(), True, None, None, (0, -0.5774, 0.5774), tol=1e-3)


def test_halfspace_plane():  # collide_halfspaceplane :3450-3566
    w = World()
    n, off = N_RANDOM, 3.14
    hf = w.halfspace(n, off)
    w.check(hf, tf(), w.plane(n, off), tf(), True, None, 0.0, n)
    w.check(hf, tf(), w.plane(n, off + 1.19841), tf(), False)
    off2 = off - 1.19841
    w.check(hf, tf(), w.plane(n, off2), tf(), True, None, off2 - off, n)
    a = w.halfspace((1, 0, 0), 3.14)
    w.check(a, tf(), w.plane((0, 0, 1), -2.13), tf(), True, None, None, (0, -1, 0))
    w.check(a, tf(), w.plane((1, 1, 1), -2.13), tf(), True, None, None, (0, -0.5774, 0.5774), tol=1e-3)
