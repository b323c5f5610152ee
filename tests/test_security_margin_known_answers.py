"""collide() with a security margin: the reference's known answers
(/root/reference/test/security_margin.cpp:182-506 -- sphere-sphere, capsule-capsule, box-box, and box / box-as-convex-
hull against a sphere), every block with its own tolerances (BOOST_CHECK_CLOSE is in percent, BOOST_CHECK_SMALL
absolute).  Checked on the oracle; the host build of the device code and, where oracle/_ref exists, the reference build must
return the same bits.  CPU only.
"""
import numpy as np
import pytest

from tests.common import P, compare_distance
from tests.test_plane_known_answers import World, tf


def close_pct(a, b, pct):
    d = abs(a - b)
    return d <= pct / 100 * abs(a) and d <= pct / 100 * abs(b)


def collide(w, h1, h2, T, margin=0.0):
    w.sc.commit()
    req = P.CollisionRequestPOD(security_margin=margin)  # CollisionRequest(CONTACT, 1)
    ro = w.sc.b["oracle"].batch_collide([h1], tf(), [h2], tf(T), req)
    re = w.sc.b["emu"].batch_collide([h1], tf(), [h2], tf(T), req)
    assert ro.tobytes() == re.tobytes()
    if "ref" in w.sc.b:  # the reference build (it leaves `distance` unset without a contact)
        rr = w.sc.b["ref"].batch_collide([h1], tf(), [h2], tf(T), req)
        a, b = rr.copy(), ro.copy()
        nc = a["num_contacts"] == 0
        a["distance"][nc] = 0
        b["distance"][nc] = 0
        compare_distance(a, b, what="security margin, reference build")
    return ro[0]


@pytest.mark.parametrize("kind,p1,p2,touch,axis", [
    ("sphere", (1,), (2,), (0, 0, 3), 2),      # sphere_sphere :182-258
    ("capsule", (0.5, 1.0), (0.5, 1.0), (0, 1.0, 0), 1),  # capsule_capsule :260-332
])
def test_margin_round_shapes(kind, p1, p2, touch, axis):
    w = World()
    a, b = w.shape(kind, *p1), w.shape(kind, *p2)
    touch = np.array(touch, dtype=float)
    e = np.zeros(3)
    e[axis] = 1
    r = collide(w, a, b, touch)  # no margin, touching
    assert r["num_contacts"] == 1 and abs(r["distance_lower_bound"]) < 1e-8 and abs(r["distance"]) < 1e-8
    r = collide(w, a, b, touch + 0.01 * e)  # no margin, 0.01 apart
    assert r["num_contacts"] == 0 and close_pct(r["distance_lower_bound"], 0.01, 1e-8)
    r = collide(w, a, b, touch + 0.01 * e, margin=0.01)  # margin reaches
    assert r["num_contacts"] == 1 and abs(r["distance_lower_bound"]) < 1e-8 and close_pct(r["distance"], 0.01, 1e-8)
    r = collide(w, a, b, touch - 0.01 * e, margin=-0.01)  # 0.01 deep, margin -0.01: in contact
    assert r["num_contacts"] == 1 and abs(r["distance_lower_bound"]) < 1e-8 and close_pct(r["distance"], -0.01, 1e-8)
    r = collide(w, a, b, touch, margin=-0.01)  # touching, margin -0.01: free
    assert r["num_contacts"] == 0 and close_pct(r["distance_lower_bound"], 0.01, 1e-8)


def shape_shape(w, a, b, touch, tol, free_bound):
    """box_box :334-418 and test_shape_shape :420-488 (they differ in the expected bound of the fourth block)"""
    touch = np.array(touch, dtype=float)
    r = collide(w, a, b, touch)
    assert r["num_contacts"] == 1 and abs(r["distance_lower_bound"]) < tol and abs(r["distance"]) < 1e-8
    r = collide(w, a, b, touch + [0, 0, 0.01])
    assert r["num_contacts"] == 0 and close_pct(r["distance_lower_bound"], 0.01, tol)
    r = collide(w, a, b, touch, margin=0.01)
    assert r["num_contacts"] == 1 and close_pct(r["distance_lower_bound"], -0.01, tol) and abs(r["distance"]) < 1e-8
    r = collide(w, a, b, touch, margin=-0.01)
    assert r["num_contacts"] == 0 and close_pct(r["distance_lower_bound"], free_bound, tol)
    r = collide(w, a, b, touch + [0, -0.01, -0.01], margin=-0.01)
    assert r["num_contacts"] == 1 and abs(r["distance_lower_bound"]) < tol and close_pct(r["distance"], -0.01, tol)


def test_margin_box_box():  # :334-418
    w = World()
    shape_shape(w, w.shape("box", 1, 1, 1), w.shape("box", 1, 1, 1), (0, 1, 1), 1e-3, 0.01)


def test_margin_box_sphere():  # sphere_box :490-506: the box, and the box as the convex hull of its mesh
    w = World()
    s = w.shape("sphere", 0.5)
    touch = (0, 0, 1)
    bound = np.linalg.norm(-0.01 * np.array(touch, dtype=float))
    shape_shape(w, w.shape("box", 1, 1, 1), s, touch, 1e-6, bound)
    # generateBVHModel(box) + buildConvexRepresentation: 8 corners, 12 triangles
    # (geometric_shape_to_BVH_model.h:56-99)
    a, b, c = 0.5, 0.5, 0.5
    pts = np.array([[a, -b, c], [a, b, c], [-a, b, c], [-a, -b, c], [a, -b, -c], [a, b, -c], [-a, b, -c], [-a, -b, -c]])
    tris = np.array([[0, 4, 1], [1, 4, 5], [2, 6, 3], [3, 6, 7], [3, 0, 2], [2, 0, 1],
                     [6, 5, 7], [7, 5, 4], [1, 5, 2], [2, 5, 6], [3, 7, 0], [0, 7, 4]])
    cid = w.sc.register_convex(pts, tris)
    hull = int(w.sc.register_shapes(P.make_shapes([P.GEOM_CONVEX], [[0, 0, 0]], data=[cid]))[0])
    shape_shape(w, hull, s, touch, 1e-6, bound)


def test_box_box_collision_with_cached_guess_and_upper_bound():  # test/box_box_collision.cpp:20-47
    w = World()
    a, b = w.shape("box", 1, 1, 1), w.shape("box", 1, 1, 1)
    w.sc.commit()
    # enable_cached_gjk_guess = true (the guess of a fresh request: (1, 0, 0)), distance_upper_bound = 1e-6
    req = P.CollisionRequestPOD(gjk_initial_guess=P.CachedGuess, distance_upper_bound=1e-6)
    for T, expect in (((0, 0, 0), True), ((2, 0, 0), False)):
        ro = w.sc.b["oracle"].batch_collide([a], tf(T), [b], tf(), req)
        re = w.sc.b["emu"].batch_collide([a], tf(T), [b], tf(), req)
        assert ro.tobytes() == re.tobytes()
        assert (ro[0]["num_contacts"] > 0) == expect
