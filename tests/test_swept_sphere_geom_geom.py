"""Swept-sphere radii through collide(): the reference's `ssr_geom_geom` (/root/reference/test/swept_sphere_radius.cpp:
279-383 with the shape draws of test/utility.cpp:556-650).  For every ordered pair of {box, sphere, ellipsoid,
capsule, cone, cylinder, convex hull, plane, halfspace} (no two of the plane family, as there) and radii
{0, 0.1, 1, 10} x {0, 0.1, 1, 10}: with the security margin at infinity every pair is "in collision", so the contact
always carries witness points; the contact of the swept shapes must be the contact of the bare shapes moved by the
radii -- depth - (r1 + r2), the same normal, p1 + r1 n, p2 - r2 n -- to 3 sqrt(tol) + max(r1, r2) / 100.

Checked on the oracle; the host build of the device code and, where oracle/_ref exists, the reference build must
return the same bits.  CPU only.
"""
import numpy as np
import pytest

from tests.common import P, compare_distance, make_scenes, ref_agrees
from hppfcl_b200 import workloads as W

RADII = (0.0, 0.1, 1.0, 10.0)
KINDS = ("box", "sphere", "ellipsoid", "capsule", "cone", "cylinder", "convex", "plane", "halfspace")
POSES = 25  # (1 in the reference)
TOL = 1e-6


def draw(sc, rng, kind):
    """-> handles of the same random geometry with each of RADII as its swept-sphere radius (utility.cpp:607-650)"""
    u = rng.uniform
    if kind in ("plane", "halfspace"):
        n = u(-1, 1, 3)
        nd = np.array([[*(n / np.linalg.norm(n)), u(0.1, 1.0)]])
        t = P.GEOM_PLANE if kind == "plane" else P.GEOM_HALFSPACE
        return [int(sc.register_halfspaces(t, nd, np.array([r]))[0]) for r in RADII]
    data = 0
    if kind == "box":
        t, p = P.GEOM_BOX, u(0.1, 1.0, 3) / 2
    elif kind == "sphere":
        t, p = P.GEOM_SPHERE, [u(0.1, 1.0), 0, 0]
    elif kind == "ellipsoid":
        t, p = P.GEOM_ELLIPSOID, u(0.1, 1.0, 3)
    elif kind == "convex":
        pts, tris = W.icosahedron_from_ellipsoid(tuple(u(0.1, 1.0, 3)))
        t, p, data = P.GEOM_CONVEX, [0, 0, 0], sc.register_convex(pts, tris)
    else:
        t = {"capsule": P.GEOM_CAPSULE, "cone": P.GEOM_CONE, "cylinder": P.GEOM_CYLINDER}[kind]
        p = [u(0.1, 0.8), u(0.2, 1.0) / 2, 0]
    return [int(sc.register_shapes(P.make_shapes([t], [p], ssr=[r], data=[data]))[0]) for r in RADII]


@pytest.mark.parametrize("k1", KINDS)
def test_swept_sphere_radius_through_collide(k1):
    rng = np.random.default_rng(300 + KINDS.index(k1))
    sc = make_scenes(ref=True)
    rows = []  # (bare 1, bare 2, swept 1, swept 2, r1, r2)
    for k2 in KINDS:
        if k1 in ("plane", "halfspace") and k2 in ("plane", "halfspace"):
            continue
        a, b = draw(sc, rng, k1), draw(sc, rng, k2)
        for i, r1 in enumerate(RADII):
            for j, r2 in enumerate(RADII):
                rows += [(a[0], b[0], a[i], b[j], r1, r2)] * POSES
    sc.commit()
    rows = np.array(rows)
    n = len(rows)
    t1 = W.random_transforms(rng, n, (-2, -2, -2), (2, 2, 2))
    t2 = W.random_transforms(rng, n, (-2, -2, -2), (2, 2, 2))
    req = P.CollisionRequestPOD(enable_contact=1, security_margin=P.DBL_MAX, gjk_tolerance=TOL, epa_tolerance=TOL)
    out = []
    for c1, c2 in ((0, 1), (2, 3)):
        h1, h2 = rows[:, c1].astype(np.uint32), rows[:, c2].astype(np.uint32)
        ro = sc.b["oracle"].batch_collide(h1, t1, h2, t2, req)
        re = sc.b["emu"].batch_collide(h1, t1, h2, t2, req)
        compare_distance(ro, re, what="%s-* %s" % (k1, "swept" if c1 else "bare"))
        ref_agrees(sc, "batch_collide", ro, (h1, t1, h2, t2, req), "%s-* %s" % (k1, "swept" if c1 else "bare"))
        out.append(ro)
    bare, swept = out
    r1, r2 = rows[:, 4], rows[:, 5]
    assert np.all(bare["num_contacts"] == 1) and np.all(swept["num_contacts"] == 1)
    prec = 3 * np.sqrt(TOL) + np.maximum(r1, r2) / 100
    assert np.all(np.abs(bare["distance"] - (r1 + r2) - swept["distance"]) < prec)
    dots = np.einsum("ij,ij->i", bare["normal"], swept["normal"])
    assert np.all(dots > 0) and np.all(np.abs(1 - dots) < prec)
    for f, r, sgn in (("p1", r1, 1.0), ("p2", r2, -1.0)):
        moved = bare[f] + sgn * r[:, None] * bare["normal"]
        assert np.all(np.abs(moved - swept[f]) <= prec[:, None]), f  # isZero(tol): every coefficient
