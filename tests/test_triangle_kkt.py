"""distance(TriangleP, TriangleP): the reference's optimality test (/root/reference/test/gjk.cpp:71-327), restated
in numpy.  For every pair -- the two literal ones of the reference (:95-132) first, then random ones with
coordinates in [-1, 1] -- the witness points must lie in the planes of their triangles (to 1e-7), inside them, and
satisfy the Karush-Kuhn-Tucker conditions of the closest-point problem with non-negative multipliers; a colliding
pair must come apart when triangle 2 is moved by (penetration depth + 6) along the normal (:169-198).

Run on the oracle; the host build of the device code must return the same bits.  CPU only.
"""
import numpy as np
import pytest

from tests.common import P, compare_distance, make_scenes
from hppfcl_b200 import workloads as W

EPS = 1e-7
N = 5000  # (the reference draws 10 000)

LITERAL = [  # gjk.cpp:95-132
    dict(P=[(0.063996093749999997, -0.15320971679687501, -0.42799999999999999),
            (0.069105957031249998, -0.150722900390625, -0.42999999999999999),
            (0.063996093749999997, -0.15320971679687501, -0.42999999999999999)],
         Q=[(-25.655000000000001, -1.2858199462890625, 3.7249809570312502),
            (-10.926, -1.284259033203125, 3.7281499023437501),
            (-10.926, -1.2866180419921875, 3.72335400390625)],
         # Quatf(w, x, y, z)
         R=W.quat_to_rot(-0.42437287410898855, -0.26862477561450587, -0.46249645019513175, 0.73064726592483387),
         T=(-12.824601270753471, -1.6840516940066426, 3.8914453043793844)),
    dict(P=[(-0.8027043342590332, -0.30276307463645935, -0.4372950792312622),
            (-0.8027043342590332, 0.30276307463645935, -0.4372950792312622),
            (0.8027043342590332, 0.30276307463645935, -0.4372950792312622)],
         Q=[(-0.224713996052742, -0.7417119741439819, 0.19999997317790985),
            (-0.5247139930725098, -0.7417119741439819, 0.19999997317790985),
            (-0.224713996052742, -0.7417119741439819, 0.09999997168779373)],
         R=np.array([[0.9657787025454787, 0.09400415350535746, 0.24173273843919627],
                     [-0.06713698817647556, 0.9908494114820345, -0.11709000206805695],
                     [-0.25052768814676646, 0.09685382227587608, 0.9632524147814993]]),
         T=(-0.13491177905469953, -1, 0.6000449621843792)),
]


def kkt_check(Pw, Qw, p1, p2, what):
    """gjk.cpp:200-305 for one pair: Pw, Qw world vertices (3x3 rows), p1, p2 the witness points"""
    u1, v1 = Pw[1] - Pw[0], Pw[2] - Pw[0]
    u2, v2 = Qw[1] - Qw[0], Qw[2] - Qw[0]
    w1, w2 = np.cross(u1, v1), np.cross(u2, v2)
    assert w1 @ w1 > EPS * EPS and w2 @ w2 > EPS * EPS
    a1 = np.linalg.solve(np.column_stack([u1, v1, w1]), p1 - Pw[0])
    a2 = np.linalg.solve(np.column_stack([u2, v2, w2]), p2 - Qw[0])

    def approx(a, b):  # EIGEN_VECTOR_IS_APPROX: isApprox(b, eps)
        return np.linalg.norm(a - b) <= EPS * min(np.linalg.norm(a), np.linalg.norm(b))
    assert approx(p1, Pw[0] + a1[0] * u1 + a1[1] * v1), what
    assert approx(p2, Qw[0] + a2[0] * u2 + a2[1] * v2), what
    assert abs(a1[2]) < EPS and abs(a2[2]) < EPS, what  # on the planes of the triangles
    d = p2 - p1
    grad_f = np.array([-d @ u1, -d @ v1, d @ u2, d @ v2])
    g = np.array([-a1[0], -a1[1], a1[0] + a1[1] - 1, -a2[0], -a2[1], a2[0] + a2[1] - 1])
    grad_g = np.zeros((4, 6))
    grad_g[0, 0] = grad_g[1, 1] = grad_g[2, 3] = grad_g[3, 4] = -1
    grad_g[0, 2] = grad_g[1, 2] = grad_g[2, 5] = grad_g[3, 5] = 1
    assert np.all(g <= EPS), (what, g)  # inside the triangles
    sat = np.abs(g) <= EPS
    if sat.any():
        c = np.linalg.lstsq(grad_g[:, sat], -grad_f, rcond=None)[0]
        assert np.all(c >= -EPS), (what, c)


@pytest.mark.parametrize("variant", [P.DefaultGJK, P.NesterovAcceleration])
def test_triangle_triangle_distance_is_optimal(variant):
    rng = np.random.default_rng(11 + variant)
    sc = make_scenes()  # (TriangleP operands are not in the reference's public distance table: no reference build here)
    Ps = rng.uniform(-1, 1, (N, 3, 3))
    Qs = rng.uniform(-1, 1, (N, 3, 3))
    R1 = np.eye(3)[None].repeat(N, 0)
    T1 = np.zeros((N, 3))
    for k, lit in enumerate(LITERAL):
        Ps[k], Qs[k], R1[k], T1[k] = lit["P"], lit["Q"], lit["R"], lit["T"]
    h1, h2 = [], []
    for k in range(N):
        for pts, out in ((Ps[k], h1), (Qs[k], h2)):
            cid = sc.register_convex(pts, None)
            out.append(int(sc.register_shapes(P.make_shapes([P.GEOM_TRIANGLE], [[0, 0, 0]], data=[cid]))[0]))
    sc.commit()
    t1 = P.make_transforms(R1, T1)
    t2 = W.identity_transforms(N)
    req = P.DistanceRequestPOD(gjk_variant=variant)
    ro = sc.b["oracle"].batch_distance(h1, t1, h2, t2, req)
    re = sc.b["emu"].batch_distance(h1, t1, h2, t2, req)
    compare_distance(ro, re, what="triangle pairs")
    hit = ro["min_distance"] <= 0
    assert 0 < hit.sum() < N // 2
    # :169-198 a colliding pair, triangle 2 moved by (depth + 10 - 4) along the normal, is free
    idx = np.nonzero(hit)[0]
    t2m = W.identity_transforms(len(idx))
    t2m["T"] = (-ro["min_distance"][idx] + 10 - 4)[:, None] * ro["normal"][idx]
    sub = lambda t: t[idx]
    rm = sc.b["oracle"].batch_distance([h1[i] for i in idx], sub(t1), [h2[i] for i in idx], t2m, req)
    compare_distance(rm, sc.b["emu"].batch_distance([h1[i] for i in idx], sub(t1), [h2[i] for i in idx], t2m, req),
                     what="moved triangle pairs")
    assert np.all(rm["min_distance"] > 0)
    for k in range(N):
        Pw = Ps[k] @ R1[k].T + T1[k]
        kkt_check(Pw, Qs[k], ro["p1"][k], ro["p2"][k], "pair %d" % k)
