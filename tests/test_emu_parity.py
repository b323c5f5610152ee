"""CPU-only parity: the per-pair DEVICE code (hpp-fcl_b200/csrc/*.cuh compiled with g++ by
tests/emu, one lane per pair) against the oracle on seeded batches.  This is what lets the
kernels' arithmetic be debugged without a GPU; the `-m gpu` tests repeat it on the real kernels.
Bar: status words, iteration counts, collide flags, warm-start outputs and every double
bit-identical.
"""
import numpy as np
import pytest

from tests.common import P, compare_distance, compare_hill_climb, make_scenes
from hppfcl_b200 import workloads as W

ALL_PRIMS = (P.GEOM_SPHERE, P.GEOM_CAPSULE, P.GEOM_BOX, P.GEOM_CYLINDER, P.GEOM_CONE, P.GEOM_ELLIPSOID)


@pytest.fixture(scope="module")
def prim():
    sc = make_scenes()
    w = W.config2_mixed_primitives(30000, pool=4096, types=ALL_PRIMS)
    sc.register_shapes(w["shapes"])
    return sc, w


@pytest.mark.parametrize("variant", [P.DefaultGJK, P.PolyakAcceleration, P.NesterovAcceleration])
@pytest.mark.parametrize("crit,ctype", [(P.Default, P.Relative), (P.DualityGap, P.Relative),
                                        (P.DualityGap, P.Absolute), (P.Hybrid, P.Relative), (P.Hybrid, P.Absolute)])
def test_distance_all_variants_and_criteria(prim, variant, crit, ctype):
    sc, w = prim
    req = P.DistanceRequestPOD(gjk_variant=variant, gjk_convergence_criterion=crit,
                               gjk_convergence_criterion_type=ctype)
    ro = sc.b["oracle"].batch_distance(w["h1"], w["tf1"], w["h2"], w["tf2"], req, nthreads=0)
    re = sc.b["emu"].batch_distance(w["h1"], w["tf1"], w["h2"], w["tf2"], req)
    compare_distance(ro, re, what="distance")
    assert (P.status_epa(ro["status"]) == P.EPA_AccuracyReached).sum() > 100


@pytest.mark.parametrize("kw", [dict(), dict(security_margin=0.05), dict(security_margin=-0.02),
                                dict(enable_contact=0, distance_upper_bound=0.0),
                                dict(enable_contact=0), dict(distance_upper_bound=0.3),
                                dict(gjk_max_iterations=4), dict(epa_max_iterations=3),
                                dict(gjk_tolerance=1e-3, epa_tolerance=1e-3)])
def test_collide_request_fields(prim, kw):
    sc, w = prim
    req = P.CollisionRequestPOD(**kw)
    ro, og, oh = sc.b["oracle"].batch_collide(w["h1"], w["tf1"], w["h2"], w["tf2"], req, want_guess=True, nthreads=0)
    re, eg, eh = sc.b["emu"].batch_collide(w["h1"], w["tf1"], w["h2"], w["tf2"], req, want_guess=True)
    compare_distance(ro, re, what="collide %s" % kw)
    assert np.array_equal(og.view(np.uint64), eg.view(np.uint64)) and np.array_equal(oh, eh)


def test_epa_two_tier_workspace(prim):
    """k_epa first runs a pair in a reduced-size polytope workspace and repeats it in the full-size one
    when the polytope outgrows it; a tight tolerance makes many curved pairs take that second path."""
    from tests.common import emu_lib
    L = emu_lib()
    L.emu_epa_retries.restype = __import__("ctypes").c_long
    sc, w = prim
    n = 6000
    req = P.DistanceRequestPOD(epa_tolerance=1e-12)
    before = L.emu_epa_retries()
    ro = sc.b["oracle"].batch_distance(w["h1"][:n], w["tf1"][:n], w["h2"][:n], w["tf2"][:n], req, nthreads=0)
    re = sc.b["emu"].batch_distance(w["h1"][:n], w["tf1"][:n], w["h2"][:n], w["tf2"][:n], req)
    compare_distance(ro, re, what="two-tier EPA")
    assert L.emu_epa_retries() - before > 20
    ep = ro["iterations"] >> 16
    assert ep.max() > 30  # runs past the reduced workspace's 24 iterations
    # tier 1 continues from the state tier 0 reached (epa_ws_grow + pair_phase2_resume); starting the pair over instead
    # must give the same bits, and so must the lane-group form (8 lanes per pair, as k_epa runs it)
    L.emu_epa_resumed.restype = __import__("ctypes").c_long
    assert L.emu_epa_resumed() > 20
    r0 = L.emu_epa_resumed()
    L.emu_set_epa_resume(0)
    try:
        re2 = sc.b["emu"].batch_distance(w["h1"][:n], w["tf1"][:n], w["h2"][:n], w["tf2"][:n], req)
    finally:
        L.emu_set_epa_resume(1)
    assert L.emu_epa_resumed() == r0
    compare_distance(ro, re2, what="two-tier EPA, tier 1 starting over")


def test_lane_group_epa_continues_in_the_full_workspace(prim):
    """the same hand-over as k_epa<8> does it: eight lanes copy the reduced workspace into the full one (every lane
    its share on the device, every lane thread its private copy here) and carry on"""
    from tests.common import emu_lib
    L = emu_lib()
    L.emu_epa_resumed.restype = __import__("ctypes").c_long
    sc, w = prim
    n = 2500
    req = P.DistanceRequestPOD(epa_tolerance=1e-12)
    emu = sc.b["emu"]
    before = L.emu_epa_resumed()
    emu.lanes = 8
    try:
        re = emu.batch_distance(w["h1"][:n], w["tf1"][:n], w["h2"][:n], w["tf2"][:n], req)
    finally:
        emu.lanes = 1
    ro = sc.b["oracle"].batch_distance(w["h1"][:n], w["tf1"][:n], w["h2"][:n], w["tf2"][:n], req, nthreads=0)
    compare_distance(ro, re, what="two-tier EPA, 8 lanes")
    assert L.emu_epa_resumed() - before > 5


def test_signed_distance_off(prim):
    sc, w = prim
    req = P.DistanceRequestPOD(enable_signed_distance=0)
    ro = sc.b["oracle"].batch_distance(w["h1"], w["tf1"], w["h2"], w["tf2"], req, nthreads=0)
    re = sc.b["emu"].batch_distance(w["h1"], w["tf1"], w["h2"], w["tf2"], req)
    compare_distance(ro, re, what="unsigned distance")
    col = P.status_gjk(ro["status"]) == P.GJK_Collision
    assert col.sum() > 100 and np.all(np.isnan(ro["p1"][col]))  # narrowphase.h:638-656


def test_initial_guess_modes(prim):
    sc, w = prim
    n = 8000
    a = [w["h1"][:n], w["tf1"][:n], w["h2"][:n], w["tf2"][:n]]
    ro, og, oh = sc.b["oracle"].batch_distance(*a, want_guess=True)
    req = P.DistanceRequestPOD(gjk_initial_guess=P.CachedGuess)
    req.q.cached_gjk_guess = og.ctypes.data
    req.q.cached_support_func_guess = oh.ctypes.data
    r2 = sc.b["oracle"].batch_distance(*a, req)
    e2 = sc.b["emu"].batch_distance(*a, req)
    compare_distance(r2, e2, what="cached guess")
    # warm-started GJK needs no more iterations than the cold start on average (README claim)
    g = P.status_path(ro["status"]) == P.PATH_GJK
    assert (r2["iterations"][g] & 0xffff).mean() < (ro["iterations"][g] & 0xffff).mean()
    req = P.DistanceRequestPOD(gjk_initial_guess=P.BoundingVolumeGuess)
    compare_distance(sc.b["oracle"].batch_distance(*a, req), sc.b["emu"].batch_distance(*a, req), what="bv guess")


def test_swept_sphere_radius(prim):
    """test/swept_sphere_radius.cpp: inflating a shape by ssr shifts the distance by ssr."""
    sc = make_scenes()
    rng = np.random.default_rng(4)
    base = W.random_primitive_shapes(rng, 64, ALL_PRIMS)
    infl = base.copy()
    infl["ssr"] = 0.07
    hb, hi = sc.register_shapes(base), sc.register_shapes(infl)
    n = 5000
    i1, i2 = rng.integers(0, 64, n), rng.integers(0, 64, n)
    t1 = W.identity_transforms(n)
    t2 = W.random_transforms(rng, n, (2.5, -1, -1), (4, 1, 1))  # separated
    ro = sc.b["oracle"].batch_distance(hi[i1], t1, hi[i2], t2, nthreads=0)
    re = sc.b["emu"].batch_distance(hi[i1], t1, hi[i2], t2)
    compare_distance(ro, re, what="ssr")
    r0 = sc.b["oracle"].batch_distance(hb[i1], t1, hb[i2], t2, nthreads=0)
    assert np.allclose(ro["min_distance"], r0["min_distance"] - 0.14, atol=2e-6)


def test_triangles_and_unsupported():
    sc = make_scenes()
    rng = np.random.default_rng(8)
    tris = []
    for _ in range(32):
        cid = sc.register_convex(rng.normal(size=(3, 3)) * 0.5, None)
        tris.append(cid)
    ht = sc.register_shapes(P.make_shapes([P.GEOM_TRIANGLE] * 32, np.zeros((32, 3)), data=tris))
    hp = sc.register_shapes(W.random_primitive_shapes(rng, 64, ALL_PRIMS))
    hx = sc.register_shapes(P.make_shapes([18, 20], np.zeros((2, 3))))  # GEOM_OCTREE, HF_AABB (collision_object.h:65-89)
    allh = np.concatenate([ht, hp])
    n = 20000
    h1, h2 = allh[rng.integers(0, len(allh), n)], allh[rng.integers(0, len(allh), n)]
    t1 = W.random_transforms(rng, n, (0, 0, 0), (0, 0, 0))
    t2 = W.random_transforms(rng, n, (-1.5, -1.5, -1.5), (1.5, 1.5, 1.5))
    # TriangleP-TriangleP ignores the request's variant / criterion (triangle_triangle.cpp:67 only resets
    # the GJK object), every other pair of the batch honours them
    for req in (P.DistanceRequestPOD(), P.DistanceRequestPOD(gjk_variant=P.NesterovAcceleration),
                P.DistanceRequestPOD(gjk_variant=P.PolyakAcceleration, gjk_convergence_criterion=P.DualityGap,
                                     gjk_convergence_criterion_type=P.Absolute)):
        ro = sc.b["oracle"].batch_distance(h1, t1, h2, t2, req, nthreads=0)
        re = sc.b["emu"].batch_distance(h1, t1, h2, t2, req)
        compare_distance(ro, re, what="triangles")
    rc = sc.b["oracle"].batch_collide(h1, t1, h2, t2, nthreads=0)
    compare_distance(rc, sc.b["emu"].batch_collide(h1, t1, h2, t2), what="triangles collide")
    assert rc["num_contacts"].sum() > 500
    # node types outside the path are reported per pair as unsupported (collision.cpp:110-117)
    h1[:10] = hx[0]
    h2[10:20] = hx[1]
    ro = sc.b["oracle"].batch_distance(h1[:40], t1[:40], h2[:40], t2[:40])
    re = sc.b["emu"].batch_distance(h1[:40], t1[:40], h2[:40], t2[:40])
    compare_distance(ro, re, what="unsupported")
    assert np.all(P.status_path(ro["status"][:20]) == P.PATH_UNSUPPORTED)


def _convex(faithful, n=20000, pool=48):
    sc = make_scenes()
    w = W.config3_convex_pairs(n, pool=pool, nv=64)
    cids = [sc.register_convex(p, t if faithful else None) for p, t in w["hulls"]]
    rng = np.random.default_rng(5)
    small = [W.icosahedron_from_ellipsoid(0.1 + rng.random(3)) for _ in range(16)]
    cids += [sc.register_convex(p, t) for p, t in small]
    hc = sc.register_shapes(P.make_shapes([P.GEOM_CONVEX] * len(cids), np.zeros((len(cids), 3)), data=cids))
    hp = sc.register_shapes(W.random_primitive_shapes(np.random.default_rng(3), 128, ALL_PRIMS))
    return sc, w, hc, hp


@pytest.mark.parametrize("variant", [P.DefaultGJK, P.NesterovAcceleration, P.PolyakAcceleration])
def test_convex_linear_scan_bit_exact(variant):
    """ConvexBase support = linear scan (support_functions.cpp:401-421): 12-vertex hulls use it in
    the reference; the 64-vertex hulls are registered in the oracle without neighbours so that it
    runs the same exhaustive argmax as the kernels."""
    sc, w, hc, hp = _convex(False)
    h1, h2 = hc[w["h1"] % len(hc)], hc[w["h2"] % len(hc)]
    req = P.CollisionRequestPOD(gjk_variant=variant)
    ro, og, oh = sc.b["oracle"].batch_collide(h1, w["tf1"], h2, w["tf2"], req, want_guess=True, nthreads=0)
    re, eg, eh = sc.b["emu"].batch_collide(h1, w["tf1"], h2, w["tf2"], req, want_guess=True)
    compare_distance(ro, re, what="convex")
    assert np.array_equal(oh, eh) and np.array_equal(og.view(np.uint64), eg.view(np.uint64))
    assert 0.3 < ro["num_contacts"].mean() < 0.7
    rng = np.random.default_rng(1)
    allh = np.concatenate([hc, hp])
    n = len(h1)
    h1, h2 = allh[rng.integers(0, len(allh), n)], allh[rng.integers(0, len(allh), n)]
    dreq = P.DistanceRequestPOD(gjk_variant=variant)
    compare_distance(sc.b["oracle"].batch_distance(h1, w["tf1"], h2, w["tf2"], dreq, nthreads=0),
                     sc.b["emu"].batch_distance(h1, w["tf1"], h2, w["tf2"], dreq), what="convex/primitive")


@pytest.mark.parametrize("variant", [P.DefaultGJK, P.NesterovAcceleration])
def test_convex_vs_reference_hill_climb(variant):
    """Against the reference's >32-vertex hill-climbing support (support_functions.cpp:324-397,
    with visited flags, warm starts and hint carry-over): the exhaustive argmax can pick a
    different vertex only when the maximum is tied to rounding, so flags/statuses are exact and
    polytope-polytope distances, witness points and normals agree to ~1e-12."""
    sc, w, hc, hp = _convex(True)
    h1, h2 = hc[w["h1"] % len(hc)], hc[w["h2"] % len(hc)]
    req = P.CollisionRequestPOD(gjk_variant=variant)
    ro = sc.b["oracle"].batch_collide(h1, w["tf1"], h2, w["tf2"], req, nthreads=0)
    re = sc.b["emu"].batch_collide(h1, w["tf1"], h2, w["tf2"], req)
    compare_hill_climb(ro, re)
    # the hill-climb and the linear scan agree on every fresh query (no state carried over)
    rng = np.random.default_rng(2)
    ids = rng.integers(0, len(w["hulls"]), 50000).astype(np.uint32)
    dirs = rng.normal(size=(50000, 3))
    li, _ = sc.b["oracle"].batch_convex_support(ids, dirs)
    assert np.array_equal(li, sc.b["oracle"].batch_convex_support(ids, dirs, log=True))
    ei, es = sc.b["emu"].batch_convex_support(ids, dirs)
    assert np.array_equal(li, ei)


def test_convex_vs_curved_primitive_hill_climb_tolerance():
    """64-vertex hull vs curved primitive against the hill-climbing oracle: near-tie support
    choices change the GJK/EPA path, and on curved shapes two valid solutions of a tol=1e-6 solve
    differ by ~tol in distance (and ~sqrt(tol) in witness points).  Bar kept: statuses and
    collide flags exact away from |d| < tol, distances within 1e-6 absolute."""
    sc, w, hc, hp = _convex(True, n=12000)
    rng = np.random.default_rng(3)
    n = len(w["h1"])
    h1, h2 = hc[rng.integers(0, 48, n)], hp[rng.integers(0, len(hp), n)]
    ro = sc.b["oracle"].batch_collide(h1, w["tf1"], h2, w["tf2"], nthreads=0)
    re = sc.b["emu"].batch_collide(h1, w["tf1"], h2, w["tf2"])
    m = np.abs(ro["distance"]) > 1e-5
    assert np.array_equal(ro["num_contacts"][m], re["num_contacts"][m])
    assert np.array_equal(P.status_gjk(ro["status"])[m], P.status_gjk(re["status"])[m])
    ok = ~np.isnan(ro["p1"][:, 0]) & ~np.isnan(re["p1"][:, 0])
    assert np.max(np.abs(ro["distance"][ok] - re["distance"][ok])) < 2e-6


def test_edge_cases():
    sc = make_scenes()
    hs = sc.register_shapes(W.random_primitive_shapes(np.random.default_rng(0), 8, ALL_PRIMS))
    e = sc.b["emu"]
    t = W.identity_transforms(4)
    # empty batch
    assert e.batch_distance(hs[:0], t[:0], hs[:0], t[:0]).shape == (0,)
    # identical poses (deep penetration, identity relative transform path, minkowski_difference.cpp:281)
    r = sc.b["oracle"].batch_distance(hs[:4], t, hs[:4], t)
    compare_distance(r, e.batch_distance(hs[:4], t, hs[:4], t), what="coincident")
    assert np.all(r["min_distance"] < 0)
    # num_max_contacts == 0 -> invalid argument; security_margin == -inf -> cleared result
    with pytest.raises(ValueError):
        e.batch_collide(hs[:4], t, hs[:4], t, P.CollisionRequestPOD(num_max_contacts=0))
    with pytest.raises(ValueError):
        sc.b["oracle"].batch_collide(hs[:4], t, hs[:4], t, P.CollisionRequestPOD(num_max_contacts=0))
    req = P.CollisionRequestPOD(security_margin=-np.inf)
    r = sc.b["oracle"].batch_collide(hs[:4], t, hs[:4], t, req)
    g = e.batch_collide(hs[:4], t, hs[:4], t, req)
    assert r["num_contacts"].sum() == 0 and g["num_contacts"].sum() == 0
    assert np.all(np.isnan(g["p1"])) and np.all(g["distance_lower_bound"] == P.DBL_MAX)
    # invalid tolerance (gjk.cpp:62) and oversize EPA budget
    with pytest.raises(ValueError):
        e.batch_distance(hs[:4], t, hs[:4], t, P.DistanceRequestPOD(gjk_tolerance=0.0))
    with pytest.raises(ValueError):
        e.batch_distance(hs[:4], t, hs[:4], t, P.DistanceRequestPOD(epa_max_iterations=1000))


def test_lane_group_argmax_equals_the_serial_scan_for_any_direction():
    """The support argmax of a ConvexBase is split over the G lanes that own a pair.  Every lane must end with
    the serial scan's vertex (start from vertex 0, move on strictly greater: support_functions.cpp:401-421)
    for ANY direction -- GJK feeds it NaN now and then (0/0 in the projection of a degenerate simplex) and
    carries on, and with NaN dots the serial scan keeps vertex 0.  Found by the GPU fuzz: the lanes of a group
    used to disagree there.  tests/emu simulates the lane group on the host (LaneSim)."""
    import ctypes as C
    from tests.common import emu_lib
    L = emu_lib()
    L.emu_lane_group_support.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(5)
    special = [np.nan, np.inf, -np.inf, 0.0, -0.0, 1.0, -1.0]

    def serial(pts, d):
        best, bi = None, 0
        for i, p in enumerate(pts):
            with np.errstate(all="ignore"):
                v = (p[0] * d[0] + p[1] * d[1]) + p[2] * d[2]
            if i == 0 or v > best:
                best, bi = v, i
        return bi

    checked = nan_cases = 0
    for trial in range(1500):
        nv = int(rng.choice([1, 2, 3, 5, 8, 20, 33, 64]))
        pts = rng.normal(size=(nv, 3))
        kind = trial % 5
        if kind == 1:  # ties: duplicated and lattice points
            pts = np.round(pts)
        elif kind == 2:  # zeros in the coordinates: inf * 0 = NaN for some vertices only
            pts[rng.random((nv, 3)) < 0.4] = 0.0
        d = rng.normal(size=3)
        if kind >= 2:
            for k in range(3):
                if rng.random() < 0.5:
                    d[k] = special[rng.integers(0, len(special))]
        pts = np.ascontiguousarray(pts)
        d = np.ascontiguousarray(d)
        want = serial(pts, d)
        nan_cases += bool(np.isnan(d).any() or np.isinf(d).any())
        for G in (1, 2, 4, 8, 16, 32):
            out = np.full(G, -99, dtype=np.int32)
            assert L.emu_lane_group_support(G, pts.ctypes.data, nv, d.ctypes.data, out.ctypes.data) == 0
            assert np.all(out == want), "G=%d nv=%d dir=%s: lanes %s, serial scan %d" % (G, nv, d, out, want)
            checked += 1
    assert checked == 9000 and nan_cases > 300


@pytest.mark.parametrize("lanes", [2, 4, 8])
def test_pairs_through_lane_groups(lanes):
    """k_pairs<G> and k_epa<G> run a pair on G cooperating lanes.  tests/emu runs the same device code with the G
    lanes of a group as G host threads meeting at barriers where the device shuffles or syncs (lanesim; in EPA
    every lane works in a private copy of the polytope workspace and fetches the faces its peers computed): every
    lane must reach the same outcome, and the results must be the oracle's bit for bit -- hulls with few and many
    vertices, triangles, primitives, with requests that drive GJK through NaN rays (relative duality gap on
    exact-zero poses at scale 1000, the configuration in which the GPU fuzz found the lanes disagreeing)."""
    from tests.common import EmuScene
    from oracle import oracle_lib
    import itertools
    rng = np.random.default_rng(11)
    orc, emu = oracle_lib.OracleScene(P), EmuScene()
    emu.lanes = lanes
    scale = 1000.0
    hulls = [W.ellipsoid_hull(rng, nv) for nv in (6, 20, 32)] + [W.icosahedron_from_ellipsoid((0.3, 0.5, 0.2))]
    cube = np.array(list(itertools.product((-1.0, 1.0), repeat=3))) * 0.4
    ids = []
    for pts in [h[0] for h in hulls] + [cube]:
        a, b = orc.register_convex(pts * scale, None), emu.register_convex(pts * scale)
        assert a == b
        ids.append(a)
    prim = W.random_primitive_shapes(rng, 64, ALL_PRIMS)
    prim["p"] *= scale
    recs = np.concatenate([P.make_shapes([P.GEOM_CONVEX] * len(ids), np.zeros((len(ids), 3)), data=ids), prim])
    h = orc.register_shapes(recs)
    assert np.array_equal(h, emu.register_shapes(recs))
    n = 6000
    h1, h2 = h[rng.integers(0, len(h), n)], h[rng.integers(0, len(h), n)]
    h1[: n // 2] = h[rng.integers(0, len(ids), n // 2)]  # a hull in at least half of the pairs
    tf1 = W.identity_transforms(n)
    tf2 = W.identity_transforms(n, rng.uniform(-2, 2, (n, 3)) * (rng.random((n, 3)) < 0.5) * scale)
    tf2[n // 2:] = W.random_transforms(rng, n - n // 2, (-2 * scale,) * 3, (2 * scale,) * 3)
    gg = rng.normal(size=(n, 3)) * scale
    gg[rng.random(n) < 0.3] = 0
    gh = np.zeros((n, 2), dtype=np.int32)
    nan_rays = 0
    for kw in (dict(), dict(gjk_convergence_criterion=P.DualityGap),
               dict(gjk_variant=P.PolyakAcceleration, gjk_convergence_criterion=P.DualityGap),
               dict(gjk_variant=P.NesterovAcceleration, gjk_convergence_criterion=P.Hybrid)):
        req = P.DistanceRequestPOD(gjk_initial_guess=P.CachedGuess, **kw)
        req.q.cached_gjk_guess = gg.ctypes.data
        req.q.cached_support_func_guess = gh.ctypes.data
        ro, og, oh = orc.batch_distance(h1, tf1, h2, tf2, req, want_guess=True, nthreads=0)
        re, eg, eh = emu.batch_distance(h1, tf1, h2, tf2, req, want_guess=True)
        compare_distance(ro, re, what="lanes=%d %s" % (lanes, kw))
        same = (og.view(np.uint64) == eg.view(np.uint64)) | (np.isnan(og) & np.isnan(eg))  # NaN payloads aside
        assert same.all() and np.array_equal(oh, eh)
        nan_rays += int(np.isnan(og).any(axis=1).sum())
        creq = P.CollisionRequestPOD(security_margin=0.01 * scale, **kw)
        compare_distance(orc.batch_collide(h1, tf1, h2, tf2, creq, nthreads=0), emu.batch_collide(h1, tf1, h2, tf2, creq),
                         what="collide lanes=%d %s" % (lanes, kw))
    assert nan_rays > 0  # the batch does contain pairs whose GJK ends on a NaN ray
