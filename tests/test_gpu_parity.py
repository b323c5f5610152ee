"""GPU parity tests: the CUDA path (through the C-ABI) against the oracle on the
same seeded inputs.  Bar (BASELINE.json north_star): collide flags, status words,
iteration counts and indices bit-exact; distances / witness points / normals
within 1e-6 relative -- and, because oracle and kernels execute the same IEEE
operation sequence (no FMA), the doubles are in fact required to be bit-identical.
"""
import numpy as np
import pytest

from tests.common import P, compare_distance, compare_hill_climb, hf, make_scenes
from hppfcl_b200 import workloads as W

pytestmark = pytest.mark.gpu

ALL_PRIMS = (P.GEOM_SPHERE, P.GEOM_CAPSULE, P.GEOM_BOX, P.GEOM_CYLINDER, P.GEOM_CONE, P.GEOM_ELLIPSOID)


@pytest.fixture(scope="module")
def prim_scene():
    sc = make_scenes(gpu=True, emu=False)
    w = W.config2_mixed_primitives(200_000, pool=8192, types=ALL_PRIMS)
    sc.register_shapes(w["shapes"])
    sc.commit()
    return sc, w


@pytest.mark.parametrize("variant", [P.DefaultGJK, P.NesterovAcceleration, P.PolyakAcceleration])
def test_distance_primitives_all_types(prim_scene, variant):
    sc, w = prim_scene
    req = P.DistanceRequestPOD(gjk_variant=variant)
    ref = sc.b["oracle"].batch_distance(w["h1"], w["tf1"], w["h2"], w["tf2"], req, nthreads=0)
    got = sc.b["gpu"].batch_distance(w["h1"], w["tf1"], w["h2"], w["tf2"], req)
    compare_distance(ref, got, what="distance variant %d" % variant)
    assert (ref["min_distance"] < 0).sum() > 100  # the penetrating (EPA) branch is exercised


@pytest.mark.parametrize("nsub", ["1", "3"])
def test_epa_two_tier_and_parts(nsub, monkeypatch):
    """Tight EPA tolerance: many pairs outgrow the reduced-size workspace of k_epa's first tier and are
    repeated by the second; HFB_NSUB=3 cuts phase 1 in parts whose EPA runs on the side stream."""
    monkeypatch.setenv("HFB_NSUB", nsub)
    sc = make_scenes(gpu=True, emu=False)
    w = W.config2_mixed_primitives(150_000, pool=4096, types=ALL_PRIMS)
    sc.register_shapes(w["shapes"])
    sc.commit()
    req = P.DistanceRequestPOD(epa_tolerance=1e-12)
    ref = sc.b["oracle"].batch_distance(w["h1"], w["tf1"], w["h2"], w["tf2"], req, nthreads=0)
    got = sc.b["gpu"].batch_distance(w["h1"], w["tf1"], w["h2"], w["tf2"], req)
    compare_distance(ref, got, what="two-tier EPA nsub=%s" % nsub)
    assert ((ref["iterations"] >> 16) > 24).sum() > 200


@pytest.mark.parametrize("margin", [0.0, 0.05, -0.02])
def test_collide_primitives(prim_scene, margin):
    sc, w = prim_scene
    req = P.CollisionRequestPOD(security_margin=margin)
    ref = sc.b["oracle"].batch_collide(w["h1"], w["tf1"], w["h2"], w["tf2"], req, nthreads=0)
    got = sc.b["gpu"].batch_collide(w["h1"], w["tf1"], w["h2"], w["tf2"], req)
    compare_distance(ref, got, what="collide margin %g" % margin)
    assert ref["num_contacts"].sum() > 100


def test_collide_no_contact_info_early_stop(prim_scene):
    """enable_contact=False + distance_upper_bound=0: GJK early exit (gjk.cpp:288-293), NaN witness."""
    sc, w = prim_scene
    req = P.CollisionRequestPOD(enable_contact=0, distance_upper_bound=0.0)
    ref = sc.b["oracle"].batch_collide(w["h1"], w["tf1"], w["h2"], w["tf2"], req, nthreads=0)
    got = sc.b["gpu"].batch_collide(w["h1"], w["tf1"], w["h2"], w["tf2"], req)
    compare_distance(ref, got, what="collide early-stop")
    assert (P.status_gjk(ref["status"]) == P.GJK_NoCollisionEarlyStopped).sum() > 100


def test_cached_guess_roundtrip(prim_scene):
    sc, w = prim_scene
    n = 20000
    sl = slice(0, n)
    a = [w["h1"][sl], w["tf1"][sl], w["h2"][sl], w["tf2"][sl]]
    ref, rg, rh = sc.b["oracle"].batch_distance(*a, want_guess=True)
    got, gg, gh = sc.b["gpu"].batch_distance(*a, want_guess=True)
    compare_distance(ref, got, what="guess pass 1")
    assert np.array_equal(rg.view(np.uint64), gg.view(np.uint64)) and np.array_equal(rh, gh)
    req = P.DistanceRequestPOD(gjk_initial_guess=P.CachedGuess)
    req.q.cached_gjk_guess = rg.ctypes.data
    req.q.cached_support_func_guess = rh.ctypes.data
    ref2 = sc.b["oracle"].batch_distance(*a, req)
    got2 = sc.b["gpu"].batch_distance(*a, req)
    compare_distance(ref2, got2, what="guess pass 2")


def _convex_scene(faithful):
    """faithful=True: the oracle uses the reference's neighbour hill-climb for hulls with more
    than 32 vertices (support_functions.cpp:324-397); faithful=False: hulls are registered in the
    oracle without neighbours, so it runs the linear scan (:401-421) the kernels implement."""
    sc = make_scenes(gpu=True, emu=False)
    w = W.config3_convex_pairs(60_000, pool=96, nv=64)
    cids = [sc.register_convex(p, t if faithful else None) for p, t in w["hulls"]]
    rng = np.random.default_rng(5)
    small = [W.icosahedron_from_ellipsoid(0.1 + rng.random(3)) for _ in range(16)]
    cids += [sc.register_convex(p, t) for p, t in small]
    hc = sc.register_shapes(P.make_shapes([P.GEOM_CONVEX] * len(cids), np.zeros((len(cids), 3)), data=cids))
    prims = W.random_primitive_shapes(np.random.default_rng(3), 256, ALL_PRIMS)
    hp = sc.register_shapes(prims)
    sc.commit()
    return sc, w, hc, hp


@pytest.fixture(scope="module")
def convex_scene():
    return _convex_scene(False)


@pytest.fixture(scope="module")
def convex_scene_faithful():
    return _convex_scene(True)


@pytest.mark.parametrize("variant", [P.DefaultGJK, P.NesterovAcceleration])
def test_convex_convex(convex_scene, variant):
    sc, w, hc, hp = convex_scene
    h1, h2 = hc[w["h1"] % len(hc)], hc[w["h2"] % len(hc)]
    req = P.CollisionRequestPOD(gjk_variant=variant)
    ref = sc.b["oracle"].batch_collide(h1, w["tf1"], h2, w["tf2"], req, nthreads=0)
    got = sc.b["gpu"].batch_collide(h1, w["tf1"], h2, w["tf2"], req)
    compare_distance(ref, got, what="convex collide variant %d" % variant)
    assert 0.2 < ref["num_contacts"].mean() < 0.8
    dreq = P.DistanceRequestPOD(gjk_variant=variant)
    ref, rg, rh = sc.b["oracle"].batch_distance(h1, w["tf1"], h2, w["tf2"], dreq, want_guess=True, nthreads=0)
    got, gg, gh = sc.b["gpu"].batch_distance(h1, w["tf1"], h2, w["tf2"], dreq, want_guess=True)
    compare_distance(ref, got, what="convex distance variant %d" % variant)
    assert np.array_equal(rh, gh), "support hints (witness vertex indices) differ"


@pytest.mark.parametrize("variant", [P.DefaultGJK, P.NesterovAcceleration])
def test_convex_convex_vs_hill_climb_reference(convex_scene_faithful, variant):
    """64-vertex hulls against the reference's hill-climbing support: the support vertex can
    differ only when the maximum is tied to rounding (direction normal to a face/edge, i.e. at
    GJK convergence), so collide flags and statuses are bit-exact and distances / witness
    points / normals agree far inside the 1e-6 bar for polytope-polytope pairs."""
    sc, w, hc, hp = convex_scene_faithful
    h1, h2 = hc[w["h1"] % len(hc)], hc[w["h2"] % len(hc)]
    req = P.CollisionRequestPOD(gjk_variant=variant)
    ref = sc.b["oracle"].batch_collide(h1, w["tf1"], h2, w["tf2"], req, nthreads=0)
    got = sc.b["gpu"].batch_collide(h1, w["tf1"], h2, w["tf2"], req)
    compare_hill_climb(ref, got)


def test_convex_vs_primitives_mixed(convex_scene):
    sc, w, hc, hp = convex_scene
    rng = np.random.default_rng(11)
    n = 40000
    allh = np.concatenate([hc, hp])
    h1 = allh[rng.integers(0, len(allh), n)]
    h2 = allh[rng.integers(0, len(allh), n)]
    ref = sc.b["oracle"].batch_distance(h1, w["tf1"][:n], h2, w["tf2"][:n], nthreads=0)
    got = sc.b["gpu"].batch_distance(h1, w["tf1"][:n], h2, w["tf2"][:n])
    compare_distance(ref, got, what="mixed convex/primitive")


@pytest.mark.parametrize("g", [1, 2, 4, 8, 16, 32])
def test_lane_group_sizes(g, monkeypatch):
    """every instantiated lanes-per-pair setting (HFB_GC / HFB_GE) gives the same bits"""
    monkeypatch.setenv("HFB_GC", str(g))
    monkeypatch.setenv("HFB_GE", str(max(g, 4)))  # the EPA kernel comes with 4, 8, 16, 32
    sc, w, hc, hp = _convex_scene(False)
    n = 20000
    rng = np.random.default_rng(g)
    allh = np.concatenate([hc, hp])
    h1, h2 = allh[rng.integers(0, len(allh), n)], allh[rng.integers(0, len(allh), n)]
    ref = sc.b["oracle"].batch_collide(h1, w["tf1"][:n], h2, w["tf2"][:n], nthreads=0)
    got = sc.b["gpu"].batch_collide(h1, w["tf1"][:n], h2, w["tf2"][:n])
    compare_distance(ref, got, what="G=%d" % g)


def test_reference_known_answers_on_gpu():
    """the reference's own literal known-answer cases (tests/test_oracle_golden.py: box_box_distance.cpp,
    gjk.cpp, capsule_*.cpp, geometric_shapes.cpp, security_margin.cpp ...) with the CUDA path as the
    backend whose numbers are asserted"""
    import os
    import subprocess
    import sys
    from tests.common import ROOT
    env = dict(os.environ, HFB_GOLDEN_BACKEND="gpu")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_oracle_golden.py"), "-q",
                        "-x", "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout


@pytest.mark.parametrize("knob", ["HFB_STAGE", "HFB_REFILL"])
def test_optional_kernel_paths(knob, monkeypatch):
    """the two measured-and-kept-optional paths give the same bits: cp.async.bulk (TMA) staging of the
    hulls' vertex blocks into shared memory, and the lane-refill GJK kernel for primitive pairs"""
    monkeypatch.setenv(knob, "1")
    sc, w, hc, hp = _convex_scene(False)
    n = 40000
    rng = np.random.default_rng(11)
    allh = np.concatenate([hc, hp])
    h1, h2 = allh[rng.integers(0, len(allh), n)], allh[rng.integers(0, len(allh), n)]
    for req in (P.DistanceRequestPOD(), P.DistanceRequestPOD(gjk_variant=P.NesterovAcceleration)):
        ref = sc.b["oracle"].batch_distance(h1, w["tf1"][:n], h2, w["tf2"][:n], req, nthreads=0)
        got = sc.b["gpu"].batch_distance(h1, w["tf1"][:n], h2, w["tf2"][:n], req)
        compare_distance(ref, got, what=knob)


def test_convex_support_kernel(convex_scene):
    sc, w, hc, hp = convex_scene
    rng = np.random.default_rng(2)
    n = 50000
    ids = rng.integers(0, len(w["hulls"]) + 16, n).astype(np.uint32)
    dirs = rng.normal(size=(n, 3))
    dirs[:100] = np.eye(3)[rng.integers(0, 3, 100)]  # axis-aligned directions (tie-prone)
    ri, rs = sc.b["oracle"].batch_convex_support(ids, dirs)
    gi, gs = sc.b["gpu"].batch_convex_support(ids, dirs)
    assert np.array_equal(ri, gi)
    assert np.array_equal(rs.view(np.uint64), gs.view(np.uint64))


def test_errors_and_edge_cases(prim_scene):
    sc, w = prim_scene
    eng = sc.b["gpu"]
    # empty batch
    out = eng.batch_distance(w["h1"][:0], w["tf1"][:0], w["h2"][:0], w["tf2"][:0])
    assert out.shape == (0,)
    # num_max_contacts == 0 -> invalid argument (collision.cpp:82-85)
    with pytest.raises(hf.EngineError):
        eng.batch_collide(w["h1"][:4], w["tf1"][:4], w["h2"][:4], w["tf2"][:4], P.CollisionRequestPOD(num_max_contacts=0))
    # bad handle
    bad = w["h1"][:4].copy()
    bad[1] = 0x7fffffff
    with pytest.raises(hf.EngineError):
        eng.batch_distance(bad, w["tf1"][:4], w["h2"][:4], w["tf2"][:4])
    # security_margin = -inf -> cleared results, no contacts (collision.cpp:73-76)
    got = eng.batch_collide(w["h1"][:64], w["tf1"][:64], w["h2"][:64], w["tf2"][:64],
                            P.CollisionRequestPOD(security_margin=-np.inf))
    assert got["num_contacts"].sum() == 0 and np.all(np.isnan(got["p1"]))
    # single pair (n = 1)
    ref = sc.b["oracle"].batch_distance(w["h1"][:1], w["tf1"][:1], w["h2"][:1], w["tf2"][:1])
    got = eng.batch_distance(w["h1"][:1], w["tf1"][:1], w["h2"][:1], w["tf2"][:1])
    compare_distance(ref, got, what="n=1")


def test_geometry_clear_and_reuse():
    """hfb_geom_clear drops the arena; queries need a new registration + commit, handles restart at 0"""
    eng = hf.Engine(0)
    h = eng.register_shapes(P.make_shapes([P.GEOM_SPHERE, P.GEOM_BOX], [[0.5, 0, 0], [0.5, 0.5, 0.5]]))
    eng.commit()
    tf = W.identity_transforms(1)
    tf2 = P.make_transforms(np.eye(3)[None], np.array([[3.0, 0, 0]]))
    r = eng.batch_distance(h[:1], tf, h[1:], tf2)
    assert abs(r[0]["min_distance"] - 2.0) < 1e-12
    eng.clear_geometry()
    with pytest.raises(hf.EngineError):
        eng.batch_distance(h[:1], tf, h[1:], tf2)  # not committed any more
    h2 = eng.register_shapes(P.make_shapes([P.GEOM_SPHERE, P.GEOM_SPHERE], [[1.0, 0, 0], [0.25, 0, 0]]))
    assert list(h2) == [0, 1]
    eng.commit()
    r = eng.batch_distance(h2[:1], tf, h2[1:], tf2)
    assert abs(r[0]["min_distance"] - 1.75) < 1e-12


def test_full_size_properties():
    """1M-pair batch (BASELINE config 2 size): size-independent properties instead of the oracle:
    p2 = p1 + d*n, |n| = 1, swapping the operands mirrors the result
    (test/normal_and_nearest_points.cpp:74-241)."""
    eng = hf.Engine(0)
    w = W.config2_mixed_primitives(1_000_000)
    eng.register_shapes(w["shapes"])
    eng.commit()
    r = eng.batch_distance(w["h1"], w["tf1"], w["h2"], w["tf2"])
    ok = ~np.isnan(r["p1"][:, 0])
    assert ok.mean() > 0.999
    d = r["min_distance"][ok, None]
    assert np.allclose(r["p1"][ok] + d * r["normal"][ok], r["p2"][ok], atol=1e-6)
    assert np.allclose(np.linalg.norm(r["normal"][ok], axis=1), 1.0, atol=1e-9)
    # a 64k slice against the oracle (all host threads)
    from oracle import oracle_lib
    orc = oracle_lib.OracleScene(P)
    orc.register_shapes(w["shapes"])
    s = slice(500_000, 565_536)
    ref = orc.batch_distance(w["h1"][s], w["tf1"][s], w["h2"][s], w["tf2"][s], nthreads=0)
    compare_distance(ref, r[s], what="1M slice")
    # reversed operands: same distance within GJK tolerance, flipped normal
    r2 = eng.batch_distance(w["h2"][s], w["tf2"][s], w["h1"][s], w["tf1"][s])
    m = ~np.isnan(ref["p1"][:, 0]) & ~np.isnan(r2["p1"][:, 0])
    assert np.allclose(ref["min_distance"][m], r2["min_distance"][m], atol=5e-4)


def test_object_table_calls_equal_the_pair_calls():
    """hfb_batch_{distance,collide}_objects: an object table (handle + pose) and index pairs, expanded on the device.
    The records must equal, bit for bit, those of the pair calls on the expanded rows -- shape pairs, convex pairs and
    mesh pairs alike -- and the compact result modes must be projections of the full records."""
    import torch
    rng = np.random.default_rng(11)
    eng = hf.Engine(0)
    w = W.config2_mixed_primitives(1000, pool=512, types=ALL_PRIMS, seed=3)
    hp = eng.register_shapes(w["shapes"])
    pts, _ = W.ellipsoid_hull(rng, 40)
    cid = eng.register_convex(pts)
    hc = eng.register_shapes(P.make_shapes([P.GEOM_CONVEX], [[0, 0, 0]], data=[cid]))
    verts, tris = W.sphere_mesh(0.6, 12, 6, noise=0.02, rng=rng)
    bid = eng.register_bvh_obbrss(None, verts, tris)
    hm = eng.register_shapes(P.make_shapes([P.BV_OBBRSS], [[0, 0, 0]], data=[bid]))
    eng.commit()
    n_obj, n_pairs = 20_000, 300_000
    allh = np.concatenate([hp, np.repeat(hc, 40), np.repeat(hm, 10)])
    oh = allh[rng.integers(0, len(allh), n_obj)].astype(np.uint32)
    otf = W.random_transforms(rng, n_obj, (-2, -2, -2), (2, 2, 2))
    pi = rng.integers(0, n_obj, n_pairs).astype(np.uint32)
    pj = rng.integers(0, n_obj, n_pairs).astype(np.uint32)
    full = eng.batch_distance(oh[pi], otf[pi], oh[pj], otf[pj])
    obj = eng.batch_distance_objects(oh, otf, pi, pj)
    assert full.tobytes() == obj.tobytes()
    dmin = eng.batch_distance_objects(oh, otf, pi, pj, min_only=True)
    assert dmin.tobytes() == np.ascontiguousarray(full["min_distance"]).tobytes()
    req = P.CollisionRequestPOD(security_margin=0.01)
    cfull = eng.batch_collide(oh[pi], otf[pi], oh[pj], otf[pj], req)
    cobj = eng.batch_collide_objects(oh, otf, pi, pj, req)
    assert cfull.tobytes() == cobj.tobytes()
    hits = np.nonzero(cfull["num_contacts"] > 0)[0]
    assert len(hits) > 500
    flags, nh, ids, recs = eng.batch_collide_objects(oh, otf, pi, pj, req, compact_capacity=len(hits) + 7)
    bits = np.unpackbits(flags.view(np.uint8), bitorder="little")[:n_pairs].astype(bool)
    assert np.array_equal(np.nonzero(bits)[0], hits) and nh == len(hits)
    order = np.argsort(ids)
    assert np.array_equal(ids[order], hits) and recs[order].tobytes() == cfull[hits].tobytes()
    # capacity below the number of colliding pairs: all counted, `capacity` kept
    _, nh2, ids2, recs2 = eng.batch_collide_objects(oh, otf, pi, pj, req, compact_capacity=100)
    assert nh2 == len(hits) and len(ids2) == 100 and np.all(np.isin(ids2, hits))
    assert all(recs2[k].tobytes() == cfull[ids2[k]].tobytes() for k in range(100))
    # device-resident scene
    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).cuda()
    d = [dev(oh), dev(otf), dev(pi), dev(pj)]
    d_out = torch.empty(n_pairs * P.distance_result_dtype.itemsize, dtype=torch.uint8, device="cuda")
    eng.batch_distance_objects_device(n_obj, d[0].data_ptr(), d[1].data_ptr(), n_pairs, d[2].data_ptr(), d[3].data_ptr(),
                                      d_out.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert d_out.cpu().numpy().tobytes() == full.tobytes()
    d_c = torch.empty(n_pairs * P.contact_dtype.itemsize, dtype=torch.uint8, device="cuda")
    eng.batch_collide_objects_device(n_obj, d[0].data_ptr(), d[1].data_ptr(), n_pairs, d[2].data_ptr(), d[3].data_ptr(),
                                     d_c.data_ptr(), req, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert d_c.cpu().numpy().tobytes() == cfull.tobytes()
    # an index past the table: invalid argument on the host path
    bad = pi.copy()
    bad[5] = n_obj
    with pytest.raises(hf.EngineError):
        eng.batch_distance_objects(oh, otf, bad, pj)
    assert eng.stats()["watchdog_trips"] == 0


@pytest.mark.gpu
def test_host_pipeline_with_one_epa_pass_equals_the_chunk_by_chunk_one():
    """large host batches run phase 1 chunk by chunk and EPA once over the whole batch, the EPA pairs' records
    following compacted (host_batch_pipelined); HFB_HOST_PIPE=0 keeps the chunk-by-chunk pipeline.  Same bits, for
    rows, distances only, collide, object tables -- with many small chunks -- and the oracle's on a sample."""
    import os
    from oracle import oracle_lib
    rng = np.random.default_rng(23)
    n = 200_000
    w = W.config2_mixed_primitives(n, pool=2048, types=ALL_PRIMS, seed=5)
    pts, _ = W.ellipsoid_hull(rng, 24)
    n_obj = 5000
    pi, pj = rng.integers(0, n_obj, n).astype(np.uint32), rng.integers(0, n_obj, n).astype(np.uint32)
    otf = W.random_transforms(rng, n_obj, (-1.5, -1.5, -1.5), (1.5, 1.5, 1.5))
    out = {}
    for pipe in ("1", "0"):
        old = {k: os.environ.get(k) for k in ("HFB_HOST_PIPE", "HFB_CHUNK")}
        os.environ["HFB_HOST_PIPE"], os.environ["HFB_CHUNK"] = ("2" if pipe == "1" else "0"), "24576"
        try:
            eng = hf.Engine(0)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        hp = eng.register_shapes(w["shapes"])
        hc = eng.register_shapes(P.make_shapes([P.GEOM_CONVEX], [[0, 0, 0]], data=[eng.register_convex(pts)]))
        eng.commit()
        h1, h2 = hp[w["h1"] % len(hp)].copy(), hp[w["h2"] % len(hp)].copy()
        h1[::17] = hc[0]  # hull pairs in every chunk
        oh = np.concatenate([hp, hc])[rng.integers(0, len(hp) + 1, n_obj)].astype(np.uint32) if pipe == "1" else out["oh"]
        res = dict(oh=oh,
                   rows=eng.batch_distance(h1, w["tf1"], h2, w["tf2"]),
                   col=eng.batch_collide(h1, w["tf1"], h2, w["tf2"], P.CollisionRequestPOD(security_margin=0.02)),
                   obj=eng.batch_distance_objects(oh, otf, pi, pj),
                   omin=eng.batch_distance_objects(oh, otf, pi, pj, min_only=True),
                   launches=eng.stats()["kernel_launches"])
        if pipe == "1":
            out = res
            first = (h1, h2)
        else:
            for k in ("rows", "col", "obj", "omin"):
                assert out[k].tobytes() == res[k].tobytes(), k
            assert out["launches"] < res["launches"]  # one EPA pass instead of one per chunk
    epa = (out["rows"]["iterations"] >> 16) > 0
    assert epa.sum() > 1000 and (out["col"]["num_contacts"] > 0).sum() > 5000
    # a bad handle / object index in a late chunk fails the call (it is looked at while the GPU already works on the
    # chunk: the device side treats such a pair as unsupported, nothing faults)
    bad_h = first[0].copy()
    bad_h[-5] = 0xfffffff0
    with pytest.raises(hf.EngineError):
        eng.batch_distance(bad_h, w["tf1"], first[1], w["tf2"])
    bad_i = pi.copy()
    bad_i[n // 2] = n_obj
    with pytest.raises(hf.EngineError):
        eng.batch_distance_objects(out["oh"], otf, bad_i, pj)
    assert eng.batch_distance(first[0], w["tf1"], first[1], w["tf2"]).tobytes() == out["rows"].tobytes()  # and the context lives on
    assert out["omin"].tobytes() == np.ascontiguousarray(out["obj"]["min_distance"]).tobytes()
    orc = oracle_lib.OracleScene(P)
    orc.register_shapes(w["shapes"])  # (same handle numbering as the engines: the primitives, then the hull)
    orc.register_shapes(P.make_shapes([P.GEOM_CONVEX], [[0, 0, 0]], data=[orc.register_convex(pts, None)]))
    m = 30_000
    sel = np.sort(rng.choice(n, m, replace=False))
    compare_distance(orc.batch_distance(first[0][sel], w["tf1"][sel], first[1][sel], w["tf2"][sel], nthreads=0), out["rows"][sel],
                     what="pipelined host batch vs oracle")
