"""Seeded synthetic workloads for the BASELINE.json configs.

Pure numpy/scipy generators; they produce POD arrays (see _pod.py) that are fed
identically to the CUDA engine, to the oracle and to the CPU emulation harness.
Generators mirror the reference's test helpers:
  generateRandomTransforms  test/utility.cpp:198-247
  constructPolytopeFromEllipsoid / random hulls  test/utility.cpp:501-557
"""
import numpy as np

from . import _pod as P


def random_rotations(rng, n):
    """Uniform unit quaternions -> rotation matrices (n,3,3) (math/transform.h:228-250)."""
    u1, u2, u3 = rng.random(n), rng.random(n), rng.random(n)
    m1, m2 = np.sqrt(1.0 - u1), np.sqrt(u1)
    w, x = m1 * np.sin(2 * np.pi * u2), m1 * np.cos(2 * np.pi * u2)
    y, z = m2 * np.sin(2 * np.pi * u3), m2 * np.cos(2 * np.pi * u3)
    return quat_to_rot(w, x, y, z)


def quat_to_rot(w, x, y, z):
    w, x, y, z = [np.asarray(a, dtype=np.float64) for a in (w, x, y, z)]
    R = np.empty(w.shape + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z)
    R[..., 0, 1] = 2 * (x * y - z * w)
    R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w)
    R[..., 1, 1] = 1 - 2 * (x * x + z * z)
    R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w)
    R[..., 2, 1] = 2 * (y * z + x * w)
    R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def random_transforms(rng, n, lo, hi):
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    T = lo + (hi - lo) * rng.random((n, 3))
    return P.make_transforms(random_rotations(rng, n), T)


def identity_transforms(n, T=None):
    R = np.broadcast_to(np.eye(3), (n, 3, 3))
    return P.make_transforms(R, np.zeros((n, 3)) if T is None else T)


def random_primitive_shapes(rng, n, types):
    """n shape records with type drawn uniformly from `types`; sizes as in
    test/normal_and_nearest_points.cpp:283-… (U[0.05,1] radii / half sides)."""
    types = np.asarray(types, dtype=np.uint32)
    t = types[rng.integers(0, len(types), n)]
    p = 0.05 + 0.95 * rng.random((n, 3))
    out = P.make_shapes(t, p)
    # capsule/cylinder/cone: p[0] = radius, p[1] = halfLength in [0.075, 0.5]
    lz = 0.15 + 0.85 * rng.random(n)
    m = (t == P.GEOM_CAPSULE) | (t == P.GEOM_CYLINDER) | (t == P.GEOM_CONE)
    out["p"][m, 1] = lz[m] / 2
    out["p"][m, 2] = 0.0
    ms = t == P.GEOM_SPHERE
    out["p"][ms, 1:] = 0.0
    return out


def config2_mixed_primitives(n_pairs, seed=0xFC1 + 2, pool=65536,
                             types=(P.GEOM_SPHERE, P.GEOM_CAPSULE, P.GEOM_BOX, P.GEOM_CYLINDER)):
    """BASELINE config 2: mixed primitive pairs (sphere/capsule/box/cylinder), random poses;
    relative translation U[-3,3]^2 x U[0,3] as test/accelerated_gjk.cpp:122."""
    rng = np.random.default_rng(seed)
    shapes = random_primitive_shapes(rng, pool, types)
    h1 = rng.integers(0, pool, n_pairs).astype(np.uint32)
    h2 = rng.integers(0, pool, n_pairs).astype(np.uint32)
    tf1 = random_transforms(rng, n_pairs, (0, 0, 0), (0, 0, 0))
    tf2 = random_transforms(rng, n_pairs, (-3, -3, 0), (3, 3, 3))
    # move the whole scene so poses are not origin-centred
    off = -5 + 10 * rng.random((n_pairs, 3))
    tf1["T"] += off
    tf2["T"] += off
    return dict(shapes=shapes, h1=h1, tf1=tf1, h2=h2, tf2=tf2)


def config2_scene(n_objects=100_000, n_pairs=1_000_000, seed=0xFC1 + 2, pool=65536,
                  types=(P.GEOM_SPHERE, P.GEOM_CAPSULE, P.GEOM_BOX, P.GEOM_CYLINDER)):
    """BASELINE config 2 as a SCENE: n_objects mixed primitives (sphere/capsule/box/cylinder, sizes as in
    config2_mixed_primitives) with random poses in a cube, and the n_pairs object pairs a broadphase would hand to the
    narrow phase -- every ordered pair (i, j) whose relative translation lies in [-3,3]^2 x [0,3], the box
    test/accelerated_gjk.cpp:122 draws relative translations from, so the pairs have config 2's statistics (uniform
    relative translation in that box, independent uniform rotations) while 1 M pairs share 100 k objects, as pairs do
    in a real scene.  -> shapes (the pool), obj_h (pool index per object), obj_tf, first, second."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    shapes = random_primitive_shapes(rng, pool, types)
    obj_h = rng.integers(0, pool, n_objects).astype(np.uint32)
    # expected pairs = N^2 / 2 * 6^3 / side^3; aim 8 % above n_pairs, then cut
    side = (0.5 * n_objects * n_objects * 216.0 / (1.08 * n_pairs)) ** (1.0 / 3.0)
    for _ in range(6):
        tf = random_transforms(rng, n_objects, (0, 0, 0), (side, side, side))
        pr = cKDTree(tf["T"]).query_pairs(3.0, p=np.inf, output_type="ndarray")
        if len(pr) >= n_pairs:
            break
        side *= 0.96
    assert len(pr) >= n_pairs, "scene too sparse"
    pr = pr[rng.permutation(len(pr))[:n_pairs]]
    dz = tf["T"][pr[:, 1], 2] - tf["T"][pr[:, 0], 2]
    first = np.where(dz >= 0, pr[:, 0], pr[:, 1]).astype(np.uint32)
    second = np.where(dz >= 0, pr[:, 1], pr[:, 0]).astype(np.uint32)
    return dict(shapes=shapes, obj_h=obj_h, obj_tf=tf, first=first, second=second)


def ellipsoid_hull(rng, nv=64, radii=None):
    """nv points on a random ellipsoid (all are hull vertices) + triangulated hull faces."""
    from scipy.spatial import ConvexHull
    if radii is None:
        radii = 0.05 + 0.95 * rng.random(3)
    while True:
        v = rng.normal(size=(nv, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        pts = v * radii
        hull = ConvexHull(pts)
        if len(hull.vertices) == nv:
            break
    tris = hull.simplices.astype(np.uint32)
    # orient faces outward (fillNeighbors only needs adjacency, convex.hxx:231-280)
    c = pts.mean(axis=0)
    a, b, cc = pts[tris[:, 0]], pts[tris[:, 1]], pts[tris[:, 2]]
    flip = np.einsum("ij,ij->i", np.cross(b - a, cc - a), a - c) < 0
    tris[flip] = tris[flip][:, [0, 2, 1]]
    return pts, tris


def config3_convex_pairs(n_pairs, seed=0xFC1 + 3, pool=1024, nv=64):
    """BASELINE config 3: ConvexBase(nv) x ConvexBase(nv); ~half of the pairs overlap."""
    rng = np.random.default_rng(seed)
    hulls = [ellipsoid_hull(rng, nv) for _ in range(pool)]
    h1 = rng.integers(0, pool, n_pairs).astype(np.uint32)
    h2 = rng.integers(0, pool, n_pairs).astype(np.uint32)
    tf1 = random_transforms(rng, n_pairs, (0, 0, 0), (0, 0, 0))
    tf2 = random_transforms(rng, n_pairs, (-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
    off = -5 + 10 * rng.random((n_pairs, 3))
    tf1["T"] += off
    tf2["T"] += off
    return dict(hulls=hulls, h1=h1, tf1=tf1, h2=h2, tf2=tf2)


def icosahedron_from_ellipsoid(radii):
    """constructPolytopeFromEllipsoid (test/utility.cpp:486-557): 12 vertices, 20 faces.
    toEllipsoid: point /= ||point||, then scaled per axis by the radii."""
    phi = (1 + np.sqrt(5)) / 2
    pts = np.array([
        [-1, phi, 0], [1, phi, 0], [-1, -phi, 0], [1, -phi, 0],
        [0, -1, phi], [0, 1, phi], [0, -1, -phi], [0, 1, -phi],
        [phi, 0, -1], [phi, 0, 1], [-phi, 0, -1], [-phi, 0, 1]], dtype=np.float64)
    pts /= np.linalg.norm(pts, axis=1, keepdims=True)
    pts *= np.asarray(radii, dtype=np.float64)
    tris = np.array([
        [0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4],
        [11, 10, 2], [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8],
        [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.uint32)
    return pts, tris


def sphere_mesh(radius=1.0, seg=100, ring=50, noise=0.0, rng=None):
    """generateBVHModel(Sphere, seg, ring) (shape/geometric_shape_to_BVH_model.h:100-150):
    2*seg*ring triangles (seg=100, ring=50 -> 10 000); optional radial noise breaks symmetric ties."""
    pts, tris = [], []
    phid, thetad = 2 * np.pi / seg, np.pi / (ring + 1)
    for i in range(ring):
        theta = thetad * (i + 1)
        for j in range(seg):
            phi = phid * j
            pts.append([radius * np.sin(theta) * np.cos(phi), radius * np.sin(theta) * np.sin(phi),
                        radius * np.cos(theta)])
    pts.append([0, 0, radius])
    pts.append([0, 0, -radius])
    for i in range(ring - 1):
        for j in range(seg):
            a = i * seg + j
            b = i * seg if j == seg - 1 else i * seg + j + 1
            c = (i + 1) * seg + j
            d = (i + 1) * seg if j == seg - 1 else (i + 1) * seg + j + 1
            tris.append([a, c, b])
            tris.append([b, c, d])
    for j in range(seg):
        a, b, c = j, (0 if j == seg - 1 else j + 1), ring * seg
        tris.append([c, a, b])
        a = (ring - 1) * seg + j
        b = (ring - 1) * seg if j == seg - 1 else (ring - 1) * seg + j + 1
        c = ring * seg + 1
        tris.append([a, c, b])
    pts = np.array(pts, dtype=np.float64)
    if noise and rng is not None:
        pts *= 1 + noise * rng.standard_normal((len(pts), 1))
    return pts, np.array(tris, dtype=np.uint32)


def config4_mesh_vs_capsules(n_queries, seed=0xFC1 + 4, seg=100, ring=50, pool=4096):
    """BASELINE config 4: 10k-triangle OBBRSS mesh vs capsules, distance + nearest points."""
    rng = np.random.default_rng(seed)
    verts, tris = sphere_mesh(1.0, seg, ring, noise=0.01, rng=rng)
    caps = P.make_shapes([P.GEOM_CAPSULE] * pool,
                         np.stack([0.02 + 0.08 * rng.random(pool), (0.05 + 0.25 * rng.random(pool)) / 2,
                                   np.zeros(pool)], axis=1))
    hc = rng.integers(0, pool, n_queries).astype(np.uint32)
    tf_mesh = random_transforms(rng, n_queries, (-0.2, -0.2, -0.2), (0.2, 0.2, 0.2))
    tf_caps = random_transforms(rng, n_queries, (-2, -2, -2), (2, 2, 2))
    return dict(verts=verts, tris=tris, capsules=caps, hc=hc, tf_mesh=tf_mesh, tf_caps=tf_caps)


def config5_moving_boxes(n_objects=100_000, env_scale=None, seed=0xFC1 + 5, target_pairs=1_000_000):
    """BASELINE config 5: n Box(5, 10, 20) objects (generateEnvironments, test/utility.cpp:392-404) with random poses in
    [-env_scale, env_scale]^3; `step(k)` moves every object by a small random delta (generateRandomTransforms with
    delta_trans / delta_rot, utility.cpp:249-...).  env_scale defaults to the value that gives about `target_pairs`
    overlapping AABB pairs (the reference's 2000 for 1000 objects would give ~50 pairs for 100 k)."""
    rng = np.random.default_rng(seed)
    shapes = P.make_shapes([P.GEOM_BOX], [[2.5, 5.0, 10.0]])  # half sides of Box(5, 10, 20)
    if env_scale is None:
        # a rotated box's AABB has a mean side of about 16.5; two cubes of side s overlap with probability (2 s / L)^3
        s = 16.5
        env_scale = 0.5 * 2 * s / (2.0 * target_pairs / (float(n_objects) ** 2)) ** (1.0 / 3.0)
    tf = random_transforms(rng, n_objects, (-env_scale,) * 3, (env_scale,) * 3)
    obj_h = np.zeros(n_objects, dtype=np.uint32)

    def step(k, delta_trans=1.0, delta_rot=0.05):
        """poses of step k: the base poses moved by a seeded random delta"""
        r = np.random.default_rng(seed + 7919 * (k + 1))
        out = tf.copy()
        out["T"] += delta_trans * (2 * r.random((n_objects, 3)) - 1)
        # small rotation about a random axis, applied on the left: R' = dR * R
        ax = r.standard_normal((n_objects, 3))
        ax /= np.linalg.norm(ax, axis=1, keepdims=True)
        ang = delta_rot * (2 * r.random(n_objects) - 1)
        K = np.zeros((n_objects, 3, 3))
        K[:, 0, 1], K[:, 0, 2], K[:, 1, 0] = -ax[:, 2], ax[:, 1], ax[:, 2]
        K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -ax[:, 0], -ax[:, 1], ax[:, 0]
        dR = np.eye(3)[None] + np.sin(ang)[:, None, None] * K + (1 - np.cos(ang))[:, None, None] * (K @ K)
        R = out["R"].reshape(n_objects, 3, 3).transpose(0, 2, 1)  # stored column-major
        Rn = dR @ R
        out["R"] = Rn.transpose(0, 2, 1).reshape(out["R"].shape)
        return out

    return dict(shapes=shapes, obj_h=obj_h, obj_tf=tf, env_scale=env_scale, step=step)
