"""Host-side mirror of the hpp-fcl query interface for the accelerated path.

Same names, argument meaning and error behaviour as the reference's Python /
C++ surface for this path (python/collision.cc, python/distance.cc;
include/hpp/fcl/{collision,distance,collision_data}.h):

    collide(o1, tf1, o2, tf2, request, result) -> number of contacts
    distance(o1, tf1, o2, tf2, request, result) -> min distance
    ComputeCollision / ComputeDistance functors (collision.h:79-117)
    BatchQuery: the batched form (many (o1,tf1,o2,tf2) tuples, one launch)

Every call runs on the GPU through the C-ABI (engine.Engine); there is no CPU
path.  Single-pair calls are batches of one.
"""
import numpy as np

from . import _pod as P
from .engine import Engine

_NAN3 = np.full(3, np.nan)


# ---------------------------------------------------------------- geometry ---
class Transform3f:
    """math/transform.h:56-216 (rotation matrix + translation)."""

    def __init__(self, R=None, T=None):
        self.R = np.eye(3) if R is None else np.asarray(R, dtype=np.float64).reshape(3, 3)
        self.T = np.zeros(3) if T is None else np.asarray(T, dtype=np.float64).reshape(3)

    @staticmethod
    def from_quat(w, x, y, z, T=None):
        from .workloads import quat_to_rot
        return Transform3f(quat_to_rot(w, x, y, z), T)

    def transform(self, v):
        return self.R @ np.asarray(v, dtype=np.float64) + self.T

    def __mul__(self, o):  # transform.h:186-188
        return Transform3f(self.R @ o.R, self.R @ o.T + self.T)

    def getRotation(self):
        return self.R

    def getTranslation(self):
        return self.T

    def setTranslation(self, T):
        self.T = np.asarray(T, dtype=np.float64).reshape(3)

    def pod(self):
        return P.make_transforms(self.R[None], self.T[None])


class CollisionGeometry:
    node_type = None

    def __init__(self):
        self._ssr = 0.0

    def setSweptSphereRadius(self, r):  # geometric_shapes.h:77-85
        if r < 0:
            raise ValueError("Swept-sphere radius must be positive.")
        self._ssr = float(r)

    def getSweptSphereRadius(self):
        return self._ssr

    def getNodeType(self):
        return self.node_type

    def _params(self):
        return (0.0, 0.0, 0.0)

    def _points(self):
        return None


class Box(CollisionGeometry):
    node_type = P.GEOM_BOX

    def __init__(self, x, y=None, z=None):  # full side lengths; halfSide = side / 2 (geometric_shapes.h:164-187)
        super().__init__()
        side = np.asarray(x if y is None else (x, y, z), dtype=np.float64)
        self.halfSide = side / 2

    def _params(self):
        return tuple(self.halfSide)


class Sphere(CollisionGeometry):
    node_type = P.GEOM_SPHERE

    def __init__(self, radius):
        super().__init__()
        self.radius = float(radius)

    def _params(self):
        return (self.radius, 0.0, 0.0)


class Ellipsoid(CollisionGeometry):
    node_type = P.GEOM_ELLIPSOID

    def __init__(self, rx, ry=None, rz=None):
        super().__init__()
        self.radii = np.asarray(rx if ry is None else (rx, ry, rz), dtype=np.float64)

    def _params(self):
        return tuple(self.radii)


class _RadiusLength(CollisionGeometry):
    def __init__(self, radius, lz):  # halfLength = lz / 2 (geometric_shapes.h:386-400)
        super().__init__()
        self.radius = float(radius)
        self.halfLength = float(lz) / 2

    def _params(self):
        return (self.radius, self.halfLength, 0.0)


class Capsule(_RadiusLength):
    node_type = P.GEOM_CAPSULE


class Cone(_RadiusLength):
    node_type = P.GEOM_CONE


class Cylinder(_RadiusLength):
    node_type = P.GEOM_CYLINDER


class _PlaneBase(CollisionGeometry):
    """n . x <= d (Halfspace) or n . x = d (Plane); the constructor normalises (n, d) (unitNormalTest,
    src/shape/geometric_shapes.cpp:121-143)."""

    def __init__(self, a, b=None, c=None, d=None):  # (n, d) or (a, b, c, d): geometric_shapes.h:887-896, 979-988
        super().__init__()
        if c is None:
            n, off = np.asarray(a, dtype=np.float64).reshape(3).copy(), float(0.0 if b is None else b)
        else:
            n, off = np.array([a, b, c], dtype=np.float64), float(d)
        l = float(np.sqrt((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]))
        if l > 0:
            inv_l = 1.0 / l
            n, off = n * inv_l, off * inv_l
        else:
            n, off = np.array([1.0, 0.0, 0.0]), 0.0
        self.n, self.d = n, off

    def _params(self):
        return tuple(self.n)

    def _plane(self):
        return np.array([self.n[0], self.n[1], self.n[2], self.d], dtype=np.float64)


class Halfspace(_PlaneBase):  # geometric_shapes.h:885-961
    node_type = P.GEOM_HALFSPACE

    def signedDistance(self, p):
        return float(np.dot(self.n, p) - (self.d + self._ssr))


class Plane(_PlaneBase):  # geometric_shapes.h:977-1050
    node_type = P.GEOM_PLANE


class Convex(CollisionGeometry):
    """ConvexBase / Convex<Triangle> (geometric_shapes.h:638-872, shape/convex.h)."""
    node_type = P.GEOM_CONVEX

    def __init__(self, points, triangles=None):
        super().__init__()
        self.points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        self.num_points = self.points.shape[0]
        self.triangles = None if triangles is None else np.ascontiguousarray(triangles, dtype=np.uint32)

    def _points(self):
        return self.points


class TriangleP(CollisionGeometry):
    node_type = P.GEOM_TRIANGLE

    def __init__(self, a, b, c):
        super().__init__()
        self.a, self.b, self.c = [np.asarray(v, dtype=np.float64).reshape(3) for v in (a, b, c)]

    def _points(self):
        return np.stack([self.a, self.b, self.c])


class BVHModelOBBRSS(CollisionGeometry):
    """BVHModel<OBBRSS> (include/hpp/fcl/BVH/BVH_model.h) with the reference's build protocol; the tree is
    built on the host by the library (hfb_bvh_build_obbrss, mean split), the queries walk it on the GPU."""
    node_type = P.BV_OBBRSS

    def __init__(self):
        super().__init__()
        self._verts, self._tris = [], []
        self.vertices = self.tri_indices = self.bvs = None
        self._begun = False

    def beginModel(self, num_tris=0, num_vertices=0):
        self._verts, self._tris = [], []
        self.vertices = self.tri_indices = self.bvs = None
        self._begun = True
        return 0

    def _need_begun(self):
        if not self._begun:
            raise ValueError("BVH construction does not follow correct sequence (beginModel first)")

    def addVertex(self, p):
        self._need_begun()
        self._verts.append(np.asarray(p, dtype=np.float64).reshape(3))
        return 0

    def addTriangle(self, p1, p2, p3):  # BVH_model.cpp:359-414: three new vertices
        self._need_begun()
        o = len(self._verts)
        for p in (p1, p2, p3):
            self.addVertex(p)
        self._tris.append((o, o + 1, o + 2))
        return 0

    def addSubModel(self, points, triangles):  # :470-538
        self._need_begun()
        o = len(self._verts)
        for p in np.asarray(points, dtype=np.float64).reshape(-1, 3):
            self._verts.append(p)
        for t in np.asarray(triangles, dtype=np.int64).reshape(-1, 3):
            self._tris.append((int(t[0]) + o, int(t[1]) + o, int(t[2]) + o))
        return 0

    def endModel(self):
        self._need_begun()
        if not self._tris:
            raise ValueError("BVH construction error: empty model")
        from .engine import build_bvh_obbrss
        self.vertices = np.ascontiguousarray(np.stack(self._verts), dtype=np.float64)
        self.tri_indices = np.ascontiguousarray(np.array(self._tris), dtype=np.uint32)
        self.bvs = build_bvh_obbrss(self.vertices, self.tri_indices)
        self.num_vertices, self.num_tris = len(self.vertices), len(self.tri_indices)
        self._begun = False
        return 0

    def getNumBVs(self):
        return 0 if self.bvs is None else len(self.bvs)


class BVHModelOBB(BVHModelOBBRSS):
    """BVHModel<OBB>: collide() only, as in the library (distance() on a plain OBB model is not offered, see
    hfb_geom_register_bvh_obb).  The plain OBB tree of a mesh is the OBB half of its OBBRSS tree, so the builder is shared."""
    node_type = P.BV_OBB


_MESH_TYPES = (P.BV_OBBRSS, P.BV_OBB)


# ---------------------------------------------------------- requests/results --
class _QueryRequest:
    _pod_cls = None

    def __init__(self, **kw):
        self._pod = self._pod_cls(**kw)

    def __getattr__(self, k):
        pod = object.__getattribute__(self, "_pod")
        if hasattr(pod.q, k):
            return getattr(pod.q, k)
        return getattr(pod, k)

    def __setattr__(self, k, v):
        if k == "_pod":
            object.__setattr__(self, k, v)
        elif hasattr(self._pod.q, k):
            setattr(self._pod.q, k, v)
        elif hasattr(self._pod, k):
            setattr(self._pod, k, v)
        else:
            raise AttributeError(k)


class CollisionRequest(_QueryRequest):  # collision_data.h:312-383
    _pod_cls = P.CollisionRequestPOD


class DistanceRequest(_QueryRequest):  # collision_data.h:987-1050
    _pod_cls = P.DistanceRequestPOD


class Contact:  # collision_data.h:59-166
    NONE = -1

    def __init__(self, o1, o2, rec):
        self.o1, self.o2 = o1, o2
        self.b1, self.b2 = int(rec["b1"]), int(rec["b2"])
        self.normal = rec["normal"].copy()
        self.nearest_points = [rec["p1"].copy(), rec["p2"].copy()]
        self.pos = rec["pos"].copy()
        self.penetration_depth = float(rec["distance"])


class CollisionResult:  # collision_data.h:391-509
    def __init__(self):
        self.clear()

    def clear(self):
        self.contacts = []
        self.distance_lower_bound = P.DBL_MAX
        self.normal = _NAN3.copy()
        self.nearest_points = [_NAN3.copy(), _NAN3.copy()]
        self.cached_gjk_guess = np.array([1.0, 0.0, 0.0])
        self.cached_support_func_guess = np.zeros(2, dtype=np.int32)

    def numContacts(self):
        return len(self.contacts)

    def isCollision(self):
        return len(self.contacts) > 0

    def getContact(self, i):
        if i >= len(self.contacts):
            raise IndexError("The index is out of range.")
        return self.contacts[i]


class DistanceResult:  # collision_data.h:1053-1174
    NONE = -1

    def __init__(self):
        self.clear()

    def clear(self):
        self.min_distance = P.DBL_MAX
        self.nearest_points = [_NAN3.copy(), _NAN3.copy()]
        self.normal = _NAN3.copy()
        self.o1 = self.o2 = None
        self.b1 = self.b2 = -1
        self.cached_gjk_guess = np.array([1.0, 0.0, 0.0])
        self.cached_support_func_guess = np.zeros(2, dtype=np.int32)


# ----------------------------------------------------------------- engine ----
_DEFAULT = {}


def default_engine(device=0):
    if device not in _DEFAULT:
        _DEFAULT[device] = _Scene(Engine(device))
    return _DEFAULT[device]


class _Scene:
    """Geometry handle cache keyed by object identity (the arena of SURVEY 8b)."""

    def __init__(self, engine):
        self.engine = engine
        self._handles = {}
        self._dirty = False

    @staticmethod
    def _content(geom):
        """what the arena holds for `geom` beyond its record: the vertices of a hull / triangle, the mesh of a model"""
        pts = geom._points()
        if geom.node_type in _MESH_TYPES:
            return (np.ascontiguousarray(geom.bvs).tobytes() if geom.bvs is not None else None,
                    np.ascontiguousarray(geom.vertices, dtype=np.float64).tobytes(),
                    np.ascontiguousarray(geom.tri_indices, dtype=np.uint32).tobytes())
        return np.ascontiguousarray(pts, dtype=np.float64).tobytes() if pts is not None else None

    def handle(self, geom):
        """the arena handle of `geom`, registered on first use.  A geometry changed in place since (new sizes, new
        swept-sphere radius, moved hull vertices) keeps its handle: the record / vertex set is updated
        (hfb_geom_update_shapes / hfb_geom_update_convex); only a change the arena cannot absorb -- another number of
        vertices, another mesh -- retires the handle and registers a new one."""
        key = id(geom)
        ent = self._handles.get(key)
        if geom.node_type in (P.GEOM_PLANE, P.GEOM_HALFSPACE):
            # (n, d) live outside the 40-byte record: a plane changed in place is registered anew
            sig = (geom.node_type, tuple(geom._plane().tolist()), geom.getSweptSphereRadius())
            if ent is not None and ent[2] is geom:
                if ent[1] == sig:
                    return ent[0]
                self.engine.release_shapes([ent[0]])
            h = int(self.engine.register_halfspaces(geom.node_type, geom._plane(), [geom.getSweptSphereRadius()])[0])
            self._handles[key] = (h, sig, geom, None, 0)
            self._dirty = True
            return h
        sig = (geom.node_type, tuple(np.asarray(geom._params(), dtype=np.float64).tolist()), geom.getSweptSphereRadius())
        content = self._content(geom)
        if ent is not None and ent[2] is geom:
            h, old_sig, _, old_content, data = ent
            if old_sig == sig and old_content == content:
                return h
            same_kind = old_sig[0] == sig[0]
            pts = geom._points()
            if same_kind and geom.node_type not in _MESH_TYPES and (
                    old_content == content or (pts is not None and old_content is not None and len(old_content) == len(content))):
                if old_content != content:
                    self.engine.update_convex(data, pts)
                rec = P.make_shapes([geom.node_type], [geom._params()], ssr=geom.getSweptSphereRadius(), data=data)
                self.engine.update_shapes([h], rec)
                self._handles[key] = (h, sig, geom, content, data)
                self._dirty = True
                return h
            self.engine.release_shapes([h])  # the storage of its vertices / mesh stays until hfb_geom_clear
        data = 0
        pts = geom._points()
        if geom.node_type in _MESH_TYPES:
            if geom.bvs is None:
                raise ValueError("BVHModel: endModel() was not called")
            reg = self.engine.register_bvh_obb if geom.node_type == P.BV_OBB else self.engine.register_bvh_obbrss
            data = reg(geom.bvs, geom.vertices, geom.tri_indices)
        elif pts is not None:
            data = self.engine.register_convex(pts)
        rec = P.make_shapes([geom.node_type], [geom._params()], ssr=geom.getSweptSphereRadius(), data=data)
        h = int(self.engine.register_shapes(rec)[0])
        self._handles[key] = (h, sig, geom, content, data)
        self._dirty = True
        return h

    def commit(self):
        if self._dirty:
            self.engine.commit()
            self._dirty = False


def _tf_array(tfs):
    if isinstance(tfs, np.ndarray) and tfs.dtype == P.transform_dtype:
        return tfs
    if isinstance(tfs, Transform3f):
        tfs = [tfs]
    R = np.stack([t.R for t in tfs])
    T = np.stack([t.T for t in tfs])
    return P.make_transforms(R, T)


class BatchQuery:
    """Batched collide()/distance(): the seam a broadphase CollisionCallBackCollect
    (broadphase/default_broadphase_callbacks.h:224-252) feeds."""

    def __init__(self, device=0):
        self.scene = default_engine(device)

    def handles(self, geoms):
        return np.array([self.scene.handle(g) for g in geoms], dtype=np.uint32)

    def distance(self, geoms1, tfs1, geoms2, tfs2, request=None):
        request = request or DistanceRequest()
        h1, h2 = self.handles(geoms1), self.handles(geoms2)
        self.scene.commit()
        return self.scene.engine.batch_distance(h1, _tf_array(tfs1), h2, _tf_array(tfs2), request._pod)

    def collide(self, geoms1, tfs1, geoms2, tfs2, request=None):
        request = request or CollisionRequest()
        if request.num_max_contacts == 0:  # collision.cpp:82-85
            raise ValueError("Invalid number of max contacts (current value is 0).")
        h1, h2 = self.handles(geoms1), self.handles(geoms2)
        self.scene.commit()
        return self.scene.engine.batch_collide(h1, _tf_array(tfs1), h2, _tf_array(tfs2), request._pod)


# ------------------------------------------------------------- broadphase seam --
class AABB:  # BV/AABB.h
    def __init__(self, min_=None, max_=None):
        self.min_ = np.full(3, np.inf) if min_ is None else np.asarray(min_, dtype=np.float64)
        self.max_ = np.full(3, -np.inf) if max_ is None else np.asarray(max_, dtype=np.float64)

    def overlap(self, other):  # :111-118, closed intervals
        return bool(np.all(self.min_ <= other.max_) and np.all(self.max_ >= other.min_))


class CollisionObject:  # collision_object.h:214-330 (geometry + pose + world-space box)
    def __init__(self, geom, tf=None):
        self.geom = geom
        self.tf = tf if tf is not None else Transform3f()
        self.aabb = AABB()

    def collisionGeometry(self):
        return self.geom

    def getTransform(self):
        return self.tf

    def setTransform(self, tf):
        self.tf = tf

    def getAABB(self):
        return self.aabb


class CollisionCallBackCollect:  # broadphase/default_broadphase_callbacks.h:224-252, .cpp:95-120
    def __init__(self, max_size=0):
        self.max_size = max_size
        self.collision_pairs = []

    def init(self):
        self.collision_pairs = []

    def collide(self, o1, o2):
        self.collision_pairs.append((o1, o2))
        return False

    __call__ = collide

    def numCollisionPairs(self):
        return len(self.collision_pairs)

    def getCollisionPairs(self):
        return self.collision_pairs

    def exist(self, o1, o2):
        return any(a is o1 and b is o2 for a, b in self.collision_pairs)


class DynamicAABBTreeCollisionManager:
    """BroadPhaseCollisionManager's interface (broadphase/broadphase_collision_manager.h:56-134) for the self-collision
    query: setup() / update() compute every object's box (CollisionObject::computeAABB, on the device), collide(callback)
    reports every pair of objects with overlapping boxes once -- the SET the reference's managers report
    (broadphase_dynamic_AABB_tree.cpp:336-407,716-721); the order is each manager's own.  No tree is kept: a uniform grid
    is rebuilt per query.  collide_batch() is what the seam is for: candidates + batched narrow phase in one go."""

    def __init__(self, device=0):
        self.scene = default_engine(device)
        self.objs = []
        self._stale = True
        self._handles = self._tfs = self._boxes = None

    def registerObject(self, obj):
        self.objs.append(obj)
        self._stale = True

    def registerObjects(self, objs):
        self.objs.extend(objs)
        self._stale = True

    def unregisterObject(self, obj):
        self.objs = [o for o in self.objs if o is not obj]
        self._stale = True

    def clear(self):
        self.objs = []
        self._stale = True

    def size(self):
        return len(self.objs)

    def empty(self):
        return not self.objs

    def getObjects(self):
        return list(self.objs)

    def _refresh(self):
        self._handles = np.array([self.scene.handle(o.geom) for o in self.objs], dtype=np.uint32)
        self.scene.commit()
        self._tfs = _tf_array([o.tf for o in self.objs]) if self.objs else np.zeros(0, dtype=P.transform_dtype)
        self._boxes = self.scene.engine.scene_aabbs(self._handles, self._tfs) if self.objs else np.zeros((0, 6))
        for o, b in zip(self.objs, self._boxes):
            o.aabb = AABB(b[:3].copy(), b[3:].copy())
        self._stale = False

    def setup(self):
        self._refresh()

    def update(self):
        self._refresh()

    def pairs(self):
        """the candidate pairs as object indices (first < second)"""
        if self._stale:
            self._refresh()
        from .engine import broadphase_pairs
        return broadphase_pairs(self._boxes)

    def collide(self, callback):
        callback.init()
        first, second = self.pairs()
        for a, b in zip(first.tolist(), second.tolist()):
            if callback(self.objs[a], self.objs[b]):
                return

    def collide_batch(self, request=None):
        """-> (first, second, contact records) of ALL candidate pairs: one hfb_batch_collide_objects call"""
        request = request or CollisionRequest()
        if request.num_max_contacts == 0:
            raise ValueError("Invalid number of max contacts (current value is 0).")
        first, second = self.pairs()
        rec = self.scene.engine.batch_collide_objects(self._handles, self._tfs, first, second, request._pod)
        return first, second, rec


def _unsupported(o1, o2, what, request=None):
    """the exception the reference throws for this pair (the record carries HFB_PATH_UNSUPPORTED)"""
    mesh = [o for o in (o1, o2) if o.node_type in _MESH_TYPES]
    shape = [o for o in (o1, o2) if o.node_type not in _MESH_TYPES]
    if len(mesh) == 1 and shape:
        if what == "Collision" and request is not None and request.security_margin < 0:
            # collision_func_matrix.cpp:109-112
            raise ValueError("Negative security margin are not handled yet for BVHModel")
        if getattr(shape[0], "getSweptSphereRadius", lambda: 0.0)() > 0:
            raise RuntimeError("Swept-sphere radius not yet supported.")  # geometric_shapes_utility.h:73-78
        if request is not None and request.gjk_initial_guess == P.BoundingVolumeGuess:
            raise RuntimeError("computeLocalAABB must have been called on the shapes before using "
                               "GJKInitialGuess::BoundingVolumeGuess.")  # narrowphase.h:368-377
    raise ValueError("%s function between node type %s and node type %s is not yet supported."
                     % (what, o1.node_type, o2.node_type))


def collide(o1, tf1, o2, tf2, request, result, device=0):
    """collide() of src/collision.cpp:69-130. Results accumulate: callers clear()."""
    if request.security_margin == -np.inf:  # :73-76
        result.clear()
        return 0
    if request.num_max_contacts == 0:
        raise ValueError("Invalid number of max contacts (current value is 0).")
    if result.isCollision() and request.num_max_contacts <= result.numContacts():  # isSatisfied
        return result.numContacts()
    bq = BatchQuery(device)
    scene = bq.scene
    h1, h2 = scene.handle(o1), scene.handle(o2)
    scene.commit()
    mesh = o1.node_type in _MESH_TYPES or o2.node_type in _MESH_TYPES
    more = []
    if mesh and request.num_max_contacts > 1:  # every contact of the pair, not only contacts[0]
        out, extra, counts = scene.engine.batch_collide_contacts([h1], _tf_array(tf1), [h2], _tf_array(tf2), request._pod,
                                                                 max_extra=int(request.num_max_contacts) - 1)
        more = [extra[0, k] for k in range(int(counts[0]) - 1)]
        gg, gh = [result.cached_gjk_guess], [result.cached_support_func_guess]
    else:
        out, gg, gh = scene.engine.batch_collide([h1], _tf_array(tf1), [h2], _tf_array(tf2), request._pod,
                                                 want_guess=True)
    rec = out[0]
    if P.status_path(rec["status"]) == P.PATH_UNSUPPORTED:
        _unsupported(o1, o2, "Collision", request)
    if rec["distance_lower_bound"] < result.distance_lower_bound:
        result.distance_lower_bound = float(rec["distance_lower_bound"])
        result.nearest_points = [rec["p1"].copy(), rec["p2"].copy()]
        result.normal = rec["normal"].copy()
    if rec["num_contacts"]:
        result.contacts.append(Contact(o1, o2, rec))
        for x in more:
            result.contacts.append(Contact(o1, o2, x))
    result.cached_gjk_guess = gg[0]
    result.cached_support_func_guess = gh[0]
    return result.numContacts() if rec["num_contacts"] else 0


def distance(o1, tf1, o2, tf2, request, result, device=0):
    """distance() of src/distance.cpp:60-109."""
    if result.min_distance <= 0:  # DistanceRequest::isSatisfied
        return result.min_distance
    bq = BatchQuery(device)
    scene = bq.scene
    h1, h2 = scene.handle(o1), scene.handle(o2)
    scene.commit()
    out, gg, gh = scene.engine.batch_distance([h1], _tf_array(tf1), [h2], _tf_array(tf2), request._pod,
                                              want_guess=True)
    rec = out[0]
    if P.status_path(rec["status"]) == P.PATH_UNSUPPORTED:
        _unsupported(o1, o2, "Distance", request)
    d = float(rec["min_distance"])
    closed = P.status_path(rec["status"]) == P.PATH_CLOSED_FORM
    if closed or result.min_distance > d:
        result.min_distance = d
        result.o1, result.o2 = o1, o2
        result.b1, result.b2 = int(rec["b1"]), int(rec["b2"])
        result.nearest_points = [rec["p1"].copy(), rec["p2"].copy()]
        result.normal = rec["normal"].copy()
    result.cached_gjk_guess = gg[0]
    result.cached_support_func_guess = gh[0]
    return d


class ComputeCollision:  # collision.h:79-117
    def __init__(self, o1, o2):
        self.o1, self.o2 = o1, o2

    def __call__(self, tf1, tf2, request, result):
        return collide(self.o1, tf1, self.o2, tf2, request, result)


class ComputeDistance:  # distance.h:74-112
    def __init__(self, o1, o2):
        self.o1, self.o2 = o1, o2

    def __call__(self, tf1, tf2, request, result):
        return distance(self.o1, tf1, self.o2, tf2, request, result)
