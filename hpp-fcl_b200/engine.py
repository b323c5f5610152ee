"""ctypes binding of the C-ABI (include/hppfcl_b200.h) -- the drop-in boundary.

`Engine` owns one hfb_ctx (one CUDA device).  Everything that computes goes
through libhppfcl_b200.so; there is no Python or CPU fallback: a missing library
or a missing GPU raises EngineError.
"""
import ctypes as C
import os

import numpy as np

from . import _pod as P
from .build import library_path

_LIB = None


class EngineError(RuntimeError):
    pass


_ERR = {1: "invalid argument", 2: "no CUDA device", 3: "CUDA error", 4: "unsupported pair",
        5: "out of memory"}


def load_library(path=None):
    """dlopen the in-tree extension. Works without a GPU (symbols only)."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    path = path or library_path()
    if not os.path.exists(path):
        raise EngineError("CUDA extension %s is not built (run __graft_entry__.build())" % path)
    L = C.CDLL(path)
    vp, sz, u32 = C.c_void_p, C.c_size_t, C.c_uint32
    L.hfb_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.hfb_ctx_destroy.argtypes = [vp]
    L.hfb_last_error.argtypes = [vp]
    L.hfb_last_error.restype = C.c_char_p
    L.hfb_version.restype = C.c_char_p
    L.hfb_default_distance_request.argtypes = [vp]
    L.hfb_default_collision_request.argtypes = [vp]
    L.hfb_geom_register_shapes.argtypes = [vp, vp, sz, vp]
    L.hfb_geom_register_halfspaces.argtypes = [vp, C.c_uint32, vp, vp, sz, vp]
    L.hfb_geom_register_convex.argtypes = [vp, vp, u32, vp]
    L.hfb_geom_register_convex_batch.argtypes = [vp, vp, u32, u32, vp]
    L.hfb_geom_register_bvh_obbrss.argtypes = [vp, vp, u32, vp, u32, vp, u32, vp]
    L.hfb_bvh_build_obbrss.argtypes = [vp, u32, vp, u32, vp, u32]
    L.hfb_geom_commit.argtypes = [vp]
    L.hfb_geom_device_arena.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
    L.hfb_geom_num_shapes.argtypes = [vp]
    L.hfb_geom_num_shapes.restype = sz
    L.hfb_geom_clear.argtypes = [vp]
    L.hfb_geom_update_shapes.argtypes = [vp, vp, vp, C.c_size_t]
    L.hfb_geom_update_convex.argtypes = [vp, C.c_uint32, vp, C.c_uint32]
    L.hfb_geom_release_shapes.argtypes = [vp, vp, C.c_size_t]
    for name in ("hfb_batch_distance", "hfb_batch_collide"):
        getattr(L, name).argtypes = [vp, sz, vp, vp, vp, vp, vp, vp, vp]
        getattr(L, name + "_device").argtypes = [vp, sz, vp, vp, vp, vp, vp, vp, vp, vp]
    L.hfb_batch_collide_contacts.argtypes = [vp, sz, vp, vp, vp, vp, vp, vp, u32, vp, vp, vp]
    L.hfb_scene_aabbs.argtypes = [vp, sz, vp, vp, vp]
    L.hfb_broadphase_pairs.argtypes = [sz, vp, vp, vp, sz, vp]
    L.hfb_scene_aabbs_device.argtypes = [vp, sz, vp, vp, vp, vp]
    L.hfb_broadphase_pairs_device.argtypes = [vp, sz, vp, sz, sz, vp, vp, sz, vp, vp]
    L.hfb_scene_collide.argtypes = [vp, sz, vp, vp, sz, sz, vp, vp]
    L.hfb_comm_unique_id.argtypes = [vp]
    L.hfb_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    L.hfb_comm_destroy.argtypes = [vp]
    L.hfb_geom_broadcast.argtypes = [vp, C.c_int]
    L.hfb_comm_wait.argtypes = [vp, vp]
    for name in ("hfb_batch_distance_sharded_device", "hfb_batch_collide_sharded_device"):
        getattr(L, name).argtypes = [vp, sz, vp, vp, vp, vp, vp, vp, vp]
    L.hfb_batch_distance_objects.argtypes = [vp, vp, vp, vp, vp, vp]
    L.hfb_batch_collide_objects.argtypes = [vp, vp, vp, vp, vp, vp]
    L.hfb_batch_distance_objects_device.argtypes = [vp, vp, vp, vp, vp, vp]
    L.hfb_batch_collide_objects_device.argtypes = [vp, vp, vp, vp, vp, vp]
    L.hfb_batch_convex_support.argtypes = [vp, sz, vp, vp, vp, vp]
    L.hfb_batch_convex_support_device.argtypes = [vp, sz, vp, vp, vp, vp, vp]
    L.hfb_get_stats.argtypes = [vp, vp]
    L.hfb_set_profiling.argtypes = [vp, C.c_int]
    L.hfb_get_kernel_times.argtypes = [vp, vp, C.c_int]
    if path == library_path():
        _LIB = L
    return L


def build_bvh_obbrss(vertices, triangles):
    """Host-side OBBRSS tree of a triangle mesh (BVHModel<OBBRSS>::endModel, mean split): the node array
    hfb_geom_register_bvh_obbrss takes.  Needs no GPU."""
    L = load_library()
    v = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 3)
    t = np.ascontiguousarray(triangles, dtype=np.uint32).reshape(-1, 3)
    nodes = np.zeros(max(2 * t.shape[0] - 1, 0), dtype=P.bvh_node_dtype)
    rc = L.hfb_bvh_build_obbrss(_ptr(v), v.shape[0], _ptr(t), t.shape[0], _ptr(nodes), nodes.shape[0])
    if rc != 0:
        raise EngineError("hfb_bvh_build_obbrss: %s" % _ERR.get(rc, rc))
    return nodes


def broadphase_pairs(aabbs, capacity=None):
    """every pair i < j of overlapping boxes (n x 6: min xyz, max xyz), in no particular order -> (first, second).
    Host function of the library: needs neither a context nor a GPU."""
    L = load_library()
    bb = np.ascontiguousarray(aabbs, dtype=np.float64).reshape(-1, 6)
    n = bb.shape[0]
    cap = int(capacity) if capacity is not None else max(16 * n, 1024)
    while True:
        first = np.empty(cap, dtype=np.uint32)
        second = np.empty(cap, dtype=np.uint32)
        cnt = C.c_size_t(0)
        rc = L.hfb_broadphase_pairs(n, _ptr(bb), _ptr(first), _ptr(second), cap, C.byref(cnt))
        if rc != 0:
            raise EngineError("hfb_broadphase_pairs: %s" % _ERR.get(rc, rc))
        if cnt.value <= cap or capacity is not None:
            k = min(cnt.value, cap)
            return first[:k], second[:k]
        cap = cnt.value


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


class Engine:
    """One device context + geometry arena (ctx_create / geom_register_* / batch_*)."""

    def __init__(self, device=0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.hfb_ctx_create(int(device), C.byref(h))
        if rc != 0:
            raise EngineError("hfb_ctx_create(device=%d) failed: %s (no CPU fallback exists)"
                              % (device, _ERR.get(rc, rc)))
        self.h = h
        self.device = device

    # -- lifetime ---------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.L.hfb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self.L.hfb_last_error(self.h)
            raise EngineError("%s: %s" % (_ERR.get(rc, rc), msg.decode() if msg else ""))

    # -- geometry ---------------------------------------------------------
    def register_convex(self, points):
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        cid = C.c_uint32()
        self._check(self.L.hfb_geom_register_convex(self.h, _ptr(pts), pts.shape[0], C.byref(cid)))
        return cid.value

    def register_shapes(self, shapes):
        shapes = np.ascontiguousarray(shapes, dtype=P.shape_dtype)
        handles = np.zeros(shapes.shape[0], dtype=np.uint32)
        self._check(self.L.hfb_geom_register_shapes(self.h, _ptr(shapes), shapes.shape[0], _ptr(handles)))
        return handles

    def register_halfspaces(self, kind, n_d, ssr=None):
        """Halfspace (kind = GEOM_HALFSPACE) or Plane (GEOM_PLANE) geometries: n_d is (count, 4) rows
        (n.x, n.y, n.z, d), normalised by the library; ssr: count swept-sphere radii or None.  Returns the handles."""
        nd = np.ascontiguousarray(n_d, dtype=np.float64).reshape(-1, 4)
        r = None if ssr is None else np.ascontiguousarray(ssr, dtype=np.float64).reshape(-1)
        if r is not None and r.shape[0] != nd.shape[0]:
            raise ValueError("one swept-sphere radius per plane")
        handles = np.zeros(nd.shape[0], dtype=np.uint32)
        self._check(self.L.hfb_geom_register_halfspaces(self.h, int(kind), _ptr(nd), None if r is None else _ptr(r),
                                                        nd.shape[0], _ptr(handles)))
        return handles

    def register_convex_batch(self, points):
        """points: (count, nv, 3); returns the id of the first hull, the others follow"""
        pts = np.ascontiguousarray(points, dtype=np.float64)
        assert pts.ndim == 3 and pts.shape[2] == 3
        first = C.c_uint32()
        self._check(self.L.hfb_geom_register_convex_batch(self.h, _ptr(pts), pts.shape[1], pts.shape[0], C.byref(first)))
        return first.value

    def register_bvh_obbrss(self, nodes, vertices, triangles):
        """nodes: BVHModel<OBBRSS>::bvs as hfb_bvh_node records (the reference's own tree, or
        build_bvh_obbrss(vertices, triangles)); None builds it here."""
        if nodes is None:
            nodes = build_bvh_obbrss(vertices, triangles)
        nodes = np.ascontiguousarray(nodes, dtype=P.bvh_node_dtype)
        v = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 3)
        t = np.ascontiguousarray(triangles, dtype=np.uint32).reshape(-1, 3)
        bid = C.c_uint32()
        self._check(self.L.hfb_geom_register_bvh_obbrss(self.h, _ptr(nodes), nodes.shape[0], _ptr(v),
                                                        v.shape[0], _ptr(t), t.shape[0], C.byref(bid)))
        return bid.value  # use as `data` of a shape record of type BV_OBBRSS

    def register_bvh_obb(self, nodes, vertices, triangles):
        """a plain BVHModel<OBB> (collide() only): the same node records, RSS half ignored -- the OBB tree of a mesh is the
        OBB half of its OBBRSS tree, so build_bvh_obbrss serves both; use as `data` of a shape record of type BV_OBB"""
        if nodes is None:
            nodes = build_bvh_obbrss(vertices, triangles)
        nodes = np.ascontiguousarray(nodes, dtype=P.bvh_node_dtype)
        v = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 3)
        t = np.ascontiguousarray(triangles, dtype=np.uint32).reshape(-1, 3)
        bid = C.c_uint32()
        self._check(self.L.hfb_geom_register_bvh_obb(self.h, _ptr(nodes), nodes.shape[0], _ptr(v), v.shape[0], _ptr(t),
                                                     t.shape[0], C.byref(bid)))
        return bid.value

    def commit(self):
        self._check(self.L.hfb_geom_commit(self.h))

    def clear_geometry(self):
        """drops every registered shape / hull / mesh; all handles become invalid"""
        self._check(self.L.hfb_geom_clear(self.h))

    def update_shapes(self, handles, shapes):
        """replace the records of registered handles (e.g. new sizes); commit() before the next query"""
        handles = np.ascontiguousarray(handles, dtype=np.uint32)
        shapes = np.ascontiguousarray(shapes, dtype=P.shape_dtype)
        assert handles.shape[0] == shapes.shape[0]
        self._check(self.L.hfb_geom_update_shapes(self.h, _ptr(handles), _ptr(shapes), handles.shape[0]))

    def update_convex(self, convex_id, points):
        """replace the vertices of a registered hull by as many new ones"""
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        self._check(self.L.hfb_geom_update_convex(self.h, int(convex_id), _ptr(pts), pts.shape[0]))

    def release_shapes(self, handles):
        """retire handles: pairs that still name them come back as unsupported"""
        handles = np.ascontiguousarray(handles, dtype=np.uint32)
        self._check(self.L.hfb_geom_release_shapes(self.h, _ptr(handles), handles.shape[0]))

    def device_arena(self):
        base = C.c_void_p()
        n = C.c_size_t()
        self._check(self.L.hfb_geom_device_arena(self.h, C.byref(base), C.byref(n)))
        return base.value, n.value

    # -- host-buffer queries (H2D + kernels + D2H inside the call) ----------
    def _host_call(self, fn, out_dtype, h1, tf1, h2, tf2, req, want_guess, out=None):
        h1 = np.ascontiguousarray(h1, dtype=np.uint32)
        h2 = np.ascontiguousarray(h2, dtype=np.uint32)
        tf1 = np.ascontiguousarray(tf1, dtype=P.transform_dtype)
        tf2 = np.ascontiguousarray(tf2, dtype=P.transform_dtype)
        n = h1.shape[0]
        if not (h2.shape[0] == n and tf1.shape[0] == n and tf2.shape[0] == n):
            raise ValueError("batch arrays must have equal length")
        if out is None:
            out = np.empty(n, dtype=out_dtype)
        g = None
        gg = gh = None
        if want_guess:
            gg = np.zeros((n, 3), dtype=np.float64)
            gh = np.zeros((n, 2), dtype=np.int32)
            g = P.GuessOut(_ptr(gg), _ptr(gh))
        self._check(fn(self.h, n, _ptr(h1), _ptr(tf1), _ptr(h2), _ptr(tf2), C.byref(req), _ptr(out),
                       C.byref(g) if g is not None else None))
        if want_guess:
            return out, gg, gh
        return out

    def batch_distance(self, h1, tf1, h2, tf2, req=None, want_guess=False, out=None):
        req = req or P.DistanceRequestPOD()
        return self._host_call(self.L.hfb_batch_distance, P.distance_result_dtype, h1, tf1, h2, tf2,
                               req, want_guess, out)

    def batch_collide(self, h1, tf1, h2, tf2, req=None, want_guess=False, out=None):
        req = req or P.CollisionRequestPOD()
        return self._host_call(self.L.hfb_batch_collide, P.contact_dtype, h1, tf1, h2, tf2, req,
                               want_guess, out)

    def batch_collide_contacts(self, h1, tf1, h2, tf2, req=None, max_extra=3):
        """collide() keeping up to 1 + max_extra contacts of a mesh pair: -> (out, extra[n, max_extra], counts);
        contacts[k] of pair i, 1 <= k < min(counts[i], max_extra + 1), is extra[i, k - 1]"""
        req = req or P.CollisionRequestPOD()
        h1 = np.ascontiguousarray(h1, dtype=np.uint32)
        h2 = np.ascontiguousarray(h2, dtype=np.uint32)
        tf1 = np.ascontiguousarray(tf1, dtype=P.transform_dtype)
        tf2 = np.ascontiguousarray(tf2, dtype=P.transform_dtype)
        n = h1.shape[0]
        out = np.zeros(n, dtype=P.contact_dtype)
        extra = np.zeros((n, max(max_extra, 1)), dtype=P.contact_dtype)
        counts = np.zeros(n, dtype=np.uint32)
        self._check(self.L.hfb_batch_collide_contacts(self.h, n, _ptr(h1), _ptr(tf1), _ptr(h2), _ptr(tf2), C.byref(req),
                                                      _ptr(out), max_extra, _ptr(extra), _ptr(counts), None))
        return out, extra[:, :max_extra], counts

    # -- broadphase feed ---------------------------------------------------------------------------------
    def scene_aabbs(self, obj_handles, obj_tfs):
        """CollisionObject::computeAABB of every object -> n x 6 (min xyz, max xyz); host function"""
        oh = np.ascontiguousarray(obj_handles, dtype=np.uint32)
        ot = np.ascontiguousarray(obj_tfs, dtype=P.transform_dtype)
        bb = np.empty((oh.shape[0], 6), dtype=np.float64)
        self._check(self.L.hfb_scene_aabbs(self.h, oh.shape[0], _ptr(oh), _ptr(ot), _ptr(bb)))
        return bb

    def scene_aabbs_device(self, n, d_handles, d_tfs, d_aabbs, stream=0):
        self._check(self.L.hfb_scene_aabbs_device(self.h, n, _ptr(d_handles), _ptr(d_tfs), _ptr(d_aabbs), _ptr(stream)))

    def broadphase_pairs_device(self, n, d_aabbs, d_first, d_second, capacity, d_n_pairs, stream=0, first_object=0,
                                num_first_objects=None):
        """pairs (i, j), i < j, with i in [first_object, first_object + num_first_objects) (default: all)"""
        self._check(self.L.hfb_broadphase_pairs_device(self.h, n, _ptr(d_aabbs), first_object,
                                                       n if num_first_objects is None else num_first_objects, _ptr(d_first),
                                                       _ptr(d_second), capacity, _ptr(d_n_pairs), _ptr(stream)))

    def scene_collide(self, obj_handles, obj_tfs, req=None, capacity=None, first_object=0, num_first_objects=None, bufs=None):
        """broadphase + narrow phase of a scene in one call (hfb_scene_collide)
        -> (first, second, contacts, n_candidates, n_colliding): the colliding pairs, in no particular order"""
        req = req or P.CollisionRequestPOD()
        oh = np.ascontiguousarray(obj_handles, dtype=np.uint32)
        ot = np.ascontiguousarray(obj_tfs, dtype=P.transform_dtype)
        n = oh.shape[0]
        cap = int(capacity) if capacity is not None else max(4 * n, 1024)
        if bufs is None:
            bufs = (np.empty(cap, dtype=np.uint32), np.empty(cap, dtype=np.uint32), np.empty(cap, dtype=P.contact_dtype))
        f, s2, rec = bufs
        nc = np.zeros(2, dtype=np.uint32)
        out = P.SceneContacts(f.ctypes.data, s2.ctypes.data, rec.ctypes.data, cap, nc[0:1].ctypes.data, nc[1:2].ctypes.data)
        self._check(self.L.hfb_scene_collide(self.h, n, _ptr(oh), _ptr(ot), first_object,
                                             n if num_first_objects is None else num_first_objects, C.byref(req), C.byref(out)))
        k = min(int(nc[0]), cap)
        return f[:k], s2[:k], rec[:k], int(nc[1]), int(nc[0])

    # -- object-table queries: the batched form of the CollisionObject overloads (collision.h:58-61) -----
    @staticmethod
    def _scene(obj_handles, obj_tfs, first, second):
        oh = np.ascontiguousarray(obj_handles, dtype=np.uint32)
        ot = np.ascontiguousarray(obj_tfs, dtype=P.transform_dtype)
        pi = np.ascontiguousarray(first, dtype=np.uint32)
        pj = np.ascontiguousarray(second, dtype=np.uint32)
        if oh.shape[0] != ot.shape[0] or pi.shape[0] != pj.shape[0]:
            raise ValueError("object table / pair list arrays must have equal length")
        sc = P.ObjectPairs(oh.shape[0], oh.ctypes.data, ot.ctypes.data, pi.shape[0], pi.ctypes.data, pj.ctypes.data)
        return sc, (oh, ot, pi, pj)

    def batch_distance_objects(self, obj_handles, obj_tfs, first, second, req=None, out=None, min_only=False):
        """distance() of the object pairs (first[k], second[k]).  min_only: only DistanceResult::min_distance
        comes back (8 B per pair over PCIe instead of 96)"""
        req = req or P.DistanceRequestPOD()
        sc, keep = self._scene(obj_handles, obj_tfs, first, second)
        n = sc.n_pairs
        if min_only:
            d = np.empty(n, dtype=np.float64) if out is None else out
            self._check(self.L.hfb_batch_distance_objects(self.h, C.byref(sc), C.byref(req), None, _ptr(d), None))
            return d
        if out is None:
            out = np.empty(n, dtype=P.distance_result_dtype)
        self._check(self.L.hfb_batch_distance_objects(self.h, C.byref(sc), C.byref(req), _ptr(out), None, None))
        return out

    def batch_collide_objects(self, obj_handles, obj_tfs, first, second, req=None, out=None, compact_capacity=None):
        """collide() of the object pairs.  compact_capacity = c: -> (flags bit per pair, n_colliding, pair_ids,
        contacts) with the records of at most c colliding pairs, in no particular order; else full records"""
        req = req or P.CollisionRequestPOD()
        sc, keep = self._scene(obj_handles, obj_tfs, first, second)
        n = sc.n_pairs
        if compact_capacity is not None:
            cap = int(compact_capacity)
            flags = np.zeros((n + 31) // 32, dtype=np.uint32)
            nhit = np.zeros(1, dtype=np.uint32)
            ids = np.zeros(max(cap, 1), dtype=np.uint32)
            recs = np.zeros(max(cap, 1), dtype=P.contact_dtype)
            cc = P.CompactContacts(flags.ctypes.data, nhit.ctypes.data, ids.ctypes.data, recs.ctypes.data, cap)
            self._check(self.L.hfb_batch_collide_objects(self.h, C.byref(sc), C.byref(req), None, C.byref(cc), None))
            k = min(int(nhit[0]), cap)
            return flags, int(nhit[0]), ids[:k], recs[:k]
        if out is None:
            out = np.empty(n, dtype=P.contact_dtype)
        self._check(self.L.hfb_batch_collide_objects(self.h, C.byref(sc), C.byref(req), _ptr(out), None, None))
        return out

    def batch_distance_objects_device(self, n_objects, d_handles, d_tfs, n_pairs, d_first, d_second, d_out, req=None,
                                      stream=0):
        req = req or P.DistanceRequestPOD()
        sc = P.ObjectPairs(n_objects, d_handles, d_tfs, n_pairs, d_first, d_second)
        self._check(self.L.hfb_batch_distance_objects_device(self.h, C.byref(sc), C.byref(req), _ptr(d_out), None,
                                                             _ptr(stream)))

    def batch_collide_objects_device(self, n_objects, d_handles, d_tfs, n_pairs, d_first, d_second, d_out, req=None,
                                     stream=0):
        req = req or P.CollisionRequestPOD()
        sc = P.ObjectPairs(n_objects, d_handles, d_tfs, n_pairs, d_first, d_second)
        self._check(self.L.hfb_batch_collide_objects_device(self.h, C.byref(sc), C.byref(req), _ptr(d_out), None,
                                                            _ptr(stream)))

    def batch_convex_support(self, convex_ids, dirs):
        ids = np.ascontiguousarray(convex_ids, dtype=np.uint32)
        d = np.ascontiguousarray(dirs, dtype=np.float64).reshape(-1, 3)
        idx = np.zeros(ids.shape[0], dtype=np.int32)
        sup = np.zeros((ids.shape[0], 3), dtype=np.float64)
        self._check(self.L.hfb_batch_convex_support(self.h, ids.shape[0], _ptr(ids), _ptr(d), _ptr(idx),
                                                    _ptr(sup)))
        return idx, sup

    # -- device-buffer queries (raw device pointers, async on `stream`) ------
    def batch_distance_device(self, n, d_h1, d_tf1, d_h2, d_tf2, d_out, req=None, stream=0, d_guess=None):
        req = req or P.DistanceRequestPOD()
        self._check(self.L.hfb_batch_distance_device(self.h, n, _ptr(d_h1), _ptr(d_tf1), _ptr(d_h2),
                                                     _ptr(d_tf2), C.byref(req), _ptr(d_out),
                                                     C.byref(d_guess) if d_guess is not None else None,
                                                     _ptr(stream)))

    def batch_collide_device(self, n, d_h1, d_tf1, d_h2, d_tf2, d_out, req=None, stream=0, d_guess=None):
        req = req or P.CollisionRequestPOD()
        self._check(self.L.hfb_batch_collide_device(self.h, n, _ptr(d_h1), _ptr(d_tf1), _ptr(d_h2),
                                                    _ptr(d_tf2), C.byref(req), _ptr(d_out),
                                                    C.byref(d_guess) if d_guess is not None else None,
                                                    _ptr(stream)))

    def batch_convex_support_device(self, n, d_ids, d_dirs, d_idx, d_sup, stream=0):
        self._check(self.L.hfb_batch_convex_support_device(self.h, n, _ptr(d_ids), _ptr(d_dirs),
                                                           _ptr(d_idx), _ptr(d_sup), _ptr(stream)))

    # -- multi-GPU: one context per rank (hfb_comm_*), NCCL behind the C-ABI -----------------------------------
    @staticmethod
    def comm_unique_id():
        """128 bytes rank 0 creates and hands to every rank (ncclGetUniqueId)"""
        L = load_library()
        buf = C.create_string_buffer(128)
        rc = L.hfb_comm_unique_id(buf)
        if rc != 0:
            raise EngineError("hfb_comm_unique_id: %s" % _ERR.get(rc, rc))
        return buf.raw

    def comm_init(self, comm_id, rank, nranks):
        self._check(self.L.hfb_comm_init(self.h, C.create_string_buffer(bytes(comm_id), 128), int(rank), int(nranks)))

    def comm_destroy(self):
        self._check(self.L.hfb_comm_destroy(self.h))

    def geom_broadcast(self, root=0):
        """the geometry registered on `root` replaces this context's, on every rank, committed"""
        self._check(self.L.hfb_geom_broadcast(self.h, int(root)))

    def batch_distance_sharded_device(self, n_local, d_h1, d_tf1, d_h2, d_tf2, req=None, stream=0):
        """-> device pointer of the records of all ranks' pairs (rank-major); complete after comm_wait(stream)"""
        req = req or P.DistanceRequestPOD()
        out = C.c_void_p()
        self._check(self.L.hfb_batch_distance_sharded_device(self.h, n_local, _ptr(d_h1), _ptr(d_tf1), _ptr(d_h2), _ptr(d_tf2),
                                                             C.byref(req), C.byref(out), _ptr(stream)))
        return out.value

    def batch_collide_sharded_device(self, n_local, d_h1, d_tf1, d_h2, d_tf2, req=None, stream=0):
        req = req or P.CollisionRequestPOD()
        out = C.c_void_p()
        self._check(self.L.hfb_batch_collide_sharded_device(self.h, n_local, _ptr(d_h1), _ptr(d_tf1), _ptr(d_h2), _ptr(d_tf2),
                                                            C.byref(req), C.byref(out), _ptr(stream)))
        return out.value

    def comm_wait(self, stream=0):
        self._check(self.L.hfb_comm_wait(self.h, _ptr(stream)))

    def set_profiling(self, on=True):
        self._check(self.L.hfb_set_profiling(self.h, int(on)))

    def kernel_times(self, reset=True):
        t = P.KernelTimes()
        self._check(self.L.hfb_get_kernel_times(self.h, C.byref(t), int(reset)))
        return {f[0]: getattr(t, f[0]) for f in P.KernelTimes._fields_}

    def stats(self):
        s = P.Stats()
        self._check(self.L.hfb_get_stats(self.h, C.byref(s)))
        return {f[0]: getattr(s, f[0]) for f in P.Stats._fields_}
