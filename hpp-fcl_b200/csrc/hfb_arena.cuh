// Geometry arena: the flattened, device-resident replacement of the reference's
// caller-owned `const CollisionGeometry*` objects (ShapeBase subclasses,
// ConvexBase::points -- include/hpp/fcl/shape/geometric_shapes.h:164-872).
//
// Layout in HBM (one allocation, replicated per GPU):
//   shapes : hfb_shape[nshapes]           40 B records, handle = index
//   cvx    : ConvexDesc[nconvex]          40 B records
//   pool   : double[]                     per convex, SoA  x[vpad] y[vpad] z[vpad],
//                                         vpad = nv rounded up to 2 so every block is
//                                         16-B aligned (cp.async.bulk granularity)
#pragma once
#include <vector>

#include "hfb_shapes.cuh"

namespace hfb {

struct ConvexDesc {
  uint32_t off;   // first double of the block in `pool`
  uint32_t nv;    // number of vertices
  uint32_t vpad;  // padded count (stride between x / y / z arrays)
  uint32_t _r;
  double cx, cy, cz;  // aabb_local.center()
};

struct BvhDesc {  // one registered BVHModel<OBBRSS>
  uint32_t node_off, num_nodes;  // into the node array
  uint32_t vert_off, num_verts;  // in vertices (3 doubles each)
  uint32_t tri_off, num_tris;    // in triangles (3 uint32 each)
  uint32_t _r0, _r1;  // depth of the tree; 0: BVHModel<OBBRSS>, 1: BVHModel<OBB> (RSS half of the nodes unused)
};

struct ArenaView {
  const hfb_shape* shapes;
  const ConvexDesc* cvx;
  const double* pool;
  uint32_t nshapes;
  uint32_t ncvx;
  // OBBRSS BVH models (nodes are the 256-B records of include/hppfcl_b200.h)
  const hfb_bvh_node* bvh_nodes;
  const double* bvh_verts;
  const uint32_t* bvh_tris;
  const BvhDesc* bvh_desc;
  uint32_t nbvh;
  // aabb_local of every shape handle (min xyz, max xyz; NaN for node types without one): the broadphase feed
  const double* local_aabbs;
};

template <int CAPS>
HFB_HD ShapeD load_shape(const ArenaView& A, uint32_t h) {
  ShapeD s;
  if (h >= A.nshapes) {  // a handle the arena never issued (device entry points trust their caller only this far):
    s.type = 0;          // no such geometry -> the pair is reported as HFB_PATH_UNSUPPORTED
    s.p0 = s.p1 = s.p2 = s.ssr = 0;
    s.cx = s.cy = s.cz = nullptr;
    s.nv = 0;
    s.center = s.ta = s.tb = s.tc = mk(0, 0, 0);
    return s;
  }
  const hfb_shape& r = A.shapes[h];
  s.type = (int)r.type;
  s.p0 = r.p[0];
  s.p1 = r.p[1];
  s.p2 = r.p[2];
  s.ssr = r.ssr;
  s.cx = s.cy = s.cz = nullptr;
  s.nv = 0;
  s.center = mk(0, 0, 0);
  s.ta = s.tb = s.tc = mk(0, 0, 0);
  if ((CAPS & CAP_PLANE) && (s.type == HFB_GEOM_PLANE || s.type == HFB_GEOM_HALFSPACE)) {
    s.center.x = A.pool[r.data];  // the offset d (the record's p holds the unit normal): HostArena::add_halfspace
  }
  if ((CAPS & (CAP_CONVEX | CAP_TRI)) && (s.type == HFB_GEOM_CONVEX || s.type == HFB_GEOM_TRIANGLE)) {
    const ConvexDesc& d = A.cvx[r.data];
    const double* b = A.pool + d.off;
    if ((CAPS & CAP_TRI) && s.type == HFB_GEOM_TRIANGLE) {
      s.ta = mk(b[0], b[d.vpad], b[2 * d.vpad]);
      s.tb = mk(b[1], b[d.vpad + 1], b[2 * d.vpad + 1]);
      s.tc = mk(b[2], b[d.vpad + 2], b[2 * d.vpad + 2]);
    } else {
      s.cx = b;
      s.cy = b + d.vpad;
      s.cz = b + 2 * d.vpad;
      s.nv = (int)d.nv;
      s.center = mk(d.cx, d.cy, d.cz);
    }
  }
  return s;
}

// host-side builder (plain C++, used by the C-ABI and by the CPU emulation harness)
struct HostArena {
  std::vector<hfb_shape> shapes;
  std::vector<ConvexDesc> cvx;
  std::vector<double> pool;
  std::vector<hfb_bvh_node> bvh_nodes;
  std::vector<double> bvh_verts;
  std::vector<uint32_t> bvh_tris;
  std::vector<BvhDesc> bvh_desc;
  bool has_convex = false, has_tri = false, has_unknown = false, has_bvh = false;

  // validates the tree (child links, every node but the root referenced once, leaf primitive ids, triangle
  // vertex ids, depth within the walks' stacks) and stores it.  The stored copy of a node carries in `_pad` the
  // number of triangles below it when its descendants are ONE contiguous block of the array -- what
  // recursiveBuildTree produces (BVH_model.cpp:860-960: the two children are allocated together, then the left
  // subtree is built completely before the right one) -- and 0 otherwise; the walk of hfb_bvhq.cuh speculates only
  // on such subtrees.
  int max_bvh_depth = 0;  // deepest registered tree (root = depth 0)
  bool add_bvh(const hfb_bvh_node* nodes, uint32_t nn, const double* verts, uint32_t nv, const uint32_t* tris,
               uint32_t nt, uint32_t* id, uint32_t kind = 0 /* 0: BVHModel<OBBRSS>, 1: BVHModel<OBB> */) {
    if (nn == 0 || nv == 0 || nt == 0) return false;
    std::vector<uint8_t> refs(nn, 0);
    for (uint32_t i = 0; i < nn; ++i) {
      const int fc = nodes[i].first_child;
      if (fc < 0) {
        if ((uint32_t)(-(fc + 1)) >= nt) return false;
      } else if ((uint32_t)fc + 1 >= nn || (uint32_t)fc <= i) {
        return false;  // children follow their parent in BVHModel::bvs (recursiveBuildTree)
      } else {
        if (refs[fc] || refs[fc + 1]) return false;
        refs[fc] = refs[fc + 1] = 1;
      }
    }
    for (uint32_t i = 1; i < nn; ++i)
      if (!refs[i]) return false;
    for (uint32_t i = 0; i < 3 * nt; ++i)
      if (tris[i] >= nv) return false;
    // depth (parents precede their children) -- the walks keep explicit stacks of HFB_BVH_STACK = 128 entries:
    // a mesh-shape walk needs depth + 2, a mesh-mesh walk depth1 + depth2 + 2
    std::vector<int> depth(nn, 0);
    int maxd = 0;
    for (uint32_t i = 0; i < nn; ++i) {
      const int fc = nodes[i].first_child;
      if (fc >= 0) {
        depth[fc] = depth[fc + 1] = depth[i] + 1;
        if (depth[i] + 1 > maxd) maxd = depth[i] + 1;
      }
    }
    if (maxd > 62) return false;
    // leaves below every node and the last index of its subtree, children first
    std::vector<uint32_t> leaves(nn, 1), last(nn, 0);
    for (uint32_t k = nn; k-- > 0;) {
      const int fc = nodes[k].first_child;
      last[k] = k;
      if (fc >= 0) {
        leaves[k] = leaves[fc] + leaves[fc + 1];
        last[k] = last[fc] > last[fc + 1] ? last[fc] : last[fc + 1];
      }
    }
    BvhDesc d;
    d.node_off = (uint32_t)bvh_nodes.size();
    d.num_nodes = nn;
    d.vert_off = (uint32_t)(bvh_verts.size() / 3);
    d.num_verts = nv;
    d.tri_off = (uint32_t)(bvh_tris.size() / 3);
    d.num_tris = nt;
    d._r0 = (uint32_t)maxd;
    d._r1 = kind;
    bvh_nodes.insert(bvh_nodes.end(), nodes, nodes + nn);
    for (uint32_t k = 0; k < nn; ++k) {
      const int fc = nodes[k].first_child;
      uint32_t mark = 0;
      if (fc < 0) mark = 1;
      else if (last[k] - (uint32_t)fc + 1 == 2 * leaves[k] - 2) mark = leaves[k];  // exactly the block [fc, last]
      bvh_nodes[d.node_off + k]._pad = mark;
    }
    if (maxd > max_bvh_depth) max_bvh_depth = maxd;
    bvh_verts.insert(bvh_verts.end(), verts, verts + 3 * (size_t)nv);
    bvh_tris.insert(bvh_tris.end(), tris, tris + 3 * (size_t)nt);
    bvh_desc.push_back(d);
    *id = (uint32_t)bvh_desc.size() - 1;
    return true;
  }

  uint32_t add_convex(const double* pts, uint32_t n) {
    ConvexDesc d;
    d.off = (uint32_t)pool.size();
    d.nv = n;
    d.vpad = (n + 1u) & ~1u;
    d._r = 0;
    double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    pool.resize(pool.size() + 3 * (size_t)d.vpad, 0.0);
    for (uint32_t i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) {
        const double v = pts[3 * i + k];
        pool[d.off + (size_t)k * d.vpad + i] = v;
        mn[k] = v < mn[k] ? v : mn[k];
        mx[k] = v > mx[k] ? v : mx[k];
      }
    // padding replicates vertex 0 so a padded lane can never win the argmax
    for (uint32_t i = n; i < d.vpad; ++i)
      for (int k = 0; k < 3; ++k) pool[d.off + (size_t)k * d.vpad + i] = pts[k];
    d.cx = (mn[0] + mx[0]) * 0.5;  // AABB::center() = (min_ + max_) * 0.5
    d.cy = (mn[1] + mx[1]) * 0.5;
    d.cz = (mn[2] + mx[2]) * 0.5;
    cvx.push_back(d);
    return (uint32_t)cvx.size() - 1;
  }
  // vertex set `id` replaced by another one of the same size (a deformed / re-scaled hull)
  bool set_convex(uint32_t id, const double* pts, uint32_t n) {
    if (id >= cvx.size() || cvx[id].nv != n || n == 0) return false;
    ConvexDesc& d = cvx[id];
    double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (uint32_t i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) {
        const double v = pts[3 * i + k];
        pool[d.off + (size_t)k * d.vpad + i] = v;
        mn[k] = v < mn[k] ? v : mn[k];
        mx[k] = v > mx[k] ? v : mx[k];
      }
    for (uint32_t i = n; i < d.vpad; ++i)
      for (int k = 0; k < 3; ++k) pool[d.off + (size_t)k * d.vpad + i] = pts[k];
    d.cx = (mn[0] + mx[0]) * 0.5;
    d.cy = (mn[1] + mx[1]) * 0.5;
    d.cz = (mn[2] + mx[2]) * 0.5;
    return true;
  }
  // record of handle `h` replaced (same validation as add_shape); a record of type 0 retires the handle:
  // pairs that name it come back as HFB_PATH_UNSUPPORTED
  bool set_shape(uint32_t h, const hfb_shape& s) {
    if (h >= shapes.size()) return false;
    uint32_t tmp;  // run the record through add_shape's checks and flags, then move it into place
    if (!add_shape(s, &tmp)) return false;
    shapes[h] = shapes.back();
    shapes.pop_back();
    return true;
  }
  bool valid_shape(const hfb_shape& s) const {
    if (s.type == HFB_BV_OBBRSS) return s.data < bvh_desc.size() && bvh_desc[s.data]._r1 == 0;
    if (s.type == HFB_BV_OBB) return s.data < bvh_desc.size() && bvh_desc[s.data]._r1 == 1;
    if (s.type == HFB_GEOM_CONVEX) return s.data < cvx.size();
    if (s.type == HFB_GEOM_TRIANGLE) return s.data < cvx.size() && cvx[s.data].nv >= 3;
    if (s.type == HFB_GEOM_PLANE || s.type == HFB_GEOM_HALFSPACE) return s.data < pool.size();  // (only add_halfspace makes these)
    return true;
  }
  // Halfspace(n, d) / Plane(n, d) (geometric_shapes.h:885-1031): the constructors normalise (n, d)
  // (unitNormalTest, geometric_shapes.cpp:121-143); the record keeps the unit normal in p and the offset in the pool
  bool add_halfspace(uint32_t type, const double* n_in, double d, double ssr, uint32_t* handle) {
    if (type != HFB_GEOM_PLANE && type != HFB_GEOM_HALFSPACE) return false;
    double n[3] = {n_in[0], n_in[1], n_in[2]};
    const double l = sqrt((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]);
    if (l > 0) {
      const double inv_l = 1.0 / l;
      for (int k = 0; k < 3; ++k) n[k] *= inv_l;
      d *= inv_l;
    } else {
      n[0] = 1;
      n[1] = n[2] = 0;
      d = 0;
    }
    hfb_shape s;
    s.type = type;
    s.data = (uint32_t)pool.size();
    pool.push_back(d);
    pool.push_back(0.0);  // (the pool stays a whole number of 16-byte units: the hull blocks are copied in bulk)
    s.p[0] = n[0];
    s.p[1] = n[1];
    s.p[2] = n[2];
    s.ssr = ssr;
    return add_shape(s, handle);
  }
  // returns false on an invalid record
  bool add_shape(const hfb_shape& s, uint32_t* handle) {
    if (!valid_shape(s)) return false;
    if (s.type == HFB_BV_OBBRSS || s.type == HFB_BV_OBB) {
      has_bvh = true;
    } else if (s.type == HFB_GEOM_CONVEX || s.type == HFB_GEOM_TRIANGLE) {
      if (s.type == HFB_GEOM_CONVEX) has_convex = true; else has_tri = true;
    } else if (!(s.type == HFB_GEOM_BOX || s.type == HFB_GEOM_SPHERE || s.type == HFB_GEOM_CAPSULE ||
                 s.type == HFB_GEOM_CONE || s.type == HFB_GEOM_CYLINDER || s.type == HFB_GEOM_ELLIPSOID ||
                 s.type == HFB_GEOM_PLANE || s.type == HFB_GEOM_HALFSPACE)) {
      has_unknown = true;  // octree / height field / ...: reported per pair as unsupported
    }
    shapes.push_back(s);
    *handle = (uint32_t)shapes.size() - 1;
    return true;
  }
  ArenaView view() const {
    ArenaView v;
    v.shapes = shapes.data();
    v.cvx = cvx.data();
    v.pool = pool.data();
    v.nshapes = (uint32_t)shapes.size();
    v.ncvx = (uint32_t)cvx.size();
    v.bvh_nodes = bvh_nodes.data();
    v.bvh_verts = bvh_verts.data();
    v.bvh_tris = bvh_tris.data();
    v.bvh_desc = bvh_desc.data();
    v.nbvh = (uint32_t)bvh_desc.size();
    v.local_aabbs = nullptr;
    return v;
  }
};

}  // namespace hfb
