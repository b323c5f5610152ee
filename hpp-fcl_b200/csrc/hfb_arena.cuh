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

struct ArenaView {
  const hfb_shape* shapes;
  const ConvexDesc* cvx;
  const double* pool;
  uint32_t nshapes;
  uint32_t ncvx;
};

template <int CAPS>
HFB_HD ShapeD load_shape(const ArenaView& A, uint32_t h) {
  ShapeD s;
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
  bool has_convex = false, has_tri = false, has_unknown = false;

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
  // returns false on an invalid record
  bool add_shape(const hfb_shape& s, uint32_t* handle) {
    if (s.type == HFB_GEOM_CONVEX || s.type == HFB_GEOM_TRIANGLE) {
      if (s.data >= cvx.size()) return false;
      if (s.type == HFB_GEOM_TRIANGLE && cvx[s.data].nv < 3) return false;
      if (s.type == HFB_GEOM_CONVEX) has_convex = true; else has_tri = true;
    } else if (!(s.type == HFB_GEOM_BOX || s.type == HFB_GEOM_SPHERE || s.type == HFB_GEOM_CAPSULE ||
                 s.type == HFB_GEOM_CONE || s.type == HFB_GEOM_CYLINDER || s.type == HFB_GEOM_ELLIPSOID)) {
      has_unknown = true;  // plane / halfspace / ...: reported per pair as unsupported
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
    return v;
  }
};

}  // namespace hfb
