// ORACLE / TEST INFRASTRUCTURE ONLY.
//
// C entry points over the REFERENCE ITSELF: the hpp-fcl sources under /root/reference, compiled in
// place and unmodified into oracle/_ref/libhppfcl_ref.so (recipe: `make -C oracle ref`).  Eigen and
// Boost are not installed in this image; oracle/ref_shim/ holds a minimal stand-in for the fixed-size
// Eigen API those sources use and for the three headers their build system would generate.  The
// stand-in fixes the floating-point conventions, the ALGORITHMS are the
// reference's own code -- which is what this library is for: checking the oracle's restatement of
// them, and serving as `cpu_baseline.kind = "reference"` in bench.py.
//
// Same PODs and the same calling shape as oracle/capi.cpp.  Per pair this is exactly
//   hpp::fcl::distance() / collide()   (src/distance.cpp:60-109, src/collision.cpp:69-130)
// on fresh result objects; the ComputeDistance / ComputeCollision functors (the same code path with the
// solver kept as a member) are used so that the solver's status and iteration counters can be read back.
#include <cstdio>
#include <cstdlib>
#include <hpp/fcl/BVH/BVH_model.h>
#include <hpp/fcl/collision.h>
#include <hpp/fcl/distance.h>
#include <hpp/fcl/shape/convex.h>
#include <hpp/fcl/shape/geometric_shapes.h>

#include <cstring>
#include <memory>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/hppfcl_b200.h"

using namespace hpp::fcl;

namespace {
struct Scene {
  std::vector<std::shared_ptr<std::vector<Vec3f>>> cvx_points;
  std::vector<std::shared_ptr<std::vector<Triangle>>> cvx_tris;
  std::vector<std::shared_ptr<BVHModel<OBBRSS>>> bvhs;
  std::vector<std::shared_ptr<BVHModel<OBB>>> bvhs_obb;  // the same meshes as plain OBB models (same ids)
  std::vector<std::shared_ptr<CollisionGeometry>> geoms;
};
struct PeekDistance : ComputeDistance {
  PeekDistance(const CollisionGeometry* a, const CollisionGeometry* b) : ComputeDistance(a, b) {}
  const GJKSolver& s() const { return solver; }
};
struct PeekCollision : ComputeCollision {
  PeekCollision(const CollisionGeometry* a, const CollisionGeometry* b) : ComputeCollision(a, b) {}
  const GJKSolver& s() const { return solver; }
};
Transform3f to_tf(const hfb_transform& t) {
  Matrix3f R;
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) R(r, c) = t.R[3 * c + r];
  return Transform3f(R, Vec3f(t.T[0], t.T[1], t.T[2]));
}
void put3(double* o, const Vec3f& v) {
  o[0] = v[0];
  o[1] = v[1];
  o[2] = v[2];
}
template <class Req>
void fill_query(const hfb_query_request& q, size_t i, Req& r) {
  r.gjk_initial_guess = (GJKInitialGuess)q.gjk_initial_guess;
  r.gjk_variant = (GJKVariant)q.gjk_variant;
  r.gjk_convergence_criterion = (GJKConvergenceCriterion)q.gjk_convergence_criterion;
  r.gjk_convergence_criterion_type = (GJKConvergenceCriterionType)q.gjk_convergence_criterion_type;
  r.gjk_max_iterations = q.gjk_max_iterations;
  r.epa_max_iterations = q.epa_max_iterations;
  r.gjk_tolerance = q.gjk_tolerance;
  r.epa_tolerance = q.epa_tolerance;
  r.collision_distance_threshold = q.collision_distance_threshold;
  if (q.gjk_initial_guess == HFB_GUESS_CACHED) {
    if (q.cached_gjk_guess)
      r.cached_gjk_guess = Vec3f(q.cached_gjk_guess[3 * i], q.cached_gjk_guess[3 * i + 1], q.cached_gjk_guess[3 * i + 2]);
    if (q.cached_support_func_guess) {
      r.cached_support_func_guess[0] = q.cached_support_func_guess[2 * i];
      r.cached_support_func_guess[1] = q.cached_support_func_guess[2 * i + 1];
    }
  }
}
// status word of the batch records (include/hppfcl_b200.h): gjk | epa << 8 | path << 16.  The
// TriangleP-TriangleP specialisation is recorded on the closed-form path with its GJK status and no
// iteration count -- a convention of the records, the reference has no such field.
unsigned pack(const GJKSolver& s, bool bvh, bool tri_tri) {
  if (bvh) return (unsigned)HFB_PATH_BVH << 16;
  const bool closed = tri_tri || s.gjk.status == details::GJK::DidNotRun;
  return ((unsigned)s.gjk.status & 0xffu) | (((unsigned)(int)s.epa.status & 0xffu) << 8) |
         ((unsigned)(closed ? HFB_PATH_CLOSED_FORM : HFB_PATH_GJK) << 16);
}
}  // namespace

extern "C" {

void* ref_scene_create() {
  // getShapeSupport<Box> keeps a function-local static `inflate` that the FIRST direction ever queried
  // in the process fixes (support_functions.cpp:146).  A process whose first box query starts from the
  // default guess (1,0,0) asks for (-1,0,0) and gets 1 + 1e-10 -- the value the oracle and the kernels
  // use.  Make that first query here, before any (possibly multi-threaded) batch does.
  static bool primed = false;
  if (!primed) {
    primed = true;
    Box a(1, 1, 1), b(1, 1, 1);
    DistanceRequest rq;
    DistanceResult rs;
    distance(&a, Transform3f(), &b, Transform3f(Vec3f(3, 0, 0)), rq, rs);
  }
  return new Scene();
}
void ref_scene_destroy(void* p) { delete static_cast<Scene*>(p); }

// ConvexBase: Convex<Triangle> (include/hpp/fcl/shape/convex.h); fillNeighbors runs in the constructor
int ref_register_convex(void* p, const double* pts, uint32_t n, const uint32_t* tris, uint32_t nt) {
  Scene* s = static_cast<Scene*>(p);
  auto P = std::make_shared<std::vector<Vec3f>>(n);
  for (uint32_t i = 0; i < n; ++i) (*P)[i] = Vec3f(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
  auto T = std::make_shared<std::vector<Triangle>>(nt);
  for (uint32_t i = 0; i < nt; ++i) (*T)[i].set(tris[3 * i], tris[3 * i + 1], tris[3 * i + 2]);
  s->cvx_points.push_back(P);
  s->cvx_tris.push_back(T);
  return (int)s->cvx_points.size() - 1;
}

int ref_register_bvh(void* p, const double* verts, uint32_t nv, const uint32_t* tris, uint32_t nt) {
  Scene* s = static_cast<Scene*>(p);
  std::vector<Vec3f> V(nv);
  for (uint32_t i = 0; i < nv; ++i) V[i] = Vec3f(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]);
  std::vector<Triangle> T(nt);
  for (uint32_t i = 0; i < nt; ++i) T[i].set(tris[3 * i], tris[3 * i + 1], tris[3 * i + 2]);
  auto m = std::make_shared<BVHModel<OBBRSS>>();
  m->beginModel();
  m->addSubModel(V, T);
  m->endModel();
  s->bvhs.push_back(m);
  auto mo = std::make_shared<BVHModel<OBB>>();
  mo->beginModel();
  mo->addSubModel(V, T);
  mo->endModel();
  s->bvhs_obb.push_back(mo);
  return (int)s->bvhs.size() - 1;
}

// 1 when the plain BVHModel<OBB> of mesh `id` equals, node for node and bit for bit, the OBB half of its
// BVHModel<OBBRSS> (what lets one node array serve both kinds in the product)
int ref_bvh_obb_is_obbrss_half(void* p, int id) {
  Scene* s = static_cast<Scene*>(p);
  const BVHModel<OBBRSS>& a = *s->bvhs[(size_t)id];
  const BVHModel<OBB>& b = *s->bvhs_obb[(size_t)id];
  if (a.getNumBVs() != b.getNumBVs()) return 0;
  for (unsigned i = 0; i < a.getNumBVs(); ++i) {
    const BVNode<OBBRSS>& x = a.getBV(i);
    const BVNode<OBB>& y = b.getBV(i);
    if (x.first_child != y.first_child || x.first_primitive != y.first_primitive || x.num_primitives != y.num_primitives) return 0;
    if (std::memcmp(x.bv.obb.axes.data(), y.bv.axes.data(), 9 * sizeof(double)) != 0) return 0;
    for (int k = 0; k < 3; ++k)
      if (x.bv.obb.To[k] != y.bv.To[k] || x.bv.obb.extent[k] != y.bv.extent[k]) return 0;
  }
  return 1;
}

// BVHModel<OBBRSS>::bvs as hfb_bvh_node records; returns the node count
int ref_bvh_export(void* p, int id, hfb_bvh_node* out, uint32_t cap) {
  Scene* s = static_cast<Scene*>(p);
  const BVHModel<OBBRSS>& m = *s->bvhs[(size_t)id];
  const unsigned n = m.getNumBVs();
  if (!out) return (int)n;
  for (unsigned i = 0; i < n && i < cap; ++i) {
    const BVNode<OBBRSS>& b = m.getBV(i);
    hfb_bvh_node& o = out[i];
    std::memset(&o, 0, sizeof(o));
    o.first_child = b.first_child;
    o.first_primitive = b.first_primitive;
    o.num_primitives = b.num_primitives;
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r) {
        o.obb_axes[3 * c + r] = b.bv.obb.axes(r, c);
        o.rss_axes[3 * c + r] = b.bv.rss.axes(r, c);
      }
    put3(o.obb_To, b.bv.obb.To);
    put3(o.obb_extent, b.bv.obb.extent);
    put3(o.rss_Tr, b.bv.rss.Tr);
    o.rss_length[0] = b.bv.rss.length[0];
    o.rss_length[1] = b.bv.rss.length[1];
    o.rss_radius = b.bv.rss.radius;
  }
  return (int)n;
}

int64_t ref_register_shapes(void* p, const hfb_shape* recs, size_t n) {
  Scene* s = static_cast<Scene*>(p);
  const int64_t first = (int64_t)s->geoms.size();
  for (size_t i = 0; i < n; ++i) {
    const hfb_shape& r = recs[i];
    std::shared_ptr<CollisionGeometry> g;
    switch (r.type) {
      case HFB_GEOM_BOX: g.reset(new Box(2 * r.p[0], 2 * r.p[1], 2 * r.p[2])); break;
      case HFB_GEOM_SPHERE: g.reset(new Sphere(r.p[0])); break;
      case HFB_GEOM_ELLIPSOID: g.reset(new Ellipsoid(r.p[0], r.p[1], r.p[2])); break;
      case HFB_GEOM_CAPSULE: g.reset(new Capsule(r.p[0], 2 * r.p[1])); break;
      case HFB_GEOM_CONE: g.reset(new Cone(r.p[0], 2 * r.p[1])); break;
      case HFB_GEOM_CYLINDER: g.reset(new Cylinder(r.p[0], 2 * r.p[1])); break;
      case HFB_GEOM_PLANE: g.reset(new Plane(Vec3f(0, 0, 1), 0)); break;
      case HFB_GEOM_HALFSPACE: g.reset(new Halfspace(Vec3f(0, 0, 1), 0)); break;
      case HFB_GEOM_CONVEX: {
        if (r.data >= s->cvx_points.size()) return -1;
        auto& P = s->cvx_points[r.data];
        auto& T = s->cvx_tris[r.data];
        g.reset(new Convex<Triangle>(P, (unsigned)P->size(), T, (unsigned)T->size()));
      } break;
      case HFB_GEOM_TRIANGLE: {
        if (r.data >= s->cvx_points.size()) return -1;
        const std::vector<Vec3f>& P = *s->cvx_points[r.data];
        g.reset(new TriangleP(P[0], P[1], P[2]));
      } break;
      case HFB_BV_OBBRSS:
        if (r.data >= s->bvhs.size()) return -1;
        g = s->bvhs[r.data];
        break;
      case HFB_BV_OBB:
        if (r.data >= s->bvhs_obb.size()) return -1;
        g = s->bvhs_obb[r.data];
        break;
      default: return -1;
    }
    if (ShapeBase* sb = dynamic_cast<ShapeBase*>(g.get()))
      if (r.ssr > 0) sb->setSweptSphereRadius(r.ssr);
    if (r.type != HFB_BV_OBBRSS && r.type != HFB_BV_OBB) g->computeLocalAABB();
    s->geoms.push_back(g);
  }
  return first;
}

// Halfspace(n, d) / Plane(n, d); the first call also fixes the `inflate` static of the box support's
// WithSweptSphere instantiation (support_functions.cpp:146) with an axis-aligned direction, the case the product
// computes with (include/hppfcl_b200.h, hfb_geom_register_halfspaces)
int64_t ref_register_halfspaces(void* p, uint32_t type, const double* nd, const double* ssr, size_t count) {
  Scene* s = static_cast<Scene*>(p);
  if (type != HFB_GEOM_PLANE && type != HFB_GEOM_HALFSPACE) return -1;
  static bool primed = false;
  if (!primed) {
    primed = true;
    Box b(1, 1, 1);
    int hint = 0;
    (void)details::getSupport<details::SupportOptions::WithSweptSphere>(&b, Vec3f(0, 0, -1), hint);
  }
  const int64_t first = (int64_t)s->geoms.size();
  for (size_t i = 0; i < count; ++i) {
    const Vec3f n(nd[4 * i], nd[4 * i + 1], nd[4 * i + 2]);
    std::shared_ptr<CollisionGeometry> g;
    if (type == HFB_GEOM_PLANE) g.reset(new Plane(n, nd[4 * i + 3]));
    else g.reset(new Halfspace(n, nd[4 * i + 3]));
    if (ssr && ssr[i] > 0) static_cast<ShapeBase*>(g.get())->setSweptSphereRadius(ssr[i]);
    g->computeLocalAABB();
    s->geoms.push_back(g);
  }
  return first;
}

// CollisionObject::computeAABB (collision_object.h:258-278) of geometry h[i] at pose tf[i]: min xyz, max xyz
int ref_object_aabbs(void* p, size_t n, const uint32_t* h, const hfb_transform* tf, double* out) {
  Scene* s = static_cast<Scene*>(p);
  for (size_t i = 0; i < n; ++i) {
    if (h[i] >= s->geoms.size()) return HFB_ERR_INVALID_ARGUMENT;
    CollisionObject o(s->geoms[h[i]], to_tf(tf[i]));
    const AABB& a = o.getAABB();
    for (int k = 0; k < 3; ++k) {
      out[6 * i + k] = a.min_[k];
      out[6 * i + 3 + k] = a.max_[k];
    }
  }
  return HFB_OK;
}

int ref_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int ref_batch_distance(void* p, size_t n, const uint32_t* h1, const hfb_transform* tf1, const uint32_t* h2,
                       const hfb_transform* tf2, const hfb_distance_request* req, hfb_distance_result* out,
                       const hfb_guess_out* go, int nthreads) {
  Scene* s = static_cast<Scene*>(p);
  for (size_t i = 0; i < n; ++i)
    if (h1[i] >= s->geoms.size() || h2[i] >= s->geoms.size()) return HFB_ERR_INVALID_ARGUMENT;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
  // debugging aids, read once: nothing but the reference's own work runs inside the timed loop
  const char* dump_env = getenv("HFB_REF_DUMP_SIMPLEX");
  const long long dump_pair = dump_env ? atoll(dump_env) : -1;
  const bool verbose = getenv("HFB_REF_VERBOSE") != nullptr;
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (long long ii = 0; ii < (long long)n; ++ii) {
    const size_t i = (size_t)ii;
    const CollisionGeometry* o1 = s->geoms[h1[i]].get();
    const CollisionGeometry* o2 = s->geoms[h2[i]].get();
    hfb_distance_result& r = out[i];
    DistanceRequest rq(req->enable_nearest_points != 0, req->enable_signed_distance != 0, req->rel_err, req->abs_err);
    fill_query(req->q, i, rq);
    DistanceResult res;
    unsigned status = (unsigned)HFB_PATH_UNSUPPORTED << 16, iters = 0;
    try {
      PeekDistance f(o1, o2);
      f(to_tf(tf1[i]), to_tf(tf2[i]), rq, res);
      const bool bvh = o1->getObjectType() == OT_BVH || o2->getObjectType() == OT_BVH;
      const bool tt = o1->getNodeType() == GEOM_TRIANGLE && o2->getNodeType() == GEOM_TRIANGLE;
      status = pack(f.s(), bvh, tt);
      if (!bvh && !tt && f.s().gjk.status != details::GJK::DidNotRun)
        iters = (unsigned)(f.s().gjk.getNumIterations() & 0xffff) | ((unsigned)(f.s().epa.getNumIterations() & 0xffff) << 16);
      if (dump_pair == ii) {  // debugging aid
        const details::GJK::Simplex* sx = f.s().gjk.getSimplex();
        fprintf(stderr, "ref simplex rank %d ray %.17g %.17g %.17g\n", (int)sx->rank, f.s().gjk.ray[0], f.s().gjk.ray[1], f.s().gjk.ray[2]);
        for (int k = 0; k < (int)sx->rank; ++k) {
          const details::GJK::SimplexV* v = sx->vertex[k];
          fprintf(stderr, "  w0 %.17g %.17g %.17g  w1 %.17g %.17g %.17g\n", v->w0[0], v->w0[1], v->w0[2], v->w1[0], v->w1[1], v->w1[2]);
        }
      }
    } catch (const std::exception& e) {
      if (verbose) fprintf(stderr, "reference threw (pair %zu): %s\n", i, e.what());
      res.clear();
    }
    r.min_distance = res.min_distance;
    put3(r.p1, res.nearest_points[0]);
    put3(r.p2, res.nearest_points[1]);
    put3(r.normal, res.normal);
    r.b1 = res.b1;
    r.b2 = res.b2;
    r.status = status;
    r.iterations = iters;
    if (go) {
      if (go->cached_gjk_guess) put3(go->cached_gjk_guess + 3 * i, res.cached_gjk_guess);
      if (go->cached_support_func_guess) {
        go->cached_support_func_guess[2 * i] = res.cached_support_func_guess[0];
        go->cached_support_func_guess[2 * i + 1] = res.cached_support_func_guess[1];
      }
    }
  }
  return HFB_OK;
}

static hfb_contact* g_extra = nullptr;  // ref_batch_collide_contacts
static uint32_t* g_counts = nullptr;
static uint32_t g_max_extra = 0;

int ref_batch_collide(void* p, size_t n, const uint32_t* h1, const hfb_transform* tf1, const uint32_t* h2,
                      const hfb_transform* tf2, const hfb_collision_request* req, hfb_contact* out,
                      const hfb_guess_out* go, int nthreads) {
  Scene* s = static_cast<Scene*>(p);
  for (size_t i = 0; i < n; ++i)
    if (h1[i] >= s->geoms.size() || h2[i] >= s->geoms.size()) return HFB_ERR_INVALID_ARGUMENT;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (long long ii = 0; ii < (long long)n; ++ii) {
    const size_t i = (size_t)ii;
    const CollisionGeometry* o1 = s->geoms[h1[i]].get();
    const CollisionGeometry* o2 = s->geoms[h2[i]].get();
    hfb_contact& r = out[i];
    std::memset(&r, 0, sizeof(r));
    CollisionRequest rq;
    rq.num_max_contacts = req->num_max_contacts;
    rq.enable_contact = req->enable_contact != 0;
    rq.security_margin = req->security_margin;
    rq.break_distance = req->break_distance;
    rq.distance_upper_bound = req->distance_upper_bound;
    fill_query(req->q, i, rq);
    CollisionResult res;
    unsigned status = (unsigned)HFB_PATH_UNSUPPORTED << 16, iters = 0;
    try {
      if (req->security_margin == -std::numeric_limits<double>::infinity()) {
        res.clear();  // collision.cpp:73-76
        status = 0;
      } else {
        PeekCollision f(o1, o2);
        f(to_tf(tf1[i]), to_tf(tf2[i]), rq, res);
        const bool bvh = o1->getObjectType() == OT_BVH || o2->getObjectType() == OT_BVH;
        const bool tt = o1->getNodeType() == GEOM_TRIANGLE && o2->getNodeType() == GEOM_TRIANGLE;
        status = pack(f.s(), bvh, tt);
        if (!bvh && !tt && f.s().gjk.status != details::GJK::DidNotRun)
          iters = (unsigned)(f.s().gjk.getNumIterations() & 0xffff) | ((unsigned)(f.s().epa.getNumIterations() & 0xffff) << 16);
      }
    } catch (const std::exception& e) {
      if (getenv("HFB_REF_VERBOSE")) fprintf(stderr, "reference threw (pair %zu): %s\n", i, e.what());
      res.clear();
    }
    r.distance = std::numeric_limits<double>::max();
    r.distance_lower_bound = res.distance_lower_bound;
    put3(r.p1, res.nearest_points[0]);
    put3(r.p2, res.nearest_points[1]);
    put3(r.normal, res.normal);
    const double q = std::numeric_limits<double>::quiet_NaN();
    r.pos[0] = r.pos[1] = r.pos[2] = q;
    r.b1 = r.b2 = -1;
    r.num_contacts = res.numContacts() ? 1u : 0u;
    if (res.numContacts()) {
      const Contact& c = res.getContact(0);
      r.distance = c.penetration_depth;
      r.b1 = c.b1;
      r.b2 = c.b2;
      put3(r.pos, c.pos);
      put3(r.p1, c.nearest_points[0]);
      put3(r.p2, c.nearest_points[1]);
      put3(r.normal, c.normal);
    }
    r.status = status;
    r.iterations = iters;
    if (g_counts) g_counts[i] = (uint32_t)res.numContacts();
    if (g_extra)
      for (size_t k = 1; k < res.numContacts() && k - 1 < g_max_extra; ++k) {
        const Contact& c = res.getContact(k);
        hfb_contact& x = g_extra[i * g_max_extra + (k - 1)];
        std::memset(&x, 0, sizeof(x));
        x.distance_lower_bound = std::numeric_limits<double>::max();
        x.num_contacts = 1;
        x.distance = c.penetration_depth;
        x.b1 = c.b1;
        x.b2 = c.b2;
        put3(x.pos, c.pos);
        put3(x.p1, c.nearest_points[0]);
        put3(x.p2, c.nearest_points[1]);
        put3(x.normal, c.normal);
        x.status = status;
      }
    if (go) {
      if (go->cached_gjk_guess) put3(go->cached_gjk_guess + 3 * i, res.cached_gjk_guess);
      if (go->cached_support_func_guess) {
        go->cached_support_func_guess[2 * i] = res.cached_support_func_guess[0];
        go->cached_support_func_guess[2 * i + 1] = res.cached_support_func_guess[1];
      }
    }
  }
  return HFB_OK;
}


int ref_batch_collide_contacts(void* p, size_t n, const uint32_t* h1, const hfb_transform* tf1, const uint32_t* h2,
                               const hfb_transform* tf2, const hfb_collision_request* req, hfb_contact* out,
                               uint32_t max_extra, hfb_contact* extra, uint32_t* counts, int nthreads) {
  g_extra = max_extra ? extra : nullptr;
  g_counts = counts;
  g_max_extra = max_extra;
  const int rc = ref_batch_collide(p, n, h1, tf1, h2, tf2, req, out, nullptr, nthreads);
  g_extra = nullptr;
  g_counts = nullptr;
  return rc;
}

}  // extern "C"
