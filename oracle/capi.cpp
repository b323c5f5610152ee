// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.hpp).
// C entry points so tests / bench.py (ctypes) can run the CPU restatement on the
// same POD buffers the product C-ABI (include/hppfcl_b200.h) takes.
// distance():  src/distance.cpp:60-109 + ShapeShapeDistancer::run (shape_shape_func.h:51-82)
// collide():   src/collision.cpp:69-130 + ShapeShapeCollider::run (shape_shape_func.h:132-164)
#include <stdexcept>
#include <cstring>
#include <memory>
#include <string>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "bvh.hpp"
#include "oracle.hpp"

using namespace oracle;

namespace {
struct Scene {
  std::vector<std::unique_ptr<Convex>> convexes;
  std::vector<std::unique_ptr<BVHModel>> bvhs;
  std::vector<Shape> shapes;  // type HFB_BV_OBBRSS: p[0] holds the BVH index
};

inline void put3(double* o, const V3& v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }

inline uint32_t pack_status(const GJKSolver& s, bool closed_form) {
  uint32_t g = (uint32_t)s.gjk.status & 0xffu;
  uint32_t e = (uint32_t)((int)s.epa.status) & 0xffu;
  uint32_t p = closed_form ? HFB_PATH_CLOSED_FORM : HFB_PATH_GJK;
  return g | (e << 8) | (p << 16);
}
inline uint32_t pack_iters(const GJKSolver& s) {
  return (uint32_t)(s.gjk.iterations & 0xffff) | ((uint32_t)(s.epa.iterations & 0xffff) << 16);
}
}  // namespace

extern "C" {

void* oracle_scene_create() { return new Scene(); }
void oracle_scene_destroy(void* s) { delete static_cast<Scene*>(s); }

// points: n x 3. tris: ntris x 3 vertex ids of the hull faces (may be NULL: no
// neighbours => always the linear support, like a ConvexBase without neighbors).
int oracle_register_convex(void* sc, const double* points, uint32_t n, const uint32_t* tris,
                           uint32_t ntris) {
  Scene* s = static_cast<Scene*>(sc);
  std::unique_ptr<Convex> c(new Convex());
  c->points.resize(n);
  V3 mn(DBL_MAX, DBL_MAX, DBL_MAX), mx(-DBL_MAX, -DBL_MAX, -DBL_MAX);
  for (uint32_t i = 0; i < n; ++i) {
    c->points[i] = V3(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
    for (int k = 0; k < 3; ++k) {
      mn[k] = std::fmin(mn[k], c->points[i][k]);
      mx[k] = std::fmax(mx[k], c->points[i][k]);
    }
  }
  c->aabb_center = (mn + mx) * 0.5;
  if (tris && ntris) {
    fillNeighborsFromTriangles(*c, tris, ntris);
    buildSupportWarmStart(*c);
  }
  s->convexes.push_back(std::move(c));
  return (int)s->convexes.size() - 1;
}

// BVHModel<OBBRSS>: beginModel/addSubModel/endModel -> buildTree (BVH_model.cpp:860-960)
int oracle_register_bvh(void* sc, const double* vertices, uint32_t nv, const uint32_t* tris, uint32_t nt) {
  Scene* s = static_cast<Scene*>(sc);
  std::unique_ptr<BVHModel> m(new BVHModel());
  m->vertices.resize(nv);
  for (uint32_t i = 0; i < nv; ++i) m->vertices[i] = V3(vertices[3 * i], vertices[3 * i + 1], vertices[3 * i + 2]);
  m->tris.resize(nt);
  for (uint32_t i = 0; i < nt; ++i)
    for (int k = 0; k < 3; ++k) m->tris[i].v[k] = tris[3 * i + k];
  m->build();
  s->bvhs.push_back(std::move(m));
  return (int)s->bvhs.size() - 1;
}
// exports the built tree in the product's node layout (what a binding would copy out of
// BVHModel<OBBRSS>::bvs); returns the node count
int oracle_bvh_export(void* sc, int id, hfb_bvh_node* out, uint32_t cap) {
  Scene* s = static_cast<Scene*>(sc);
  const BVHModel& m = *s->bvhs[(size_t)id];
  if (!out) return (int)m.num_bvs;
  for (uint32_t i = 0; i < m.num_bvs && i < cap; ++i) {
    const BVNode& n = m.bvs[i];
    hfb_bvh_node& o = out[i];
    o.first_child = n.first_child;
    o.first_primitive = n.first_primitive;
    o.num_primitives = n.num_primitives;
    o._pad = 0;
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r) {
        o.obb_axes[c * 3 + r] = n.bv.obb.axes.m[r][c];
        o.rss_axes[c * 3 + r] = n.bv.rss.axes.m[r][c];
      }
    put3(o.obb_To, n.bv.obb.To);
    put3(o.obb_extent, n.bv.obb.extent);
    put3(o.rss_Tr, n.bv.rss.Tr);
    o.rss_length[0] = n.bv.rss.length[0];
    o.rss_length[1] = n.bv.rss.length[1];
    o.rss_radius = n.bv.rss.radius;
  }
  return (int)m.num_bvs;
}

// Halfspace(n, d) / Plane(n, d) (geometric_shapes.h:885-1031): `nd` holds count x (n.x, n.y, n.z, d), normalised here
// as the constructors' unitNormalTest does (geometric_shapes.cpp:121-143); ssr may be NULL.  Returns the first handle.
int64_t oracle_register_halfspaces(void* sc, uint32_t type, const double* nd, const double* ssr, size_t count) {
  Scene* s = static_cast<Scene*>(sc);
  if (type != HFB_GEOM_PLANE && type != HFB_GEOM_HALFSPACE) return -1;
  const size_t first = s->shapes.size();
  for (size_t i = 0; i < count; ++i) {
    Shape sh;
    sh.type = (int)type;
    V3 n(nd[4 * i], nd[4 * i + 1], nd[4 * i + 2]);
    double d = nd[4 * i + 3];
    const double l = norm(n);
    if (l > 0) {
      const double inv_l = 1.0 / l;
      n *= inv_l;
      d *= inv_l;
    } else {
      n = V3(1, 0, 0);
      d = 0;
    }
    sh.p[0] = n.x;
    sh.p[1] = n.y;
    sh.p[2] = n.z;
    sh.d = d;
    sh.ssr = ssr ? ssr[i] : 0.0;
    s->shapes.push_back(sh);
  }
  return (int64_t)first;
}

// returns the handle of shapes[0]; handles are consecutive. -1 on error.
int64_t oracle_register_shapes(void* sc, const hfb_shape* shapes, size_t n) {
  Scene* s = static_cast<Scene*>(sc);
  size_t first = s->shapes.size();
  for (size_t i = 0; i < n; ++i) {
    Shape sh;
    sh.type = (int)shapes[i].type;
    if (sh.type == HFB_GEOM_PLANE || sh.type == HFB_GEOM_HALFSPACE) return -1;  // oracle_register_halfspaces (n and d)
    sh.p[0] = shapes[i].p[0];
    sh.p[1] = shapes[i].p[1];
    sh.p[2] = shapes[i].p[2];
    sh.ssr = shapes[i].ssr;
    if (sh.type == HFB_BV_OBBRSS || sh.type == HFB_BV_OBB) {  // (a model serves both kinds: the plain OBB tree is the OBB half)
      if (shapes[i].data >= s->bvhs.size()) return -1;
      sh.p[0] = (double)shapes[i].data;
    }
    if (sh.type == HFB_GEOM_CONVEX || sh.type == HFB_GEOM_TRIANGLE) {
      if (shapes[i].data >= s->convexes.size()) return -1;
      sh.cvx = s->convexes[shapes[i].data].get();
      if (sh.type == HFB_GEOM_TRIANGLE) {
        if (sh.cvx->points.size() < 3) return -1;
        for (int k = 0; k < 3; ++k) sh.tri[k] = sh.cvx->points[k];
      }
    }
    s->shapes.push_back(sh);
  }
  return (int64_t)first;
}

int oracle_batch_distance(void* sc, size_t n, const uint32_t* h1, const hfb_transform* tf1,
                          const uint32_t* h2, const hfb_transform* tf2,
                          const hfb_distance_request* req, hfb_distance_result* out,
                          const hfb_guess_out* guess_out, int nthreads) {
  Scene* s = static_cast<Scene*>(sc);
  const size_t ns = s->shapes.size();
  for (size_t i = 0; i < n; ++i)
    if (h1[i] >= ns || h2[i] >= ns) return HFB_ERR_INVALID_ARGUMENT;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
  {
    GJKSolver solver(*req);
#pragma omp for schedule(static)
    for (long long ii = 0; ii < (long long)n; ++ii) {
      const size_t i = (size_t)ii;
      const Shape& s1 = s->shapes[h1[i]];
      const Shape& s2 = s->shapes[h2[i]];
      const Tf T1 = tf_from_pod(tf1[i]), T2 = tf_from_pod(tf2[i]);
      // GJKSolver solver(request) -- distance.cpp:63 (fresh per call)
      solver.set_query(req->q);
      solver.cached_guess = V3(1, 0, 0);
      solver.support_func_cached_guess[0] = solver.support_func_cached_guess[1] = 0;
      if (req->q.gjk_initial_guess == HFB_GUESS_CACHED) {  // narrowphase.h:167-171
        if (req->q.cached_gjk_guess)
          solver.cached_guess = V3(req->q.cached_gjk_guess[3 * i], req->q.cached_gjk_guess[3 * i + 1],
                                   req->q.cached_gjk_guess[3 * i + 2]);
        if (req->q.cached_support_func_guess) {
          solver.support_func_cached_guess[0] = req->q.cached_support_func_guess[2 * i];
          solver.support_func_cached_guess[1] = req->q.cached_support_func_guess[2 * i + 1];
        }
      }
      solver.distance_upper_bound = std::numeric_limits<double>::max();
      solver.gjk.iterations = 0;
      solver.epa.iterations = 0;
      // the GJK object of a fresh solver carries its constructor's settings (gjk.cpp:51-57) until
      // runGJKAndEPA overwrites them; the TriangleP-TriangleP specialisation only calls reset() and so
      // runs with these whatever the request says (triangle_triangle.cpp:67)
      solver.gjk.distance_upper_bound = std::numeric_limits<double>::max();
      solver.gjk.gjk_variant = HFB_GJK_DEFAULT;
      solver.gjk.convergence_criterion = HFB_CRIT_DEFAULT;
      solver.gjk.convergence_criterion_type = HFB_CRIT_RELATIVE;

      hfb_distance_result& r = out[i];
      // DistanceResult::clear (collision_data.h:1140-1151)
      r.min_distance = std::numeric_limits<double>::max();
      put3(r.p1, nan3());
      put3(r.p2, nan3());
      put3(r.normal, nan3());
      r.b1 = r.b2 = -1;
      V3 p1, p2, normal;
      double distance;
      bool closed = false;
      if (s1.type == HFB_BV_OBB || s2.type == HFB_BV_OBB) {
        // BVHModel<OBB>: distance() is not restated (the reference rebuilds the tree from the transformed vertices on
        // every call and walks all of it, OBB::distance being unimplemented: traversal_node_setup.h:709-723, BV/OBB.cpp)
        r.status = (uint32_t)HFB_PATH_UNSUPPORTED << 16;
        r.iterations = 0;
        continue;
      }
      if (s1.type == HFB_BV_OBBRSS || s2.type == HFB_BV_OBBRSS) {
        // distance(): (GEOM, BVH) calls the (BVH, GEOM) entry with swapped operands and swaps
        // o1/o2, the nearest points and the normal back -- not b1/b2 (distance.cpp:74-89)
        const bool swap = s1.type != HFB_BV_OBBRSS;
        const Shape& sm = swap ? s2 : s1;
        const Shape& ss = swap ? s1 : s2;
        if (s1.type == HFB_BV_OBBRSS && s2.type == HFB_BV_OBBRSS) {  // BVHDistance<OBBRSS> (distance_func_matrix.cpp:259-268)
          BvhQueryResult q;
          bvhBvhDistance(*s->bvhs[(size_t)s1.p[0]], T1, *s->bvhs[(size_t)s2.p[0]], T2, 0.0, 0.0 /* see below */,
                         req->enable_nearest_points != 0, q);
          r.min_distance = q.distance;
          put3(r.p1, q.p1);
          put3(r.p2, q.p2);
          put3(r.normal, nan3());
          r.b1 = q.b1;
          r.b2 = q.b2;
          r.status = (uint32_t)HFB_PATH_BVH << 16;
          r.iterations = (uint32_t)(q.num_bv_tests & 0xffff) | ((uint32_t)(q.num_leaf_tests & 0xffff) << 16);
        } else if (ss.type == HFB_BV_OBBRSS || !(ss.type == HFB_GEOM_BOX || ss.type == HFB_GEOM_SPHERE ||
            ss.type == HFB_GEOM_CAPSULE || ss.type == HFB_GEOM_CONE || ss.type == HFB_GEOM_CYLINDER ||
            ss.type == HFB_GEOM_ELLIPSOID || ss.type == HFB_GEOM_CONVEX) ||
            // computeBV<OBBRSS, S> throws "Swept-sphere radius not yet supported" (geometric_shapes_utility.h:73-78)
            ss.ssr > 0) {
          r.status = (uint32_t)HFB_PATH_UNSUPPORTED << 16;  // plane, halfspace, triangle partners: not covered
          r.iterations = 0;
        } else {
          BvhQueryResult q;
          try {
          bvhShapeDistance(*s->bvhs[(size_t)sm.p[0]], swap ? T2 : T1, ss, swap ? T1 : T2, solver,
                           req->enable_signed_distance != 0,
                           // rel_err / abs_err never reach the walk: the mesh-shape node zeroes its copies in
                           // its constructor (traversal_node_bvh_shape.h:294-295) and nothing sets them; the
                           // mesh-mesh node copies them from its own default-constructed request
                           // (traversal_node_bvhs.h:409-410), before initialize() stores the caller's
                           0.0, 0.0, q);
          } catch (const std::logic_error&) {  // BoundingVolumeGuess at a mesh leaf (bvh.cpp leafGuessCheck)
            r.status = (uint32_t)HFB_PATH_UNSUPPORTED << 16;
            r.iterations = 0;
            continue;
          }
          r.min_distance = q.distance;
          put3(r.p1, swap ? q.p2 : q.p1);
          put3(r.p2, swap ? q.p1 : q.p2);
          put3(r.normal, swap ? -q.normal : q.normal);
          r.b1 = q.b1;
          r.b2 = -1;
          r.status = (uint32_t)HFB_PATH_BVH << 16;
          r.iterations = (uint32_t)(q.num_bv_tests & 0xffff) | ((uint32_t)(q.num_leaf_tests & 0xffff) << 16);
        }
        if (guess_out) {
          if (guess_out->cached_gjk_guess) put3(guess_out->cached_gjk_guess + 3 * i, solver.cached_guess);
          if (guess_out->cached_support_func_guess) {
            guess_out->cached_support_func_guess[2 * i] = solver.support_func_cached_guess[0];
            guess_out->cached_support_func_guess[2 * i + 1] = solver.support_func_cached_guess[1];
          }
        }
        continue;
      }
      bool ok = shapeShapeDistance(s1, T1, s2, T2, solver, req->enable_signed_distance != 0, distance,
                                   p1, p2, normal, closed);
      if (!ok) {
        r.status = (uint32_t)HFB_PATH_UNSUPPORTED << 16;
        r.iterations = 0;
      } else {
        // result.update(...) with strict '>' (collision_data.h:1111-1124); the
        // closed-form specialisations assign unconditionally (shape_shape_func.h:228-241)
        if (closed || r.min_distance > distance) {
          r.min_distance = distance;
          put3(r.p1, p1);
          put3(r.p2, p2);
          put3(r.normal, normal);
        }
        r.status = pack_status(solver, closed);
        r.iterations = closed ? 0 : pack_iters(solver);
      }
      if (guess_out) {  // distance.cpp:105-107
        if (guess_out->cached_gjk_guess) put3(guess_out->cached_gjk_guess + 3 * i, solver.cached_guess);
        if (guess_out->cached_support_func_guess) {
          guess_out->cached_support_func_guess[2 * i] = solver.support_func_cached_guess[0];
          guess_out->cached_support_func_guess[2 * i + 1] = solver.support_func_cached_guess[1];
        }
      }
    }
  }
  return HFB_OK;
}

// contacts[1..] and numContacts() of mesh pairs (oracle_batch_collide_contacts); unset otherwise
static hfb_contact* g_extra = nullptr;
static uint32_t* g_counts = nullptr;
static uint32_t g_max_extra = 0;
static void sink_contacts(size_t i, const std::vector<BvhContact>& cs, bool swap, bool mesh_mesh) {
  if (g_counts) g_counts[i] = (uint32_t)cs.size();
  if (!g_extra) return;
  for (size_t k = 1; k < cs.size() && k - 1 < g_max_extra; ++k) {
    const BvhContact& c = cs[k];
    hfb_contact& r = g_extra[i * g_max_extra + (k - 1)];
    std::memset(&r, 0, sizeof(r));
    r.distance_lower_bound = std::numeric_limits<double>::max();
    r.num_contacts = 1;
    r.distance = c.distance;
    const int cb2 = mesh_mesh ? c.b2 : -1;
    r.b1 = swap ? cb2 : c.b1;
    r.b2 = swap ? c.b1 : cb2;
    put3(r.pos, (c.p1 + c.p2) / 2);
    put3(r.p1, swap ? c.p2 : c.p1);
    put3(r.p2, swap ? c.p1 : c.p2);
    put3(r.normal, swap ? -c.normal : c.normal);
    r.status = (uint32_t)HFB_PATH_BVH << 16;
  }
}

int oracle_batch_collide(void* sc, size_t n, const uint32_t* h1, const hfb_transform* tf1,
                         const uint32_t* h2, const hfb_transform* tf2,
                         const hfb_collision_request* req, hfb_contact* out,
                         const hfb_guess_out* guess_out, int nthreads) {
  Scene* s = static_cast<Scene*>(sc);
  const size_t ns = s->shapes.size();
  for (size_t i = 0; i < n; ++i)
    if (h1[i] >= ns || h2[i] >= ns) return HFB_ERR_INVALID_ARGUMENT;
  // collision.cpp:82-85
  const bool minus_inf_margin = (req->security_margin == -std::numeric_limits<double>::infinity());
  if (!minus_inf_margin && req->num_max_contacts == 0) return HFB_ERR_INVALID_ARGUMENT;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
  {
    GJKSolver solver(*req);
#pragma omp for schedule(static)
    for (long long ii = 0; ii < (long long)n; ++ii) {
      const size_t i = (size_t)ii;
      hfb_contact& r = out[i];
      // CollisionResult::clear (collision_data.h:486-495)
      r.distance = std::numeric_limits<double>::max();
      r.distance_lower_bound = std::numeric_limits<double>::max();
      put3(r.p1, nan3());
      put3(r.p2, nan3());
      put3(r.normal, nan3());
      put3(r.pos, nan3());
      r.b1 = r.b2 = -1;
      r.status = 0;
      r.num_contacts = 0;
      r.iterations = 0;
      r._pad = 0;
      if (minus_inf_margin) continue;  // collision.cpp:73-76

      const Shape& s1 = s->shapes[h1[i]];
      const Shape& s2 = s->shapes[h2[i]];
      const Tf T1 = tf_from_pod(tf1[i]), T2 = tf_from_pod(tf2[i]);
      solver.set_query(req->q);
      solver.cached_guess = V3(1, 0, 0);
      solver.support_func_cached_guess[0] = solver.support_func_cached_guess[1] = 0;
      if (req->q.gjk_initial_guess == HFB_GUESS_CACHED) {
        if (req->q.cached_gjk_guess)
          solver.cached_guess = V3(req->q.cached_gjk_guess[3 * i], req->q.cached_gjk_guess[3 * i + 1],
                                   req->q.cached_gjk_guess[3 * i + 2]);
        if (req->q.cached_support_func_guess) {
          solver.support_func_cached_guess[0] = req->q.cached_support_func_guess[2 * i];
          solver.support_func_cached_guess[1] = req->q.cached_support_func_guess[2 * i + 1];
        }
      }
      solver.distance_upper_bound =
          std::max(0., std::max(req->distance_upper_bound, req->security_margin));
      solver.gjk.iterations = 0;
      solver.epa.iterations = 0;
      // the GJK object of a fresh solver carries its constructor's settings (gjk.cpp:51-57) until
      // runGJKAndEPA overwrites them; the TriangleP-TriangleP specialisation only calls reset() and so
      // runs with these whatever the request says (triangle_triangle.cpp:67)
      solver.gjk.distance_upper_bound = std::numeric_limits<double>::max();
      solver.gjk.gjk_variant = HFB_GJK_DEFAULT;
      solver.gjk.convergence_criterion = HFB_CRIT_DEFAULT;
      solver.gjk.convergence_criterion_type = HFB_CRIT_RELATIVE;

      // BVHModel<OBB> collides through the same walks: MeshShapeCollisionTraversalNodeOBB / MeshCollisionTraversalNodeOBB
      // test OBBs exactly as the OBBRSS nodes do (collision_func_matrix.cpp:141-166, 236-246)
      const bool bv1 = s1.type == HFB_BV_OBBRSS || s1.type == HFB_BV_OBB, bv2 = s2.type == HFB_BV_OBBRSS || s2.type == HFB_BV_OBB;
      if (bv1 && bv2 && s1.type != s2.type) {  // no collision_matrix[BV_OBB][BV_OBBRSS]
        r.status = (uint32_t)HFB_PATH_UNSUPPORTED << 16;
        r.iterations = 0;
        continue;
      }
      if (bv1 || bv2) {
        // collide(): (GEOM, BVH) is run as (BVH, GEOM) and swapObjects() swaps o1/o2, b1/b2, the
        // nearest points and negates the normal of each contact and of the result (collision.cpp:92-108)
        const bool swap = !bv1;
        const Shape& sm = swap ? s2 : s1;
        const Shape& ss = swap ? s1 : s2;
        if (bv1 && bv2) {  // BVHCollide<OBBRSS> (collision_func_matrix.cpp:248-257)
          const GJKSolver proto = solver;  // every leaf builds GJKSolver(request) (traversal_node_bvhs.h:197)
          BvhCollideResult q;
          GJKSolver leaf_solver = proto;
          bvhBvhCollide(*s->bvhs[(size_t)s1.p[0]], T1, *s->bvhs[(size_t)s2.p[0]], T2, leaf_solver, *req, q);
          r.distance_lower_bound = q.distance_lower_bound;
          put3(r.p1, q.lb_p1);
          put3(r.p2, q.lb_p2);
          put3(r.normal, q.lb_normal);
          if (!q.contacts.empty()) {
            const BvhContact& c = q.contacts[0];
            r.num_contacts = 1;
            r.distance = c.distance;
            r.b1 = c.b1;
            r.b2 = c.b2;
            put3(r.pos, (c.p1 + c.p2) / 2);
            put3(r.p1, c.p1);
            put3(r.p2, c.p2);
            put3(r.normal, c.normal);
          }
          r.status = (uint32_t)HFB_PATH_BVH << 16;
          r.iterations = (uint32_t)(q.num_bv_tests & 0xffff) | ((uint32_t)(q.num_leaf_tests & 0xffff) << 16);
          sink_contacts(i, q.contacts, false, true);
          continue;
        }
        const bool shape_ok = ss.type == HFB_GEOM_BOX || ss.type == HFB_GEOM_SPHERE || ss.type == HFB_GEOM_CAPSULE ||
                              ss.type == HFB_GEOM_CONE || ss.type == HFB_GEOM_CYLINDER ||
                              ss.type == HFB_GEOM_ELLIPSOID || ss.type == HFB_GEOM_CONVEX;
        // a swept-sphere radius on the shape throws in computeBV (geometric_shapes_utility.h:73-78)
        if (!shape_ok || ss.ssr > 0 || req->security_margin < 0) {  // negative margin throws for BVH (collision_func_matrix.cpp:109-112)
          r.status = (uint32_t)HFB_PATH_UNSUPPORTED << 16;
          continue;
        }
        BvhCollideResult q;
        try {
          bvhShapeCollide(*s->bvhs[(size_t)sm.p[0]], swap ? T2 : T1, ss, swap ? T1 : T2, solver, *req, q, sm.type == HFB_BV_OBB);
        } catch (const std::logic_error&) {  // BoundingVolumeGuess at a mesh leaf (bvh.cpp leafGuessCheck)
          r.status = (uint32_t)HFB_PATH_UNSUPPORTED << 16;
          continue;
        }
        r.distance_lower_bound = q.distance_lower_bound;
        put3(r.p1, swap ? q.lb_p2 : q.lb_p1);
        put3(r.p2, swap ? q.lb_p1 : q.lb_p2);
        put3(r.normal, swap ? -q.lb_normal : q.lb_normal);
        if (!q.contacts.empty()) {
          const BvhContact& c = q.contacts[0];
          r.num_contacts = 1;
          r.distance = c.distance;
          r.b1 = swap ? -1 : c.b1;
          r.b2 = swap ? c.b1 : -1;
          put3(r.pos, (c.p1 + c.p2) / 2);
          // the record carries one set of witness fields: the contact's when there is a contact
          put3(r.p1, swap ? c.p2 : c.p1);
          put3(r.p2, swap ? c.p1 : c.p2);
          put3(r.normal, swap ? -c.normal : c.normal);
        }
        r.status = (uint32_t)HFB_PATH_BVH << 16;
        r.iterations = (uint32_t)(q.num_bv_tests & 0xffff) | ((uint32_t)(q.num_leaf_tests & 0xffff) << 16);
        sink_contacts(i, q.contacts, swap, false);
        continue;
      }
      // ShapeShapeCollider::run (shape_shape_func.h:134-163)
      const bool compute_penetration = (req->enable_contact != 0) || (req->security_margin < 0);
      V3 p1, p2, normal;
      double distance;
      bool closed = false;
      bool ok = shapeShapeDistance(s1, T1, s2, T2, solver, compute_penetration, distance, p1, p2,
                                   normal, closed);
      if (!ok) {
        r.status = (uint32_t)HFB_PATH_UNSUPPORTED << 16;
        continue;
      }
      const double distToCollision = distance - req->security_margin;
      // updateDistanceLowerBoundFromLeaf (collision_data.h:1186-1197)
      if (distToCollision < r.distance_lower_bound) {
        r.distance_lower_bound = distToCollision;
        put3(r.p1, p1);
        put3(r.p2, p2);
        put3(r.normal, normal);
      }
      r.distance = distance;
      if (distToCollision <= req->q.collision_distance_threshold) {
        // Contact(o1,o2,NONE,NONE,p1,p2,normal,distance): pos = (p1+p2)/2 (collision_data.h:138-148)
        put3(r.pos, (p1 + p2) / 2);
        r.num_contacts = 1;
      }
      r.status = pack_status(solver, closed);
      r.iterations = closed ? 0 : pack_iters(solver);
      if (guess_out) {
        if (guess_out->cached_gjk_guess) put3(guess_out->cached_gjk_guess + 3 * i, solver.cached_guess);
        if (guess_out->cached_support_func_guess) {
          guess_out->cached_support_func_guess[2 * i] = solver.support_func_cached_guess[0];
          guess_out->cached_support_func_guess[2 * i + 1] = solver.support_func_cached_guess[1];
        }
      }
    }
  }
  return HFB_OK;
}

// getShapeSupportLinear over registered convexes (support_functions.cpp:401-421)

// oracle_batch_collide plus every contact of a mesh pair (see hfb_batch_collide_contacts in include/hppfcl_b200.h)
int oracle_batch_collide_contacts(void* sc, size_t n, const uint32_t* h1, const hfb_transform* tf1, const uint32_t* h2,
                                  const hfb_transform* tf2, const hfb_collision_request* req, hfb_contact* out,
                                  uint32_t max_extra, hfb_contact* extra, uint32_t* counts, int nthreads) {
  for (size_t i = 0; i < n; ++i) counts[i] = 0xffffffffu;
  g_extra = max_extra ? extra : nullptr;
  g_counts = counts;
  g_max_extra = max_extra;
  const int rc = oracle_batch_collide(sc, n, h1, tf1, h2, tf2, req, out, nullptr, nthreads);
  g_extra = nullptr;
  g_counts = nullptr;
  if (rc) return rc;
  for (size_t i = 0; i < n; ++i)
    if (counts[i] == 0xffffffffu) counts[i] = out[i].num_contacts;
  return HFB_OK;
}

int oracle_batch_convex_support(void* sc, size_t n, const uint32_t* convex_ids, const double* dirs,
                                int32_t* index_out, double* support_out) {
  Scene* s = static_cast<Scene*>(sc);
  for (size_t i = 0; i < n; ++i) {
    if (convex_ids[i] >= s->convexes.size()) return HFB_ERR_INVALID_ARGUMENT;
    const Convex* c = s->convexes[convex_ids[i]].get();
    V3 dir(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
    int hint = 0;
    double maxdot = dot(c->points[0], dir);
    for (int k = 1; k < (int)c->points.size(); ++k) {
      double d = dot(c->points[(size_t)k], dir);
      if (d > maxdot) { maxdot = d; hint = k; }
    }
    index_out[i] = hint;
    put3(support_out + 3 * i, c->points[(size_t)hint]);
  }
  return HFB_OK;
}

// hill-climbing support (support_functions.cpp:324-397) with a fresh
// ShapeSupportData per query; used to check it agrees with the linear scan.
int oracle_batch_convex_support_log(void* sc, size_t n, const uint32_t* convex_ids,
                                    const double* dirs, int32_t* index_out) {
  Scene* s = static_cast<Scene*>(sc);
  for (size_t i = 0; i < n; ++i) {
    const Convex* c = s->convexes[convex_ids[i]].get();
    Shape sh;
    sh.type = HFB_GEOM_CONVEX;
    sh.cvx = c;
    SupportData sd;
    V3 sup;
    int hint = 0;
    getShapeSupport(&sh, V3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), sup, hint, sd);
    index_out[i] = hint;
  }
  return HFB_OK;
}

// Project::*Origin exposed for the known-answer tests of test/simple.cpp
void oracle_project_line_origin(const double* a, const double* b, double* param, double* sqr,
                                unsigned* encode) {
  ProjectResult r = projectLineOrigin(V3(a[0], a[1], a[2]), V3(b[0], b[1], b[2]));
  for (int i = 0; i < 4; ++i) param[i] = r.parameterization[i];
  *sqr = r.sqr_distance;
  *encode = r.encode;
}
void oracle_project_triangle_origin(const double* a, const double* b, const double* c, double* param,
                                    double* sqr, unsigned* encode) {
  ProjectResult r = projectTriangleOrigin(V3(a[0], a[1], a[2]), V3(b[0], b[1], b[2]), V3(c[0], c[1], c[2]));
  for (int i = 0; i < 4; ++i) param[i] = r.parameterization[i];
  *sqr = r.sqr_distance;
  *encode = r.encode;
}
void oracle_project_tetrahedra_origin(const double* a, const double* b, const double* c,
                                      const double* d, double* param, double* sqr, unsigned* encode) {
  ProjectResult r = projectTetrahedraOrigin(V3(a[0], a[1], a[2]), V3(b[0], b[1], b[2]),
                                            V3(c[0], c[1], c[2]), V3(d[0], d[1], d[2]));
  for (int i = 0; i < 4; ++i) param[i] = r.parameterization[i];
  *sqr = r.sqr_distance;
  *encode = r.encode;
}

// Low-level hook for the reference's GJK/EPA unit tests (test/gjk.cpp:337-490):
// MinkowskiDiff::set(s0,s1,tf0,tf1); GJK gjk(max_it,tol); gjk.evaluate(shape,guess);
// then gjk.getWitnessPointsAndNormal or EPA epa(epa_it,epa_tol).evaluate(gjk, epa_guess).
// out[0..2]=w0, out[3..5]=w1, out[6..8]=normal, out[9..11]=gjk.ray, out[12]=gjk.distance,
// out[13]=epa.depth; istat[0]=gjk status, istat[1]=epa status, istat[2]=gjk iterations,
// istat[3]=epa iterations, istat[4]=simplex rank
int oracle_gjk_lowlevel(void* sc, uint32_t h0, const hfb_transform* tf0, uint32_t h1,
                        const hfb_transform* tf1, uint32_t gjk_max_it, double gjk_tol, int variant,
                        int criterion, int criterion_type, const double* guess, int run_epa_if_collision,
                        uint32_t epa_max_it, double epa_tol, const double* epa_guess, double* out,
                        int* istat) {
  Scene* s = static_cast<Scene*>(sc);
  const Shape& s0 = s->shapes[h0];
  const Shape& s1 = s->shapes[h1];
  MinkowskiDiff md;
  md.set(&s0, &s1, tf_from_pod(*tf0), tf_from_pod(*tf1));
  GJK gjk(gjk_max_it, gjk_tol);
  gjk.gjk_variant = variant;
  gjk.convergence_criterion = criterion;
  gjk.convergence_criterion_type = criterion_type;
  int hint[2] = {0, 0};
  GJK::Status st = gjk.evaluate(md, V3(guess[0], guess[1], guess[2]), hint);
  V3 w0, w1, normal;
  istat[0] = (int)st;
  istat[1] = (int)EPA::DidNotRun;
  istat[2] = (int)gjk.iterations;
  istat[3] = 0;
  istat[4] = gjk.simplex->rank;
  out[13] = 0;
  if (run_epa_if_collision && st == GJK::Collision) {
    EPA epa(epa_max_it, epa_tol);
    EPA::Status es = epa.evaluate(gjk, V3(epa_guess[0], epa_guess[1], epa_guess[2]));
    istat[1] = (int)es;
    istat[3] = (int)epa.iterations;
    epa.getWitnessPointsAndNormal(md, w0, w1, normal);
    out[13] = epa.depth;
  } else {
    gjk.getWitnessPointsAndNormal(md, w0, w1, normal);
  }
  put3(out, w0);
  put3(out + 3, w1);
  put3(out + 6, normal);
  put3(out + 9, gjk.ray);
  out[12] = gjk.distance;
  return HFB_OK;
}

// number of Project::*Origin calls so far that the reference answers from uninitialised memory (gjk_epa.cpp)
unsigned long long oracle_undefined_projections() { return g_undefined_projections.load(); }

int oracle_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
