// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.hpp).
// Restatement of: support functions (src/narrowphase/support_functions.cpp:111-437),
// MinkowskiDiff (src/narrowphase/minkowski_difference.cpp:47-305), GJKSolver
// (include/hpp/fcl/narrowphase/narrowphase.h:58-724), the closed-form pair
// distances (src/narrowphase/details.h, src/distance/*.cpp) and the
// ShapeShapeDistance dispatch (include/hpp/fcl/internal/shape_shape_func.h).
#include <algorithm>
#include <cassert>
#include <set>

#include "oracle.hpp"

namespace oracle {

static const double kDummyPrecision = 1e-12;
// Box::inflate is a function-local static decided by the first support direction
// a process ever queries (support_functions.cpp:146).  With the default GJK
// guess (1,0,0) that first direction is (-1,0,0), which has zero components, so
// the value every realistic process ends up with is 1 + 1e-10.  Fixed here.
static const double kBoxInflate = 1 + 1e-10;

// ---------------------------------------------------------------- supports --
static void supportTriangle(const Shape* s, const V3& dir, V3& support) {  // :111-134
  const V3& a = s->tri[0];
  const V3& b = s->tri[1];
  const V3& c = s->tri[2];
  double dota = dot(dir, a), dotb = dot(dir, b), dotc = dot(dir, c);
  if (dota > dotb) {
    if (dotc > dota) support = c; else support = a;
  } else {
    if (dotc > dotb) support = c; else support = b;
  }
}

static void supportBox(const Shape* s, const V3& dir, V3& support) {  // :141-157
  for (int i = 0; i < 3; ++i) {
    double s1 = (dir[i] > kDummyPrecision) ? s->p[i] : 0.0;
    double s2 = (dir[i] < -kDummyPrecision) ? (-kBoxInflate * s->p[i]) : 0.0;
    support[i] = s1 + s2;
  }
}

static void supportEllipsoid(const Shape* s, const V3& dir, V3& support) {  // :183-199
  double a2 = s->p[0] * s->p[0];
  double b2 = s->p[1] * s->p[1];
  double c2 = s->p[2] * s->p[2];
  V3 v(a2 * dir.x, b2 * dir.y, c2 * dir.z);
  double d = std::sqrt(dot(v, dir));
  support = v / d;
}

static void supportCapsule(const Shape* s, const V3& dir, V3& support) {  // :206-222
  support = V3(0, 0, 0);
  if (dir.z > kDummyPrecision) support.z = s->p[1];
  else if (dir.z < -kDummyPrecision) support.z = -s->p[1];
}

static void supportCone(const Shape* s, const V3& dir, V3& support) {  // :229-274
  static const double inflate = 1 + 1e-10;
  double h = s->p[1];
  double r = s->p[0];
  if (std::fabs(dir.x) <= kDummyPrecision && std::fabs(dir.y) <= kDummyPrecision) {
    support.x = 0;
    support.y = 0;
    if (dir.z > kDummyPrecision) support.z = h;
    else support.z = -inflate * h;
  } else {
    double zdist = dir.x * dir.x + dir.y * dir.y;
    double len = zdist + dir.z * dir.z;
    zdist = std::sqrt(zdist);
    if (dir.z <= 0) {
      double rad = r / zdist;
      support.x = rad * dir.x;
      support.y = rad * dir.y;
      support.z = -h;
    } else {
      len = std::sqrt(len);
      double sin_a = r / std::sqrt(r * r + 4 * h * h);
      if (dir.z > len * sin_a)
        support = V3(0, 0, h);
      else {
        double rad = r / zdist;
        support.x = rad * dir.x;
        support.y = rad * dir.y;
        support.z = -h;
      }
    }
  }
}

static void supportCylinder(const Shape* s, const V3& dir, V3& support) {  // :281-317
  static const double inflate = 1 + 1e-10;
  double half_h = s->p[1];
  double r = s->p[0];
  const bool aligned = std::fabs(dir.x) <= kDummyPrecision && std::fabs(dir.y) <= kDummyPrecision;
  if (aligned) half_h *= inflate;
  if (dir.z > kDummyPrecision) support.z = half_h;
  else if (dir.z < -kDummyPrecision) support.z = -half_h;
  else { support.z = 0; r *= inflate; }
  if (aligned) {
    support.x = 0;
    support.y = 0;
  } else {
    // dir.head<2>().normalized() * r
    double z = dir.x * dir.x + dir.y * dir.y;
    double nx = dir.x, ny = dir.y;
    if (z > 0) { double q = std::sqrt(z); nx = dir.x / q; ny = dir.y / q; }
    support.x = nx * r;
    support.y = ny * r;
  }
}

static void supportConvexLinear(const Convex* c, const V3& dir, V3& support, int& hint) {  // :401-421
  const std::vector<V3>& pts = c->points;
  hint = 0;
  double maxdot = dot(pts[0], dir);
  for (int i = 1; i < (int)pts.size(); ++i) {
    double d = dot(pts[(size_t)i], dir);
    if (d > maxdot) { maxdot = d; hint = i; }
  }
  support = pts[(size_t)hint];
}

static void supportConvexLog(const Convex* c, const V3& dir, V3& support, int& hint,
                             SupportData& sd) {  // :324-397
  const double use_warm_start_threshold = 0.9;
  V3 dir_normalized = normalized(dir);
  if (!is_zero(sd.last_dir) && !c->warm_points.empty() &&
      dot(sd.last_dir, dir_normalized) < use_warm_start_threshold) {
    double maxdot = dot(c->warm_points[0], dir);
    hint = c->warm_indices[0];
    for (size_t i = 1; i < c->warm_points.size(); ++i) {
      double d = dot(c->warm_points[i], dir);
      if (d > maxdot) { maxdot = d; hint = c->warm_indices[i]; }
    }
  }
  sd.last_dir = dir_normalized;

  const std::vector<V3>& pts = c->points;
  const size_t np = pts.size();
  if (hint < 0 || hint >= (int)np) hint = 0;
  double maxdot = dot(pts[(size_t)hint], dir);
  std::vector<int8_t>& visited = sd.visited;
  if (visited.size() == np) std::fill(visited.begin(), visited.end(), false);
  else visited.assign(np, false);
  visited[(size_t)hint] = true;
  bool found = true;
  bool loose_check = true;
  while (found) {
    const std::vector<unsigned>& n = c->neighbors[(size_t)hint];
    found = false;
    for (size_t in = 0; in < n.size(); ++in) {
      const unsigned ip = n[in];
      if (visited[ip]) continue;
      visited[ip] = true;
      const double d = dot(pts[ip], dir);
      bool better = false;
      if (d > maxdot) { better = true; loose_check = false; }
      else if (loose_check && d == maxdot) better = true;
      if (better) { maxdot = d; hint = (int)ip; found = true; }
    }
  }
  support = pts[(size_t)hint];
}

void getShapeSupport(const Shape* s, const V3& dir, V3& support, int& hint, SupportData& data) {
  switch (s->type) {
    case HFB_GEOM_TRIANGLE: supportTriangle(s, dir, support); break;
    case HFB_GEOM_BOX: supportBox(s, dir, support); break;
    case HFB_GEOM_SPHERE: support = V3(0, 0, 0); break;  // :164-176
    case HFB_GEOM_ELLIPSOID: supportEllipsoid(s, dir, support); break;
    case HFB_GEOM_CAPSULE: supportCapsule(s, dir, support); break;
    case HFB_GEOM_CONE: supportCone(s, dir, support); break;
    case HFB_GEOM_CYLINDER: supportCylinder(s, dir, support); break;
    case HFB_GEOM_CONVEX:
      // switch at num_vertices_large_convex_threshold = 32 (:425-437, minkowski_difference.cpp:136-152)
      if (s->cvx->points.size() > 32 && s->cvx->has_neighbors())
        supportConvexLog(s->cvx, dir, support, hint, data);
      else
        supportConvexLinear(s->cvx, dir, support, hint);
      break;
    default: support = V3(0, 0, 0);
  }
}

void fillNeighborsFromTriangles(Convex& c, const uint32_t* tris, size_t ntris) {  // convex.hxx:231-280
  std::vector<std::set<unsigned>> nn(c.points.size());
  for (size_t l = 0; l < ntris; ++l) {
    const uint32_t* poly = tris + 3 * l;
    const int n = 3;
    for (int j = 0; j < n; ++j) {
      int i = (j == 0) ? n - 1 : j - 1;
      int k = (j == n - 1) ? 0 : j + 1;
      unsigned pi = poly[i], pj = poly[j], pk = poly[k];
      nn[pj].insert(pi);
      nn[pj].insert(pk);
    }
  }
  c.neighbors.resize(c.points.size());
  for (size_t i = 0; i < c.points.size(); ++i) c.neighbors[i].assign(nn[i].begin(), nn[i].end());
}

void buildSupportWarmStart(Convex& c) {  // gjk.cpp:1470-1534
  c.warm_points.clear();
  c.warm_indices.clear();
  if (c.points.size() < 32) return;
  Shape sh;
  sh.type = HFB_GEOM_CONVEX;
  sh.cvx = &c;
  int support_hint = 0;
  // ONE ShapeSupportData is declared before the loops (:1480) and passed to every
  // call, so last_dir / visited persist across the 14 queries.
  SupportData support_data;
  V3 axiis(0, 0, 0);
  for (int i = 0; i < 3; ++i) {
    axiis[i] = 1;
    {
      V3 support;
      getShapeSupport(&sh, axiis, support, support_hint, support_data);
      c.warm_points.push_back(support);
      c.warm_indices.push_back(support_hint);
    }
    axiis[i] = -1;
    {
      V3 support;
      getShapeSupport(&sh, axiis, support, support_hint, support_data);
      c.warm_points.push_back(support);
      c.warm_indices.push_back(support_hint);
    }
    axiis[i] = 0;
  }
  const V3 eis[4] = {V3(1, 1, 1), V3(-1, 1, 1), V3(-1, -1, 1), V3(1, -1, 1)};
  for (size_t k = 0; k < 4; ++k) {
    {
      V3 support;
      getShapeSupport(&sh, eis[k], support, support_hint, support_data);
      c.warm_points.push_back(support);
      c.warm_indices.push_back(support_hint);
    }
    {
      V3 support;
      getShapeSupport(&sh, -eis[k], support, support_hint, support_data);
      c.warm_points.push_back(support);
      c.warm_indices.push_back(support_hint);
    }
  }
}

// ----------------------------------------------------------- MinkowskiDiff --
static void setup_md(MinkowskiDiff& md, const Shape* s0, const Shape* s1, bool identity) {
  // makeGetSupportFunction0/1 (minkowski_difference.cpp:78-226), NoSweptSphere
  md.shapes[0] = s0;
  md.shapes[1] = s1;
  // getNormalizeSupportDirectionFromShapes (:261-266): only ConvexBase needs it
  md.normalize_support_direction = (s0->type == HFB_GEOM_CONVEX) && (s1->type == HFB_GEOM_CONVEX);
  md.identity = identity;
  for (int k = 0; k < 2; ++k) {
    const Shape* s = md.shapes[k];
    md.swept_sphere_radius[k] = s->ssr;
    if (s->type == HFB_GEOM_SPHERE || s->type == HFB_GEOM_CAPSULE) md.swept_sphere_radius[k] += s->p[0];
    if (s->type == HFB_GEOM_CONVEX && s->cvx->points.size() > 32) {
      md.data[k].visited.assign(s->cvx->points.size(), false);
      md.data[k].last_dir = V3(0, 0, 0);
    }
  }
}

void MinkowskiDiff::set(const Shape* s0, const Shape* s1, const Tf& tf0, const Tf& tf1) {  // :269-285
  oR1 = tmul(tf0.R, tf1.R);
  ot1 = tmul(tf0.R, tf1.T - tf0.T);
  bool id = is_identity(oR1) && is_zero(ot1);
  setup_md(*this, s0, s1, id);
}

void MinkowskiDiff::set(const Shape* s0, const Shape* s1) {  // :293-305
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) oR1.m[i][j] = (i == j) ? 1.0 : 0.0;
  ot1 = V3(0, 0, 0);
  setup_md(*this, s0, s1, true);
}

void MinkowskiDiff::support(const V3& dir, V3& supp0, V3& supp1, int hint[2]) {  // getSupportTpl :47-63
  getShapeSupport(shapes[0], dir, supp0, hint[0], data[0]);
  if (identity) {
    getShapeSupport(shapes[1], -dir, supp1, hint[1], data[1]);
  } else {
    getShapeSupport(shapes[1], -tmul(oR1, dir), supp1, hint[1], data[1]);
    supp1 = mul(oR1, supp1) + ot1;
  }
}

// ---------------------------------------------------------------- GJKSolver --
void GJKSolver::set_query(const hfb_query_request& q) {  // narrowphase.h:162-190 / :214-244
  gjk_initial_guess = q.gjk_initial_guess;
  gjk_max_iterations = q.gjk_max_iterations;
  gjk_tolerance = q.gjk_tolerance;
  gjk_variant = q.gjk_variant;
  gjk_convergence_criterion = q.gjk_convergence_criterion;
  gjk_convergence_criterion_type = q.gjk_convergence_criterion_type;
  epa_max_iterations = q.epa_max_iterations;
  epa_tolerance = q.epa_tolerance;
  epa.status = EPA::DidNotRun;
  gjk.status = GJK::DidNotRun;
}

GJKSolver::GJKSolver(const hfb_distance_request& r)
    : gjk(r.q.gjk_max_iterations, r.q.gjk_tolerance), epa(0, r.q.epa_tolerance) {
  cached_guess = V3(1, 0, 0);
  support_func_cached_guess[0] = support_func_cached_guess[1] = 0;
  set_query(r.q);
  distance_upper_bound = std::numeric_limits<double>::max();  // :175
}

GJKSolver::GJKSolver(const hfb_collision_request& r)
    : gjk(r.q.gjk_max_iterations, r.q.gjk_tolerance), epa(0, r.q.epa_tolerance) {
  cached_guess = V3(1, 0, 0);
  support_func_cached_guess[0] = support_func_cached_guess[1] = 0;
  set_query(r.q);
  distance_upper_bound = std::max(0., std::max(r.distance_upper_bound, r.security_margin));  // :228-229
}

double GJKSolver::shapeDistance(const Shape& s1, const Tf& tf1, const Shape& s2, const Tf& tf2,
                                bool compute_penetration, V3& p1, V3& p2, V3& normal) {
  double distance;
  if (s2.type == HFB_GEOM_TRIANGLE) {  // :322-336
    const Tf tf_1M2 = tf1.inverse_times(tf2);
    Shape tri = s2;
    tri.tri[0] = tf_1M2.transform(s2.tri[0]);
    tri.tri[1] = tf_1M2.transform(s2.tri[1]);
    tri.tri[2] = tf_1M2.transform(s2.tri[2]);
    runGJKAndEPA(s1, tf1, tri, tf_1M2, compute_penetration, distance, p1, p2, normal, true);
    return distance;
  }
  if (s1.type == HFB_GEOM_TRIANGLE) {  // :339-348
    distance = shapeDistance(s2, tf2, s1, tf1, compute_penetration, p2, p1, normal);
    normal = -normal;
    return distance;
  }
  runGJKAndEPA(s1, tf1, s2, tf2, compute_penetration, distance, p1, p2, normal, false);  // :308-317
  return distance;
}

void GJKSolver::runGJKAndEPA(const Shape& s1, const Tf& tf1, const Shape& s2, const Tf& tf2,
                             bool compute_penetration, double& distance, V3& p1, V3& p2,
                             V3& normal, bool relative_already) {  // :420-587
  if (relative_already)
    minkowski_difference.set(&s1, &s2);
  else
    minkowski_difference.set(&s1, &s2, tf1, tf2);
  gjk.reset(gjk_max_iterations, gjk_tolerance);
  gjk.distance_upper_bound = distance_upper_bound;
  gjk.gjk_variant = gjk_variant;
  gjk.convergence_criterion = gjk_convergence_criterion;
  gjk.convergence_criterion_type = gjk_convergence_criterion_type;
  epa.status = EPA::DidNotRun;

  // getGJKInitialGuess (:353-391)
  V3 guess;
  int support_hint[2] = {support_func_cached_guess[0], support_func_cached_guess[1]};
  switch (gjk_initial_guess) {
    case HFB_GUESS_DEFAULT: guess = V3(1, 0, 0); break;
    case HFB_GUESS_CACHED: guess = cached_guess; break;
    case HFB_GUESS_BOUNDING_VOLUME: {
      V3 c1 = (s1.type == HFB_GEOM_CONVEX) ? s1.cvx->aabb_center : V3(0, 0, 0);
      V3 c2 = (s2.type == HFB_GEOM_CONVEX) ? s2.cvx->aabb_center : V3(0, 0, 0);
      guess = c1 - (mul(minkowski_difference.oR1, c2) + minkowski_difference.ot1);
    } break;
    default: guess = V3(1, 0, 0);
  }

  gjk.evaluate(minkowski_difference, guess, support_hint);

  auto extract_gjk = [&]() {  // GJKExtractWitnessPointsAndNormal :610-636
    cached_guess = gjk.ray;
    support_func_cached_guess[0] = gjk.support_hint[0];
    support_func_cached_guess[1] = gjk.support_hint[1];
    distance = gjk.distance;
    gjk.getWitnessPointsAndNormal(minkowski_difference, p1, p2, normal);
    V3 p = tf1.transform(0.5 * (p1 + p2));
    normal = mul(tf1.R, normal);
    p1 = p - 0.5 * distance * normal;
    p2 = p + 0.5 * distance * normal;
  };
  auto extract_epa = [&]() {  // EPAExtractWitnessPointsAndNormal :658-711
    cached_guess = -(epa.depth * epa.normal);
    support_func_cached_guess[0] = epa.support_hint[0];
    support_func_cached_guess[1] = epa.support_hint[1];
    distance = std::min(0., -epa.depth);
    epa.getWitnessPointsAndNormal(minkowski_difference, p1, p2, normal);
    V3 p = tf1.transform(0.5 * (p1 + p2));
    normal = mul(tf1.R, normal);
    p1 = p - 0.5 * distance * normal;
    p2 = p + 0.5 * distance * normal;
  };
  auto extract_epa_failed = [&]() {  // :713-723
    cached_guess = V3(1, 0, 0);
    support_func_cached_guess[0] = support_func_cached_guess[1] = 0;
    distance = -std::numeric_limits<double>::max();
    p1 = p2 = normal = nan3();
  };

  switch (gjk.status) {
    case GJK::DidNotRun:
      cached_guess = V3(1, 0, 0);
      support_func_cached_guess[0] = support_func_cached_guess[1] = 0;
      distance = -std::numeric_limits<double>::max();
      p1 = p2 = normal = nan3();
      break;
    case GJK::Failed:
      extract_gjk();
      break;
    case GJK::NoCollisionEarlyStopped:  // :589-608
      cached_guess = gjk.ray;
      support_func_cached_guess[0] = gjk.support_hint[0];
      support_func_cached_guess[1] = gjk.support_hint[1];
      distance = gjk.distance;
      p1 = p2 = normal = nan3();
      break;
    case GJK::NoCollision:
      extract_gjk();
      break;
    case GJK::CollisionWithPenetrationInformation:
      extract_gjk();
      break;
    case GJK::Collision:
      if (!compute_penetration) {  // :638-656
        support_func_cached_guess[0] = gjk.support_hint[0];
        support_func_cached_guess[1] = gjk.support_hint[1];
        distance = gjk.distance;
        p1 = p2 = normal = nan3();
      } else {
        epa.reset(epa_max_iterations, epa_tolerance);
        epa.evaluate(gjk, -guess);
        switch (epa.status) {
          case EPA::OutOfFaces:
          case EPA::OutOfVertices:
          case EPA::Failed:
          case EPA::Valid:
          case EPA::AccuracyReached:
          case EPA::Degenerated:
          case EPA::NonConvex:
          case EPA::InvalidHull:
            extract_epa();
            break;
          case EPA::DidNotRun:
          case EPA::FallBack:
            extract_epa_failed();
            break;
        }
      }
      break;
  }
}

// ------------------------------------------------------ closed-form pairs --
// details.h:52-70
static inline void lineSegmentPointClosestToPoint(const V3& p, const V3& s1, const V3& s2, V3& sp) {
  V3 v = s2 - s1;
  V3 w = p - s1;
  double c1 = dot(w, v);
  double c2 = dot(v, v);
  if (c1 <= 0) sp = s1;
  else if (c2 <= c1) sp = s2;
  else { double b = c1 / c2; V3 Pb = s1 + v * b; sp = Pb; }
}

// details.h:76-101
static double sphereCapsuleDistance(const Shape& s1, const Tf& tf1, const Shape& s2, const Tf& tf2,
                                    V3& p1, V3& p2, V3& normal) {
  V3 pos1(tf2.transform(V3(0., 0., s2.p[1])));
  V3 pos2(tf2.transform(V3(0., 0., -s2.p[1])));
  V3 s_c = tf1.T;
  V3 segment_point;
  lineSegmentPointClosestToPoint(s_c, pos1, pos2, segment_point);
  normal = segment_point - s_c;
  double nrm = norm(normal);
  double r1 = s1.p[0] + s1.ssr;
  double r2 = s2.p[0] + s2.ssr;
  double dist = nrm - r1 - r2;
  static const double eps(std::numeric_limits<double>::epsilon());
  if (nrm > eps) normal = normalized(normal);
  else normal = V3(1, 0, 0);
  p1 = s_c + normal * r1;
  p2 = segment_point - normal * r2;
  return dist;
}

// details.h:107-209
static double sphereCylinderDistance(const Shape& s1, const Tf& tf1, const Shape& s2, const Tf& tf2,
                                     V3& p1, V3& p2, V3& normal) {
  static const double eps(std::sqrt(std::numeric_limits<double>::epsilon()));
  double r1(s1.p[0]);
  double r2(s2.p[0]);
  double lz2(s2.p[1]);
  V3 A(tf2.transform(V3(0, 0, -lz2)));
  V3 B(tf2.transform(V3(0, 0, lz2)));
  V3 S(tf1.T);
  V3 u(tf2.R.col(2));
  V3 AS(S - A);
  double s(dot(u, AS));
  V3 P(A + s * u);
  V3 PS(S - P);
  double dPS = norm(PS);
  V3 v(0, 0, 0);
  double dist;
  if (dPS > eps) v = (1 / dPS) * PS;
  if (s <= 0) {
    if (dPS <= r2) {
      dist = -s - r1;
      p1 = S + r1 * u;
      p2 = A + dPS * v;
      normal = u;
    } else {
      p2 = A + r2 * v;
      V3 Sp2(p2 - S);
      double dSp2 = norm(Sp2);
      if (dSp2 > eps) {
        normal = (1 / dSp2) * Sp2;
        p1 = S + r1 * normal;
        dist = dSp2 - r1;
      } else {
        normal = p2 - .5 * (A + B);
        normal = normalized(normal);
        dist = -r1;
        p1 = S + r1 * normal;
      }
    }
  } else if (s <= (s2.p[1] * 2)) {
    normal = -v;
    dist = dPS - r1 - r2;
    p2 = P + r2 * v;
    p1 = S - r1 * v;
  } else {
    if (dPS <= r2) {
      dist = s - (s2.p[1] * 2) - r1;
      p1 = S - r1 * u;
      p2 = B + dPS * v;
      normal = -u;
    } else {
      p2 = B + r2 * v;
      V3 Sp2(p2 - S);
      double dSp2 = norm(Sp2);
      if (dSp2 > eps) {
        normal = (1 / dSp2) * Sp2;
        p1 = S + r1 * normal;
        dist = dSp2 - r1;
      } else {
        normal = p2 - .5 * (A + B);
        normal = normalized(normal);
        p1 = S + r1 * normal;
        dist = -r1;
      }
    }
  }
  const double ssr1 = s1.ssr;
  const double ssr2 = s2.ssr;
  if (ssr1 > 0 || ssr2 > 0) {
    p1 += ssr1 * normal;
    p2 -= ssr2 * normal;
    dist -= (ssr1 + ssr2);
  }
  return dist;
}

// details.h:215-231
static double sphereSphereDistance(const Shape& s1, const Tf& tf1, const Shape& s2, const Tf& tf2,
                                   V3& p1, V3& p2, V3& normal) {
  const V3& center1 = tf1.T;
  const V3& center2 = tf2.T;
  double r1 = (s1.p[0] + s1.ssr);
  double r2 = (s2.p[0] + s2.ssr);
  V3 c1c2 = center2 - center1;
  double cdist = norm(c1c2);
  V3 unit(1, 0, 0);
  if (cdist > DBL_EPSILON) unit = c1c2 / cdist;
  double dist = cdist - r1 - r2;
  normal = unit;
  p1 = center1 + r1 * unit;
  p2 = center2 - r2 * unit;
  return dist;
}

// details.h:235-255
static double segmentSqrDistance(const V3& from, const V3& to, const V3& p, V3& nearest) {
  V3 diff = p - from;
  V3 v = to - from;
  double t = dot(v, diff);
  if (t > 0) {
    double dotVV = sqnorm(v);
    if (t < dotVV) { t /= dotVV; diff -= v * t; }
    else { t = 1; diff -= v; }
  } else
    t = 0;
  nearest = from + v * t;
  return sqnorm(diff);
}

// details.h:258-280
static bool projectInTriangle(const V3& p1, const V3& p2, const V3& p3, const V3& normal, const V3& p) {
  V3 edge1(p2 - p1), edge2(p3 - p2), edge3(p1 - p3);
  V3 p1_to_p(p - p1), p2_to_p(p - p2), p3_to_p(p - p3);
  V3 e1n(cross(edge1, normal)), e2n(cross(edge2, normal)), e3n(cross(edge3, normal));
  double r1 = dot(e1n, p1_to_p), r2 = dot(e2n, p2_to_p), r3 = dot(e3n, p3_to_p);
  if ((r1 > 0 && r2 > 0 && r3 > 0) || (r1 <= 0 && r2 <= 0 && r3 <= 0)) return true;
  return false;
}

// details.h:286-342
static double sphereTriangleDistance(const Shape& s, const Tf& tf1, const Shape& tri, const Tf& tf2,
                                     V3& p1, V3& p2, V3& normal) {
  const V3 P1 = tf2.transform(tri.tri[0]);
  const V3 P2 = tf2.transform(tri.tri[1]);
  const V3 P3 = tf2.transform(tri.tri[2]);
  V3 tri_normal = cross(P2 - P1, P3 - P1);
  tri_normal = normalized(tri_normal);
  const V3& center = tf1.T;
  const double radius = s.p[0] + s.ssr + tri.ssr;
  V3 p1_to_center = center - P1;
  double distance_from_plane = dot(p1_to_center, tri_normal);
  V3 closest_point(nan3());
  double min_distance_sqr, distance_sqr;
  if (distance_from_plane < 0) {
    distance_from_plane *= -1;
    tri_normal *= -1;
  }
  if (projectInTriangle(P1, P2, P3, tri_normal, center)) {
    closest_point = center - tri_normal * distance_from_plane;
    min_distance_sqr = distance_from_plane * distance_from_plane;
  } else {
    V3 nearest_on_edge;
    min_distance_sqr = segmentSqrDistance(P1, P2, center, closest_point);
    distance_sqr = segmentSqrDistance(P2, P3, center, nearest_on_edge);
    if (distance_sqr < min_distance_sqr) { min_distance_sqr = distance_sqr; closest_point = nearest_on_edge; }
    distance_sqr = segmentSqrDistance(P3, P1, center, nearest_on_edge);
    if (distance_sqr < min_distance_sqr) { min_distance_sqr = distance_sqr; closest_point = nearest_on_edge; }
  }
  normal = normalized(closest_point - center);
  p1 = center + normal * (s.p[0] + s.ssr);
  p2 = closest_point - normal * tri.ssr;
  const double distance = std::sqrt(min_distance_sqr) - radius;
  return distance;
}

// details.h:435-496
static double boxSphereDistance(const Shape& b, const Tf& tfb, const Shape& s, const Tf& tfs,
                                V3& pb, V3& ps, V3& normal) {
  const V3& os = tfs.T;
  const V3& ob = tfb.T;
  const M3& Rb = tfb.R;
  pb = ob;
  bool outside = false;
  const V3 os_in_b_frame(tmul(Rb, os - ob));
  int axis = -1;
  double min_d = std::numeric_limits<double>::max();
  for (int i = 0; i < 3; ++i) {
    double facedist;
    if (os_in_b_frame[i] < -b.p[i]) {
      pb -= b.p[i] * Rb.col(i);
      outside = true;
    } else if (os_in_b_frame[i] > b.p[i]) {
      pb += b.p[i] * Rb.col(i);
      outside = true;
    } else {
      pb += os_in_b_frame[i] * Rb.col(i);
      if (!outside && (facedist = b.p[i] - std::fabs(os_in_b_frame[i])) < min_d) {
        axis = i;
        min_d = facedist;
      }
    }
  }
  normal = pb - os;
  double pdist = norm(normal);
  double dist;
  if (outside) {
    dist = pdist - s.p[0];
    normal /= -pdist;
  } else {
    if (os_in_b_frame[axis] >= 0) normal = Rb.col(axis);
    else normal = -Rb.col(axis);
    dist = -min_d - s.p[0];
  }
  ps = os - s.p[0] * normal;
  if (!outside || dist <= 0) pb = ps - dist * normal;
  const double ssrb = b.ssr;
  const double ssrs = s.ssr;
  if (ssrb > 0 || ssrs > 0) {
    pb += ssrb * normal;
    ps -= ssrs * normal;
    dist -= (ssrb + ssrs);
  }
  return dist;
}

// src/distance/capsule_capsule.cpp:51-166
static double clamp01(const double& num, const double& denom) {
  if (num <= 0.) return 0.;
  else if (num >= denom) return 1.;
  else return num / denom;
}
static void clamped_linear(V3& a_sd, const V3& a, const double& s_n, const double& s_d, const V3& d) {
  if (s_n <= 0.) a_sd = a;
  else if (s_n >= s_d) a_sd = a + d;
  else a_sd = a + s_n / s_d * d;
}
static double capsuleCapsuleDistance(const Shape& c1s, const Tf& tf1, const Shape& c2s, const Tf& tf2,
                                     V3& wp1, V3& wp2, V3& normal) {
  double EPSILON = std::numeric_limits<double>::epsilon() * 100;
  const V3& c1 = tf1.T;
  const V3& c2 = tf2.T;
  double halfLength1 = c1s.p[1];
  double halfLength2 = c2s.p[1];
  double radius1 = (c1s.p[0] + c1s.ssr);
  double radius2 = (c2s.p[0] + c2s.ssr);
  const V3 d1 = 2 * halfLength1 * tf1.R.col(2);
  const V3 d2 = 2 * halfLength2 * tf2.R.col(2);
  const V3 p1 = c1 - d1 / 2;
  const V3 p2 = c2 - d2 / 2;
  const V3 r = p1 - p2;
  double a = dot(d1, d1);
  double b = dot(d1, d2);
  double c = dot(d1, r);
  double e = dot(d2, d2);
  double f = dot(d2, r);
  V3 w1, w2;
  if (a <= EPSILON) {
    w1 = p1;
    if (e <= EPSILON) w2 = p2;
    else clamped_linear(w2, p2, f, e, d2);
  } else if (e <= EPSILON) {
    clamped_linear(w1, p1, -c, a, d1);
    w2 = p2;
  } else {
    double denom = std::fmax(a * e - b * b, 0);
    double s, t;
    if (denom > EPSILON) {
      s = clamp01((b * f - c * e), denom);
      t = b * s + f;
    } else {
      s = 0.;
      t = f;
    }
    if (t <= 0.0) {
      w2 = p2;
      clamped_linear(w1, p1, -c, a, d1);
    } else if (t >= e) {
      clamped_linear(w1, p1, (b - c), a, d1);
      w2 = p2 + d2;
    } else {
      w1 = p1 + s * d1;
      w2 = p2 + t / e * d2;
    }
  }
  double distance = norm(w1 - w2);
  distance = distance - (radius1 + radius2);
  normal = normalized(w2 - w1);
  wp1 = w1 + radius1 * normal;
  wp2 = w2 - radius2 * normal;
  return distance;
}

// details.h:699-711
static double computePenetration(const V3& P1, const V3& P2, const V3& P3, const V3& Q1,
                                 const V3& Q2, const V3& Q3, V3& normal) {
  V3 u(cross(P2 - P1, P3 - P1));
  normal = normalized(u);
  double depth1(dot(P1 - Q1, normal));
  double depth2(dot(P1 - Q2, normal));
  double depth3(dot(P1 - Q3, normal));
  return std::max(depth1, std::max(depth2, depth3));
}

// src/distance/triangle_triangle.cpp:47-104
static double triangleTriangleDistance(const Shape& s1, const Tf& tf1, const Shape& s2, const Tf& tf2,
                                       GJKSolver& solver, V3& p1, V3& p2, V3& normal) {
  Shape t1 = s1, t2 = s2;
  for (int i = 0; i < 3; ++i) {
    t1.tri[i] = tf1.transform(s1.tri[i]);
    t2.tri[i] = tf2.transform(s2.tri[i]);
  }
  solver.minkowski_difference.set(&t1, &t2);
  solver.gjk.reset(solver.gjk_max_iterations, solver.gjk_tolerance);
  V3 guess;
  if (solver.gjk_initial_guess == HFB_GUESS_CACHED) {
    guess = solver.cached_guess;
  } else {
    guess = (t1.tri[0] + t1.tri[1] + t1.tri[2] - t2.tri[0] - t2.tri[1] - t2.tri[2]) / 3;
  }
  // `support_func_guess_t support_hint;` is left uninitialised by the reference
  // (:78); only ConvexBase supports read it, triangles do not. Zero here.
  int support_hint[2] = {0, 0};
  solver.epa.status = EPA::DidNotRun;
  GJK::Status gjk_status = solver.gjk.evaluate(solver.minkowski_difference, guess, support_hint);
  solver.cached_guess = solver.gjk.ray;
  solver.support_func_cached_guess[0] = solver.gjk.support_hint[0];
  solver.support_func_cached_guess[1] = solver.gjk.support_hint[1];
  solver.gjk.getWitnessPointsAndNormal(solver.minkowski_difference, p1, p2, normal);
  double distance = solver.gjk.distance;
  if (gjk_status == GJK::Collision) {
    double penetrationDepth =
        computePenetration(t1.tri[0], t1.tri[1], t1.tri[2], t2.tri[0], t2.tri[1], t2.tri[2], normal);
    distance = -penetrationDepth;
  }
  return distance;
}

// ------------------------------------------------------------- dispatch ----
// ---- Plane / Halfspace family (src/narrowphase/details.h:343-693, src/distance/*_halfspace.cpp, *_plane.cpp) ----
// getSupport<WithSweptSphere> (support_functions.cpp:51-91): the support of the shape proper plus
// (radius of a sphere / capsule +) swept-sphere radius along dir.normalized(); a fresh ShapeSupportData and hint 0.
// Box: the `inflate` factor of this instantiation is a function-local static fixed by the first direction the
// process asks for (:146) -- 1 + 1e-10 when that direction has a zero component (a floor with an axis-aligned
// normal), which is the value restated here; tests prime the reference the same way.
static V3 supportWithSweptSphere(const Shape& s, const V3& dir) {
  V3 support;
  int hint = 0;
  SupportData sd;
  sd.last_dir = V3(0, 0, 0);
  switch (s.type) {
    case HFB_GEOM_SPHERE: return (s.p[0] + s.ssr) * normalized(dir);  // :164-176
    case HFB_GEOM_CAPSULE:                                            // :206-222
      getShapeSupport(&s, dir, support, hint, sd);
      support += (s.p[0] + s.ssr) * normalized(dir);
      return support;
    case HFB_GEOM_TRIANGLE:
    case HFB_GEOM_BOX:
    case HFB_GEOM_ELLIPSOID:
    case HFB_GEOM_CONE:
    case HFB_GEOM_CYLINDER:
    case HFB_GEOM_CONVEX:
      getShapeSupport(&s, dir, support, hint, sd);
      support += s.ssr * normalized(dir);
      return support;
    default: return V3(0, 0, 0);  // GEOM_PLANE, GEOM_HALFSPACE, ...: support.setZero() (:82-86)
  }
}
// transform(Halfspace / Plane, tf) (geometric_shapes_utility.cpp:249-277): n' = R n, d' = d + n'.T, then the
// constructor's unitNormalTest (geometric_shapes.cpp:121-143) normalises (n', d') once more
struct WorldPlane { V3 n; double d; };
static WorldPlane unitNormal(V3 n, double d) {
  WorldPlane w;
  const double l = norm(n);
  if (l > 0) {
    const double inv_l = 1.0 / l;
    w.n = n * inv_l;
    w.d = d * inv_l;
  } else {
    w.n = V3(1, 0, 0);
    w.d = 0;
  }
  return w;
}
static WorldPlane transformPlane(const Shape& h, const Tf& tf, bool negate = false) {
  const V3 n = mul(tf.R, V3(h.p[0], h.p[1], h.p[2]));
  const double d = h.d + dot(n, tf.T);
  return negate ? unitNormal(-n, -d) : unitNormal(n, d);  // transformToHalfspaces :279-290
}
// details.h:347-375 (h halfspace, s any other shape)
static double halfspaceDistance(const Shape& h, const Tf& tf1, const Shape& s, const Tf& tf2, V3& p1, V3& p2, V3& normal) {
  const WorldPlane new_h = transformPlane(h, tf1);
  const V3 n_2 = tmul(tf2.R, new_h.n);
  p2 = supportWithSweptSphere(s, -n_2);
  p2 = tf2.transform(p2);
  const double dist = dot(new_h.n, p2) - (new_h.d + h.ssr);  // Halfspace::signedDistance :913-915
  p1 = p2 - dist * new_h.n;
  normal = new_h.n;
  return dist;
}
// details.h:381-428 (plane, s any other shape)
static double planeDistance(const Shape& plane, const Tf& tf1, const Shape& s, const Tf& tf2, V3& p1, V3& p2, V3& normal) {
  const WorldPlane h0 = transformPlane(plane, tf1), h1 = transformPlane(plane, tf1, true);
  const V3 n_h1 = tmul(tf2.R, h0.n), n_h2 = tmul(tf2.R, h1.n);
  V3 p2h1 = supportWithSweptSphere(s, -n_h1);
  p2h1 = tf2.transform(p2h1);
  V3 p2h2 = supportWithSweptSphere(s, -n_h2);
  p2h2 = tf2.transform(p2h2);
  const double dist1 = dot(h0.n, p2h1) - (h0.d + plane.ssr);
  const double dist2 = dot(h1.n, p2h2) - (h1.d + plane.ssr);
  double dist;
  if (dist1 >= dist2) {
    dist = dist1;
    p2 = p2h1;
    p1 = p2 - dist * h0.n;
    normal = h0.n;
  } else {
    dist = dist2;
    p2 = p2h2;
    p1 = p2 - dist * h1.n;
    normal = h1.n;
  }
  return dist;
}
// the non-parallel branch the three pairs below share (details.h:546-560, 607-621, 671-685)
static double intersectionLine(const WorldPlane& a, const WorldPlane& b, const V3& dir, double dir_sq_norm, V3& p1, V3& p2,
                               V3& normal) {
  normal = dir;
  p1 = p2 = cross(b.n * a.d - a.n * b.d, dir) / dir_sq_norm;
  return -(std::numeric_limits<double>::max)();
}
static void planeSweptSpheres(const Shape& s1, const Shape& s2, double& distance, V3& p1, V3& p2, const V3& normal) {
  if (s1.ssr > 0 || s2.ssr > 0) {  // :562-568 etc.
    p1 += s1.ssr * normal;
    p2 -= s2.ssr * normal;
    distance -= (s1.ssr + s2.ssr);
  }
}
// details.h:509-571
static double halfspaceHalfspaceDistance(const Shape& s1, const Tf& tf1, const Shape& s2, const Tf& tf2, V3& p1, V3& p2,
                                         V3& normal) {
  const WorldPlane a = transformPlane(s1, tf1), b = transformPlane(s2, tf2);
  double distance;
  const V3 dir = cross(a.n, b.n);
  const double dir_sq_norm = sqnorm(dir);
  if (dir_sq_norm < std::numeric_limits<double>::epsilon()) {
    if (dot(a.n, b.n) > 0) {
      distance = -(std::numeric_limits<double>::max)();
      if (a.d <= b.d) {
        normal = a.n;
        p1 = normal * distance;
        p2 = b.n * b.d;
      } else {
        normal = -a.n;
        p1 = a.n * a.d;
        p2 = -(normal * distance);
      }
    } else {
      distance = -(a.d + b.d);
      normal = a.n;
      p1 = a.n * a.d;
      p2 = b.n * b.d;
    }
  } else {
    distance = intersectionLine(a, b, dir, dir_sq_norm, p1, p2, normal);
  }
  planeSweptSpheres(s1, s2, distance, p1, p2, normal);
  return distance;
}
// details.h:585-632 (s1 halfspace, s2 plane)
static double halfspacePlaneDistance(const Shape& s1, const Tf& tf1, const Shape& s2, const Tf& tf2, V3& p1, V3& p2,
                                     V3& normal) {
  const WorldPlane a = transformPlane(s1, tf1), b = transformPlane(s2, tf2);
  double distance;
  const V3 dir = cross(a.n, b.n);
  const double dir_sq_norm = sqnorm(dir);
  if (dir_sq_norm < std::numeric_limits<double>::epsilon()) {
    normal = a.n;
    distance = dot(a.n, b.n) > 0 ? (b.d - a.d) : -(a.d + b.d);
    p1 = a.n * a.d;
    p2 = b.n * b.d;
  } else {
    distance = intersectionLine(a, b, dir, dir_sq_norm, p1, p2, normal);
  }
  planeSweptSpheres(s1, s2, distance, p1, p2, normal);
  return distance;
}
// details.h:646-693
static double planePlaneDistance(const Shape& s1, const Tf& tf1, const Shape& s2, const Tf& tf2, V3& p1, V3& p2, V3& normal) {
  const WorldPlane a = transformPlane(s1, tf1), b = transformPlane(s2, tf2);
  double distance;
  const V3 dir = cross(a.n, b.n);
  const double dir_sq_norm = sqnorm(dir);
  if (dir_sq_norm < std::numeric_limits<double>::epsilon()) {
    p1 = a.n * a.d;
    p2 = b.n * b.d;
    distance = norm(p1 - p2);
    if (distance > kDummyPrecision) normal = normalized(p2 - p1);
    else normal = a.n;
  } else {
    distance = intersectionLine(a, b, dir, dir_sq_norm, p1, p2, normal);
  }
  planeSweptSpheres(s1, s2, distance, p1, p2, normal);
  return distance;
}
// the ShapeShapeDistance<S, Halfspace / Plane> specialisations (src/distance/*_halfspace.cpp, *_plane.cpp)
static double planeFamilyDistance(const Shape& s1, const Tf& tf1, const Shape& s2, const Tf& tf2, V3& p1, V3& p2, V3& normal) {
  const int t1 = s1.type, t2 = s2.type;
  const bool h1 = t1 == HFB_GEOM_HALFSPACE, h2 = t2 == HFB_GEOM_HALFSPACE;
  const bool q1 = t1 == HFB_GEOM_PLANE, q2 = t2 == HFB_GEOM_PLANE;
  double distance;
  if (h1 && h2) return halfspaceHalfspaceDistance(s1, tf1, s2, tf2, p1, p2, normal);
  if (q1 && q2) return planePlaneDistance(s1, tf1, s2, tf2, p1, p2, normal);
  if (h1 && q2) return halfspacePlaneDistance(s1, tf1, s2, tf2, p1, p2, normal);
  if (q1 && h2) {
    distance = halfspacePlaneDistance(s2, tf2, s1, tf1, p2, p1, normal);
    normal = -normal;
    return distance;
  }
  if (h1) return halfspaceDistance(s1, tf1, s2, tf2, p1, p2, normal);
  if (q1) return planeDistance(s1, tf1, s2, tf2, p1, p2, normal);
  if (h2) distance = halfspaceDistance(s2, tf2, s1, tf1, p2, p1, normal);
  else distance = planeDistance(s2, tf2, s1, tf1, p2, p1, normal);
  normal = -normal;
  return distance;
}

bool shapeShapeDistance(const Shape& s1, const Tf& tf1, const Shape& s2, const Tf& tf2,
                        GJKSolver& solver, bool compute_signed_distance, double& distance, V3& p1,
                        V3& p2, V3& normal, bool& closed_form) {
  const int t1 = s1.type, t2 = s2.type;
  auto known = [](int t) {
    return t == HFB_GEOM_BOX || t == HFB_GEOM_SPHERE || t == HFB_GEOM_CAPSULE || t == HFB_GEOM_CONE ||
           t == HFB_GEOM_CYLINDER || t == HFB_GEOM_CONVEX || t == HFB_GEOM_TRIANGLE ||
           t == HFB_GEOM_ELLIPSOID || t == HFB_GEOM_PLANE || t == HFB_GEOM_HALFSPACE;
  };
  if (!known(t1) || !known(t2)) return false;
  closed_form = true;
  if (t1 == HFB_GEOM_PLANE || t1 == HFB_GEOM_HALFSPACE || t2 == HFB_GEOM_PLANE || t2 == HFB_GEOM_HALFSPACE) {
    distance = planeFamilyDistance(s1, tf1, s2, tf2, p1, p2, normal);
    return true;
  }
  // specialisations listed in shape_shape_func.h:281-306
  if (t1 == HFB_GEOM_SPHERE && t2 == HFB_GEOM_SPHERE) {
    distance = sphereSphereDistance(s1, tf1, s2, tf2, p1, p2, normal);  // sphere_sphere.cpp
  } else if (t1 == HFB_GEOM_SPHERE && t2 == HFB_GEOM_CAPSULE) {
    distance = sphereCapsuleDistance(s1, tf1, s2, tf2, p1, p2, normal);  // sphere_capsule.cpp
  } else if (t1 == HFB_GEOM_CAPSULE && t2 == HFB_GEOM_SPHERE) {
    distance = sphereCapsuleDistance(s2, tf2, s1, tf1, p2, p1, normal);
    normal = -normal;
  } else if (t1 == HFB_GEOM_SPHERE && t2 == HFB_GEOM_CYLINDER) {
    distance = sphereCylinderDistance(s1, tf1, s2, tf2, p1, p2, normal);  // sphere_cylinder.cpp
  } else if (t1 == HFB_GEOM_CYLINDER && t2 == HFB_GEOM_SPHERE) {
    distance = sphereCylinderDistance(s2, tf2, s1, tf1, p2, p1, normal);
    normal = -normal;
  } else if (t1 == HFB_GEOM_BOX && t2 == HFB_GEOM_SPHERE) {
    distance = boxSphereDistance(s1, tf1, s2, tf2, p1, p2, normal);  // box_sphere.cpp
  } else if (t1 == HFB_GEOM_SPHERE && t2 == HFB_GEOM_BOX) {
    distance = boxSphereDistance(s2, tf2, s1, tf1, p2, p1, normal);
    normal = -normal;
  } else if (t1 == HFB_GEOM_CAPSULE && t2 == HFB_GEOM_CAPSULE) {
    distance = capsuleCapsuleDistance(s1, tf1, s2, tf2, p1, p2, normal);  // capsule_capsule.cpp
  } else if (t1 == HFB_GEOM_TRIANGLE && t2 == HFB_GEOM_SPHERE) {
    distance = sphereTriangleDistance(s2, tf2, s1, tf1, p2, p1, normal);  // triangle_sphere.cpp
    normal = -normal;
  } else if (t1 == HFB_GEOM_SPHERE && t2 == HFB_GEOM_TRIANGLE) {
    distance = sphereTriangleDistance(s1, tf1, s2, tf2, p1, p2, normal);
  } else if (t1 == HFB_GEOM_TRIANGLE && t2 == HFB_GEOM_TRIANGLE) {
    distance = triangleTriangleDistance(s1, tf1, s2, tf2, solver, p1, p2, normal);
  } else {
    closed_form = false;
    distance = solver.shapeDistance(s1, tf1, s2, tf2, compute_signed_distance, p1, p2, normal);
  }
  return true;
}

}  // namespace oracle
