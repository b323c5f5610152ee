// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.hpp).
// Restatement of src/narrowphase/gjk.cpp (GJK :51-1010, EPA :1012-1466) and of
// src/intersect.cpp Project::*Origin (:570-705), operation by operation.
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <atomic>
#include <cassert>

#include "oracle.hpp"

namespace oracle {

static const double kDummyPrecision = 1e-12;  // Eigen::NumTraits<double>::dummy_precision()

// Project::ProjectResult's constructor leaves parameterization[4] uninitialised (internal/intersect.h:58-70), and
// the projections return it untouched for an exactly degenerate simplex (l == 0 / vl == 0): the reference then
// builds its witness points from stack garbage.  The restatement (and the device code) define that case as
// all-zero weights; this counter lets the tests recognise the pairs on which the reference is undefined.
std::atomic<unsigned long long> g_undefined_projections{0};

// ======================================================= Project (intersect.cpp)
ProjectResult projectLineOrigin(const V3& a, const V3& b) {  // :570-595
  ProjectResult res;
  const V3 d = b - a;
  const double l = sqnorm(d);
  if (l > 0) {
    const double t = -dot(a, d);
    res.parameterization[1] = (t >= l) ? 1 : ((t <= 0) ? 0 : (t / l));
    res.parameterization[0] = 1 - res.parameterization[1];
    if (t >= l) {
      res.sqr_distance = sqnorm(b);
      res.encode = 2;
    } else if (t <= 0) {
      res.sqr_distance = sqnorm(a);
      res.encode = 1;
    } else {
      res.sqr_distance = sqnorm(a + d * res.parameterization[1]);
      res.encode = 3;
    }
  } else {
    g_undefined_projections.fetch_add(1, std::memory_order_relaxed);
  }
  return res;
}

ProjectResult projectTriangleOrigin(const V3& a, const V3& b, const V3& c) {  // :596-645
  ProjectResult res;
  static const size_t nexti[3] = {1, 2, 0};
  const V3* vt[] = {&a, &b, &c};
  const V3 dl[] = {a - b, b - c, c - a};
  const V3 n = cross(dl[0], dl[1]);
  const double l = sqnorm(n);
  if (l > 0) {
    double mindist = -1;
    for (size_t i = 0; i < 3; ++i) {
      if (dot(*vt[i], cross(dl[i], n)) > 0) {
        size_t j = nexti[i];
        ProjectResult res_line = projectLineOrigin(*vt[i], *vt[j]);
        if (mindist < 0 || res_line.sqr_distance < mindist) {
          mindist = res_line.sqr_distance;
          res.encode = static_cast<unsigned>(((res_line.encode & 1) ? 1 << i : 0) +
                                             ((res_line.encode & 2) ? 1 << j : 0));
          res.parameterization[i] = res_line.parameterization[0];
          res.parameterization[j] = res_line.parameterization[1];
          res.parameterization[nexti[j]] = 0;
        }
      }
    }
    if (mindist < 0) {
      double d = dot(a, n);
      double s = std::sqrt(l);
      V3 o_to_project = n * (d / l);
      mindist = sqnorm(o_to_project);
      res.encode = 7;
      res.parameterization[0] = norm(cross(dl[1], b - o_to_project)) / s;
      res.parameterization[1] = norm(cross(dl[2], c - o_to_project)) / s;
      res.parameterization[2] = 1 - res.parameterization[0] - res.parameterization[1];
    }
    res.sqr_distance = mindist;
  } else {
    g_undefined_projections.fetch_add(1, std::memory_order_relaxed);
  }
  return res;
}

ProjectResult projectTetrahedraOrigin(const V3& a, const V3& b, const V3& c, const V3& d) {  // :647-705
  ProjectResult res;
  static const size_t nexti[] = {1, 2, 0};
  const V3* vt[] = {&a, &b, &c, &d};
  const V3 dl[3] = {a - d, b - d, c - d};
  double vl = triple(dl[0], dl[1], dl[2]);
  bool ng = (vl * dot(a, cross(b - c, a - b))) <= 0;
  if (ng && std::fabs(vl) > 0) {
    double mindist = -1;
    for (size_t i = 0; i < 3; ++i) {
      size_t j = nexti[i];
      double s = vl * dot(d, cross(dl[i], dl[j]));
      if (s > 0) {
        ProjectResult rt = projectTriangleOrigin(*vt[i], *vt[j], d);
        if (mindist < 0 || rt.sqr_distance < mindist) {
          mindist = rt.sqr_distance;
          res.encode = static_cast<unsigned>((rt.encode & 1 ? 1 << i : 0) +
                                             (rt.encode & 2 ? 1 << j : 0) + (rt.encode & 4 ? 8 : 0));
          res.parameterization[i] = rt.parameterization[0];
          res.parameterization[j] = rt.parameterization[1];
          res.parameterization[nexti[j]] = 0;
          res.parameterization[3] = rt.parameterization[2];
        }
      }
    }
    if (mindist < 0) {
      mindist = 0;
      res.encode = 15;
      res.parameterization[0] = triple(c, b, d) / vl;
      res.parameterization[1] = triple(a, c, d) / vl;
      res.parameterization[2] = triple(b, a, d) / vl;
      res.parameterization[3] =
          1 - (res.parameterization[0] + res.parameterization[1] + res.parameterization[2]);
    }
    res.sqr_distance = mindist;
  } else if (!ng) {
    res = projectTriangleOrigin(a, b, c);
    res.parameterization[3] = 0;
  } else {
    g_undefined_projections.fetch_add(1, std::memory_order_relaxed);
  }
  return res;
}

// ======================================================================= GJK
GJK::GJK(size_t max_it, double tol) : max_iterations(max_it), tolerance(tol) {  // gjk.cpp:51-57
  distance_upper_bound = std::numeric_limits<double>::max();
  gjk_variant = HFB_GJK_DEFAULT;
  convergence_criterion = HFB_CRIT_DEFAULT;
  convergence_criterion_type = HFB_CRIT_RELATIVE;
  shape = nullptr;
  distance = 0;
  support_hint[0] = support_hint[1] = 0;
  current = 0;
  reset(max_it, tol);
}

void GJK::reset(size_t max_it, double tol) {  // :59-69
  max_iterations = max_it;
  tolerance = tol;
  status = DidNotRun;
  nfree = 0;
  simplex = nullptr;
  iterations = 0;
  iterations_momentum_stop = 0;
}

void GJK::getSupport(const V3& d, SimplexV& sv, int hint[2]) const {  // gjk.h:163-167
  shape->support(d, sv.w0, sv.w1, hint);
  sv.w = sv.w0 - sv.w1;
}

// details::getClosestPoints, gjk.cpp:94-151
static void getClosestPoints(const GJK::Simplex& simplex, V3& w0, V3& w1) {
  SimplexV* const* vs = simplex.vertex;
  ProjectResult projection;
  switch (simplex.rank) {
    case 1:
      w0 = vs[0]->w0;
      w1 = vs[0]->w1;
      return;
    case 2: {
      const V3 &a = vs[0]->w, a0 = vs[0]->w0, a1 = vs[0]->w1, b = vs[1]->w, b0 = vs[1]->w0,
               b1 = vs[1]->w1;
      double la, lb;
      V3 N(b - a);
      la = dot(N, -a);
      if (la <= 0) {
        w0 = a0;
        w1 = a1;
      } else {
        lb = sqnorm(N);
        if (la > lb) {
          w0 = b0;
          w1 = b1;
        } else {
          lb = la / lb;
          la = 1 - lb;
          w0 = la * a0 + lb * b0;
          w1 = la * a1 + lb * b1;
        }
      }
    }
      return;
    case 3:
      projection = projectTriangleOrigin(vs[0]->w, vs[1]->w, vs[2]->w);
      break;
    case 4:
      projection = projectTetrahedraOrigin(vs[0]->w, vs[1]->w, vs[2]->w, vs[3]->w);
      break;
    default:
      assert(false);
  }
  w0 = V3();
  w1 = V3();
  for (unsigned i = 0; i < simplex.rank; ++i) {
    w0 += projection.parameterization[i] * vs[i]->w0;
    w1 += projection.parameterization[i] * vs[i]->w1;
  }
}

// details::inflate, gjk.cpp:158-173
static void inflate(const MinkowskiDiff& shape, const V3& normal, V3& w0, V3& w1) {
  const double* I = shape.swept_sphere_radius;
  bool i0 = I[0] > 0, i1 = I[1] > 0;
  if (!(i0 || i1)) return;
  if (i0) w0 += I[0] * normal;
  if (i1) w1 -= I[1] * normal;
}

void GJK::getWitnessPointsAndNormal(const MinkowskiDiff& shape_, V3& w0, V3& w1,
                                    V3& normal) const {  // :177-186
  getClosestPoints(*simplex, w0, w1);
  if (getenv("HFB_ORACLE_DUMP_SIMPLEX")) {  // debugging aid (single-pair runs)
    fprintf(stderr, "oracle simplex rank %d ray %.17g %.17g %.17g\n", (int)simplex->rank, ray.x, ray.y, ray.z);
    for (int k = 0; k < (int)simplex->rank; ++k) {
      const SimplexV* v = simplex->vertex[k];
      fprintf(stderr, "  w0 %.17g %.17g %.17g  w1 %.17g %.17g %.17g\n", v->w0.x, v->w0.y, v->w0.z, v->w1.x, v->w1.y, v->w1.z);
    }
    fprintf(stderr, "  closest w0 %.17g %.17g %.17g  w1 %.17g %.17g %.17g\n", w0.x, w0.y, w0.z, w1.x, w1.y, w1.z);
  }
  if (norm(w1 - w0) > kDummyPrecision) {
    normal = normalized(w1 - w0);
  } else {
    normal = -normalized(ray);
  }
  inflate(shape_, normal, w0, w1);
}

GJK::Status GJK::evaluate(MinkowskiDiff& shape_, const V3& guess, const int hint_in[2]) {  // :188-370
  double alpha = 0;
  iterations = 0;
  const double swept_sphere_radius = shape_.swept_sphere_radius[0] + shape_.swept_sphere_radius[1];
  const double upper_bound = distance_upper_bound + swept_sphere_radius;

  free_v[0] = &store_v[0];
  free_v[1] = &store_v[1];
  free_v[2] = &store_v[2];
  free_v[3] = &store_v[3];

  nfree = 4;
  status = NoCollision;
  shape = &shape_;
  distance = 0.0;
  current = 0;
  simplices[current].rank = 0;
  support_hint[0] = hint_in[0];
  support_hint[1] = hint_in[1];

  double rl = norm(guess);
  if (rl < tolerance) {
    ray = V3(-1, 0, 0);
    rl = 1;
  } else
    ray = guess;

  int current_gjk_variant = gjk_variant;
  V3 w = ray;
  V3 dir = ray;
  V3 y;
  double momentum;
  bool normalize_support_direction = shape->normalize_support_direction;
  do {
    unsigned char next = (unsigned char)(1 - current);
    Simplex& curr_simplex = simplices[current];
    Simplex& next_simplex = simplices[next];

    // check A
    if (rl < tolerance) {
      status = Collision;
      distance = rl;
      break;
    }

    switch (current_gjk_variant) {
      case HFB_GJK_DEFAULT:
        dir = ray;
        break;
      case HFB_GJK_NESTEROV:
        if (normalize_support_direction) {
          momentum = (double(iterations) + 2) / (double(iterations) + 3);
          y = momentum * ray + (1 - momentum) * w;
          double y_norm = norm(y);
          dir = momentum * dir / norm(dir) + (1 - momentum) * y / y_norm;
        } else {
          momentum = (double(iterations) + 1) / (double(iterations) + 3);
          y = momentum * ray + (1 - momentum) * w;
          dir = momentum * dir + (1 - momentum) * y;
        }
        break;
      case HFB_GJK_POLYAK:
        momentum = 1 / (double(iterations) + 1);
        dir = momentum * dir + (1 - momentum) * ray;
        break;
      default:
        assert(false);
    }

    appendVertex(curr_simplex, -dir, support_hint);
    w = curr_simplex.vertex[curr_simplex.rank - 1]->w;

    // check B
    double omega = dot(dir, w) / norm(dir);
    if (omega > upper_bound) {
      distance = omega - swept_sphere_radius;
      status = NoCollisionEarlyStopped;
      break;
    }

    if (current_gjk_variant != HFB_GJK_DEFAULT) {
      double frank_wolfe_duality_gap = 2 * dot(ray, ray - w);
      if (frank_wolfe_duality_gap - tolerance <= 0) {
        removeVertex(simplices[current]);
        current_gjk_variant = HFB_GJK_DEFAULT;
        iterations_momentum_stop = iterations;
        continue;
      }
    }

    // check C
    bool cv_check_passed = checkConvergence(w, rl, alpha, omega);
    if (iterations > 0 && cv_check_passed) {
      if (iterations > 0) removeVertex(simplices[current]);
      if (current_gjk_variant != HFB_GJK_DEFAULT) {
        current_gjk_variant = HFB_GJK_DEFAULT;
        iterations_momentum_stop = iterations;
        continue;
      }
      distance = rl - swept_sphere_radius;
      if (distance < tolerance) {
        status = CollisionWithPenetrationInformation;
      } else {
        status = NoCollision;
      }
      break;
    }

    bool inside;
    switch (curr_simplex.rank) {
      case 1:
        ray = w;
        inside = false;
        next_simplex.rank = 1;
        next_simplex.vertex[0] = curr_simplex.vertex[0];
        break;
      case 2:
        inside = projectLineOrigin(curr_simplex, next_simplex);
        break;
      case 3:
        inside = projectTriangleOrigin(curr_simplex, next_simplex);
        break;
      case 4:
        inside = projectTetrahedraOrigin(curr_simplex, next_simplex);
        break;
      default:
        inside = false;
        assert(false);
    }
    current = next;
    rl = norm(ray);
    if (inside || rl == 0) {
      status = Collision;
      distance = rl;
      break;
    }

    status = ((++iterations) < max_iterations) ? status : Failed;

  } while (status == NoCollision);

  simplex = &simplices[current];
  return status;
}

bool GJK::checkConvergence(const V3& w, const double& rl, double& alpha,
                           const double& omega) const {  // :372-425
  switch (convergence_criterion) {
    case HFB_CRIT_DEFAULT: {
      alpha = std::max(alpha, omega);
      const double diff = rl - alpha;
      return ((diff - (tolerance + tolerance * rl)) <= 0);
    }
    case HFB_CRIT_DUALITY_GAP: {
      const double diff = 2 * dot(ray, ray - w);
      switch (convergence_criterion_type) {
        case HFB_CRIT_ABSOLUTE:
          return ((diff - tolerance) <= 0);
        case HFB_CRIT_RELATIVE:
          return (((diff / tolerance * rl) - tolerance * rl) <= 0);
      }
    } break;
    case HFB_CRIT_HYBRID: {
      alpha = std::max(alpha, omega);
      const double diff = rl * rl - alpha * alpha;
      switch (convergence_criterion_type) {
        case HFB_CRIT_ABSOLUTE:
          return ((diff - tolerance) <= 0);
        case HFB_CRIT_RELATIVE:
          return (((diff / tolerance * rl) - tolerance * rl) <= 0);
      }
    } break;
  }
  assert(false);
  return false;
}

void GJK::removeVertex(Simplex& s) { free_v[nfree++] = s.vertex[--s.rank]; }  // :427-429

void GJK::appendVertex(Simplex& s, const V3& v, int hint[2]) {  // :431-435
  s.vertex[s.rank] = free_v[--nfree];
  getSupport(v, *s.vertex[s.rank++], hint);
}

bool GJK::encloseOrigin() {  // :437-492
  V3 axis(0, 0, 0);
  int hint[2] = {0, 0};
  switch (simplex->rank) {
    case 1:
      for (int i = 0; i < 3; ++i) {
        axis[i] = 1;
        appendVertex(*simplex, axis, hint);
        if (encloseOrigin()) return true;
        removeVertex(*simplex);
        axis[i] = -1;
        appendVertex(*simplex, -axis, hint);
        if (encloseOrigin()) return true;
        removeVertex(*simplex);
        axis[i] = 0;
      }
      break;
    case 2: {
      V3 d = simplex->vertex[1]->w - simplex->vertex[0]->w;
      for (int i = 0; i < 3; ++i) {
        axis[i] = 1;
        V3 p = cross(d, axis);
        if (!is_zero(p)) {
          appendVertex(*simplex, p, hint);
          if (encloseOrigin()) return true;
          removeVertex(*simplex);
          appendVertex(*simplex, -p, hint);
          if (encloseOrigin()) return true;
          removeVertex(*simplex);
        }
        axis[i] = 0;
      }
    } break;
    case 3:
      axis = cross(simplex->vertex[1]->w - simplex->vertex[0]->w,
                   simplex->vertex[2]->w - simplex->vertex[0]->w);
      if (!is_zero(axis)) {
        appendVertex(*simplex, axis, hint);
        if (encloseOrigin()) return true;
        removeVertex(*simplex);
        appendVertex(*simplex, -axis, hint);
        if (encloseOrigin()) return true;
        removeVertex(*simplex);
      }
      break;
    case 4:
      if (std::fabs(triple(simplex->vertex[0]->w - simplex->vertex[3]->w,
                           simplex->vertex[1]->w - simplex->vertex[3]->w,
                           simplex->vertex[2]->w - simplex->vertex[3]->w)) > 0)
        return true;
      break;
  }
  return false;
}

// :494-541
static inline void originToPoint(const GJK::Simplex& current, unsigned a, const V3& A,
                                 GJK::Simplex& next, V3& ray) {
  ray = A;
  next.vertex[0] = current.vertex[a];
  next.rank = 1;
}
static inline void originToSegment(const GJK::Simplex& current, unsigned a, unsigned b,
                                   const V3& A, const V3& B, const V3& AB, const double& ABdotAO,
                                   GJK::Simplex& next, V3& ray) {
  ray = dot(AB, B) * A + ABdotAO * B;
  next.vertex[0] = current.vertex[b];
  next.vertex[1] = current.vertex[a];
  next.rank = 2;
  ray /= sqnorm(AB);
}
static inline bool originToTriangle(const GJK::Simplex& current, unsigned a, unsigned b,
                                    unsigned c, const V3& ABC, const double& ABCdotAO,
                                    GJK::Simplex& next, V3& ray) {
  next.rank = 3;
  next.vertex[2] = current.vertex[a];
  if (ABCdotAO == 0) {
    next.vertex[0] = current.vertex[c];
    next.vertex[1] = current.vertex[b];
    ray = V3(0, 0, 0);
    return true;
  }
  if (ABCdotAO > 0) {
    next.vertex[0] = current.vertex[c];
    next.vertex[1] = current.vertex[b];
  } else {
    next.vertex[0] = current.vertex[b];
    next.vertex[1] = current.vertex[c];
  }
  ray = -ABCdotAO / sqnorm(ABC) * ABC;
  return false;
}

bool GJK::projectLineOrigin(const Simplex& current, Simplex& next) {  // :543-569
  const unsigned a = 1, b = 0;
  const V3& A = current.vertex[a]->w;
  const V3& B = current.vertex[b]->w;
  const V3 AB = B - A;
  const double d = dot(AB, -A);
  if (d == 0) {
    originToPoint(current, a, A, next, ray);
    free_v[nfree++] = current.vertex[b];
    return is_zero(A);
  } else if (d < 0) {
    originToPoint(current, a, A, next, ray);
    free_v[nfree++] = current.vertex[b];
  } else
    originToSegment(current, a, b, A, B, AB, d, next, ray);
  return false;
}

bool GJK::projectTriangleOrigin(const Simplex& current, Simplex& next) {  // :571-611
  const unsigned a = 2, b = 1, c = 0;
  const V3 &A = current.vertex[a]->w, B = current.vertex[b]->w, C = current.vertex[c]->w;
  const V3 AB = B - A, AC = C - A, ABC = cross(AB, AC);

  double edgeAC2o = dot(cross(ABC, AC), -A);
  if (edgeAC2o >= 0) {
    double towardsC = dot(AC, -A);
    if (towardsC >= 0) {
      originToSegment(current, a, c, A, C, AC, towardsC, next, ray);
      free_v[nfree++] = current.vertex[b];
    } else {
      double towardsB = dot(AB, -A);
      if (towardsB < 0) {
        originToPoint(current, a, A, next, ray);
        free_v[nfree++] = current.vertex[b];
      } else
        originToSegment(current, a, b, A, B, AB, towardsB, next, ray);
      free_v[nfree++] = current.vertex[c];
    }
  } else {
    double edgeAB2o = dot(cross(AB, ABC), -A);
    if (edgeAB2o >= 0) {
      double towardsB = dot(AB, -A);
      if (towardsB < 0) {
        originToPoint(current, a, A, next, ray);
        free_v[nfree++] = current.vertex[b];
      } else
        originToSegment(current, a, b, A, B, AB, towardsB, next, ray);
      free_v[nfree++] = current.vertex[c];
    } else {
      return originToTriangle(current, a, b, c, ABC, dot(ABC, -A), next, ray);
    }
  }
  return false;
}

// :613-1010.  The reference's 24-leaf decision tree (generated by doc/gjk.py)
// is a pure function of twelve sign predicates; it is restated here with the
// same nesting, each leaf naming the Voronoi region it selects.
bool GJK::projectTetrahedraOrigin(const Simplex& current, Simplex& next) {
  const unsigned a = 3, b = 2, c = 1, d = 0;
  const V3& A = current.vertex[a]->w;
  const V3& B = current.vertex[b]->w;
  const V3& C = current.vertex[c]->w;
  const V3& D = current.vertex[d]->w;
  const double aa = sqnorm(A);
  const double da = dot(D, A);
  const double db = dot(D, B);
  const double dc = dot(D, C);
  const double dd = dot(D, D);
  const double da_aa = da - aa;
  const double ca = dot(C, A);
  const double cb = dot(C, B);
  const double cc = dot(C, C);
  const double& cd = dc;
  const double ca_aa = ca - aa;
  const double ba = dot(B, A);
  const double bb = dot(B, B);
  const double& bc = cb;
  const double& bd = db;
  const double ba_aa = ba - aa;
  const double ba_ca = ba - ca;
  const double ca_da = ca - da;
  const double da_ba = da - ba;
  const V3 a_cross_b = cross(A, B);
  const V3 a_cross_c = cross(A, C);

  // predicates, named as in the generated comments of the reference
#define P_a10 (ba_aa <= 0)
#define P_a11 (ca_aa <= 0)
#define P_a12 (da_aa <= 0)
#define P_a3 (-dot(D, a_cross_b) <= 0)
#define P_a1 (dot(C, a_cross_b) <= 0)
#define P_a2 (dot(D, a_cross_c) <= 0)
#define P_a9 (ba * da_ba + bd * ba_aa - bb * da_aa <= 0)
#define P_a8 (da * da_ba + dd * ba_aa - db * da_aa <= 0)
#define P_a4 (ba * ba_ca + bb * ca_aa - bc * ba_aa <= 0)
#define P_a5 (ca * ba_ca + cb * ca_aa - cc * ba_aa <= 0)
#define P_a6 (ca * ca_da + cc * da_aa - cd * ca_aa <= 0)
#define P_a7 (da * ca_da + dc * da_aa - dd * ca_aa <= 0)

  auto R_ABC = [&]() {
    originToTriangle(current, a, b, c, cross(B - A, C - A), -dot(C, a_cross_b), next, ray);
    free_v[nfree++] = current.vertex[d];
  };
  auto R_ACD = [&]() {
    originToTriangle(current, a, c, d, cross(C - A, D - A), -dot(D, a_cross_c), next, ray);
    free_v[nfree++] = current.vertex[b];
  };
  auto R_ADB = [&]() {
    originToTriangle(current, a, d, b, cross(D - A, B - A), dot(D, a_cross_b), next, ray);
    free_v[nfree++] = current.vertex[c];
  };
  auto R_AB = [&]() {
    originToSegment(current, a, b, A, B, B - A, -ba_aa, next, ray);
    free_v[nfree++] = current.vertex[c];
    free_v[nfree++] = current.vertex[d];
  };
  auto R_AC = [&]() {
    originToSegment(current, a, c, A, C, C - A, -ca_aa, next, ray);
    free_v[nfree++] = current.vertex[b];
    free_v[nfree++] = current.vertex[d];
  };
  auto R_AD = [&]() {
    originToSegment(current, a, d, A, D, D - A, -da_aa, next, ray);
    free_v[nfree++] = current.vertex[b];
    free_v[nfree++] = current.vertex[c];
  };
  auto R_A = [&]() {
    originToPoint(current, a, A, next, ray);
    free_v[nfree++] = current.vertex[b];
    free_v[nfree++] = current.vertex[c];
    free_v[nfree++] = current.vertex[d];
  };
#define R_INSIDE()                    \
  ray = V3(0, 0, 0);                  \
  next.vertex[0] = current.vertex[d]; \
  next.vertex[1] = current.vertex[c]; \
  next.vertex[2] = current.vertex[b]; \
  next.vertex[3] = current.vertex[a]; \
  next.rank = 4;                      \
  return true;

  if (P_a10) {
    if (P_a3) {
      if (P_a9) {
        if (P_a12) {
          if (P_a4) R_ABC(); else R_AB();
        } else {
          if (P_a4) {
            if (P_a5) {
              if (P_a6) R_ACD(); else R_AC();
            } else R_ABC();
          } else R_AB();
        }
      } else {
        if (P_a8) R_ADB();
        else {
          if (P_a6) {
            if (P_a7) R_AD(); else R_ACD();
          } else {
            if (P_a7) R_AD(); else R_AC();
          }
        }
      }
    } else {
      if (P_a1) {
        if (P_a4) {
          if (P_a5) {
            if (P_a6) R_ACD(); else R_AC();
          } else R_ABC();
        } else R_AB();
      } else {
        if (P_a2) {
          if (P_a6) {
            if (P_a7) R_AD(); else R_ACD();
          } else {
            if (P_a11) R_AC(); else R_AD();
          }
        } else { R_INSIDE() }
      }
    }
  } else {
    if (P_a11) {
      if (P_a2) {
        if (P_a12) {
          if (P_a6) {
            if (P_a7) {
              if (P_a8) R_ADB(); else R_AD();
            } else R_ACD();
          } else {
            if (P_a5) R_AC(); else R_ABC();
          }
        } else {
          if (P_a5) {
            if (P_a6) R_ACD(); else R_AC();
          } else {
            if (P_a1) R_ABC(); else R_ACD();
          }
        }
      } else {
        if (P_a1) {
          if (P_a5) R_AC(); else R_ABC();
        } else {
          if (P_a3) {
            if (P_a8) R_ADB(); else R_AD();
          } else { R_INSIDE() }
        }
      }
    } else {
      if (P_a12) {
        if (P_a3) {
          if (P_a7) {
            if (P_a8) R_ADB(); else R_AD();
          } else {
            if (P_a2) R_ACD();
            else R_ADB();  // both sub-branches on ABC.AO select ADB (:958-972)
          }
        } else {
          if (P_a2) {
            if (P_a7) R_AD(); else R_ACD();
          } else { R_INSIDE() }
        }
      } else R_A();
    }
  }
#undef R_INSIDE
#undef P_a10
#undef P_a11
#undef P_a12
#undef P_a3
#undef P_a1
#undef P_a2
#undef P_a9
#undef P_a8
#undef P_a4
#undef P_a5
#undef P_a6
#undef P_a7
  return false;
}

// ======================================================================= EPA
void EPA::FaceList::append(Face* face) {  // gjk.h:292-298
  face->prev_face = nullptr;
  face->next_face = root;
  if (root != nullptr) root->prev_face = face;
  root = face;
  ++count;
}
void EPA::FaceList::remove(Face* face) {  // gjk.h:300-307
  if (face->next_face != nullptr) face->next_face->prev_face = face->prev_face;
  if (face->prev_face != nullptr) face->prev_face->next_face = face->next_face;
  if (face == root) root = face->next_face;
  --count;
}
static inline void bind(EPA::Face* fa, size_t ea, EPA::Face* fb, size_t eb) {  // gjk.h:312-320
  fa->adjacent_edge[ea] = eb;
  fa->adjacent_faces[ea] = fb;
  fb->adjacent_edge[eb] = ea;
  fb->adjacent_faces[eb] = fa;
}

EPA::EPA(size_t max_it, double tol) : max_iterations(max_it), tolerance(tol) { reset(max_it, tol); }

void EPA::reset(size_t max_it, double tol) {  // gjk.cpp:1014-1037
  max_iterations = max_it;
  tolerance = tol;
  sv_store.resize(max_iterations + 4);
  fc_store.resize(2 * max_iterations + 4);
  status = DidNotRun;
  normal = V3(0, 0, 0);
  support_hint[0] = support_hint[1] = 0;
  depth = 0;
  closest_face = nullptr;
  result.rank = 0;
  for (int i = 0; i < 4; ++i) result.vertex[i] = nullptr;
  hull.reset();
  num_vertices = 0;
  stock.reset();
  for (size_t i = 0; i < fc_store.size(); ++i) stock.append(&fc_store[fc_store.size() - i - 1]);
  iterations = 0;
}

EPA::Face* EPA::newFace(size_t id_a, size_t id_b, size_t id_c, bool force) {  // :1068-1138
  if (stock.root != nullptr) {
    Face* face = stock.root;
    stock.remove(face);
    hull.append(face);
    face->pass = 0;
    face->vertex_id[0] = id_a;
    face->vertex_id[1] = id_b;
    face->vertex_id[2] = id_c;
    const SimplexV& a = sv_store[id_a];
    const SimplexV& b = sv_store[id_b];
    const SimplexV& c = sv_store[id_c];
    face->n = cross(b.w - a.w, c.w - a.w);

    if (norm(face->n) > DBL_EPSILON) {
      face->n = normalized(face->n);  // n.normalize()
      double a_dot_nab = dot(a.w, cross(b.w - a.w, face->n));
      double b_dot_nbc = dot(b.w, cross(c.w - b.w, face->n));
      double c_dot_nca = dot(c.w, cross(a.w - c.w, face->n));
      if (a_dot_nab >= -tolerance && b_dot_nbc >= -tolerance && c_dot_nca >= -tolerance) {
        face->d = dot(a.w, face->n);
        face->ignore = false;
      } else {
        face->d = std::numeric_limits<double>::max();
        face->ignore = true;
      }
      if (face->d >= -tolerance || force)
        return face;
      else
        status = NonConvex;
    } else
      status = Degenerated;

    hull.remove(face);
    stock.append(face);
    return nullptr;
  }
  status = OutOfFaces;
  return nullptr;
}

EPA::Face* EPA::findClosestFace() {  // :1141-1154
  Face* minf = hull.root;
  double mind = std::numeric_limits<double>::max();
  for (Face* f = minf; f; f = f->next_face) {
    if (f->ignore) continue;
    double sqd = f->d * f->d;
    if (sqd < mind) {
      minf = f;
      mind = sqd;
    }
  }
  return minf;
}

EPA::Status EPA::evaluate(GJK& gjk, const V3& guess) {  // :1156-1316
  GJK::Simplex& simplex = *gjk.simplex;
  support_hint[0] = gjk.support_hint[0];
  support_hint[1] = gjk.support_hint[1];

  bool enclosed_origin = gjk.encloseOrigin();
  if ((simplex.rank > 1) && enclosed_origin) {
    while (hull.root) {
      Face* f = hull.root;
      hull.remove(f);
      stock.append(f);
    }
    status = Valid;
    num_vertices = 0;

    if (dot(simplex.vertex[0]->w - simplex.vertex[3]->w,
            cross(simplex.vertex[1]->w - simplex.vertex[3]->w,
                  simplex.vertex[2]->w - simplex.vertex[3]->w)) < 0) {
      SimplexV* tmp = simplex.vertex[0];
      simplex.vertex[0] = simplex.vertex[1];
      simplex.vertex[1] = tmp;
    }
    for (size_t i = 0; i < 4; ++i) sv_store[num_vertices++] = *simplex.vertex[i];

    // brace-init-list: evaluated left to right
    Face* t0 = newFace(0, 1, 2, true);
    Face* t1 = newFace(1, 0, 3, true);
    Face* t2 = newFace(2, 1, 3, true);
    Face* t3 = newFace(0, 2, 3, true);
    Face* tetrahedron[] = {t0, t1, t2, t3};

    if (hull.count == 4) {
      bind(tetrahedron[0], 0, tetrahedron[1], 0);
      bind(tetrahedron[0], 1, tetrahedron[2], 0);
      bind(tetrahedron[0], 2, tetrahedron[3], 0);
      bind(tetrahedron[1], 1, tetrahedron[3], 2);
      bind(tetrahedron[1], 2, tetrahedron[2], 1);
      bind(tetrahedron[2], 2, tetrahedron[3], 1);

      closest_face = findClosestFace();
      Face outer = *closest_face;

      status = Valid;
      iterations = 0;
      size_t pass = 0;
      for (; iterations < max_iterations; ++iterations) {
        if (num_vertices >= sv_store.size()) {
          status = OutOfVertices;
          break;
        }
        Horizon horizon;
        SimplexV& w = sv_store[num_vertices++];
        bool valid = true;
        closest_face->pass = ++pass;
        gjk.getSupport(closest_face->n, w, support_hint);

        const SimplexV& vf1 = sv_store[closest_face->vertex_id[0]];
        const SimplexV& vf2 = sv_store[closest_face->vertex_id[1]];
        const SimplexV& vf3 = sv_store[closest_face->vertex_id[2]];
        double fdist = dot(closest_face->n, w.w - vf1.w);
        double wnorm = norm(w.w);
        if (fdist <= tolerance + tolerance * wnorm) {
          status = AccuracyReached;
          break;
        }
        if (norm(w.w - vf1.w) <= tolerance + tolerance * wnorm ||
            norm(w.w - vf2.w) <= tolerance + tolerance * wnorm ||
            norm(w.w - vf3.w) <= tolerance + tolerance * wnorm) {
          status = AccuracyReached;
          break;
        }

        for (size_t j = 0; (j < 3) && valid; ++j)
          valid &= expand(pass, w, closest_face->adjacent_faces[j],
                          closest_face->adjacent_edge[j], horizon);

        if (!valid || horizon.num_faces < 3) {
          break;
        }
        bind(horizon.first_face, 2, horizon.current_face, 1);
        hull.remove(closest_face);
        stock.append(closest_face);
        closest_face = findClosestFace();
        outer = *closest_face;
      }

      status = ((iterations) < max_iterations) ? status : Failed;
      normal = outer.n;
      depth = outer.d + (gjk.shape->swept_sphere_radius[0] + gjk.shape->swept_sphere_radius[1]);
      result.rank = 3;
      result.vertex[0] = &sv_store[outer.vertex_id[0]];
      result.vertex[1] = &sv_store[outer.vertex_id[1]];
      result.vertex[2] = &sv_store[outer.vertex_id[2]];
      return status;
    }
  }

  // FallBack (:1299-1315)
  status = FallBack;
  normal = -guess;
  double nl = norm(normal);
  if (nl > 0)
    normal /= nl;
  else
    normal = V3(1, 0, 0);
  depth = 0;
  result.rank = 1;
  result.vertex[0] = simplex.vertex[0];
  return status;
}

bool EPA::expand(size_t pass, const SimplexV& w, Face* f, size_t e, Horizon& horizon) {  // :1361-1449
  static const size_t nexti[] = {1, 2, 0};
  static const size_t previ[] = {2, 0, 1};
  const size_t id_w = num_vertices - 1;

  if (f->pass == pass) {
    status = InvalidHull;
    return false;
  }
  const size_t e1 = nexti[e];
  const double dummy_precision(3 * std::sqrt(std::numeric_limits<double>::epsilon()));
  const SimplexV& vf = sv_store[f->vertex_id[e]];
  if (dot(f->n, w.w - vf.w) < dummy_precision) {
    Face* new_face = newFace(f->vertex_id[e1], f->vertex_id[e], id_w);
    if (new_face != nullptr) {
      bind(new_face, 0, f, e);
      if (horizon.current_face != nullptr) {
        bind(new_face, 2, horizon.current_face, 1);
      } else {
        horizon.first_face = new_face;
      }
      horizon.current_face = new_face;
      ++horizon.num_faces;
      return true;
    }
    return false;
  }

  const size_t e2 = previ[e];
  f->pass = pass;
  if (expand(pass, w, f->adjacent_faces[e1], f->adjacent_edge[e1], horizon) &&
      expand(pass, w, f->adjacent_faces[e2], f->adjacent_edge[e2], horizon)) {
    hull.remove(f);
    stock.append(f);
    return true;
  }
  return false;
}

void EPA::getWitnessPointsAndNormal(const MinkowskiDiff& shape, V3& w0, V3& w1,
                                    V3& normal_) const {  // :1451-1466
  getClosestPoints(result, w0, w1);
  if (norm(w0 - w1) > kDummyPrecision) {
    if (this->depth >= 0) {
      normal_ = normalized(w0 - w1);
    } else {
      normal_ = normalized(w1 - w0);
    }
  } else {
    normal_ = this->normal;
  }
  inflate(shape, normal_, w0, w1);
}

}  // namespace oracle
