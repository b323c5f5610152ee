// ORACLE -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the hpp-fcl hot path (reference snapshot 3224f2ba, package
// 3.0.0) used as the parity checker for the CUDA engine.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may build, link or call anything in this directory; the product library
// (hpp-fcl_b200/) never does.
//
// Parity status: PINNED (1) against the reference's own known-answer tests
// (tests/test_oracle_golden.py cites each test/file:line) and (2) bit for bit against
// the reference's own sources compiled in place into oracle/_ref/ (`make ref`;
// Eigen3 and Boost, hard requirements of the reference's headers, are not installed:
// ref_shim/ is a minimal stand-in for the fixed-size Eigen API they use) --
// tests/test_reference_build.py.
//
// Floating-point convention ("Floating-point order
// caveat"): every 3-vector reduction is evaluated left to right,
// (x*x + y*y) + z*z, which is what Eigen >= 3.3 emits for fixed 3-vectors on
// x86-64 (packet of two + scalar remainder); no FMA contraction
// (-ffp-contract=off), IEEE sqrt/div.
#pragma once
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <vector>

#include "../include/hppfcl_b200.h"

namespace oracle {

// ---------------------------------------------------------------- math -----
struct V3 {
  double x, y, z;
  V3() : x(0), y(0), z(0) {}
  V3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
  double& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(const V3& a, const V3& b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(const V3& a, const V3& b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(const V3& a) { return V3(-a.x, -a.y, -a.z); }
inline V3 operator*(double s, const V3& a) { return V3(s * a.x, s * a.y, s * a.z); }
inline V3 operator*(const V3& a, double s) { return V3(a.x * s, a.y * s, a.z * s); }
inline V3 operator/(const V3& a, double s) { return V3(a.x / s, a.y / s, a.z / s); }
inline V3& operator+=(V3& a, const V3& b) { a = a + b; return a; }
inline V3& operator-=(V3& a, const V3& b) { a = a - b; return a; }
inline V3& operator*=(V3& a, double s) { a = a * s; return a; }
inline V3& operator/=(V3& a, double s) { a = a / s; return a; }
inline double dot(const V3& a, const V3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline double sqnorm(const V3& a) { return dot(a, a); }
inline double norm(const V3& a) { return std::sqrt(sqnorm(a)); }
inline V3 cross(const V3& a, const V3& b) {
  return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// Eigen MatrixBase::normalized(): divide by sqrt(squaredNorm) iff squaredNorm > 0
inline V3 normalized(const V3& a) {
  double z = sqnorm(a);
  if (z > 0) return a / std::sqrt(z);
  return a;
}
inline double triple(const V3& a, const V3& b, const V3& c) { return dot(a, cross(b, c)); }
// Eigen isZero(prec): every |coeff| <= prec
inline bool is_zero(const V3& a, double prec = 1e-12) {
  return std::fabs(a.x) <= prec && std::fabs(a.y) <= prec && std::fabs(a.z) <= prec;
}
inline V3 nan3() {
  double q = std::numeric_limits<double>::quiet_NaN();
  return V3(q, q, q);
}

// 3x3, m[r][c]
struct M3 {
  double m[3][3];
  V3 col(int c) const { return V3(m[0][c], m[1][c], m[2][c]); }
};
inline V3 mul(const M3& A, const V3& v) {  // A * v
  return V3((A.m[0][0] * v.x + A.m[0][1] * v.y) + A.m[0][2] * v.z,
            (A.m[1][0] * v.x + A.m[1][1] * v.y) + A.m[1][2] * v.z,
            (A.m[2][0] * v.x + A.m[2][1] * v.y) + A.m[2][2] * v.z);
}
inline V3 tmul(const M3& A, const V3& v) {  // A^T * v
  return V3((A.m[0][0] * v.x + A.m[1][0] * v.y) + A.m[2][0] * v.z,
            (A.m[0][1] * v.x + A.m[1][1] * v.y) + A.m[2][1] * v.z,
            (A.m[0][2] * v.x + A.m[1][2] * v.y) + A.m[2][2] * v.z);
}
inline M3 tmul(const M3& A, const M3& B) {  // A^T * B
  M3 C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C.m[i][j] = (A.m[0][i] * B.m[0][j] + A.m[1][i] * B.m[1][j]) + A.m[2][i] * B.m[2][j];
  return C;
}
inline M3 mul(const M3& A, const M3& B) {  // A * B
  M3 C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C.m[i][j] = (A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j]) + A.m[i][2] * B.m[2][j];
  return C;
}
// Eigen isIdentity(prec) (MatrixBase): diagonal isApprox(1), off-diagonal
// isMuchSmallerThan(1)
inline bool is_identity(const M3& A, double prec = 1e-12) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      if (i == j) {
        double x = A.m[i][j];
        if (!(std::fabs(x - 1.0) <= std::fmin(std::fabs(x), 1.0) * prec)) return false;
      } else if (!(std::fabs(A.m[i][j]) <= prec))
        return false;
    }
  return true;
}

// Transform3f (include/hpp/fcl/math/transform.h:56-216)
struct Tf {
  M3 R;
  V3 T;
  V3 transform(const V3& v) const { return mul(R, v) + T; }      // :152-155
  Tf inverse_times(const Tf& o) const {                          // :176-178
    Tf r;
    r.R = tmul(R, o.R);
    r.T = tmul(R, o.T - T);
    return r;
  }
};
inline Tf tf_from_pod(const hfb_transform& t) {
  Tf r;
  for (int c = 0; c < 3; ++c)
    for (int k = 0; k < 3; ++k) r.R.m[k][c] = t.R[c * 3 + k];  // column-major
  r.T = V3(t.T[0], t.T[1], t.T[2]);
  return r;
}

// ---------------------------------------------------------------- shapes ---
// ConvexBase (include/hpp/fcl/shape/geometric_shapes.h:638-872)
struct Convex {
  std::vector<V3> points;
  std::vector<std::vector<unsigned>> neighbors;  // ascending ids (std::set order, convex.hxx:231-280)
  std::vector<V3> warm_points;                   // support_warm_starts (gjk.cpp:1470-1534)
  std::vector<int> warm_indices;
  V3 aabb_center;
  bool has_neighbors() const { return !neighbors.empty(); }
};

struct Shape {
  int type;       // HFB_GEOM_*
  double p[3];    // see hfb_shape
  double ssr;     // swept sphere radius
  const Convex* cvx;  // CONVEX: vertex set; TRIANGLE: cvx->points[0..2] = a,b,c
  V3 tri[3];      // TRIANGLE by value (used for transformed copies)
  double d;       // PLANE / HALFSPACE: offset (p = unit normal), geometric_shapes.h:885-1031
  Shape() : type(0), p{0, 0, 0}, ssr(0), cvx(nullptr), d(0) {}
};

// ShapeSupportData (include/hpp/fcl/narrowphase/support_functions.h:75-94)
struct SupportData {
  std::vector<int8_t> visited;
  V3 last_dir;
};

// ---------------------------------------------------------------- solver ---
struct SimplexV { V3 w0, w1, w; };  // gjk.h:55-61

struct MinkowskiDiff {  // minkowski_difference.h:57-186
  const Shape* shapes[2];
  SupportData data[2];
  M3 oR1;
  V3 ot1;
  double swept_sphere_radius[2];
  bool normalize_support_direction;
  bool identity;
  void set(const Shape* s0, const Shape* s1, const Tf& tf0, const Tf& tf1);
  void set(const Shape* s0, const Shape* s1);
  void support(const V3& dir, V3& supp0, V3& supp1, int hint[2]);
};

struct GJK {
  enum Status { DidNotRun, Failed, NoCollisionEarlyStopped, NoCollision,
                CollisionWithPenetrationInformation, Collision };
  struct Simplex { SimplexV* vertex[4]; unsigned char rank; };
  double distance_upper_bound;
  Status status;
  int gjk_variant, convergence_criterion, convergence_criterion_type;
  MinkowskiDiff* shape;
  V3 ray;
  int support_hint[2];
  double distance;
  Simplex* simplex;
  size_t max_iterations;
  double tolerance;
  SimplexV store_v[4];
  SimplexV* free_v[4];
  unsigned char nfree, current;
  Simplex simplices[2];
  size_t iterations, iterations_momentum_stop;

  GJK(size_t max_it, double tol);
  void reset(size_t max_it, double tol);
  Status evaluate(MinkowskiDiff& shape, const V3& guess, const int hint[2]);
  void getSupport(const V3& d, SimplexV& sv, int hint[2]) const;
  bool encloseOrigin();
  void getWitnessPointsAndNormal(const MinkowskiDiff& shape, V3& w0, V3& w1, V3& normal) const;
  bool checkConvergence(const V3& w, const double& rl, double& alpha, const double& omega) const;
  void removeVertex(Simplex& s);
  void appendVertex(Simplex& s, const V3& v, int hint[2]);
  bool projectLineOrigin(const Simplex& current, Simplex& next);
  bool projectTriangleOrigin(const Simplex& current, Simplex& next);
  bool projectTetrahedraOrigin(const Simplex& current, Simplex& next);
};

struct EPA {
  enum Status { DidNotRun = -1, Failed = 0, Valid = 1, AccuracyReached = 1 << 1 | Valid,
                Degenerated = 1 << 1 | Failed, NonConvex = 2 << 1 | Failed,
                InvalidHull = 3 << 1 | Failed, OutOfFaces = 4 << 1 | Failed,
                OutOfVertices = 5 << 1 | Failed, FallBack = 6 << 1 | Failed };
  struct Face {
    V3 n; double d; bool ignore; size_t vertex_id[3];
    Face* adjacent_faces[3]; Face* prev_face; Face* next_face;
    size_t adjacent_edge[3]; size_t pass;
    Face() : d(0), ignore(false) {}
  };
  struct FaceList {
    Face* root; size_t count;
    FaceList() : root(nullptr), count(0) {}
    void reset() { root = nullptr; count = 0; }
    void append(Face* f);
    void remove(Face* f);
  };
  struct Horizon { Face* current_face; Face* first_face; size_t num_faces;
                   Horizon() : current_face(nullptr), first_face(nullptr), num_faces(0) {} };
  Status status;
  GJK::Simplex result;
  V3 normal;
  int support_hint[2];
  double depth;
  Face* closest_face;
  size_t max_iterations;
  double tolerance;
  std::vector<SimplexV> sv_store;
  std::vector<Face> fc_store;
  FaceList hull, stock;
  size_t num_vertices, iterations;

  EPA(size_t max_it, double tol);
  void reset(size_t max_it, double tol);
  Status evaluate(GJK& gjk, const V3& guess);
  void getWitnessPointsAndNormal(const MinkowskiDiff& shape, V3& w0, V3& w1, V3& normal) const;
  Face* newFace(size_t a, size_t b, size_t c, bool force = false);
  Face* findClosestFace();
  bool expand(size_t pass, const SimplexV& w, Face* f, size_t e, Horizon& horizon);
};

// GJKSolver (include/hpp/fcl/narrowphase/narrowphase.h:58-724)
struct GJKSolver {
  GJK gjk;
  EPA epa;
  size_t gjk_max_iterations; double gjk_tolerance; int gjk_initial_guess;
  V3 cached_guess; int support_func_cached_guess[2];
  double distance_upper_bound;
  int gjk_variant, gjk_convergence_criterion, gjk_convergence_criterion_type;
  size_t epa_max_iterations; double epa_tolerance;
  MinkowskiDiff minkowski_difference;

  explicit GJKSolver(const hfb_distance_request& r);
  explicit GJKSolver(const hfb_collision_request& r);
  void set_query(const hfb_query_request& q);
  // shapeDistance incl. the TriangleP partial specialisations (:308-348)
  double shapeDistance(const Shape& s1, const Tf& tf1, const Shape& s2, const Tf& tf2,
                       bool compute_penetration, V3& p1, V3& p2, V3& normal);
  void runGJKAndEPA(const Shape& s1, const Tf& tf1, const Shape& s2, const Tf& tf2,
                    bool compute_penetration, double& distance, V3& p1, V3& p2, V3& normal,
                    bool relative_transformation_already_computed);
};

// per-shape support (support_functions.cpp:111-437), NoSweptSphere option
void getShapeSupport(const Shape* s, const V3& dir, V3& support, int& hint, SupportData& data);
void buildSupportWarmStart(Convex& c);       // gjk.cpp:1470-1534
void fillNeighborsFromTriangles(Convex& c, const uint32_t* tris, size_t ntris);  // convex.hxx:231-280

// internal::ShapeShapeDistance<S1,S2> dispatch (shape_shape_func.h:185-306):
// closed-form specialisations or GJKSolver::shapeDistance.
// returns false when the pair type is unsupported.
bool shapeShapeDistance(const Shape& s1, const Tf& tf1, const Shape& s2, const Tf& tf2,
                        GJKSolver& solver, bool compute_signed_distance, double& distance,
                        V3& p1, V3& p2, V3& normal, bool& closed_form);

// Project (src/intersect.cpp:570-705)
struct ProjectResult { double parameterization[4]; double sqr_distance; unsigned encode;
                       ProjectResult() : parameterization{0, 0, 0, 0}, sqr_distance(-1), encode(0) {} };
extern std::atomic<unsigned long long> g_undefined_projections;  // see gjk_epa.cpp
ProjectResult projectLineOrigin(const V3& a, const V3& b);
ProjectResult projectTriangleOrigin(const V3& a, const V3& b, const V3& c);
ProjectResult projectTetrahedraOrigin(const V3& a, const V3& b, const V3& c, const V3& d);

}  // namespace oracle
