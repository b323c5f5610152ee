// C++ host-side mirror of the hpp-fcl query interface for the accelerated hot
// path, above the C-ABI of include/hppfcl_b200.h.  Header-only, C++17, no Eigen.
//
// Same names, argument meaning and error behaviour as the reference:
//   collide(o1, tf1, o2, tf2, CollisionRequest, CollisionResult&)   include/hpp/fcl/collision.h:65-70
//   distance(o1, tf1, o2, tf2, DistanceRequest, DistanceResult&)    include/hpp/fcl/distance.h:60-65
//   ComputeCollision / ComputeDistance functors                     collision.h:79-117, distance.h:74-112
//   CollisionRequest/Result, DistanceRequest/Result, Contact        include/hpp/fcl/collision_data.h
//   shapes (Box, Sphere, Capsule, Cone, Cylinder, Ellipsoid,
//           ConvexBase, TriangleP)                                  include/hpp/fcl/shape/geometric_shapes.h
// plus the batched form (`BatchNarrowPhase`) a broadphase pair collector
// (CollisionCallBackCollect, broadphase/default_broadphase_callbacks.h:224-252)
// feeds.  Every query runs on the GPU; unsupported pairs and invalid requests
// throw std::invalid_argument like HPP_FCL_THROW_PRETTY does (fwd.hh:63-72).
#pragma once
#include <array>
#include <algorithm>
#include <cmath>
#include <limits>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/hppfcl_b200.h"

namespace hppfcl_b200 {

typedef double FCL_REAL;

struct Vec3f {
  FCL_REAL v[3];
  Vec3f() : v{0, 0, 0} {}
  Vec3f(FCL_REAL x, FCL_REAL y, FCL_REAL z) : v{x, y, z} {}
  FCL_REAL& operator[](int i) { return v[i]; }
  FCL_REAL operator[](int i) const { return v[i]; }
  static Vec3f Constant(FCL_REAL c) { return Vec3f(c, c, c); }
  FCL_REAL norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
};
inline Vec3f operator+(const Vec3f& a, const Vec3f& b) { return Vec3f(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline Vec3f operator-(const Vec3f& a, const Vec3f& b) { return Vec3f(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline Vec3f operator*(FCL_REAL s, const Vec3f& a) { return Vec3f(s * a[0], s * a[1], s * a[2]); }

// column-major 3x3 like Eigen::Matrix3d
struct Matrix3f {
  FCL_REAL m[9];
  Matrix3f() : m{1, 0, 0, 0, 1, 0, 0, 0, 1} {}
  FCL_REAL& operator()(int r, int c) { return m[c * 3 + r]; }
  FCL_REAL operator()(int r, int c) const { return m[c * 3 + r]; }
  static Matrix3f Identity() { return Matrix3f(); }
};
inline Vec3f operator*(const Matrix3f& R, const Vec3f& x) {
  return Vec3f(R(0, 0) * x[0] + R(0, 1) * x[1] + R(0, 2) * x[2], R(1, 0) * x[0] + R(1, 1) * x[1] + R(1, 2) * x[2],
               R(2, 0) * x[0] + R(2, 1) * x[1] + R(2, 2) * x[2]);
}
inline Matrix3f operator*(const Matrix3f& A, const Matrix3f& B) {
  Matrix3f C;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C(i, j) = A(i, 0) * B(0, j) + A(i, 1) * B(1, j) + A(i, 2) * B(2, j);
  return C;
}

struct Quatf {
  FCL_REAL w, x, y, z;
  Quatf(FCL_REAL w_, FCL_REAL x_, FCL_REAL y_, FCL_REAL z_) : w(w_), x(x_), y(y_), z(z_) {}
  Matrix3f toRotationMatrix() const {
    Matrix3f R;
    R(0, 0) = 1 - 2 * (y * y + z * z); R(0, 1) = 2 * (x * y - z * w); R(0, 2) = 2 * (x * z + y * w);
    R(1, 0) = 2 * (x * y + z * w); R(1, 1) = 1 - 2 * (x * x + z * z); R(1, 2) = 2 * (y * z - x * w);
    R(2, 0) = 2 * (x * z - y * w); R(2, 1) = 2 * (y * z + x * w); R(2, 2) = 1 - 2 * (x * x + y * y);
    return R;
  }
};
inline Quatf makeQuat(FCL_REAL w, FCL_REAL x, FCL_REAL y, FCL_REAL z) { return Quatf(w, x, y, z); }

// math/transform.h:56-216
class Transform3f {
  Matrix3f R;
  Vec3f T;

 public:
  Transform3f() {}
  Transform3f(const Matrix3f& R_, const Vec3f& T_) : R(R_), T(T_) {}
  Transform3f(const Quatf& q, const Vec3f& T_) : R(q.toRotationMatrix()), T(T_) {}
  explicit Transform3f(const Vec3f& T_) : T(T_) {}
  const Vec3f& getTranslation() const { return T; }
  const Matrix3f& getRotation() const { return R; }
  void setTranslation(const Vec3f& t) { T = t; }
  void setRotation(const Matrix3f& r) { R = r; }
  void setQuatRotation(const Quatf& q) { R = q.toRotationMatrix(); }
  Vec3f transform(const Vec3f& v) const { return R * v + T; }
  Transform3f operator*(const Transform3f& o) const { return Transform3f(R * o.R, R * o.T + T); }
  hfb_transform pod() const {
    hfb_transform t;
    for (int i = 0; i < 9; ++i) t.R[i] = R.m[i];
    for (int i = 0; i < 3; ++i) t.T[i] = T[i];
    return t;
  }
};

// ------------------------------------------------------------------ geometry --
enum NODE_TYPE {  // collision_object.h:65-89 (subset)
  GEOM_BOX = HFB_GEOM_BOX, GEOM_SPHERE = HFB_GEOM_SPHERE, GEOM_CAPSULE = HFB_GEOM_CAPSULE,
  GEOM_CONE = HFB_GEOM_CONE, GEOM_CYLINDER = HFB_GEOM_CYLINDER, GEOM_CONVEX = HFB_GEOM_CONVEX,
  GEOM_PLANE = HFB_GEOM_PLANE, GEOM_HALFSPACE = HFB_GEOM_HALFSPACE, GEOM_TRIANGLE = HFB_GEOM_TRIANGLE,
  GEOM_ELLIPSOID = HFB_GEOM_ELLIPSOID, BV_OBB = HFB_BV_OBB, BV_OBBRSS = HFB_BV_OBBRSS
};

class CollisionGeometry {
 public:
  virtual ~CollisionGeometry() {}
  virtual NODE_TYPE getNodeType() const = 0;
  // flattening hooks used by the arena
  virtual void params(double p[3]) const { p[0] = p[1] = p[2] = 0; }
  virtual const std::vector<Vec3f>* vertices() const { return nullptr; }
  // triangle meshes: flattened tree + geometry (null for shapes)
  struct MeshData {
    const std::vector<hfb_bvh_node>* nodes;
    const std::vector<double>* verts;
    const std::vector<uint32_t>* tris;
  };
  virtual bool mesh(MeshData&) const { return false; }
  // Plane / Halfspace: unit normal and offset (the 40-byte record has no room for them)
  virtual bool plane(double /*n_d*/[4]) const { return false; }
};

class ShapeBase : public CollisionGeometry {  // geometric_shapes.h:59-102
 public:
  void setSweptSphereRadius(FCL_REAL radius) {
    if (radius < 0) throw std::invalid_argument("Swept-sphere radius must be positive.");
    m_swept_sphere_radius = radius;
  }
  FCL_REAL getSweptSphereRadius() const { return m_swept_sphere_radius; }

 protected:
  FCL_REAL m_swept_sphere_radius{0};
};

class Box : public ShapeBase {  // :164-236
 public:
  Box(FCL_REAL x, FCL_REAL y, FCL_REAL z) : halfSide(x / 2, y / 2, z / 2) {}
  Vec3f halfSide;
  NODE_TYPE getNodeType() const override { return GEOM_BOX; }
  void params(double p[3]) const override { p[0] = halfSide[0]; p[1] = halfSide[1]; p[2] = halfSide[2]; }
};
class Sphere : public ShapeBase {  // :238-301
 public:
  explicit Sphere(FCL_REAL r) : radius(r) {}
  FCL_REAL radius;
  NODE_TYPE getNodeType() const override { return GEOM_SPHERE; }
  void params(double p[3]) const override { p[0] = radius; p[1] = p[2] = 0; }
};
class Ellipsoid : public ShapeBase {  // :303-379
 public:
  Ellipsoid(FCL_REAL rx, FCL_REAL ry, FCL_REAL rz) : radii(rx, ry, rz) {}
  Vec3f radii;
  NODE_TYPE getNodeType() const override { return GEOM_ELLIPSOID; }
  void params(double p[3]) const override { p[0] = radii[0]; p[1] = radii[1]; p[2] = radii[2]; }
};
#define HFB_RADIUS_LENGTH_SHAPE(Name, Type)                                         \
  class Name : public ShapeBase {                                                   \
   public:                                                                          \
    Name(FCL_REAL radius_, FCL_REAL lz_) : radius(radius_), halfLength(lz_ / 2) {}  \
    FCL_REAL radius, halfLength;                                                    \
    NODE_TYPE getNodeType() const override { return Type; }                         \
    void params(double p[3]) const override { p[0] = radius; p[1] = halfLength; p[2] = 0; } \
  };
HFB_RADIUS_LENGTH_SHAPE(Capsule, GEOM_CAPSULE)    // :381-463
HFB_RADIUS_LENGTH_SHAPE(Cone, GEOM_CONE)          // :465-553
HFB_RADIUS_LENGTH_SHAPE(Cylinder, GEOM_CYLINDER)  // :555-634
#undef HFB_RADIUS_LENGTH_SHAPE

// Halfspace n . x <= d (:885-961) and Plane n . x = d (:977-1050); the constructors normalise (n, d) like the
// reference's unitNormalTest (src/shape/geometric_shapes.cpp:121-143)
#define HFB_PLANE_SHAPE(Name, Type)                                                        \
  class Name : public ShapeBase {                                                          \
   public:                                                                                 \
    Name(const Vec3f& n_, FCL_REAL d_) : n(n_), d(d_) { unitNormalTest(); }                \
    Name(FCL_REAL a, FCL_REAL b, FCL_REAL c, FCL_REAL d_) : n(a, b, c), d(d_) { unitNormalTest(); } \
    Name() : n(1, 0, 0), d(0) {}                                                           \
    Vec3f n;                                                                               \
    FCL_REAL d;                                                                            \
    NODE_TYPE getNodeType() const override { return Type; }                                \
    void params(double p[3]) const override { p[0] = n[0]; p[1] = n[1]; p[2] = n[2]; }     \
    bool plane(double nd[4]) const override {                                              \
      nd[0] = n[0]; nd[1] = n[1]; nd[2] = n[2]; nd[3] = d;                                 \
      return true;                                                                         \
    }                                                                                      \
                                                                                           \
   protected:                                                                              \
    void unitNormalTest() {                                                                \
      const FCL_REAL l = std::sqrt((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]);             \
      if (l > 0) {                                                                         \
        const FCL_REAL inv_l = 1.0 / l;                                                    \
        n = Vec3f(n[0] * inv_l, n[1] * inv_l, n[2] * inv_l);                               \
        d *= inv_l;                                                                        \
      } else {                                                                             \
        n = Vec3f(1, 0, 0);                                                                \
        d = 0;                                                                             \
      }                                                                                    \
    }                                                                                      \
  };
HFB_PLANE_SHAPE(Halfspace, GEOM_HALFSPACE)
HFB_PLANE_SHAPE(Plane, GEOM_PLANE)
#undef HFB_PLANE_SHAPE

class ConvexBase : public ShapeBase {  // :638-872 (vertex set; faces/neighbours are not needed on the device)
 public:
  explicit ConvexBase(std::vector<Vec3f> pts) : points(std::make_shared<std::vector<Vec3f>>(std::move(pts))) {
    num_points = (unsigned)points->size();
  }
  std::shared_ptr<std::vector<Vec3f>> points;
  unsigned num_points;
  NODE_TYPE getNodeType() const override { return GEOM_CONVEX; }
  const std::vector<Vec3f>* vertices() const override { return points.get(); }
};
class TriangleP : public ShapeBase {  // :109-162
 public:
  TriangleP(const Vec3f& a_, const Vec3f& b_, const Vec3f& c_) : a(a_), b(b_), c(c_), abc{a_, b_, c_} {}
  Vec3f a, b, c;
  NODE_TYPE getNodeType() const override { return GEOM_TRIANGLE; }
  const std::vector<Vec3f>* vertices() const override { return &abc; }

 private:
  std::vector<Vec3f> abc;
};
// BVHModel<OBBRSS> (include/hpp/fcl/BVH/BVH_model.h): the build protocol of the reference
// (beginModel / addVertex / addTriangle / addSubModel / endModel), triangles only, SPLIT_METHOD_MEAN.
// endModel() builds the tree on the host (hfb_bvh_build_obbrss: bit-identical to the reference's
// BVHModel::bvs); queries walk it on the GPU.  BVHModel<OBB> (collide() only, as in the library) shares the builder:
// the plain OBB tree of a mesh is the OBB half of its OBBRSS tree (oracle/ref_driver.cpp: ref_bvh_obb_is_obbrss_half).
// Other BV types are not provided.
struct OBBRSS {};
struct OBB {};
struct Triangle {  // data_types.h
  size_t vids[3];
  Triangle() : vids{0, 0, 0} {}
  Triangle(size_t a, size_t b, size_t c) : vids{a, b, c} {}
  size_t operator[](int i) const { return vids[i]; }
};
enum BVHBuildState { BVH_BUILD_STATE_EMPTY, BVH_BUILD_STATE_BEGUN, BVH_BUILD_STATE_PROCESSED };
enum BVHReturnCode { BVH_OK = 0, BVH_ERR_BUILD_OUT_OF_SEQUENCE = -2, BVH_ERR_BUILD_EMPTY_MODEL = -3 };
class BVHModelTriangles : public CollisionGeometry {  // what the two models share (BVHModelBase + the tree)
 public:
  unsigned num_tris = 0, num_vertices = 0;
  BVHBuildState build_state = BVH_BUILD_STATE_EMPTY;
  int beginModel(unsigned num_tris_ = 0, unsigned num_vertices_ = 0) {  // BVH_model.cpp:226-263
    if (build_state != BVH_BUILD_STATE_EMPTY) {
      verts_.clear();
      tris_.clear();
      nodes_.clear();
      num_tris = num_vertices = 0;
    }
    tris_.reserve(3 * (size_t)num_tris_);
    verts_.reserve(3 * (size_t)num_vertices_);
    build_state = BVH_BUILD_STATE_BEGUN;
    return BVH_OK;
  }
  int addVertex(const Vec3f& p) {
    if (build_state != BVH_BUILD_STATE_BEGUN) return BVH_ERR_BUILD_OUT_OF_SEQUENCE;
    for (int k = 0; k < 3; ++k) verts_.push_back(p[k]);
    ++num_vertices;
    return BVH_OK;
  }
  int addTriangle(const Vec3f& p1, const Vec3f& p2, const Vec3f& p3) {  // :359-414: three new vertices
    if (build_state != BVH_BUILD_STATE_BEGUN) return BVH_ERR_BUILD_OUT_OF_SEQUENCE;
    const uint32_t offset = num_vertices;
    addVertex(p1);
    addVertex(p2);
    addVertex(p3);
    tris_.push_back(offset);
    tris_.push_back(offset + 1);
    tris_.push_back(offset + 2);
    ++num_tris;
    return BVH_OK;
  }
  int addSubModel(const std::vector<Vec3f>& ps, const std::vector<Triangle>& ts) {  // :470-538
    if (build_state != BVH_BUILD_STATE_BEGUN) return BVH_ERR_BUILD_OUT_OF_SEQUENCE;
    const uint32_t offset = num_vertices;
    for (const Vec3f& p : ps) addVertex(p);
    for (const Triangle& t : ts) {
      for (int k = 0; k < 3; ++k) tris_.push_back((uint32_t)t[k] + offset);
      ++num_tris;
    }
    return BVH_OK;
  }
  int endModel() {  // :540-586 + buildTree
    if (build_state != BVH_BUILD_STATE_BEGUN) return BVH_ERR_BUILD_OUT_OF_SEQUENCE;
    if (num_tris == 0) return BVH_ERR_BUILD_EMPTY_MODEL;
    nodes_.resize(2 * (size_t)num_tris - 1);
    if (hfb_bvh_build_obbrss(verts_.data(), num_vertices, tris_.data(), num_tris, nodes_.data(),
                             (uint32_t)nodes_.size()) != HFB_OK)
      throw std::invalid_argument("BVHModel: invalid triangle indices");
    build_state = BVH_BUILD_STATE_PROCESSED;
    return BVH_OK;
  }
  unsigned getNumBVs() const { return (unsigned)nodes_.size(); }
  const hfb_bvh_node& getBV(unsigned i) const { return nodes_[i]; }
  bool mesh(MeshData& m) const override {
    if (build_state != BVH_BUILD_STATE_PROCESSED) throw std::invalid_argument("BVHModel: endModel() was not called");
    m.nodes = &nodes_;
    m.verts = &verts_;
    m.tris = &tris_;
    return true;
  }

 private:
  std::vector<double> verts_;
  std::vector<uint32_t> tris_;
  std::vector<hfb_bvh_node> nodes_;
};
template <typename BV>
class BVHModel;
template <>
class BVHModel<OBBRSS> : public BVHModelTriangles {
 public:
  NODE_TYPE getNodeType() const override { return BV_OBBRSS; }
};
template <>
class BVHModel<OBB> : public BVHModelTriangles {
 public:
  NODE_TYPE getNodeType() const override { return BV_OBB; }
};
inline bool isMeshType(NODE_TYPE t) { return t == BV_OBBRSS || t == BV_OBB; }
typedef std::shared_ptr<CollisionGeometry> CollisionGeometryPtr_t;

struct AABB {  // BV/AABB.h
  Vec3f min_, max_;
  bool overlap(const AABB& o) const {  // :111-118, closed intervals
    for (int k = 0; k < 3; ++k)
      if (min_[k] > o.max_[k] || max_[k] < o.min_[k]) return false;
    return true;
  }
};

class CollisionObject {  // collision_object.h
 public:
  CollisionObject(const CollisionGeometryPtr_t& g, const Transform3f& tf) : geom(g), t(tf) {}
  const CollisionGeometry* collisionGeometryPtr() const { return geom.get(); }
  const CollisionGeometryPtr_t& collisionGeometry() const { return geom; }
  const Transform3f& getTransform() const { return t; }
  void setTransform(const Transform3f& tf) { t = tf; }
  /// world-space box (:258-278); filled by the broadphase manager's setup() / update() for all its objects at once
  const AABB& getAABB() const { return aabb; }
  AABB& getAABB() { return aabb; }

 private:
  CollisionGeometryPtr_t geom;
  Transform3f t;
  AABB aabb;
};

// -------------------------------------------------------------- requests ------
enum GJKInitialGuess { DefaultGuess, CachedGuess, BoundingVolumeGuess };
enum GJKVariant { DefaultGJK, PolyakAcceleration, NesterovAcceleration };
enum GJKConvergenceCriterion { Default, DualityGap, Hybrid };
enum GJKConvergenceCriterionType { Relative, Absolute };

struct QueryResult;
struct QueryRequest {  // collision_data.h:171-274
  GJKInitialGuess gjk_initial_guess = DefaultGuess;
  GJKVariant gjk_variant = DefaultGJK;
  GJKConvergenceCriterion gjk_convergence_criterion = Default;
  GJKConvergenceCriterionType gjk_convergence_criterion_type = Relative;
  FCL_REAL gjk_tolerance = 1e-6;
  size_t gjk_max_iterations = 128;
  mutable Vec3f cached_gjk_guess = Vec3f(1, 0, 0);
  mutable std::array<int, 2> cached_support_func_guess{{0, 0}};
  size_t epa_max_iterations = 64;
  FCL_REAL epa_tolerance = 1e-6;
  bool enable_timings = false;
  FCL_REAL collision_distance_threshold = 1e-12;
  inline void updateGuess(const QueryResult& result) const;  // :291-300
  void fill(hfb_query_request& q) const {
    q.gjk_initial_guess = gjk_initial_guess;
    q.gjk_variant = gjk_variant;
    q.gjk_convergence_criterion = gjk_convergence_criterion;
    q.gjk_convergence_criterion_type = gjk_convergence_criterion_type;
    q.gjk_max_iterations = (uint32_t)gjk_max_iterations;
    q.epa_max_iterations = (uint32_t)epa_max_iterations;
    q.gjk_tolerance = gjk_tolerance;
    q.epa_tolerance = epa_tolerance;
    q.collision_distance_threshold = collision_distance_threshold;
    q.cached_gjk_guess = nullptr;
    q.cached_support_func_guess = nullptr;
  }
};
struct QueryResult {
  Vec3f cached_gjk_guess = Vec3f(1, 0, 0);
  std::array<int, 2> cached_support_func_guess{{0, 0}};
};
inline void QueryRequest::updateGuess(const QueryResult& result) const {
  if (gjk_initial_guess == CachedGuess) {
    cached_gjk_guess = result.cached_gjk_guess;
    cached_support_func_guess = result.cached_support_func_guess;
  }
}

struct Contact {  // collision_data.h:59-166
  const CollisionGeometry* o1 = nullptr;
  const CollisionGeometry* o2 = nullptr;
  int b1 = -1, b2 = -1;
  Vec3f normal;
  std::array<Vec3f, 2> nearest_points;
  Vec3f pos;
  FCL_REAL penetration_depth = (std::numeric_limits<FCL_REAL>::max)();
  static const int NONE = -1;
};

struct CollisionResult;
struct CollisionRequest : QueryRequest {  // :312-383
  size_t num_max_contacts = 1;
  bool enable_contact = true;
  FCL_REAL security_margin = 0;
  FCL_REAL break_distance = 1e-3;
  FCL_REAL distance_upper_bound = (std::numeric_limits<FCL_REAL>::max)();
  inline bool isSatisfied(const CollisionResult& result) const;
};
struct CollisionResult : QueryResult {  // :391-509
  std::vector<Contact> contacts;
  FCL_REAL distance_lower_bound = (std::numeric_limits<FCL_REAL>::max)();
  Vec3f normal = Vec3f::Constant(std::numeric_limits<FCL_REAL>::quiet_NaN());
  std::array<Vec3f, 2> nearest_points{{normal, normal}};
  bool isCollision() const { return contacts.size() > 0; }
  size_t numContacts() const { return contacts.size(); }
  const Contact& getContact(size_t i) const {
    if (contacts.size() == 0) throw std::invalid_argument("The number of contacts is zero. No Contact can be returned.");
    return i < contacts.size() ? contacts[i] : contacts.back();
  }
  void addContact(const Contact& c) { contacts.push_back(c); }
  void clear() { *this = CollisionResult(); }
};
inline bool CollisionRequest::isSatisfied(const CollisionResult& r) const {  // collision_data.cpp:44-46
  return r.isCollision() && (num_max_contacts <= r.numContacts());
}

struct DistanceResult;
struct DistanceRequest : QueryRequest {  // :987-1050
  bool enable_nearest_points = true;
  bool enable_signed_distance = true;
  FCL_REAL rel_err = 0, abs_err = 0;
  DistanceRequest(bool enable_nearest_points_ = true, bool enable_signed_distance_ = true, FCL_REAL rel_err_ = 0,
                  FCL_REAL abs_err_ = 0)
      : enable_nearest_points(enable_nearest_points_), enable_signed_distance(enable_signed_distance_),
        rel_err(rel_err_), abs_err(abs_err_) {}
  inline bool isSatisfied(const DistanceResult& result) const;
};
struct DistanceResult : QueryResult {  // :1053-1174
  FCL_REAL min_distance = (std::numeric_limits<FCL_REAL>::max)();
  Vec3f normal = Vec3f::Constant(std::numeric_limits<FCL_REAL>::quiet_NaN());
  std::array<Vec3f, 2> nearest_points{{normal, normal}};
  const CollisionGeometry* o1 = nullptr;
  const CollisionGeometry* o2 = nullptr;
  int b1 = -1, b2 = -1;
  static const int NONE = -1;
  void clear() { *this = DistanceResult(); }
};
inline bool DistanceRequest::isSatisfied(const DistanceResult& r) const { return r.min_distance <= 0; }

// ------------------------------------------------------------- device arena ----
// One context per device; geometry handle cache keyed by object address (the
// caller keeps ownership, like the reference; re-registration on parameter change
// is the caller's job via Context::invalidate).
class Context {
 public:
  static Context& instance(int device = 0) {
    static std::mutex mu;
    static std::unordered_map<int, std::unique_ptr<Context>> all;
    std::lock_guard<std::mutex> lk(mu);
    auto& c = all[device];
    if (!c) c.reset(new Context(device));
    return *c;
  }
  ~Context() { hfb_ctx_destroy(ctx); }
  hfb_ctx* raw() { return ctx; }
  uint32_t handle(const CollisionGeometry* g) {
    CollisionGeometry::MeshData md;
    if (g->mesh(md)) return mesh_handle(g, md);
    double nd[4];
    if (g->plane(nd)) return plane_handle(g, nd);
    hfb_shape rec;
    rec.type = (uint32_t)g->getNodeType();
    rec.data = 0;
    g->params(rec.p);
    const ShapeBase* sb = dynamic_cast<const ShapeBase*>(g);
    rec.ssr = sb ? sb->getSweptSphereRadius() : 0.0;
    const std::vector<Vec3f>* verts = g->vertices();
    // the cache is keyed by address, so an entry is only valid while the flattened record still
    // matches (a freed geometry's address can be reused by a different shape)
    auto it = handles.find(g);
    if (it != handles.end()) {
      const Entry& e = it->second;
      bool same = e.rec.type == rec.type && e.rec.p[0] == rec.p[0] && e.rec.p[1] == rec.p[1] &&
                  e.rec.p[2] == rec.p[2] && e.rec.ssr == rec.ssr && e.nverts == (verts ? verts->size() : 0);
      if (same && verts)
        for (size_t i = 0; i < verts->size() && same; ++i)
          for (int k = 0; k < 3; ++k) same = same && e.verts[3 * i + k] == (*verts)[i][k];
      if (same) return e.handle;
      // changed in place: the handle stays when the arena can absorb the change (same node type, same number of
      // vertices) -- record and vertex set are updated; otherwise it is retired and a new one registered
      if (e.rec.type == rec.type && e.nverts == (verts ? verts->size() : 0)) {
        Entry& m = it->second;
        rec.data = m.rec.data;
        if (verts) {
          for (size_t i = 0; i < verts->size(); ++i)
            for (int k = 0; k < 3; ++k) m.verts[3 * i + k] = (*verts)[i][k];
          check(hfb_geom_update_convex(ctx, rec.data, m.verts.data(), (uint32_t)verts->size()));
        }
        check(hfb_geom_update_shapes(ctx, &m.handle, &rec, 1));
        m.rec = rec;
        dirty = true;
        return m.handle;
      }
      check(hfb_geom_release_shapes(ctx, &e.handle, 1));
      handles.erase(it);
    }
    Entry ent;
    if (const std::vector<Vec3f>* v = verts) {
      std::vector<double>& flat = ent.verts;
      flat.resize(3 * v->size());
      for (size_t i = 0; i < v->size(); ++i)
        for (int k = 0; k < 3; ++k) flat[3 * i + k] = (*v)[i][k];
      check(hfb_geom_register_convex(ctx, flat.data(), (uint32_t)v->size(), &rec.data));
      ent.nverts = v->size();
    }
    uint32_t h;
    check(hfb_geom_register_shapes(ctx, &rec, 1, &h));
    ent.rec = rec;
    ent.handle = h;
    handles[g] = ent;
    dirty = true;
    return h;
  }
  // Plane / Halfspace: (n, d) live outside the record, so one changed in place is registered anew
  uint32_t plane_handle(const CollisionGeometry* g, const double nd[4]) {
    const uint32_t type = (uint32_t)g->getNodeType();
    const double ssr = static_cast<const ShapeBase*>(g)->getSweptSphereRadius();
    auto it = handles.find(g);
    if (it != handles.end()) {
      const Entry& e = it->second;
      if (e.rec.type == type && e.rec.ssr == ssr && e.verts.size() == 4 && e.verts[0] == nd[0] && e.verts[1] == nd[1] &&
          e.verts[2] == nd[2] && e.verts[3] == nd[3])
        return e.handle;
      check(hfb_geom_release_shapes(ctx, &e.handle, 1));
      handles.erase(it);
    }
    Entry ent;
    ent.rec = hfb_shape{};
    ent.rec.type = type;
    ent.rec.ssr = ssr;
    ent.verts.assign(nd, nd + 4);
    uint32_t h;
    check(hfb_geom_register_halfspaces(ctx, type, nd, &ssr, 1, &h));
    ent.handle = h;
    handles[g] = ent;
    dirty = true;
    return h;
  }
  uint32_t mesh_handle(const CollisionGeometry* g, const CollisionGeometry::MeshData& md) {
    auto it = handles.find(g);
    if (it != handles.end()) {
      const Entry& e = it->second;
      if (e.rec.type == (uint32_t)g->getNodeType() && e.verts == *md.verts && e.tris == *md.tris) return e.handle;
      check(hfb_geom_release_shapes(ctx, &e.handle, 1));  // a rebuilt model: the old handle is retired
      handles.erase(it);
    }
    Entry ent;
    ent.rec = hfb_shape{};
    ent.rec.type = (uint32_t)g->getNodeType();
    check((ent.rec.type == HFB_BV_OBB ? hfb_geom_register_bvh_obb : hfb_geom_register_bvh_obbrss)(
        ctx, md.nodes->data(), (uint32_t)md.nodes->size(), md.verts->data(), (uint32_t)(md.verts->size() / 3), md.tris->data(),
        (uint32_t)(md.tris->size() / 3), &ent.rec.data));
    uint32_t h;
    check(hfb_geom_register_shapes(ctx, &ent.rec, 1, &h));
    ent.handle = h;
    ent.verts = *md.verts;
    ent.tris = *md.tris;
    handles[g] = ent;
    dirty = true;
    return h;
  }
  void invalidate(const CollisionGeometry* g) { handles.erase(g); }
  void commit() {
    if (dirty) check(hfb_geom_commit(ctx));
    dirty = false;
  }
  void check(int rc) {
    if (rc == HFB_OK) return;
    const std::string msg = hfb_last_error(ctx);
    if (rc == HFB_ERR_INVALID_ARGUMENT || rc == HFB_ERR_UNSUPPORTED_PAIR) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
  }

 private:
  explicit Context(int device) {
    const int rc = hfb_ctx_create(device, &ctx);
    if (rc != HFB_OK) throw std::runtime_error("hppfcl_b200: no usable CUDA device (there is no CPU fallback)");
  }
  struct Entry {
    hfb_shape rec{};
    uint32_t handle = 0;
    size_t nverts = 0;
    std::vector<double> verts;
    std::vector<uint32_t> tris;
  };
  hfb_ctx* ctx = nullptr;
  std::unordered_map<const CollisionGeometry*, Entry> handles;
  bool dirty = false;
};

namespace detail {
inline Vec3f v3(const double* p) { return Vec3f(p[0], p[1], p[2]); }
inline void fill(const CollisionRequest& r, hfb_collision_request& q) {
  r.fill(q.q);
  q.num_max_contacts = (uint32_t)r.num_max_contacts;
  q.enable_contact = r.enable_contact;
  q.security_margin = r.security_margin;
  q.break_distance = r.break_distance;
  q.distance_upper_bound = r.distance_upper_bound;
}
inline void fill(const DistanceRequest& r, hfb_distance_request& q) {
  r.fill(q.q);
  q.enable_signed_distance = r.enable_signed_distance;
  q.enable_nearest_points = r.enable_nearest_points;
  q.rel_err = r.rel_err;
  q.abs_err = r.abs_err;
}
inline void unsupported(const char* what, const CollisionGeometry* o1, const CollisionGeometry* o2) {
  throw std::invalid_argument(std::string(what) + " function between node type " +
                              std::to_string((int)o1->getNodeType()) + " and node type " +
                              std::to_string((int)o2->getNodeType()) + " is not yet supported.");
}
}  // namespace detail

// ------------------------------------------------------------------ batched ----
struct BatchNarrowPhase {
  explicit BatchNarrowPhase(int device = 0) : C(Context::instance(device)) {}
  void add(const CollisionGeometry* o1, const Transform3f& tf1, const CollisionGeometry* o2, const Transform3f& tf2) {
    h1.push_back(C.handle(o1));
    h2.push_back(C.handle(o2));
    t1.push_back(tf1.pod());
    t2.push_back(tf2.pod());
  }
  size_t size() const { return h1.size(); }
  void clear() { h1.clear(); h2.clear(); t1.clear(); t2.clear(); }
  std::vector<hfb_contact> collide(const CollisionRequest& request) {
    if (request.num_max_contacts == 0 && request.security_margin != -std::numeric_limits<FCL_REAL>::infinity())
      throw std::invalid_argument("Invalid number of max contacts (current value is 0).");
    hfb_collision_request q;
    detail::fill(request, q);
    C.commit();
    std::vector<hfb_contact> out(h1.size());
    C.check(hfb_batch_collide(C.raw(), h1.size(), h1.data(), t1.data(), h2.data(), t2.data(), &q, out.data(), nullptr));
    return out;
  }
  std::vector<hfb_distance_result> distance(const DistanceRequest& request) {
    hfb_distance_request q;
    detail::fill(request, q);
    C.commit();
    std::vector<hfb_distance_result> out(h1.size());
    C.check(hfb_batch_distance(C.raw(), h1.size(), h1.data(), t1.data(), h2.data(), t2.data(), &q, out.data(), nullptr));
    return out;
  }
  Context& C;
  std::vector<uint32_t> h1, h2;
  std::vector<hfb_transform> t1, t2;
};

// ------------------------------------------------------------ broadphase feed ----
// The seam BASELINE config 5 goes through: a broadphase manager hands candidate pairs to a callback
// (broadphase/broadphase_callbacks.h:54-75); the collecting callback (default_broadphase_callbacks.h:224-252,
// src/broadphase/default_broadphase_callbacks.cpp:43-60... CollisionCallBackCollect) keeps them, and the batch then
// goes to the narrow phase in ONE call (hfb_batch_collide_objects) instead of one collide() per callback.
struct CollisionCallBackBase {
  virtual void init() {}
  virtual bool collide(CollisionObject* o1, CollisionObject* o2) = 0;  // true: stop the broadphase
  virtual bool operator()(CollisionObject* o1, CollisionObject* o2) { return collide(o1, o2); }
  virtual ~CollisionCallBackBase() {}
};
struct CollisionCallBackCollect : CollisionCallBackBase {
  typedef std::pair<CollisionObject*, CollisionObject*> CollisionPair;
  explicit CollisionCallBackCollect(size_t max_size_) : max_size(max_size_) { collision_pairs.resize(max_size); }
  bool collide(CollisionObject* o1, CollisionObject* o2) override {
    collision_pairs.push_back(std::make_pair(o1, o2));
    return false;
  }
  size_t numCollisionPairs() const { return collision_pairs.size(); }
  const std::vector<CollisionPair>& getCollisionPairs() const { return collision_pairs; }
  void init() override { collision_pairs.clear(); }
  bool exist(const CollisionPair& pair) const {
    return std::find(collision_pairs.begin(), collision_pairs.end(), pair) != collision_pairs.end();
  }
  bool exist(CollisionObject* o1, CollisionObject* o2) const { return exist(std::make_pair(o1, o2)); }

 protected:
  std::vector<CollisionPair> collision_pairs;
  size_t max_size;
};

// The manager: BroadPhaseCollisionManager's interface (broadphase/broadphase_collision_manager.h:56-134) for the
// self-collision query of config 5.  setup() / update() compute every object's box (CollisionObject::computeAABB)
// and collide(callback) reports every pair of objects with overlapping boxes exactly once -- the set the reference's
// DynamicAABBTreeCollisionManager reports (broadphase_dynamic_AABB_tree.cpp:336-407,716-721); the order is each
// manager's own.  No tree is kept: a uniform grid is rebuilt per query (hfb_broadphase.cuh), so update() costs a
// re-computation of the boxes and nothing else.
class DynamicAABBTreeCollisionManager {
 public:
  explicit DynamicAABBTreeCollisionManager(int device = 0) : C(Context::instance(device)) {}
  void registerObjects(const std::vector<CollisionObject*>& other_objs) {
    objs.insert(objs.end(), other_objs.begin(), other_objs.end());
    stale = true;
  }
  void registerObject(CollisionObject* obj) {
    objs.push_back(obj);
    stale = true;
  }
  void unregisterObject(CollisionObject* obj) {
    objs.erase(std::remove(objs.begin(), objs.end(), obj), objs.end());
    stale = true;
  }
  void clear() {
    objs.clear();
    stale = true;
  }
  size_t size() const { return objs.size(); }
  bool empty() const { return objs.empty(); }
  void getObjects(std::vector<CollisionObject*>& out) const { out = objs; }
  void setup() { refresh(); }
  void update() { refresh(); }
  /// self collision: callback(o1, o2) for every pair with overlapping boxes until it returns true
  void collide(CollisionCallBackBase* callback) {
    callback->init();
    if (stale) refresh();
    std::vector<uint32_t> first, second;
    pairs(first, second);
    for (size_t k = 0; k < first.size(); ++k)
      if ((*callback)(objs[first[k]], objs[second[k]])) return;
  }
  /// the candidate pairs as object indices (what collide(callback) iterates over)
  void pairs(std::vector<uint32_t>& first, std::vector<uint32_t>& second) {
    if (stale) refresh();
    size_t cap = std::max<size_t>(16 * objs.size(), 1024), cnt = 0;
    for (;;) {
      first.resize(cap);
      second.resize(cap);
      C.check(hfb_broadphase_pairs(objs.size(), boxes.data(), first.data(), second.data(), cap, &cnt));
      if (cnt <= cap) break;
      cap = cnt;
    }
    first.resize(cnt);
    second.resize(cnt);
  }
  /// the object table of the scene in the form hfb_batch_*_objects takes
  const std::vector<uint32_t>& handles() const { return hs; }
  const std::vector<hfb_transform>& transforms() const { return tfs; }
  Context& context() { return C; }

 private:
  void refresh() {
    hs.resize(objs.size());
    tfs.resize(objs.size());
    for (size_t i = 0; i < objs.size(); ++i) {
      hs[i] = C.handle(objs[i]->collisionGeometryPtr());
      tfs[i] = objs[i]->getTransform().pod();
    }
    C.commit();
    boxes.resize(6 * objs.size());
    C.check(hfb_scene_aabbs(C.raw(), objs.size(), hs.data(), tfs.data(), boxes.data()));
    for (size_t i = 0; i < objs.size(); ++i) {
      AABB& b = objs[i]->getAABB();
      b.min_ = Vec3f(boxes[6 * i], boxes[6 * i + 1], boxes[6 * i + 2]);
      b.max_ = Vec3f(boxes[6 * i + 3], boxes[6 * i + 4], boxes[6 * i + 5]);
    }
    stale = false;
  }
  Context& C;
  std::vector<CollisionObject*> objs;
  std::vector<uint32_t> hs;
  std::vector<hfb_transform> tfs;
  std::vector<double> boxes;
  bool stale = true;
};

/// config 5 in one call: the manager's candidate pairs through the batched narrow phase.  Returns, per candidate
/// pair, the record hfb_batch_collide writes (first / second: object indices in registration order).
inline std::vector<hfb_contact> collide(DynamicAABBTreeCollisionManager& manager, const CollisionRequest& request,
                                        std::vector<uint32_t>& first, std::vector<uint32_t>& second) {
  if (request.num_max_contacts == 0 && request.security_margin != -std::numeric_limits<FCL_REAL>::infinity())
    throw std::invalid_argument("Invalid number of max contacts (current value is 0).");
  manager.pairs(first, second);
  hfb_collision_request q;
  detail::fill(request, q);
  hfb_object_pairs sc;
  sc.n_objects = manager.handles().size();
  sc.object_handles = manager.handles().data();
  sc.object_tfs = manager.transforms().data();
  sc.n_pairs = first.size();
  sc.first = first.data();
  sc.second = second.data();
  std::vector<hfb_contact> out(first.size());
  manager.context().check(hfb_batch_collide_objects(manager.context().raw(), &sc, &q, out.data(), nullptr, nullptr));
  return out;
}

// ---------------------------------------------------------------- free functions
// collide(): src/collision.cpp:69-130 (results accumulate; callers clear()).
inline std::size_t collide(const CollisionGeometry* o1, const Transform3f& tf1, const CollisionGeometry* o2,
                           const Transform3f& tf2, const CollisionRequest& request, CollisionResult& result) {
  if (request.security_margin == -std::numeric_limits<FCL_REAL>::infinity()) {
    result.clear();
    return false;
  }
  if (request.num_max_contacts == 0) throw std::invalid_argument("Invalid number of max contacts (current value is 0).");
  if (request.isSatisfied(result)) return result.numContacts();  // shape_shape_func.h:139
  Context& C = Context::instance();
  const uint32_t h1 = C.handle(o1), h2 = C.handle(o2);
  C.commit();
  hfb_collision_request q;
  detail::fill(request, q);
  double gin[3] = {request.cached_gjk_guess[0], request.cached_gjk_guess[1], request.cached_gjk_guess[2]};
  int32_t hin[2] = {request.cached_support_func_guess[0], request.cached_support_func_guess[1]};
  q.q.cached_gjk_guess = gin;
  q.q.cached_support_func_guess = hin;
  double gout[3];
  int32_t hout[2];
  hfb_guess_out go{gout, hout};
  hfb_contact rec;
  const hfb_transform t1 = tf1.pod(), t2 = tf2.pod();
  // a mesh pair can have more than one contact (CollisionResult::contacts): fetch them all when asked to
  std::vector<hfb_contact> more;
  uint32_t count = 0;
  const bool mesh = isMeshType(o1->getNodeType()) || isMeshType(o2->getNodeType());
  if (mesh && request.num_max_contacts > 1) {
    more.resize(request.num_max_contacts - 1);
    C.check(hfb_batch_collide_contacts(C.raw(), 1, &h1, &t1, &h2, &t2, &q, &rec, (uint32_t)more.size(), more.data(),
                                       &count, &go));
    more.resize(count > 1 ? count - 1 : 0);
  } else {
    C.check(hfb_batch_collide(C.raw(), 1, &h1, &t1, &h2, &t2, &q, &rec, &go));
  }
  if (HFB_STATUS_PATH(rec.status) == HFB_PATH_UNSUPPORTED) detail::unsupported("Collision", o1, o2);
  if (rec.distance_lower_bound < result.distance_lower_bound) {  // collision_data.h:1186-1197
    result.distance_lower_bound = rec.distance_lower_bound;
    result.nearest_points = {{detail::v3(rec.p1), detail::v3(rec.p2)}};
    result.normal = detail::v3(rec.normal);
  }
  std::size_t res = 0;
  if (rec.num_contacts && result.numContacts() < request.num_max_contacts) {
    Contact c;
    c.o1 = o1;
    c.o2 = o2;
    c.b1 = rec.b1;
    c.b2 = rec.b2;
    c.normal = detail::v3(rec.normal);
    c.nearest_points = {{detail::v3(rec.p1), detail::v3(rec.p2)}};
    c.pos = detail::v3(rec.pos);
    c.penetration_depth = rec.distance;
    result.addContact(c);
    for (const hfb_contact& x : more) {
      if (!(result.numContacts() < request.num_max_contacts)) break;
      Contact d;
      d.o1 = o1;
      d.o2 = o2;
      d.b1 = x.b1;
      d.b2 = x.b2;
      d.normal = detail::v3(x.normal);
      d.nearest_points = {{detail::v3(x.p1), detail::v3(x.p2)}};
      d.pos = detail::v3(x.pos);
      d.penetration_depth = x.distance;
      result.addContact(d);
    }
    res = result.numContacts();
  }
  result.cached_gjk_guess = detail::v3(gout);  // collision.cpp:125-127
  result.cached_support_func_guess = {{hout[0], hout[1]}};
  request.updateGuess(result);
  return res;
}
inline std::size_t collide(const CollisionObject* o1, const CollisionObject* o2, const CollisionRequest& request,
                           CollisionResult& result) {
  return collide(o1->collisionGeometryPtr(), o1->getTransform(), o2->collisionGeometryPtr(), o2->getTransform(),
                 request, result);
}

// distance(): src/distance.cpp:60-109
inline FCL_REAL distance(const CollisionGeometry* o1, const Transform3f& tf1, const CollisionGeometry* o2,
                         const Transform3f& tf2, const DistanceRequest& request, DistanceResult& result) {
  if (request.isSatisfied(result)) return result.min_distance;  // shape_shape_func.h:57
  Context& C = Context::instance();
  const uint32_t h1 = C.handle(o1), h2 = C.handle(o2);
  C.commit();
  hfb_distance_request q;
  detail::fill(request, q);
  double gin[3] = {request.cached_gjk_guess[0], request.cached_gjk_guess[1], request.cached_gjk_guess[2]};
  int32_t hin[2] = {request.cached_support_func_guess[0], request.cached_support_func_guess[1]};
  q.q.cached_gjk_guess = gin;
  q.q.cached_support_func_guess = hin;
  double gout[3];
  int32_t hout[2];
  hfb_guess_out go{gout, hout};
  hfb_distance_result rec;
  const hfb_transform t1 = tf1.pod(), t2 = tf2.pod();
  C.check(hfb_batch_distance(C.raw(), 1, &h1, &t1, &h2, &t2, &q, &rec, &go));
  if (HFB_STATUS_PATH(rec.status) == HFB_PATH_UNSUPPORTED) detail::unsupported("Distance", o1, o2);
  const bool closed = HFB_STATUS_PATH(rec.status) == HFB_PATH_CLOSED_FORM;
  if (closed || result.min_distance > rec.min_distance) {  // DistanceResult::update, collision_data.h:1111-1124
    result.min_distance = rec.min_distance;
    result.o1 = o1;
    result.o2 = o2;
    result.b1 = rec.b1;
    result.b2 = rec.b2;
    result.nearest_points = {{detail::v3(rec.p1), detail::v3(rec.p2)}};
    result.normal = detail::v3(rec.normal);
  }
  result.cached_gjk_guess = detail::v3(gout);
  result.cached_support_func_guess = {{hout[0], hout[1]}};
  request.updateGuess(result);
  return rec.min_distance;
}
inline FCL_REAL distance(const CollisionObject* o1, const CollisionObject* o2, const DistanceRequest& request,
                         DistanceResult& result) {
  return distance(o1->collisionGeometryPtr(), o1->getTransform(), o2->collisionGeometryPtr(), o2->getTransform(),
                  request, result);
}

// functors (collision.h:79-117, distance.h:74-112): geometry handles are resolved once
class ComputeCollision {
 public:
  ComputeCollision(const CollisionGeometry* o1_, const CollisionGeometry* o2_) : o1(o1_), o2(o2_) {
    Context::instance().handle(o1);
    Context::instance().handle(o2);
  }
  std::size_t operator()(const Transform3f& tf1, const Transform3f& tf2, const CollisionRequest& request,
                         CollisionResult& result) const {
    return collide(o1, tf1, o2, tf2, request, result);
  }

 private:
  const CollisionGeometry* o1;
  const CollisionGeometry* o2;
};
class ComputeDistance {
 public:
  ComputeDistance(const CollisionGeometry* o1_, const CollisionGeometry* o2_) : o1(o1_), o2(o2_) {
    Context::instance().handle(o1);
    Context::instance().handle(o2);
  }
  FCL_REAL operator()(const Transform3f& tf1, const Transform3f& tf2, const DistanceRequest& request,
                      DistanceResult& result) const {
    return distance(o1, tf1, o2, tf2, request, result);
  }

 private:
  const CollisionGeometry* o1;
  const CollisionGeometry* o2;
};

}  // namespace hppfcl_b200
