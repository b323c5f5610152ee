// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.hpp). OBBRSS BVH types.
#pragma once
#include <array>
#include <vector>

#include "oracle.hpp"

namespace oracle {

struct OBB { M3 axes; V3 To; V3 extent; };                      // include/hpp/fcl/BV/OBB.h
struct RSS { M3 axes; V3 Tr; double length[2]; double radius; };  // include/hpp/fcl/BV/RSS.h
struct OBBRSS { OBB obb; RSS rss; };                             // include/hpp/fcl/BV/OBBRSS.h
struct Tri { unsigned v[3]; };
struct BVNode {  // include/hpp/fcl/BV/BV_node.h:52-148
  int first_child = 0;
  unsigned first_primitive = 0, num_primitives = 0;
  OBBRSS bv;
};
struct BVHModel {  // include/hpp/fcl/BVH/BVH_model.h (triangles, OBBRSS, SPLIT_METHOD_MEAN)
  std::vector<V3> vertices;
  std::vector<Tri> tris;
  std::vector<BVNode> bvs;
  std::vector<unsigned> primitive_indices;
  unsigned num_bvs = 0;
  void build();
};

struct BvhQueryResult {
  double distance;
  V3 p1, p2, normal;
  int b1;
  uint64_t num_bv_tests, num_leaf_tests;
  int b2 = -1;
};
struct BvhContact { int b1; V3 p1, p2, normal; double distance; int b2 = -1; };
struct BvhCollideResult {
  std::vector<BvhContact> contacts;
  double distance_lower_bound = std::numeric_limits<double>::max();
  V3 lb_p1 = nan3(), lb_p2 = nan3(), lb_normal = nan3();
  uint64_t num_bv_tests = 0, num_leaf_tests = 0;
};

void computeBV_OBBRSS(const Shape& s, const Tf& tf, OBBRSS& bv);
double rectDistance(const M3& Rab, const V3& Tab, const double a[2], const double b[2]);
double rss_distance(const M3& R0, const V3& T0, const RSS& b1, const RSS& b2);
bool obb_overlap(const M3& R0, const V3& T0, const OBB& b1, const OBB& b2, double security_margin,
                 double break_distance, double& sqrDistLowerBound);
void bvhShapeDistance(const BVHModel& m, const Tf& tf1, const Shape& s, const Tf& tf2, GJKSolver& solver,
                      bool signed_distance, double rel_err, double abs_err, BvhQueryResult& out);
void bvhShapeCollide(const BVHModel& m, const Tf& tf1, const Shape& s, const Tf& tf2, GJKSolver& solver,
                     const hfb_collision_request& req, BvhCollideResult& out, bool plain_obb = false);
void computeBV_OBB(const Shape& s, const Tf& tf, OBB& bv);

// mesh-mesh (BVHModel<OBBRSS> x BVHModel<OBBRSS>)
double sqrTriDistance(const V3 S[3], const V3 T[3], V3& P, V3& Q);
void bvhBvhDistance(const BVHModel& m1, const Tf& tf1, const BVHModel& m2, const Tf& tf2, double rel_err,
                    double abs_err, bool enable_nearest_points, BvhQueryResult& out);
void bvhBvhCollide(const BVHModel& m1, const Tf& tf1, const BVHModel& m2, const Tf& tf2, GJKSolver& solver,
                   const hfb_collision_request& req, BvhCollideResult& out);

}  // namespace oracle
