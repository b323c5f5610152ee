// C++ host-API test: the reference's own box_box_distance / capsule tests, written against
// the hppfcl_b200 mirror of the hpp-fcl interface (compare test/box_box_distance.cpp:62-254,
// test/capsule_capsule.cpp:218-356, test/collision.cpp error paths).  Runs on the GPU box.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "../../hpp-fcl_b200/host/hppfcl_b200.hpp"

using namespace hppfcl_b200;

static int failures = 0;
#define CHECK(c)                                                          \
  do {                                                                    \
    if (!(c)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #c); ++failures; } \
  } while (0)
#define CHECK_CLOSE(a, b, pct) CHECK(std::fabs((a) - (b)) <= (pct) / 100.0 * std::fmin(std::fabs(a), std::fabs(b)))

int main() {
  {  // BVHModel<OBB>: the same builder, collide() only (collision_func_matrix.cpp:141-166, 488-501)
    BVHModel<OBB> mo;
    BVHModel<OBBRSS> mr;
    const Vec3f v[4] = {Vec3f(0, 0, 0), Vec3f(1, 0, 0), Vec3f(0, 1, 0), Vec3f(0, 0, 1)};
    const int f[4][3] = {{0, 2, 1}, {0, 1, 3}, {0, 3, 2}, {1, 2, 3}};
    for (BVHModelTriangles* m : {static_cast<BVHModelTriangles*>(&mo), static_cast<BVHModelTriangles*>(&mr)}) {
      m->beginModel();
      for (int k = 0; k < 4; ++k) m->addTriangle(v[f[k][0]], v[f[k][1]], v[f[k][2]]);
      CHECK(m->endModel() == BVH_OK);
    }
    CHECK(mo.getNodeType() == BV_OBB && mo.getNumBVs() == 7);
    Sphere ball(0.2);
    CollisionRequest creq;
    CollisionResult ra, rb;
    const Transform3f tfs(Vec3f(0.3, 0.3, 0.45));
    CHECK(collide(&mo, Transform3f(), &ball, tfs, creq, ra) == collide(&mr, Transform3f(), &ball, tfs, creq, rb));
    CHECK(ra.numContacts() == 1 && ra.getContact(0).b1 == rb.getContact(0).b1);
    CollisionResult rc;
    CHECK(collide(&mo, Transform3f(), &ball, Transform3f(Vec3f(3, 3, 3)), creq, rc) == 0);
  }
  {  // a Halfspace floor and a Plane against a sphere (details::halfspaceDistance / planeDistance, details.h:347-428)
    Halfspace floor(Vec3f(0, 0, 2), 1.0);  // normalised: n = (0, 0, 1), d = 0.5
    CHECK(floor.n[2] == 1.0 && floor.d == 0.5);
    Sphere ball(0.25);
    DistanceRequest req(true, true, 0, 0);
    DistanceResult res;
    distance(&floor, Transform3f(), &ball, Transform3f(Vec3f(0.3, -0.2, 1.5)), req, res);
    CHECK_CLOSE(res.min_distance, 1.5 - 0.25 - 0.5, 1e-12);
    CHECK_CLOSE(res.nearest_points[1][2], 1.25, 1e-12);
    CHECK(res.normal[2] == 1.0);
    Plane wall(1, 0, 0, 2.0);
    DistanceResult r2;
    distance(&ball, Transform3f(Vec3f(5, 0, 0)), &wall, Transform3f(), req, r2);
    CHECK_CLOSE(r2.min_distance, 3 - 0.25, 1e-12);
    CHECK(r2.normal[0] == -1.0);  // from the sphere towards the wall
    CollisionRequest creq;
    CollisionResult cres;
    CHECK(collide(&ball, Transform3f(Vec3f(2.1, 0, 0)), &wall, Transform3f(), creq, cres) == 1);
    floor.d = 1.3;  // changed in place: registered anew
    DistanceResult r3;
    distance(&floor, Transform3f(), &ball, Transform3f(Vec3f(0.3, -0.2, 1.5)), req, r3);
    CHECK_CLOSE(r3.min_distance, 1.5 - 0.25 - 1.3, 1e-12);
  }
  {  // distance_box_box_1 (box_box_distance.cpp:62-101)
    CollisionGeometryPtr_t s1(new Box(6, 10, 2));
    CollisionGeometryPtr_t s2(new Box(2, 2, 2));
    Transform3f tf1;
    Transform3f tf2(Vec3f(25, 20, 5));
    CollisionObject o1(s1, tf1);
    CollisionObject o2(s2, tf2);
    DistanceRequest distanceRequest(true, true, 0, 0);
    DistanceResult distanceResult;
    distance(&o1, &o2, distanceRequest, distanceResult);
    double dx = 25 - 3 - 1, dy = 20 - 5 - 1, dz = 5 - 1 - 1;
    const Vec3f& p1 = distanceResult.nearest_points[0];
    const Vec3f& p2 = distanceResult.nearest_points[1];
    CHECK_CLOSE(distanceResult.min_distance, std::sqrt(dx * dx + dy * dy + dz * dz), 1e-4);
    CHECK_CLOSE(p1[0], 3, 1e-6); CHECK_CLOSE(p1[1], 5, 1e-6); CHECK_CLOSE(p1[2], 1, 1e-6);
    CHECK_CLOSE(p2[0], 24, 1e-6); CHECK_CLOSE(p2[1], 19, 1e-6); CHECK_CLOSE(p2[2], 4, 1e-6);
  }
  {  // distance_box_box_2 (:103-142)
    CollisionGeometryPtr_t s1(new Box(6, 10, 2));
    CollisionGeometryPtr_t s2(new Box(2, 2, 2));
    static double pi = M_PI;
    Transform3f tf1;
    Transform3f tf2(makeQuat(cos(pi / 8), sin(pi / 8) / sqrt(3), sin(pi / 8) / sqrt(3), sin(pi / 8) / sqrt(3)),
                    Vec3f(0, 0, 10));
    CollisionObject o1(s1, tf1), o2(s2, tf2);
    DistanceRequest distanceRequest(true, true, 0, 0);
    DistanceResult distanceResult;
    distance(&o1, &o2, distanceRequest, distanceResult);
    const Vec3f& p1 = distanceResult.nearest_points[0];
    const Vec3f& p2 = distanceResult.nearest_points[1];
    CHECK_CLOSE(distanceResult.min_distance, -1.62123444 + 10 - 1, 1e-4);
    CHECK_CLOSE(p1[0], 0.60947571, 1e-4); CHECK_CLOSE(p1[1], 0.01175873, 1e-4); CHECK_CLOSE(p1[2], 1, 1e-6);
    CHECK_CLOSE(p2[0], 0.60947571, 1e-4); CHECK_CLOSE(p2[1], 0.01175873, 1e-4);
    CHECK_CLOSE(p2[2], -1.62123444 + 10, 1e-4);
  }
  {  // distance_box_box_4 (:218-254)
    Box s1(1, 1, 1), s2(1, 1, 1);
    DistanceRequest distanceRequest(true, true, 0, 0);
    DistanceResult distanceResult;
    Transform3f tf1(Vec3f(2, 0, 0)), tf2;
    distance(&s1, tf1, &s2, tf2, distanceRequest, distanceResult);
    CHECK_CLOSE(distanceResult.min_distance, 1., 1e-4);
    tf1.setTranslation(Vec3f(1.01, 0, 0));
    distanceResult.clear();
    distance(&s1, tf1, &s2, tf2, distanceRequest, distanceResult);
    CHECK_CLOSE(distanceResult.min_distance, 0.01, 2e-3);
    tf1.setTranslation(Vec3f(0.99, 0, 0));
    distanceResult.clear();
    distance(&s1, tf1, &s2, tf2, distanceRequest, distanceResult);
    CHECK_CLOSE(distanceResult.min_distance, -0.01, 2e-3);
    tf1.setTranslation(Vec3f(0, 0, 0));
    distanceResult.clear();
    distance(&s1, tf1, &s2, tf2, distanceRequest, distanceResult);
    CHECK_CLOSE(distanceResult.min_distance, -1., 2e-3);
    // results accumulate: without clear() a satisfied request returns early (shape_shape_func.h:57)
    tf1.setTranslation(Vec3f(5, 0, 0));
    CHECK(distance(&s1, tf1, &s2, tf2, distanceRequest, distanceResult) < 0);
  }
  {  // distance_capsulecapsule_transformZ2 (capsule_capsule.cpp:322-356)
    CollisionGeometryPtr_t s1(new Capsule(5, 10)), s2(new Capsule(5, 10));
    Transform3f tf1;
    Transform3f tf2(makeQuat(sqrt(2) / 2, 0, sqrt(2) / 2, 0), Vec3f(0, 0, 25.1));
    CollisionObject o1(s1, tf1), o2(s2, tf2);
    DistanceRequest distanceRequest(true);
    DistanceResult distanceResult;
    distance(&o1, &o2, distanceRequest, distanceResult);
    CHECK_CLOSE(distanceResult.min_distance, 10.1, 1e-6);
    CHECK_CLOSE(distanceResult.nearest_points[0][2], 10, 1e-4);
    CHECK_CLOSE(distanceResult.nearest_points[1][2], 20.1, 1e-4);
  }
  {  // collide(): contact, security margin, functor, error paths (collision.cpp:69-130)
    Sphere s(1.0);
    CollisionRequest req;
    CollisionResult res;
    Transform3f a, b(Vec3f(1.5, 0, 0));
    CHECK(collide(&s, a, &s, b, req, res) == 1);
    CHECK(res.isCollision() && res.numContacts() == 1);
    const Contact& c = res.getContact(0);
    CHECK(std::fabs(c.penetration_depth + 0.5) < 1e-12 && std::fabs(c.normal[0] - 1) < 1e-12);
    CHECK(std::fabs(c.pos[0] - 0.75) < 1e-12 && c.b1 == Contact::NONE && c.o1 == &s);
    res.clear();
    b.setTranslation(Vec3f(2.1, 0, 0));
    CHECK(collide(&s, a, &s, b, req, res) == 0 && std::fabs(res.distance_lower_bound - 0.1) < 1e-9);
    req.security_margin = 0.2;
    res.clear();
    CHECK(ComputeCollision(&s, &s)(a, b, req, res) == 1);
    req.security_margin = -std::numeric_limits<double>::infinity();
    CHECK(collide(&s, a, &s, b, req, res) == 0 && !res.isCollision());
    req.security_margin = 0;
    req.num_max_contacts = 0;
    bool thrown = false;
    try { res.clear(); collide(&s, a, &s, b, req, res); } catch (const std::invalid_argument&) { thrown = true; }
    CHECK(thrown);
    // convex vs box through the batched form
    std::vector<Vec3f> pts;
    for (int i = 0; i < 8; ++i) pts.push_back(Vec3f((i & 1) ? .5 : -.5, (i & 2) ? .5 : -.5, (i & 4) ? .5 : -.5));
    ConvexBase cube(pts);
    Box bx(1, 1, 1);
    BatchNarrowPhase batch;
    for (int k = 0; k < 100; ++k) batch.add(&cube, Transform3f(), &bx, Transform3f(Vec3f(0.5 + 0.01 * k, 0, 0)));
    std::vector<hfb_distance_result> r = batch.distance(DistanceRequest());
    for (int k = 0; k < 100; ++k) CHECK(std::fabs(r[k].min_distance - (-0.5 + 0.01 * k)) < 1e-6);
    std::vector<hfb_contact> cc = batch.collide(CollisionRequest());
    CHECK(cc[10].num_contacts == 1 && cc[90].num_contacts == 0);
  }
  {  // BVHModel<OBBRSS>: a unit cube as 12 triangles, the build protocol of the reference
     // (test/utility.cpp generateBVHModel(Box) style); mesh-shape and mesh-mesh queries
    BVHModel<OBBRSS> cube;
    std::vector<Vec3f> ps;
    for (int i = 0; i < 8; ++i) ps.push_back(Vec3f((i & 1) ? .5 : -.5, (i & 2) ? .5 : -.5, (i & 4) ? .5 : -.5));
    const int q[6][4] = {{0, 2, 3, 1}, {4, 5, 7, 6}, {0, 1, 5, 4}, {2, 6, 7, 3}, {0, 4, 6, 2}, {1, 3, 7, 5}};
    std::vector<Triangle> ts;
    for (int f = 0; f < 6; ++f) {
      ts.push_back(Triangle(q[f][0], q[f][1], q[f][2]));
      ts.push_back(Triangle(q[f][0], q[f][2], q[f][3]));
    }
    CHECK(cube.addVertex(Vec3f(0, 0, 0)) == BVH_ERR_BUILD_OUT_OF_SEQUENCE);
    cube.beginModel();
    cube.addSubModel(ps, ts);
    CHECK(cube.endModel() == BVH_OK);
    CHECK(cube.getNodeType() == BV_OBBRSS && cube.getNumBVs() == 23 && cube.num_tris == 12);
    Sphere sph(0.25);
    DistanceRequest dreq(true, true, 0, 0);
    DistanceResult dres;
    // sphere 1 away from the +x face: distance 1 - 0.5 - 0.25
    double d = distance(&cube, Transform3f(), &sph, Transform3f(Vec3f(1, 0.1, -0.2)), dreq, dres);
    CHECK(std::fabs(d - 0.25) < 1e-9);
    CHECK(dres.b1 >= 0 && dres.b1 < 12 && dres.b2 == -1);
    CHECK(std::fabs(dres.nearest_points[0][0] - 0.5) < 1e-9 && std::fabs(dres.nearest_points[1][0] - 0.75) < 1e-9);
    // operands swapped: same distance, points swapped, triangle id still reported in b1 (distance.cpp:74-89)
    DistanceResult dres2;
    d = distance(&sph, Transform3f(Vec3f(1, 0.1, -0.2)), &cube, Transform3f(), dreq, dres2);
    CHECK(std::fabs(d - 0.25) < 1e-9 && std::fabs(dres2.nearest_points[0][0] - 0.75) < 1e-9);
    CollisionRequest creq;
    CollisionResult cres;
    CHECK(collide(&cube, Transform3f(), &sph, Transform3f(Vec3f(0.7, 0, 0)), creq, cres) == 1);
    CHECK(cres.getContact(0).b1 >= 0 && cres.getContact(0).b2 == -1);
    cres.clear();
    CHECK(collide(&cube, Transform3f(), &sph, Transform3f(Vec3f(0.8, 0, 0)), creq, cres) == 0);
    // mesh-mesh: two cubes, gap 0.5 along x
    DistanceResult mm;
    d = distance(&cube, Transform3f(), &cube, Transform3f(Vec3f(1.5, 0.2, 0.1)), dreq, mm);
    CHECK(std::fabs(d - 0.5) < 1e-9 && mm.b1 >= 0 && mm.b2 >= 0);
    cres.clear();
    CHECK(collide(&cube, Transform3f(), &cube, Transform3f(Vec3f(0.9, 0.2, 0.1)), creq, cres) == 1);
    CHECK(cres.getContact(0).b1 >= 0 && cres.getContact(0).b2 >= 0);
    cres.clear();
    CHECK(collide(&cube, Transform3f(), &cube, Transform3f(Vec3f(1.1, 0.2, 0.1)), creq, cres) == 0);
  }
  {  // the broadphase seam of BASELINE config 5 (compare test/broadphase_collision_1.cpp: a manager, a collecting
     // callback, and the narrow phase over what it collected): boxes as generateEnvironments makes them
    std::vector<std::unique_ptr<CollisionObject>> own;
    std::vector<CollisionObject*> env;
    CollisionGeometryPtr_t box(new Box(5, 10, 20));
    unsigned long long rng = 12345;
    auto rnd = [&]() {
      rng = rng * 6364136223846793005ull + 1442695040888963407ull;
      return (double)(rng >> 11) / 9007199254740992.0;
    };
    const int n = 3000;
    for (int i = 0; i < n; ++i) {
      const double a = 6.28 * rnd(), b = 6.28 * rnd();
      Quatf q = makeQuat(std::cos(a / 2), std::sin(a / 2) * std::cos(b), std::sin(a / 2) * std::sin(b), 0);
      own.emplace_back(new CollisionObject(box, Transform3f(q.toRotationMatrix(), Vec3f(120 * rnd(), 120 * rnd(), 120 * rnd()))));
      env.push_back(own.back().get());
    }
    DynamicAABBTreeCollisionManager manager;
    manager.registerObjects(env);
    manager.setup();
    CollisionCallBackCollect collect(0);
    manager.collide(&collect);
    // brute force over the boxes the manager computed: the same set of pairs
    size_t brute = 0;
    for (int i = 0; i < n; ++i)
      for (int j = i + 1; j < n; ++j)
        if (env[i]->getAABB().overlap(env[j]->getAABB())) {
          ++brute;
          CHECK(collect.exist(env[i], env[j]) || collect.exist(env[j], env[i]));
        }
    CHECK(brute == collect.numCollisionPairs() && brute > 1000);
    // every box of a rotated Box(5, 10, 20) holds the object: |extent| >= the half diagonal's projection
    CHECK(env[0]->getAABB().max_[0] - env[0]->getAABB().min_[0] >= 5 - 1e-9);
    // the batch: the candidate pairs through the narrow phase in one call == collide() pair by pair
    std::vector<uint32_t> first, second;
    CollisionRequest creq;
    std::vector<hfb_contact> recs = collide(manager, creq, first, second);
    CHECK(recs.size() == brute);
    size_t hits = 0, agree = 0;
    for (size_t k = 0; k < recs.size(); k += 7) {
      CollisionResult one;
      const size_t nc = collide(env[first[k]], env[second[k]], creq, one);
      agree += (nc > 0) == (recs[k].num_contacts > 0);
      hits += nc > 0;
      if (nc > 0) CHECK(one.getContact(0).penetration_depth == recs[k].distance);
    }
    CHECK(agree == (recs.size() + 6) / 7 && hits > 10);
    // moving one object and update(): its box follows
    env[5]->setTransform(Transform3f(Vec3f(-500, -500, -500)));
    manager.update();
    CHECK(env[5]->getAABB().min_[0] == -502.5 && env[5]->getAABB().max_[2] == -490);  // identity rotation: translate(aabb_local)
    collect.init();
    manager.collide(&collect);
    for (size_t k = 0; k < collect.numCollisionPairs(); ++k)
      CHECK(collect.getCollisionPairs()[k].first != env[5] && collect.getCollisionPairs()[k].second != env[5]);
  }
  std::printf(failures ? "HOST-API-FAILED %d\n" : "HOST-API-OK\n", failures);
  return failures ? 1 : 0;
}
