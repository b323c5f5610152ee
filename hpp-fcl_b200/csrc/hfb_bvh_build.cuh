// Host-side OBBRSS tree builder: what BVHModel<OBBRSS>::endModel() does in the reference for a
// triangle model with the default splitter (SPLIT_METHOD_MEAN):
//   buildTree / recursiveBuildTree      src/BVH/BVH_model.cpp:860-960
//   BVFitter<OBBRSS>::fit(prims, n)     src/BVH/BV_fitter.cpp:501-531 (one eigen decomposition for both halves)
//   getCovariance (triangle branch)     src/BVH/BVH_utility.cpp:183-259
//   getExtentAndCenter (mesh branch)    src/BVH/BVH_utility.cpp:529-584
//   getRadiusAndOriginAndRectangleSize  src/BVH/BVH_utility.cpp:264-482 (fit_rss_rectangle in hfb_bvh.cuh)
//   BVSplitter<OBBRSS> mean rule/apply  src/BVH/BV_splitter.cpp:80-118,242-278
// The walks (k_bvh) only ever read the tree, so building stays a host job as in the reference; a
// binding can equally hand over the reference's own BVHModel::bvs.  The output is
// bit-identical to the reference's tree: the traversal order, and with it every query result, depends
// on it.
#pragma once
#include <vector>

#include "hfb_bvh.cuh"

namespace hfb {

inline void tri_covariance(const TriPts& P, unsigned n, double M[6]) {
  v3 S1 = mk(0, 0, 0);
  double s00 = 0, s11 = 0, s22 = 0, s01 = 0, s02 = 0, s12 = 0;
  for (unsigned i = 0; i < n; ++i) {
    const v3 p1 = P.at(3 * (int)i), p2 = P.at(3 * (int)i + 1), p3 = P.at(3 * (int)i + 2);
    S1.x += (p1.x + p2.x + p3.x);
    S1.y += (p1.y + p2.y + p3.y);
    S1.z += (p1.z + p2.z + p3.z);
    s00 += (p1.x * p1.x + p2.x * p2.x + p3.x * p3.x);
    s11 += (p1.y * p1.y + p2.y * p2.y + p3.y * p3.y);
    s22 += (p1.z * p1.z + p2.z * p2.z + p3.z * p3.z);
    s01 += (p1.x * p1.y + p2.x * p2.y + p3.x * p3.y);
    s02 += (p1.x * p1.z + p2.x * p2.z + p3.x * p3.z);
    s12 += (p1.y * p1.z + p2.y * p2.z + p3.y * p3.z);
  }
  const unsigned np = 3 * n;
  M[0] = s00 - S1.x * S1.x / np;
  M[1] = s11 - S1.y * S1.y / np;
  M[2] = s22 - S1.z * S1.z / np;
  M[3] = s01 - S1.x * S1.y / np;
  M[4] = s12 - S1.y * S1.z / np;
  M[5] = s02 - S1.x * S1.z / np;
}

inline void put_colmajor(double* o, const m3& A) {
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) o[3 * c + r] = mel(A, r, c);
}

inline void fit_triangles(const TriPts& P, unsigned n, hfb_bvh_node& node, m3& axes_out) {
  double M[6];
  tri_covariance(P, n, M);
  ObbD obb;
  RssD rss;
  fit_axes_from_covariance(M, obb.axes);
  rss.axes = obb.axes;
  // getExtentAndCenter, mesh branch: centre = axes * ((max + min) / 2)
  v3 mn = mk(DBL_MAX, DBL_MAX, DBL_MAX), mx = mk(-DBL_MAX, -DBL_MAX, -DBL_MAX);
  for (unsigned i = 0; i < 3 * n; ++i) {
    const v3 proj = mtmul(obb.axes, P.at((int)i));
    if (proj.x > mx.x) mx.x = proj.x;
    if (proj.x < mn.x) mn.x = proj.x;
    if (proj.y > mx.y) mx.y = proj.y;
    if (proj.y < mn.y) mn.y = proj.y;
    if (proj.z > mx.z) mx.z = proj.z;
    if (proj.z < mn.z) mn.z = proj.z;
  }
  obb.To = mmul(obb.axes, (mx + mn) / 2);
  obb.extent = (mx - mn) / 2;
  fit_rss_rectangle(P, (int)(3 * n), rss);
  put_colmajor(node.obb_axes, obb.axes);
  node.obb_To[0] = obb.To.x; node.obb_To[1] = obb.To.y; node.obb_To[2] = obb.To.z;
  node.obb_extent[0] = obb.extent.x; node.obb_extent[1] = obb.extent.y; node.obb_extent[2] = obb.extent.z;
  put_colmajor(node.rss_axes, rss.axes);
  node.rss_Tr[0] = rss.Tr.x; node.rss_Tr[1] = rss.Tr.y; node.rss_Tr[2] = rss.Tr.z;
  node.rss_length[0] = rss.l0;
  node.rss_length[1] = rss.l1;
  node.rss_radius = rss.radius;
  axes_out = obb.axes;
}

// nodes: 2 * nt - 1 entries.  Returns false on invalid input (no triangles, vertex index out of range).
inline bool build_obbrss_tree(const double* verts, uint32_t nv, const uint32_t* tris, uint32_t nt, hfb_bvh_node* nodes) {
  if (nt == 0 || !verts || !tris || !nodes) return false;
  for (size_t k = 0; k < 3 * (size_t)nt; ++k)
    if (tris[k] >= nv) return false;
  std::vector<uint32_t> prim(nt);
  for (uint32_t i = 0; i < nt; ++i) prim[i] = i;
  struct Frame { int bv_id; uint32_t first, num; };
  std::vector<Frame> stack;
  stack.push_back({0, 0, nt});
  int num_bvs = 1;
  auto vert = [&](uint32_t t, int k) {
    const double* p = verts + 3 * (size_t)tris[3 * (size_t)t + (unsigned)k];
    return mk(p[0], p[1], p[2]);
  };
  while (!stack.empty()) {
    const Frame f = stack.back();
    stack.pop_back();
    uint32_t* cur = prim.data() + f.first;
    hfb_bvh_node& node = nodes[f.bv_id];
    std::memset(&node, 0, sizeof(node));
    m3 axes;
    fit_triangles(TriPts{verts, tris, cur}, f.num, node, axes);
    node.first_primitive = f.first;
    node.num_primitives = f.num;
    if (f.num == 1) {
      node.first_child = -((int)cur[0] + 1);
      continue;
    }
    // computeRule_mean: split along the first OBB axis at the mean of the triangle vertices
    const v3 split_vector = mcol(axes, 0);
    v3 c = mk(0, 0, 0);
    for (uint32_t i = 0; i < f.num; ++i) c = c + (vert(cur[i], 0) + vert(cur[i], 1) + vert(cur[i], 2));
    const double split_value = dot(c, split_vector) / (3 * f.num);
    node.first_child = num_bvs;
    num_bvs += 2;
    uint32_t c1 = 0;
    for (uint32_t i = 0; i < f.num; ++i) {
      const v3 p = (vert(cur[i], 0) + vert(cur[i], 1) + vert(cur[i], 2)) / 3.;
      if (dot(split_vector, p) > split_value) {  // BVSplitter::apply: stays in the right part
      } else {
        const uint32_t tmp = cur[i];
        cur[i] = cur[c1];
        cur[c1] = tmp;
        c1++;
      }
    }
    if (c1 == 0 || c1 == f.num) c1 = f.num / 2;
    // the left subtree is built (and numbers its nodes) before the right one
    stack.push_back({node.first_child + 1, f.first + c1, f.num - c1});
    stack.push_back({node.first_child, f.first, c1});
  }
  return true;
}

}  // namespace hfb
