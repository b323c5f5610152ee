// Broadphase feed of the batched narrow phase (BASELINE config 5): world-space AABBs of a scene's
// objects and the pairs whose AABBs overlap -- what a BroadPhaseCollisionManager hands to its collision callback.
//
// Replaces, for this path only:
//   computeLocalAABB of the shapes / BVHModel            src/shape/geometric_shapes.cpp:145-260, src/BVH/BVH_model.cpp
//       (computeBV<AABB, S> with the identity transform: src/shape/geometric_shapes_utility.cpp:264-420)
//   CollisionObject::computeAABB                          include/hpp/fcl/collision_object.h:258-278
//   AABB::overlap                                         include/hpp/fcl/BV/AABB.h:111-118
//   DynamicAABBTreeCollisionManager::collide(callback)    src/broadphase/broadphase_dynamic_AABB_tree.cpp:336-407,716-721
//       -- the SET of pairs it reports (every pair of objects with overlapping AABBs, once); the order is the
//       manager's own business in the reference as well (its managers disagree among themselves).
// Design: no tree.  A uniform grid whose cell is as large as the largest AABB: objects are bucketed by the cell of
// their centre, so a partner can only sit in the 27 cells around -- a counting sort and one sweep, the same three
// steps on the host (here) and on the device (hfb_broadphase.cu).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "hfb_arena.cuh"

namespace hfb {

struct LocalAabb {  // CollisionGeometry::aabb_local
  double mn[3], mx[3];
};

// aabb_local of a shape record (computeLocalAABB); false for node types without one here
inline bool shape_local_aabb(const HostArena& A, const hfb_shape& s, LocalAabb& b) {
  double h[3];
  switch (s.type) {
    case HFB_GEOM_BOX: h[0] = s.p[0]; h[1] = s.p[1]; h[2] = s.p[2]; break;            // half sides
    case HFB_GEOM_SPHERE: h[0] = h[1] = h[2] = s.p[0]; break;
    case HFB_GEOM_ELLIPSOID: h[0] = s.p[0]; h[1] = s.p[1]; h[2] = s.p[2]; break;
    case HFB_GEOM_CAPSULE: h[0] = h[1] = s.p[0]; h[2] = s.p[1] + s.p[0]; break;      // 0.5 * lz + radius
    case HFB_GEOM_CONE:
    case HFB_GEOM_CYLINDER: h[0] = h[1] = s.p[0]; h[2] = s.p[1]; break;
    case HFB_GEOM_CONVEX:
    case HFB_GEOM_TRIANGLE: {
      if (s.data >= A.cvx.size()) return false;
      const ConvexDesc& d = A.cvx[s.data];
      const uint32_t nv = s.type == HFB_GEOM_TRIANGLE ? 3u : d.nv;
      for (int k = 0; k < 3; ++k) {
        b.mn[k] = DBL_MAX;
        b.mx[k] = -DBL_MAX;
        for (uint32_t i = 0; i < nv; ++i) {
          const double v = A.pool[d.off + (size_t)k * d.vpad + i];
          b.mn[k] = v < b.mn[k] ? v : b.mn[k];
          b.mx[k] = v > b.mx[k] ? v : b.mx[k];
        }
      }
      h[0] = -1;
    } break;
    case HFB_BV_OBB:
    case HFB_BV_OBBRSS: {
      if (s.data >= A.bvh_desc.size()) return false;
      const BvhDesc& d = A.bvh_desc[s.data];
      for (int k = 0; k < 3; ++k) {
        b.mn[k] = DBL_MAX;
        b.mx[k] = -DBL_MAX;
      }
      for (uint32_t i = 0; i < d.num_verts; ++i)
        for (int k = 0; k < 3; ++k) {
          const double v = A.bvh_verts[3 * ((size_t)d.vert_off + i) + k];
          b.mn[k] = v < b.mn[k] ? v : b.mn[k];
          b.mx[k] = v > b.mx[k] ? v : b.mx[k];
        }
      h[0] = -1;
    } break;
    case HFB_GEOM_PLANE:
    case HFB_GEOM_HALFSPACE: {
      // computeBV<AABB, Halfspace / Plane> with the identity pose (geometric_shapes_utility.cpp:391-456): everything,
      // except along an axis the normal is exactly aligned with.  transform() builds a new Halfspace / Plane, whose
      // constructor normalises (n, d) once more (unitNormalTest)
      if (s.data >= A.pool.size()) return false;
      double n[3] = {s.p[0], s.p[1], s.p[2]};
      double d = A.pool[s.data] + ((n[0] * 0.0 + n[1] * 0.0) + n[2] * 0.0);
      const double l = sqrt((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]);
      if (l > 0) {
        const double inv_l = 1.0 / l;
        for (int k = 0; k < 3; ++k) n[k] *= inv_l;
        d *= inv_l;
      } else {
        n[0] = 1;
        n[1] = n[2] = 0;
        d = 0;
      }
      for (int k = 0; k < 3; ++k) {
        b.mn[k] = -DBL_MAX;
        b.mx[k] = DBL_MAX;
      }
      int axis = -1;
      if (n[1] == 0.0 && n[2] == 0.0) axis = 0;
      else if (n[0] == 0.0 && n[2] == 0.0) axis = 1;
      else if (n[0] == 0.0 && n[1] == 0.0) axis = 2;
      if (axis >= 0) {
        if (s.type == HFB_GEOM_HALFSPACE) {
          if (n[axis] < 0) b.mn[axis] = -d;
          else if (n[axis] > 0) b.mx[axis] = d;
        } else {
          if (n[axis] < 0) b.mn[axis] = b.mx[axis] = -d;
          else if (n[axis] > 0) b.mn[axis] = b.mx[axis] = d;
        }
      }
      h[0] = -1;
    } break;
    default: return false;
  }
  if (h[0] >= 0)
    for (int k = 0; k < 3; ++k) {
      b.mn[k] = -h[k];
      b.mx[k] = h[k];
    }
  if (s.ssr > 0)  // the swept-sphere radius inflates the local box (geometric_shapes.cpp:147-151)
    for (int k = 0; k < 3; ++k) {
      b.mn[k] -= s.ssr;
      b.mx[k] += s.ssr;
    }
  return true;
}

// CollisionObject::computeAABB: the box around the rotated local box, translated.  `out`: min xyz, max xyz
HFB_HD void object_aabb(const double* lmn, const double* lmx, const hfb_transform& tf, double* out) {
  // Matrix3f::isIdentity(): |R - I| <= 1e-12 element-wise (Eigen's fuzzy test against the identity's norm of 1)
  bool ident = true;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      const double v = tf.R[c * 3 + r] - (r == c ? 1.0 : 0.0);
      if (fabs(v) > 1e-12) ident = false;
    }
  if (ident) {
    for (int k = 0; k < 3; ++k) {
      out[k] = lmn[k] + tf.T[k];
      out[3 + k] = lmx[k] + tf.T[k];
    }
    return;
  }
  for (int k = 0; k < 3; ++k) {
    double a[3], b[3];
    for (int j = 0; j < 3; ++j) {
      const double r = tf.R[j * 3 + k];  // R(k, j), column-major storage
      const double lo = r * lmn[j], hi = r * lmx[j];
      a[j] = lo < hi ? lo : hi;
      b[j] = lo < hi ? hi : lo;
    }
    out[k] = tf.T[k] + ((a[0] + a[1]) + a[2]);
    out[3 + k] = tf.T[k] + ((b[0] + b[1]) + b[2]);
  }
}

HFB_HD bool aabb_overlap(const double* a, const double* b) {  // AABB::overlap: closed intervals
  if (a[0] > b[3] || a[1] > b[4] || a[2] > b[5]) return false;
  if (a[3] < b[0] || a[4] < b[1] || a[5] < b[2]) return false;
  return true;
}

// A box that reaches past this (a Halfspace or Plane: +-DBL_MAX, or +-inf once rotated) is "unbounded": it takes no part
// in the grid -- a cell as large as the largest box would be the whole scene -- and is tested against every object.
#define HFB_BP_UNBOUNDED 1e300
HFB_HD bool aabb_unbounded(const double* b) {
  for (int k = 0; k < 6; ++k)
    if (!(b[k] > -HFB_BP_UNBOUNDED && b[k] < HFB_BP_UNBOUNDED)) return true;
  return false;
}
// the marker hfb_scene_aabbs writes for an object without a box (no such geometry, node type without aabb_local)
HFB_HD bool aabb_empty(const double* b) { return b[0] == DBL_MAX && b[3] == -DBL_MAX; }

// grid geometry shared by the host and device pair finders
struct BroadGrid {
  double origin[3];
  double inv_cell;
  int dim[3];
};
HFB_HD int grid_coord(const BroadGrid& g, double c, int k) {
  int v = (int)floor((c - g.origin[k]) * g.inv_cell);
  return v < 0 ? 0 : (v >= g.dim[k] ? g.dim[k] - 1 : v);
}
HFB_HD unsigned grid_cell(const BroadGrid& g, const double* bb) {
  const int x = grid_coord(g, 0.5 * (bb[0] + bb[3]), 0), y = grid_coord(g, 0.5 * (bb[1] + bb[4]), 1),
            z = grid_coord(g, 0.5 * (bb[2] + bb[5]), 2);
  return (unsigned)((z * g.dim[1] + y) * g.dim[0] + x);
}
// a grid for these boxes: cell = the largest box extent (so partners sit in adjacent cells), at most ~2^21 cells
inline BroadGrid make_grid(size_t n, const double* bb) {
  BroadGrid g;
  double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX}, ext = 0;
  for (size_t i = 0; i < n; ++i) {
    if (aabb_empty(bb + 6 * i) || aabb_unbounded(bb + 6 * i)) continue;
    for (int k = 0; k < 3; ++k) {
      const double c = 0.5 * (bb[6 * i + k] + bb[6 * i + 3 + k]), e = bb[6 * i + 3 + k] - bb[6 * i + k];
      lo[k] = c < lo[k] ? c : lo[k];
      hi[k] = c > hi[k] ? c : hi[k];
      ext = e > ext ? e : ext;
    }
  }
  if (!(lo[0] <= hi[0])) lo[0] = lo[1] = lo[2] = hi[0] = hi[1] = hi[2] = 0;  // nothing bounded
  double cell = ext > 0 ? ext : 1.0;
  for (;;) {
    double cells = 1;
    for (int k = 0; k < 3; ++k) {
      g.dim[k] = (int)floor((hi[k] - lo[k]) / cell) + 1;
      cells *= g.dim[k];
    }
    if (cells <= 2097152.0) break;
    cell *= 1.26;  // halves the number of cells
  }
  for (int k = 0; k < 3; ++k) g.origin[k] = lo[k];
  g.inv_cell = 1.0 / cell;
  return g;
}

// every pair i < j with overlapping boxes, in no particular order; returns their number (all of them are counted,
// at most `capacity` stored)
inline size_t broadphase_pairs_host(size_t n, const double* bb, uint32_t* first, uint32_t* second, size_t capacity) {
  if (n < 2) return 0;
  const BroadGrid g = make_grid(n, bb);
  const size_t ncell = (size_t)g.dim[0] * g.dim[1] * g.dim[2];
  std::vector<uint32_t> cell(n), start(ncell + 1, 0), order(n), unbounded;
  for (size_t i = 0; i < n; ++i) {
    if (aabb_empty(bb + 6 * i)) {
      cell[i] = 0xffffffffu;
    } else if (aabb_unbounded(bb + 6 * i)) {
      cell[i] = 0xfffffffeu;
      unbounded.push_back((uint32_t)i);
    } else {
      cell[i] = grid_cell(g, bb + 6 * i);
      ++start[cell[i] + 1];
    }
  }
  for (size_t c = 0; c < ncell; ++c) start[c + 1] += start[c];
  {
    std::vector<uint32_t> cur(start.begin(), start.end() - 1);
    for (size_t i = 0; i < n; ++i)
      if (cell[i] < 0xfffffffeu) order[cur[cell[i]]++] = (uint32_t)i;
  }
  size_t count = 0;
  // the unbounded objects against everything (a pair of two of them once)
  for (uint32_t u : unbounded)
    for (size_t j = 0; j < n; ++j) {
      if (j == u || cell[j] == 0xffffffffu || (cell[j] == 0xfffffffeu && j < u)) continue;
      if (!aabb_overlap(bb + 6 * (size_t)u, bb + 6 * j)) continue;
      if (count < capacity) {
        first[count] = u < j ? u : (uint32_t)j;
        second[count] = u < j ? (uint32_t)j : u;
      }
      ++count;
    }
  for (size_t i = 0; i < n; ++i) {
    if (cell[i] >= 0xfffffffeu) continue;
    const unsigned c = cell[i];
    const int cx = (int)(c % (unsigned)g.dim[0]), cy = (int)((c / (unsigned)g.dim[0]) % (unsigned)g.dim[1]),
              cz = (int)(c / ((unsigned)g.dim[0] * (unsigned)g.dim[1]));
    for (int dz = -1; dz <= 1; ++dz)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int x = cx + dx, y = cy + dy, z = cz + dz;
          if (x < 0 || y < 0 || z < 0 || x >= g.dim[0] || y >= g.dim[1] || z >= g.dim[2]) continue;
          const size_t nc = ((size_t)z * g.dim[1] + y) * g.dim[0] + x;
          for (uint32_t q = start[nc]; q < start[nc + 1]; ++q) {
            const uint32_t j = order[q];
            if (j <= i) continue;
            if (!aabb_overlap(bb + 6 * i, bb + 6 * (size_t)j)) continue;
            if (count < capacity) {
              first[count] = (uint32_t)i;
              second[count] = j;
            }
            ++count;
          }
        }
  }
  return count;
}

}  // namespace hfb
