// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.hpp).
// Restatement of the OBBRSS BVH path:
//   builder      src/BVH/BVH_model.cpp:860-960 (buildTree / recursiveBuildTree),
//                src/BVH/BV_fitter.cpp:49-135,145-200,455-531 (fit1/2/3/n, BVFitter<OBBRSS>::fit),
//                src/BVH/BVH_utility.cpp:183-600 (getCovariance, getRadiusAndOriginAndRectangleSize,
//                getExtentAndCenter), include/hpp/fcl/internal/tools.h:60-203 (generateCoordinateSystem,
//                Jacobi eigen), src/BVH/BV_splitter.cpp:80-118,242-280 (mean split)
//   shape BV     src/shape/geometric_shapes_utility.cpp:47-262 (getBoundVertices),
//                include/hpp/fcl/shape/geometric_shapes_utility.h:73-83 (computeBV = fit(bound vertices))
//   BV tests     src/BV/RSS.cpp:67-713,995-1005 (segCoords, inVoronoi, rectDistance, distance),
//                src/BV/OBB.cpp:290-393,475-483 (obbDisjointAndLowerBoundDistance, overlap)
//   traversal    src/traversal/traversal_recurse.cpp:44-85,153-203, src/collision_node.cpp:64-91,
//                include/hpp/fcl/internal/traversal_node_bvh_shape.h:97-194,286-478,
//                include/hpp/fcl/internal/traversal_node_setup.h:655-694,746-810
#include <stdexcept>
#include <algorithm>
#include <cassert>
#include <cstdio>

#include "bvh.hpp"

namespace oracle {

// ----------------------------------------------------------------- tools.h ----
// generateCoordinateSystem (tools.h:60-96)
static void generateCoordinateSystem(const V3& w, V3& u, V3& v) {
  double inv_length;
  if (std::fabs(w[0]) >= std::fabs(w[1])) {
    inv_length = 1.0 / std::sqrt(w[0] * w[0] + w[2] * w[2]);
    u = V3(-w[2] * inv_length, 0, w[0] * inv_length);
    v = V3(w[1] * u[2], w[2] * u[0] - w[0] * u[2], -w[1] * u[0]);
  } else {
    inv_length = 1.0 / std::sqrt(w[1] * w[1] + w[2] * w[2]);
    u = V3(0, w[2] * inv_length, -w[1] * inv_length);
    v = V3(w[1] * u[2] - w[2] * u[1], -w[0] * u[2], w[0] * u[1]);
  }
}

// eigen (tools.h:103-203): cyclic Jacobi, at most 50 sweeps. vout[i] is the i-th ROW of the
// eigenvector matrix v (vout[i][k] = v[i][k]).
static void eigen(const M3& m, double dout[3], V3 vout[3]) {
  M3 R = m;
  const int n = 3;
  int j, iq, ip, i;
  double tresh, theta, tau, t, sm, s, h, g, c;
  double b[3], z[3], d[3];
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (ip = 0; ip < n; ++ip) {
    b[ip] = d[ip] = R.m[ip][ip];
    z[ip] = 0;
  }
  for (i = 0; i < 50; ++i) {
    sm = 0;
    for (ip = 0; ip < n; ++ip)
      for (iq = ip + 1; iq < n; ++iq) sm += std::fabs(R.m[ip][iq]);
    if (sm == 0.0) {
      for (int k = 0; k < 3; ++k) {
        vout[k] = V3(v[k][0], v[k][1], v[k][2]);
        dout[k] = d[k];
      }
      return;
    }
    if (i < 3) tresh = 0.2 * sm / (n * n);
    else tresh = 0.0;
    for (ip = 0; ip < n; ++ip) {
      for (iq = ip + 1; iq < n; ++iq) {
        g = 100.0 * std::fabs(R.m[ip][iq]);
        if (i > 3 && std::fabs(d[ip]) + g == std::fabs(d[ip]) && std::fabs(d[iq]) + g == std::fabs(d[iq]))
          R.m[ip][iq] = 0.0;
        else if (std::fabs(R.m[ip][iq]) > tresh) {
          h = d[iq] - d[ip];
          if (std::fabs(h) + g == std::fabs(h))
            t = (R.m[ip][iq]) / h;
          else {
            theta = 0.5 * h / (R.m[ip][iq]);
            t = 1.0 / (std::fabs(theta) + std::sqrt(1.0 + theta * theta));
            if (theta < 0.0) t = -t;
          }
          c = 1.0 / std::sqrt(1 + t * t);
          s = t * c;
          tau = s / (1.0 + c);
          h = t * R.m[ip][iq];
          z[ip] -= h;
          z[iq] += h;
          d[ip] -= h;
          d[iq] += h;
          R.m[ip][iq] = 0.0;
          for (j = 0; j < ip; ++j) {
            g = R.m[j][ip];
            h = R.m[j][iq];
            R.m[j][ip] = g - s * (h + g * tau);
            R.m[j][iq] = h + s * (g - h * tau);
          }
          for (j = ip + 1; j < iq; ++j) {
            g = R.m[ip][j];
            h = R.m[j][iq];
            R.m[ip][j] = g - s * (h + g * tau);
            R.m[j][iq] = h + s * (g - h * tau);
          }
          for (j = iq + 1; j < n; ++j) {
            g = R.m[ip][j];
            h = R.m[iq][j];
            R.m[ip][j] = g - s * (h + g * tau);
            R.m[iq][j] = h + s * (g - h * tau);
          }
          for (j = 0; j < n; ++j) {
            g = v[j][ip];
            h = v[j][iq];
            v[j][ip] = g - s * (h + g * tau);
            v[j][iq] = h + s * (g - h * tau);
          }
        }
      }
    }
    for (ip = 0; ip < n; ++ip) {
      b[ip] += z[ip];
      d[ip] = b[ip];
      z[ip] = 0.0;
    }
  }
  // "too many iterations": outputs are left untouched by the reference
}

// axisFromEigen (BV_fitter.cpp:49-76); axes columns = (max, mid, max x mid)
static void axisFromEigen(const V3 eigenV[3], const double eigenS[3], M3& axes) {
  int min, mid, max;
  if (eigenS[0] > eigenS[1]) { max = 0; min = 1; } else { min = 0; max = 1; }
  if (eigenS[2] < eigenS[min]) { mid = min; min = 2; }
  else if (eigenS[2] > eigenS[max]) { mid = max; max = 2; }
  else { mid = 2; }
  (void)min;
  axes.m[0][0] = eigenV[0][max]; axes.m[1][0] = eigenV[1][max]; axes.m[2][0] = eigenV[2][max];
  axes.m[0][1] = eigenV[0][mid]; axes.m[1][1] = eigenV[1][mid]; axes.m[2][1] = eigenV[2][mid];
  axes.m[0][2] = eigenV[1][max] * eigenV[2][mid] - eigenV[1][mid] * eigenV[2][max];
  axes.m[1][2] = eigenV[0][mid] * eigenV[2][max] - eigenV[0][max] * eigenV[2][mid];
  axes.m[2][2] = eigenV[0][max] * eigenV[1][mid] - eigenV[0][mid] * eigenV[1][max];
}

static inline void set_col(M3& A, int c, const V3& v) { A.m[0][c] = v.x; A.m[1][c] = v.y; A.m[2][c] = v.z; }
static inline M3 identity3() { M3 I; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) I.m[i][j] = i == j; return I; }

// ------------------------------------------------------------ BVH_utility.cpp ---
// a "point source": either a plain point list or triangles (ts) indexed by `indices`
struct PointSrc {
  const V3* ps;
  const Tri* ts;            // may be null
  const unsigned* indices;  // may be null
  unsigned n;
};

// getCovariance (BVH_utility.cpp:183-259), ps2 == NULL
static void getCovariance(const PointSrc& S, M3& M) {
  V3 S1(0, 0, 0);
  double S2[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  if (S.ts) {
    for (unsigned i = 0; i < S.n; ++i) {
      const Tri& t = S.indices ? S.ts[S.indices[i]] : S.ts[i];
      const V3& p1 = S.ps[t.v[0]];
      const V3& p2 = S.ps[t.v[1]];
      const V3& p3 = S.ps[t.v[2]];
      S1.x += (p1[0] + p2[0] + p3[0]);
      S1.y += (p1[1] + p2[1] + p3[1]);
      S1.z += (p1[2] + p2[2] + p3[2]);
      S2[0][0] += (p1[0] * p1[0] + p2[0] * p2[0] + p3[0] * p3[0]);
      S2[1][1] += (p1[1] * p1[1] + p2[1] * p2[1] + p3[1] * p3[1]);
      S2[2][2] += (p1[2] * p1[2] + p2[2] * p2[2] + p3[2] * p3[2]);
      S2[0][1] += (p1[0] * p1[1] + p2[0] * p2[1] + p3[0] * p3[1]);
      S2[0][2] += (p1[0] * p1[2] + p2[0] * p2[2] + p3[0] * p3[2]);
      S2[1][2] += (p1[1] * p1[2] + p2[1] * p2[2] + p3[1] * p3[2]);
    }
  } else {
    for (unsigned i = 0; i < S.n; ++i) {
      const V3& p = S.indices ? S.ps[S.indices[i]] : S.ps[i];
      S1 += p;
      S2[0][0] += (p[0] * p[0]);
      S2[1][1] += (p[1] * p[1]);
      S2[2][2] += (p[2] * p[2]);
      S2[0][1] += (p[0] * p[1]);
      S2[0][2] += (p[0] * p[2]);
      S2[1][2] += (p[1] * p[2]);
    }
  }
  const unsigned n_points = (S.ts ? 3 : 1) * S.n;
  M.m[0][0] = S2[0][0] - S1[0] * S1[0] / n_points;
  M.m[1][1] = S2[1][1] - S1[1] * S1[1] / n_points;
  M.m[2][2] = S2[2][2] - S1[2] * S1[2] / n_points;
  M.m[0][1] = S2[0][1] - S1[0] * S1[1] / n_points;
  M.m[1][2] = S2[1][2] - S1[1] * S1[2] / n_points;
  M.m[0][2] = S2[0][2] - S1[0] * S1[2] / n_points;
  M.m[1][0] = M.m[0][1];
  M.m[2][0] = M.m[0][2];
  M.m[2][1] = M.m[1][2];
}

// getExtentAndCenter (BVH_utility.cpp:487-584)
static void getExtentAndCenter(const PointSrc& S, const M3& axes, V3& center, V3& extent) {
  const double real_max = std::numeric_limits<double>::max();
  V3 min_coord(real_max, real_max, real_max), max_coord(-real_max, -real_max, -real_max);
  auto add = [&](const V3& p) {
    V3 proj = tmul(axes, p);
    for (int j = 0; j < 3; ++j) {
      if (proj[j] > max_coord[j]) max_coord[j] = proj[j];
      if (proj[j] < min_coord[j]) min_coord[j] = proj[j];
    }
  };
  for (unsigned i = 0; i < S.n; ++i) {
    const unsigned index = S.indices ? S.indices[i] : i;
    if (S.ts) {
      const Tri& t = S.ts[index];
      for (int j = 0; j < 3; ++j) add(S.ps[t.v[j]]);
    } else {
      add(S.ps[index]);
    }
  }
  if (S.ts) {
    V3 o((max_coord + min_coord) / 2);
    center = mul(axes, o);
  } else {
    center = mul(axes, max_coord + min_coord) / 2;  // axes * (max+min) / 2
  }
  extent = (max_coord - min_coord) / 2;
}

// getRadiusAndOriginAndRectangleSize (BVH_utility.cpp:264-482), ps2 == NULL
static void getRadiusAndOriginAndRectangleSize(const PointSrc& S, const M3& axes, V3& origin, double l[2],
                                               double& r) {
  const unsigned size_P = (S.ts ? 3 : 1) * S.n;
  std::vector<std::array<double, 3>> P(size_P);
  unsigned P_id = 0;
  const V3 c0 = axes.col(0), c1 = axes.col(1), c2 = axes.col(2);
  for (unsigned i = 0; i < S.n; ++i) {
    const unsigned index = S.indices ? S.indices[i] : i;
    if (S.ts) {
      const Tri& t = S.ts[index];
      for (int j = 0; j < 3; ++j) {
        const V3& v = S.ps[t.v[j]];
        P[P_id][0] = dot(c0, v);
        P[P_id][1] = dot(c1, v);
        P[P_id][2] = dot(c2, v);
        P_id++;
      }
    } else {
      const V3& v = S.ps[index];
      P[P_id][0] = dot(c0, v);
      P[P_id][1] = dot(c1, v);
      P[P_id][2] = dot(c2, v);
      P_id++;
    }
  }
  double minx, maxx, miny, maxy, minz, maxz;
  double cz, radsqr;
  minz = maxz = P[0][2];
  for (unsigned i = 1; i < size_P; ++i) {
    double z_value = P[i][2];
    if (z_value < minz) minz = z_value;
    else if (z_value > maxz) maxz = z_value;
  }
  r = 0.5 * (maxz - minz);
  radsqr = r * r;
  cz = 0.5 * (maxz + minz);

  unsigned minindex = 0, maxindex = 0;
  double mintmp, maxtmp;
  mintmp = maxtmp = P[0][0];
  for (unsigned i = 1; i < size_P; ++i) {
    double x_value = P[i][0];
    if (x_value < mintmp) { minindex = i; mintmp = x_value; }
    else if (x_value > maxtmp) { maxindex = i; maxtmp = x_value; }
  }
  double x, dz;
  dz = P[minindex][2] - cz;
  minx = P[minindex][0] + std::sqrt(std::max<double>(radsqr - dz * dz, 0));
  dz = P[maxindex][2] - cz;
  maxx = P[maxindex][0] - std::sqrt(std::max<double>(radsqr - dz * dz, 0));
  for (unsigned i = 0; i < size_P; ++i) {
    if (P[i][0] < minx) {
      dz = P[i][2] - cz;
      x = P[i][0] + std::sqrt(std::max<double>(radsqr - dz * dz, 0));
      if (x < minx) minx = x;
    } else if (P[i][0] > maxx) {
      dz = P[i][2] - cz;
      x = P[i][0] - std::sqrt(std::max<double>(radsqr - dz * dz, 0));
      if (x > maxx) maxx = x;
    }
  }
  minindex = maxindex = 0;
  mintmp = maxtmp = P[0][1];
  for (unsigned i = 1; i < size_P; ++i) {
    double y_value = P[i][1];
    if (y_value < mintmp) { minindex = i; mintmp = y_value; }
    else if (y_value > maxtmp) { maxindex = i; maxtmp = y_value; }
  }
  double y;
  dz = P[minindex][2] - cz;
  miny = P[minindex][1] + std::sqrt(std::max<double>(radsqr - dz * dz, 0));
  dz = P[maxindex][2] - cz;
  maxy = P[maxindex][1] - std::sqrt(std::max<double>(radsqr - dz * dz, 0));
  for (unsigned i = 0; i < size_P; ++i) {
    if (P[i][1] < miny) {
      dz = P[i][2] - cz;
      y = P[i][1] + std::sqrt(std::max<double>(radsqr - dz * dz, 0));
      if (y < miny) miny = y;
    } else if (P[i][1] > maxy) {
      dz = P[i][2] - cz;
      y = P[i][1] - std::sqrt(std::max<double>(radsqr - dz * dz, 0));
      if (y > maxy) maxy = y;
    }
  }
  // corners
  double dx, dy, u, t;
  double a = std::sqrt(0.5);
  for (unsigned i = 0; i < size_P; ++i) {
    if (P[i][0] > maxx) {
      if (P[i][1] > maxy) {
        dx = P[i][0] - maxx;
        dy = P[i][1] - maxy;
        u = dx * a + dy * a;
        t = (a * u - dx) * (a * u - dx) + (a * u - dy) * (a * u - dy) + (cz - P[i][2]) * (cz - P[i][2]);
        u = u - std::sqrt(std::max<double>(radsqr - t, 0));
        if (u > 0) { maxx += u * a; maxy += u * a; }
      } else if (P[i][1] < miny) {
        dx = P[i][0] - maxx;
        dy = P[i][1] - miny;
        u = dx * a - dy * a;
        t = (a * u - dx) * (a * u - dx) + (-a * u - dy) * (-a * u - dy) + (cz - P[i][2]) * (cz - P[i][2]);
        u = u - std::sqrt(std::max<double>(radsqr - t, 0));
        if (u > 0) { maxx += u * a; miny -= u * a; }
      }
    } else if (P[i][0] < minx) {
      if (P[i][1] > maxy) {
        dx = P[i][0] - minx;
        dy = P[i][1] - maxy;
        u = dy * a - dx * a;
        t = (-a * u - dx) * (-a * u - dx) + (a * u - dy) * (a * u - dy) + (cz - P[i][2]) * (cz - P[i][2]);
        u = u - std::sqrt(std::max<double>(radsqr - t, 0));
        if (u > 0) { minx -= u * a; maxy += u * a; }
      } else if (P[i][1] < miny) {
        dx = P[i][0] - minx;
        dy = P[i][1] - miny;
        u = -dx * a - dy * a;
        t = (-a * u - dx) * (-a * u - dx) + (-a * u - dy) * (-a * u - dy) + (cz - P[i][2]) * (cz - P[i][2]);
        u = u - std::sqrt(std::max<double>(radsqr - t, 0));
        if (u > 0) { minx -= u * a; miny -= u * a; }
      }
    }
  }
  origin = mul(axes, V3(minx, miny, cz));
  l[0] = std::max<double>(maxx - minx, 0);
  l[1] = std::max<double>(maxy - miny, 0);
}

// ------------------------------------------------------------- BV_fitter.cpp ----
// fit(Vec3f* ps, n, OBBRSS&) (:455-470): the OBB and the RSS halves are fitted independently
static void fit_points_obb(const V3* ps, unsigned n, OBB& bv) {
  switch (n) {
    case 1:  // OBB_fit_functions::fit1 :80-84
      bv.To = ps[0];
      bv.axes = identity3();
      bv.extent = V3(0, 0, 0);
      break;
    case 2: {  // fit2 :86-98
      const V3& p1 = ps[0];
      const V3& p2 = ps[1];
      V3 p1p2 = p1 - p2;
      double len_p1p2 = norm(p1p2);
      p1p2 = normalized(p1p2);
      set_col(bv.axes, 0, p1p2);
      V3 u, v;
      generateCoordinateSystem(p1p2, u, v);
      set_col(bv.axes, 1, u);
      set_col(bv.axes, 2, v);
      bv.extent = V3(len_p1p2 * 0.5, 0, 0);
      bv.To = (p1 + p2) / 2;
    } break;
    case 3: {  // fit3 :100-123
      const V3 &p1 = ps[0], &p2 = ps[1], &p3 = ps[2];
      V3 e[3] = {p1 - p2, p2 - p3, p3 - p1};
      double len[3] = {sqnorm(e[0]), sqnorm(e[1]), sqnorm(e[2])};
      int imax = 0;
      if (len[1] > len[0]) imax = 1;
      if (len[2] > len[imax]) imax = 2;
      set_col(bv.axes, 2, normalized(cross(e[0], e[1])));
      set_col(bv.axes, 0, normalized(e[imax]));
      set_col(bv.axes, 1, cross(bv.axes.col(2), bv.axes.col(0)));
      PointSrc S{ps, nullptr, nullptr, 3};
      getExtentAndCenter(S, bv.axes, bv.To, bv.extent);
    } break;
    default: {  // fitn :131-143
      M3 M;
      V3 E[3];
      double s[3] = {0, 0, 0};
      PointSrc S{ps, nullptr, nullptr, n};
      getCovariance(S, M);
      eigen(M, s, E);
      axisFromEigen(E, s, bv.axes);
      getExtentAndCenter(S, bv.axes, bv.To, bv.extent);
    }
  }
}
static void fit_points_rss(const V3* ps, unsigned n, RSS& bv) {
  switch (n) {
    case 1:  // RSS_fit_functions::fit1 :147-153
      bv.Tr = ps[0];
      bv.axes = identity3();
      bv.length[0] = bv.length[1] = 0;
      bv.radius = 0;
      break;
    case 2: {  // fit2 :155-168
      const V3 &p1 = ps[0], &p2 = ps[1];
      V3 a0 = p1 - p2;
      double len_p1p2 = norm(a0);
      a0 /= len_p1p2;
      set_col(bv.axes, 0, a0);
      V3 u, v;
      generateCoordinateSystem(a0, u, v);
      set_col(bv.axes, 1, u);
      set_col(bv.axes, 2, v);
      bv.length[0] = len_p1p2;
      bv.length[1] = 0;
      bv.Tr = p2;
      bv.radius = 0;
    } break;
    case 3: {  // fit3 :170-194
      const V3 &p1 = ps[0], &p2 = ps[1], &p3 = ps[2];
      V3 e[3] = {p1 - p2, p2 - p3, p3 - p1};
      double len[3] = {sqnorm(e[0]), sqnorm(e[1]), sqnorm(e[2])};
      int imax = 0;
      if (len[1] > len[0]) imax = 1;
      if (len[2] > len[imax]) imax = 2;
      set_col(bv.axes, 2, normalized(cross(e[0], e[1])));
      set_col(bv.axes, 0, normalized(e[imax]));
      set_col(bv.axes, 1, cross(bv.axes.col(2), bv.axes.col(0)));
      PointSrc S{ps, nullptr, nullptr, 3};
      getRadiusAndOriginAndRectangleSize(S, bv.axes, bv.Tr, bv.length, bv.radius);
    } break;
    default: {  // fitn :203-216
      M3 M;
      V3 E[3];
      double s[3] = {0, 0, 0};
      PointSrc S{ps, nullptr, nullptr, n};
      getCovariance(S, M);
      eigen(M, s, E);
      axisFromEigen(E, s, bv.axes);
      getRadiusAndOriginAndRectangleSize(S, bv.axes, bv.Tr, bv.length, bv.radius);
    }
  }
}

// getBoundVertices (geometric_shapes_utility.cpp:47-262)
static std::vector<V3> getBoundVertices(const Shape& s, const Tf& tf) {
  std::vector<V3> r;
  auto T = [&](double x, double y, double z) { r.push_back(tf.transform(V3(x, y, z))); };
  switch (s.type) {
    case HFB_GEOM_BOX: {
      double a = s.p[0], b = s.p[1], c = s.p[2];
      T(a, b, c); T(a, b, -c); T(a, -b, c); T(a, -b, -c); T(-a, b, c); T(-a, b, -c); T(-a, -b, c); T(-a, -b, -c);
    } break;
    case HFB_GEOM_SPHERE: {
      const double m = (1 + std::sqrt(5.0)) / 2.0;
      double edge_size = s.p[0] * 6 / (std::sqrt(27.0) + std::sqrt(15.0));
      double a = edge_size, b = m * edge_size;
      T(0, a, b); T(0, -a, b); T(0, a, -b); T(0, -a, -b); T(a, b, 0); T(-a, b, 0); T(a, -b, 0); T(-a, -b, 0);
      T(b, 0, a); T(b, 0, -a); T(-b, 0, a); T(-b, 0, -a);
    } break;
    case HFB_GEOM_ELLIPSOID: {
      const double phi = (1 + std::sqrt(5.0)) / 2.0;
      const double a = std::sqrt(3.0) / (phi * phi);
      const double b = phi * a;
      const double A = s.p[0], B = s.p[1], C = s.p[2];
      double Aa = A * a, Ab = A * b, Ba = B * a, Bb = B * b, Ca = C * a, Cb = C * b;
      T(0, Ba, Cb); T(0, -Ba, Cb); T(0, Ba, -Cb); T(0, -Ba, -Cb); T(Aa, Bb, 0); T(-Aa, Bb, 0); T(Aa, -Bb, 0);
      T(-Aa, -Bb, 0); T(Ab, 0, Ca); T(Ab, 0, -Ca); T(-Ab, 0, Ca); T(-Ab, 0, -Ca);
    } break;
    case HFB_GEOM_CAPSULE: {
      const double m = (1 + std::sqrt(5.0)) / 2.0;
      double hl = s.p[1];
      double edge_size = s.p[0] * 6 / (std::sqrt(27.0) + std::sqrt(15.0));
      double a = edge_size, b = m * edge_size;
      double r2 = s.p[0] * 2 / std::sqrt(3.0);
      T(0, a, b + hl); T(0, -a, b + hl); T(0, a, -b + hl); T(0, -a, -b + hl); T(a, b, hl); T(-a, b, hl);
      T(a, -b, hl); T(-a, -b, hl); T(b, 0, a + hl); T(b, 0, -a + hl); T(-b, 0, a + hl); T(-b, 0, -a + hl);
      T(0, a, b - hl); T(0, -a, b - hl); T(0, a, -b - hl); T(0, -a, -b - hl); T(a, b, -hl); T(-a, b, -hl);
      T(a, -b, -hl); T(-a, -b, -hl); T(b, 0, a - hl); T(b, 0, -a - hl); T(-b, 0, a - hl); T(-b, 0, -a - hl);
      double c = 0.5 * r2, d = s.p[0];
      T(r2, 0, hl); T(c, d, hl); T(-c, d, hl); T(-r2, 0, hl); T(-c, -d, hl); T(c, -d, hl);
      T(r2, 0, -hl); T(c, d, -hl); T(-c, d, -hl); T(-r2, 0, -hl); T(-c, -d, -hl); T(c, -d, -hl);
    } break;
    case HFB_GEOM_CONE: {
      double hl = s.p[1], r2 = s.p[0] * 2 / std::sqrt(3.0), a = 0.5 * r2, b = s.p[0];
      T(r2, 0, -hl); T(a, b, -hl); T(-a, b, -hl); T(-r2, 0, -hl); T(-a, -b, -hl); T(a, -b, -hl); T(0, 0, hl);
    } break;
    case HFB_GEOM_CYLINDER: {
      double hl = s.p[1], r2 = s.p[0] * 2 / std::sqrt(3.0), a = 0.5 * r2, b = s.p[0];
      T(r2, 0, -hl); T(a, b, -hl); T(-a, b, -hl); T(-r2, 0, -hl); T(-a, -b, -hl); T(a, -b, -hl);
      T(r2, 0, hl); T(a, b, hl); T(-a, b, hl); T(-r2, 0, hl); T(-a, -b, hl); T(a, -b, hl);
    } break;
    case HFB_GEOM_CONVEX:
      for (const V3& p : s.cvx->points) r.push_back(tf.transform(p));
      break;
    case HFB_GEOM_TRIANGLE:
      for (int k = 0; k < 3; ++k) r.push_back(tf.transform(s.tri[k]));
      break;
    default:
      break;
  }
  return r;
}

void computeBV_OBBRSS(const Shape& s, const Tf& tf, OBBRSS& bv) {  // geometric_shapes_utility.h:73-83
  std::vector<V3> v = getBoundVertices(s, tf);
  fit_points_obb(v.data(), (unsigned)v.size(), bv.obb);
  fit_points_rss(v.data(), (unsigned)v.size(), bv.rss);
}

// computeBV<OBB, S>: the specialisations the plain BVHModel<OBB> walks use for the shape's box
// (geometric_shapes_utility.cpp:458-545); Ellipsoid has none and takes the generic fit of the bound vertices
void computeBV_OBB(const Shape& s, const Tf& tf, OBB& bv) {
  M3 I;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) I.m[r][c] = r == c ? 1.0 : 0.0;
  switch (s.type) {
    case HFB_GEOM_BOX:
      bv.To = tf.T;
      bv.axes = tf.R;
      bv.extent = V3(s.p[0], s.p[1], s.p[2]);
      return;
    case HFB_GEOM_SPHERE:
      bv.To = tf.T;
      bv.axes = I;
      bv.extent = V3(s.p[0], s.p[0], s.p[0]);
      return;
    case HFB_GEOM_CAPSULE:
      bv.To = tf.T;
      bv.axes = tf.R;
      bv.extent = V3(s.p[0], s.p[0], s.p[1] + s.p[0]);
      return;
    case HFB_GEOM_CONE:
    case HFB_GEOM_CYLINDER:
      bv.To = tf.T;
      bv.axes = tf.R;
      bv.extent = V3(s.p[0], s.p[0], s.p[1]);
      return;
    case HFB_GEOM_CONVEX: {
      fit_points_obb(s.cvx->points.data(), (unsigned)s.cvx->points.size(), bv);
      M3 A;  // bv.axes.applyOnTheLeft(R)
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
          A.m[r][c] = (tf.R.m[r][0] * bv.axes.m[0][c] + tf.R.m[r][1] * bv.axes.m[1][c]) + tf.R.m[r][2] * bv.axes.m[2][c];
      bv.axes = A;
      bv.To = mul(tf.R, bv.To) + tf.T;
      return;
    }
    default: {
      std::vector<V3> v = getBoundVertices(s, tf);
      fit_points_obb(v.data(), (unsigned)v.size(), bv);
    }
  }
}

// BVFitter<OBBRSS>::fit (BV_fitter.cpp:501-531): one eigen decomposition shared by both halves
static OBBRSS fit_primitives(const BVHModel& m, const unsigned* prim, unsigned n) {
  OBBRSS bv;
  M3 M;
  V3 E[3];
  double s[3];
  PointSrc S{m.vertices.data(), m.tris.data(), prim, n};
  getCovariance(S, M);
  eigen(M, s, E);
  axisFromEigen(E, s, bv.obb.axes);
  bv.rss.axes = bv.obb.axes;
  getExtentAndCenter(S, bv.obb.axes, bv.obb.To, bv.obb.extent);
  V3 origin;
  double l[2], r;
  getRadiusAndOriginAndRectangleSize(S, bv.rss.axes, origin, l, r);
  bv.rss.Tr = origin;
  bv.rss.length[0] = l[0];
  bv.rss.length[1] = l[1];
  bv.rss.radius = r;
  return bv;
}

// recursiveBuildTree (BVH_model.cpp:892-960), SPLIT_METHOD_MEAN (BV_splitter.cpp:80-118,242-250)
static void recursiveBuildTree(BVHModel& m, int bv_id, unsigned first_primitive, unsigned num_primitives) {
  unsigned* cur = m.primitive_indices.data() + first_primitive;
  OBBRSS bv = fit_primitives(m, cur, num_primitives);
  // computeRule_mean: split_vector = obb.axes.col(0); split_value = c.dot(split_vector) / (3 n)
  V3 split_vector = bv.obb.axes.col(0);
  V3 c(0, 0, 0);
  for (unsigned i = 0; i < num_primitives; ++i) {
    const Tri& t = m.tris[cur[i]];
    c += m.vertices[t.v[0]] + m.vertices[t.v[1]] + m.vertices[t.v[2]];
  }
  double split_value = dot(c, split_vector) / (3 * num_primitives);

  BVNode* node = &m.bvs[bv_id];
  node->bv = bv;
  node->first_primitive = first_primitive;
  node->num_primitives = num_primitives;
  if (num_primitives == 1) {
    node->first_child = -((int)(*cur) + 1);
  } else {
    node->first_child = (int)m.num_bvs;
    m.num_bvs += 2;
    unsigned c1 = 0;
    for (unsigned i = 0; i < num_primitives; ++i) {
      const Tri& t = m.tris[cur[i]];
      V3 p = (m.vertices[t.v[0]] + m.vertices[t.v[1]] + m.vertices[t.v[2]]) / 3.;
      if (dot(split_vector, p) > split_value) {  // BVSplitter<OBBRSS>::apply :275-278
      } else {
        unsigned temp = cur[i];
        cur[i] = cur[c1];
        cur[c1] = temp;
        c1++;
      }
    }
    if ((c1 == 0) || (c1 == num_primitives)) c1 = num_primitives / 2;
    const int left = m.bvs[bv_id].first_child;
    recursiveBuildTree(m, left, first_primitive, c1);
    recursiveBuildTree(m, left + 1, first_primitive + c1, num_primitives - c1);
  }
}

void BVHModel::build() {  // buildTree (BVH_model.cpp:860-890)
  const unsigned n = (unsigned)tris.size();
  bvs.assign(2 * n - 1, BVNode());
  primitive_indices.resize(n);
  for (unsigned i = 0; i < n; ++i) primitive_indices[i] = i;
  num_bvs = 1;
  recursiveBuildTree(*this, 0, 0, n);
}

// --------------------------------------------------------------------- RSS.cpp ---
static inline void clipToRange(double& val, double a, double b) {
  if (val < a) val = a;
  else if (val > b) val = b;
}
static void segCoords(double& t, double& u, double a, double b, double A_dot_B, double A_dot_T, double B_dot_T) {
  double denom = 1 - A_dot_B * A_dot_B;
  if (denom == 0) t = 0;
  else {
    t = (A_dot_T - B_dot_T * A_dot_B) / denom;
    clipToRange(t, 0, a);
  }
  u = t * A_dot_B - B_dot_T;
  if (u < 0) {
    u = 0;
    t = A_dot_T;
    clipToRange(t, 0, a);
  } else if (u > b) {
    u = b;
    t = u * A_dot_B + A_dot_T;
    clipToRange(t, 0, a);
  }
}
static bool inVoronoi(double a, double b, double Anorm_dot_B, double Anorm_dot_T, double A_dot_B, double A_dot_T,
                      double B_dot_T) {
  if (std::fabs(Anorm_dot_B) < 1e-7) return false;
  double t, u, v;
  u = -Anorm_dot_T / Anorm_dot_B;
  clipToRange(u, 0, b);
  t = u * A_dot_B + A_dot_T;
  clipToRange(t, 0, a);
  v = t * A_dot_B - B_dot_T;
  if (Anorm_dot_B > 0) {
    if (v > (u + 1e-7)) return true;
  } else {
    if (v < (u - 1e-7)) return true;
  }
  return false;
}

// rectDistance (RSS.cpp:121-713) without the optional closest points P, Q
#define RAB(i, j) Rab.m[i][j]
double rectDistance(const M3& Rab, const V3& Tab, const double a[2], const double b[2]) {
  double A0_dot_B0 = RAB(0, 0), A0_dot_B1 = RAB(0, 1), A1_dot_B0 = RAB(1, 0), A1_dot_B1 = RAB(1, 1);
  double aA0_dot_B0 = a[0] * A0_dot_B0, aA0_dot_B1 = a[0] * A0_dot_B1, aA1_dot_B0 = a[1] * A1_dot_B0,
         aA1_dot_B1 = a[1] * A1_dot_B1;
  double bA0_dot_B0 = b[0] * A0_dot_B0, bA1_dot_B0 = b[0] * A1_dot_B0, bA0_dot_B1 = b[1] * A0_dot_B1,
         bA1_dot_B1 = b[1] * A1_dot_B1;
  V3 Tba = tmul(Rab, Tab);
  V3 S;
  double t, u;

  double ALL_x, ALU_x, AUL_x, AUU_x, BLL_x, BLU_x, BUL_x, BUU_x;
  double LA1_lx, LA1_ux, UA1_lx, UA1_ux, LB1_lx, LB1_ux, UB1_lx, UB1_ux;
  ALL_x = -Tba[0];
  ALU_x = ALL_x + aA1_dot_B0;
  AUL_x = ALL_x + aA0_dot_B0;
  AUU_x = ALU_x + aA0_dot_B0;
  if (ALL_x < ALU_x) { LA1_lx = ALL_x; LA1_ux = ALU_x; UA1_lx = AUL_x; UA1_ux = AUU_x; }
  else { LA1_lx = ALU_x; LA1_ux = ALL_x; UA1_lx = AUU_x; UA1_ux = AUL_x; }
  BLL_x = Tab[0];
  BLU_x = BLL_x + bA0_dot_B1;
  BUL_x = BLL_x + bA0_dot_B0;
  BUU_x = BLU_x + bA0_dot_B0;
  if (BLL_x < BLU_x) { LB1_lx = BLL_x; LB1_ux = BLU_x; UB1_lx = BUL_x; UB1_ux = BUU_x; }
  else { LB1_lx = BLU_x; LB1_ux = BLL_x; UB1_lx = BUU_x; UB1_ux = BUL_x; }

  // UA1, UB1
  if ((UA1_ux > b[0]) && (UB1_ux > a[0])) {
    if (((UA1_lx > b[0]) || inVoronoi(b[1], a[1], A1_dot_B0, aA0_dot_B0 - b[0] - Tba[0], A1_dot_B1,
                                      aA0_dot_B1 - Tba[1], -Tab[1] - bA1_dot_B0)) &&
        ((UB1_lx > a[0]) || inVoronoi(a[1], b[1], A0_dot_B1, Tab[0] + bA0_dot_B0 - a[0], A1_dot_B1,
                                      Tab[1] + bA1_dot_B0, Tba[1] - aA0_dot_B1))) {
      segCoords(t, u, a[1], b[1], A1_dot_B1, Tab[1] + bA1_dot_B0, Tba[1] - aA0_dot_B1);
      S.x = Tab[0] + RAB(0, 0) * b[0] + RAB(0, 1) * u - a[0];
      S.y = Tab[1] + RAB(1, 0) * b[0] + RAB(1, 1) * u - t;
      S.z = Tab[2] + RAB(2, 0) * b[0] + RAB(2, 1) * u;
      return norm(S);
    }
  }
  // UA1, LB1
  if ((UA1_lx < 0) && (LB1_ux > a[0])) {
    if (((UA1_ux < 0) || inVoronoi(b[1], a[1], -A1_dot_B0, Tba[0] - aA0_dot_B0, A1_dot_B1, aA0_dot_B1 - Tba[1],
                                   -Tab[1])) &&
        ((LB1_lx > a[0]) || inVoronoi(a[1], b[1], A0_dot_B1, Tab[0] - a[0], A1_dot_B1, Tab[1], Tba[1] - aA0_dot_B1))) {
      segCoords(t, u, a[1], b[1], A1_dot_B1, Tab[1], Tba[1] - aA0_dot_B1);
      S.x = Tab[0] + RAB(0, 1) * u - a[0];
      S.y = Tab[1] + RAB(1, 1) * u - t;
      S.z = Tab[2] + RAB(2, 1) * u;
      return norm(S);
    }
  }
  // LA1, UB1
  if ((LA1_ux > b[0]) && (UB1_lx < 0)) {
    if (((LA1_lx > b[0]) || inVoronoi(b[1], a[1], A1_dot_B0, -Tba[0] - b[0], A1_dot_B1, -Tba[1], -Tab[1] - bA1_dot_B0)) &&
        ((UB1_ux < 0) || inVoronoi(a[1], b[1], -A0_dot_B1, -Tab[0] - bA0_dot_B0, A1_dot_B1, Tab[1] + bA1_dot_B0, Tba[1]))) {
      segCoords(t, u, a[1], b[1], A1_dot_B1, Tab[1] + bA1_dot_B0, Tba[1]);
      S.x = Tab[0] + RAB(0, 0) * b[0] + RAB(0, 1) * u;
      S.y = Tab[1] + RAB(1, 0) * b[0] + RAB(1, 1) * u - t;
      S.z = Tab[2] + RAB(2, 0) * b[0] + RAB(2, 1) * u;
      return norm(S);
    }
  }
  // LA1, LB1
  if ((LA1_lx < 0) && (LB1_lx < 0)) {
    if (((LA1_ux < 0) || inVoronoi(b[1], a[1], -A1_dot_B0, Tba[0], A1_dot_B1, -Tba[1], -Tab[1])) &&
        ((LB1_ux < 0) || inVoronoi(a[1], b[1], -A0_dot_B1, -Tab[0], A1_dot_B1, Tab[1], Tba[1]))) {
      segCoords(t, u, a[1], b[1], A1_dot_B1, Tab[1], Tba[1]);
      S.x = Tab[0] + RAB(0, 1) * u;
      S.y = Tab[1] + RAB(1, 1) * u - t;
      S.z = Tab[2] + RAB(2, 1) * u;
      return norm(S);
    }
  }

  double ALL_y, ALU_y, AUL_y, AUU_y;
  ALL_y = -Tba[1];
  ALU_y = ALL_y + aA1_dot_B1;
  AUL_y = ALL_y + aA0_dot_B1;
  AUU_y = ALU_y + aA0_dot_B1;
  double LA1_ly, LA1_uy, UA1_ly, UA1_uy, LB0_lx, LB0_ux, UB0_lx, UB0_ux;
  if (ALL_y < ALU_y) { LA1_ly = ALL_y; LA1_uy = ALU_y; UA1_ly = AUL_y; UA1_uy = AUU_y; }
  else { LA1_ly = ALU_y; LA1_uy = ALL_y; UA1_ly = AUU_y; UA1_uy = AUL_y; }
  if (BLL_x < BUL_x) { LB0_lx = BLL_x; LB0_ux = BUL_x; UB0_lx = BLU_x; UB0_ux = BUU_x; }
  else { LB0_lx = BUL_x; LB0_ux = BLL_x; UB0_lx = BUU_x; UB0_ux = BLU_x; }

  // UA1, UB0
  if ((UA1_uy > b[1]) && (UB0_ux > a[0])) {
    if (((UA1_ly > b[1]) || inVoronoi(b[0], a[1], A1_dot_B1, aA0_dot_B1 - Tba[1] - b[1], A1_dot_B0,
                                      aA0_dot_B0 - Tba[0], -Tab[1] - bA1_dot_B1)) &&
        ((UB0_lx > a[0]) || inVoronoi(a[1], b[0], A0_dot_B0, Tab[0] - a[0] + bA0_dot_B1, A1_dot_B0,
                                      Tab[1] + bA1_dot_B1, Tba[0] - aA0_dot_B0))) {
      segCoords(t, u, a[1], b[0], A1_dot_B0, Tab[1] + bA1_dot_B1, Tba[0] - aA0_dot_B0);
      S.x = Tab[0] + RAB(0, 1) * b[1] + RAB(0, 0) * u - a[0];
      S.y = Tab[1] + RAB(1, 1) * b[1] + RAB(1, 0) * u - t;
      S.z = Tab[2] + RAB(2, 1) * b[1] + RAB(2, 0) * u;
      return norm(S);
    }
  }
  // UA1, LB0
  if ((UA1_ly < 0) && (LB0_ux > a[0])) {
    if (((UA1_uy < 0) || inVoronoi(b[0], a[1], -A1_dot_B1, Tba[1] - aA0_dot_B1, A1_dot_B0, aA0_dot_B0 - Tba[0],
                                   -Tab[1])) &&
        ((LB0_lx > a[0]) || inVoronoi(a[1], b[0], A0_dot_B0, Tab[0] - a[0], A1_dot_B0, Tab[1], Tba[0] - aA0_dot_B0))) {
      segCoords(t, u, a[1], b[0], A1_dot_B0, Tab[1], Tba[0] - aA0_dot_B0);
      S.x = Tab[0] + RAB(0, 0) * u - a[0];
      S.y = Tab[1] + RAB(1, 0) * u - t;
      S.z = Tab[2] + RAB(2, 0) * u;
      return norm(S);
    }
  }
  // LA1, UB0
  if ((LA1_uy > b[1]) && (UB0_lx < 0)) {
    if (((LA1_ly > b[1]) || inVoronoi(b[0], a[1], A1_dot_B1, -Tba[1] - b[1], A1_dot_B0, -Tba[0], -Tab[1] - bA1_dot_B1)) &&
        ((UB0_ux < 0) || inVoronoi(a[1], b[0], -A0_dot_B0, -Tab[0] - bA0_dot_B1, A1_dot_B0, Tab[1] + bA1_dot_B1, Tba[0]))) {
      segCoords(t, u, a[1], b[0], A1_dot_B0, Tab[1] + bA1_dot_B1, Tba[0]);
      S.x = Tab[0] + RAB(0, 1) * b[1] + RAB(0, 0) * u;
      S.y = Tab[1] + RAB(1, 1) * b[1] + RAB(1, 0) * u - t;
      S.z = Tab[2] + RAB(2, 1) * b[1] + RAB(2, 0) * u;
      return norm(S);
    }
  }
  // LA1, LB0
  if ((LA1_ly < 0) && (LB0_lx < 0)) {
    if (((LA1_uy < 0) || inVoronoi(b[0], a[1], -A1_dot_B1, Tba[1], A1_dot_B0, -Tba[0], -Tab[1])) &&
        ((LB0_ux < 0) || inVoronoi(a[1], b[0], -A0_dot_B0, -Tab[0], A1_dot_B0, Tab[1], Tba[0]))) {
      segCoords(t, u, a[1], b[0], A1_dot_B0, Tab[1], Tba[0]);
      S.x = Tab[0] + RAB(0, 0) * u;
      S.y = Tab[1] + RAB(1, 0) * u - t;
      S.z = Tab[2] + RAB(2, 0) * u;
      return norm(S);
    }
  }

  double BLL_y, BLU_y, BUL_y, BUU_y;
  BLL_y = Tab[1];
  BLU_y = BLL_y + bA1_dot_B1;
  BUL_y = BLL_y + bA1_dot_B0;
  BUU_y = BLU_y + bA1_dot_B0;
  double LA0_lx, LA0_ux, UA0_lx, UA0_ux, LB1_ly, LB1_uy, UB1_ly, UB1_uy;
  if (ALL_x < AUL_x) { LA0_lx = ALL_x; LA0_ux = AUL_x; UA0_lx = ALU_x; UA0_ux = AUU_x; }
  else { LA0_lx = AUL_x; LA0_ux = ALL_x; UA0_lx = AUU_x; UA0_ux = ALU_x; }
  if (BLL_y < BLU_y) { LB1_ly = BLL_y; LB1_uy = BLU_y; UB1_ly = BUL_y; UB1_uy = BUU_y; }
  else { LB1_ly = BLU_y; LB1_uy = BLL_y; UB1_ly = BUU_y; UB1_uy = BUL_y; }

  // UA0, UB1
  if ((UA0_ux > b[0]) && (UB1_uy > a[1])) {
    if (((UA0_lx > b[0]) || inVoronoi(b[1], a[0], A0_dot_B0, aA1_dot_B0 - Tba[0] - b[0], A0_dot_B1,
                                      aA1_dot_B1 - Tba[1], -Tab[0] - bA0_dot_B0)) &&
        ((UB1_ly > a[1]) || inVoronoi(a[0], b[1], A1_dot_B1, Tab[1] - a[1] + bA1_dot_B0, A0_dot_B1,
                                      Tab[0] + bA0_dot_B0, Tba[1] - aA1_dot_B1))) {
      segCoords(t, u, a[0], b[1], A0_dot_B1, Tab[0] + bA0_dot_B0, Tba[1] - aA1_dot_B1);
      S.x = Tab[0] + RAB(0, 0) * b[0] + RAB(0, 1) * u - t;
      S.y = Tab[1] + RAB(1, 0) * b[0] + RAB(1, 1) * u - a[1];
      S.z = Tab[2] + RAB(2, 0) * b[0] + RAB(2, 1) * u;
      return norm(S);
    }
  }
  // UA0, LB1
  if ((UA0_lx < 0) && (LB1_uy > a[1])) {
    if (((UA0_ux < 0) || inVoronoi(b[1], a[0], -A0_dot_B0, Tba[0] - aA1_dot_B0, A0_dot_B1, aA1_dot_B1 - Tba[1],
                                   -Tab[0])) &&
        ((LB1_ly > a[1]) || inVoronoi(a[0], b[1], A1_dot_B1, Tab[1] - a[1], A0_dot_B1, Tab[0], Tba[1] - aA1_dot_B1))) {
      segCoords(t, u, a[0], b[1], A0_dot_B1, Tab[0], Tba[1] - aA1_dot_B1);
      S.x = Tab[0] + RAB(0, 1) * u - t;
      S.y = Tab[1] + RAB(1, 1) * u - a[1];
      S.z = Tab[2] + RAB(2, 1) * u;
      return norm(S);
    }
  }
  // LA0, UB1
  if ((LA0_ux > b[0]) && (UB1_ly < 0)) {
    if (((LA0_lx > b[0]) || inVoronoi(b[1], a[0], A0_dot_B0, -b[0] - Tba[0], A0_dot_B1, -Tba[1], -bA0_dot_B0 - Tab[0])) &&
        ((UB1_uy < 0) || inVoronoi(a[0], b[1], -A1_dot_B1, -Tab[1] - bA1_dot_B0, A0_dot_B1, Tab[0] + bA0_dot_B0, Tba[1]))) {
      segCoords(t, u, a[0], b[1], A0_dot_B1, Tab[0] + bA0_dot_B0, Tba[1]);
      S.x = Tab[0] + RAB(0, 0) * b[0] + RAB(0, 1) * u - t;
      S.y = Tab[1] + RAB(1, 0) * b[0] + RAB(1, 1) * u;
      S.z = Tab[2] + RAB(2, 0) * b[0] + RAB(2, 1) * u;
      return norm(S);
    }
  }
  // LA0, LB1
  if ((LA0_lx < 0) && (LB1_ly < 0)) {
    if (((LA0_ux < 0) || inVoronoi(b[1], a[0], -A0_dot_B0, Tba[0], A0_dot_B1, -Tba[1], -Tab[0])) &&
        ((LB1_uy < 0) || inVoronoi(a[0], b[1], -A1_dot_B1, -Tab[1], A0_dot_B1, Tab[0], Tba[1]))) {
      segCoords(t, u, a[0], b[1], A0_dot_B1, Tab[0], Tba[1]);
      S.x = Tab[0] + RAB(0, 1) * u - t;
      S.y = Tab[1] + RAB(1, 1) * u;
      S.z = Tab[2] + RAB(2, 1) * u;
      return norm(S);
    }
  }

  double LA0_ly, LA0_uy, UA0_ly, UA0_uy, LB0_ly, LB0_uy, UB0_ly, UB0_uy;
  if (ALL_y < AUL_y) { LA0_ly = ALL_y; LA0_uy = AUL_y; UA0_ly = ALU_y; UA0_uy = AUU_y; }
  else { LA0_ly = AUL_y; LA0_uy = ALL_y; UA0_ly = AUU_y; UA0_uy = ALU_y; }
  if (BLL_y < BUL_y) { LB0_ly = BLL_y; LB0_uy = BUL_y; UB0_ly = BLU_y; UB0_uy = BUU_y; }
  else { LB0_ly = BUL_y; LB0_uy = BLL_y; UB0_ly = BUU_y; UB0_uy = BLU_y; }

  // UA0, UB0
  if ((UA0_uy > b[1]) && (UB0_uy > a[1])) {
    if (((UA0_ly > b[1]) || inVoronoi(b[0], a[0], A0_dot_B1, aA1_dot_B1 - Tba[1] - b[1], A0_dot_B0,
                                      aA1_dot_B0 - Tba[0], -Tab[0] - bA0_dot_B1)) &&
        ((UB0_ly > a[1]) || inVoronoi(a[0], b[0], A1_dot_B0, Tab[1] - a[1] + bA1_dot_B1, A0_dot_B0,
                                      Tab[0] + bA0_dot_B1, Tba[0] - aA1_dot_B0))) {
      segCoords(t, u, a[0], b[0], A0_dot_B0, Tab[0] + bA0_dot_B1, Tba[0] - aA1_dot_B0);
      S.x = Tab[0] + RAB(0, 1) * b[1] + RAB(0, 0) * u - t;
      S.y = Tab[1] + RAB(1, 1) * b[1] + RAB(1, 0) * u - a[1];
      S.z = Tab[2] + RAB(2, 1) * b[1] + RAB(2, 0) * u;
      return norm(S);
    }
  }
  // UA0, LB0
  if ((UA0_ly < 0) && (LB0_uy > a[1])) {
    if (((UA0_uy < 0) || inVoronoi(b[0], a[0], -A0_dot_B1, Tba[1] - aA1_dot_B1, A0_dot_B0, aA1_dot_B0 - Tba[0],
                                   -Tab[0])) &&
        ((LB0_ly > a[1]) || inVoronoi(a[0], b[0], A1_dot_B0, Tab[1] - a[1], A0_dot_B0, Tab[0], Tba[0] - aA1_dot_B0))) {
      segCoords(t, u, a[0], b[0], A0_dot_B0, Tab[0], Tba[0] - aA1_dot_B0);
      S.x = Tab[0] + RAB(0, 0) * u - t;
      S.y = Tab[1] + RAB(1, 0) * u - a[1];
      S.z = Tab[2] + RAB(2, 0) * u;
      return norm(S);
    }
  }
  // LA0, UB0
  if ((LA0_uy > b[1]) && (UB0_ly < 0)) {
    if (((LA0_ly > b[1]) || inVoronoi(b[0], a[0], A0_dot_B1, -Tba[1] - b[1], A0_dot_B0, -Tba[0], -Tab[0] - bA0_dot_B1)) &&
        ((UB0_uy < 0) || inVoronoi(a[0], b[0], -A1_dot_B0, -Tab[1] - bA1_dot_B1, A0_dot_B0, Tab[0] + bA0_dot_B1, Tba[0]))) {
      segCoords(t, u, a[0], b[0], A0_dot_B0, Tab[0] + bA0_dot_B1, Tba[0]);
      S.x = Tab[0] + RAB(0, 1) * b[1] + RAB(0, 0) * u - t;
      S.y = Tab[1] + RAB(1, 1) * b[1] + RAB(1, 0) * u;
      S.z = Tab[2] + RAB(2, 1) * b[1] + RAB(2, 0) * u;
      return norm(S);
    }
  }
  // LA0, LB0
  if ((LA0_ly < 0) && (LB0_ly < 0)) {
    if (((LA0_uy < 0) || inVoronoi(b[0], a[0], -A0_dot_B1, Tba[1], A0_dot_B0, -Tba[0], -Tab[0])) &&
        ((LB0_uy < 0) || inVoronoi(a[0], b[0], -A1_dot_B0, -Tab[1], A0_dot_B0, Tab[0], Tba[0]))) {
      segCoords(t, u, a[0], b[0], A0_dot_B0, Tab[0], Tba[0]);
      S.x = Tab[0] + RAB(0, 0) * u - t;
      S.y = Tab[1] + RAB(1, 0) * u;
      S.z = Tab[2] + RAB(2, 0) * u;
      return norm(S);
    }
  }

  // no edges passed, take max separation along face normals
  double sep1, sep2;
  if (Tab[2] > 0.0) {
    sep1 = Tab[2];
    if (RAB(2, 0) < 0.0) sep1 += b[0] * RAB(2, 0);
    if (RAB(2, 1) < 0.0) sep1 += b[1] * RAB(2, 1);
  } else {
    sep1 = -Tab[2];
    if (RAB(2, 0) > 0.0) sep1 -= b[0] * RAB(2, 0);
    if (RAB(2, 1) > 0.0) sep1 -= b[1] * RAB(2, 1);
  }
  if (Tba[2] < 0) {
    sep2 = -Tba[2];
    if (RAB(0, 2) < 0.0) sep2 += a[0] * RAB(0, 2);
    if (RAB(1, 2) < 0.0) sep2 += a[1] * RAB(1, 2);
  } else {
    sep2 = Tba[2];
    if (RAB(0, 2) > 0.0) sep2 -= a[0] * RAB(0, 2);
    if (RAB(1, 2) > 0.0) sep2 -= a[1] * RAB(1, 2);
  }
  double sep = (sep1 > sep2 ? sep1 : sep2);
  return (sep > 0 ? sep : 0);
}
#undef RAB

// distance(R0, T0, b1, b2) (RSS.cpp:995-1005): b1 in configuration (R0,T0)?  NB the reference's
// doc comment and its arithmetic disagree; the arithmetic is restated: R = b1.axes^T R0 b2.axes,
// T = b1.axes^T (R0 b2.Tr + T0 - b1.Tr)
double rss_distance(const M3& R0, const V3& T0, const RSS& b1, const RSS& b2) {
  M3 R = mul(tmul(b1.axes, R0), b2.axes);
  V3 Ttemp = mul(R0, b2.Tr) + T0 - b1.Tr;
  V3 T = tmul(b1.axes, Ttemp);
  double dist = rectDistance(R, T, b1.length, b2.length);
  dist -= (b1.radius + b2.radius);
  return (dist < 0.0) ? 0.0 : dist;
}

// --------------------------------------------------------------------- OBB.cpp ---
// obbDisjointAndLowerBoundDistance (OBB.cpp:290-393)
static bool obbDisjointAndLowerBoundDistance(const M3& B, const V3& T, const V3& a_, const V3& b_,
                                             double security_margin, double break_distance,
                                             double& squaredLowerBoundDistance) {
  const double breakDistance2 = break_distance * break_distance;
  M3 Bf;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Bf.m[i][j] = std::fabs(B.m[i][j]);
  const double hm = security_margin / 2;
  const V3 a(std::fmax(a_.x + hm, 0.0), std::fmax(a_.y + hm, 0.0), std::fmax(a_.z + hm, 0.0));
  const V3 b(std::fmax(b_.x + hm, 0.0), std::fmax(b_.y + hm, 0.0), std::fmax(b_.z + hm, 0.0));
  {  // obbDisjoint_check_A_axis
    V3 corner(std::fabs(T.x) - a.x, std::fabs(T.y) - a.y, std::fabs(T.z) - a.z);
    corner -= mul(Bf, b);
    V3 c(std::fmax(corner.x, 0.0), std::fmax(corner.y, 0.0), std::fmax(corner.z, 0.0));
    squaredLowerBoundDistance = sqnorm(c);
  }
  if (squaredLowerBoundDistance > breakDistance2) return true;
  {  // obbDisjoint_check_B_axis
    double s, t = 0;
    s = std::fabs(dot(B.col(0), T)) - dot(Bf.col(0), a) - b[0];
    if (s > 0) t += s * s;
    s = std::fabs(dot(B.col(1), T)) - dot(Bf.col(1), a) - b[1];
    if (s > 0) t += s * s;
    s = std::fabs(dot(B.col(2), T)) - dot(Bf.col(2), a) - b[2];
    if (s > 0) t += s * s;
    squaredLowerBoundDistance = t;
  }
  if (squaredLowerBoundDistance > breakDistance2) return true;
  int ja = 1, ka = 2;
  for (int ia = 0; ia < 3; ++ia) {
    for (int ib = 0; ib < 3; ++ib) {  // obbDisjoint_check_Ai_cross_Bi<ib>
      const int jb = (ib + 1) % 3, kb = (ib + 2) % 3;
      double sinus2 = 1 - Bf.m[ia][ib] * Bf.m[ia][ib];
      if (sinus2 < 1e-6) continue;
      const double s = T[ka] * B.m[ja][ib] - T[ja] * B.m[ka][ib];
      const double diff = std::fabs(s) - (a[ja] * Bf.m[ka][ib] + a[ka] * Bf.m[ja][ib] + b[jb] * Bf.m[ia][kb] +
                                          b[kb] * Bf.m[ia][jb]);
      if (diff > 0) {
        squaredLowerBoundDistance = diff * diff / sinus2;
        if (squaredLowerBoundDistance > breakDistance2) return true;
      }
    }
    ja = ka;
    ka = ia;
  }
  return false;
}

// overlap(R0, T0, b1, b2, request, sqrDistLowerBound) (OBB.cpp:475-483)
bool obb_overlap(const M3& R0, const V3& T0, const OBB& b1, const OBB& b2, double security_margin,
                 double break_distance, double& sqrDistLowerBound) {
  V3 Ttemp = tmul(R0, b2.To - T0) - b1.To;
  V3 T = tmul(b1.axes, Ttemp);
  // b1.axes^T * R0^T * b2.axes
  M3 R0t;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R0t.m[i][j] = R0.m[j][i];
  M3 R = mul(tmul(b1.axes, R0t), b2.axes);
  return !obbDisjointAndLowerBoundDistance(R, T, b1.extent, b2.extent, security_margin, break_distance,
                                           sqrDistLowerBound);
}

// ------------------------------------------------------------------ traversal ----
struct DistNode {
  const BVHModel* model1;
  const Shape* model2;
  Tf tf1, tf2;
  OBBRSS model2_bv;
  GJKSolver* solver;
  bool signed_distance;
  double rel_err, abs_err;
  // result
  double min_distance;
  V3 p1, p2, normal;
  int b1;
  uint64_t num_bv_tests = 0, num_leaf_tests = 0;
  unsigned last_status = 0, last_iters = 0;
  bool last_closed = false;
};

// GJKSolver::getGJKInitialGuess with BoundingVolumeGuess requires aabb_local of both shapes
// (narrowphase.h:368-377); the TriangleP a leaf builds on the fly never had computeLocalAABB called, so the
// reference throws std::logic_error at the first leaf that reaches GJK (sphere partners use the closed form)
static inline void leafGuessCheck(const GJKSolver& solver, const Shape& partner) {
  if (solver.gjk_initial_guess == HFB_GUESS_BOUNDING_VOLUME && partner.type != HFB_GEOM_SPHERE)
    throw std::logic_error("computeLocalAABB must have been called on the shapes before using BoundingVolumeGuess");
}

static void leafDistance(DistNode& n, int primitive_id) {  // traversal_node_bvh_shape.h:342-364
  const Tri& t = n.model1->tris[primitive_id];
  Shape tri;
  tri.type = HFB_GEOM_TRIANGLE;
  tri.tri[0] = n.model1->vertices[t.v[0]];
  tri.tri[1] = n.model1->vertices[t.v[1]];
  tri.tri[2] = n.model1->vertices[t.v[2]];
  V3 p1, p2, normal;
  double distance;
  bool closed = false;
  leafGuessCheck(*n.solver, *n.model2);
  shapeShapeDistance(tri, n.tf1, *n.model2, n.tf2, *n.solver, n.signed_distance, distance, p1, p2, normal, closed);
  n.num_leaf_tests++;
  if (n.min_distance > distance) {  // DistanceResult::update (collision_data.h:1111-1124)
    n.min_distance = distance;
    n.b1 = primitive_id;
    n.p1 = p1;
    n.p2 = p2;
    n.normal = normal;
    n.last_closed = closed;
  }
}
static inline bool canStop(const DistNode& n, double c) {  // :322-327
  return (c >= n.min_distance - n.abs_err) && (c * (1 + n.rel_err) >= n.min_distance);
}
static inline double bvLowerBound(DistNode& n, unsigned b1) {  // :465-469
  n.num_bv_tests++;
  return rss_distance(n.tf1.R, n.tf1.T, n.model2_bv.rss, n.model1->bvs[b1].bv.rss);
}
static void distanceRecurse(DistNode& n, unsigned b1) {  // traversal_recurse.cpp:153-203
  const BVNode& node = n.model1->bvs[b1];
  if (node.first_child < 0) {
    leafDistance(n, -(node.first_child + 1));
    return;
  }
  const unsigned a1 = (unsigned)node.first_child, c1 = (unsigned)node.first_child + 1;
  double d1 = bvLowerBound(n, a1);
  double d2 = bvLowerBound(n, c1);
  if (d2 < d1) {
    if (!canStop(n, d2)) distanceRecurse(n, c1);
    if (!canStop(n, d1)) distanceRecurse(n, a1);
  } else {
    if (!canStop(n, d1)) distanceRecurse(n, a1);
    if (!canStop(n, d2)) distanceRecurse(n, c1);
  }
}

// orientedBVHShapeDistance (distance_func_matrix.cpp:109-127) on a fresh DistanceResult
void bvhShapeDistance(const BVHModel& m, const Tf& tf1, const Shape& s, const Tf& tf2, GJKSolver& solver,
                      bool signed_distance, double rel_err, double abs_err, BvhQueryResult& out) {
  DistNode n;
  n.model1 = &m;
  n.model2 = &s;
  n.tf1 = tf1;
  n.tf2 = tf2;
  n.solver = &solver;
  n.signed_distance = signed_distance;
  n.rel_err = rel_err;
  n.abs_err = abs_err;
  n.min_distance = std::numeric_limits<double>::max();
  n.p1 = n.p2 = n.normal = nan3();
  n.b1 = -1;
  computeBV_OBBRSS(s, tf2, n.model2_bv);  // traversal_node_setup.h:765
  leafDistance(n, 0);                      // preprocess(): triangle 0 (:457-461)
  n.num_leaf_tests = 0;                    // the seed is not counted by enable_statistics
  distanceRecurse(n, 0);
  out.distance = n.min_distance;
  out.p1 = n.p1;
  out.p2 = n.p2;
  out.normal = n.normal;
  out.b1 = n.b1;
  out.num_bv_tests = n.num_bv_tests;
  out.num_leaf_tests = n.num_leaf_tests;
}

// ---- collision: MeshShapeCollisionTraversalNodeOBBRSS + collisionRecurse ----------------------
struct ColNode {
  const BVHModel* model1;
  const Shape* model2;
  Tf tf1, tf2;
  OBBRSS model2_bv;
  GJKSolver* solver;
  const hfb_collision_request* req;
  BvhCollideResult* res;
};

static void leafCollides(ColNode& n, unsigned b1, double& sqrDistLowerBound) {  // traversal_node_bvh_shape.h:139-188
  const BVNode& node = n.model1->bvs[b1];
  const int primitive_id = -(node.first_child + 1);
  const Tri& t = n.model1->tris[primitive_id];
  Shape tri;
  tri.type = HFB_GEOM_TRIANGLE;
  tri.tri[0] = n.model1->vertices[t.v[0]];
  tri.tri[1] = n.model1->vertices[t.v[1]];
  tri.tri[2] = n.model1->vertices[t.v[2]];
  const bool compute_penetration = n.req->enable_contact || (n.req->security_margin < 0);
  V3 c1, c2, normal;
  double distance;
  bool closed;
  leafGuessCheck(*n.solver, *n.model2);
  shapeShapeDistance(tri, n.tf1, *n.model2, n.tf2, *n.solver, compute_penetration, distance, c1, c2, normal, closed);
  n.res->num_leaf_tests++;
  const double distToCollision = distance - n.req->security_margin;
  BvhCollideResult& r = *n.res;
  if (distToCollision < r.distance_lower_bound) {  // updateDistanceLowerBoundFromLeaf
    r.distance_lower_bound = distToCollision;
    r.lb_p1 = c1;
    r.lb_p2 = c2;
    r.lb_normal = normal;
  }
  if (distToCollision <= n.req->q.collision_distance_threshold) {
    sqrDistLowerBound = 0;
    if (r.contacts.size() < n.req->num_max_contacts) {
      BvhContact c;
      c.b1 = primitive_id;
      c.p1 = c1;
      c.p2 = c2;
      c.normal = normal;
      c.distance = distance;
      r.contacts.push_back(c);
    }
  } else {
    sqrDistLowerBound = distToCollision * distToCollision;
  }
}

static bool bvDisjoints(ColNode& n, unsigned b1, double& sqrDistLowerBound) {  // :120-136
  n.res->num_bv_tests++;
  bool disjoint = !obb_overlap(n.tf1.R, n.tf1.T, n.model1->bvs[b1].bv.obb, n.model2_bv.obb, n.req->security_margin,
                               n.req->break_distance, sqrDistLowerBound);
  if (disjoint) {  // updateDistanceLowerBoundFromBV (collision_data.h:1177-1184)
    BvhCollideResult& r = *n.res;
    if (r.distance_lower_bound > 0) {
      double new_dlb = std::sqrt(sqrDistLowerBound);
      if (new_dlb < r.distance_lower_bound) r.distance_lower_bound = new_dlb;
    }
  }
  return disjoint;
}

static inline bool colCanStop(const ColNode& n) {  // request.isSatisfied(result)
  return !n.res->contacts.empty() && (n.req->num_max_contacts <= n.res->contacts.size());
}

static void collisionRecurse(ColNode& n, unsigned b1, double& sqrDistLowerBound) {  // traversal_recurse.cpp:44-85
  double sqrDistLowerBound1 = 0, sqrDistLowerBound2 = 0;
  const BVNode& node = n.model1->bvs[b1];
  if (node.first_child < 0) {
    leafCollides(n, b1, sqrDistLowerBound);
    return;
  }
  if (bvDisjoints(n, b1, sqrDistLowerBound)) return;
  const unsigned c1 = (unsigned)node.first_child, c2 = c1 + 1;
  collisionRecurse(n, c1, sqrDistLowerBound1);
  if (colCanStop(n)) return;  // front_list == NULL
  collisionRecurse(n, c2, sqrDistLowerBound2);
  sqrDistLowerBound = std::min(sqrDistLowerBound1, sqrDistLowerBound2);
}

// BVHShapeCollider<OBBRSS,S>::oriented (collision_func_matrix.cpp:141-155) + collide(node)
// (collision_node.cpp:64-79) on a fresh CollisionResult
void bvhShapeCollide(const BVHModel& m, const Tf& tf1, const Shape& s, const Tf& tf2, GJKSolver& solver,
                     const hfb_collision_request& req, BvhCollideResult& out, bool plain_obb) {
  ColNode n;
  n.model1 = &m;
  n.model2 = &s;
  n.tf1 = tf1;
  n.tf2 = tf2;
  n.solver = &solver;
  n.req = &req;
  n.res = &out;
  if (plain_obb) computeBV_OBB(s, tf2, n.model2_bv.obb);  // MeshShapeCollisionTraversalNodeOBB: computeBV<OBB, S>
  else computeBV_OBBRSS(s, tf2, n.model2_bv);  // traversal_node_setup.h:655-694
  double sqrDistLowerBound = 0;
  collisionRecurse(n, 0, sqrDistLowerBound);
}


// ===================== mesh-mesh: MeshDistanceTraversalNodeOBBRSS / MeshCollisionTraversalNodeOBBRSS ==========

// TriangleDistance::segPoints (src/intersect.cpp:60-153)
static void segPoints(const V3& P, const V3& A, const V3& Q, const V3& B, V3& VEC, V3& X, V3& Y) {
  V3 T = Q - P;
  const double A_dot_A = dot(A, A), B_dot_B = dot(B, B), A_dot_B = dot(A, B);
  const double A_dot_T = dot(A, T), B_dot_T = dot(B, T);
  const double denom = A_dot_A * B_dot_B - A_dot_B * A_dot_B;
  double t = (A_dot_T * B_dot_B - B_dot_T * A_dot_B) / denom;
  if ((t < 0) || std::isnan(t)) t = 0;
  else if (t > 1) t = 1;
  const double u = (t * A_dot_B - B_dot_T) / B_dot_B;
  if ((u <= 0) || std::isnan(u)) {
    Y = Q;
    t = A_dot_T / A_dot_A;
    if ((t <= 0) || std::isnan(t)) {
      X = P;
      VEC = Q - P;
    } else if (t >= 1) {
      X = P + A;
      VEC = Q - X;
    } else {
      X = P + A * t;
      VEC = cross(A, cross(T, A));
    }
  } else if (u >= 1) {
    Y = Q + B;
    t = (A_dot_B + A_dot_T) / A_dot_A;
    if ((t <= 0) || std::isnan(t)) {
      X = P;
      VEC = Y - P;
    } else if (t >= 1) {
      X = P + A;
      VEC = Y - X;
    } else {
      X = P + A * t;
      T = Y - P;
      VEC = cross(A, cross(T, A));
    }
  } else {
    Y = Q + B * u;
    if ((t <= 0) || std::isnan(t)) {
      X = P;
      VEC = cross(B, cross(T, B));
    } else if (t >= 1) {
      X = P + A;
      T = Q - X;
      VEC = cross(B, cross(T, B));
    } else {
      X = P + A * t;
      VEC = cross(A, B);
      if (dot(VEC, T) < 0) VEC = VEC * (-1.0);
    }
  }
}

// TriangleDistance::sqrTriDistance (src/intersect.cpp:156-357).  P and Q are written by every segPoints
// call, so on the "triangles overlap" exit (return 0) they hold the last edge pair's points.
double sqrTriDistance(const V3 S[3], const V3 T[3], V3& P, V3& Q) {
  V3 Sv[3], Tv[3], VEC;
  Sv[0] = S[1] - S[0];
  Sv[1] = S[2] - S[1];
  Sv[2] = S[0] - S[2];
  Tv[0] = T[1] - T[0];
  Tv[1] = T[2] - T[1];
  Tv[2] = T[0] - T[2];
  V3 V, Z, minP = nan3(), minQ = nan3();
  int shown_disjoint = 0;
  double mindd = sqnorm(S[0] - T[0]) + 1;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {
      segPoints(S[i], Sv[i], T[j], Tv[j], VEC, P, Q);
      V = Q - P;
      const double dd = dot(V, V);
      if (dd <= mindd) {
        minP = P;
        minQ = Q;
        mindd = dd;
        Z = S[(i + 2) % 3] - P;
        double a = dot(Z, VEC);
        Z = T[(j + 2) % 3] - Q;
        double b = dot(Z, VEC);
        if ((a <= 0) && (b >= 0)) return dd;
        const double p = dot(V, VEC);
        if (a < 0) a = 0;
        if (b > 0) b = 0;
        if ((p - a + b) > 0) shown_disjoint = 1;
      }
    }
  }
  const V3 Sn = cross(Sv[0], Sv[1]);
  const double Snl = dot(Sn, Sn);
  if (Snl > 1e-15) {
    double Tp[3];
    Tp[0] = dot(S[0] - T[0], Sn);
    Tp[1] = dot(S[0] - T[1], Sn);
    Tp[2] = dot(S[0] - T[2], Sn);
    int point = -1;
    if ((Tp[0] > 0) && (Tp[1] > 0) && (Tp[2] > 0)) {
      point = (Tp[0] < Tp[1]) ? 0 : 1;
      if (Tp[2] < Tp[point]) point = 2;
    } else if ((Tp[0] < 0) && (Tp[1] < 0) && (Tp[2] < 0)) {
      point = (Tp[0] > Tp[1]) ? 0 : 1;
      if (Tp[2] > Tp[point]) point = 2;
    }
    if (point >= 0) {
      shown_disjoint = 1;
      if (dot(T[point] - S[0], cross(Sn, Sv[0])) > 0 && dot(T[point] - S[1], cross(Sn, Sv[1])) > 0 &&
          dot(T[point] - S[2], cross(Sn, Sv[2])) > 0) {
        P = T[point] + Sn * (Tp[point] / Snl);
        Q = T[point];
        return sqnorm(P - Q);
      }
    }
  }
  const V3 Tn = cross(Tv[0], Tv[1]);
  const double Tnl = dot(Tn, Tn);
  if (Tnl > 1e-15) {
    double Sp[3];
    Sp[0] = dot(T[0] - S[0], Tn);
    Sp[1] = dot(T[0] - S[1], Tn);
    Sp[2] = dot(T[0] - S[2], Tn);
    int point = -1;
    if ((Sp[0] > 0) && (Sp[1] > 0) && (Sp[2] > 0)) {
      point = (Sp[0] < Sp[1]) ? 0 : 1;
      if (Sp[2] < Sp[point]) point = 2;
    } else if ((Sp[0] < 0) && (Sp[1] < 0) && (Sp[2] < 0)) {
      point = (Sp[0] > Sp[1]) ? 0 : 1;
      if (Sp[2] > Sp[point]) point = 2;
    }
    if (point >= 0) {
      shown_disjoint = 1;
      if (dot(S[point] - T[0], cross(Tn, Tv[0])) > 0 && dot(S[point] - T[1], cross(Tn, Tv[1])) > 0 &&
          dot(S[point] - T[2], cross(Tn, Tv[2])) > 0) {
        P = S[point];
        Q = S[point] + Tn * (Sp[point] / Tnl);
        return sqnorm(P - Q);
      }
    }
  }
  if (shown_disjoint) {
    P = minP;
    Q = minQ;
    return mindd;
  }
  return 0;
}

struct MMDistNode {
  const BVHModel *model1, *model2;
  Tf tf1;
  M3 R;  // RT: model 2 in the frame of model 1 (internal/tools.h:91-99)
  V3 T;
  double rel_err, abs_err;
  double min_distance;
  V3 p1, p2;
  int b1, b2;
  uint64_t num_bv_tests = 0, num_leaf_tests = 0;
};
static inline double obbSize(const OBBRSS& bv) { return sqnorm(bv.obb.extent); }  // OBBRSS.h:114, OBB.h:110
// BVHDistanceTraversalNode::firstOverSecond (traversal_node_bvhs.h:333-344, same rule as :87-98)
template <class N>
static inline bool firstOverSecond(const N& n, unsigned b1, unsigned b2) {
  const BVNode& n1 = n.model1->bvs[b1];
  const BVNode& n2 = n.model2->bvs[b2];
  const double sz1 = obbSize(n1.bv), sz2 = obbSize(n2.bv);
  const bool l1 = n1.first_child < 0, l2 = n2.first_child < 0;
  return l2 || (!l1 && (sz1 > sz2));
}
static void mmTriDistance(MMDistNode& n, int id1, int id2, bool count) {  // :437-470 (and :487-510 for the seed)
  const Tri& t1 = n.model1->tris[id1];
  const Tri& t2 = n.model2->tris[id2];
  V3 S[3], Tt[3];
  for (int k = 0; k < 3; ++k) {
    S[k] = n.model1->vertices[t1.v[k]];
    Tt[k] = mul(n.R, n.model2->vertices[t2.v[k]]) + n.T;  // intersect.cpp:401-409
  }
  V3 P1 = nan3(), P2 = nan3();
  const double d = std::sqrt(sqrTriDistance(S, Tt, P1, P2));
  if (count) n.num_leaf_tests++;
  if (n.min_distance > d) {  // DistanceResult::update with b2 (collision_data.h:1126-1138)
    n.min_distance = d;
    n.b1 = id1;
    n.b2 = id2;
    n.p1 = P1;
    n.p2 = P2;
  }
}
static inline bool mmCanStop(const MMDistNode& n, double c) {  // :473-478
  return (c >= n.min_distance - n.abs_err) && (c * (1 + n.rel_err) >= n.min_distance);
}
static inline double mmBvLowerBound(MMDistNode& n, unsigned b1, unsigned b2) {  // :426-434, :258-262
  n.num_bv_tests++;
  return rss_distance(n.R, n.T, n.model1->bvs[b1].bv.rss, n.model2->bvs[b2].bv.rss);
}
static void mmDistanceRecurse(MMDistNode& n, unsigned b1, unsigned b2) {  // traversal_recurse.cpp:153-203
  const BVNode& n1 = n.model1->bvs[b1];
  const BVNode& n2 = n.model2->bvs[b2];
  if (n1.first_child < 0 && n2.first_child < 0) {
    mmTriDistance(n, -(n1.first_child + 1), -(n2.first_child + 1), true);
    return;
  }
  unsigned a1, a2, c1, c2;
  if (firstOverSecond(n, b1, b2)) {
    a1 = (unsigned)n1.first_child;
    a2 = b2;
    c1 = (unsigned)n1.first_child + 1;
    c2 = b2;
  } else {
    a1 = b1;
    a2 = (unsigned)n2.first_child;
    c1 = b1;
    c2 = (unsigned)n2.first_child + 1;
  }
  const double d1 = mmBvLowerBound(n, a1, a2);
  const double d2 = mmBvLowerBound(n, c1, c2);
  if (d2 < d1) {
    if (!mmCanStop(n, d2)) mmDistanceRecurse(n, c1, c2);
    if (!mmCanStop(n, d1)) mmDistanceRecurse(n, a1, a2);
  } else {
    if (!mmCanStop(n, d1)) mmDistanceRecurse(n, a1, a2);
    if (!mmCanStop(n, d2)) mmDistanceRecurse(n, c1, c2);
  }
}

// orientedMeshDistance<MeshDistanceTraversalNodeOBBRSS> (distance_func_matrix.cpp:221-239) + distance(node)
// (collision_node.cpp:81-91) on a fresh DistanceResult.  `normal` is never written by this node
// (leafComputeDistance passes an uninitialised Vec3f): reported as NaN here.
void bvhBvhDistance(const BVHModel& m1, const Tf& tf1, const BVHModel& m2, const Tf& tf2, double rel_err,
                    double abs_err, bool enable_nearest_points, BvhQueryResult& out) {
  MMDistNode n;
  n.model1 = &m1;
  n.model2 = &m2;
  n.tf1 = tf1;
  n.R = tmul(tf1.R, tf2.R);
  n.T = tmul(tf1.R, tf2.T - tf1.T);
  n.rel_err = rel_err;
  n.abs_err = abs_err;
  n.min_distance = std::numeric_limits<double>::max();
  n.p1 = n.p2 = nan3();
  n.b1 = n.b2 = -1;
  mmTriDistance(n, 0, 0, false);  // preprocessOrientedNode (:487-510)
  mmDistanceRecurse(n, 0, 0);
  if (enable_nearest_points) {  // postprocessOrientedNode (:527-536)
    n.p1 = tf1.transform(n.p1);
    n.p2 = tf1.transform(n.p2);
  }
  out.distance = n.min_distance;
  out.p1 = n.p1;
  out.p2 = n.p2;
  out.normal = nan3();
  out.b1 = n.b1;
  out.b2 = n.b2;
  out.num_bv_tests = n.num_bv_tests;
  out.num_leaf_tests = n.num_leaf_tests;
}

struct MMColNode {
  const BVHModel *model1, *model2;
  Tf tf1, tf2;
  M3 R;
  V3 T;
  GJKSolver* solver;
  const hfb_collision_request* req;
  BvhCollideResult* res;
};
static void mmLeafCollides(MMColNode& n, unsigned b1, unsigned b2, double& sqrDistLowerBound) {  // :173-232
  const int id1 = -(n.model1->bvs[b1].first_child + 1), id2 = -(n.model2->bvs[b2].first_child + 1);
  const Tri& t1 = n.model1->tris[id1];
  const Tri& t2 = n.model2->tris[id2];
  Shape tri1, tri2;
  tri1.type = tri2.type = HFB_GEOM_TRIANGLE;
  for (int k = 0; k < 3; ++k) {
    tri1.tri[k] = n.model1->vertices[t1.v[k]];
    tri2.tri[k] = n.model2->vertices[t2.v[k]];
  }
  const bool compute_penetration = n.req->enable_contact || (n.req->security_margin < 0);
  V3 p1, p2, normal;
  double distance;
  bool closed;
  GJKSolver solver = *n.solver;  // "GJKSolver solver(this->request)" (:197): a fresh one per leaf
  shapeShapeDistance(tri1, n.tf1, tri2, n.tf2, solver, compute_penetration, distance, p1, p2, normal, closed);
  n.res->num_leaf_tests++;
  const double distToCollision = distance - n.req->security_margin;
  BvhCollideResult& r = *n.res;
  if (distToCollision < r.distance_lower_bound) {  // updateDistanceLowerBoundFromLeaf
    r.distance_lower_bound = distToCollision;
    r.lb_p1 = p1;
    r.lb_p2 = p2;
    r.lb_normal = normal;
  }
  if (distToCollision <= n.req->q.collision_distance_threshold) {
    sqrDistLowerBound = 0;
    if (r.contacts.size() < n.req->num_max_contacts) {
      BvhContact c;
      c.b1 = id1;
      c.b2 = id2;
      c.p1 = p1;
      c.p2 = p2;
      c.normal = normal;
      c.distance = distance;
      r.contacts.push_back(c);
    }
  } else {
    sqrDistLowerBound = distToCollision * distToCollision;
  }
}
static bool mmBvDisjoints(MMColNode& n, unsigned b1, unsigned b2, double& sqrDistLowerBound) {  // :147-162
  n.res->num_bv_tests++;
  // note the operand order of the reference: (RT, bv of model 2, bv of model 1)
  const bool disjoint = !obb_overlap(n.R, n.T, n.model2->bvs[b2].bv.obb, n.model1->bvs[b1].bv.obb,
                                     n.req->security_margin, n.req->break_distance, sqrDistLowerBound);
  if (disjoint) {
    BvhCollideResult& r = *n.res;
    if (r.distance_lower_bound > 0) {
      const double new_dlb = std::sqrt(sqrDistLowerBound);
      if (new_dlb < r.distance_lower_bound) r.distance_lower_bound = new_dlb;
    }
  }
  return disjoint;
}
static inline bool mmColCanStop(const MMColNode& n) {
  return !n.res->contacts.empty() && (n.req->num_max_contacts <= n.res->contacts.size());
}
static void mmCollisionRecurse(MMColNode& n, unsigned b1, unsigned b2, double& sqrDistLowerBound) {  // :44-85
  double lb1 = 0, lb2 = 0;
  const BVNode& n1 = n.model1->bvs[b1];
  const BVNode& n2 = n.model2->bvs[b2];
  if (n1.first_child < 0 && n2.first_child < 0) {
    mmLeafCollides(n, b1, b2, sqrDistLowerBound);
    return;
  }
  if (mmBvDisjoints(n, b1, b2, sqrDistLowerBound)) return;
  if (firstOverSecond(n, b1, b2)) {
    const unsigned c1 = (unsigned)n1.first_child, c2 = c1 + 1;
    mmCollisionRecurse(n, c1, b2, lb1);
    if (mmColCanStop(n)) return;
    mmCollisionRecurse(n, c2, b2, lb2);
  } else {
    const unsigned c1 = (unsigned)n2.first_child, c2 = c1 + 1;
    mmCollisionRecurse(n, b1, c1, lb1);
    if (mmColCanStop(n)) return;
    mmCollisionRecurse(n, b1, c2, lb2);
  }
  sqrDistLowerBound = std::min(lb1, lb2);
}

// orientedMeshCollide<MeshCollisionTraversalNodeOBBRSS> (collision_func_matrix.cpp:187-206) on a fresh result
void bvhBvhCollide(const BVHModel& m1, const Tf& tf1, const BVHModel& m2, const Tf& tf2, GJKSolver& solver,
                   const hfb_collision_request& req, BvhCollideResult& out) {
  MMColNode n;
  n.model1 = &m1;
  n.model2 = &m2;
  n.tf1 = tf1;
  n.tf2 = tf2;
  n.R = tmul(tf1.R, tf2.R);  // traversal_node_setup.h:561-563
  n.T = tmul(tf1.R, tf2.T - tf1.T);
  n.solver = &solver;
  n.req = &req;
  n.res = &out;
  double sqrDistLowerBound = 0;
  mmCollisionRecurse(n, 0, 0, sqrDistLowerBound);
}

}  // namespace oracle
