// Closed-form pair distances (the 26-of-55 pair types that never reach GJK,
// include/hpp/fcl/internal/shape_shape_func.h:185-306): signed distance, witness
// points and normal (shape 1 -> shape 2) in the world frame.
//
// Replaces details::sphereSphereDistance / sphereCapsuleDistance /
// sphereCylinderDistance / boxSphereDistance / sphereTriangleDistance
// (src/narrowphase/details.h:76-101,107-209,215-231,286-342,435-496) and
// ShapeShapeDistance<Capsule,Capsule> (src/distance/capsule_capsule.cpp:81-166).
#pragma once
#include "hfb_shapes.cuh"

namespace hfb {

struct Wit {  // distance + witness points + normal
  double d;
  v3 p1, p2, n;
};

HFB_HD Wit flip(const Wit& a) {  // the <T2,T1> specialisations: swap points, negate normal
  Wit r;
  r.d = a.d;
  r.p1 = a.p2;
  r.p2 = a.p1;
  r.n = -a.n;
  return r;
}

// details.h:215-231
HFB_HD Wit sphere_sphere(const ShapeD& s1, const xf& tf1, const ShapeD& s2, const xf& tf2) {
  Wit r;
  const double r1 = s1.p0 + s1.ssr, r2 = s2.p0 + s2.ssr;
  const v3 c1c2 = tf2.T - tf1.T;
  const double cdist = nrm(c1c2);
  v3 u = mk(1, 0, 0);
  if (cdist > DBL_EPSILON) u = c1c2 / cdist;
  r.d = cdist - r1 - r2;
  r.n = u;
  r.p1 = tf1.T + r1 * u;
  r.p2 = tf2.T - r2 * u;
  return r;
}

// details.h:52-70
HFB_HD v3 seg_point_closest(v3 p, v3 s1, v3 s2) {
  const v3 v = s2 - s1;
  const v3 w = p - s1;
  const double c1 = dot(w, v);
  const double c2 = dot(v, v);
  if (c1 <= 0) return s1;
  if (c2 <= c1) return s2;
  const double b = c1 / c2;
  return s1 + v * b;
}

// details.h:76-101 (s1 sphere, s2 capsule)
HFB_HD Wit sphere_capsule(const ShapeD& s1, const xf& tf1, const ShapeD& s2, const xf& tf2) {
  Wit r;
  const v3 pos1 = xform(tf2, mk(0., 0., s2.p1));
  const v3 pos2 = xform(tf2, mk(0., 0., -s2.p1));
  const v3 s_c = tf1.T;
  const v3 segp = seg_point_closest(s_c, pos1, pos2);
  v3 n = segp - s_c;
  const double nn = nrm(n);
  const double r1 = s1.p0 + s1.ssr;
  const double r2 = s2.p0 + s2.ssr;
  r.d = nn - r1 - r2;
  if (nn > DBL_EPSILON) n = unit(n);
  else n = mk(1, 0, 0);
  r.n = n;
  r.p1 = s_c + n * r1;
  r.p2 = segp - n * r2;
  return r;
}

// details.h:107-209 (s1 sphere, s2 cylinder)
HFB_HD Wit sphere_cylinder(const ShapeD& s1, const xf& tf1, const ShapeD& s2, const xf& tf2) {
  Wit r;
  const double eps = sqrt(DBL_EPSILON);
  const double r1 = s1.p0, r2 = s2.p0, lz2 = s2.p1;
  const v3 A = xform(tf2, mk(0, 0, -lz2));
  const v3 B = xform(tf2, mk(0, 0, lz2));
  const v3 S = tf1.T;
  const v3 u = mcol(tf2.R, 2);
  const v3 AS = S - A;
  const double s = dot(u, AS);
  const v3 P = A + s * u;
  const v3 PS = S - P;
  const double dPS = nrm(PS);
  v3 v = mk(0, 0, 0);
  if (dPS > eps) v = (1 / dPS) * PS;
  const bool lo = s <= 0;
  const bool mid = !lo && (s <= (s2.p1 * 2));
  if (mid) {
    r.n = -v;
    r.d = dPS - r1 - r2;
    r.p2 = P + r2 * v;
    r.p1 = S - r1 * v;
  } else {
    const v3 E = lo ? A : B;  // end-cap centre
    if (dPS <= r2) {
      // closest point on the disc
      if (lo) {
        r.d = -s - r1;
        r.p1 = S + r1 * u;
        r.n = u;
      } else {
        r.d = s - (s2.p1 * 2) - r1;
        r.p1 = S - r1 * u;
        r.n = -u;
      }
      r.p2 = E + dPS * v;
    } else {
      // closest point on the rim circle
      r.p2 = E + r2 * v;
      const v3 Sp2 = r.p2 - S;
      const double dSp2 = nrm(Sp2);
      if (dSp2 > eps) {
        r.n = (1 / dSp2) * Sp2;
        r.p1 = S + r1 * r.n;
        r.d = dSp2 - r1;
      } else {
        v3 n = r.p2 - .5 * (A + B);
        n = unit(n);
        r.n = n;
        r.d = -r1;
        r.p1 = S + r1 * n;
      }
    }
  }
  const double ssr1 = s1.ssr, ssr2 = s2.ssr;
  if (ssr1 > 0 || ssr2 > 0) {
    r.p1 = r.p1 + ssr1 * r.n;
    r.p2 = r.p2 - ssr2 * r.n;
    r.d -= (ssr1 + ssr2);
  }
  return r;
}

// details.h:435-496 (b box, s sphere)
HFB_HD Wit box_sphere(const ShapeD& b, const xf& tfb, const ShapeD& s, const xf& tfs) {
  Wit r;
  const v3 os = tfs.T, ob = tfb.T;
  v3 pb = ob;
  bool outside = false;
  const v3 q = mtmul(tfb.R, os - ob);  // os in box frame
  int axis = -1;
  double min_d = DBL_MAX;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double qi = comp(q, i);
    const double hi = i == 0 ? b.p0 : (i == 1 ? b.p1 : b.p2);
    const v3 ci = mcol(tfb.R, i);
    if (qi < -hi) {
      pb = pb - hi * ci;
      outside = true;
    } else if (qi > hi) {
      pb = pb + hi * ci;
      outside = true;
    } else {
      pb = pb + qi * ci;
      if (!outside) {
        const double facedist = hi - fabs(qi);
        if (facedist < min_d) {
          axis = i;
          min_d = facedist;
        }
      }
    }
  }
  v3 n = pb - os;
  const double pdist = nrm(n);
  double dist;
  if (outside) {
    dist = pdist - s.p0;
    n = n / (-pdist);
  } else {
    const v3 ca = mcol(tfb.R, axis);
    n = (comp(q, axis) >= 0) ? ca : -ca;
    dist = -min_d - s.p0;
  }
  v3 ps = os - s.p0 * n;
  if (!outside || dist <= 0) pb = ps - dist * n;
  const double ssrb = b.ssr, ssrs = s.ssr;
  if (ssrb > 0 || ssrs > 0) {
    pb = pb + ssrb * n;
    ps = ps - ssrs * n;
    dist -= (ssrb + ssrs);
  }
  r.d = dist;
  r.p1 = pb;
  r.p2 = ps;
  r.n = n;
  return r;
}

// capsule_capsule.cpp:51-74
HFB_HD v3 clamped_linear(v3 a, double s_n, double s_d, v3 d) {
  if (s_n <= 0.) return a;
  if (s_n >= s_d) return a + d;
  return a + s_n / s_d * d;
}
// capsule_capsule.cpp:81-166
HFB_HD Wit capsule_capsule(const ShapeD& s1, const xf& tf1, const ShapeD& s2, const xf& tf2) {
  Wit r;
  const double EPS = DBL_EPSILON * 100;
  const v3 c1 = tf1.T, c2 = tf2.T;
  const double radius1 = s1.p0 + s1.ssr, radius2 = s2.p0 + s2.ssr;
  const v3 d1 = 2 * s1.p1 * mcol(tf1.R, 2);
  const v3 d2 = 2 * s2.p1 * mcol(tf2.R, 2);
  const v3 p1 = c1 - d1 / 2;
  const v3 p2 = c2 - d2 / 2;
  const v3 rr = p1 - p2;
  const double a = dot(d1, d1), b = dot(d1, d2), c = dot(d1, rr), e = dot(d2, d2), f = dot(d2, rr);
  v3 w1, w2;
  if (a <= EPS) {
    w1 = p1;
    if (e <= EPS) w2 = p2;
    else w2 = clamped_linear(p2, f, e, d2);
  } else if (e <= EPS) {
    w1 = clamped_linear(p1, -c, a, d1);
    w2 = p2;
  } else {
    const double denom = fmax(a * e - b * b, 0.0);
    double s, t;
    if (denom > EPS) {
      const double num = b * f - c * e;
      s = (num <= 0.) ? 0. : ((num >= denom) ? 1. : num / denom);
      t = b * s + f;
    } else {
      s = 0.;
      t = f;
    }
    if (t <= 0.0) {
      w2 = p2;
      w1 = clamped_linear(p1, -c, a, d1);
    } else if (t >= e) {
      w1 = clamped_linear(p1, (b - c), a, d1);
      w2 = p2 + d2;
    } else {
      w1 = p1 + s * d1;
      w2 = p2 + t / e * d2;
    }
  }
  double distance = nrm(w1 - w2);
  distance = distance - (radius1 + radius2);
  r.n = unit(w2 - w1);
  r.p1 = w1 + radius1 * r.n;
  r.p2 = w2 - radius2 * r.n;
  r.d = distance;
  return r;
}

// details.h:235-255
HFB_HD double seg_sqr_distance(v3 from, v3 to, v3 p, v3& nearest) {
  v3 diff = p - from;
  const v3 v = to - from;
  double t = dot(v, diff);
  if (t > 0) {
    const double dotVV = sqn(v);
    if (t < dotVV) {
      t /= dotVV;
      diff = diff - v * t;
    } else {
      t = 1;
      diff = diff - v;
    }
  } else {
    t = 0;
  }
  nearest = from + v * t;
  return sqn(diff);
}

// details.h:286-342 (s sphere with tf1, tri with tf2)
HFB_HD Wit sphere_triangle(const ShapeD& s, const xf& tf1, const ShapeD& tri, const xf& tf2) {
  Wit r;
  const v3 P1 = xform(tf2, tri.ta), P2 = xform(tf2, tri.tb), P3 = xform(tf2, tri.tc);
  v3 tn = unit(cross(P2 - P1, P3 - P1));
  const v3 center = tf1.T;
  const double radius = s.p0 + s.ssr + tri.ssr;
  const v3 p1c = center - P1;
  double dplane = dot(p1c, tn);
  v3 closest = nan3();
  double min_sq;
  if (dplane < 0) {
    dplane *= -1;
    tn = tn * -1.0;
  }
  // projectInTriangle (details.h:258-280)
  const v3 e1 = P2 - P1, e2 = P3 - P2, e3 = P1 - P3;
  const double r1 = dot(cross(e1, tn), center - P1);
  const double r2 = dot(cross(e2, tn), center - P2);
  const double r3 = dot(cross(e3, tn), center - P3);
  if ((r1 > 0 && r2 > 0 && r3 > 0) || (r1 <= 0 && r2 <= 0 && r3 <= 0)) {
    closest = center - tn * dplane;
    min_sq = dplane * dplane;
  } else {
    v3 ne;
    min_sq = seg_sqr_distance(P1, P2, center, closest);
    double dsq = seg_sqr_distance(P2, P3, center, ne);
    if (dsq < min_sq) {
      min_sq = dsq;
      closest = ne;
    }
    dsq = seg_sqr_distance(P3, P1, center, ne);
    if (dsq < min_sq) {
      min_sq = dsq;
      closest = ne;
    }
  }
  r.n = unit(closest - center);
  r.p1 = center + r.n * (s.p0 + s.ssr);
  r.p2 = closest - r.n * tri.ssr;
  r.d = sqrt(min_sq) - radius;
  return r;
}

}  // namespace hfb
