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

// ---- Plane / Halfspace family (src/narrowphase/details.h:343-693; the ShapeShapeDistance<S, Halfspace / Plane>
// specialisations of src/distance/*_halfspace.cpp and *_plane.cpp).  Record: p0..p2 = unit normal,
// center.x = offset d (hfb_geom_register_halfspaces normalises like the constructors' unitNormalTest).
HFB_HD bool is_plane_type(int t) { return t == HFB_GEOM_PLANE || t == HFB_GEOM_HALFSPACE; }

// getSupport<WithSweptSphere> (support_functions.cpp:51-91) with hint 0 and a fresh ShapeSupportData: the support
// of the shape proper, plus (radius of a sphere / capsule +) swept-sphere radius along dir.normalized().
// Box: `inflate` of this instantiation is a function-local static of the reference, fixed by the first direction
// the process asks for (:146); 1 + 1e-10 here, its value once that direction had a zero component (a floor).
template <int G, int CAPS>
HFB_HD v3 support_swept(const ShapeD& s, v3 dir) {
  if (s.type == HFB_GEOM_SPHERE) return (s.p0 + s.ssr) * unit(dir);  // :164-176 (assigned, not added to zero)
  int hint = 0;
  const v3 sup = shape_support<G, CAPS>(s, dir, hint);
  const double r = (s.type == HFB_GEOM_CAPSULE) ? (s.p0 + s.ssr) : s.ssr;  // :206-222
  return sup + r * unit(dir);
}

struct PlaneW {  // Halfspace / Plane in the world frame
  v3 n;
  double d;
};
// transform(Halfspace / Plane, tf) (geometric_shapes_utility.cpp:249-290): n' = R n, d' = d + n'.T, and the
// constructor normalises (n', d') once more (unitNormalTest, geometric_shapes.cpp:121-143)
HFB_HD PlaneW plane_world(const ShapeD& h, const xf& tf, bool negate) {
  v3 n = mmul(tf.R, mk(h.p0, h.p1, h.p2));
  double d = h.center.x + dot(n, tf.T);
  if (negate) {  // the second halfspace of a plane: Halfspace(-n, -d)
    n = -n;
    d = -d;
  }
  PlaneW w;
  const double l = nrm(n);
  if (l > 0) {
    const double inv_l = 1.0 / l;
    w.n = n * inv_l;
    w.d = d * inv_l;
  } else {
    w.n = mk(1, 0, 0);
    w.d = 0;
  }
  return w;
}

// details.h:347-375: h halfspace (frame tf1), s any other shape (frame tf2); p1 on the halfspace, p2 on the shape
template <int G, int CAPS>
HFB_HD Wit halfspace_shape(const ShapeD& h, const xf& tf1, const ShapeD& s, const xf& tf2) {
  Wit r;
  const PlaneW w = plane_world(h, tf1, false);
  const v3 n_2 = mtmul(tf2.R, w.n);
  const v3 p2 = xform(tf2, support_swept<G, CAPS>(s, -n_2));
  const double dist = dot(w.n, p2) - (w.d + h.ssr);  // Halfspace::signedDistance
  r.d = dist;
  r.p2 = p2;
  r.p1 = p2 - dist * w.n;
  r.n = w.n;
  return r;
}
// details.h:381-428: the plane as two halfspaces, the larger signed distance wins (the first on a tie)
template <int G, int CAPS>
HFB_HD Wit plane_shape(const ShapeD& pl, const xf& tf1, const ShapeD& s, const xf& tf2) {
  Wit r;
  const PlaneW h0 = plane_world(pl, tf1, false), h1 = plane_world(pl, tf1, true);
  const v3 p2h1 = xform(tf2, support_swept<G, CAPS>(s, -mtmul(tf2.R, h0.n)));
  const v3 p2h2 = xform(tf2, support_swept<G, CAPS>(s, -mtmul(tf2.R, h1.n)));
  const double dist1 = dot(h0.n, p2h1) - (h0.d + pl.ssr);
  const double dist2 = dot(h1.n, p2h2) - (h1.d + pl.ssr);
  const bool first = dist1 >= dist2;
  r.d = first ? dist1 : dist2;
  r.p2 = first ? p2h1 : p2h2;
  r.n = first ? h0.n : h1.n;
  r.p1 = r.p2 - r.d * r.n;
  return r;
}
// the non-parallel branch of the three pairs below: infinite penetration, both points on the intersection line,
// "normal" its direction (details.h:546-560, 607-621, 671-685)
HFB_HD void plane_line(const PlaneW& a, const PlaneW& b, v3 dir, double dir_sq_norm, Wit& r) {
  r.d = -DBL_MAX;
  r.n = dir;
  r.p1 = r.p2 = cross(b.n * a.d - a.n * b.d, dir) / dir_sq_norm;
}
HFB_HD void plane_swept(const ShapeD& s1, const ShapeD& s2, Wit& r) {
  if (s1.ssr > 0 || s2.ssr > 0) {
    r.p1 = r.p1 + s1.ssr * r.n;
    r.p2 = r.p2 - s2.ssr * r.n;
    r.d -= (s1.ssr + s2.ssr);
  }
}
// details.h:509-571
HFB_HD Wit halfspace_halfspace(const ShapeD& s1, const xf& tf1, const ShapeD& s2, const xf& tf2) {
  Wit r;
  const PlaneW a = plane_world(s1, tf1, false), b = plane_world(s2, tf2, false);
  const v3 dir = cross(a.n, b.n);
  const double dir_sq_norm = sqn(dir);
  if (dir_sq_norm < DBL_EPSILON) {
    if (dot(a.n, b.n) > 0) {  // same normal: one inside the other
      r.d = -DBL_MAX;
      if (a.d <= b.d) {
        r.n = a.n;
        r.p1 = r.n * r.d;
        r.p2 = b.n * b.d;
      } else {
        r.n = -a.n;
        r.p1 = a.n * a.d;
        r.p2 = -(r.n * r.d);
      }
    } else {
      r.d = -(a.d + b.d);
      r.n = a.n;
      r.p1 = a.n * a.d;
      r.p2 = b.n * b.d;
    }
  } else {
    plane_line(a, b, dir, dir_sq_norm, r);
  }
  plane_swept(s1, s2, r);
  return r;
}
// details.h:585-632 (s1 halfspace, s2 plane)
HFB_HD Wit halfspace_plane(const ShapeD& s1, const xf& tf1, const ShapeD& s2, const xf& tf2) {
  Wit r;
  const PlaneW a = plane_world(s1, tf1, false), b = plane_world(s2, tf2, false);
  const v3 dir = cross(a.n, b.n);
  const double dir_sq_norm = sqn(dir);
  if (dir_sq_norm < DBL_EPSILON) {
    r.n = a.n;
    r.d = dot(a.n, b.n) > 0 ? (b.d - a.d) : -(a.d + b.d);
    r.p1 = a.n * a.d;
    r.p2 = b.n * b.d;
  } else {
    plane_line(a, b, dir, dir_sq_norm, r);
  }
  plane_swept(s1, s2, r);
  return r;
}
// details.h:646-693
HFB_HD Wit plane_plane(const ShapeD& s1, const xf& tf1, const ShapeD& s2, const xf& tf2) {
  Wit r;
  const PlaneW a = plane_world(s1, tf1, false), b = plane_world(s2, tf2, false);
  const v3 dir = cross(a.n, b.n);
  const double dir_sq_norm = sqn(dir);
  if (dir_sq_norm < DBL_EPSILON) {
    r.p1 = a.n * a.d;
    r.p2 = b.n * b.d;
    r.d = nrm(r.p1 - r.p2);
    r.n = (r.d > HFB_DUMMY_PRECISION) ? unit(r.p2 - r.p1) : a.n;
  } else {
    plane_line(a, b, dir, dir_sq_norm, r);
  }
  plane_swept(s1, s2, r);
  return r;
}
// dispatch of a pair with a plane or halfspace on either side
template <int G, int CAPS>
HFB_HD Wit plane_family(const ShapeD& s1, const xf& tf1, const ShapeD& s2, const xf& tf2) {
  const bool h1 = s1.type == HFB_GEOM_HALFSPACE, h2 = s2.type == HFB_GEOM_HALFSPACE;
  const bool q1 = s1.type == HFB_GEOM_PLANE, q2 = s2.type == HFB_GEOM_PLANE;
  if (h1 && h2) return halfspace_halfspace(s1, tf1, s2, tf2);
  if (q1 && q2) return plane_plane(s1, tf1, s2, tf2);
  if (h1 && q2) return halfspace_plane(s1, tf1, s2, tf2);
  if (q1 && h2) return flip(halfspace_plane(s2, tf2, s1, tf1));
  if (h1) return halfspace_shape<G, CAPS>(s1, tf1, s2, tf2);
  if (q1) return plane_shape<G, CAPS>(s1, tf1, s2, tf2);
  if (h2) return flip(halfspace_shape<G, CAPS>(s2, tf2, s1, tf1));
  return flip(plane_shape<G, CAPS>(s2, tf2, s1, tf1));
}

}  // namespace hfb
