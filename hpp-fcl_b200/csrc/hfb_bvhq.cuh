// Mesh-shape distance walk as a task system (kernel k_bvhq, hfb_bvhq_kernel.cuh).
//
// Replaces, like bvh_shape_distance_stream (hfb_bvh.cuh), for distance():
//   distanceRecurse                                   src/traversal/traversal_recurse.cpp:153-203
//   MeshShapeDistanceTraversalNodeOBBRSS              include/hpp/fcl/internal/traversal_node_bvh_shape.h:286-478
//   BVDistanceLowerBound = RSS distance               src/BV/RSS.cpp:995-1005
//
// Why another walk.  The reference recurses one query at a time; a query is a chain of dependent steps
// (bounding-volume test of two children -> push -> pop/prune -> ... -> triangle-shape GJK) and config 4
// holds one query of 9 360 steps next to a mean of 246.  A lane per query (k_bvh) leaves most lanes of a
// warp idle -- every lane is at a different step -- and is bound by that longest chain.  Here a query is a
// small state machine whose steps are ITEMS handed to whichever warp is free:
//   * a BV item  = the RSS distances of the two children of a node: 8 lanes, the sixteen edge-pair cases of
//     rectDistance split four per lane (rect_case), first passing case by ballot -- the reference's
//     first-return order;
//   * a leaf item = one triangle-shape GJK (+EPA): one lane, 32 items of 32 different queries per warp;
//   * the walk (pop / prune / push, canStop, both counters) runs on the lane that finished the query's
//     last outstanding item, in exactly the recursion's order.
// Items of one kind come from one queue, so every warp instruction works on 32 leaf tests or 4 x 8 lanes of
// bounding-volume tests, whatever the queries' individual progress.
//
// Speculation for long walks.  A bounding-volume distance is a pure function of (query, node); a leaf test is
// one too unless the request chains GJK guesses from leaf to leaf (CachedGuess).  Once a query has used
// `spec_after` items, reaching a node with at most HFB_Q_TREELET_MAX triangles below it (a "treelet"; the
// builder lays a subtree out as one contiguous block of node pairs, checked at registration) issues ALL the
// subtree's BV items and leaf items at once; when the last one lands the walk replays the recursion over the
// cached values -- same visiting order, same pruning, same counters (tests never visited are not counted) --
// and the chain of ~2 x 62 dependent steps costs two.  Results cannot differ from the unspeculated walk;
// tests/emu runs this file on the host with items executed in random order against the oracle.
#pragma once
#include "hfb_bvh.cuh"

namespace hfb {

#define HFB_Q_TREELET_MAX 128   // triangles of the largest subtree a walk speculates on
#define HFB_Q_TREELET_SMALL 32  // ... before it has used spec_big_after items
#define HFB_Q_ITEM_SPEC 0x100000u
#define HFB_Q_ITEM_RESUME 0x200000u  // leaf item: the solver state is parked, continue from it
#define HFB_Q_ITEM_VALID 0x80000000u
#define HFB_Q_SLOT_MASK 0xfffu

struct QStackEnt {  // a node still to be visited + the lower bound canStop() re-checks at pop time
  double dlow;
  int fc;    // first_child of the node (< 0: leaf, primitive -(fc + 1))
  int node;  // its index
};

struct QPrep {  // per query, written by k_bvhq_prep: the shape's RSS (computeBV) and b1^T * R0
  double sbv[15];  // axes (rows of the m3), Tr, l0, l1, radius
  double M[9];
  int ok, swapped;
};

// cached values of one speculated subtree; entry k belongs to descendant node tfc + k (k < 2 L - 2)
struct QTreeletHot {  // what the replay reads at every step: shared memory on the device
  double d[2 * HFB_Q_TREELET_MAX];      // RSS distance of the node (as a child of its parent)
  double leafd[2 * HFB_Q_TREELET_MAX];  // leaf nodes: distance of the triangle
  int fc[2 * HFB_Q_TREELET_MAX];        // first_child
};
struct QTreelet {  // read only for a leaf that improves the minimum: global memory
  double wit[2 * HFB_Q_TREELET_MAX][9];  // p1, p2, normal of the leaf tests
};

struct QSlot {  // one query in flight (shared memory on the device; 83 eight-byte words: odd stride)
  double tfm[12], tfs[12];
  double sbv[15], M[9];
  double shp[7];  // p0 p1 p2 ssr center
  double guess[3];
  double out[10];  // min_distance, p1, p2, normal
  const void* ptr[6];  // nodes, verts, tris, cx, cy, cz
  int type, nv, hint0, hint1, b1, sp, bv_tests, leaf_tests, pair, swapped, pending, rounds, tfc, tend, scr, cur;
  int seed, tbase;
};

struct QCtx {  // uniform per launch
  SolverP P;
  double rel_err, abs_err;
  int spec_after;      // items a query must have used before it may speculate (subtrees of up to HFB_Q_TREELET_SMALL
                       // triangles); < 0: never
  int spec_big_after;  // ... before it speculates on subtrees of up to HFB_Q_TREELET_MAX triangles
};

HFB_HD void q_put_m3(double* o, const m3& A) {
  o[0] = A.r0.x; o[1] = A.r0.y; o[2] = A.r0.z;
  o[3] = A.r1.x; o[4] = A.r1.y; o[5] = A.r1.z;
  o[6] = A.r2.x; o[7] = A.r2.y; o[8] = A.r2.z;
}
HFB_HD m3 q_get_m3(const double* o) {
  m3 A;
  A.r0 = mk(o[0], o[1], o[2]);
  A.r1 = mk(o[3], o[4], o[5]);
  A.r2 = mk(o[6], o[7], o[8]);
  return A;
}
HFB_HD void q_put_xf(double* o, const xf& t) {
  q_put_m3(o, t.R);
  o[9] = t.T.x; o[10] = t.T.y; o[11] = t.T.z;
}
HFB_HD xf q_get_xf(const double* o) {
  xf t;
  t.R = q_get_m3(o);
  t.T = mk(o[9], o[10], o[11]);
  return t;
}
HFB_HD void q_put_rss(double* o, const RssD& r) {
  q_put_m3(o, r.axes);
  o[9] = r.Tr.x; o[10] = r.Tr.y; o[11] = r.Tr.z;
  o[12] = r.l0; o[13] = r.l1; o[14] = r.radius;
}
HFB_HD RssD q_get_rss(const double* o) {
  RssD r;
  r.axes = q_get_m3(o);
  r.Tr = mk(o[9], o[10], o[11]);
  r.l0 = o[12]; r.l1 = o[13]; r.radius = o[14];
  return r;
}

// the per-query set-up of orientedBVHShapeDistance (traversal_node_setup.h:765): computeBV<OBBRSS>(shape)
// and the constant factor of every RSS distance of this query
HFB_HD void q_make_prep(const BvhQuery& q, bool swapped, QPrep& pr) {
  RssD sbv;
  compute_shape_rss(q.shape, q.tf_shape, sbv);
  q_put_rss(pr.sbv, sbv);
  q_put_m3(pr.M, mmulm(mtrans(sbv.axes), q.tf_mesh.R));
  pr.ok = 1;
  pr.swapped = swapped ? 1 : 0;
}

// a fresh walk: DistanceResult cleared, root on the stack, preprocess() seeds with triangle 0
// (traversal_node_bvh_shape.h:457-461) -- the caller pushes that leaf item
HFB_HD void q_start(QSlot& s, QStackEnt* stk, const BvhQuery& q, const QPrep& pr, unsigned pair, v3 guess, int h0,
                    int h1) {
  q_put_xf(s.tfm, q.tf_mesh);
  q_put_xf(s.tfs, q.tf_shape);
  for (int k = 0; k < 15; ++k) s.sbv[k] = pr.sbv[k];
  for (int k = 0; k < 9; ++k) s.M[k] = pr.M[k];
  s.shp[0] = q.shape.p0; s.shp[1] = q.shape.p1; s.shp[2] = q.shape.p2; s.shp[3] = q.shape.ssr;
  s.shp[4] = q.shape.center.x; s.shp[5] = q.shape.center.y; s.shp[6] = q.shape.center.z;
  s.guess[0] = guess.x; s.guess[1] = guess.y; s.guess[2] = guess.z;
  s.out[0] = DBL_MAX;
  for (int k = 1; k < 10; ++k) s.out[k] = nan3().x;
  s.ptr[0] = q.nodes; s.ptr[1] = q.verts; s.ptr[2] = q.tris;
  s.ptr[3] = q.shape.cx; s.ptr[4] = q.shape.cy; s.ptr[5] = q.shape.cz;
  s.type = q.shape.type;
  s.nv = q.shape.nv;
  s.hint0 = h0; s.hint1 = h1;
  s.b1 = -1;
  s.bv_tests = s.leaf_tests = 0;
  s.pair = (int)pair;
  s.swapped = pr.swapped;
  s.rounds = 0;
  s.tfc = s.tend = -1;
  s.tbase = 0;
  s.scr = -1;
  s.seed = 1;
  stk[0].dlow = -1.0;  // root: visited unconditionally
  stk[0].fc = q.nodes[0].first_child;
  stk[0].node = 0;
  s.sp = 1;
  s.cur = 0;
  s.pending = 1;
}

struct QLeafRes {
  double distance;
  v3 p1, p2, normal, guess;
  int hint0, hint1;
};
// operands of leafComputeDistance (traversal_node_bvh_shape.h:342-364): TriangleP(mesh triangle) vs the shape.
// `chained`: the warm start of this query's previous leaf goes in (normal items; it only matters to
// CachedGuess requests), otherwise the request's default.
HFB_HD void q_leaf_inputs(const QSlot& s, int prim, bool chained, PairIn& in) {
  const uint32_t* t = static_cast<const uint32_t*>(s.ptr[2]) + 3 * (size_t)prim;
  const double* verts = static_cast<const double*>(s.ptr[1]);
  const double* a = verts + 3 * (size_t)t[0];
  const double* b = verts + 3 * (size_t)t[1];
  const double* c = verts + 3 * (size_t)t[2];
  in.s1.type = HFB_GEOM_TRIANGLE;
  in.s1.ssr = 0;
  in.s1.p0 = in.s1.p1 = in.s1.p2 = 0;
  in.s1.nv = 0;
  in.s1.cx = in.s1.cy = in.s1.cz = nullptr;
  in.s1.center = mk(0, 0, 0);
  in.s1.ta = mk(a[0], a[1], a[2]);
  in.s1.tb = mk(b[0], b[1], b[2]);
  in.s1.tc = mk(c[0], c[1], c[2]);
  in.s2.type = s.type;
  in.s2.p0 = s.shp[0]; in.s2.p1 = s.shp[1]; in.s2.p2 = s.shp[2]; in.s2.ssr = s.shp[3];
  in.s2.center = mk(s.shp[4], s.shp[5], s.shp[6]);
  in.s2.cx = static_cast<const double*>(s.ptr[3]);
  in.s2.cy = static_cast<const double*>(s.ptr[4]);
  in.s2.cz = static_cast<const double*>(s.ptr[5]);
  in.s2.nv = s.nv;
  in.s2.ta = in.s2.tb = in.s2.tc = mk(0, 0, 0);
  in.tf1 = q_get_xf(s.tfm);
  in.tf2 = q_get_xf(s.tfs);
  in.cached_guess = chained ? mk(s.guess[0], s.guess[1], s.guess[2]) : mk(1, 0, 0);
  in.hint0 = chained ? s.hint0 : 0;
  in.hint1 = chained ? s.hint1 : 0;
}
HFB_HD void q_leaf_result(const PairOut& o, QLeafRes& r) {
  r.distance = o.distance;
  r.p1 = o.p1; r.p2 = o.p2; r.normal = o.normal;
  r.guess = o.cached_guess;
  r.hint0 = o.hint0; r.hint1 = o.hint1;
}
// the whole leaf test in one go (host emulation of an unsuspended item; the device splits it, see q_leaf_gjk)
template <int CAPS>
HFB_HD void q_leaf_eval(const QSlot& s, int prim, const SolverP& P, EpaWs* ws, bool chained, QLeafRes& r) {
  PairIn in;
  q_leaf_inputs(s, prim, chained, in);
  PairOut o;
  GjkState g;
  if (pair_phase1<1, CAPS, PATH_BOTH>(in, P, o, g)) pair_phase2<1, CAPS>(in, P, g, ws, o);
  q_leaf_result(o, r);
}

// A leaf test in pieces.  GJK takes 3 to 26 iterations on a triangle-capsule pair, and the block runs its leaf
// items in a phase that lasts as long as its longest item: a leaf item therefore runs at most `max_steps` GJK
// iterations, then parks the solver state (QLeafSave, global memory) and queues itself again (QL_SUSPENDED);
// EPA -- rare, up to 64 long iterations -- is an item kind of its own (QL_NEED_EPA, q_leaf_epa).  The pieces
// are pair_phase1 / pair_phase2 of hfb_pair.cuh cut at their loop boundaries: same operations in the same order.
struct QLeafSave {
  GjkState g;
  GjkLoop L;
};
enum { QL_DONE = 0, QL_SUSPENDED = 1, QL_NEED_EPA = 2 };
template <int CAPS>
HFB_HD int q_leaf_gjk(const QSlot& s, int prim, const SolverP& P, bool chained, bool resume, int max_steps,
                      QLeafSave& sv, QLeafRes& r) {
  PairIn in;
  q_leaf_inputs(s, prim, chained, in);
  PairOut o;
  GjkState g;
  if (is_closed_form(HFB_GEOM_TRIANGLE, in.s2.type)) {  // sphere partner: details.h:286-342, no solver
    pair_phase1<1, CAPS, PATH_BOTH>(in, P, o, g);
    q_leaf_result(o, r);
    return QL_DONE;
  }
  GjkSetup S;
  GjkLoop L;
  if (!resume) {
    pair_gjk_begin<CAPS>(in, P, S, L, g, o);
  } else {
    o.cached_guess = (P.initial_guess == HFB_GUESS_CACHED) ? in.cached_guess : mk(1, 0, 0);
    o.hint0 = in.hint0;
    o.hint1 = in.hint1;
    o.iterations = 0;
    make_setup<CAPS>(in, S);
    g = sv.g;
    L = sv.L;
  }
  bool more = true;
  for (int k = 0; k < max_steps && more; ++k) more = gjk_step<1, CAPS>(S.a, S.b, S.md, P.gjk, g, L);
  if (more) {
    sv.g = g;
    sv.L = L;
    return QL_SUSPENDED;
  }
  if (pair_gjk_end(P, S, g, o)) {
    sv.g = g;
    return QL_NEED_EPA;
  }
  q_leaf_result(o, r);
  return QL_DONE;
}
template <int CAPS>
HFB_HD void q_leaf_epa(const QSlot& s, int prim, const SolverP& P, bool chained, EpaWs* ws, const QLeafSave& sv,
                       QLeafRes& r) {
  PairIn in;
  q_leaf_inputs(s, prim, chained, in);
  PairOut o;
  o.cached_guess = mk(1, 0, 0);
  o.hint0 = o.hint1 = 0;
  GjkState g = sv.g;
  pair_phase2<1, CAPS>(in, P, g, ws, o);
  q_leaf_result(o, r);
}
// where the parked state of an item lives: one entry per slot (its one unspeculated leaf) and one per node of every
// treelet buffer
HFB_HD unsigned q_save_index(const QSlot& s, unsigned item, unsigned nslots) {
  return (item & HFB_Q_ITEM_SPEC) ? nslots + (unsigned)s.scr * 2 * HFB_Q_TREELET_MAX + ((item >> 12) & 0xffu)
                                  : (item & HFB_Q_SLOT_MASK);
}

// RSS distance of one node against the query's shape RSS: rss_distance (hfb_bvh.cuh) with the factor
// b1^T * R0 taken from the slot.  The serial form (host emulation, and the reference for the lane-group form)
HFB_HD void q_rss_operands(const QSlot& s, const hfb_bvh_node& nd, m3& R, v3& T, double& b0, double& b1, double& rad) {
  const m3 ax2 = load_colmajor(nd.rss_axes);
  const v3 Tr2 = mk(nd.rss_Tr[0], nd.rss_Tr[1], nd.rss_Tr[2]);
  const m3 M = q_get_m3(s.M);
  const m3 R0 = q_get_m3(s.tfm);
  const v3 T0 = mk(s.tfm[9], s.tfm[10], s.tfm[11]);
  const m3 ax1 = q_get_m3(s.sbv);
  const v3 Tr1 = mk(s.sbv[9], s.sbv[10], s.sbv[11]);
  R = mmulm(M, ax2);
  const v3 Ttemp = mmul(R0, Tr2) + T0 - Tr1;
  T = mtmul(ax1, Ttemp);
  b0 = nd.rss_length[0];
  b1 = nd.rss_length[1];
  rad = s.sbv[14] + nd.rss_radius;
}
HFB_HD double q_rss_child(const QSlot& s, const hfb_bvh_node& nd) {
  m3 R;
  v3 T;
  double b0, b1, rad;
  q_rss_operands(s, nd, R, T, b0, b1, rad);
  double dist = rect_distance(R, T, s.sbv[12], s.sbv[13], b0, b1, 0);
  dist -= rad;
  return (dist < 0.0) ? 0.0 : dist;
}

enum { Q_ISSUED = 0, Q_DONE = 1 };

// the two children of a node in the reference's order: the farther one goes on the stack, the nearer one is
// visited next (returned; the recursion would push it and pop it again at once)
HFB_HD QStackEnt q_push_children(QStackEnt* stk, int& sp, int base, double d1, double d2, int f1, int f2) {
  QStackEnt a, c;
  a.dlow = d1; a.fc = f1; a.node = base;
  c.dlow = d2; c.fc = f2; c.node = base + 1;
  if (d2 < d1) {
    stk[sp++] = a;
    return c;
  }
  stk[sp++] = c;
  return a;
}
// DistanceResult::update: strict '>' keeps the first minimum
HFB_HD void q_take_leaf(QSlot& s, int prim, const double* leaf10) {
  if (s.out[0] > leaf10[0]) {
    s.b1 = prim;
    for (int k = 0; k < 10; ++k) s.out[k] = leaf10[k];
  }
}

// The recursion, from the query's stack: pops / prunes (canStop, traversal_node_bvh_shape.h:322-327) until a
// node needs a value that is not at hand, issues the item(s) for it and returns Q_ISSUED; Q_DONE when the
// stack is empty.  Called by whoever completed the query's last outstanding item; `e0` (when `have`) is the node
// to visit before anything is popped -- the nearer child of the node a BV item just finished.
// Replay of a speculated subtree: its root sits on the query's stack at index s.tbase (re-pushed by the set-up);
// everything the replay pushes goes to a stack of its own (`lstk`, thread-local: the subtree is finished before this
// call returns, and the query's stack lives in global memory), every value it needs comes from `hot`.
#define HFB_Q_LOCAL_STACK 64
template <class Sink>
HFB_HD int q_advance(QSlot& s, unsigned slot_id, QStackEnt* stk, QTreelet* tls, QTreeletHot* hot, const QCtx& c,
                     Sink& sink, bool have, QStackEnt e) {
  int sp = s.sp;
  const hfb_bvh_node* nodes = static_cast<const hfb_bvh_node*>(s.ptr[0]);
  QStackEnt lstk[HFB_Q_LOCAL_STACK];
  int lsp = 0;
  bool inside = false;  // (of the entry in `e`)
  for (;;) {
    if (!have) {
      if (lsp > 0) {
        e = lstk[--lsp];
        inside = true;
      } else {
        if (sp == 0) break;
        e = stk[--sp];
        inside = s.scr >= 0 && sp == s.tbase;  // the speculated subtree's root
      }
    }
    have = false;
    if (e.dlow >= 0) {
      if ((e.dlow >= s.out[0] - c.abs_err) && (e.dlow * (1 + c.rel_err) >= s.out[0])) continue;
    }
    if (s.scr >= 0 && !inside) {  // the walk has left the speculated subtree
      sink.treelet_release(s.scr);
      s.scr = -1;
    }
    if (e.fc < 0) {  // leaf
      const int prim = -(e.fc + 1);
      if (inside) {
        const int k = e.node - s.tfc;
        s.leaf_tests++;
        const QTreeletHot& H = hot[s.scr];
        if (s.out[0] > H.leafd[k]) {  // DistanceResult::update
          s.b1 = prim;
          s.out[0] = H.leafd[k];
          const double* wv = tls[s.scr].wit[k];
          for (int q = 0; q < 9; ++q) s.out[1 + q] = wv[q];
        }
        continue;
      }
      s.cur = prim;
      s.pending = 1;
      s.rounds++;
      s.sp = sp;
      sink.push_leaf(slot_id);
      return Q_ISSUED;
    }
    if (inside) {  // both children cached
      const int k = e.fc - s.tfc;
      const QTreeletHot& H = hot[s.scr];
      s.bv_tests += 2;
      QStackEnt a, b;
      a.dlow = H.d[k]; a.fc = H.fc[k]; a.node = e.fc;
      b.dlow = H.d[k + 1]; b.fc = H.fc[k + 1]; b.node = e.fc + 1;
      if (b.dlow < a.dlow) {  // the nearer child is visited first
        lstk[lsp++] = a;
        e = b;
      } else {
        lstk[lsp++] = b;
        e = a;
      }
      have = true;
      inside = true;
      continue;
    }
    if (c.spec_after >= 0 && s.rounds >= c.spec_after) {
      const unsigned L = nodes[e.node]._pad;  // triangles below; 0 unless the subtree is one contiguous block
      const unsigned lmax = s.rounds >= c.spec_big_after ? HFB_Q_TREELET_MAX : HFB_Q_TREELET_SMALL;
      if (L >= 2 && L <= lmax) {
        const int id = sink.treelet_acquire();
        if (id >= 0) {
          const int nd = 2 * (int)L - 2;
          s.scr = id;
          s.tfc = e.fc;
          s.tbase = sp;
          e.dlow = -1.0;  // (it passed canStop just now, and nothing changes the minimum before the replay)
          stk[sp++] = e;  // the replay starts by popping this node again; its children are cached then
          s.sp = sp;
          // one BV item per pair of descendants; each spawns the leaf items of its own leaves (q_bv_store)
          s.pending = nd / 2 + (int)L;
          s.rounds++;
          sink.push_bv_pairs(slot_id, nd / 2);
          return Q_ISSUED;
        }
      }
    }
    s.cur = e.fc;
    s.pending = 1;
    s.rounds++;
    s.sp = sp;
    sink.push_bv(slot_id);
    return Q_ISSUED;
  }
  if (s.scr >= 0) {
    sink.treelet_release(s.scr);
    s.scr = -1;
  }
  s.sp = 0;
  return Q_DONE;
}

// completion of a BV item whose values are in hand (d of children base, base + 1 and their first_child).
// Returns true when this was the query's last outstanding item: the caller continues the walk (q_advance) with
// `near` (if `have`) as the node to visit first.  A speculated item stores its values and queues the leaf tests of
// the children that are leaves.
template <class Sink>
HFB_HD bool q_bv_store(QSlot& s, unsigned item, QStackEnt* stk, QTreeletHot* hot, Sink& sink, double d1, double d2, int f1,
                       int f2, bool& have, QStackEnt& near) {
  have = false;
  if (item & HFB_Q_ITEM_SPEC) {
    QTreeletHot& H = hot[s.scr];
    const unsigned k = (item >> 12) & 0xffu;
    H.d[k] = d1;
    H.d[k + 1] = d2;
    H.fc[k] = f1;
    H.fc[k + 1] = f2;
    const unsigned slot_id = item & HFB_Q_SLOT_MASK;
    if (f1 < 0) sink.push_leaf(slot_id | (k << 12) | HFB_Q_ITEM_SPEC);
    if (f2 < 0) sink.push_leaf(slot_id | ((k + 1) << 12) | HFB_Q_ITEM_SPEC);
    return sink.dec_pending(s) == 1;
  }
  s.bv_tests += 2;  // BVDistanceLowerBound of both children (:465-469)
  int sp = s.sp;
  near = q_push_children(stk, sp, s.cur, d1, d2, f1, f2);
  s.sp = sp;
  have = true;
  return true;
}
// node pair a BV item is about
HFB_HD int q_bv_base(const QSlot& s, unsigned item) {
  return (item & HFB_Q_ITEM_SPEC) ? s.tfc + (int)((item >> 12) & 0xffu) : s.cur;
}
HFB_HD int q_leaf_prim(const QSlot& s, const QTreeletHot* hot, unsigned item) {
  return (item & HFB_Q_ITEM_SPEC) ? -(hot[s.scr].fc[(item >> 12) & 0xffu] + 1) : s.cur;
}

// completion of a leaf item; true: the caller continues the walk
template <class Sink>
HFB_HD bool q_leaf_store(QSlot& s, unsigned item, QTreelet* tls, QTreeletHot* hot, Sink& sink, const QLeafRes& r) {
  if (item & HFB_Q_ITEM_SPEC) {
    const unsigned k = (item >> 12) & 0xffu;
    double* wv = tls[s.scr].wit[k];
    wv[0] = r.p1.x; wv[1] = r.p1.y; wv[2] = r.p1.z;
    wv[3] = r.p2.x; wv[4] = r.p2.y; wv[5] = r.p2.z;
    wv[6] = r.normal.x; wv[7] = r.normal.y; wv[8] = r.normal.z;
    hot[s.scr].leafd[k] = r.distance;
    return sink.dec_pending(s) == 1;
  }
  double v[10] = {r.distance, r.p1.x, r.p1.y, r.p1.z, r.p2.x, r.p2.y, r.p2.z, r.normal.x, r.normal.y, r.normal.z};
  if (!s.seed) s.leaf_tests++;  // the seed triangle of preprocess() is not a counted leaf test
  s.seed = 0;
  q_take_leaf(s, s.cur, v);
  // GJKSolver keeps cached_guess / support_func_cached_guess between calls (narrowphase.h:353-391, 625-626)
  s.guess[0] = r.guess.x; s.guess[1] = r.guess.y; s.guess[2] = r.guess.z;
  s.hint0 = r.hint0;
  s.hint1 = r.hint1;
  return true;
}

HFB_HD void q_write_result(const QSlot& s, hfb_distance_result* rec) {
  BvhDistOut o;
  o.min_distance = s.out[0];
  o.p1 = mk(s.out[1], s.out[2], s.out[3]);
  o.p2 = mk(s.out[4], s.out[5], s.out[6]);
  o.normal = mk(s.out[7], s.out[8], s.out[9]);
  o.b1 = s.b1;
  o.bv_tests = (unsigned)s.bv_tests;
  o.leaf_tests = (unsigned)s.leaf_tests;
  bvh_write_shape_distance(rec, s.swapped != 0, o);
}

}  // namespace hfb
