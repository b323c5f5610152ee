// TEST HARNESS ONLY -- not part of the product and never loaded by it.
//
// Compiles the per-pair DEVICE code of hpp-fcl_b200/csrc/*.cuh with g++ (G = 1
// lane per pair, HFB_HD expands to `inline`; lane groups only for the support argmax, see LaneSim) so that the exact arithmetic the
// CUDA kernels execute can be checked against the oracle in the CPU-only
// container (`-m "not gpu"` tests).  The product library has no such path: its
// entry points fail with HFB_ERR_NO_DEVICE when there is no GPU.
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "../../hpp-fcl_b200/csrc/hfb_arena.cuh"
#include "../../hpp-fcl_b200/csrc/hfb_bvh.cuh"
#include "../../hpp-fcl_b200/csrc/hfb_request.cuh"

// ---- lane groups on the host: Coop<G> for G > 1 (device-only in the product) simulated in two passes --------
// Pass 0 runs every lane of a group up to its cross-lane argmax and records what the lane brings to it; the
// butterfly of Coop<G>::argmax (hfb_shapes.cuh: xor-shuffle, `greater, or equal with the lower index`) is
// then replayed on the recorded values; pass 1 runs every lane again and hands it its reduced value.  Good
// for code with one reduction per call (shape_support): checks that every lane of a group ends with the
// answer of the serial scan.
namespace hfb {
struct LaneSim {
  static inline int lane = 0, pass = 0;
  static inline double v[32];
  static inline int idx[32];
  static void reduce(int G) {
    for (int off = G / 2; off > 0; off >>= 1) {
      double nv[32];
      int ni[32];
      for (int l = 0; l < G; ++l) {
        const double ov = v[l ^ off];
        const int oi = idx[l ^ off];
        nv[l] = v[l];
        ni[l] = idx[l];
        if (ov > v[l] || (ov == v[l] && oi < idx[l])) {
          nv[l] = ov;
          ni[l] = oi;
        }
      }
      for (int l = 0; l < G; ++l) {
        v[l] = nv[l];
        idx[l] = ni[l];
      }
    }
  }
};
#define HFB_EMU_COOP(G_)                                      \
  template <>                                                 \
  struct Coop<G_> {                                           \
    static int lane() { return LaneSim::lane; }               \
    static unsigned mask() { return 0; }                      \
    static void argmax(double& v, int& idx) {                 \
      if (LaneSim::pass == 0) {                               \
        LaneSim::v[LaneSim::lane] = v;                        \
        LaneSim::idx[LaneSim::lane] = idx;                    \
      } else {                                                \
        v = LaneSim::v[LaneSim::lane];                        \
        idx = LaneSim::idx[LaneSim::lane];                    \
      }                                                       \
    }                                                         \
    static void sync() {}                                     \
  };
HFB_EMU_COOP(2)
HFB_EMU_COOP(4)
HFB_EMU_COOP(8)
HFB_EMU_COOP(16)
HFB_EMU_COOP(32)
}  // namespace hfb

using namespace hfb;

namespace {
template <int G>
void lane_group_support(const ShapeD& s, v3 dir, int32_t* idx_per_lane) {
  int hint = 0;
  for (LaneSim::pass = 0; LaneSim::pass < 2; ++LaneSim::pass) {
    for (LaneSim::lane = 0; LaneSim::lane < G; ++LaneSim::lane) {
      hint = -7;
      shape_support<G, CAP_CONVEX>(s, dir, hint);
      if (LaneSim::pass == 1) idx_per_lane[LaneSim::lane] = hint;
    }
    if (LaneSim::pass == 0) LaneSim::reduce(G);
  }
  LaneSim::lane = LaneSim::pass = 0;
}

struct Emu {
  HostArena arena;
};
constexpr int CAPS_ALL = CAP_PRIM | CAP_CONVEX | CAP_TRI;

inline PairIn load_pair(const ArenaView& A, size_t i, const uint32_t* h1, const hfb_transform* tf1,
                        const uint32_t* h2, const hfb_transform* tf2, const hfb_query_request& q) {
  PairIn in;
  in.s1 = load_shape<CAPS_ALL>(A, h1[i]);
  in.s2 = load_shape<CAPS_ALL>(A, h2[i]);
  in.tf1 = load_xf(tf1[i].R);
  in.tf2 = load_xf(tf2[i].R);
  in.cached_guess = mk(1, 0, 0);
  in.hint0 = in.hint1 = 0;
  if (q.gjk_initial_guess == HFB_GUESS_CACHED) {
    if (q.cached_gjk_guess) in.cached_guess = mk(q.cached_gjk_guess[3 * i], q.cached_gjk_guess[3 * i + 1], q.cached_gjk_guess[3 * i + 2]);
    if (q.cached_support_func_guess) {
      in.hint0 = q.cached_support_func_guess[2 * i];
      in.hint1 = q.cached_support_func_guess[2 * i + 1];
    }
  }
  return in;
}

long g_retries = 0;
inline void run_pair(const PairIn& in, const SolverP& P, EpaWs* ws, PairOut& o) {
  GjkState g;
  std::memset(&g, 0, sizeof(g));
  if (pair_phase1<1, CAPS_ALL>(in, P, o, g)) {
    // the two tiers of k_epa: reduced-size workspace first, full size when the polytope outgrows it
    const GjkState g0 = g;
    EpaWsSmall small;
    if (!pair_phase2<1, CAPS_ALL>(in, P, g, &small, o)) {
      ++g_retries;
      g = g0;
      pair_phase2<1, CAPS_ALL>(in, P, g, ws, o);
    }
  }
}

inline void put_guess(const hfb_guess_out* go, size_t i, const PairOut& o) {
  if (!go) return;
  if (go->cached_gjk_guess) {
    go->cached_gjk_guess[3 * i] = o.cached_guess.x;
    go->cached_gjk_guess[3 * i + 1] = o.cached_guess.y;
    go->cached_gjk_guess[3 * i + 2] = o.cached_guess.z;
  }
  if (go->cached_support_func_guess) {
    go->cached_support_func_guess[2 * i] = o.hint0;
    go->cached_support_func_guess[2 * i + 1] = o.hint1;
  }
}
}  // namespace

extern "C" {

void* emu_create() { return new Emu(); }
long emu_epa_retries() { return g_retries; }
void emu_destroy(void* e) { delete static_cast<Emu*>(e); }

int emu_register_convex(void* e, const double* pts, uint32_t n) {
  return (int)static_cast<Emu*>(e)->arena.add_convex(pts, n);
}
int emu_register_bvh(void* e, const hfb_bvh_node* nodes, uint32_t nn, const double* verts, uint32_t nv,
                     const uint32_t* tris, uint32_t nt) {
  uint32_t id;
  if (!static_cast<Emu*>(e)->arena.add_bvh(nodes, nn, verts, nv, tris, nt, &id)) return -1;
  return (int)id;
}
int64_t emu_register_shapes(void* e, const hfb_shape* shapes, size_t n) {
  Emu* E = static_cast<Emu*>(e);
  int64_t first = (int64_t)E->arena.shapes.size();
  for (size_t i = 0; i < n; ++i) {
    uint32_t h;
    if (!E->arena.add_shape(shapes[i], &h)) return -1;
  }
  return first;
}

int emu_batch_distance(void* e, size_t n, const uint32_t* h1, const hfb_transform* tf1,
                       const uint32_t* h2, const hfb_transform* tf2, const hfb_distance_request* req,
                       hfb_distance_result* out, const hfb_guess_out* go) {
  Emu* E = static_cast<Emu*>(e);
  if (int rc = validate_query(req->q)) return rc;
  const SolverP P = solver_from_distance_request(*req);
  const ArenaView A = E->arena.view();
  std::unique_ptr<EpaWs> ws(new EpaWs());
  for (size_t i = 0; i < n; ++i) {
    if (h1[i] >= A.nshapes || h2[i] >= A.nshapes) return HFB_ERR_INVALID_ARGUMENT;
    if (A.shapes[h1[i]].type == HFB_BV_OBBRSS || A.shapes[h2[i]].type == HFB_BV_OBBRSS) {
      BvhReq R{/* rel_err, abs_err: see hfb_distance_request */ 0, 0, 0, 0, 0, 1, req->enable_nearest_points != 0, req->q.gjk_initial_guess};
      unsigned bt, lt;
      v3 guess = mk(1, 0, 0);
      int hh0 = 0, hh1 = 0;
      if (req->q.gjk_initial_guess == HFB_GUESS_CACHED) {
        if (req->q.cached_gjk_guess) guess = mk(req->q.cached_gjk_guess[3 * i], req->q.cached_gjk_guess[3 * i + 1], req->q.cached_gjk_guess[3 * i + 2]);
        if (req->q.cached_support_func_guess) { hh0 = req->q.cached_support_func_guess[2 * i]; hh1 = req->q.cached_support_func_guess[2 * i + 1]; }
      }
      const xf t1 = load_xf(tf1[i].R), t2 = load_xf(tf2[i].R);
      if (A.shapes[h1[i]].type == HFB_BV_OBBRSS && A.shapes[h2[i]].type == HFB_BV_OBBRSS) {
        bvh_mesh_pair_distance(A, h1[i], t1, h2[i], t2, R, &out[i], bt, lt);
      } else {
        BvhSingleSrc src;
        src.pending = bvh_make_job<CAPS_ALL, 0>(A, h1[i], t1, h2[i], t2, R, guess, hh0, hh1, &out[i], src.job);
        unsigned long long b2 = 0, l2 = 0;
        bvh_shape_distance_stream<CAPS_ALL>(src, P, R.rel_err, R.abs_err, ws.get(), b2, l2);
      }
      continue;
    }
    const PairIn in = load_pair(A, i, h1, tf1, h2, tf2, req->q);
    PairOut o;
    run_pair(in, P, ws.get(), o);
    write_distance(o, &out[i]);
    put_guess(go, i, o);
  }
  return HFB_OK;
}

int emu_batch_collide(void* e, size_t n, const uint32_t* h1, const hfb_transform* tf1,
                      const uint32_t* h2, const hfb_transform* tf2, const hfb_collision_request* req,
                      hfb_contact* out, const hfb_guess_out* go) {
  Emu* E = static_cast<Emu*>(e);
  if (int rc = validate_query(req->q)) return rc;
  const bool minus_inf = req->security_margin == -INFINITY;
  if (!minus_inf && req->num_max_contacts == 0) return HFB_ERR_INVALID_ARGUMENT;
  const SolverP P = solver_from_collision_request(*req);
  CollideP C;
  C.security_margin = req->security_margin;
  C.collision_distance_threshold = req->q.collision_distance_threshold;
  const ArenaView A = E->arena.view();
  std::unique_ptr<EpaWs> ws(new EpaWs());
  for (size_t i = 0; i < n; ++i) {
    if (h1[i] >= A.nshapes || h2[i] >= A.nshapes) return HFB_ERR_INVALID_ARGUMENT;
    PairOut o;
    if (minus_inf) {  // collision.cpp:73-76: result.clear(); return
      o.status = pack_status(0, 0, HFB_PATH_UNSUPPORTED);
      o.iterations = 0;
      write_contact(o, C, &out[i]);
      out[i].status = 0;
      continue;
    }
    if (A.shapes[h1[i]].type == HFB_BV_OBBRSS || A.shapes[h2[i]].type == HFB_BV_OBBRSS) {
      BvhReq R{0, 0, req->security_margin, req->break_distance, req->q.collision_distance_threshold,
               req->num_max_contacts, true, req->q.gjk_initial_guess};
      unsigned bt, lt;
      v3 guess = mk(1, 0, 0);
      int hh0 = 0, hh1 = 0;
      if (req->q.gjk_initial_guess == HFB_GUESS_CACHED) {
        if (req->q.cached_gjk_guess) guess = mk(req->q.cached_gjk_guess[3 * i], req->q.cached_gjk_guess[3 * i + 1], req->q.cached_gjk_guess[3 * i + 2]);
        if (req->q.cached_support_func_guess) { hh0 = req->q.cached_support_func_guess[2 * i]; hh1 = req->q.cached_support_func_guess[2 * i + 1]; }
      }
      const xf t1 = load_xf(tf1[i].R), t2 = load_xf(tf2[i].R);
      if (A.shapes[h1[i]].type == HFB_BV_OBBRSS && A.shapes[h2[i]].type == HFB_BV_OBBRSS) {
        bvh_mesh_pair_collide<CAPS_ALL>(A, h1[i], t1, h2[i], t2, P, R, guess, hh0, hh1, ws.get(), &out[i], bt, lt);
      } else {
        BvhSingleSrc src;
        src.pending = bvh_make_job<CAPS_ALL, 1>(A, h1[i], t1, h2[i], t2, R, guess, hh0, hh1, &out[i], src.job);
        unsigned long long b2 = 0, l2 = 0;
        bvh_shape_collide_stream<CAPS_ALL>(src, P, R.security_margin, R.break_distance, R.collision_distance_threshold,
                                           R.num_max_contacts, ws.get(), b2, l2);
      }
      continue;
    }
    const PairIn in = load_pair(A, i, h1, tf1, h2, tf2, req->q);
    run_pair(in, P, ws.get(), o);
    write_contact(o, C, &out[i]);
    put_guess(go, i, o);
  }
  return HFB_OK;
}

int emu_batch_convex_support(void* e, size_t n, const uint32_t* ids, const double* dirs, int32_t* idx,
                             double* sup) {
  Emu* E = static_cast<Emu*>(e);
  const ArenaView A = E->arena.view();
  for (size_t i = 0; i < n; ++i) {
    if (ids[i] >= A.ncvx) return HFB_ERR_INVALID_ARGUMENT;
    ShapeD s;
    const ConvexDesc& d = A.cvx[ids[i]];
    s.type = HFB_GEOM_CONVEX;
    s.cx = A.pool + d.off;
    s.cy = s.cx + d.vpad;
    s.cz = s.cy + d.vpad;
    s.nv = (int)d.nv;
    int hint = 0;
    const v3 r = shape_support<1, CAP_CONVEX>(s, mk(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), hint);
    idx[i] = hint;
    sup[3 * i] = r.x;
    sup[3 * i + 1] = r.y;
    sup[3 * i + 2] = r.z;
  }
  return HFB_OK;
}

// support vertex of a point set along `dir` as each of the G lanes of a group computes it (see LaneSim)
int emu_lane_group_support(int G, const double* points, int nv, const double* dir, int32_t* idx_per_lane) {
  std::vector<double> x(nv), y(nv), z(nv);
  for (int i = 0; i < nv; ++i) {
    x[i] = points[3 * i];
    y[i] = points[3 * i + 1];
    z[i] = points[3 * i + 2];
  }
  ShapeD s;
  s.type = HFB_GEOM_CONVEX;
  s.cx = x.data();
  s.cy = y.data();
  s.cz = z.data();
  s.nv = nv;
  const v3 d = mk(dir[0], dir[1], dir[2]);
  switch (G) {
    case 1: {
      int hint = 0;
      shape_support<1, CAP_CONVEX>(s, d, hint);
      idx_per_lane[0] = hint;
    } break;
    case 2: lane_group_support<2>(s, d, idx_per_lane); break;
    case 4: lane_group_support<4>(s, d, idx_per_lane); break;
    case 8: lane_group_support<8>(s, d, idx_per_lane); break;
    case 16: lane_group_support<16>(s, d, idx_per_lane); break;
    case 32: lane_group_support<32>(s, d, idx_per_lane); break;
    default: return HFB_ERR_INVALID_ARGUMENT;
  }
  return HFB_OK;
}

}  // extern "C"
