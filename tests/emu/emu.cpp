// TEST HARNESS ONLY -- not part of the product and never loaded by it.
//
// Compiles the per-pair DEVICE code of hpp-fcl_b200/csrc/*.cuh with g++ (G = 1
// lane per pair, HFB_HD expands to `inline`; lane groups of G threads: see lanesim below) so that the exact arithmetic the
// CUDA kernels execute can be checked against the oracle in the CPU-only
// container (`-m "not gpu"` tests).  The product library has no such path: its
// entry points fail with HFB_ERR_NO_DEVICE when there is no GPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <queue>
#include <vector>

#define HFB_LANE_SIM 1  // host lane groups, see below
#include "../../hpp-fcl_b200/csrc/hfb_arena.cuh"
#include "../../hpp-fcl_b200/csrc/hfb_bvh.cuh"
#include "../../hpp-fcl_b200/csrc/hfb_bvhq.cuh"
#include "../../hpp-fcl_b200/csrc/hfb_request.cuh"
#include "../../hpp-fcl_b200/csrc/hfb_broadphase.cuh"

// ---- lane groups on the host ---------------------------------------------------------------------------
// Coop<G> for G > 1 is warp intrinsics on the device.  Here (HFB_LANE_SIM) the G lanes of a group are G threads
// that run the same device function and meet at a barrier wherever the device code shuffles or syncs; whatever
// the lanes share on the device (the EPA workspace) they share here.  Between two barriers the threads interleave
// freely, so code that only works because a warp happens to run in lock step shows up as a mismatch (or a lane
// that never arrives: the barrier gives up after a while and the batch call reports it).
#include <atomic>
#include <chrono>
#include <thread>

namespace hfb {
namespace lanesim {
struct Group {
  int G = 1;
  std::atomic<int> count{0};
  std::atomic<int> gen{0};
  std::atomic<bool> failed{false};
  double dv[32];
  int iv[32];
  const void* ws[32];  // the lanes' private EPA workspaces (see hfb_epa.cuh, epa_flush_pending)
};
thread_local Group* tl_group = nullptr;
thread_local int tl_lane = 0;
int lane() { return tl_lane; }
void sync() {
  Group* g = tl_group;
  if (!g || g->G == 1 || g->failed.load(std::memory_order_relaxed)) return;
  const int gen = g->gen.load(std::memory_order_acquire);
  if (g->count.fetch_add(1, std::memory_order_acq_rel) + 1 == g->G) {
    g->count.store(0, std::memory_order_relaxed);
    g->gen.store(gen + 1, std::memory_order_release);
    return;
  }
  const auto t0 = std::chrono::steady_clock::now();
  for (unsigned spins = 0; g->gen.load(std::memory_order_acquire) == gen; ++spins) {
    if (g->failed.load(std::memory_order_relaxed)) return;
    if (spins > 200) std::this_thread::yield();
    if ((spins & 0xffff) == 0xffff && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(20)) {
      g->failed.store(true);  // a lane took another path: on the device this is a hang
      return;
    }
  }
}
void register_workspace(const void* p);
double shfl_xor(double v, int off) {
  Group* g = tl_group;
  g->dv[tl_lane] = v;
  sync();
  const double r = g->dv[tl_lane ^ off];
  sync();
  return r;
}
int shfl_xor(int v, int off) {
  Group* g = tl_group;
  g->iv[tl_lane] = v;
  sync();
  const int r = g->iv[tl_lane ^ off];
  sync();
  return r;
}
const void* peer_workspace(int l) { return tl_group->ws[l]; }
// phase sequences of BVH queries for the offline scheduling model (emu_bvh_trace_distance)
thread_local std::vector<uint8_t>* tl_trace = nullptr;
void trace_bvh_state(int state) {
  if (tl_trace) tl_trace->push_back((uint8_t)state);
}
void register_workspace(const void* p) {
  tl_group->ws[tl_lane] = p;
  sync();
}
}  // namespace lanesim
}  // namespace hfb

using namespace hfb;

namespace {
// runs fn(lane) on G threads that form one lane group; false when a lane never arrived at a barrier
template <class Fn>
bool run_lane_group(int G, Fn fn) {
  lanesim::Group grp;
  grp.G = G;
  std::vector<std::thread> th;
  for (int l = 0; l < G; ++l)
    th.emplace_back([&grp, l, &fn]() {
      lanesim::tl_group = &grp;
      lanesim::tl_lane = l;
      fn(l);
      lanesim::tl_group = nullptr;
    });
  for (auto& t : th) t.join();
  return !grp.failed.load();
}

template <int G>
bool lane_group_support(const ShapeD& s, v3 dir, int32_t* idx_per_lane) {
  return run_lane_group(G, [&](int l) {
    int hint = -7;
    shape_support<G, CAP_CONVEX>(s, dir, hint);
    idx_per_lane[l] = hint;
  });
}

struct Emu {
  HostArena arena;
};
constexpr int CAPS_ALL = CAP_PRIM | CAP_CONVEX | CAP_TRI;
constexpr int CAPS_ALLP = CAPS_ALL | CAP_PLANE;  // phase 1 of the shape-pair kernels (hfb_batch.cuh)

inline PairIn load_pair(const ArenaView& A, size_t i, const uint32_t* h1, const hfb_transform* tf1,
                        const uint32_t* h2, const hfb_transform* tf2, const hfb_query_request& q) {
  PairIn in;
  in.s1 = load_shape<CAPS_ALLP>(A, h1[i]);
  in.s2 = load_shape<CAPS_ALLP>(A, h2[i]);
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

long g_retries = 0, g_resumed = 0;
bool g_no_resume = false;  // HFB_EMU_NO_RESUME: tier 1 always starts over (the two must agree bit for bit)
inline void run_pair(const PairIn& in, const SolverP& P, EpaWs* ws, PairOut& o) {
  GjkState g;
  std::memset(&g, 0, sizeof(g));
  if (pair_phase1<1, CAPS_ALLP>(in, P, o, g)) {
    // the two tiers of k_epa: reduced-size workspace first, full size when the polytope outgrows it
    // (tier 1 continues from the state tier 0 reached when it stopped at the top of an iteration, else starts over)
    const GjkState g0 = g;
    EpaWsSmall small;
    EpaResume rs;
    rs.L.resumable = 0;
    if (!pair_phase2<1, CAPS_ALL>(in, P, g, &small, o, &rs)) {
      ++g_retries;
      if (rs.L.resumable && !g_no_resume) {
        ++g_resumed;
        epa_ws_grow<1>(&small, ws, rs.E);
        pair_phase2_resume<1, CAPS_ALL>(in, P, g0, ws, rs, o);
      } else {
        g = g0;
        pair_phase2<1, CAPS_ALL>(in, P, g, ws, o);
      }
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

// what k_pairs queues for k_epa and k_epa rebuilds (hfb_kernels.cu): rank, hints, GJK iteration count and the
// simplex's w0 / w1, w recomputed as w0 - w1
inline GjkState requeue(const GjkState& q) {
  GjkState g;
  std::memset(&g, 0, sizeof(g));
  g.rank = q.rank;
  g.hint0 = q.hint0;
  g.hint1 = q.hint1;
  g.iterations = q.iterations;
  g.status = HFB_GJK_COLLISION;
  g.distance = 0;
  g.ray = mk(0, 0, 0);
  g.s0.w0 = q.s0.w0; g.s0.w1 = q.s0.w1;
  g.s1.w0 = q.s1.w0; g.s1.w1 = q.s1.w1;
  g.s2.w0 = q.s2.w0; g.s2.w1 = q.s2.w1;
  g.s3.w0 = q.s3.w0; g.s3.w1 = q.s3.w1;
  g.s0.w = g.s0.w0 - g.s0.w1;
  g.s1.w = g.s1.w0 - g.s1.w1;
  g.s2.w = g.s2.w0 - g.s2.w1;
  g.s3.w = g.s3.w0 - g.s3.w1;
  return g;
}

// A batch of shape pairs run by lane groups of G threads: phase 1 (closed forms, GJK: k_pairs<G>) and phase 2
// (EPA: k_epa<G>, reduced workspace first, full size when the polytope outgrows it).  In phase 2 every lane
// works in a private copy of the polytope workspace -- EPA's lanes re-execute the serial parts of an iteration
// redundantly on one shared workspace, which is sound for the converged lanes of a warp and not for free-running
// threads -- and fetches the face geometry its peers computed (hfb_epa.cuh, epa_flush_pending).
// Returns the number of pairs on which some lane's outcome differed from lane 0's (must be 0), or -1 when a
// lane never reached a barrier.
template <int G, int MODE>
long batch_lanes(Emu* E, size_t n, const uint32_t* h1, const hfb_transform* tf1, const uint32_t* h2,
                 const hfb_transform* tf2, const hfb_query_request& q, const SolverP& P, const CollideP& C, void* out,
                 const hfb_guess_out* go) {
  const ArenaView A = E->arena.view();
  struct LaneOut {
    PairOut o;
    GjkState g;
    bool need_epa;
  };
  std::vector<LaneOut> outs(G);
  long disagree = 0;
  const bool ok = run_lane_group(G, [&](int l) {
    std::unique_ptr<EpaWs> ws(new EpaWs());
    std::unique_ptr<EpaWsSmall> small(new EpaWsSmall());
    for (size_t i = 0; i < n; ++i) {
      const PairIn in = load_pair(A, i, h1, tf1, h2, tf2, q);
      LaneOut& mine = outs[l];
      std::memset(&mine, 0, sizeof(mine));
      mine.need_epa = pair_phase1<G, CAPS_ALLP>(in, P, mine.o, mine.g);
      if (mine.need_epa) {  // k_epa: the queued state, tier 0, then tier 1
        const GjkState queued = mine.g;
        GjkState g = requeue(queued);
        mine.o.cached_guess = mk(1, 0, 0);
        mine.o.hint0 = mine.o.hint1 = 0;
        lanesim::register_workspace(small.get());
        Coop<G>::sync();
        EpaResume rs;
        rs.L.resumable = 0;
        const bool done = pair_phase2<G, CAPS_ALL>(in, P, g, small.get(), mine.o, &rs);
        Coop<G>::sync();
        if (!done) {
          if (l == 0) ++g_retries;
          mine.o.cached_guess = mk(1, 0, 0);
          mine.o.hint0 = mine.o.hint1 = 0;
          if (rs.L.resumable && !g_no_resume) {
            if (l == 0) ++g_resumed;
            epa_ws_grow<G>(small.get(), ws.get(), rs.E);
            lanesim::register_workspace(ws.get());
            Coop<G>::sync();
            pair_phase2_resume<G, CAPS_ALL>(in, P, requeue(queued), ws.get(), rs, mine.o);
          } else {
            g = requeue(queued);
            lanesim::register_workspace(ws.get());
            Coop<G>::sync();
            pair_phase2<G, CAPS_ALL>(in, P, g, ws.get(), mine.o);
          }
        }
      }
      Coop<G>::sync();
      if (l == 0) {
        for (int k = 1; k < G; ++k)
          if (outs[k].need_epa != mine.need_epa || std::memcmp(&outs[k].o, &mine.o, sizeof(PairOut)) != 0) {
            ++disagree;
            break;
          }
        if (MODE == 0) write_distance(mine.o, static_cast<hfb_distance_result*>(out) + i);
        else write_contact(mine.o, C, static_cast<hfb_contact*>(out) + i);
        put_guess(go, i, mine.o);
      }
      Coop<G>::sync();
    }
  });
  return ok ? disagree : -1;
}

template <int MODE>
long batch_lanes_g(int G, Emu* E, size_t n, const uint32_t* h1, const hfb_transform* tf1, const uint32_t* h2,
                   const hfb_transform* tf2, const hfb_query_request& q, const SolverP& P, const CollideP& C, void* out,
                   const hfb_guess_out* go) {
  const ArenaView A = E->arena.view();
  for (size_t i = 0; i < n; ++i) {
    if (h1[i] >= A.nshapes || h2[i] >= A.nshapes) return -2;
    if (is_bvh_type(A.shapes[h1[i]].type) || is_bvh_type(A.shapes[h2[i]].type)) return -2;  // one lane per query
  }
  switch (G) {
    case 2: return batch_lanes<2, MODE>(E, n, h1, tf1, h2, tf2, q, P, C, out, go);
    case 4: return batch_lanes<4, MODE>(E, n, h1, tf1, h2, tf2, q, P, C, out, go);
    case 8: return batch_lanes<8, MODE>(E, n, h1, tf1, h2, tf2, q, P, C, out, go);
    case 16: return batch_lanes<16, MODE>(E, n, h1, tf1, h2, tf2, q, P, C, out, go);
    default: return -2;
  }
}
}  // namespace


// ---- the task-system walk of hfb_bvhq.cuh on the host ---------------------------------------------------
// Same functions as the kernel k_bvhq; the block's queues are two vectors here and the items -- of several
// queries in flight, speculated subtrees included -- are executed ONE AT A TIME IN RANDOM ORDER, which is every
// interleaving the device's warps can produce as far as the walk logic can tell.
struct HostQSink {
  std::vector<unsigned> leafq, bvq, epaq;
  void push_epa(unsigned it) { epaq.push_back(it); }
  void push_bv_pairs(unsigned slot_id, int count) {
    for (int p = 0; p < count; ++p) bvq.push_back(slot_id | ((unsigned)(2 * p) << 12) | HFB_Q_ITEM_SPEC);
  }
  std::vector<int> free_tl;
  void push_leaf(unsigned it) { leafq.push_back(it); }
  void push_bv(unsigned it) { bvq.push_back(it); }
  int treelet_acquire() {
    if (free_tl.empty()) return -1;
    const int id = free_tl.back();
    free_tl.pop_back();
    return id;
  }
  void treelet_release(int id) { free_tl.push_back(id); }
  int dec_pending(QSlot& s) { return s.pending--; }
};
long g_q_spec_items = 0, g_q_items = 0;
// every (mesh, shape) pair of `todo` through the walk; spec_after < 0: never speculate
static void host_bvhq_distance(const ArenaView& A, const std::vector<size_t>& todo, const uint32_t* h1,
                               const hfb_transform* tf1, const uint32_t* h2, const hfb_transform* tf2,
                               const hfb_distance_request* req, const SolverP& P, hfb_distance_result* out,
                               int spec_after, unsigned seed, int nslots, int ntl, int gjk_chunk, int big_after) {
  BvhReq R{0, 0, 0, 0, 0, 1, req->enable_nearest_points != 0, req->q.gjk_initial_guess};
  QCtx c;
  c.P = P;
  c.rel_err = R.rel_err;
  c.abs_err = R.abs_err;
  c.spec_after = (P.initial_guess == HFB_GUESS_CACHED) ? -1 : spec_after;
  c.spec_big_after = spec_after + big_after;
  std::vector<QSlot> slots((size_t)nslots);
  std::vector<QStackEnt> stacks((size_t)nslots * 64);
  std::vector<QTreelet> tls((size_t)(ntl > 0 ? ntl : 1));
  std::vector<QTreeletHot> hot((size_t)(ntl > 0 ? ntl : 1));
  std::vector<QLeafSave> saves((size_t)nslots + (size_t)(ntl > 0 ? ntl : 1) * 2 * HFB_Q_TREELET_MAX);
  HostQSink sink;
  for (int k = 0; k < ntl; ++k) sink.free_tl.push_back(k);
  std::unique_ptr<EpaWs> ws(new EpaWs());
  size_t next = 0;
  int active = 0;
  unsigned long long rng = 0x9E3779B97F4A7C15ull * (seed + 1);
  auto rnd = [&]() {
    rng = rng * 6364136223846793005ull + 1442695040888963407ull;
    return (unsigned)(rng >> 33);
  };
  auto fetch = [&](int sl) -> bool {
    while (next < todo.size()) {
      const size_t i = todo[next++];
      v3 guess = mk(1, 0, 0);
      int hh0 = 0, hh1 = 0;
      if (req->q.gjk_initial_guess == HFB_GUESS_CACHED) {
        if (req->q.cached_gjk_guess) guess = mk(req->q.cached_gjk_guess[3 * i], req->q.cached_gjk_guess[3 * i + 1], req->q.cached_gjk_guess[3 * i + 2]);
        if (req->q.cached_support_func_guess) { hh0 = req->q.cached_support_func_guess[2 * i]; hh1 = req->q.cached_support_func_guess[2 * i + 1]; }
      }
      const xf t1 = load_xf(tf1[i].R), t2 = load_xf(tf2[i].R);
      BvhJob job;
      if (!bvh_make_job<CAPS_ALL, 0>(A, h1[i], t1, h2[i], t2, R, guess, hh0, hh1, &out[i], job)) continue;
      QPrep pr;
      q_make_prep(job.q, job.swapped, pr);
      q_start(slots[sl], &stacks[(size_t)sl * 64], job.q, pr, (unsigned)i, guess, hh0, hh1);
      sink.push_leaf((unsigned)sl);
      return true;
    }
    return false;
  };
  for (int sl = 0; sl < nslots; ++sl)
    if (fetch(sl)) ++active;
  while (active > 0) {
    const size_t nl = sink.leafq.size(), nb = sink.bvq.size(), ne = sink.epaq.size();
    if (nl + nb + ne == 0) std::abort();  // a query in flight always has an item outstanding
    size_t pick = rnd() % (nl + nb + ne);
    unsigned item;
    int rc = Q_ISSUED;
    ++g_q_items;
    if (pick < nl + ne) {
      const bool epa = pick >= nl;
      std::vector<unsigned>& qv = epa ? sink.epaq : sink.leafq;
      if (epa) pick -= nl;
      item = qv[pick];
      qv[pick] = qv.back();
      qv.pop_back();
      const unsigned sl = item & HFB_Q_SLOT_MASK;
      QSlot& s = slots[sl];
      const bool spec = (item & HFB_Q_ITEM_SPEC) != 0;
      g_q_spec_items += spec;
      QLeafRes r;
      QLeafSave& sv = saves[q_save_index(s, item, (unsigned)nslots)];
      const int prim = q_leaf_prim(s, hot.data(), item);
      int st = QL_DONE;
      if (epa) q_leaf_epa<CAPS_ALL>(s, prim, P, !spec, ws.get(), sv, r);
      else st = q_leaf_gjk<CAPS_ALL>(s, prim, P, !spec, (item & HFB_Q_ITEM_RESUME) != 0, gjk_chunk, sv, r);
      if (st == QL_SUSPENDED) sink.push_leaf(item | HFB_Q_ITEM_RESUME);
      else if (st == QL_NEED_EPA) sink.push_epa(item & ~HFB_Q_ITEM_RESUME);
      else if (q_leaf_store(s, item, tls.data(), hot.data(), sink, r))
        rc = q_advance(s, sl, &stacks[(size_t)sl * 64], tls.data(), hot.data(), c, sink, false, QStackEnt());
    } else {
      pick -= nl + ne;
      item = sink.bvq[pick];
      sink.bvq[pick] = sink.bvq.back();
      sink.bvq.pop_back();
      QSlot& s = slots[item & HFB_Q_SLOT_MASK];
      g_q_spec_items += (item & HFB_Q_ITEM_SPEC) != 0;
      const hfb_bvh_node* nodes = static_cast<const hfb_bvh_node*>(s.ptr[0]);
      const int base = q_bv_base(s, item);
      const double d1 = q_rss_child(s, nodes[base]), d2 = q_rss_child(s, nodes[base + 1]);
      bool have;
      QStackEnt near;
      if (q_bv_store(s, item, &stacks[(size_t)(item & HFB_Q_SLOT_MASK) * 64], hot.data(), sink, d1, d2,
                     nodes[base].first_child, nodes[base + 1].first_child, have, near))
        rc = q_advance(s, item & HFB_Q_SLOT_MASK, &stacks[(size_t)(item & HFB_Q_SLOT_MASK) * 64], tls.data(), hot.data(), c, sink, have, near);
    }
    if (rc == Q_DONE) {
      const int sl = (int)(item & HFB_Q_SLOT_MASK);
      q_write_result(slots[sl], &out[slots[sl].pair]);
      if (!fetch(sl)) --active;
    }
  }
  if (!sink.leafq.empty() || !sink.bvq.empty() || !sink.epaq.empty() || (int)sink.free_tl.size() != ntl) std::abort();
}

extern "C" {

long emu_q_spec_items() { return g_q_spec_items; }
long emu_q_items() { return g_q_items; }
void* emu_create() { return new Emu(); }
long emu_epa_retries() { return g_retries; }
long emu_epa_resumed() { return g_resumed; }
void emu_set_epa_resume(int on) { g_no_resume = on == 0; }
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
int emu_register_bvh_obb(void* e, const hfb_bvh_node* nodes, uint32_t nn, const double* verts, uint32_t nv,
                         const uint32_t* tris, uint32_t nt) {
  uint32_t id;
  if (!static_cast<Emu*>(e)->arena.add_bvh(nodes, nn, verts, nv, tris, nt, &id, 1)) return -1;
  return (int)id;
}
int64_t emu_register_shapes(void* e, const hfb_shape* shapes, size_t n) {
  Emu* E = static_cast<Emu*>(e);
  int64_t first = (int64_t)E->arena.shapes.size();
  for (size_t i = 0; i < n; ++i) {
    uint32_t h;
    if (shapes[i].type == HFB_GEOM_PLANE || shapes[i].type == HFB_GEOM_HALFSPACE) return -1;  // as hfb_geom_register_shapes
    if (!E->arena.add_shape(shapes[i], &h)) return -1;
  }
  return first;
}

int64_t emu_register_halfspaces(void* e, uint32_t type, const double* nd, const double* ssr, size_t count) {
  Emu* E = static_cast<Emu*>(e);
  int64_t first = (int64_t)E->arena.shapes.size();
  for (size_t i = 0; i < count; ++i) {
    uint32_t h;
    if (!E->arena.add_halfspace(type, nd + 4 * i, nd[4 * i + 3], ssr ? ssr[i] : 0.0, &h)) return -1;
  }
  return first;
}

// hfb_scene_aabbs on the host: aabb_local of the shape records (what hfb_geom_commit tabulates) + object_aabb
int emu_scene_aabbs(void* e, size_t n, const uint32_t* handles, const hfb_transform* tfs, double* out) {
  Emu* E = static_cast<Emu*>(e);
  for (size_t i = 0; i < n; ++i) {
    LocalAabb b;
    if (handles[i] >= E->arena.shapes.size() || !shape_local_aabb(E->arena, E->arena.shapes[handles[i]], b)) {
      for (int k = 0; k < 3; ++k) {
        out[6 * i + k] = DBL_MAX;
        out[6 * i + 3 + k] = -DBL_MAX;
      }
      continue;
    }
    object_aabb(b.mn, b.mx, tfs[i], out + 6 * i);
  }
  return HFB_OK;
}

int emu_update_shapes(void* e, const uint32_t* handles, const hfb_shape* shapes, size_t n) {
  Emu* E = static_cast<Emu*>(e);
  for (size_t i = 0; i < n; ++i)
    if (handles[i] >= E->arena.shapes.size() || !E->arena.valid_shape(shapes[i])) return HFB_ERR_INVALID_ARGUMENT;
  for (size_t i = 0; i < n; ++i) E->arena.set_shape(handles[i], shapes[i]);
  return HFB_OK;
}
int emu_update_convex(void* e, uint32_t id, const double* pts, uint32_t n) {
  return static_cast<Emu*>(e)->arena.set_convex(id, pts, n) ? HFB_OK : HFB_ERR_INVALID_ARGUMENT;
}

int emu_batch_distance(void* e, size_t n, const uint32_t* h1, const hfb_transform* tf1,
                       const uint32_t* h2, const hfb_transform* tf2, const hfb_distance_request* req,
                       hfb_distance_result* out, const hfb_guess_out* go) {
  Emu* E = static_cast<Emu*>(e);
  if (int rc = validate_query(req->q)) return rc;
  const SolverP P = solver_from_distance_request(*req);
  const ArenaView A = E->arena.view();
  std::unique_ptr<EpaWs> ws(new EpaWs());
  // HFB_EMU_BVHQ="spec_after[,seed[,slots[,treelets[,GJK iterations per leaf item[,items more before big subtrees]]]]]": (mesh, shape) pairs through the task-system walk
  const char* qenv = getenv("HFB_EMU_BVHQ");
  std::vector<size_t> qtodo;
  for (size_t i = 0; i < n; ++i) {
    if (h1[i] >= A.nshapes || h2[i] >= A.nshapes) return HFB_ERR_INVALID_ARGUMENT;
    const bool m1 = is_bvh_type(A.shapes[h1[i]].type), m2 = is_bvh_type(A.shapes[h2[i]].type);
    if (qenv && (m1 != m2)) {
      qtodo.push_back(i);
      continue;
    }
    if (is_bvh_type(A.shapes[h1[i]].type) || is_bvh_type(A.shapes[h2[i]].type)) {
      BvhReq R{/* rel_err, abs_err: see hfb_distance_request */ 0, 0, 0, 0, 0, 1, req->enable_nearest_points != 0, req->q.gjk_initial_guess};
      unsigned bt, lt;
      v3 guess = mk(1, 0, 0);
      int hh0 = 0, hh1 = 0;
      if (req->q.gjk_initial_guess == HFB_GUESS_CACHED) {
        if (req->q.cached_gjk_guess) guess = mk(req->q.cached_gjk_guess[3 * i], req->q.cached_gjk_guess[3 * i + 1], req->q.cached_gjk_guess[3 * i + 2]);
        if (req->q.cached_support_func_guess) { hh0 = req->q.cached_support_func_guess[2 * i]; hh1 = req->q.cached_support_func_guess[2 * i + 1]; }
      }
      const xf t1 = load_xf(tf1[i].R), t2 = load_xf(tf2[i].R);
      if (is_bvh_type(A.shapes[h1[i]].type) && is_bvh_type(A.shapes[h2[i]].type)) {
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
  if (!qtodo.empty()) {
    int spec_after = 0, slots = 6, ntl = 2, chunk = 3, big_after = 30;
    unsigned seed = 1;
    sscanf(qenv, "%d,%u,%d,%d,%d,%d", &spec_after, &seed, &slots, &ntl, &chunk, &big_after);
    host_bvhq_distance(A, qtodo, h1, tf1, h2, tf2, req, P, out, spec_after, seed, slots, ntl, chunk, big_after);
  }
  return HFB_OK;
}

static int emu_collide_impl(void* e, size_t n, const uint32_t* h1, const hfb_transform* tf1, const uint32_t* h2,
                            const hfb_transform* tf2, const hfb_collision_request* req, hfb_contact* out,
                            const hfb_guess_out* go, uint32_t max_extra, hfb_contact* extra, uint32_t* counts);

int emu_batch_collide(void* e, size_t n, const uint32_t* h1, const hfb_transform* tf1,
                      const uint32_t* h2, const hfb_transform* tf2, const hfb_collision_request* req,
                      hfb_contact* out, const hfb_guess_out* go) {
  return emu_collide_impl(e, n, h1, tf1, h2, tf2, req, out, go, 0, nullptr, nullptr);
}
// hfb_batch_collide_contacts of the library on the host build of the device code
int emu_batch_collide_contacts(void* e, size_t n, const uint32_t* h1, const hfb_transform* tf1, const uint32_t* h2,
                               const hfb_transform* tf2, const hfb_collision_request* req, hfb_contact* out,
                               uint32_t max_extra, hfb_contact* extra, uint32_t* counts) {
  for (size_t i = 0; i < n; ++i) counts[i] = 0xffffffffu;
  const int rc = emu_collide_impl(e, n, h1, tf1, h2, tf2, req, out, nullptr, max_extra, extra, counts);
  if (rc) return rc;
  for (size_t i = 0; i < n; ++i)
    if (counts[i] == 0xffffffffu) counts[i] = out[i].num_contacts;
  return HFB_OK;
}

static int emu_collide_impl(void* e, size_t n, const uint32_t* h1, const hfb_transform* tf1, const uint32_t* h2,
                            const hfb_transform* tf2, const hfb_collision_request* req, hfb_contact* out,
                            const hfb_guess_out* go, uint32_t max_extra, hfb_contact* extra, uint32_t* counts) {
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
    if (is_bvh_type(A.shapes[h1[i]].type) || is_bvh_type(A.shapes[h2[i]].type)) {
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
      BvhContactSink sink;
      sink.extra = (extra && max_extra) ? extra + i * (size_t)max_extra : nullptr;
      sink.cap = sink.extra ? max_extra : 0u;
      sink.count = counts ? counts + i : nullptr;
      if (is_bvh_type(A.shapes[h1[i]].type) && is_bvh_type(A.shapes[h2[i]].type)) {
        bvh_mesh_pair_collide<CAPS_ALL>(A, h1[i], t1, h2[i], t2, P, R, guess, hh0, hh1, ws.get(), &out[i], bt, lt, sink);
      } else {
        BvhSingleColSrc src;
        src.job.sink = sink;
        if (sink.count) *sink.count = 0;
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

// The sequence of phases (BVS_NEED_INIT / BVS_NEED_BV / BVS_NEED_LEAF, one byte per scheduling round) every
// (mesh, shape) distance query goes through, for tests/tools/bvh_sched_model.py.  offsets: n + 1 entries.
// Returns the total length, or -1 if `cap` bytes do not hold it.
long emu_bvh_trace_distance(void* e, size_t n, const uint32_t* h1, const hfb_transform* tf1, const uint32_t* h2,
                            const hfb_transform* tf2, const hfb_distance_request* req, uint8_t* states, size_t cap,
                            uint64_t* offsets) {
  Emu* E = static_cast<Emu*>(e);
  const SolverP P = solver_from_distance_request(*req);
  const ArenaView A = E->arena.view();
  std::unique_ptr<EpaWs> ws(new EpaWs());
  std::vector<uint8_t> tr;
  size_t total = 0;
  BvhReq R{0, 0, 0, 0, 0, 1, req->enable_nearest_points != 0, req->q.gjk_initial_guess};
  for (size_t i = 0; i < n; ++i) {
    offsets[i] = total;
    tr.clear();
    hfb_distance_result rec;
    BvhSingleSrc src;
    src.pending = bvh_make_job<CAPS_ALL, 0>(A, h1[i], load_xf(tf1[i].R), h2[i], load_xf(tf2[i].R), R, mk(1, 0, 0), 0, 0,
                                            &rec, src.job);
    unsigned long long b2 = 0, l2 = 0;
    lanesim::tl_trace = &tr;
    bvh_shape_distance_stream<CAPS_ALL>(src, P, 0.0, 0.0, ws.get(), b2, l2);
    lanesim::tl_trace = nullptr;
    for (uint8_t st : tr)
      if (st == BVS_NEED_INIT || st == BVS_NEED_BV || st == BVS_NEED_LEAF) {
        if (total >= cap) return -1;
        states[total++] = st;
      }
  }
  offsets[n] = total;
  return (long)total;
}

// Model of k_bvh's warps (see tests/tools/bvh_sched_model.py): `warps` warps of 32 lanes pull queries in `order`
// from one counter, each lane holding up to `slots` queries at a time (1 = the kernel as built); each scheduling
// round a warp votes like bvh_vote over the lanes that have a query waiting for a phase and runs one phase: every
// such lane advances one of its queries by one step.  A round of phase p takes cost[p] of the warp's time.
// policy 0: bvh_vote as built; 1: bounding volumes win ties.
// stats: [0] makespan, [1..3] rounds per phase, [4..6] lane-steps per phase, [7] time at which the counter ran dry
int emu_bvh_sched_sim(size_t nq, const uint8_t* states, const uint64_t* offsets, const uint32_t* order, int warps,
                      const double* cost, int init_quorum, int policy, int slots, double* stats) {
  struct Slot {
    int64_t q = -1;
    uint64_t pos = 0, end = 0;
  };
  if (slots < 1 || slots > 4) return HFB_ERR_INVALID_ARGUMENT;
  std::vector<Slot> st((size_t)warps * 32 * slots);
  size_t next = 0;
  double dry_at = -1;
  for (int k = 0; k < 8; ++k) stats[k] = 0;
  typedef std::pair<double, int> Ev;
  std::priority_queue<Ev, std::vector<Ev>, std::greater<Ev>> pq;
  for (int w = 0; w < warps; ++w) pq.push(Ev(0.0, w));
  while (!pq.empty()) {
    const int w = pq.top().second;
    const double t = pq.top().first;
    pq.pop();
    Slot* L = &st[(size_t)w * 32 * slots];
    int cnt[3] = {0, 0, 0};
    for (int l = 0; l < 32; ++l) {
      bool want[3] = {false, false, false};
      for (int k = 0; k < slots; ++k) {
        Slot& s = L[l * slots + k];
        if (s.q < 0 && next < nq) {  // FETCH
          const uint32_t q = order ? order[next] : (uint32_t)next;
          ++next;
          if (next == nq) dry_at = t;
          s.q = q;
          s.pos = offsets[q];
          s.end = offsets[q + 1];
          if (s.pos == s.end) s.q = -1;
        }
        if (s.q >= 0) want[states[s.pos] - BVS_NEED_INIT] = true;
      }
      for (int p = 0; p < 3; ++p) cnt[p] += want[p];
    }
    const int ni = cnt[0], nb = cnt[1], nl = cnt[2];
    if (ni + nb + nl == 0) {
      if (t > stats[0]) stats[0] = t;
      continue;
    }
    int phase;
    if (ni >= init_quorum || nb + nl == 0) phase = 0;
    else if (policy == 1) phase = (nb >= nl) ? 1 : 2;
    else phase = (nl >= nb) ? 2 : 1;
    for (int l = 0; l < 32; ++l)
      for (int k = 0; k < slots; ++k) {
        Slot& s = L[l * slots + k];
        if (s.q >= 0 && states[s.pos] - BVS_NEED_INIT == phase) {
          if (++s.pos == s.end) s.q = -1;
          stats[4 + phase] += 1;
          break;  // one step per lane and round
        }
      }
    stats[1 + phase] += 1;
    pq.push(Ev(t + cost[phase], w));
  }
  stats[7] = dry_at;
  return HFB_OK;
}

// shape pairs (no meshes) through lane groups of G threads, see batch_lanes
long emu_batch_distance_lanes(void* e, int G, size_t n, const uint32_t* h1, const hfb_transform* tf1,
                              const uint32_t* h2, const hfb_transform* tf2, const hfb_distance_request* req,
                              hfb_distance_result* out, const hfb_guess_out* go) {
  if (validate_query(req->q)) return -2;
  CollideP C;
  C.security_margin = 0;
  C.collision_distance_threshold = 0;
  return batch_lanes_g<0>(G, static_cast<Emu*>(e), n, h1, tf1, h2, tf2, req->q, solver_from_distance_request(*req), C,
                          out, go);
}
long emu_batch_collide_lanes(void* e, int G, size_t n, const uint32_t* h1, const hfb_transform* tf1,
                             const uint32_t* h2, const hfb_transform* tf2, const hfb_collision_request* req,
                             hfb_contact* out, const hfb_guess_out* go) {
  if (validate_query(req->q) || req->security_margin == -INFINITY || req->num_max_contacts == 0) return -2;
  CollideP C;
  C.security_margin = req->security_margin;
  C.collision_distance_threshold = req->q.collision_distance_threshold;
  return batch_lanes_g<1>(G, static_cast<Emu*>(e), n, h1, tf1, h2, tf2, req->q, solver_from_collision_request(*req), C,
                          out, go);
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

// support vertex of a point set along `dir` as each of the G lanes of a group computes it (lanesim)
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
    case 2: return lane_group_support<2>(s, d, idx_per_lane) ? HFB_OK : HFB_ERR_CUDA;
    case 4: return lane_group_support<4>(s, d, idx_per_lane) ? HFB_OK : HFB_ERR_CUDA;
    case 8: return lane_group_support<8>(s, d, idx_per_lane) ? HFB_OK : HFB_ERR_CUDA;
    case 16: return lane_group_support<16>(s, d, idx_per_lane) ? HFB_OK : HFB_ERR_CUDA;
    case 32: return lane_group_support<32>(s, d, idx_per_lane) ? HFB_OK : HFB_ERR_CUDA;
    default: return HFB_ERR_INVALID_ARGUMENT;
  }
  return HFB_OK;
}

}  // extern "C"
