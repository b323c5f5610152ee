// k_bvhq_prep / k_bvhq: mesh-shape distance queries as a per-block task system (design: hfb_bvhq.cuh).
//
// One persistent block of 8 warps per SM.  Shared memory holds the block's HFB_Q_NSLOTS queries in flight
// (QSlot: poses, the shape's RSS, result so far, counters), two item queues (rings of 32-bit items) and a
// free mask of the treelet buffers; the traversal stacks and the treelet value caches live in global memory
// (L1/L2-resident, written and read by the same SM).  Every warp loops: take up to 32 leaf items or up to 4
// bounding-volume items, run them, and let the lane that completed a query's last outstanding item continue
// that query's walk (q_advance) -- which pushes the next item(s) or retires the query and starts another.
#include <cfloat>

#include "hfb_bvhq_launch.h"

namespace hfb {

#define CAPS_BVHQ (CAP_PRIM | CAP_CONVEX | CAP_TRI | CAP_INLINE_PRIM)

struct QSched {
  int lhead, ltail, bhead, btail;
  int active;          // queries in flight in this block
  unsigned tl_free;    // free treelet buffers
  int abort_;
  int _pad;
  unsigned leafq[HFB_Q_QCAP];
  unsigned bvq[HFB_Q_QCAP];
};

__device__ __forceinline__ int vload(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

struct DevSink {
  QSched* sc;
  __device__ __forceinline__ void push(unsigned* buf, int* tail, unsigned item) {
    __threadfence_block();  // the slot / stack / treelet writes of this item's producer come first
    const int pos = atomicAdd(tail, 1) & (HFB_Q_QCAP - 1);
    volatile unsigned* q = buf;
    // the ring is larger than the number of items that can exist: the previous occupant was taken long ago
    for (unsigned spins = 0; q[pos] != 0u; ++spins)
      if (spins > (1u << 26)) {  // cannot happen; never hang the GPU over it
        atomicExch(&sc->abort_, 1);
        break;
      }
    q[pos] = item | HFB_Q_ITEM_VALID;
  }
  __device__ __forceinline__ void push_leaf(unsigned it) { push(sc->leafq, &sc->ltail, it); }
  __device__ __forceinline__ void push_bv(unsigned it) { push(sc->bvq, &sc->btail, it); }
  __device__ __forceinline__ int treelet_acquire() {
    unsigned m = *reinterpret_cast<volatile unsigned*>(&sc->tl_free);
    while (m) {
      const int b = __ffs(m) - 1;
      const unsigned old = atomicAnd(&sc->tl_free, ~(1u << b));
      if (old & (1u << b)) return b;
      m = old & ~(1u << b);
    }
    return -1;
  }
  __device__ __forceinline__ void treelet_release(int id) { atomicOr(&sc->tl_free, 1u << id); }
  __device__ __forceinline__ int dec_pending(QSlot& s) {
    __threadfence_block();
    const int old = atomicSub(&s.pending, 1);
    __threadfence_block();
    return old;
  }
};

// waits for the producer of a reserved ring entry (it stores right after reserving)
__device__ __forceinline__ unsigned q_take(volatile unsigned* q, int pos, QSched* sc) {
  unsigned item;
  for (unsigned spins = 0; ((item = q[pos]) & HFB_Q_ITEM_VALID) == 0u; ++spins)
    if (spins > (1u << 26)) {  // cannot happen; never hang the GPU over it
      atomicExch(&sc->abort_, 1);
      break;
    }
  return item;
}

// next query of the slice into slot `sl` (+ its seed leaf item); false when the slice is exhausted
__device__ bool q_fetch(const BvhqLaunch& L, unsigned lo, unsigned hi, QSlot& s, QStackEnt* stk, unsigned sl,
                        DevSink& sink) {
  for (;;) {
    const unsigned k = atomicAdd(L.work, 1u);
    if (k >= hi - lo) return false;
    const QPrep& pr = L.prep[k];
    if (!pr.ok) continue;  // unsupported pair: k_bvhq_prep wrote its record
    const unsigned i = L.index_list[lo + k];
    const xf t1 = load_xf(L.tf1[i].R), t2 = load_xf(L.tf2[i].R);
    v3 guess = mk(1, 0, 0);
    int h0 = 0, h1 = 0;
    if (L.P.initial_guess == HFB_GUESS_CACHED) {
      if (L.guess_in) guess = mk(L.guess_in[3 * i], L.guess_in[3 * i + 1], L.guess_in[3 * i + 2]);
      if (L.hint_in) {
        h0 = L.hint_in[2 * i];
        h1 = L.hint_in[2 * i + 1];
      }
    }
    BvhQuery q;
    bool swapped;
    bvh_make_query<CAPS_BVHQ>(L.A, L.h1[i], t1, L.h2[i], t2, q, swapped);
    q_start(s, stk, q, pr, i, guess, h0, h1);
    sink.push_leaf(sl);
    return true;
  }
}

// The RSS distances of the two children `pair[0]`, `pair[1]` of a node by the 8 lanes gbase .. gbase + 7 of the
// warp: lanes 0-3 of the group take child 0, lanes 4-7 child 1; lane u of a child evaluates the edge-pair cases
// u, 4 + u, 8 + u, 12 + u of rectDistance in turn until one of the child's four lanes has a passing case; the
// lowest passing case index is the reference's first return (RSS.cpp:121-713), its lane computes the distance.
__device__ __forceinline__ void q_bv_group(const QSlot& s, const hfb_bvh_node* pair, unsigned gmask, unsigned gbase,
                                           unsigned sub, double& d1, double& d2, int& f1, int& f2) {
  const unsigned child = sub >> 2, u = sub & 3u;
  const hfb_bvh_node& nd = pair[child];
  m3 R;
  v3 T;
  double b0, b1, rad;
  q_rss_operands(s, nd, R, T, b0, b1, rad);
  const int fc = nd.first_child;
  RectPre p;
  rect_prelude(R, T, s.sbv[12], s.sbv[13], b0, b1, p);
  int kmine = 16;
  bool sub_found = false;
#pragma unroll 1
  for (int t = 0; t < 4; ++t) {
    if (!sub_found && rect_case(p, 4 * t + (int)u)) kmine = 4 * t + (int)u;
    const unsigned b = __ballot_sync(gmask, kmine < 16) >> gbase;
    sub_found = ((b >> (4u * child)) & 0xfu) != 0u;
    if ((b & 0xfu) != 0u && (b & 0xf0u) != 0u) break;  // both children have their case (uniform over the group)
  }
  int kwin = kmine;
  kwin = min(kwin, __shfl_xor_sync(gmask, kwin, 1));
  kwin = min(kwin, __shfl_xor_sync(gmask, kwin, 2));
  double dist = 0;
  if (kwin < 16 ? (kmine == kwin) : (u == 0u)) {
    dist = rect_finish(p, kwin < 16 ? kwin : -1);
    dist -= rad;
    dist = (dist < 0.0) ? 0.0 : dist;
  }
  const unsigned src0 = gbase + (kwin < 16 ? (unsigned)(kwin & 3) : 0u);
  // every lane of a child asks its own child's winner; then the group leader collects both
  const double dc = __shfl_sync(gmask, dist, (int)(src0 + 4u * child));
  d1 = __shfl_sync(gmask, dc, (int)gbase);
  d2 = __shfl_sync(gmask, dc, (int)gbase + 4);
  f1 = __shfl_sync(gmask, fc, (int)gbase);
  f2 = __shfl_sync(gmask, fc, (int)gbase + 4);
}

__global__ void __launch_bounds__(128) k_bvhq_prep(const BvhqLaunch L) {
  const unsigned lo = *L.range_lo, hi = *L.range_hi;
  for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < hi - lo; k += gridDim.x * blockDim.x) {
    const unsigned i = L.index_list[lo + k];
    const xf t1 = load_xf(L.tf1[i].R), t2 = load_xf(L.tf2[i].R);
    BvhJob job;
    QPrep& pr = L.prep[k];
    if (!bvh_make_job<CAPS_BVHQ, 0>(L.A, L.h1[i], t1, L.h2[i], t2, L.B, mk(1, 0, 0), 0, 0, L.out + i, job)) {
      pr.ok = 0;
      continue;
    }
    q_make_prep(job.q, job.swapped, pr);
  }
}

__global__ void __launch_bounds__(HFB_Q_THREADS, 1) k_bvhq(const BvhqLaunch L) {
  extern __shared__ __align__(16) unsigned char q_smem[];
  QSlot* slots = reinterpret_cast<QSlot*>(q_smem);
  QSched* sc = reinterpret_cast<QSched*>(q_smem + HFB_Q_NSLOTS * sizeof(QSlot));
  const unsigned lo = *L.range_lo, hi = *L.range_hi;
  const unsigned lane = threadIdx.x & 31u;
  if (threadIdx.x == 0) {
    sc->lhead = sc->ltail = sc->bhead = sc->btail = 0;
    sc->active = 0;
    sc->tl_free = (1u << HFB_Q_NTREELETS) - 1u;
    sc->abort_ = 0;
  }
  for (unsigned k = threadIdx.x; k < HFB_Q_QCAP; k += blockDim.x) {
    sc->leafq[k] = 0u;
    sc->bvq[k] = 0u;
  }
  __syncthreads();
  QStackEnt* stacks = L.stacks + (size_t)blockIdx.x * HFB_Q_NSLOTS * (size_t)L.stack_cap;
  QTreelet* tls = L.treelets + (size_t)blockIdx.x * HFB_Q_NTREELETS;
  EpaWs* ws = L.ws + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  DevSink sink{sc};
  QCtx c;
  c.P = L.P;
  c.rel_err = L.B.rel_err;
  c.abs_err = L.B.abs_err;
  c.spec_after = (L.P.initial_guess == HFB_GUESS_CACHED) ? -1 : L.spec_after;
  unsigned long long bv_total = 0, leaf_total = 0;
  {  // thread t fills slot t
    const unsigned sl = threadIdx.x;
    if (sl < HFB_Q_NSLOTS && q_fetch(L, lo, hi, slots[sl], stacks + (size_t)sl * L.stack_cap, sl, sink))
      atomicAdd(&sc->active, 1);
  }
  __syncthreads();
  // what a lane does when q_advance / q_*_done reports the walk finished
  auto retire = [&](unsigned sl) {
    QSlot& s = slots[sl];
    q_write_result(s, L.out + s.pair);
    bv_total += (unsigned)s.bv_tests;
    leaf_total += (unsigned)s.leaf_tests;
    if (!q_fetch(L, lo, hi, s, stacks + (size_t)sl * L.stack_cap, sl, sink)) atomicSub(&sc->active, 1);
  };
  for (;;) {
    int kind = 0, cnt = 0, base = 0;
    if (lane == 0) {
      unsigned spins = 0;
      for (;;) {
        const int lh = vload(&sc->lhead), lt = vload(&sc->ltail), bh = vload(&sc->bhead), bt = vload(&sc->btail);
        const int nl = lt - lh, nb = bt - bh;
        if (nl >= 32 || (nb <= 0 && nl > 0)) {
          cnt = nl < 32 ? nl : 32;
          if (atomicCAS(&sc->lhead, lh, lh + cnt) == lh) {
            kind = 1;
            base = lh;
            break;
          }
        } else if (nb > 0) {
          cnt = nb < 4 ? nb : 4;
          if (atomicCAS(&sc->bhead, bh, bh + cnt) == bh) {
            kind = 2;
            base = bh;
            break;
          }
        } else {
          if (vload(&sc->active) <= 0 || vload(&sc->abort_)) {
            kind = -1;
            break;
          }
          __nanosleep(64);
          if (++spins > (1u << 24)) {  // watchdog: seconds without an item while queries are in flight
            atomicExch(&sc->abort_, 1);
            kind = -1;
            break;
          }
        }
      }
    }
    kind = __shfl_sync(0xffffffffu, kind, 0);
    cnt = __shfl_sync(0xffffffffu, cnt, 0);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (kind < 0) break;
    if (kind == 1) {  // up to 32 leaf items, one per lane
      if ((int)lane < cnt) {
        volatile unsigned* q = sc->leafq;
        const int pos = (base + (int)lane) & (HFB_Q_QCAP - 1);
        unsigned item = q_take(q, pos, sc);
        q[pos] = 0u;
        __threadfence_block();
        const bool valid = (item & HFB_Q_ITEM_VALID) != 0u;
        item &= ~HFB_Q_ITEM_VALID;
        const unsigned sl = item & HFB_Q_SLOT_MASK;
        QSlot& s = slots[sl];
        const bool spec = (item & HFB_Q_ITEM_SPEC) != 0u;
        if (valid) {
          QLeafRes r;
          q_leaf_eval<CAPS_BVHQ>(s, q_leaf_prim(s, tls, item), L.P, ws, !spec, r);
          if (q_leaf_done(s, item, stacks + (size_t)sl * L.stack_cap, tls, c, sink, r) == Q_DONE) retire(sl);
        }
      }
    } else {  // up to 4 bounding-volume items, 8 lanes each
      const unsigned g = lane >> 3, sub = lane & 7u;
      if ((int)g < cnt) {
        const unsigned gbase = g * 8u, gmask = 0xffu << gbase;
        volatile unsigned* q = sc->bvq;
        const int pos = (base + (int)g) & (HFB_Q_QCAP - 1);
        unsigned item = q_take(q, pos, sc);
        __syncwarp(gmask);  // every lane of the group has read the item before its leader clears the entry
        item = __shfl_sync(gmask, item, (int)gbase);
        if (sub == 0u) q[pos] = 0u;
        __threadfence_block();
        if (item & HFB_Q_ITEM_VALID) {  // (uniform over the group)
          item &= ~HFB_Q_ITEM_VALID;
          const unsigned sl = item & HFB_Q_SLOT_MASK;
          QSlot& s = slots[sl];
          const hfb_bvh_node* nodes = static_cast<const hfb_bvh_node*>(s.ptr[0]);
          double d1, d2;
          int f1, f2;
          q_bv_group(s, nodes + q_bv_base(s, item), gmask, gbase, sub, d1, d2, f1, f2);
          if (sub == 0u) {
            if (q_bv_done(s, item, stacks + (size_t)sl * L.stack_cap, tls, c, sink, d1, d2, f1, f2) == Q_DONE)
              retire(sl);
          }
        }
      }
    }
    __syncwarp();
  }
  if (bv_total) atomicAdd(L.counters, bv_total);
  if (leaf_total) atomicAdd(L.counters + 1, leaf_total);
  __syncthreads();
  if (threadIdx.x == 0 && sc->abort_ && sc->active > 0) atomicAdd(L.counters + 2, 1ull);
}

static size_t bvhq_smem_bytes() { return HFB_Q_NSLOTS * sizeof(QSlot) + sizeof(QSched); }

unsigned bvhq_blocks(int num_sms, size_t n) {
  // a block is worth launching for every 64 queries; never more than one per SM (persistent)
  size_t b = (n + 63) / 64;
  if (b > (size_t)num_sms) b = (size_t)num_sms;
  return b ? (unsigned)b : 1u;
}

BvhqSizes bvhq_sizes(unsigned blocks, size_t n, int stack_cap) {
  BvhqSizes z;
  z.prep = n * sizeof(QPrep);
  z.stacks = (size_t)blocks * HFB_Q_NSLOTS * (size_t)stack_cap * sizeof(QStackEnt);
  z.treelets = (size_t)blocks * HFB_Q_NTREELETS * sizeof(QTreelet);
  z.ws = (size_t)blocks * HFB_Q_THREADS * sizeof(EpaWs);
  return z;
}

int bvhq_launch(const BvhqLaunch& L, unsigned blocks, size_t n, cudaStream_t s) {
  static_assert(HFB_Q_NSLOTS <= HFB_Q_THREADS, "thread t fills slot t");
  static_assert(HFB_Q_NSLOTS + 32 * HFB_Q_NTREELETS <= HFB_Q_QCAP, "item rings hold every item that can exist");
  static_assert(sizeof(QSlot) % 16 == 8, "odd stride in 8-byte words: lanes reading one field of 32 slots spread over the banks");
  const size_t smem = bvhq_smem_bytes();
  cudaError_t e = cudaFuncSetAttribute(k_bvhq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  unsigned pb = (unsigned)((n + 127) / 128);
  if (pb > blocks * 16u) pb = blocks * 16u;
  if (pb == 0) pb = 1;
  k_bvhq_prep<<<pb, 128, 0, s>>>(L);
  k_bvhq<<<blocks, HFB_Q_THREADS, smem, s>>>(L);
  return (int)cudaGetLastError();
}

}  // namespace hfb
