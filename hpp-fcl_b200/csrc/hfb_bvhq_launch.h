// Host-side interface of the task-system mesh-shape walk (kernels in hfb_bvhq.cu, logic in hfb_bvhq.cuh).
#pragma once
#include <cuda_runtime.h>

#include "hfb_bvhq.cuh"

namespace hfb {

#define HFB_Q_MAX_THREADS 512  // the kernel comes with 8 warps per block (255 registers per thread) and with 16 (128)
#define HFB_Q_NSLOTS 224     // queries in flight per block (shared memory: 224 slots + the treelets' hot values + the rings)
#define HFB_Q_NTREELETS 8    // speculated subtrees in flight per block
#define HFB_Q_QCAP 2048      // ring size of each item queue (>= NSLOTS + 2 * HFB_Q_TREELET_MAX * NTREELETS)

struct BvhqLaunch {
  ArenaView A;
  const uint32_t* h1;
  const hfb_transform* tf1;
  const uint32_t* h2;
  const hfb_transform* tf2;
  const double* guess_in;
  const int32_t* hint_in;
  hfb_distance_result* out;
  const uint32_t* index_list;  // class-sorted pair ids; the (mesh, shape) slice is [*range_lo, *range_hi)
  const unsigned* range_lo;
  const unsigned* range_hi;
  SolverP P;
  BvhReq B;
  QPrep* prep;           // one per query of the slice
  QStackEnt* stacks;     // blocks x NSLOTS x stack_cap
  QTreelet* treelets;    // blocks x NTREELETS
  EpaWs* ws;             // one per thread
  QLeafSave* saves;      // blocks x (NSLOTS + 2 * TREELET_MAX * NTREELETS): parked solver state of suspended leaf items
  unsigned* work;        // hand-out counter of this launch (zeroed by the caller)
  unsigned long long* counters;  // [0] bv tests, [1] leaf tests, [2] watchdog trips (running totals)
  int stack_cap;
  int spec_after;        // QCtx::spec_after / spec_big_after
  int spec_big_after;
  int bv_gens;           // generations of BV items a cycle runs before its leaf phase
  int warps;             // 8 or 16 warps per block
  int gjk_chunk;         // GJK iterations a leaf item runs before it parks its state and queues itself again
};

// scratch the launch needs for `blocks` blocks and up to n queries
struct BvhqSizes {
  size_t prep, stacks, treelets, ws, saves;
};
BvhqSizes bvhq_sizes(unsigned blocks, size_t n, int stack_cap);
unsigned bvhq_blocks(int num_sms, size_t n);
// enqueues the set-up pass and the walk on `s`; returns a cudaError_t as int (0 = ok)
int bvhq_launch(const BvhqLaunch& L, unsigned blocks, size_t n, cudaStream_t s);

}  // namespace hfb
