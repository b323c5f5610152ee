// Device broadphase: scene boxes, uniform grid, pair sweep (design and the reference it stands in for:
// hfb_broadphase.cuh).  Everything stays on the device: the pair list feeds hfb_batch_*_objects_device directly.
#include <cfloat>

#include <cub/cub.cuh>

#include "hfb_broadphase.cuh"
#include "hfb_broadphase.h"

namespace hfb {

namespace {

constexpr unsigned kMaxCells = 2097152u;

__global__ void __launch_bounds__(256) k_scene_aabbs(const double* local, uint32_t nshapes, unsigned n,
                                                     const uint32_t* handles, const hfb_transform* tfs, double* out) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t h = handles[i];
  double* o = out + 6 * (size_t)i;
  if (h >= nshapes) {  // no such geometry: an empty box that overlaps nothing
    for (int k = 0; k < 3; ++k) {
      o[k] = DBL_MAX;
      o[3 + k] = -DBL_MAX;
    }
    return;
  }
  const double* l = local + 6 * (size_t)h;
  if (!(l[0] <= l[3])) {  // node type without a local box (NaN / inverted marker)
    for (int k = 0; k < 3; ++k) {
      o[k] = DBL_MAX;
      o[3 + k] = -DBL_MAX;
    }
    return;
  }
  object_aabb(l, l + 3, tfs[i], o);
}

struct Bounds {  // centres' box and the largest extent
  double lo[3], hi[3], ext;
};
__device__ __forceinline__ void bounds_merge(Bounds& a, const Bounds& b) {
  for (int k = 0; k < 3; ++k) {
    a.lo[k] = b.lo[k] < a.lo[k] ? b.lo[k] : a.lo[k];
    a.hi[k] = b.hi[k] > a.hi[k] ? b.hi[k] : a.hi[k];
  }
  a.ext = b.ext > a.ext ? b.ext : a.ext;
}
__device__ __forceinline__ Bounds bounds_empty() {
  Bounds b;
  for (int k = 0; k < 3; ++k) {
    b.lo[k] = DBL_MAX;
    b.hi[k] = -DBL_MAX;
  }
  b.ext = 0;
  return b;
}
__device__ Bounds block_bounds(Bounds v) {
  __shared__ Bounds sh[8];
  for (int off = 16; off > 0; off >>= 1) {
    Bounds o;
    for (int k = 0; k < 3; ++k) {
      o.lo[k] = __shfl_xor_sync(0xffffffffu, v.lo[k], off);
      o.hi[k] = __shfl_xor_sync(0xffffffffu, v.hi[k], off);
    }
    o.ext = __shfl_xor_sync(0xffffffffu, v.ext, off);
    bounds_merge(v, o);
  }
  if ((threadIdx.x & 31u) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : bounds_empty();
    for (int off = 4; off > 0; off >>= 1) {
      Bounds o;
      for (int k = 0; k < 3; ++k) {
        o.lo[k] = __shfl_xor_sync(0xffffffffu, v.lo[k], off);
        o.hi[k] = __shfl_xor_sync(0xffffffffu, v.hi[k], off);
      }
      o.ext = __shfl_xor_sync(0xffffffffu, v.ext, off);
      bounds_merge(v, o);
    }
  }
  return v;  // valid in thread 0
}
__global__ void __launch_bounds__(256) k_bounds_partial(const double* bb, unsigned n, Bounds* part) {
  Bounds v = bounds_empty();
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double* b = bb + 6 * (size_t)i;
    if (aabb_empty(b) || aabb_unbounded(b)) continue;  // no box / a halfspace or plane: not in the grid
    Bounds o;
    for (int k = 0; k < 3; ++k) {
      o.lo[k] = o.hi[k] = 0.5 * (b[k] + b[3 + k]);
      const double e = b[3 + k] - b[k];
      o.ext = (k == 0 || e > o.ext) ? e : o.ext;
    }
    bounds_merge(v, o);
  }
  v = block_bounds(v);
  if (threadIdx.x == 0) part[blockIdx.x] = v;
}
// one block: final merge, then the grid (make_grid of hfb_broadphase.cuh) and the zeroed pair counter
__global__ void __launch_bounds__(256) k_make_grid(const Bounds* part, unsigned nparts, BroadGrid* grid, unsigned* n_pairs,
                                                   unsigned* n_unbounded) {
  Bounds v = bounds_empty();
  for (unsigned i = threadIdx.x; i < nparts; i += blockDim.x) bounds_merge(v, part[i]);
  v = block_bounds(v);
  if (threadIdx.x == 0) {
    BroadGrid g;
    if (!(v.lo[0] <= v.hi[0])) {  // no boxes at all
      for (int k = 0; k < 3; ++k) {
        v.lo[k] = v.hi[k] = 0;
      }
    }
    double cell = v.ext > 0 ? v.ext : 1.0;
    for (;;) {
      double cells = 1;
      for (int k = 0; k < 3; ++k) {
        g.dim[k] = (int)floor((v.hi[k] - v.lo[k]) / cell) + 1;
        cells *= g.dim[k];
      }
      if (cells <= (double)kMaxCells) break;
      cell *= 1.26;
    }
    for (int k = 0; k < 3; ++k) g.origin[k] = v.lo[k];
    g.inv_cell = 1.0 / cell;
    *grid = g;
    *n_pairs = 0u;
    *n_unbounded = 0u;
  }
}
// cell of every object; the unbounded ones (cell 0xfffffffe) are listed from the END of `order` backwards
__global__ void __launch_bounds__(256) k_cell_hist(const double* bb, unsigned n, const BroadGrid* grid, unsigned* cell,
                                                   unsigned* count, unsigned* order, unsigned* n_unbounded) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* b = bb + 6 * (size_t)i;
  if (aabb_empty(b)) {
    cell[i] = 0xffffffffu;  // empty box: in no cell
    return;
  }
  if (aabb_unbounded(b)) {
    cell[i] = 0xfffffffeu;
    order[n - 1u - atomicAdd(n_unbounded, 1u)] = i;
    return;
  }
  const unsigned c = grid_cell(*grid, b);
  cell[i] = c;
  atomicAdd(count + c, 1u);
}
__global__ void __launch_bounds__(256) k_cell_scatter(const unsigned* cell, unsigned n, const unsigned* start,
                                                      unsigned* cursor, unsigned* order) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || cell[i] >= 0xfffffffeu) return;
  order[start[cell[i]] + atomicAdd(cursor + cell[i], 1u)] = i;
}
// one object per thread: the 27 cells around its own, partners with a larger index
__global__ void __launch_bounds__(128) k_sweep(const double* bb, unsigned n, unsigned i_lo, const BroadGrid* grid,
                                               const unsigned* cell, const unsigned* start, const unsigned* order,
                                               uint32_t* first, uint32_t* second, unsigned capacity, unsigned* n_pairs) {
  // n: one past the last object this launch reports pairs for (as their smaller index); i_lo: the first one
  const unsigned i = i_lo + blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31u;
  const BroadGrid g = *grid;
  const bool live = i < n && cell[i] < 0xfffffffeu;
  double mine[6] = {0, 0, 0, 0, 0, 0};
  int cx = 0, cy = 0, cz = 0;
  if (live) {
    for (int k = 0; k < 6; ++k) mine[k] = bb[6 * (size_t)i + k];
    const unsigned c = cell[i];
    cx = (int)(c % (unsigned)g.dim[0]);
    cy = (int)((c / (unsigned)g.dim[0]) % (unsigned)g.dim[1]);
    cz = (int)(c / ((unsigned)g.dim[0] * (unsigned)g.dim[1]));
  }
  for (int d = 0; d < 27; ++d) {
    const int x = cx + d % 3 - 1, y = cy + (d / 3) % 3 - 1, z = cz + d / 9 - 1;
    const bool in = live && x >= 0 && y >= 0 && z >= 0 && x < g.dim[0] && y < g.dim[1] && z < g.dim[2];
    unsigned q = 0, qe = 0;
    if (in) {
      const size_t nc = ((size_t)z * g.dim[1] + y) * g.dim[0] + x;
      q = start[nc];
      qe = start[nc + 1];
    }
    // the lanes of a warp walk their cells side by side; hits are appended with one counter bump per warp and round
    while (__any_sync(0xffffffffu, q < qe)) {
      bool hit = false;
      unsigned j = 0;
      if (q < qe) {
        j = order[q];
        ++q;
        hit = j > i && aabb_overlap(mine, bb + 6 * (size_t)j);
      }
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m) {
        unsigned base = 0;
        if (lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(n_pairs, (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        const unsigned pos = base + (unsigned)__popc(m & ((1u << lane) - 1u));
        if (hit && pos < capacity) {
          first[pos] = i;
          second[pos] = j;
        }
      }
    }
  }
}

// the unbounded objects (halfspaces, planes) against every object: thread per object j, loop over the few unbounded u;
// the pair is this launch's when its smaller index lies in [i_lo, i_hi), a pair of two unbounded objects is reported
// from its smaller index only
__global__ void __launch_bounds__(128) k_sweep_unbounded(const double* bb, unsigned n, unsigned i_lo, unsigned i_hi,
                                                         const unsigned* cell, const unsigned* order, const unsigned* n_unbounded,
                                                         uint32_t* first, uint32_t* second, unsigned capacity, unsigned* n_pairs) {
  const unsigned nu = *n_unbounded;
  if (nu == 0) return;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned j0 = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned jend = (n + 31u) & ~31u;  // (warp-uniform trip count for the ballots)
  for (unsigned j = j0; j < jend; j += gridDim.x * blockDim.x) {
    const bool jl = j < n && cell[j] != 0xffffffffu;
    double mine[6] = {0, 0, 0, 0, 0, 0};
    if (jl)
      for (int k = 0; k < 6; ++k) mine[k] = bb[6 * (size_t)j + k];
    for (unsigned q = 0; q < nu; ++q) {
      const unsigned u = order[n - 1u - q];
      const unsigned lo = u < j ? u : j, hi = u < j ? j : u;
      const bool hit = jl && j != u && !(cell[j] == 0xfffffffeu && j < u) && lo >= i_lo && lo < i_hi &&
                       aabb_overlap(mine, bb + 6 * (size_t)u);
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m) {
        unsigned base = 0;
        if (lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(n_pairs, (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        const unsigned pos = base + (unsigned)__popc(m & ((1u << lane) - 1u));
        if (hit && pos < capacity) {
          first[pos] = lo;
          second[pos] = hi;
        }
      }
    }
  }
}

size_t up256(size_t b) { return (b + 255) & ~(size_t)255; }
struct Scratch {
  Bounds* part;
  BroadGrid* grid;
  unsigned* n_unbounded;
  unsigned* cell;
  unsigned* order;
  unsigned* count;   // kMaxCells + 1 (count, then exclusive scan in place via `start`)
  unsigned* start;   // kMaxCells + 1
  unsigned* cursor;  // kMaxCells
  void* cub_tmp;
  size_t cub_bytes;
};
constexpr unsigned kParts = 592;
size_t cub_scan_bytes() {
  size_t b = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, b, (unsigned*)nullptr, (unsigned*)nullptr, (int)(kMaxCells + 1));
  return b;
}
Scratch carve(void* p, size_t n) {
  unsigned char* c = static_cast<unsigned char*>(p);
  Scratch s;
  s.part = reinterpret_cast<Bounds*>(c); c += up256(kParts * sizeof(Bounds));
  s.grid = reinterpret_cast<BroadGrid*>(c); c += up256(sizeof(BroadGrid));
  s.n_unbounded = reinterpret_cast<unsigned*>(c); c += 256;
  s.cell = reinterpret_cast<unsigned*>(c); c += up256(n * 4);
  s.order = reinterpret_cast<unsigned*>(c); c += up256(n * 4);
  s.count = reinterpret_cast<unsigned*>(c); c += up256((size_t)(kMaxCells + 1) * 4);
  s.start = reinterpret_cast<unsigned*>(c); c += up256((size_t)(kMaxCells + 1) * 4);
  s.cursor = reinterpret_cast<unsigned*>(c); c += up256((size_t)kMaxCells * 4);
  s.cub_tmp = c;
  s.cub_bytes = cub_scan_bytes();
  return s;
}

}  // namespace

int bp_scene_aabbs_launch(const double* d_local_aabbs, uint32_t nshapes, size_t n, const uint32_t* d_handles,
                          const hfb_transform* d_tfs, double* d_aabbs, cudaStream_t s) {
  if (n == 0) return 0;
  k_scene_aabbs<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_local_aabbs, nshapes, (unsigned)n, d_handles, d_tfs, d_aabbs);
  return (int)cudaGetLastError();
}

size_t bp_scratch_bytes(size_t n) {
  return up256(kParts * sizeof(Bounds)) + up256(sizeof(BroadGrid)) + 256 + 2 * up256(n * 4) +
         2 * up256((size_t)(kMaxCells + 1) * 4) + up256((size_t)kMaxCells * 4) + up256(cub_scan_bytes()) + 256;
}

int bp_pairs_launch(size_t n, const double* d_aabbs, size_t i_lo, size_t i_hi, uint32_t* d_first, uint32_t* d_second,
                    size_t capacity, unsigned* d_n_pairs, void* scratch, int num_sms, cudaStream_t s, int* launches) {
  if (num_sms < 1) num_sms = 1;
  Scratch w = carve(scratch, n);
  const unsigned nn = (unsigned)n;
  unsigned pb = (nn + 255) / 256;
  if (pb > kParts) pb = kParts;
  if (pb == 0) pb = 1;
  cudaError_t e;
  k_bounds_partial<<<pb, 256, 0, s>>>(d_aabbs, nn, w.part);
  k_make_grid<<<1, 256, 0, s>>>(w.part, pb, w.grid, d_n_pairs, w.n_unbounded);
  *launches += 2;
  if (n < 2) return (int)cudaGetLastError();
  // the number of cells is only known on the device: counters are cleared and scanned for the maximum
  if ((e = cudaMemsetAsync(w.count, 0, (size_t)(kMaxCells + 1) * 4, s)) != cudaSuccess) return (int)e;
  if ((e = cudaMemsetAsync(w.cursor, 0, (size_t)kMaxCells * 4, s)) != cudaSuccess) return (int)e;
  k_cell_hist<<<(nn + 255) / 256, 256, 0, s>>>(d_aabbs, nn, w.grid, w.cell, w.count, w.order, w.n_unbounded);
  size_t tb = w.cub_bytes;
  if ((e = cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.count, w.start, (int)(kMaxCells + 1), s)) != cudaSuccess) return (int)e;
  k_cell_scatter<<<(nn + 255) / 256, 256, 0, s>>>(w.cell, nn, w.start, w.cursor, w.order);
  if (i_hi > n) i_hi = n;
  if (i_lo < i_hi)
    k_sweep<<<(unsigned)((i_hi - i_lo + 127) / 128), 128, 0, s>>>(d_aabbs, (unsigned)i_hi, (unsigned)i_lo, w.grid, w.cell, w.start,
                                                                 w.order, d_first, d_second,
                                                                 (unsigned)(capacity > 0xffffffffull ? 0xffffffffull : capacity),
                                                                 d_n_pairs);
  if (i_lo < i_hi) {
    unsigned ub = (nn + 127) / 128;
    if (ub > (unsigned)num_sms * 8u) ub = (unsigned)num_sms * 8u;
    k_sweep_unbounded<<<ub ? ub : 1u, 128, 0, s>>>(d_aabbs, nn, (unsigned)i_lo, (unsigned)i_hi, w.cell, w.order, w.n_unbounded, d_first,
                                                   d_second, (unsigned)(capacity > 0xffffffffull ? 0xffffffffull : capacity), d_n_pairs);
  }
  *launches += 5;
  return (int)cudaGetLastError();
}

}  // namespace hfb
