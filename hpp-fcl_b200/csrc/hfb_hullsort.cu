// see hfb_hullsort.h
#include "hfb_hullsort.h"

#include <cub/cub.cuh>

namespace {

size_t up256(size_t b) { return (b + 255) & ~(size_t)255; }

// key of position k of the class-sorted list: 0 before the class, 1 << 24 | handle inside it, 2 << 24 after it -- a
// stable sort by this key moves nothing outside the class
__global__ void __launch_bounds__(256) k_hull_keys(const uint32_t* h1, unsigned n, const uint32_t* perm, const unsigned* offsets,
                                                   int bin, uint32_t* keys) {
  const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const unsigned lo = offsets[bin], hi = offsets[bin + 1];
  uint32_t key = 0u;
  if (k >= hi) key = 2u << 24;
  else if (k >= lo) key = (1u << 24) | (h1[perm[k]] & 0xffffffu);
  keys[k] = key;
}

}  // namespace

size_t hull_sort_bytes(size_t n) {
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (int)n, 0, 26);
  return 3 * up256(n * sizeof(uint32_t)) + up256(tmp) + 256;
}

int hull_sort_launch(const uint32_t* h1, unsigned n, const uint32_t* perm, const unsigned* offsets, int bin, void* ws,
                     const uint32_t** sorted, cudaStream_t s, int* launches) {
  unsigned char* c = static_cast<unsigned char*>(ws);
  uint32_t* keys = reinterpret_cast<uint32_t*>(c); c += up256((size_t)n * 4);
  uint32_t* keys_out = reinterpret_cast<uint32_t*>(c); c += up256((size_t)n * 4);
  uint32_t* vals_out = reinterpret_cast<uint32_t*>(c); c += up256((size_t)n * 4);
  size_t tmp = hull_sort_bytes(n) - 3 * up256((size_t)n * 4) - 256;
  k_hull_keys<<<(n + 255) / 256, 256, 0, s>>>(h1, n, perm, offsets, bin, keys);
  cudaError_t e = cub::DeviceRadixSort::SortPairs(c, tmp, keys, keys_out, perm, vals_out, (int)n, 0, 26, s);
  if (e != cudaSuccess) return (int)e;
  *sorted = vals_out;
  *launches += 5;
  return (int)cudaGetLastError();
}
