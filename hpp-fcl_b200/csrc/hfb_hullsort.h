// Second key of the pair-class sort for the hull / triangle class: the pairs of that class ordered by the handle of
// their first operand, so that the lane groups of a warp scan the SAME vertex block for shape 1 (one L1 transaction
// per load instead of one per pair).  Kernels in hfb_hullsort.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

// scratch bytes for n pairs
size_t hull_sort_bytes(size_t n);
// perm: the class-sorted index list (k_bin_scatter); offsets: first position of every class (device); [offsets[bin],
// offsets[bin + 1]) is re-ordered by h1.  Returns the re-ordered list (inside `ws`) through *sorted; the other classes
// keep their order.  cudaError_t as int.
int hull_sort_launch(const uint32_t* h1, unsigned n, const uint32_t* perm, const unsigned* offsets, int bin, void* ws,
                     const uint32_t** sorted, cudaStream_t s, int* launches);
