// Host-side interface of the device broadphase (kernels in hfb_broadphase.cu; design: hfb_broadphase.cuh).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "../../include/hppfcl_b200.h"

namespace hfb {

// world-space boxes of n objects (6 doubles each) from the per-handle local boxes of the committed arena
int bp_scene_aabbs_launch(const double* d_local_aabbs, uint32_t nshapes, size_t n, const uint32_t* d_handles,
                          const hfb_transform* d_tfs, double* d_aabbs, cudaStream_t s);
// scratch the pair finder needs for n objects
size_t bp_scratch_bytes(size_t n);
// every pair i < j with overlapping boxes and i in [i_lo, i_hi) -> (d_first, d_second), at most `capacity` stored,
// all counted in *d_n_pairs; `scratch`: bp_scratch_bytes(n) bytes.  Returns a cudaError_t as int; *launches += kernels
// launched.
int bp_pairs_launch(size_t n, const double* d_aabbs, size_t i_lo, size_t i_hi, uint32_t* d_first, uint32_t* d_second,
                    size_t capacity, unsigned* d_n_pairs, void* scratch, int num_sms, cudaStream_t s, int* launches);

}  // namespace hfb
