// Host-side interface of the GJK passes (kernels in hfb_gjkpass.cu).
#pragma once
#include <cuda_runtime.h>

#include "hfb_batch.cuh"

// bytes of solver state for n pairs
size_t gjk_pass_state_bytes(size_t n);
// bytes of `select_ws` (flags + scratch of the ordered compaction of the first pass's survivors)
size_t gjk_pass_select_bytes(size_t n);
// GJK over the class-sorted range [*a.range_lo, *a.range_hi) of a.index_list (at most n pairs) in `npass` passes of
// steps[0], steps[1], ... iterations (the last pass runs to convergence).  gjk_passes_first: the first pass and the
// extraction (result records, or EPA queue items) of the pairs it finished; gjk_passes_rest: the other passes and
// the extraction of theirs.  Between the two the caller may start EPA over the queue so far on a side stream.
// `counts`: npass words.  Return a cudaError_t as int.
int gjk_passes_first(const BatchArgs& a, int mode, unsigned n, void* state, uint32_t* list_a, unsigned* counts,
                     const int* steps, int npass, int num_sms, cudaStream_t s, int* launches, void* select_ws,
                     bool extract_now);
int gjk_passes_rest(const BatchArgs& a, int mode, unsigned n, void* state, uint32_t* list_a, uint32_t* list_b,
                    unsigned* counts, const int* steps, int npass, int num_sms, cudaStream_t s, int* launches,
                    bool first_extracted);
