// Internal interface between the head/loss entry points (head_loss.cu) and the tensor-core
// forward of the segmentation head (head_tc.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "vmtl_common.cuh"

namespace vmtl {

// Writes one (loss sum, valid count) pair of doubles per launched CTA into `partial` (at most
// `max_blocks` CTAs, the count is returned in *grid_out) -- the layout ce_finalize consumes.
// VMTL_EUNSUPPORTED when the shape is outside the kernel's envelope (caller falls back).
int head_ce_tc_fwd(const float* feat, const float* W, const float* b, const int64_t* target, int64_t P, int C,
                   int64_t ignore_index, double* partial, int max_blocks, int* grid_out, uint8_t* pred,
                   int64_t* conf, cudaStream_t st);

// Backward (head_tc_bwd.cu): dfeat (optional) and one [CPAD][33] float partial (dW rows + db) per launched
// CTA, the layout head_ce_bwd_finalize consumes (CPAD = 16 / 20 / 32 bucket of C).
int head_ce_tc_bwd(const float* feat, const float* W, const float* b, const int64_t* target, int64_t P, int C,
                   int64_t ignore_index, const double* fwd_out, const float* gscale, float* dfeat, float* partial,
                   int max_blocks, int* grid_out, cudaStream_t st);

}  // namespace vmtl
