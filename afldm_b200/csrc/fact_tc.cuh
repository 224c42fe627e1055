// Internal interface between the resampling ABI (resample.cu) and the tcgen05 filtered-activation back end.
#pragma once
#include <cuda_runtime.h>

#include "fact_common.cuh"

namespace afldm {

// n = side of the plane.  Whether the tcgen05 kernel is selected for it (AFLDM_FACT_TC).
bool fact_tc_enabled(int n);
// Filtered activation of NHWC x [B][n][n][C] on tcgen05.  AFLDM_E_NOKERNEL when the shape / alignment is outside the
// kernel's family (n not in {16, 32}, C not a multiple of 256 / n, sources not 16-byte aligned): the caller then uses
// the mma.sync / FMA kernels.
int fact_tc_launch(int n, int act, const float* x, float* y, int B, int C, const Affine& af, cudaStream_t st);

}  // namespace afldm
