// Internal interface between the resampling ABI (resample.cu) and the large-plane back end.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace afldm {

// mode: 0 = filtered activation, 1 = up2, 2 = lpf + down2; n = side of the SMALL plane (64 or 128).
size_t resample_large_workspace_floats(int mode, int B, int n, int C);
int resample_large(int mode, int act, const float* x, float* y, int B, int n, int C, const float* scale,
                   const float* shift, float* ws, size_t ws_floats, cudaStream_t st, int y_half = 0);
// y_half: the pass that writes y stores IEEE binary16 (modes 0 and 1, tensor-core passes; AFLDM_E_NOKERNEL otherwise)

}  // namespace afldm
