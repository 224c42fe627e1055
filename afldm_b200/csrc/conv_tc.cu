// tcgen05 (TF32, TMEM accumulators, TMA operand staging) implicit-GEMM convolution.
// Placeholder until the tensor-core path lands: every shape reports "no kernel".
#include "common.cuh"
#include "conv.cuh"

namespace afldm {

bool conv_tc_workspace_floats(int, int, int, int, int, int, size_t*) { return false; }

int conv_tc_launch(const float*, int, const float*, const float*, const float*, int, const float*, int, float*,
                   int, int, int, int, int, int, int, float*, size_t, cudaStream_t) {
    return AFLDM_E_NOKERNEL;
}

}  // namespace afldm
