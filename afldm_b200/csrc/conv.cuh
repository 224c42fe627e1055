// Internal interface between the convolution ABI (conv_api.cu) and its two back ends.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace afldm {

struct ConvPlan {
    int M, K, mtiles, ntiles, kchunks, splitk, chunks_per_split;
};

ConvPlan conv_simt_plan(int B, int H, int W, int Cin, int Cout, int ks);

int conv_simt_launch(const float* x, int x_pitch, const float* w, const float* bias,
                     const float* row_add, int row_add_pitch, const float* residual, int res_pitch, float* y,
                     int y_pitch, int B, int H, int W, int Cin, int Cout, int ks, float* workspace,
                     size_t workspace_floats, cudaStream_t st);

// y = sum_z ws[z] + bias + row_add + residual  (fixed summation order: deterministic split-K).
// Launches one kernel on st (the caller counts it).
void splitk_reduce_launch(const float* ws, int splitk, const float* bias, const float* row_add,
                          int row_add_pitch, const float* residual, int res_pitch, float* y, int y_pitch, int M, int Cout,
                          int HW, cudaStream_t st);

// tcgen05 / TMA path.  Returns false (and leaves *floats alone) when the shape is not covered.
bool conv_tc_workspace_floats(int B, int H, int W, int Cin, int Cout, int ks, size_t* floats);
int conv_tc_launch(const float* x, int x_pitch, const float* w, const float* bias, const float* row_add,
                   int row_add_pitch, const float* residual, int res_pitch, float* y, int y_pitch, int B, int H, int W,
                   int Cin, int Cout, int ks, float* workspace, size_t workspace_floats, cudaStream_t st);

}  // namespace afldm
