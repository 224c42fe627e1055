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

// Rows of one image handled by one CTA of the split-K reduce == pixels per GroupNorm partial slot it emits.
constexpr int SPLITK_REDUCE_ROWS = 16;
inline int splitk_reduce_slots(int HW) { return (HW + SPLITK_REDUCE_ROWS - 1) / SPLITK_REDUCE_ROWS; }

// y = sum_z ws[z] + bias + row_add + residual  (fixed summation order: deterministic split-K).
// Launches one kernel on st (the caller counts it).
void splitk_reduce_launch(const float* ws, int splitk, const float* bias, const float* row_add,
                          int row_add_pitch, const float* residual, int res_pitch, float* y, int y_pitch, int M, int Cout,
                          int HW, float* gn_partial, cudaStream_t st);

// Whether the tcgen05 / TMA path covers this shape (x_half: fp16 operands, 64-channel stages).
bool conv_tc_supported(int B, int H, int W, int Cin, int Cout, int ks, int x_half = 0);
// host-side view of the launch plan: {TMEM columns, dynamic shared-memory bytes, CTAs, CTA-pair mode, BN, K splits, ring stages, halo mode}
bool conv_tc_plan_query(int B, int H, int W, int Cin, int Cout, int ks, int x_half, int out[8]);
// tcgen05 / TMA path.  Returns false (and leaves *floats alone) when the shape is not covered.
bool conv_tc_workspace_floats(int B, int H, int W, int Cin, int Cout, int ks, size_t* floats, int x_half = 0);
// Number of GroupNorm partial-sum slots per image the tensor-core path emits for this shape
// (gn_partial [B][slots][Cout] float2), 0 when it cannot (shape outside the family, ragged tiles).
int conv_tc_gn_slots(int B, int H, int W, int Cin, int Cout, int ks, int x_half = 0);
int conv_tc_launch(const float* x, int x_pitch, const float* w, const float* bias, const float* row_add,
                   int row_add_pitch, const float* residual, int res_pitch, float* y, int y_pitch, int B, int H, int W,
                   int Cin, int Cout, int ks, float* workspace, size_t workspace_floats, float* gn_partial,
                   cudaStream_t st, int y_half = 0, const float* x2 = nullptr, int x2_pitch = 0, int Cin1 = 0,
                   int x_half = 0);   // x_half: x / x2 / w are fp16 (tcgen05.mma.kind::f16), pitches in elements

}  // namespace afldm
