// C-ABI entry points of the convolution family: argument checks and algorithm dispatch.
#include "common.cuh"
#include "conv.cuh"

using namespace afldm;

namespace {
bool conv_bad_args(const float* x, int x_pitch, const float* w, const float* y, int y_pitch,
                   const float* row_add, int row_add_pitch,
                   const float* residual, int res_pitch, int B, int H, int W, int Cin, int Cout, int ks) {
    if (x == nullptr || w == nullptr || y == nullptr) return true;
    if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) return true;
    if (ks != 1 && ks != 3) return true;
    if (x_pitch < Cin || y_pitch < Cout) return true;
    if (residual != nullptr && res_pitch < Cout) return true;
    if (row_add != nullptr && row_add_pitch < Cout) return true;
    if (x == y) return true;
    return false;
}
}  // namespace

extern "C" size_t afldm_conv2d_workspace_floats(int B, int H, int W, int Cin, int Cout, int ksize, int algo) {
    if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || (ksize != 1 && ksize != 3)) return 0;
    if (algo == AFLDM_CONV_TCGEN05_TF32 || algo == AFLDM_CONV_TCGEN05_F16) {
        size_t need = 0;
        if (conv_tc_workspace_floats(B, H, W, Cin, Cout, ksize, &need, algo == AFLDM_CONV_TCGEN05_F16)) return need;
        if (algo == AFLDM_CONV_TCGEN05_F16) return 0;
        // shape not covered by the tensor-core path: callers fall back to SIMT
    }
    const ConvPlan p = conv_simt_plan(B, H, W, Cin, Cout, ksize);
    return p.splitk > 1 ? (size_t)p.splitk * p.M * Cout : 0;
}

extern "C" int afldm_conv2d_supported(int B, int H, int W, int Cin, int Cout, int ksize, int algo) {
    if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || (ksize != 1 && ksize != 3)) return 0;
    if (algo == AFLDM_CONV_SIMT_F32) return 1;
    if (algo == AFLDM_CONV_TCGEN05_TF32 || algo == AFLDM_CONV_TCGEN05_F16)
        return conv_tc_supported(B, H, W, Cin, Cout, ksize, algo == AFLDM_CONV_TCGEN05_F16) ? 1 : 0;
    return 0;
}

extern "C" int afldm_conv2d_plan(int B, int H, int W, int Cin, int Cout, int ksize, int algo, int* plan8) {
    if (plan8 == nullptr || B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || (ksize != 1 && ksize != 3)) return AFLDM_E_ARG;
    if (algo != AFLDM_CONV_TCGEN05_TF32 && algo != AFLDM_CONV_TCGEN05_F16) return AFLDM_E_NOKERNEL;
    return conv_tc_plan_query(B, H, W, Cin, Cout, ksize, algo == AFLDM_CONV_TCGEN05_F16, plan8) ? 0 : AFLDM_E_NOKERNEL;
}

extern "C" int afldm_conv2d_gn_slots(int B, int H, int W, int Cin, int Cout, int ksize, int algo) {
    if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || (ksize != 1 && ksize != 3)) return 0;
    if (algo != AFLDM_CONV_TCGEN05_TF32 && algo != AFLDM_CONV_TCGEN05_F16) return 0;
    return conv_tc_gn_slots(B, H, W, Cin, Cout, ksize, algo == AFLDM_CONV_TCGEN05_F16);
}

extern "C" int afldm_conv2d_f32(const float* x, int x_pitch, const float* w, const float* bias,
                                const float* row_add, int row_add_pitch, const float* residual, int res_pitch, float* y,
                                int y_pitch, int B, int H, int W, int Cin, int Cout, int ksize, int algo,
                                float* workspace, size_t workspace_floats, float* gn_partial,
                                afldm_stream_t stream) {
    if (conv_bad_args(x, x_pitch, w, y, y_pitch, row_add, row_add_pitch, residual, res_pitch, B, H, W, Cin, Cout, ksize))
        return AFLDM_E_ARG;
    cudaStream_t st = as_stream(stream);
    if (algo == AFLDM_CONV_SIMT_F32) {
        if (gn_partial != nullptr) return AFLDM_E_ARG;   // statistics are a tensor-core epilogue feature
        return conv_simt_launch(x, x_pitch, w, bias, row_add, row_add_pitch, residual, res_pitch, y, y_pitch, B, H, W, Cin,
                                Cout, ksize, workspace, workspace_floats, st);
    }
    if (algo == AFLDM_CONV_TCGEN05_TF32)
        return conv_tc_launch(x, x_pitch, w, bias, row_add, row_add_pitch, residual, res_pitch, y, y_pitch, B, H, W, Cin,
                              Cout, ksize, workspace, workspace_floats, gn_partial, st);
    return AFLDM_E_ARG;
}

extern "C" int afldm_conv2d_f16out(const float* x, int x_pitch, const float* w, const float* bias, void* y, int y_pitch,
                                   int B, int H, int W, int Cin, int Cout, int ksize, afldm_stream_t stream) {
    if (conv_bad_args(x, x_pitch, w, static_cast<const float*>(y), y_pitch, nullptr, 0, nullptr, 0, B, H, W, Cin, Cout,
                      ksize))
        return AFLDM_E_ARG;
    return conv_tc_launch(x, x_pitch, w, bias, nullptr, 0, nullptr, 0, static_cast<float*>(y), y_pitch, B, H, W, Cin, Cout,
                          ksize, nullptr, 0, nullptr, as_stream(stream), 1);
}

extern "C" int afldm_conv2d_f16in_f32(const void* x, int x_pitch, const void* w, const float* bias,
                                      const float* row_add, int row_add_pitch, const float* residual, int res_pitch,
                                      float* y, int y_pitch, int B, int H, int W, int Cin, int Cout, int ksize,
                                      float* workspace, size_t workspace_floats, float* gn_partial,
                                      afldm_stream_t stream) {
    const float* xf = static_cast<const float*>(x);
    const float* wf = static_cast<const float*>(w);
    if (conv_bad_args(xf, x_pitch, wf, y, y_pitch, row_add, row_add_pitch, residual, res_pitch, B, H, W, Cin, Cout, ksize))
        return AFLDM_E_ARG;
    return conv_tc_launch(xf, x_pitch, wf, bias, row_add, row_add_pitch, residual, res_pitch, y, y_pitch, B, H, W, Cin,
                          Cout, ksize, workspace, workspace_floats, gn_partial, as_stream(stream), 0, nullptr, 0, 0, 1);
}

extern "C" int afldm_conv2d_f16in_f16out(const void* x, int x_pitch, const void* w, const float* bias, void* y, int y_pitch,
                                         int B, int H, int W, int Cin, int Cout, int ksize, afldm_stream_t stream) {
    const float* xf = static_cast<const float*>(x);
    const float* wf = static_cast<const float*>(w);
    if (conv_bad_args(xf, x_pitch, wf, static_cast<const float*>(y), y_pitch, nullptr, 0, nullptr, 0, B, H, W, Cin, Cout, ksize))
        return AFLDM_E_ARG;
    return conv_tc_launch(xf, x_pitch, wf, bias, nullptr, 0, nullptr, 0, static_cast<float*>(y), y_pitch, B, H, W, Cin, Cout,
                          ksize, nullptr, 0, nullptr, as_stream(stream), 1, nullptr, 0, 0, 1);
}

extern "C" int afldm_conv2d_cat_f32(const float* xa, int xa_pitch, int Ca, const float* xb, int xb_pitch, int Cb,
                                    const float* w, const float* bias, const float* row_add, int row_add_pitch,
                                    const float* residual, int res_pitch, float* y, int y_pitch, int B, int H, int W,
                                    int Cout, int ksize, float* workspace, size_t workspace_floats, float* gn_partial,
                                    afldm_stream_t stream) {
    if (xb == nullptr || Ca <= 0 || Cb <= 0 || xa_pitch < Ca || xb_pitch < Cb) return AFLDM_E_ARG;
    if (conv_bad_args(xa, xa_pitch, w, y, y_pitch, row_add, row_add_pitch, residual, res_pitch, B, H, W, Ca, Cout, ksize))
        return AFLDM_E_ARG;
    if (xb == y) return AFLDM_E_ARG;
    return conv_tc_launch(xa, xa_pitch, w, bias, row_add, row_add_pitch, residual, res_pitch, y, y_pitch, B, H, W, Ca + Cb,
                          Cout, ksize, workspace, workspace_floats, gn_partial, as_stream(stream), 0, xb, xb_pitch, Ca);
}
