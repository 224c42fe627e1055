// GroupNorm statistics folded into a per-(b,c) affine, and the plain affine + activation pass.
//
// torch.nn.GroupNorm(32, C, eps) sits in front of every ResnetBlock2D conv and every attention
// block.  Instead of writing a normalised copy of the activation, this build
// computes only the statistics (one read of x, 4 B/element) and hands the consumer kernel
// (filtered_act / affine_act) the folded affine  y = x*scale[b,c] + shift[b,c].
//
//   gn_partial : grid (chunks, B), one thread per channel, fp32 sum / sum-of-squares over a
//                chunk of <= 32 pixels (coalesced: consecutive threads = consecutive channels)
//   gn_finalize: one warp per (b, group): fp64 reduction of the partials, mean / rstd,
//                then scale/shift for the group's channels.
#include "common.cuh"

namespace afldm {
namespace {

constexpr int GN_PIX = 32;  // pixels per partial chunk

__global__ void __launch_bounds__(256)
gn_partial_kernel(const float* __restrict__ x, float* __restrict__ partial, int HW, int C, int nchunk) {
    const int chunk = blockIdx.x, b = blockIdx.y;
    const int p0 = chunk * GN_PIX;
    const int p1 = min(HW, p0 + GN_PIX);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float* xp = x + ((size_t)b * HW + p0) * C + c;
        float s = 0.f, q = 0.f;
#pragma unroll 8
        for (int p = p0; p < p1; ++p) {
            const float v = *xp;
            xp += C;
            s += v;
            q = fmaf(v, v, q);
        }
        float2* out = reinterpret_cast<float2*>(partial) + ((size_t)b * nchunk + chunk) * C + c;
        *out = make_float2(s, q);
    }
}

__global__ void __launch_bounds__(128)
gn_finalize_kernel(const float* __restrict__ partial, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float* __restrict__ scale, float* __restrict__ shift,
                   int B, int HW, int C, int groups, int nchunk, float eps) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= B * groups) return;
    const int b = warp / groups, g = warp % groups;
    const int cpg = C / groups;
    const float2* p = reinterpret_cast<const float2*>(partial) + (size_t)b * nchunk * C + g * cpg;
    double s = 0.0, q = 0.0;
    const int items = nchunk * cpg;
    for (int it = lane; it < items; it += 32) {
        const int ch = it / cpg, cc = it % cpg;
        const float2 v = p[(size_t)ch * C + cc];
        s += (double)v.x;
        q += (double)v.y;
    }
    s = warp_sum(s);
    q = warp_sum(q);
    const double n = (double)HW * (double)cpg;
    const double mean = s / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float fmean = (float)mean;
    for (int cc = lane; cc < cpg; cc += 32) {
        const int c = g * cpg + cc;
        const float ga = gamma != nullptr ? gamma[c] : 1.f;
        const float be = beta != nullptr ? beta[c] : 0.f;
        const float sc = ga * rstd;
        scale[(size_t)b * C + c] = sc;
        shift[(size_t)b * C + c] = fmaf(-fmean, sc, be);
    }
}

template <int ACT>
__global__ void __launch_bounds__(256)
affine_act_kernel(const float4* __restrict__ x, float4* __restrict__ y, long long n4, int C4,
                  long long per_image4, const float4* __restrict__ scale, const float4* __restrict__ shift) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v = x[i];
        if (scale != nullptr) {
            const long long b = i / per_image4;
            const int c4 = (int)(i % C4);
            const float4 sc = scale[b * C4 + c4], sh = shift[b * C4 + c4];
            v.x = fmaf(v.x, sc.x, sh.x);
            v.y = fmaf(v.y, sc.y, sh.y);
            v.z = fmaf(v.z, sc.z, sh.z);
            v.w = fmaf(v.w, sc.w, sh.w);
        }
        v.x = apply_act<ACT>(v.x);
        v.y = apply_act<ACT>(v.y);
        v.z = apply_act<ACT>(v.z);
        v.w = apply_act<ACT>(v.w);
        y[i] = v;
    }
}

inline int gn_chunks(int HW) { return ceil_div(HW, GN_PIX); }

}  // namespace
}  // namespace afldm

using namespace afldm;

extern "C" size_t afldm_groupnorm_scratch_floats(int B, int HW, int C) {
    if (B <= 0 || HW <= 0 || C <= 0) return 0;
    return (size_t)B * gn_chunks(HW) * C * 2;
}

extern "C" int afldm_groupnorm_affine_f32(const float* x, int B, int HW, int C, int groups, float eps,
                                          const float* gamma, const float* beta, float* scale,
                                          float* shift, float* partial, afldm_stream_t stream) {
    if (x == nullptr || scale == nullptr || shift == nullptr || partial == nullptr) return AFLDM_E_ARG;
    if (B <= 0 || HW <= 0 || C <= 0 || groups <= 0) return AFLDM_E_ARG;
    if (C % groups != 0) return AFLDM_E_SHAPE;
    if ((reinterpret_cast<uintptr_t>(partial) & 7u) != 0) return AFLDM_E_ARG;
    cudaStream_t st = as_stream(stream);
    const int nchunk = gn_chunks(HW);
    const int threads = C >= 256 ? 256 : ((C + 31) / 32) * 32;
    gn_partial_kernel<<<dim3(nchunk, B), threads, 0, st>>>(x, partial, HW, C, nchunk);
    const int warps = B * groups;
    gn_finalize_kernel<<<ceil_div(warps, 4), 128, 0, st>>>(partial, gamma, beta, scale, shift, B, HW, C,
                                                          groups, nchunk, eps);
    return launched(2);
}

extern "C" int afldm_affine_act_f32(const float* x, float* y, int B, int HW, int C, int act,
                                    const float* scale, const float* shift, afldm_stream_t stream) {
    if (x == nullptr || y == nullptr || B <= 0 || HW <= 0 || C <= 0) return AFLDM_E_ARG;
    if ((scale == nullptr) != (shift == nullptr)) return AFLDM_E_ARG;
    if (C % 4 != 0) return AFLDM_E_SHAPE;
    if (!aligned16(x) || !aligned16(y) || (scale && (!aligned16(scale) || !aligned16(shift)))) return AFLDM_E_ARG;
    const long long n4 = (long long)B * HW * C / 4;
    const long long per_image4 = (long long)HW * C / 4;
    const int blocks = (int)min((long long)148 * 8, (n4 + 255) / 256);
    cudaStream_t st = as_stream(stream);
    auto sc = reinterpret_cast<const float4*>(scale);
    auto sh = reinterpret_cast<const float4*>(shift);
    if (act == AFLDM_ACT_SILU)
        affine_act_kernel<AFLDM_ACT_SILU><<<blocks, 256, 0, st>>>(
            reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), n4, C / 4, per_image4, sc, sh);
    else if (act == AFLDM_ACT_IDENTITY)
        affine_act_kernel<AFLDM_ACT_IDENTITY><<<blocks, 256, 0, st>>>(
            reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), n4, C / 4, per_image4, sc, sh);
    else
        return AFLDM_E_ARG;
    return launched();
}
