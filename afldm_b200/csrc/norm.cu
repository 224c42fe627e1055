// GroupNorm statistics folded into a per-(b,c) affine, and the plain affine + activation pass.
//
// torch.nn.GroupNorm(32, C, eps) sits in front of every ResnetBlock2D conv and every attention
// block.  Instead of writing a normalised copy of the activation, this build
// computes only the statistics (one read of x, 4 B/element) and hands the consumer kernel
// (filtered_act / affine_act) the folded affine  y = x*scale[b,c] + shift[b,c].
//
//   gn_stats_kernel: one CTA per (b, group).  Threads stride over the group's HW x cpg elements
//                (consecutive threads = consecutive channels of a pixel, then the next pixel), fp32
//                partial sums per thread, fp64 block reduction, then the group's channels get their
//                scale / shift.  One launch, no scratch, fixed reduction order (deterministic).
#include <algorithm>

#include <cuda_fp16.h>

#include "common.cuh"

namespace afldm {
namespace {

__global__ void __launch_bounds__(256)
gn_stats_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                float* __restrict__ scale, float* __restrict__ shift, int HW, int C, int groups, float eps) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.x / groups, g = blockIdx.x % groups;
    const int cpg = C / groups;
    const int total = HW * cpg;
    const float* xb = x + (size_t)b * HW * C + g * cpg;
    float s = 0.f, q = 0.f;
    int e = threadIdx.x;
    // 4 independent loads in flight per thread
    for (; e + 3 * 256 < total; e += 4 * 256) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int ee = e + u * 256;
            const int p = ee / cpg, cc = ee - p * cpg;
            v[u] = xb[(size_t)p * C + cc];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            s += v[u];
            q = fmaf(v[u], v[u], q);
        }
    }
    for (; e < total; e += 256) {
        const int p = e / cpg, cc = e - p * cpg;
        const float v = xb[(size_t)p * C + cc];
        s += v;
        q = fmaf(v, v, q);
    }
    double ds = warp_sum((double)s), dq = warp_sum((double)q);
    __shared__ double red[2][8];
    __shared__ float stat[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        red[0][warp] = ds;
        red[1][warp] = dq;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0.0, tq = 0.0;
        for (int w = 0; w < 8; ++w) {
            ts += red[0][w];
            tq += red[1][w];
        }
        const double n = (double)total;
        const double mean = ts / n;
        double var = tq / n - mean * mean;
        if (var < 0.0) var = 0.0;
        stat[0] = (float)mean;
        stat[1] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    const float fmean = stat[0], rstd = stat[1];
    for (int cc = threadIdx.x; cc < cpg; cc += 256) {
        const int c = g * cpg + cc;
        const float ga = gamma != nullptr ? gamma[c] : 1.f;
        const float be = beta != nullptr ? beta[c] : 0.f;
        const float sc = ga * rstd;
        scale[(size_t)b * C + c] = sc;
        shift[(size_t)b * C + c] = fmaf(-fmean, sc, be);
    }
}

// GroupNorm finalize from partial sums that the producing kernels emitted (conv epilogue / split-K
// reduce): partial [B][slots][Cx] float2 per source; channels [0, Ca) come from source a, the rest from
// source b (the two halves of a skip-connection concat).  One warp per (b, group), fp64 combine.
__global__ void __launch_bounds__(128)
gn_finalize_kernel(const float2* __restrict__ pa, int slots_a, int Ca, const float2* __restrict__ pb, int slots_b,
                   int Cb, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ scale, float* __restrict__ shift, int B, int HW, int groups, float eps) {
    pdl_trigger();
    pdl_wait();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B * groups) return;
    const int C = Ca + Cb;
    const int b = warp / groups, g = warp % groups;
    const int cpg = C / groups;
    double s = 0.0, q = 0.0;
    // lanes stride over the group's (slot, channel) partials: channel fastest = contiguous float2 loads
    const int slots_max = max(slots_a, slots_b);
    for (int it = lane; it < slots_max * cpg; it += 32) {
        const int sl = it / cpg, c = g * cpg + (it - sl * cpg);
        float2 v = make_float2(0.f, 0.f);
        if (c < Ca) {
            if (sl < slots_a) v = pa[((size_t)b * slots_a + sl) * Ca + c];
        } else {
            if (sl < slots_b) v = pb[((size_t)b * slots_b + sl) * Cb + (c - Ca)];
        }
        s += (double)v.x;
        q += (double)v.y;
    }
    s = warp_sum(s);
    q = warp_sum(q);
    const double n = (double)HW * (double)cpg;
    const double mean = s / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps)), fmean = (float)mean;
    for (int cc = lane; cc < cpg; cc += 32) {
        const int c = g * cpg + cc;
        const float sc = (gamma != nullptr ? gamma[c] : 1.f) * rstd;
        scale[(size_t)b * C + c] = sc;
        shift[(size_t)b * C + c] = fmaf(-fmean, sc, beta != nullptr ? beta[c] : 0.f);
    }
}

template <int ACT>
__global__ void __launch_bounds__(256)
affine_act_kernel(const float4* __restrict__ x, float4* __restrict__ y, long long n4, int C4,
                  long long per_image4, const float4* __restrict__ scale, const float4* __restrict__ shift, int y_half) {
    pdl_trigger();
    pdl_wait();
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
        if (y_half) {
            const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t*>(&h0);
            pk.y = *reinterpret_cast<const uint32_t*>(&h1);
            reinterpret_cast<uint2*>(y)[i] = pk;
        } else {
            y[i] = v;
        }
    }
}

// The same pass for channel counts that are no multiple of 4 (or unaligned pointers): one element per thread.  Only the
// general-plane form of the ideal resamplers and odd-channel callers reach it.
template <int ACT>
__global__ void __launch_bounds__(256)
affine_act_scalar_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, int C, long long per_image,
                         const float* __restrict__ scale, const float* __restrict__ shift, int y_half) {
    pdl_trigger();
    pdl_wait();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float v = x[i];
        if (scale != nullptr) {
            const long long b = i / per_image;
            const int c = (int)(i % C);
            v = fmaf(v, scale[b * C + c], shift[b * C + c]);
        }
        v = apply_act<ACT>(v);
        if (y_half) reinterpret_cast<__half*>(y)[i] = __float2half_rn(v);
        else y[i] = v;
    }
}

// GroupNorm finalisation + y = act(x * scale + shift) in ONE launch (normalised input of an attention block):
// every CTA first turns the producer's partial sums of ITS image into per-channel scale / shift in shared memory
// (one warp per group, fp32 adds, fp64 only for E[x^2] - mean^2; a few KB of L2-resident loads), then streams its
// slice of the image.  Replaces gn_finalize_kernel + affine_act_kernel (two launches and their gap).
template <int ACT>
__global__ void __launch_bounds__(1024)
affine_act_gn_kernel(const float4* __restrict__ x, float4* __restrict__ y, int HW, int C,
                     const float2* __restrict__ pa, int slots_a, int Ca, const float2* __restrict__ pb, int slots_b,
                     int Cb, const float* __restrict__ gamma, const float* __restrict__ beta, int groups, float eps,
                     double inv_n, int y_half) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float s_aff[];                  // [C] scale | [C] shift | [groups] mean | [groups] rstd
    float* s_sc = s_aff;
    float* s_sh = s_aff + C;
    float* s_mean = s_aff + 2 * C;
    float* s_rstd = s_mean + groups;
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int cpg = C / groups;
    const int slots_max = max(slots_a, slots_b);
    const int total = slots_max * cpg;
    for (int g = warp; g < groups; g += nwarps) {
        float s = 0.f, q = 0.f;
        for (int it0 = lane; it0 < total; it0 += 128) {
            float2 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int it = it0 + 32 * u;
                const int sl = it / cpg, cc = g * cpg + (it - sl * cpg);
                v[u] = make_float2(0.f, 0.f);
                if (it < total) {
                    if (cc < Ca) {
                        if (sl < slots_a) v[u] = __ldg(&pa[((size_t)b * slots_a + sl) * Ca + cc]);
                    } else {
                        if (sl < slots_b) v[u] = __ldg(&pb[((size_t)b * slots_b + sl) * Cb + (cc - Ca)]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                s += v[u].x;
                q += v[u].y;
            }
        }
        s = warp_sum(s);
        q = warp_sum(q);
        if (lane == 0) {
            const double mean = (double)s * inv_n;
            double var = (double)q * inv_n - mean * mean;
            if (var < 0.0) var = 0.0;
            s_mean[g] = (float)mean;
            s_rstd[g] = rsqrtf((float)var + eps);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        const float sc = (gamma != nullptr ? gamma[c] : 1.f) * s_rstd[g];
        s_sc[c] = sc;
        s_sh[c] = fmaf(-s_mean[g], sc, beta != nullptr ? beta[c] : 0.f);
    }
    __syncthreads();
    const int C4 = C >> 2;
    const long long per_image4 = (long long)HW * C4;
    const float4* xb = x + (size_t)b * per_image4;
    float4* yb = y + (size_t)b * per_image4;
    const float4* sc4 = reinterpret_cast<const float4*>(s_sc);
    const float4* sh4 = reinterpret_cast<const float4*>(s_sh);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_image4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = xb[i];
        const int c4 = (int)(i % C4);
        const float4 sc = sc4[c4], sh = sh4[c4];
        v.x = apply_act<ACT>(fmaf(v.x, sc.x, sh.x));
        v.y = apply_act<ACT>(fmaf(v.y, sc.y, sh.y));
        v.z = apply_act<ACT>(fmaf(v.z, sc.z, sh.z));
        v.w = apply_act<ACT>(fmaf(v.w, sc.w, sh.w));
        if (y_half) {
            // fp16 result (the A operand of the q | k | v projection): 8-byte stores, same element indexing
            const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t*>(&h0);
            pk.y = *reinterpret_cast<const uint32_t*>(&h1);
            reinterpret_cast<uint2*>(y)[(size_t)b * per_image4 + i] = pk;
        } else {
            yb[i] = v;
        }
    }
}

}  // namespace
}  // namespace afldm

using namespace afldm;

extern "C" size_t afldm_groupnorm_scratch_floats(int, int, int) { return 0; }

extern "C" int afldm_groupnorm_affine_f32(const float* x, int B, int HW, int C, int groups, float eps,
                                          const float* gamma, const float* beta, float* scale,
                                          float* shift, float* /*partial: unused in this build*/,
                                          afldm_stream_t stream) {
    if (x == nullptr || scale == nullptr || shift == nullptr) return AFLDM_E_ARG;
    if (B <= 0 || HW <= 0 || C <= 0 || groups <= 0) return AFLDM_E_ARG;
    if (C % groups != 0) return AFLDM_E_SHAPE;
    if ((long long)HW * (C / groups) > 0x7fffffffLL) return AFLDM_E_SHAPE;
    launch_k(gn_stats_kernel, dim3(B * groups), dim3(256), 0, as_stream(stream), x, gamma, beta, scale, shift, HW, C, groups, eps);
    return launched();
}

extern "C" int afldm_groupnorm_finalize_f32(const float* partial_a, int slots_a, int Ca, const float* partial_b,
                                            int slots_b, int Cb, int B, int HW, int groups, float eps,
                                            const float* gamma, const float* beta, float* scale, float* shift,
                                            afldm_stream_t stream) {
    if (partial_a == nullptr || scale == nullptr || shift == nullptr) return AFLDM_E_ARG;
    if (B <= 0 || HW <= 0 || Ca <= 0 || slots_a <= 0 || groups <= 0 || Cb < 0) return AFLDM_E_ARG;
    if (Cb > 0 && (partial_b == nullptr || slots_b <= 0)) return AFLDM_E_ARG;
    if ((Ca + Cb) % groups != 0) return AFLDM_E_SHAPE;
    const int warps = B * groups;
    launch_k(gn_finalize_kernel, dim3(ceil_div(warps, 4)), dim3(128), 0, as_stream(stream),
             reinterpret_cast<const float2*>(partial_a), slots_a, Ca, reinterpret_cast<const float2*>(partial_b),
             slots_b, Cb, gamma, beta, scale, shift, B, HW, groups, eps);
    return launched();
}

static int affine_act_impl(const float* x, float* y, int y_half, int B, int HW, int C, int act,
                           const float* scale, const float* shift, afldm_stream_t stream) {
    if (x == nullptr || y == nullptr || B <= 0 || HW <= 0 || C <= 0) return AFLDM_E_ARG;
    if ((scale == nullptr) != (shift == nullptr)) return AFLDM_E_ARG;
    if (act != AFLDM_ACT_SILU && act != AFLDM_ACT_IDENTITY) return AFLDM_E_ARG;
    if (C % 4 != 0 || !aligned16(x) || !aligned16(y) || (scale && (!aligned16(scale) || !aligned16(shift)))) {
        const long long n = (long long)B * HW * C;
        const int blocks1 = (int)min((long long)148 * 8, (n + 255) / 256);
        if (act == AFLDM_ACT_SILU)
            launch_k(affine_act_scalar_kernel<AFLDM_ACT_SILU>, dim3(blocks1), dim3(256), 0, as_stream(stream),
                x, y, n, C, (long long)HW * C, scale, shift, y_half);
        else
            launch_k(affine_act_scalar_kernel<AFLDM_ACT_IDENTITY>, dim3(blocks1), dim3(256), 0, as_stream(stream),
                x, y, n, C, (long long)HW * C, scale, shift, y_half);
        return launched();
    }
    const long long n4 = (long long)B * HW * C / 4;
    const long long per_image4 = (long long)HW * C / 4;
    const int blocks = (int)min((long long)148 * 8, (n4 + 255) / 256);
    cudaStream_t st = as_stream(stream);
    auto sc = reinterpret_cast<const float4*>(scale);
    auto sh = reinterpret_cast<const float4*>(shift);
    if (act == AFLDM_ACT_SILU)
        launch_k(affine_act_kernel<AFLDM_ACT_SILU>, dim3(blocks), dim3(256), 0, st, 
            reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), n4, C / 4, per_image4, sc, sh, y_half);
    else if (act == AFLDM_ACT_IDENTITY)
        launch_k(affine_act_kernel<AFLDM_ACT_IDENTITY>, dim3(blocks), dim3(256), 0, st, 
            reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), n4, C / 4, per_image4, sc, sh, y_half);
    else
        return AFLDM_E_ARG;
    return launched();
}

extern "C" int afldm_affine_act_f32(const float* x, float* y, int B, int HW, int C, int act,
                                    const float* scale, const float* shift, afldm_stream_t stream) {
    return affine_act_impl(x, y, 0, B, HW, C, act, scale, shift, stream);
}

extern "C" int afldm_affine_act_f16out(const float* x, void* y, int B, int HW, int C, int act,
                                       const float* scale, const float* shift, afldm_stream_t stream) {
    if (static_cast<const void*>(x) == y) return AFLDM_E_ARG;
    return affine_act_impl(x, static_cast<float*>(y), 1, B, HW, C, act, scale, shift, stream);
}

static int affine_act_gn_impl(const float* x, float* y, int y_half, int B, int HW, int C, int act,
                              const float* partial_a, int slots_a, int Ca, const float* partial_b,
                              int slots_b, int Cb, int groups, float eps, const float* gamma,
                              const float* beta, afldm_stream_t stream) {
    if (x == nullptr || y == nullptr || partial_a == nullptr || B <= 0 || HW <= 0 || C <= 0) return AFLDM_E_ARG;
    if (slots_a <= 0 || Ca <= 0 || Cb < 0 || groups <= 0 || (Cb > 0 && (partial_b == nullptr || slots_b <= 0)))
        return AFLDM_E_ARG;
    if (Ca + Cb != C || C % groups != 0 || C % 4 != 0 || B > 65535) return AFLDM_E_SHAPE;
    if (!aligned16(x) || !aligned16(y)) return AFLDM_E_ARG;
    const size_t smem = (size_t)(2 * C + 2 * groups) * sizeof(float);
    if (smem > 48 * 1024) return AFLDM_E_SHAPE;
    const long long per_image4 = (long long)HW * C / 4;
    // about two waves of 1024-thread CTAs over the batch; at least one CTA per image
    int chunks = (int)std::min<long long>((per_image4 + 1023) / 1024, std::max(1, 2 * 148 / B));
    if (chunks < 1) chunks = 1;
    const double inv_n = 1.0 / ((double)HW * (double)(C / groups));
    cudaStream_t st = as_stream(stream);
    auto pa = reinterpret_cast<const float2*>(partial_a);
    auto pb = reinterpret_cast<const float2*>(partial_b);
    auto x4 = reinterpret_cast<const float4*>(x);
    auto y4 = reinterpret_cast<float4*>(y);
    if (act == AFLDM_ACT_SILU)
        launch_k(affine_act_gn_kernel<AFLDM_ACT_SILU>, dim3(chunks, B), dim3(1024), smem, st, x4, y4, HW, C, pa, slots_a, Ca,
                 pb, Cb > 0 ? slots_b : 0, Cb, gamma, beta, groups, eps, inv_n, y_half);
    else if (act == AFLDM_ACT_IDENTITY)
        launch_k(affine_act_gn_kernel<AFLDM_ACT_IDENTITY>, dim3(chunks, B), dim3(1024), smem, st, x4, y4, HW, C, pa, slots_a,
                 Ca, pb, Cb > 0 ? slots_b : 0, Cb, gamma, beta, groups, eps, inv_n, y_half);
    else
        return AFLDM_E_ARG;
    return launched();
}

extern "C" int afldm_affine_act_gn_f32(const float* x, float* y, int B, int HW, int C, int act,
                                       const float* partial_a, int slots_a, int Ca, const float* partial_b,
                                       int slots_b, int Cb, int groups, float eps, const float* gamma,
                                       const float* beta, afldm_stream_t stream) {
    return affine_act_gn_impl(x, y, 0, B, HW, C, act, partial_a, slots_a, Ca, partial_b, slots_b, Cb, groups, eps, gamma, beta,
                              stream);
}

extern "C" int afldm_affine_act_gn_f16out(const float* x, void* y, int B, int HW, int C, int act,
                                          const float* partial_a, int slots_a, int Ca, const float* partial_b,
                                          int slots_b, int Cb, int groups, float eps, const float* gamma,
                                          const float* beta, afldm_stream_t stream) {
    if (static_cast<const void*>(x) == y) return AFLDM_E_ARG;
    return affine_act_gn_impl(x, static_cast<float*>(y), 1, B, HW, C, act, partial_a, slots_a, Ca, partial_b, slots_b, Cb,
                              groups, eps, gamma, beta, stream);
}
