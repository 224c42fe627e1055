// Pieces shared by the filtered-activation kernels (resample.cu: FMA / mma.sync forms, fact_tc.cu: tcgen05 form):
// the per-(b, c) affine that folds GroupNorm into the load, the un-materialised concat source, the in-kernel
// GroupNorm finalisation and the fp16 hi / lo split.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace afldm {

// Per-(b, c) affine applied on load (GroupNorm folded into the resampler).  Either precomputed scale/shift
// vectors, or the GroupNorm partial sums that the producer of x emitted (conv epilogue): then every CTA
// finalises mean / rstd for its own CG channels in its prologue and no separate kernel runs at all.
struct Affine {
    const float* scale;
    const float* shift;
    const float2* pa;     // [B][slots_a][Ca] (sum, sumsq); channels [0, Ca)
    const float2* pb;     // [B][slots_b][Cb]; channels [Ca, Ca + Cb) (second half of a skip concat) | NULL
    const float* gamma;
    const float* beta;
    int slots_a, Ca, slots_b, Cb, groups, HW;
    float eps;
    double inv_n;         // 1 / (HW * channels per group), from the host: no fp64 division in the prologue
    float2* gn_out;       // MODE_DOWN2 only: [B][1][C] (sum, sum of squares) of every output plane | NULL
    int y_half;           // MODE_FACT / MODE_UP2: y holds IEEE binary16 (same indexing, in elements) - the operand of the next conv
    const float* x2;      // second input source: channels [xCa, C) are read from x2 (pixel pitch C - xCa), channels
    int xCa;              // [0, xCa) from x (pixel pitch xCa) - a skip-connection concat that is never materialised
};

// One warp per GroupNorm group touched by this CTA's CG channels (at most CG of them): the partial sums are added
// in fp32 (they are fp32 sums over <= 128 pixels already), only the final E[x^2] - mean^2 is formed in fp64.
// (The first version ran a full fp64 reduction per CHANNEL in every CTA and lost 5 us per call to the separate
// finalize kernel; this one adds ~1 us of latency to the first wave of CTAs and removes a 3 us launch.)
// Input source of the channel group that starts at c0: (base pointer incl. channel offset, pixel pitch).
struct XSrc {
    const float* p;
    int pitch;
};
__device__ __forceinline__ XSrc x_source(const float* x, int C, const Affine& af, int c0) {
    if (af.x2 == nullptr) return {x + c0, C};
    if (c0 < af.xCa) return {x + c0, af.xCa};
    return {af.x2 + (c0 - af.xCa), C - af.xCa};
}

template <int CG>
__device__ __forceinline__ void gn_prologue(const Affine& af, int b, int c0, int C, float* s_sc, float* s_sh) {
    __shared__ float s_mean[CG], s_rstd[CG];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int cpg = C / af.groups;
    const int g_first = c0 / cpg, g_last = (c0 + CG - 1) / cpg;
    const int slots_max = max(af.slots_a, af.slots_b);
    for (int gi = warp; gi <= g_last - g_first; gi += nwarps) {
        const int g = g_first + gi;
        float s = 0.f, q = 0.f;
        const int total = slots_max * cpg;
        for (int it0 = lane; it0 < total; it0 += 128) {      // four independent loads in flight per lane
            float2 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int it = it0 + 32 * u;
                const int sl = it / cpg, cc = g * cpg + (it - sl * cpg);
                v[u] = make_float2(0.f, 0.f);
                if (it < total) {
                    if (cc < af.Ca) {
                        if (sl < af.slots_a) v[u] = __ldg(&af.pa[((size_t)b * af.slots_a + sl) * af.Ca + cc]);
                    } else {
                        if (sl < af.slots_b) v[u] = __ldg(&af.pb[((size_t)b * af.slots_b + sl) * af.Cb + (cc - af.Ca)]);
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
            const double mean = (double)s * af.inv_n;
            double var = (double)q * af.inv_n - mean * mean;
            if (var < 0.0) var = 0.0;
            s_mean[gi] = (float)mean;
            s_rstd[gi] = rsqrtf((float)var + af.eps);
        }
    }
    __syncthreads();
    if (threadIdx.x < CG) {
        const int c = c0 + threadIdx.x;
        const int gi = c / cpg - g_first;
        const float sc = (af.gamma != nullptr ? af.gamma[c] : 1.f) * s_rstd[gi];
        s_sc[threadIdx.x] = sc;
        s_sh[threadIdx.x] = fmaf(-s_mean[gi], sc, af.beta != nullptr ? af.beta[c] : 0.f);
    }
    __syncthreads();
}

__device__ __forceinline__ void split_pack(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(v0, v1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}


}  // namespace afldm
