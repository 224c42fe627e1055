// Small ops of one denoising step: few-row linear layers (time embedding), sinusoidal timestep
// embedding, channel concat, NCHW <-> NHWC, the DDIM update, and StyleGAN3's upfirdn2d.
#include <algorithm>

#include <cstdlib>

#include "common.cuh"

namespace afldm {
namespace {

// ------------------------------------------------------------------ linear on a few rows
// y[m][n] = act_out( sum_k act_in(x[m][k]) * w[n][k] + bias[n] ),  M <= 64.
// x is staged (activated) in shared memory; one warp per output column streams its weight row
// with coalesced loads - the op is bound by the N*K*4 B weight stream.
constexpr int LIN_MT = 16;  // rows handled per accumulator pass

template <int ACT_IN, int ACT_OUT>
__global__ void __launch_bounds__(256)
linear_rows_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ y, int M, int K, int N) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float xs[];  // [M][K]
    for (int i = threadIdx.x; i < M * K; i += blockDim.x) xs[i] = apply_act<ACT_IN>(x[i]);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (int n = blockIdx.x * warps_per_block + (threadIdx.x >> 5); n < N; n += gridDim.x * warps_per_block) {
        const float* wr = w + (size_t)n * K;
        for (int m0 = 0; m0 < M; m0 += LIN_MT) {
            float acc[LIN_MT];
#pragma unroll
            for (int i = 0; i < LIN_MT; ++i) acc[i] = 0.f;
            for (int kk = lane; kk < K; kk += 32) {
                const float wv = wr[kk];
#pragma unroll
                for (int i = 0; i < LIN_MT; ++i)
                    if (m0 + i < M) acc[i] = fmaf(xs[(m0 + i) * K + kk], wv, acc[i]);
            }
#pragma unroll
            for (int i = 0; i < LIN_MT; ++i) {
                const float s = warp_sum(acc[i]);
                if (lane == 0 && m0 + i < M) {
                    float r = s + (bias != nullptr ? bias[n] : 0.f);
                    y[(size_t)(m0 + i) * N + n] = apply_act<ACT_OUT>(r);
                }
            }
        }
    }
}

// Vectorised variant (K % 4 == 0, 16-byte aligned rows): one warp owns LIN_NR output columns at a time, so every
// staged x value read from shared memory feeds LIN_NR FMAs, and each lane keeps LIN_NR independent 128-bit weight
// loads in flight per step.  The scalar kernel above re-read x from shared memory once per column and issued one
// dependent 4-byte load per step: 55 us for the 43 MB time_emb_proj stream (profiles/r01 breakdown) vs ~7 us of HBM time.
constexpr int LIN_NR = 4;

template <int ACT_IN, int ACT_OUT>
__global__ void __launch_bounds__(256, 2)
linear_rows_vec_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                       float* __restrict__ y, int M, int K, int N) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float4 xs4[];  // [M][K / 4], activated
    const int K4 = K >> 2;
    for (int i = threadIdx.x; i < M * K4; i += blockDim.x) {
        float4 v = reinterpret_cast<const float4*>(x)[i];
        v.x = apply_act<ACT_IN>(v.x); v.y = apply_act<ACT_IN>(v.y);
        v.z = apply_act<ACT_IN>(v.z); v.w = apply_act<ACT_IN>(v.w);
        xs4[i] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int tasks = (N + LIN_NR - 1) / LIN_NR;
    for (int task = blockIdx.x * warps_per_block + (threadIdx.x >> 5); task < tasks; task += gridDim.x * warps_per_block) {
        const int n0 = task * LIN_NR;
        const float4* wr[LIN_NR];
#pragma unroll
        for (int r = 0; r < LIN_NR; ++r) wr[r] = reinterpret_cast<const float4*>(w + (size_t)min(n0 + r, N - 1) * K);
        for (int m0 = 0; m0 < M; m0 += LIN_MT) {
            float acc[LIN_NR][LIN_MT];
#pragma unroll
            for (int r = 0; r < LIN_NR; ++r)
#pragma unroll
                for (int i = 0; i < LIN_MT; ++i) acc[r][i] = 0.f;
            // software pipeline: the next step's weights are in flight while this step's 256 FMAs issue
            float4 wn[LIN_NR];
#pragma unroll
            for (int r = 0; r < LIN_NR; ++r) wn[r] = lane < K4 ? __ldg(wr[r] + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
            for (int q = lane; q < K4; q += 32) {
                float4 wv[LIN_NR];
#pragma unroll
                for (int r = 0; r < LIN_NR; ++r) wv[r] = wn[r];
                if (q + 32 < K4) {
#pragma unroll
                    for (int r = 0; r < LIN_NR; ++r) wn[r] = __ldg(wr[r] + q + 32);
                }
#pragma unroll
                for (int i = 0; i < LIN_MT; ++i) {
                    const float4 xv = xs4[min(m0 + i, M - 1) * K4 + q];
#pragma unroll
                    for (int r = 0; r < LIN_NR; ++r) {
                        acc[r][i] = fmaf(xv.x, wv[r].x, acc[r][i]);
                        acc[r][i] = fmaf(xv.y, wv[r].y, acc[r][i]);
                        acc[r][i] = fmaf(xv.z, wv[r].z, acc[r][i]);
                        acc[r][i] = fmaf(xv.w, wv[r].w, acc[r][i]);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < LIN_NR; ++r) {
#pragma unroll
                for (int i = 0; i < LIN_MT; ++i) {
                    const float s = warp_sum(acc[r][i]);
                    if (lane == 0 && m0 + i < M && n0 + r < N) {
                        const float v = s + (bias != nullptr ? bias[n0 + r] : 0.f);
                        y[(size_t)(m0 + i) * N + n0 + r] = apply_act<ACT_OUT>(v);
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------ timestep embedding
__global__ void timestep_embedding_kernel(const float* __restrict__ t, float* __restrict__ out, int B, int dim) {
    pdl_trigger();
    pdl_wait();
    const int half = dim / 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * half) return;
    const int b = i / half, kf = i % half;
    // same fp32 operation order as the PyTorch expression exp(-ln(1e4) * k / half)
    const float e = (-9.210340371976184f * (float)kf) / (float)half;
    const float arg = t[b] * expf(e);
    out[(size_t)b * dim + kf] = cosf(arg);
    out[(size_t)b * dim + half + kf] = sinf(arg);
}

// ------------------------------------------------------------------ concat along channels
__global__ void __launch_bounds__(256)
concat_kernel(const float4* __restrict__ a, int Ca4, const float4* __restrict__ b, int Cb4,
              float4* __restrict__ y, long long total4) {
    pdl_trigger();
    pdl_wait();
    const int C4 = Ca4 + Cb4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
        const long long p = i / C4;
        const int c = (int)(i - p * C4);
        y[i] = c < Ca4 ? a[p * Ca4 + c] : b[p * Cb4 + (c - Ca4)];
    }
}

// ------------------------------------------------------------------ channel padding
// y[p][0 .. C) = x[p][0 .. C), y[p][C .. Cp) = 0: brings a narrow NHWC tensor (the 4-channel latents in front of
// conv_in) to the 32-channel granularity of the tensor-core convolution's TMA boxes.
__global__ void __launch_bounds__(256)
pad_channels_kernel(const float* __restrict__ x, int C, float* __restrict__ y, int Cp, long long total) {
    pdl_trigger();
    pdl_wait();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long p = i / Cp;
        const int c = (int)(i - p * Cp);
        y[i] = c < C ? x[p * C + c] : 0.f;
    }
}

// ------------------------------------------------------------------ layout transposes
// in [B][R][S] -> out [B][S][R]  (NCHW->NHWC: R = C, S = HW; NHWC->NCHW: R = HW, S = C)
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int S) {
    pdl_trigger();
    pdl_wait();
    __shared__ float t[32][33];
    const int b = blockIdx.z;
    const int s0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const float* ip = in + (size_t)b * R * S;
    float* op = out + (size_t)b * R * S;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int r = r0 + ty + j, s = s0 + tx;
        if (r < R && s < S) t[ty + j][tx] = ip[(size_t)r * S + s];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int s = s0 + ty + j, r = r0 + tx;
        if (r < R && s < S) op[(size_t)s * R + r] = t[tx][ty + j];
    }
}

// ------------------------------------------------------------------ DDIM update
__global__ void __launch_bounds__(256)
axpby_kernel(const float* __restrict__ x, const float* __restrict__ e, float* __restrict__ out, float cx,
             float ce, long long n) {
    pdl_trigger();
    pdl_wait();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = fmaf(cx, x[i], ce * e[i]);
}

__global__ void __launch_bounds__(256)
axpby_dev_kernel(const float* __restrict__ x, const float* __restrict__ e, float* __restrict__ out,
                 const float* __restrict__ coef, long long n) {
    pdl_trigger();
    pdl_wait();
    const float cx = coef[0], ce = coef[1];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = fmaf(cx, x[i], ce * e[i]);
}

// ------------------------------------------------------------------ GEGLU (SD-1.5 feed-forward)
// diffusers GEGLU (BasicTransformerBlock.ff.net.0 of UNet2DConditionModel, reached through
// afldm/pipelines/video_equiv_editing_pipeline.py:680-686): y[m][h] = p[m][h] * gelu(p[m][H + h]), exact (erf) GELU.
__global__ void __launch_bounds__(256)
geglu_kernel(const float* __restrict__ p, float* __restrict__ y, long long rows, int H) {
    pdl_trigger();
    pdl_wait();
    const long long total4 = rows * (H / 4);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
        const long long m = i / (H / 4);
        const int h4 = (int)(i - m * (H / 4));
        const float4 a = *reinterpret_cast<const float4*>(p + m * 2 * H + 4 * h4);
        const float4 g = *reinterpret_cast<const float4*>(p + m * 2 * H + H + 4 * h4);
        float4 o;
        o.x = a.x * (0.5f * g.x * (1.0f + erff(g.x * 0.70710678118654752f)));
        o.y = a.y * (0.5f * g.y * (1.0f + erff(g.y * 0.70710678118654752f)));
        o.z = a.z * (0.5f * g.z * (1.0f + erff(g.z * 0.70710678118654752f)));
        o.w = a.w * (0.5f * g.w * (1.0f + erff(g.w * 0.70710678118654752f)));
        *reinterpret_cast<float4*>(y + m * H + 4 * h4) = o;
    }
}

// ------------------------------------------------------------------ activation derivative (backward of the filtered activation)
// out = g * act'(z): the elementwise factor between the two linear halves of d/dx [D act(U x U^T) D^T]
// (training-side use: afldm/trainers/ldm_trainer.py:240-272).  silu'(z) = s (1 + z (1 - s)), s = sigmoid(z).
__global__ void __launch_bounds__(256)
act_bwd_mul_kernel(const float* __restrict__ z, const float* __restrict__ g, float* __restrict__ out, long long n, int act) {
    pdl_trigger();
    pdl_wait();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float d = 1.f;
        if (act == AFLDM_ACT_SILU) {
            const float zz = z[i];
            const float sg = 1.0f / (1.0f + __expf(-zz));
            d = sg * fmaf(zz, 1.0f - sg, 1.0f);
        }
        out[i] = g[i] * d;
    }
}

// ------------------------------------------------------------------ slot copy (cross-frame attention maps)
// table[slot][n] <-> buf[n] with the slot index read from DEVICE memory: a captured denoising step can keep one map per
// timestep (CrossFrameAttnProcessor, afldm/pipelines/cross_frame_attn.py:78-97, keys its dictionaries by the host value
// t.item(); here the step index is a device scalar refreshed between graph replays).
__global__ void __launch_bounds__(256)
slot_copy_kernel(float* __restrict__ table, float* __restrict__ buf, long long n4, const int* __restrict__ slot, int store) {
    pdl_trigger();
    pdl_wait();
    float4* t4 = reinterpret_cast<float4*>(table) + (long long)slot[0] * n4;
    float4* b4 = reinterpret_cast<float4*>(buf);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        if (store) t4[i] = b4[i];
        else b4[i] = t4[i];
    }
}

// ------------------------------------------------------------------ upfirdn2d (NCHW)
// out[oy][ox] = gain * sum_{fy,fx} f'[fy][fx] * xup[oy*downy + fy - pady0][ox*downx + fx - padx0]
// where xup is x zero-stuffed by (upy, upx) and f' = f flipped unless `flip`
// (afldm/af_libs/torch_utils/ops/upfirdn2d.py:166-211; true convolution by default).
__global__ void __launch_bounds__(256)
upfirdn2d_kernel(const float* __restrict__ x, const float* __restrict__ f, float* __restrict__ y,
                 int H, int W, int fh, int fw, int upx, int upy, int downx, int downy, int padx0, int pady0,
                 int outH, int outW, int flip, float gain, long long total) {
    pdl_trigger();
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int ox = (int)(i % outW);
    const int oy = (int)((i / outW) % outH);
    const long long plane = i / ((long long)outW * outH);
    const float* xp = x + plane * (long long)H * W;
    float acc = 0.f;
    for (int fy = 0; fy < fh; ++fy) {
        const int uy = oy * downy + fy - pady0;  // coordinate on the zero-stuffed grid
        if (uy < 0 || uy % upy != 0) continue;
        const int iy = uy / upy;
        if (iy >= H) continue;
        for (int fx = 0; fx < fw; ++fx) {
            const int ux = ox * downx + fx - padx0;
            if (ux < 0 || ux % upx != 0) continue;
            const int ix = ux / upx;
            if (ix >= W) continue;
            const float tap = flip ? f[fy * fw + fx] : f[(fh - 1 - fy) * fw + (fw - 1 - fx)];
            acc = fmaf(tap, xp[(size_t)iy * W + ix], acc);
        }
    }
    y[i] = acc * gain;
}

inline int grid_for(long long n, int threads = 256, int cap = 148 * 8) {
    return (int)std::max<long long>(1, std::min<long long>(cap, (n + threads - 1) / threads));
}


// K = 768 (the time-embedding width of the FFHQ UNet: Linear(768, 768) and the 27 fused time_emb_proj rows,
// 768 -> 14016), M <= 16.  The projection is a 43 MB weight stream; the generic kernel above keeps one 64-byte load
// per row in flight per lane and reached 1.1 TB/s.  Here a warp owns TWO output columns and every lane issues its whole
// share of both weight rows up front (12 independent 128-bit loads = 6 KB in flight per warp), then runs the 16 x 2
// dot products out of registers against the staged activations: the stream is bandwidth-, not latency-bound.
constexpr int L768_K4 = 192, L768_IT = 6, L768_NR = 2, L768_M = 16;

template <int ACT_IN, int ACT_OUT>
__global__ void __launch_bounds__(256)
linear_rows_k768_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                        float* __restrict__ y, int M, int N) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float4 xs4[];  // [M][192], activated
    for (int i = threadIdx.x; i < M * L768_K4; i += blockDim.x) {
        float4 v = reinterpret_cast<const float4*>(x)[i];
        v.x = apply_act<ACT_IN>(v.x); v.y = apply_act<ACT_IN>(v.y);
        v.z = apply_act<ACT_IN>(v.z); v.w = apply_act<ACT_IN>(v.w);
        xs4[i] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int tasks = (N + L768_NR - 1) / L768_NR;
    for (int task = blockIdx.x * warps_per_block + (threadIdx.x >> 5); task < tasks; task += gridDim.x * warps_per_block) {
        const int n0 = task * L768_NR;
        float4 wv[L768_NR][L768_IT];
#pragma unroll
        for (int r = 0; r < L768_NR; ++r) {
            const float4* wr = reinterpret_cast<const float4*>(w + (size_t)min(n0 + r, N - 1) * (4 * L768_K4));
#pragma unroll
            for (int j = 0; j < L768_IT; ++j) wv[r][j] = __ldg(wr + lane + 32 * j);
        }
        float acc[L768_NR][L768_M];
#pragma unroll
        for (int i = 0; i < L768_M; ++i) {
            const float4* xr = xs4 + min(i, M - 1) * L768_K4 + lane;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int j = 0; j < L768_IT; ++j) {
                const float4 xv = xr[32 * j];
                a0 = fmaf(xv.x, wv[0][j].x, a0); a0 = fmaf(xv.y, wv[0][j].y, a0);
                a0 = fmaf(xv.z, wv[0][j].z, a0); a0 = fmaf(xv.w, wv[0][j].w, a0);
                a1 = fmaf(xv.x, wv[1][j].x, a1); a1 = fmaf(xv.y, wv[1][j].y, a1);
                a1 = fmaf(xv.z, wv[1][j].z, a1); a1 = fmaf(xv.w, wv[1][j].w, a1);
            }
            acc[0][i] = a0;
            acc[1][i] = a1;
        }
#pragma unroll
        for (int r = 0; r < L768_NR; ++r) {
#pragma unroll
            for (int i = 0; i < L768_M; ++i) {
                const float sum = warp_sum(acc[r][i]);
                if (lane == 0 && i < M && n0 + r < N) {
                    const float v = sum + (bias != nullptr ? bias[n0 + r] : 0.f);
                    y[(size_t)i * N + n0 + r] = apply_act<ACT_OUT>(v);
                }
            }
        }
    }
}
}  // namespace
}  // namespace afldm

using namespace afldm;

extern "C" int afldm_linear_rows_f32(const float* x, const float* w, const float* bias, float* y, int M, int K,
                                     int N, int act_in, int act_out, afldm_stream_t stream) {
    if (x == nullptr || w == nullptr || y == nullptr || M <= 0 || K <= 0 || N <= 0) return AFLDM_E_ARG;
    if (M > 64) return AFLDM_E_SHAPE;
    const size_t smem = (size_t)M * K * sizeof(float);
    if (smem > 200 * 1024) return AFLDM_E_SHAPE;
    cudaStream_t st = as_stream(stream);
    const bool vec = (K % 4 == 0) && aligned16(x) && aligned16(w);
    const int blocks = vec ? std::min(ceil_div(ceil_div(N, LIN_NR), 8), 148 * 2) : std::min(ceil_div(N, 8), 148 * 4);
    static const bool k768 = !(getenv("AFLDM_LIN_K768") && atoi(getenv("AFLDM_LIN_K768")) == 0);
    const bool fast768 = vec && k768 && K == 4 * L768_K4 && M <= L768_M;
    const int blocks768 = std::min(ceil_div(ceil_div(N, L768_NR), 8), 148 * 2);
#define AFLDM_LIN(AI, AO)                                                                              \
    {                                                                                                  \
        if (fast768) {                                                                                 \
            static std::atomic<unsigned long long> cfg768{0};                                          \
            {                                                                                          \
                cudaError_t e = set_max_dyn_smem(linear_rows_k768_kernel<AI, AO>, 64 * 1024, cfg768);  \
                if (e != cudaSuccess) return (int)e;                                                   \
            }                                                                                          \
            launch_k(linear_rows_k768_kernel<AI, AO>, dim3(blocks768), dim3(256), smem, st, x, w, bias, y, M, N); \
            return launched();                                                                         \
        }                                                                                              \
        auto kern = vec ? linear_rows_vec_kernel<AI, AO> : linear_rows_kernel<AI, AO>;                 \
        if (smem > 48 * 1024) {                                                                        \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
            if (e != cudaSuccess) return (int)e;                                                       \
        }                                                                                              \
        launch_k(kern, dim3(blocks), dim3(256), smem, st, x, w, bias, y, M, K, N);                                       \
        return launched();                                                                             \
    }
    const bool ai = act_in == AFLDM_ACT_SILU, ao = act_out == AFLDM_ACT_SILU;
    if ((act_in != AFLDM_ACT_SILU && act_in != AFLDM_ACT_IDENTITY) ||
        (act_out != AFLDM_ACT_SILU && act_out != AFLDM_ACT_IDENTITY))
        return AFLDM_E_ARG;
    if (ai && ao) AFLDM_LIN(AFLDM_ACT_SILU, AFLDM_ACT_SILU)
    if (ai && !ao) AFLDM_LIN(AFLDM_ACT_SILU, AFLDM_ACT_IDENTITY)
    if (!ai && ao) AFLDM_LIN(AFLDM_ACT_IDENTITY, AFLDM_ACT_SILU)
    AFLDM_LIN(AFLDM_ACT_IDENTITY, AFLDM_ACT_IDENTITY)
#undef AFLDM_LIN
}

extern "C" int afldm_timestep_embedding_f32(const float* t, float* out, int B, int dim, afldm_stream_t stream) {
    if (t == nullptr || out == nullptr || B <= 0 || dim <= 0) return AFLDM_E_ARG;
    if (dim % 2 != 0) return AFLDM_E_SHAPE;
    const int n = B * (dim / 2);
    launch_k(timestep_embedding_kernel, dim3(ceil_div(n, 128)), dim3(128), 0, as_stream(stream), t, out, B, dim);
    return launched();
}

extern "C" int afldm_concat_channels_f32(const float* a, int Ca, const float* b, int Cb, float* y,
                                         long long pixels, afldm_stream_t stream) {
    if (a == nullptr || b == nullptr || y == nullptr || Ca <= 0 || Cb <= 0 || pixels <= 0) return AFLDM_E_ARG;
    if (Ca % 4 != 0 || Cb % 4 != 0) return AFLDM_E_SHAPE;
    if (!aligned16(a) || !aligned16(b) || !aligned16(y)) return AFLDM_E_ARG;
    const long long total4 = pixels * (Ca + Cb) / 4;
    launch_k(concat_kernel, dim3(grid_for(total4)), dim3(256), 0, as_stream(stream), 
        reinterpret_cast<const float4*>(a), Ca / 4, reinterpret_cast<const float4*>(b), Cb / 4,
        reinterpret_cast<float4*>(y), total4);
    return launched();
}

static int transpose_launch(const float* x, float* y, int B, int R, int S, afldm_stream_t stream) {
    if (x == nullptr || y == nullptr || B <= 0 || R <= 0 || S <= 0 || x == y) return AFLDM_E_ARG;
    if (B > 65535 || ceil_div(R, 32) > 65535) return AFLDM_E_SHAPE;
    launch_k(transpose_kernel, dim3(ceil_div(S, 32), ceil_div(R, 32), B), dim3(256), 0, as_stream(stream), x, y, R, S);
    return launched();
}

extern "C" int afldm_nchw_to_nhwc_f32(const float* x, float* y, int B, int C, int HW, afldm_stream_t stream) {
    return transpose_launch(x, y, B, C, HW, stream);
}

extern "C" int afldm_nhwc_to_nchw_f32(const float* x, float* y, int B, int C, int HW, afldm_stream_t stream) {
    return transpose_launch(x, y, B, HW, C, stream);
}

extern "C" int afldm_axpby_f32(const float* x, const float* eps, float* out, float cx, float ce, long long n,
                               afldm_stream_t stream) {
    if (x == nullptr || eps == nullptr || out == nullptr || n <= 0) return AFLDM_E_ARG;
    launch_k(axpby_kernel, dim3(grid_for(n)), dim3(256), 0, as_stream(stream), x, eps, out, cx, ce, n);
    return launched();
}

extern "C" int afldm_axpby_dev_f32(const float* x, const float* eps, float* out, const float* coef,
                                   long long n, afldm_stream_t stream) {
    if (x == nullptr || eps == nullptr || out == nullptr || coef == nullptr || n <= 0) return AFLDM_E_ARG;
    launch_k(axpby_dev_kernel, dim3(grid_for(n)), dim3(256), 0, as_stream(stream), x, eps, out, coef, n);
    return launched();
}

extern "C" int afldm_act_bwd_mul_f32(const float* z, const float* g, float* out, long long n, int act, afldm_stream_t stream) {
    if (z == nullptr || g == nullptr || out == nullptr || n <= 0) return AFLDM_E_ARG;
    if (act != AFLDM_ACT_SILU && act != AFLDM_ACT_IDENTITY) return AFLDM_E_ARG;
    launch_k(act_bwd_mul_kernel, dim3(grid_for(n)), dim3(256), 0, as_stream(stream), z, g, out, n, act);
    return launched();
}

extern "C" int afldm_geglu_f32(const float* proj, float* y, long long rows, int H, afldm_stream_t stream) {
    if (proj == nullptr || y == nullptr || rows <= 0 || H <= 0 || (H & 3) != 0) return AFLDM_E_ARG;
    if (!aligned16(proj) || !aligned16(y)) return AFLDM_E_ARG;
    launch_k(geglu_kernel, dim3(grid_for(rows * (H / 4))), dim3(256), 0, as_stream(stream), proj, y, rows, H);
    return launched();
}

extern "C" int afldm_slot_copy_f32(float* table, float* buf, long long n, const int* slot, int store, afldm_stream_t stream) {
    if (table == nullptr || buf == nullptr || slot == nullptr || n <= 0 || (n & 3) != 0) return AFLDM_E_ARG;
    if (!aligned16(table) || !aligned16(buf)) return AFLDM_E_ARG;
    launch_k(slot_copy_kernel, dim3(grid_for(n / 4)), dim3(256), 0, as_stream(stream), table, buf, n / 4, slot, store);
    return launched();
}

extern "C" int afldm_upfirdn2d_f32(const float* x, const float* f, float* y, int B, int C, int H, int W, int fh,
                                   int fw, int upx, int upy, int downx, int downy, int padx0, int padx1,
                                   int pady0, int pady1, int flip, float gain, afldm_stream_t stream) {
    if (x == nullptr || f == nullptr || y == nullptr) return AFLDM_E_ARG;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || fh <= 0 || fw <= 0) return AFLDM_E_ARG;
    if (upx <= 0 || upy <= 0 || downx <= 0 || downy <= 0) return AFLDM_E_ARG;
    const int outW = (W * upx + padx0 + padx1 - fw + downx) / downx;
    const int outH = (H * upy + pady0 + pady1 - fh + downy) / downy;
    if (outW < 1 || outH < 1) return AFLDM_E_SHAPE;
    const long long total = (long long)B * C * outH * outW;
    const int blocks = (int)((total + 255) / 256);
    launch_k(upfirdn2d_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), x, f, y, H, W, fh, fw, upx, upy, downx, downy, padx0,
                                                           pady0, outH, outW, flip, gain, total);
    return launched();
}

// ------------------------------------------------------------------ separable plane operator (NCHW)
// y[p] = My . x[p] . Mx^T for every plane p: the fused form of ImageShifter('ideal' | 'ideal_crop', r).shift
// (afldm/shift_utils/shifters.py:157-191): ideal r-fold up-sampling, circular roll by round(t r), validity mask and
// [::r, ::r] decimation are all linear and separable, so per axis they collapse into ONE n x n matrix
// S[i][j] = d_r[((i - j) r - s) mod n r] (rows of masked positions zeroed) that the host builds in fp64
// (afldm_b200/shift_utils/shifters.py).  The r-fold up-sampled tensor (64x the data at r = 8) never exists, and a
// sweep of shifts is one launch pair: plane p uses matrix pair p / planes_per_matrix.
// Two passes of a 32 x 32-tiled fp32 GEMM with sequential k order (deterministic): C[r][c] = sum_k A[r][k] B(k, c),
// B stored [c][k] (BT: the row pass, Mx) or [k][c] (the column pass reads the row pass's result).
namespace afldm {
namespace {
template <bool BT>
__global__ void __launch_bounds__(256)
sep_gemm_kernel(const float* __restrict__ A, long long a_plane, long long a_matrix, const float* __restrict__ Bm,
                long long b_plane, long long b_matrix, float* __restrict__ Cm, int R, int Cc, int K,
                int planes_per_matrix) {
    pdl_trigger();
    pdl_wait();
    __shared__ float As[32][33], Bs[32][33];
    const int p = blockIdx.z, m = p / planes_per_matrix;
    const float* a = A + (size_t)p * a_plane + (size_t)m * a_matrix;
    const float* b = Bm + (size_t)p * b_plane + (size_t)m * b_matrix;
    float* c = Cm + (size_t)p * R * Cc;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int rr = ty + 8 * u;
            const int ar = r0 + rr, ak = k0 + tx;
            As[rr][tx] = (ar < R && ak < K) ? a[(size_t)ar * K + ak] : 0.f;
            if (BT) {       // B(k, c) = b[c][k]: tile rows = c, columns = k
                const int bc = c0 + rr, bk = k0 + tx;
                Bs[rr][tx] = (bc < Cc && bk < K) ? b[(size_t)bc * K + bk] : 0.f;
            } else {        // B(k, c) = b[k][c]: tile rows = k, columns = c
                const int bk = k0 + rr, bc = c0 + tx;
                Bs[rr][tx] = (bk < K && bc < Cc) ? b[(size_t)bk * Cc + bc] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < 32; ++kk) {
            const float bv = BT ? Bs[tx][kk] : Bs[kk][tx];
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] = fmaf(As[ty + 8 * u][kk], bv, acc[u]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int rr = r0 + ty + 8 * u, cc = c0 + tx;
        if (rr < R && cc < Cc) c[(size_t)rr * Cc + cc] = acc[u];
    }
}
}  // namespace
}  // namespace afldm

extern "C" size_t afldm_plane_sep_transform_workspace_floats(int planes, int Hin, int Wout) {
    if (planes <= 0 || Hin <= 0 || Wout <= 0) return 0;
    return (size_t)planes * Hin * Wout;
}

extern "C" int afldm_plane_sep_transform_f32(const float* x, const float* my, const float* mx, float* y,
                                             float* workspace, size_t workspace_floats, int planes,
                                             int planes_per_matrix, int Hin, int Win, int Hout, int Wout,
                                             afldm_stream_t stream) {
    if (x == nullptr || my == nullptr || mx == nullptr || y == nullptr || x == y) return AFLDM_E_ARG;
    if (planes <= 0 || planes_per_matrix <= 0 || Hin <= 0 || Win <= 0 || Hout <= 0 || Wout <= 0) return AFLDM_E_ARG;
    if (planes > 65535) return AFLDM_E_SHAPE;
    if (workspace == nullptr || workspace_floats < (size_t)planes * Hin * Wout) return AFLDM_E_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    // rows: T[p] (Hin x Wout) = x[p] (Hin x Win) . Mx[m]^T, Mx stored [Wout][Win]
    launch_k(sep_gemm_kernel<true>, dim3(ceil_div(Wout, 32), ceil_div(Hin, 32), planes), dim3(256), 0, st,
             x, (long long)Hin * Win, 0LL, mx, 0LL, (long long)Wout * Win, workspace, Hin, Wout, Win, planes_per_matrix);
    // columns: y[p] (Hout x Wout) = My[m] (Hout x Hin) . T[p]
    launch_k(sep_gemm_kernel<false>, dim3(ceil_div(Wout, 32), ceil_div(Hout, 32), planes), dim3(256), 0, st,
             my, 0LL, (long long)Hout * Hin, (const float*)workspace, (long long)Hin * Wout, 0LL, y, Hout, Wout, Hin,
             planes_per_matrix);
    return launched(2);
}

extern "C" int afldm_pad_channels_f32(const float* x, int C, float* y, int Cpad, long long pixels, afldm_stream_t stream) {
    if (x == nullptr || y == nullptr || C <= 0 || Cpad < C || pixels <= 0 || x == y) return AFLDM_E_ARG;
    const long long total = pixels * Cpad;
    launch_k(pad_channels_kernel, dim3(grid_for(total)), dim3(256), 0, as_stream(stream), x, C, y, Cpad, total);
    return launched();
}
