// Ideal (circular-sinc) resampling kernels: filtered activation, x2 up-sample, LPF + decimate.
//
// The reference runs these through cuFFT (afldm/af_libs/ideal_lpf.py:69-93, 112-134, 148-158
// and afldm/af_modules/af_blocks.py:19-28).  Per (b, c) plane of side n the same operator is
//      up2(x)        = U x U^T            U in R^{2n x n}: even rows identity, odd rows circulant d
//      lpf_down2(a)  = D a D^T            D in R^{n x 2n}: D[i, m] = g[(2i - m) mod 2n]
//      filtered_act  = D act(U x U^T) D^T
// (SURVEY.md 8(a) identities 1-4).  Each 1-D circular convolution of one line is done by ONE
// thread entirely in registers: the loops are fully unrolled, so every tap index is a
// compile-time constant and the tap becomes a constant-bank operand of the FFMA - no shared
// memory or register traffic for the filter at all.  Lines are exchanged between the row and
// column passes through one shared-memory tile [n][2n][CG] (CG channels of the NHWC tensor,
// channel fastest, row pitch padded by CG words so that every pass is bank-conflict free).
//
//   pass 1  thread (c, i):  row i of x  ->  row i of T = x U^T                 (n^2 FMA)
//   pass 2  thread (c, jj): column jj of T -> column of Z = U T, act()          (n^2 FMA)
//   pass 3  same thread:    column of A -> column jj of Y1 = D A  (in place)   (2 n^2 FMA)
//   pass 4  thread (c, i):  row i of Y1 -> row i of y = Y1 D^T                 (2 n^2 FMA)
//
// Algorithmic HBM traffic: filtered_act 8 B/element, up2 20 B per input element,
// lpf_down2 5 B per input element (fp32).
#include "common.cuh"
#include "resample.cuh"
#include "taps.inc"

namespace afldm {
namespace {

enum { MODE_FACT = 0, MODE_UP2 = 1, MODE_DOWN2 = 2 };

// o[i] = sum_j d[(i - j) mod N] x[j]
template <int N>
__device__ __forceinline__ void up_odd(const float (&x)[N], float (&o)[N]) {
    constexpr int IB = N >= 8 ? 8 : N;  // independent accumulators for ILP
#pragma unroll
    for (int i0 = 0; i0 < N; i0 += IB) {
        float acc[IB];
#pragma unroll
        for (int u = 0; u < IB; ++u) acc[u] = 0.f;
#pragma unroll
        for (int j = 0; j < N; ++j) {
#pragma unroll
            for (int u = 0; u < IB; ++u) acc[u] = fmaf(tap_d<N>((i0 + u - j) & (N - 1)), x[j], acc[u]);
        }
#pragma unroll
        for (int u = 0; u < IB; ++u) o[i0 + u] = acc[u];
    }
}

// y[i] = sum_m g[(2i - m) mod 2N] a[m].
// The even-indexed taps of g need no convolution: the pass band of LPF_RFFT(.5) on 2N points is the N - 1
// bins |k| < N/2 (create_lpf_rect zeroes bin N/2, ideal_lpf.py:12-24), so
//      g[2r] = (1/2N) sum_{|k| < N/2} e^{2 pi i k r / N} = (1/2) delta[r mod N] - (-1)^r / (2N),
// i.e. the even samples contribute  a[2i] / 2 - (-1)^i S / (2N)  with  S = sum_m (-1)^m a[2m]  (O(N) per
// line); only the odd samples go through an N-tap circular convolution.  2N^2 -> N^2 + 2N FMAs per line.
template <int N>
__device__ __forceinline__ float alt_sum_even(const float (&a)[2 * N]) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int m = 0; m < N; m += 2) {
        s0 += a[2 * m];
        s1 += a[2 * m + 2];
    }
    return (s0 - s1) * (1.0f / (2 * N));
}

template <int N>
__device__ __forceinline__ void down_line(const float (&a)[2 * N], float (&y)[N]) {
    constexpr int IB = N >= 8 ? 8 : N;
    const float sc = alt_sum_even<N>(a);
#pragma unroll
    for (int i0 = 0; i0 < N; i0 += IB) {
        float acc[IB];
#pragma unroll
        for (int u = 0; u < IB; ++u) acc[u] = fmaf(0.5f, a[2 * (i0 + u)], ((i0 + u) & 1) ? sc : -sc);
#pragma unroll
        for (int m = 1; m < 2 * N; m += 2) {
#pragma unroll
            for (int u = 0; u < IB; ++u)
                acc[u] = fmaf(tap_g<N>((2 * (i0 + u) - m) & (2 * N - 1)), a[m], acc[u]);
        }
#pragma unroll
        for (int u = 0; u < IB; ++u) y[i0 + u] = acc[u];
    }
}

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
};

template <int CG>
__device__ __forceinline__ void gn_prologue(const Affine& af, int b, int c0, int C, float* s_sc, float* s_sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int cpg = C / af.groups;
    const int slots_max = max(af.slots_a, af.slots_b);
    for (int ch = warp; ch < CG; ch += nwarps) {
        const int c = c0 + ch;
        const int g = c / cpg;
        double s = 0.0, q = 0.0;
        for (int it = lane; it < slots_max * cpg; it += 32) {
            const int sl = it / cpg, cc = g * cpg + (it - sl * cpg);
            float2 v = make_float2(0.f, 0.f);
            if (cc < af.Ca) {
                if (sl < af.slots_a) v = af.pa[((size_t)b * af.slots_a + sl) * af.Ca + cc];
            } else {
                if (sl < af.slots_b) v = af.pb[((size_t)b * af.slots_b + sl) * af.Cb + (cc - af.Ca)];
            }
            s += (double)v.x;
            q += (double)v.y;
        }
        s = warp_sum(s);
        q = warp_sum(q);
        if (lane == 0) {
            const double n = (double)af.HW * (double)cpg;
            const double mean = s / n;
            double var = q / n - mean * mean;
            if (var < 0.0) var = 0.0;
            const float rstd = (float)(1.0 / sqrt(var + (double)af.eps));
            const float sc = (af.gamma != nullptr ? af.gamma[c] : 1.f) * rstd;
            s_sc[ch] = sc;
            s_sh[ch] = fmaf(-(float)mean, sc, af.beta != nullptr ? af.beta[c] : 0.f);
        }
    }
    __syncthreads();
}

template <int N, int CG>
struct Tile {
    static constexpr int PITCH = (2 * N + 1) * CG;  // floats per tile row (padded)
    static constexpr int SMEM_BYTES = N * PITCH * 4;
};

// N: side of the SMALL plane (input of FACT / UP2, output of DOWN2).
template <int N, int CG, int MODE, int ACT>
__global__ void __launch_bounds__(256, 2)
resample_kernel(const float* __restrict__ x, float* __restrict__ y, int C, const Affine af) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float tile[];
    __shared__ float s_sc[CG], s_sh[CG];
    if (MODE != MODE_DOWN2 && af.pa != nullptr) gn_prologue<CG>(af, blockIdx.y, blockIdx.x * CG, C, s_sc, s_sh);
    constexpr int PITCH = Tile<N, CG>::PITCH;
    constexpr int M = 2 * N;
    const int b = blockIdx.y;
    const int c0 = blockIdx.x * CG;

    if constexpr (MODE != MODE_DOWN2) {
        // pass 1: rows
        for (int t = threadIdx.x; t < N * CG; t += blockDim.x) {
            const int c = t % CG, i = t / CG;
            float sc = 1.f, sh = 0.f;
            if (af.pa != nullptr) {
                sc = s_sc[c];
                sh = s_sh[c];
            } else if (af.scale != nullptr) {
                sc = af.scale[(size_t)b * C + c0 + c];
                sh = af.shift[(size_t)b * C + c0 + c];
            }
            const float* xp = x + ((size_t)(b * N + i) * N) * C + c0 + c;
            float xr[N], od[N];
#pragma unroll
            for (int j = 0; j < N; ++j) xr[j] = fmaf(xp[(size_t)j * C], sc, sh);
            up_odd<N>(xr, od);
            float* row = tile + i * PITCH + c;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                row[(2 * j) * CG] = xr[j];
                row[(2 * j + 1) * CG] = od[j];
            }
        }
        __syncthreads();
    }

    // pass 2 (+3): columns
    for (int t = threadIdx.x; t < M * CG; t += blockDim.x) {
        const int c = t % CG, jj = t / CG;
        float a[M];
        if constexpr (MODE != MODE_DOWN2) {
            float col[N], od[N];
#pragma unroll
            for (int i = 0; i < N; ++i) col[i] = tile[i * PITCH + jj * CG + c];
            up_odd<N>(col, od);
            if constexpr (MODE == MODE_UP2) {
                float* yp = y + ((size_t)(b * M) * M + jj) * C + c0 + c;
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    yp[(size_t)(2 * i) * M * C] = apply_act<ACT>(col[i]);
                    yp[(size_t)(2 * i + 1) * M * C] = apply_act<ACT>(od[i]);
                }
                continue;
            } else {
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    a[2 * i] = apply_act<ACT>(col[i]);
                    a[2 * i + 1] = apply_act<ACT>(od[i]);
                }
            }
        } else {
            const float* xp = x + ((size_t)(b * M) * M + jj) * C + c0 + c;
#pragma unroll
            for (int m = 0; m < M; ++m) a[m] = xp[(size_t)m * M * C];
        }
        float yl[N];
        down_line<N>(a, yl);
#pragma unroll
        for (int i = 0; i < N; ++i) tile[i * PITCH + jj * CG + c] = yl[i];
    }
    if constexpr (MODE == MODE_UP2) return;
    __syncthreads();

    // pass 4: rows
    for (int t = threadIdx.x; t < N * CG; t += blockDim.x) {
        const int c = t % CG, i = t / CG;
        float a[M], yl[N];
        const float* row = tile + i * PITCH + c;
#pragma unroll
        for (int m = 0; m < M; ++m) a[m] = row[m * CG];
        down_line<N>(a, yl);
        float* yp = y + ((size_t)(b * N + i) * N) * C + c0 + c;
#pragma unroll
        for (int j = 0; j < N; ++j) yp[(size_t)j * C] = yl[j];
    }
}

// ------------------------------------------------------------------------------------------
// 32 x 32 planes: same algorithm, restructured for the instruction cache.  The fully unrolled kernel
// above is ~100 KB of straight-line FFMA for n = 32 and ncu shows it starved for instructions
// (stall_no_instruction ~ 4 per issue, FMA pipe 40 %; profiles/r01_ncu_filtered_act.md).  Here
//   * the row / column passes run as iterations of ONE phase loop, so the up-sampling body and the
//     down-sampling body each exist once in the binary instead of twice, and
//   * the down-sampler is rolled over blocks of 8 outputs: the tap pattern of a block is fixed
//     (immediates), and the line is rotated by 16 registers between blocks instead.
// Code size drops to ~30 KB; arithmetic and summation order are unchanged (bitwise-identical results).
template <int N>
__device__ __forceinline__ void down_rolled(float (&a)[2 * N], float* __restrict__ sdst, float* __restrict__ gdst,
                                            int stride, bool to_smem) {
    constexpr int M = 2 * N;
    const float sc = alt_sum_even<N>(a);     // rotation-invariant: the line is rotated by multiples of 16
#pragma unroll 1
    for (int blk = 0; blk < N / 8; ++blk) {
        float acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = fmaf(0.5f, a[2 * u], (u & 1) ? sc : -sc);
#pragma unroll
        for (int m = 1; m < M; m += 2) {
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u] = fmaf(tap_g<N>((2 * u - m) & (M - 1)), a[m], acc[u]);
        }
        if (to_smem) {
#pragma unroll
            for (int u = 0; u < 8; ++u) sdst[(blk * 8 + u) * stride] = acc[u];
        } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) gdst[(size_t)(blk * 8 + u) * stride] = acc[u];
        }
        // y[blk*8 + u] = sum_m g[2u - m] a[m + 16 blk]: rotate the line by 16 for the next block
        float tmp[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) tmp[k] = a[k];
#pragma unroll
        for (int m = 0; m < M - 16; ++m) a[m] = a[m + 16];
#pragma unroll
        for (int k = 0; k < 16; ++k) a[M - 16 + k] = tmp[k];
    }
}

template <int N, int CG, int MODE, int ACT>
__global__ void __launch_bounds__(256, 2)
resample_phased_kernel(const float* __restrict__ x, float* __restrict__ y, int C, const Affine af) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float tile[];
    __shared__ float s_sc[CG], s_sh[CG];
    if (MODE != MODE_DOWN2 && af.pa != nullptr) gn_prologue<CG>(af, blockIdx.y, blockIdx.x * CG, C, s_sc, s_sh);
    constexpr int PITCH = Tile<N, CG>::PITCH;
    constexpr int M = 2 * N;
    const int b = blockIdx.y;
    const int c0 = blockIdx.x * CG;
    constexpr int PH_BEGIN = (MODE == MODE_DOWN2) ? 1 : 0;
    constexpr int PH_END = (MODE == MODE_UP2) ? 2 : 3;

#pragma unroll 1
    for (int ph = PH_BEGIN; ph < PH_END; ++ph) {
        // phase 0: rows up (global -> tile); phase 1: columns up, act, down (tile -> tile, in place);
        // phase 2: rows down (tile -> global)
        const int ntask = (ph == 1 ? M : N) * CG;
        for (int t = threadIdx.x; t < ntask; t += blockDim.x) {
            const int c = t % CG, line = t / CG;
            // `phv` hides the phase from the optimiser inside the task loop: without it the loop is unswitched
            // per phase and every phase gets its own copy of the unrolled bodies again.
            int phv = ph;
            asm volatile("" : "+r"(phv));
            float a[M];
            if (phv < 2 && MODE != MODE_DOWN2) {
                float xr[N], od[N];
                if (phv == 0) {
                    float sc = 1.f, sh = 0.f;
                    if (af.pa != nullptr) {
                        sc = s_sc[c];
                        sh = s_sh[c];
                    } else if (af.scale != nullptr) {
                        sc = af.scale[(size_t)b * C + c0 + c];
                        sh = af.shift[(size_t)b * C + c0 + c];
                    }
                    const float* xp = x + ((size_t)(b * N + line) * N) * C + c0 + c;
#pragma unroll
                    for (int j = 0; j < N; ++j) xr[j] = fmaf(xp[(size_t)j * C], sc, sh);
                } else {
#pragma unroll
                    for (int i = 0; i < N; ++i) xr[i] = tile[i * PITCH + line * CG + c];
                }
                up_odd<N>(xr, od);
                if (phv == 0) {
                    float* row = tile + line * PITCH + c;
#pragma unroll
                    for (int j = 0; j < N; ++j) {
                        row[(2 * j) * CG] = xr[j];
                        row[(2 * j + 1) * CG] = od[j];
                    }
                    continue;
                }
                if constexpr (MODE == MODE_UP2) {
                    float* yp = y + ((size_t)(b * M) * M + line) * C + c0 + c;
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        yp[(size_t)(2 * i) * M * C] = apply_act<ACT>(xr[i]);
                        yp[(size_t)(2 * i + 1) * M * C] = apply_act<ACT>(od[i]);
                    }
                    continue;
                }
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    a[2 * i] = apply_act<ACT>(xr[i]);
                    a[2 * i + 1] = apply_act<ACT>(od[i]);
                }
            } else if (phv == 1) {      // MODE_DOWN2: the 2n x 2n input comes straight from global memory
                const float* xp = x + ((size_t)(b * M) * M + line) * C + c0 + c;
#pragma unroll
                for (int m = 0; m < M; ++m) a[m] = xp[(size_t)m * M * C];
            } else {
                const float* row = tile + line * PITCH + c;
#pragma unroll
                for (int m = 0; m < M; ++m) a[m] = row[m * CG];
            }
            if constexpr (MODE != MODE_UP2) {
                const bool col_phase = phv == 1;
                down_rolled<N>(a, tile + line * CG + c, y + ((size_t)(b * N + line) * N) * C + c0 + c,
                               col_phase ? PITCH : C, col_phase);
            }
        }
        __syncthreads();
    }
}

template <int N, int CG, int MODE, int ACT>
int launch_one(const float* x, float* y, int B, int C, const Affine& af, cudaStream_t st) {
    if (C % CG != 0) return AFLDM_E_SHAPE;
    auto kern = (N >= 32) ? resample_phased_kernel<N, CG, MODE, ACT> : resample_kernel<N, CG, MODE, ACT>;
    constexpr int smem = Tile<N, CG>::SMEM_BYTES;
    static bool configured = false;  // attribute is per-function, set once (idempotent)
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    constexpr int tasks = 2 * N * CG;
    const int threads = tasks >= 256 ? 256 : (tasks < 32 ? 32 : tasks);
    launch_k(kern, dim3(C / CG, B), dim3(threads), smem, st, x, y, C, af);
    return launched();
}

template <int MODE, int ACT>
int dispatch_n(const float* x, float* y, int B, int n, int C, const Affine& af, cudaStream_t st) {
    switch (n) {
        case 2: return launch_one<2, 32, MODE, ACT>(x, y, B, C, af, st);
        case 4: return launch_one<4, 32, MODE, ACT>(x, y, B, C, af, st);
        case 8: return launch_one<8, 32, MODE, ACT>(x, y, B, C, af, st);
        case 16: return launch_one<16, 16, MODE, ACT>(x, y, B, C, af, st);
        case 32: return launch_one<32, 8, MODE, ACT>(x, y, B, C, af, st);
        default: return AFLDM_E_NOKERNEL;
    }
}

Affine plain_affine(const float* scale, const float* shift) {
    Affine af{};
    af.scale = scale;
    af.shift = shift;
    return af;
}

bool bad_args(const float* x, const float* y, int B, int H, int W, int C, const float* scale,
              const float* shift) {
    return x == nullptr || y == nullptr || B <= 0 || H <= 0 || W <= 0 || C <= 0 ||
           ((scale == nullptr) != (shift == nullptr));
}

}  // namespace
}  // namespace afldm

using namespace afldm;

extern "C" size_t afldm_resample_workspace_floats(int op, int B, int H, int W, int C) {
    if (op < 0 || op > 2 || B <= 0 || H <= 0 || W <= 0 || C <= 0 || H != W) return 0;
    return resample_large_workspace_floats(op, B, H, C);   // 0 for the planes that run fused in one kernel
}

extern "C" int afldm_filtered_act_f32(const float* x, float* y, int B, int H, int W, int C, int act,
                                      const float* scale, const float* shift, float* workspace,
                                      size_t workspace_floats, afldm_stream_t stream) {
    if (bad_args(x, y, B, H, W, C, scale, shift)) return AFLDM_E_ARG;
    if (act != AFLDM_ACT_SILU && act != AFLDM_ACT_IDENTITY) return AFLDM_E_ARG;
    if (H != W) return AFLDM_E_SHAPE;  // the reference's mask is built from W only (ideal_lpf.py:81-88)
    cudaStream_t st = as_stream(stream);
    if (H > 32) return resample_large(MODE_FACT, act, x, y, B, H, C, scale, shift, workspace, workspace_floats, st);
    const Affine af = plain_affine(scale, shift);
    if (act == AFLDM_ACT_SILU) return dispatch_n<MODE_FACT, AFLDM_ACT_SILU>(x, y, B, H, C, af, st);
    return dispatch_n<MODE_FACT, AFLDM_ACT_IDENTITY>(x, y, B, H, C, af, st);
}

extern "C" int afldm_filtered_act_gn_f32(const float* x, float* y, int B, int H, int W, int C, int act,
                                         const float* partial_a, int slots_a, int Ca, const float* partial_b,
                                         int slots_b, int Cb, int groups, float eps, const float* gamma,
                                         const float* beta, afldm_stream_t stream) {
    if (bad_args(x, y, B, H, W, C, nullptr, nullptr) || partial_a == nullptr) return AFLDM_E_ARG;
    if (act != AFLDM_ACT_SILU && act != AFLDM_ACT_IDENTITY) return AFLDM_E_ARG;
    if (slots_a <= 0 || Ca <= 0 || Cb < 0 || groups <= 0 || (Cb > 0 && (partial_b == nullptr || slots_b <= 0)))
        return AFLDM_E_ARG;
    if (Ca + Cb != C || C % groups != 0) return AFLDM_E_SHAPE;
    if (H != W) return AFLDM_E_SHAPE;
    if (H > 32) return AFLDM_E_NOKERNEL;    // large planes: afldm_groupnorm_finalize_f32 + afldm_filtered_act_f32
    Affine af{};
    af.pa = reinterpret_cast<const float2*>(partial_a);
    af.pb = reinterpret_cast<const float2*>(partial_b);
    af.gamma = gamma; af.beta = beta;
    af.slots_a = slots_a; af.Ca = Ca; af.slots_b = Cb > 0 ? slots_b : 0; af.Cb = Cb;
    af.groups = groups; af.HW = H * W; af.eps = eps;
    cudaStream_t st = as_stream(stream);
    if (act == AFLDM_ACT_SILU) return dispatch_n<MODE_FACT, AFLDM_ACT_SILU>(x, y, B, H, C, af, st);
    return dispatch_n<MODE_FACT, AFLDM_ACT_IDENTITY>(x, y, B, H, C, af, st);
}

extern "C" int afldm_up2_ideal_f32(const float* x, float* y, int B, int H, int W, int C,
                                   const float* scale, const float* shift, float* workspace,
                                   size_t workspace_floats, afldm_stream_t stream) {
    if (bad_args(x, y, B, H, W, C, scale, shift) || x == y) return AFLDM_E_ARG;
    if (H != W) return AFLDM_E_SHAPE;
    cudaStream_t st = as_stream(stream);
    if (H > 32)
        return resample_large(MODE_UP2, AFLDM_ACT_IDENTITY, x, y, B, H, C, scale, shift, workspace, workspace_floats, st);
    return dispatch_n<MODE_UP2, AFLDM_ACT_IDENTITY>(x, y, B, H, C, plain_affine(scale, shift), st);
}

extern "C" int afldm_lpf_down2_f32(const float* x, float* y, int B, int H, int W, int C, float* workspace,
                                   size_t workspace_floats, afldm_stream_t stream) {
    if (bad_args(x, y, B, H, W, C, nullptr, nullptr) || x == y) return AFLDM_E_ARG;
    if (H != W) return AFLDM_E_SHAPE;
    cudaStream_t st = as_stream(stream);
    if (H > 32)
        return resample_large(MODE_DOWN2, AFLDM_ACT_IDENTITY, x, y, B, H, C, nullptr, nullptr, workspace,
                              workspace_floats, st);
    return dispatch_n<MODE_DOWN2, AFLDM_ACT_IDENTITY>(x, y, B, H, C, plain_affine(nullptr, nullptr), st);
}
