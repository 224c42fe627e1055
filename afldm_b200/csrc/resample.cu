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
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"
#include "fact_common.cuh"
#include "fact_tc.cuh"
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
    __shared__ float2 s_gn[MODE == MODE_DOWN2 ? N * CG : 1];
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
            const XSrc xs = x_source(x, C, af, c0);
            const float* xp = xs.p + ((size_t)(b * N + i) * N) * xs.pitch + c;
            float xr[N], od[N];
#pragma unroll
            for (int j = 0; j < N; ++j) xr[j] = fmaf(xp[(size_t)j * xs.pitch], sc, sh);
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
                const size_t yo = ((size_t)(b * M) * M + jj) * C + c0 + c;
                if (af.y_half) {
                    __half* yh = reinterpret_cast<__half*>(y) + yo;
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        yh[(size_t)(2 * i) * M * C] = __float2half_rn(apply_act<ACT>(col[i]));
                        yh[(size_t)(2 * i + 1) * M * C] = __float2half_rn(apply_act<ACT>(od[i]));
                    }
                    continue;
                }
                float* yp = y + yo;
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
        const size_t yo = ((size_t)(b * N + i) * N) * C + c0 + c;
        if (MODE == MODE_FACT && af.y_half) {
            __half* yh = reinterpret_cast<__half*>(y) + yo;
#pragma unroll
            for (int j = 0; j < N; ++j) yh[(size_t)j * C] = __float2half_rn(yl[j]);
        } else {
            float* yp = y + yo;
#pragma unroll
            for (int j = 0; j < N; ++j) yp[(size_t)j * C] = yl[j];
        }
        if constexpr (MODE == MODE_DOWN2) {
            if (af.gn_out != nullptr) {
                float ps = 0.f, pq = 0.f;
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    ps += yl[j];
                    pq = fmaf(yl[j], yl[j], pq);
                }
                s_gn[i * CG + c] = make_float2(ps, pq);
            }
        }
    }
    if constexpr (MODE == MODE_DOWN2) {
        // GroupNorm partial sums of the output plane (one slot per image): the next resnet's norm needs no pass
        // over y.  Rows are added in a fixed order (deterministic).
        if (af.gn_out != nullptr) {
            __syncthreads();
            for (int c = threadIdx.x; c < CG; c += blockDim.x) {
                float ps = 0.f, pq = 0.f;
                for (int i = 0; i < N; ++i) {
                    ps += s_gn[i * CG + c].x;
                    pq += s_gn[i * CG + c].y;
                }
                af.gn_out[(size_t)b * C + c0 + c] = make_float2(ps, pq);
            }
        }
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
                    const XSrc xs = x_source(x, C, af, c0);
                    const float* xp = xs.p + ((size_t)(b * N + line) * N) * xs.pitch + c;
#pragma unroll
                    for (int j = 0; j < N; ++j) xr[j] = fmaf(xp[(size_t)j * xs.pitch], sc, sh);
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
                    const size_t yo = ((size_t)(b * M) * M + line) * C + c0 + c;
                    if (af.y_half) {
                        __half* yh = reinterpret_cast<__half*>(y) + yo;
#pragma unroll
                        for (int i = 0; i < N; ++i) {
                            yh[(size_t)(2 * i) * M * C] = __float2half_rn(apply_act<ACT>(xr[i]));
                            yh[(size_t)(2 * i + 1) * M * C] = __float2half_rn(apply_act<ACT>(od[i]));
                        }
                        continue;
                    }
                    float* yp = y + yo;
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

// ------------------------------------------------------------------------------------------
// Filtered activation on the tensor cores (n = 16, 32): every 1-D circular convolution of the separable operator
// is a small GEMM  out[line][i] = sum_j in[line][j] * F[(i - j) mod n]  with 16 lines (2 rows or columns x 8
// channels) as the M dimension of a warp-level mma.m16n8k16.  fp32 accuracy comes from a 3-term split into fp16
// halves (x = xh + xl, F = Fh + Fl; out = xh Fh + xl Fh + xh Fl, fp32 accumulate): 22 significant bits per operand,
// measured error 3e-7 of max |y| - the same class as the exact-FMA kernels above (tests compare both with the
// reference's FFT result at 5e-6 / 1e-5).  Inputs are post-GroupNorm activations; |x| must stay below the fp16
// range (65504) - the host wrapper documents this and the SIMT kernel remains selectable (AFLDM_FACT_MMA=0).
//
//   * the circulant is held as B fragments in registers: fragment (k-step ks, n-tile nt) depends only on
//     (nt - 2 ks) mod n/8, so n/8 fragment pairs (hi, lo) cover the whole filter;
//   * an accumulator fragment (cols 2t, 2t+1 of n-tile j) is exactly half of the A fragment of the next product
//     (k = 2t, 2t+1 of k-step j/2): up-sample -> act -> down-sample of a column chain through registers;
//   * rows and columns are exchanged through one fp32 shared-memory tile T[i][j'][c] with
//     addr = i * (20 n + 4) + 10 * (j' / 4) * 4 ... (see tix()): every access pattern below is bank-conflict free;
//   * the even-indexed half of the down-sampler is the O(n) identity of down_line().
constexpr int FM_CG = 8;        // channels per CTA
constexpr int FM_THREADS = 256;

template <int N>
struct FmTile {
    static constexpr int ROW = 20 * N + 4;                 // floats per image row i: 2N * 8 + (2N / 4) * 8 pad + 4
    static constexpr int SMEM_BYTES = N * ROW * 4;
};
// offset of (row i, up-sampled column jp, channel 0)
template <int N>
__device__ __forceinline__ int tix(int i, int jp) { return i * FmTile<N>::ROW + jp * 8 + (jp >> 2) * 8; }

__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// B fragments of the circulant F[(i - j) mod N]: variant q = (nt - 2 ks) mod N/8.
template <int N>
struct CircB {
    uint32_t h[N / 8][2], l[N / 8][2];
};
template <int N, bool DOWN>
__device__ __forceinline__ float circ_tap(int r) {
    // up: odd-phase taps d[r];  down (odd samples): G[r] = g[(2r - 1) mod 2N]
    if constexpr (DOWN) return tap_g<N>((2 * r - 1) & (2 * N - 1));
    return tap_d<N>(r & (N - 1));
}
template <int N, bool DOWN>
__device__ __forceinline__ void make_circ(CircB<N>& f, int g, int t) {
#pragma unroll
    for (int q = 0; q < N / 8; ++q) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int kk = 2 * t + 8 * p;                  // B[k][n = g], k = kk, kk + 1
            const float v0 = circ_tap<N, DOWN>((8 * q + g - kk) & (N - 1));
            const float v1 = circ_tap<N, DOWN>((8 * q + g - kk - 1) & (N - 1));
            split_pack(v0, v1, f.h[q][p], f.l[q][p]);
        }
    }
}

// acc[nt][.] += sum_j in[.][j] F[(i - j) mod N] for the 16 lines of the warp; `in` is in accumulator layout:
// in[r][nt][e] = line (g + 8 r), position 8 nt + 2 t + e.
// TERMS = 3: x F ~ xh Fh + xl Fh + xh Fl (22 significant bits per operand: fp32 accuracy, the default);
// TERMS = 2: xh (Fh + Fl) - exact filter, operand rounded to 11 bits; TERMS = 1: xh Fh - the plain TF32-class product
// (opt-in experiments, AFLDM_FACT_TERMS).
template <int N, int TERMS>
__device__ __forceinline__ void circ_mma(const float (&in)[2][N / 8][2], const CircB<N>& f, float (&acc)[N / 8][4]) {
    uint32_t ah[N / 16][4], al[N / 16][4];
#pragma unroll
    for (int ks = 0; ks < N / 16; ++ks) {
#pragma unroll
        for (int idx = 0; idx < 4; ++idx) {
            const int r = idx & 1, nt = 2 * ks + (idx >> 1);
            if constexpr (TERMS == 3) {
                split_pack(in[r][nt][0], in[r][nt][1], ah[ks][idx], al[ks][idx]);
            } else {
                const __half2 h = __floats2half2_rn(in[r][nt][0], in[r][nt][1]);
                ah[ks][idx] = *reinterpret_cast<const uint32_t*>(&h);
                al[ks][idx] = 0u;
            }
        }
    }
    // consecutive MMAs go to different accumulators (N/8 independent chains); small terms first
#pragma unroll
    for (int term = 3 - TERMS; term < 3; ++term) {
#pragma unroll
        for (int ks = 0; ks < N / 16; ++ks) {
#pragma unroll
            for (int nt = 0; nt < N / 8; ++nt) {
                const int q = (nt - 2 * ks) & (N / 8 - 1);
                if (term == 0) mma_f16(acc[nt], al[ks], f.h[q][0], f.h[q][1]);
                else if (term == 1) mma_f16(acc[nt], ah[ks], f.l[q][0], f.l[q][1]);
                else mma_f16(acc[nt], ah[ks], f.h[q][0], f.h[q][1]);
            }
        }
    }
}

// y = down-sample of the line whose even samples are e[] and odd samples o[] (both activated):
// y[i] = e[i] / 2 - (-1)^i altsum(e) / (2N) + sum_m G[i - m] o[m]
template <int N, int TERMS>
__device__ __forceinline__ void down_mma(const float (&e)[2][N / 8][2], const float (&o)[2][N / 8][2],
                                         const CircB<N>& fd, float (&acc)[N / 8][4]) {
    float s[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        float a = 0.f;
#pragma unroll
        for (int nt = 0; nt < N / 8; ++nt) a += e[r][nt][0] - e[r][nt][1];
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        s[r] = a * (1.0f / (2 * N));
    }
#pragma unroll
    for (int nt = 0; nt < N / 8; ++nt) {
        acc[nt][0] = fmaf(0.5f, e[0][nt][0], -s[0]);
        acc[nt][1] = fmaf(0.5f, e[0][nt][1], s[0]);
        acc[nt][2] = fmaf(0.5f, e[1][nt][0], -s[1]);
        acc[nt][3] = fmaf(0.5f, e[1][nt][1], s[1]);
    }
    circ_mma<N, TERMS>(o, fd, acc);
}

template <int N, int ACT, int TERMS>
__global__ void __launch_bounds__(FM_THREADS, 2)
fact_mma_kernel(const float* __restrict__ x, float* __restrict__ y, int C, const Affine af) {
    pdl_trigger();
    extern __shared__ float T[];
    __shared__ float s_sc[FM_CG], s_sh[FM_CG];
    const int b = blockIdx.y, c0 = blockIdx.x * FM_CG;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    constexpr int NT = N / 8;
    constexpr int NWARPS = FM_THREADS / 32;

    // filter fragments depend on nothing the previous kernel wrote: built before the dependency wait, under its tail
    CircB<N> fu, fd;
    make_circ<N, false>(fu, g, t);
    make_circ<N, true>(fd, g, t);
    pdl_wait();
    if (af.pa != nullptr) gn_prologue<FM_CG>(af, b, c0, C, s_sc, s_sh);

    float sc = 1.f, sh = 0.f;
    if (af.pa != nullptr) {
        sc = s_sc[g];
        sh = s_sh[g];
    } else if (af.scale != nullptr) {
        sc = af.scale[(size_t)b * C + c0 + g];
        sh = af.shift[(size_t)b * C + c0 + g];
    }

    // ---- stage: x[b][:, :, c0 .. c0 + 8) -> the even columns of T, 2 x 16 B per pixel with cp.async, all in
    // flight at once (one thread per half pixel), so that no MMA chain below waits on a global load.
    const XSrc xs = x_source(x, C, af, c0);
    if ((xs.pitch & 3) == 0 && (reinterpret_cast<uintptr_t>(xs.p) & 15u) == 0) {
        for (int idx = threadIdx.x; idx < N * N * 2; idx += FM_THREADS) {
            const int half = idx & 1, pix = idx >> 1, i = pix / N, j = pix - i * N;
            const float* src = xs.p + ((size_t)(b * N + i) * N + j) * xs.pitch + 4 * half;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(T + tix<N>(i, 2 * j) + 4 * half);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
        for (int idx = threadIdx.x; idx < N * N * FM_CG; idx += FM_THREADS) {
            const int c = idx & 7, pix = idx >> 3, i = pix / N, j = pix - i * N;
            T[tix<N>(i, 2 * j) + c] = xs.p[((size_t)(b * N + i) * N + j) * xs.pitch + c];
        }
    }
    __syncthreads();

    // ---- phase 0: rows up.  m-tile = rows (i0, i0 + 1) x 8 channels; line g -> (i0, c = g), g + 8 -> (i0 + 1, g)
    for (int mt = warp; mt < N / 2; mt += NWARPS) {
        const int i0 = 2 * mt;
        float e[2][NT][2];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    e[r][nt][q] = fmaf(T[tix<N>(i0 + r, 16 * nt + 4 * t + 2 * q) + g], sc, sh);
        float acc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        circ_mma<N, TERMS>(e, fu, acc);
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    float* tp = T + tix<N>(i0 + r, 16 * nt + 4 * t + 2 * q) + g;
                    tp[0] = e[r][nt][q];
                    tp[8] = acc[nt][2 * r + q];
                }
    }
    __syncthreads();

    // ---- phase 1: columns up -> act -> down, in place.  m-tile = columns (jp0, jp0 + 1) x 8 channels
    for (int mt = warp; mt < N; mt += NWARPS) {
        const int jp0 = 2 * mt;
        float e[2][NT][2], o[2][NT][2];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int q = 0; q < 2; ++q) e[r][nt][q] = T[tix<N>(8 * nt + 2 * t + q, jp0 + r) + g];
        {
            float acc[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
            circ_mma<N, TERMS>(e, fu, acc);
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        o[r][nt][q] = apply_act<ACT>(acc[nt][2 * r + q]);
                        e[r][nt][q] = apply_act<ACT>(e[r][nt][q]);
                    }
        }
        float acc[NT][4];
        down_mma<N, TERMS>(e, o, fd, acc);
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int q = 0; q < 2; ++q) T[tix<N>(8 * nt + 2 * t + q, jp0 + r) + g] = acc[nt][2 * r + q];
    }
    __syncthreads();

    // ---- phase 2: rows down -> global
    for (int mt = warp; mt < N / 2; mt += NWARPS) {
        const int i0 = 2 * mt;
        float e[2][NT][2], o[2][NT][2];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const float* tp = T + tix<N>(i0 + r, 16 * nt + 4 * t + 2 * q) + g;
                    e[r][nt][q] = tp[0];
                    o[r][nt][q] = tp[8];
                }
        float acc[NT][4];
        down_mma<N, TERMS>(e, o, fd, acc);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const size_t yo = ((size_t)(b * N + i0 + r) * N) * C + c0 + g;
            if (af.y_half) {
                __half* yh = reinterpret_cast<__half*>(y) + yo;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int q = 0; q < 2; ++q) yh[(size_t)(8 * nt + 2 * t + q) * C] = __float2half_rn(acc[nt][2 * r + q]);
            } else {
                float* yp = y + yo;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int q = 0; q < 2; ++q) yp[(size_t)(8 * nt + 2 * t + q) * C] = acc[nt][2 * r + q];
            }
        }
    }
}

template <int N, int ACT, int TERMS>
int launch_fact_mma_t(const float* x, float* y, int B, int C, const Affine& af, cudaStream_t st) {
    auto kern = fact_mma_kernel<N, ACT, TERMS>;
    constexpr int smem = FmTile<N>::SMEM_BYTES;
    static std::atomic<unsigned long long> configured{0};      // one bit per device
    {
        cudaError_t e = set_max_dyn_smem(kern, smem, configured);
        if (e != cudaSuccess) return (int)e;
    }
    launch_k(kern, dim3(C / FM_CG, B), dim3(FM_THREADS), smem, st, x, y, C, af);
    return launched();
}

template <int N, int ACT>
int launch_fact_mma(const float* x, float* y, int B, int C, const Affine& af, cudaStream_t st) {
    // split terms per product: 3 = fp32 accuracy (default); 1 / 2 only when the result is stored as fp16 anyway
    static const int terms = getenv("AFLDM_FACT_TERMS") ? atoi(getenv("AFLDM_FACT_TERMS")) : 3;
    if (af.y_half && terms == 1) return launch_fact_mma_t<N, ACT, 1>(x, y, B, C, af, st);
    if (af.y_half && terms == 2) return launch_fact_mma_t<N, ACT, 2>(x, y, B, C, af, st);
    return launch_fact_mma_t<N, ACT, 3>(x, y, B, C, af, st);
}

bool fact_mma_enabled() {
    static const bool on = !(getenv("AFLDM_FACT_MMA") && atoi(getenv("AFLDM_FACT_MMA")) == 0);
    return on;
}

template <int N, int CG, int MODE, int ACT>
int launch_one(const float* x, float* y, int B, int C, const Affine& af, cudaStream_t st) {
    if (C % CG != 0) return AFLDM_E_SHAPE;
    auto kern = (N >= 32) ? resample_phased_kernel<N, CG, MODE, ACT> : resample_kernel<N, CG, MODE, ACT>;
    constexpr int smem = Tile<N, CG>::SMEM_BYTES;
    static std::atomic<unsigned long long> configured{0};      // one bit per device
    {
        cudaError_t e = set_max_dyn_smem(kern, smem, configured);
        if (e != cudaSuccess) return (int)e;
    }
    constexpr int tasks = 2 * N * CG;
    const int threads = tasks >= 256 ? 256 : (tasks < 32 ? 32 : tasks);
    launch_k(kern, dim3(C / CG, B), dim3(threads), smem, st, x, y, C, af);
    return launched();
}

template <int MODE, int ACT>
int dispatch_n(const float* x, float* y, int B, int n, int C, const Affine& af, cudaStream_t st) {
    if constexpr (MODE == MODE_FACT) {
        if (fact_tc_enabled(n)) {           // tcgen05 form (fact_tc.cu); falls through when the shape is outside its family
            const int r = fact_tc_launch(n, ACT, x, y, B, C, af, st);
            if (r != AFLDM_E_NOKERNEL) return r;
        }
        if (fact_mma_enabled() && C % FM_CG == 0) {
            if (n == 32) return launch_fact_mma<32, ACT>(x, y, B, C, af, st);
            if (n == 16) return launch_fact_mma<16, ACT>(x, y, B, C, af, st);
        }
    }
    // fp16 stores: fact_mma (n = 16, 32), the register-resident kernels (n <= 16) and the phased up-sampler (n = 32)
    if (af.y_half && (MODE == MODE_DOWN2 || (n >= 32 && MODE != MODE_UP2))) return AFLDM_E_NOKERNEL;
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

static int filtered_act_gn_impl(const float* x, float* y, int y_half, int B, int H, int W, int C, int act,
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
    af.inv_n = 1.0 / ((double)(H * W) * (double)(C / groups));
    af.y_half = y_half;
    cudaStream_t st = as_stream(stream);
    if (act == AFLDM_ACT_SILU) return dispatch_n<MODE_FACT, AFLDM_ACT_SILU>(x, y, B, H, C, af, st);
    return dispatch_n<MODE_FACT, AFLDM_ACT_IDENTITY>(x, y, B, H, C, af, st);
}

extern "C" int afldm_filtered_act_gn_f32(const float* x, float* y, int B, int H, int W, int C, int act,
                                         const float* partial_a, int slots_a, int Ca, const float* partial_b,
                                         int slots_b, int Cb, int groups, float eps, const float* gamma,
                                         const float* beta, afldm_stream_t stream) {
    return filtered_act_gn_impl(x, y, 0, B, H, W, C, act, partial_a, slots_a, Ca, partial_b, slots_b, Cb, groups, eps, gamma,
                                beta, stream);
}

extern "C" int afldm_filtered_act_gn_f16out(const float* x, void* y, int B, int H, int W, int C, int act,
                                            const float* partial_a, int slots_a, int Ca, const float* partial_b,
                                            int slots_b, int Cb, int groups, float eps, const float* gamma,
                                            const float* beta, afldm_stream_t stream) {
    if (static_cast<const void*>(x) == y) return AFLDM_E_ARG;          // element sizes differ: no in-place form
    return filtered_act_gn_impl(x, static_cast<float*>(y), 1, B, H, W, C, act, partial_a, slots_a, Ca, partial_b, slots_b, Cb,
                                groups, eps, gamma, beta, stream);
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

extern "C" int afldm_up2_ideal_f16out(const float* x, void* y, int B, int H, int W, int C, float* workspace,
                                      size_t workspace_floats, afldm_stream_t stream) {
    if (bad_args(x, static_cast<const float*>(y), B, H, W, C, nullptr, nullptr) || static_cast<const void*>(x) == y)
        return AFLDM_E_ARG;
    if (H != W) return AFLDM_E_SHAPE;
    cudaStream_t st = as_stream(stream);
    if (H > 32)
        return resample_large(MODE_UP2, AFLDM_ACT_IDENTITY, x, static_cast<float*>(y), B, H, C, nullptr, nullptr, workspace,
                              workspace_floats, st, 1);
    Affine af = plain_affine(nullptr, nullptr);
    af.y_half = 1;
    return dispatch_n<MODE_UP2, AFLDM_ACT_IDENTITY>(x, static_cast<float*>(y), B, H, C, af, st);
}

extern "C" int afldm_filtered_act_f16out(const float* x, void* y, int B, int H, int W, int C, int act,
                                         const float* scale, const float* shift, float* workspace,
                                         size_t workspace_floats, afldm_stream_t stream) {
    float* yf = static_cast<float*>(y);
    if (bad_args(x, yf, B, H, W, C, scale, shift) || static_cast<const void*>(x) == y) return AFLDM_E_ARG;
    if (act != AFLDM_ACT_SILU && act != AFLDM_ACT_IDENTITY) return AFLDM_E_ARG;
    if (H != W) return AFLDM_E_SHAPE;
    cudaStream_t st = as_stream(stream);
    if (H > 32) return resample_large(MODE_FACT, act, x, yf, B, H, C, scale, shift, workspace, workspace_floats, st, 1);
    Affine af = plain_affine(scale, shift);
    af.y_half = 1;
    if (act == AFLDM_ACT_SILU) return dispatch_n<MODE_FACT, AFLDM_ACT_SILU>(x, yf, B, H, C, af, st);
    return dispatch_n<MODE_FACT, AFLDM_ACT_IDENTITY>(x, yf, B, H, C, af, st);
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

extern "C" int afldm_lpf_down2_gn_f32(const float* x, float* y, int B, int H, int W, int C, float* gn_partial,
                                      afldm_stream_t stream) {
    if (bad_args(x, y, B, H, W, C, nullptr, nullptr) || x == y || gn_partial == nullptr) return AFLDM_E_ARG;
    if (H != W) return AFLDM_E_SHAPE;
    if (H > 16) return AFLDM_E_NOKERNEL;      // statistics are emitted by the register-resident kernels (n <= 16)
    Affine af = plain_affine(nullptr, nullptr);
    af.gn_out = reinterpret_cast<float2*>(gn_partial);
    return dispatch_n<MODE_DOWN2, AFLDM_ACT_IDENTITY>(x, y, B, H, C, af, as_stream(stream));
}

static int filtered_act_gn_cat_impl(const float* xa, const float* xb, float* y, int y_half, int B, int H, int W, int Ca,
                                    int Cb, int act, const float* partial_a, int slots_a,
                                    const float* partial_b, int slots_b, int groups, float eps,
                                    const float* gamma, const float* beta, afldm_stream_t stream) {
    const int C = Ca + Cb;
    if (xa == nullptr || xb == nullptr || y == nullptr || partial_a == nullptr || partial_b == nullptr) return AFLDM_E_ARG;
    if (B <= 0 || H <= 0 || W <= 0 || Ca <= 0 || Cb <= 0 || slots_a <= 0 || slots_b <= 0 || groups <= 0) return AFLDM_E_ARG;
    if (act != AFLDM_ACT_SILU && act != AFLDM_ACT_IDENTITY) return AFLDM_E_ARG;
    if (C % groups != 0 || H != W) return AFLDM_E_SHAPE;
    if (H > 32) return AFLDM_E_NOKERNEL;
    // a channel group of one CTA (8 / 32 channels) must not straddle the two sources
    const int cg = (H >= 16 && fact_mma_enabled()) ? FM_CG : (H == 32 ? 8 : (H == 16 ? 16 : 32));
    if (Ca % cg != 0 || Cb % cg != 0) return AFLDM_E_NOKERNEL;
    Affine af{};
    af.pa = reinterpret_cast<const float2*>(partial_a);
    af.pb = reinterpret_cast<const float2*>(partial_b);
    af.gamma = gamma; af.beta = beta;
    af.slots_a = slots_a; af.Ca = Ca; af.slots_b = slots_b; af.Cb = Cb;
    af.groups = groups; af.HW = H * W; af.eps = eps;
    af.inv_n = 1.0 / ((double)(H * W) * (double)(C / groups));
    af.x2 = xb; af.xCa = Ca;
    af.y_half = y_half;
    cudaStream_t st = as_stream(stream);
    if (act == AFLDM_ACT_SILU) return dispatch_n<MODE_FACT, AFLDM_ACT_SILU>(xa, y, B, H, C, af, st);
    return dispatch_n<MODE_FACT, AFLDM_ACT_IDENTITY>(xa, y, B, H, C, af, st);
}

extern "C" int afldm_filtered_act_gn_cat_f32(const float* xa, const float* xb, float* y, int B, int H, int W, int Ca,
                                             int Cb, int act, const float* partial_a, int slots_a,
                                             const float* partial_b, int slots_b, int groups, float eps,
                                             const float* gamma, const float* beta, afldm_stream_t stream) {
    return filtered_act_gn_cat_impl(xa, xb, y, 0, B, H, W, Ca, Cb, act, partial_a, slots_a, partial_b, slots_b, groups, eps,
                                    gamma, beta, stream);
}

extern "C" int afldm_filtered_act_gn_cat_f16out(const float* xa, const float* xb, void* y, int B, int H, int W, int Ca,
                                                int Cb, int act, const float* partial_a, int slots_a,
                                                const float* partial_b, int slots_b, int groups, float eps,
                                                const float* gamma, const float* beta, afldm_stream_t stream) {
    return filtered_act_gn_cat_impl(xa, xb, static_cast<float*>(y), 1, B, H, W, Ca, Cb, act, partial_a, slots_a, partial_b,
                                    slots_b, groups, eps, gamma, beta, stream);
}

// The tcgen05 form of the filtered activation, selected explicitly (tests and the bench's side measurement; the default
// dispatch of afldm_filtered_act_* uses it only under AFLDM_FACT_TC=1).  y_half: y holds IEEE binary16.
extern "C" int afldm_filtered_act_tc(const float* x, void* y, int y_half, int B, int H, int W, int C, int act,
                                     const float* scale, const float* shift, afldm_stream_t stream) {
    float* yf = static_cast<float*>(y);
    if (bad_args(x, yf, B, H, W, C, scale, shift) || static_cast<const void*>(x) == y) return AFLDM_E_ARG;
    if (act != AFLDM_ACT_SILU && act != AFLDM_ACT_IDENTITY) return AFLDM_E_ARG;
    if (H != W) return AFLDM_E_SHAPE;
    Affine af = plain_affine(scale, shift);
    af.y_half = y_half ? 1 : 0;
    return fact_tc_launch(H, act, x, yf, B, C, af, as_stream(stream));
}
