// Ideal resamplers for large planes (n = 64, 128: the VAE decoder and 64x64-latent UNets), where a
// plane no longer fits the register / shared-memory scheme of resample.cu.
//
// The separable operator  y = D act(U x U^T) D^T  is run as three line passes with fp32
// intermediates in global memory (rows up -> columns up+act+down -> rows down); every pass is a set
// of independent 1-D circular convolutions, one line per thread:
//      out[i] = sum_j taps[(i - j - sh) mod n] * in[j]
// The thread's line sits in shared memory ([pos][line][channel], conflict-free), outputs are produced
// 8 at a time in registers, and the taps - indexed by a warp-uniform run-time value - come from
// constant memory (15 loads per 64 FMAs).  The down-sampler uses its polyphase form
//      y[i] = sum_s ge[s] Ae[i - s] + go[s] Ao[i - s - 1],   Ae = a[0::2], Ao = a[1::2],
// of which only the odd phase is a convolution of length n: ge is delta/2 minus a rank-one alternating term.
//
// This is the exact-fp32 SIMT path: correct for every supported n, FMA-bound.  (A tensor-core
// formulation with split operands is the planned replacement for the VAE-sized planes.)
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"
#include "resample.cuh"
#include "taps.inc"

namespace afldm {
namespace {

enum { TAB_D = 0, TAB_GE = 1, TAB_GO = 2 };

template <int N, int TAB>
__device__ __forceinline__ float ctap(int r) {
    const int i = r & (N - 1);
    if constexpr (N == 64) {
        if constexpr (TAB == TAB_D) return cc_tap_d64[i];
        if constexpr (TAB == TAB_GE) return cc_tap_ge64[i];
        return cc_tap_go64[i];
    } else {
        if constexpr (TAB == TAB_D) return cc_tap_d128[i];
        if constexpr (TAB == TAB_GE) return cc_tap_ge128[i];
        return cc_tap_go128[i];
    }
}

// acc[u] += sum_j taps[(i0 + u - j - sh) mod N] * in[j * stride],  u = 0..7
template <int N, int TAB>
__device__ __forceinline__ void circ8(float (&acc)[8], const float* __restrict__ in, int stride, int i0, int sh) {
    for (int j0 = 0; j0 < N; j0 += 8) {
        float t[15];
        const int base = i0 - j0 - sh - 7;
#pragma unroll
        for (int k = 0; k < 15; ++k) t[k] = ctap<N, TAB>(base + k);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const float a = in[(j0 + jj) * stride];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u] = fmaf(t[7 + u - jj], a, acc[u]);
        }
    }
}

enum { OP_UP = 0, OP_UPACTDOWN = 1, OP_DOWN = 2 };

struct LineArgs {
    const float* x;
    float* y;
    const float* scale;   // per-(b, c) affine applied on load (OP_UP rows of the filtered activation), or NULL
    const float* shift;
    long long in_outer, in_inner, in_pos;      // element strides: line = outer * n_inner + inner
    long long out_outer, out_inner, out_pos;
    int n_inner, n_lines, C, lines_per_image;  // lines_per_image: lines sharing one batch index (for scale/shift)
    int y_half;                                // y holds IEEE binary16 (same element strides); tensor-core passes only
};

// N: SMALL length (input of UP / UPACTDOWN, output of DOWN).  CTA = 32 channels x L lines.
template <int N, int OP, int ACT, int L>
__global__ void __launch_bounds__(32 * L)
line_kernel(const LineArgs a) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sm[];
    constexpr int LEN_A = (OP == OP_DOWN) ? 2 * N : N;
    float* bufA = sm;                               // [LEN_A][L][32]
    float* bufB = sm + (size_t)LEN_A * L * 32;      // [N][L][32]   (odd phase; OP_UPACTDOWN only)
    const int c = threadIdx.x & 31, l = threadIdx.x >> 5;
    const int line = blockIdx.y * L + l;
    const int ch = blockIdx.x * 32 + c;
    if (line >= a.n_lines) return;                  // no block-wide barrier below: threads are independent
    const int outer = line / a.n_inner, inner = line - outer * a.n_inner;
    const float* xin = a.x + outer * a.in_outer + inner * a.in_inner + ch;
    float* yout = a.y + outer * a.out_outer + inner * a.out_inner + ch;
    constexpr int S = L * 32;                       // smem stride between consecutive positions of a line
    float* A = bufA + l * 32 + c;
    float* Bq = bufB + l * 32 + c;

    float sc = 1.f, sh = 0.f;
    if (a.scale != nullptr) {
        const int b = line / a.lines_per_image;
        sc = a.scale[(size_t)b * a.C + ch];
        sh = a.shift[(size_t)b * a.C + ch];
    }
    if constexpr (OP == OP_DOWN) {
        // de-interleave on load: A[0..N) = even samples, A[N..2N) = odd samples
        for (int m = 0; m < 2 * N; ++m) A[((m >> 1) + (m & 1) * N) * S] = xin[m * a.in_pos];
    } else {
        for (int j = 0; j < N; ++j) A[j * S] = fmaf(xin[j * a.in_pos], sc, sh);
    }
    // (each thread reads back only what it wrote itself: no barrier needed)

    if constexpr (OP == OP_UP || OP == OP_UPACTDOWN) {
        for (int i0 = 0; i0 < N; i0 += 8) {
            float acc[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u] = 0.f;
            circ8<N, TAB_D>(acc, A, S, i0, 0);
            if constexpr (OP == OP_UP) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    yout[(2 * (i0 + u)) * a.out_pos] = apply_act<ACT>(A[(i0 + u) * S]);
                    yout[(2 * (i0 + u) + 1) * a.out_pos] = apply_act<ACT>(acc[u]);
                }
            } else {
#pragma unroll
                for (int u = 0; u < 8; ++u) Bq[(i0 + u) * S] = apply_act<ACT>(acc[u]);
            }
        }
        if constexpr (OP == OP_UP) return;
        for (int j = 0; j < N; ++j) A[j * S] = apply_act<ACT>(A[j * S]);   // even phase, after the odd phase used it raw
    }
    // polyphase down: Ae = A[0..N), Ao = (OP_DOWN ? A[N..2N) : Bq[0..N))
    const float* Ao = (OP == OP_DOWN) ? (A + (size_t)N * S) : Bq;
    // even phase without a convolution: ge[r] = delta[r] / 2 - (-1)^r / (2N)  (pass band = bins |k| < N/2, see
    // resample.cu down_line), so  sum_s ge[s] Ae[i - s] = Ae[i] / 2 - (-1)^i * altsum(Ae) / (2N)
    float s0 = 0.f, s1 = 0.f;
    for (int j = 0; j < N; j += 2) {
        s0 += A[j * S];
        s1 += A[(j + 1) * S];
    }
    const float scorr = (s0 - s1) * (1.0f / (2 * N));
    for (int i0 = 0; i0 < N; i0 += 8) {
        float acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = fmaf(0.5f, A[(i0 + u) * S], (u & 1) ? scorr : -scorr);
        circ8<N, TAB_GO>(acc, Ao, S, i0, 1);
#pragma unroll
        for (int u = 0; u < 8; ++u) yout[(i0 + u) * a.out_pos] = acc[u];
    }
}

template <int N, int OP, int ACT>
int launch_lines(const LineArgs& a, cudaStream_t st) {
    constexpr int L = (N == 64) ? 8 : 4;
    constexpr int LEN_A = (OP == OP_DOWN) ? 2 * N : N;
    constexpr int smem = (LEN_A + (OP == OP_UPACTDOWN ? N : 0)) * L * 32 * 4;
    auto kern = line_kernel<N, OP, ACT, L>;
    static std::atomic<unsigned long long> configured{0};      // one bit per device
    {
        cudaError_t e = set_max_dyn_smem(kern, smem, configured);
        if (e != cudaSuccess) return (int)e;
    }
    launch_k(kern, dim3(a.C / 32, ceil_div(a.n_lines, L)), dim3(32 * L), smem, st, a);
    return 0;
}

// rows of an NHWC tensor [B][H][Win][C] -> [B][H][Wout][C]
LineArgs rows(const float* x, float* y, int B, int H, int Win, int Wout, int C) {
    LineArgs a{};
    a.x = x; a.y = y;
    a.in_outer = (long long)Win * C; a.in_inner = 0; a.in_pos = C;
    a.out_outer = (long long)Wout * C; a.out_inner = 0; a.out_pos = C;
    a.n_inner = 1; a.n_lines = B * H; a.C = C; a.lines_per_image = H;
    return a;
}
// columns of [B][Hin][W][C] -> [B][Hout][W][C]
LineArgs cols(const float* x, float* y, int B, int Hin, int Hout, int W, int C) {
    LineArgs a{};
    a.x = x; a.y = y;
    a.in_outer = (long long)Hin * W * C; a.in_inner = C; a.in_pos = (long long)W * C;
    a.out_outer = (long long)Hout * W * C; a.out_inner = C; a.out_pos = (long long)W * C;
    a.n_inner = W; a.n_lines = B * W; a.C = C; a.lines_per_image = W;
    return a;
}

// ------------------------------------------------------------------------------------------
// Tensor-core line passes (default): the same three passes, but every 1-D circular convolution is a GEMM on
// mma.sync.m16n8k16 with the 3-term fp16 split of resample.cu's fact_mma_kernel (fp32 accuracy, inputs < 65504):
// 16 lines (2 adjacent lines x 8 channels) are the M dimension of a warp, the circulant's B fragments sit in shared
// memory ([variant q = (nt - 2 ks) mod n/8][lane] = hi/lo pairs, one LDS.128 per 3 MMAs), outputs are produced four
// n-tiles at a time.  The even half of the down-sampler is the O(n) half-band identity.  6 n^3 MACs per plane at
// tensor-core rate instead of FMA rate: the 128 x 128 x 256-channel activation of the VAE decoder drops from 4.4 ms
// to ~1 ms per call at B = 16.
constexpr int LM_WARPS = 4;          // warps per CTA = m-tiles (2 lines x 8 channels) per CTA
constexpr int LM_EP = 20;            // floats per position of a warp's even-sample tile (16 lines + 4 pad: conflict-free)

__device__ __forceinline__ void lm_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void lm_split(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(v0, v1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
template <int N, bool DOWN>
__device__ __forceinline__ float lm_tap(int r) {
    // up: d[r];  down (odd samples): G[r] = g[2r - 1] = go[r - 1]  (go[s] = g[2s + 1])
    if constexpr (DOWN) return ctap<N, TAB_GO>(r - 1);
    return ctap<N, TAB_D>(r);
}
// B fragments of the circulant F[(i - j) mod N] for this lane, variant q: {h(k 2t..), h(k 2t+8..), l(..), l(..)}
template <int N, bool DOWN>
__device__ __forceinline__ void lm_fill_filter(uint4* f, int tid, int nthreads) {
    for (int idx = tid; idx < (N / 8) * 32; idx += nthreads) {
        const int q = idx >> 5, lane = idx & 31, g = lane >> 2, t = lane & 3;
        uint32_t h[2], l[2];
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int kk = 2 * t + 8 * p;
            lm_split(lm_tap<N, DOWN>((8 * q + g - kk) & (N - 1)), lm_tap<N, DOWN>((8 * q + g - kk - 1) & (N - 1)), h[p], l[p]);
        }
        f[idx] = make_uint4(h[0], h[1], l[0], l[1]);
    }
}
// acc[c][.] (n-tiles nt0 .. nt0 + 3) += circulant * line, A fragments ah / al of the whole line
template <int N>
__device__ __forceinline__ void lm_circ4(const uint32_t (&ah)[N / 16][4], const uint32_t (&al)[N / 16][4],
                                         const uint4* __restrict__ f, int lane, int nt0, float (&acc)[4][4]) {
#pragma unroll
    for (int ks = 0; ks < N / 16; ++ks) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int q = (nt0 + c - 2 * ks) & (N / 8 - 1);
            const uint4 b = f[q * 32 + lane];
            lm_mma(acc[c], al[ks], b.x, b.y);
            lm_mma(acc[c], ah[ks], b.z, b.w);
            lm_mma(acc[c], ah[ks], b.x, b.y);
        }
    }
}

// N: SMALL length.  CTA = LM_WARPS m-tiles of one 8-channel group; m-tile = lines (2 mt, 2 mt + 1) x 8 channels.
// Shared memory: up filter | down filter | per-warp fp32 line tile [N][16] (even samples kept for the down-sampler).
template <int N, int OP, int ACT>
__global__ void __launch_bounds__(32 * LM_WARPS)
line_mma_kernel(const LineArgs a) {
    pdl_trigger();
    extern __shared__ __align__(16) unsigned char lm_smem[];
    uint4* fu = reinterpret_cast<uint4*>(lm_smem);                       // [(N/8)][32]
    uint4* fd = fu + (N / 8) * 32;
    float* etile = reinterpret_cast<float*>(fd + (N / 8) * 32) + (threadIdx.x >> 5) * (N * LM_EP);   // [N pos][16 lines + pad]
    if constexpr (OP != OP_DOWN) lm_fill_filter<N, false>(fu, threadIdx.x, blockDim.x);
    if constexpr (OP != OP_UP) lm_fill_filter<N, true>(fd, threadIdx.x, blockDim.x);
    __syncthreads();
    pdl_wait();                                      // the filter tables above do not depend on the previous kernel
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int mt = blockIdx.y * LM_WARPS + warp;
    const int ch = blockIdx.x * 8 + g;
    if (2 * mt >= a.n_lines) return;                 // no block-wide barrier below
    constexpr int NT = N / 8, KS = N / 16;
    const float* xin[2];
    float* yout[2];
    __half* youth[2];                                // the same positions when y holds fp16 (a.y_half)
    bool ok[2];
    float sc[2] = {1.f, 1.f}, sh[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int line = min(2 * mt + r, a.n_lines - 1);
        ok[r] = 2 * mt + r < a.n_lines;
        const int outer = line / a.n_inner, inner = line - outer * a.n_inner;
        xin[r] = a.x + outer * a.in_outer + inner * a.in_inner + ch;
        yout[r] = a.y + outer * a.out_outer + inner * a.out_inner + ch;
        youth[r] = reinterpret_cast<__half*>(a.y) + outer * a.out_outer + inner * a.out_inner + ch;
        if (a.scale != nullptr) {
            const int b = line / a.lines_per_image;
            sc[r] = a.scale[(size_t)b * a.C + ch];
            sh[r] = a.shift[(size_t)b * a.C + ch];
        }
    }

    uint32_t ah[KS][4], al[KS][4];
    if constexpr (OP == OP_UP || OP == OP_UPACTDOWN) {
        // load the whole line first (positions 8 nt + 2 t + q): all loads in flight at once - the stores below may alias
        // the input as far as the compiler knows, so interleaving them serialises load -> store -> load
        float v[KS][4][2];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
            for (int idx = 0; idx < 4; ++idx) {
                const int r = idx & 1, j = 8 * (2 * ks + (idx >> 1)) + 2 * t;
                v[ks][idx][0] = xin[r][(size_t)j * a.in_pos];
                v[ks][idx][1] = xin[r][(size_t)(j + 1) * a.in_pos];
            }
        }
        // keep the even samples: in global memory (UP) or the warp's tile
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
            for (int idx = 0; idx < 4; ++idx) {
                const int r = idx & 1, nt = 2 * ks + (idx >> 1);
                const int j = 8 * nt + 2 * t;
                const float v0 = fmaf(v[ks][idx][0], sc[r], sh[r]);
                const float v1 = fmaf(v[ks][idx][1], sc[r], sh[r]);
                lm_split(v0, v1, ah[ks][idx], al[ks][idx]);
                if constexpr (OP == OP_UP) {
                    if (ok[r]) {
                        if (a.y_half) {
                            youth[r][(size_t)(2 * j) * a.out_pos] = __float2half_rn(apply_act<ACT>(v0));
                            youth[r][(size_t)(2 * j + 2) * a.out_pos] = __float2half_rn(apply_act<ACT>(v1));
                        } else {
                            yout[r][(size_t)(2 * j) * a.out_pos] = apply_act<ACT>(v0);
                            yout[r][(size_t)(2 * j + 2) * a.out_pos] = apply_act<ACT>(v1);
                        }
                    }
                } else {
                    etile[j * LM_EP + g + 8 * r] = apply_act<ACT>(v0);
                    etile[(j + 1) * LM_EP + g + 8 * r] = apply_act<ACT>(v1);
                }
            }
        }
        // odd phase, four n-tiles at a time
        uint32_t oh[KS][4], ol[KS][4];
#pragma unroll
        for (int nt0 = 0; nt0 < NT; nt0 += 4) {
            float acc[4][4];
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;
            lm_circ4<N>(ah, al, fu, lane, nt0, acc);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int nt = nt0 + c, j = 8 * nt + 2 * t;
                if constexpr (OP == OP_UP) {
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        if (ok[r]) {
                            if (a.y_half) {
                                youth[r][(size_t)(2 * j + 1) * a.out_pos] = __float2half_rn(apply_act<ACT>(acc[c][2 * r]));
                                youth[r][(size_t)(2 * j + 3) * a.out_pos] = __float2half_rn(apply_act<ACT>(acc[c][2 * r + 1]));
                            } else {
                                yout[r][(size_t)(2 * j + 1) * a.out_pos] = apply_act<ACT>(acc[c][2 * r]);
                                yout[r][(size_t)(2 * j + 3) * a.out_pos] = apply_act<ACT>(acc[c][2 * r + 1]);
                            }
                        }
                    }
                } else {
                    // activated odd samples become the A fragments of the down-sampler: n-tile nt = half (nt & 1) of k-step nt / 2
                    lm_split(apply_act<ACT>(acc[c][0]), apply_act<ACT>(acc[c][1]), oh[nt >> 1][2 * (nt & 1)], ol[nt >> 1][2 * (nt & 1)]);
                    lm_split(apply_act<ACT>(acc[c][2]), apply_act<ACT>(acc[c][3]), oh[nt >> 1][2 * (nt & 1) + 1], ol[nt >> 1][2 * (nt & 1) + 1]);
                }
            }
        }
        if constexpr (OP == OP_UP) return;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int idx = 0; idx < 4; ++idx) { ah[ks][idx] = oh[ks][idx]; al[ks][idx] = ol[ks][idx]; }
    } else {
        // OP_DOWN: the line has 2N samples; odd ones -> A fragments, even ones -> the warp's tile (two load batches)
        {
            float v[KS][4][2];
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                for (int idx = 0; idx < 4; ++idx) {
                    const int r = idx & 1, j = 8 * (2 * ks + (idx >> 1)) + 2 * t;
                    v[ks][idx][0] = xin[r][(size_t)(2 * j + 1) * a.in_pos];
                    v[ks][idx][1] = xin[r][(size_t)(2 * j + 3) * a.in_pos];
                }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                for (int idx = 0; idx < 4; ++idx) lm_split(v[ks][idx][0], v[ks][idx][1], ah[ks][idx], al[ks][idx]);
        }
        {
            float v[KS][4][2];
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                for (int idx = 0; idx < 4; ++idx) {
                    const int r = idx & 1, j = 8 * (2 * ks + (idx >> 1)) + 2 * t;
                    v[ks][idx][0] = xin[r][(size_t)(2 * j) * a.in_pos];
                    v[ks][idx][1] = xin[r][(size_t)(2 * j + 2) * a.in_pos];
                }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                for (int idx = 0; idx < 4; ++idx) {
                    const int r = idx & 1, j = 8 * (2 * ks + (idx >> 1)) + 2 * t;
                    etile[j * LM_EP + g + 8 * r] = v[ks][idx][0];
                    etile[(j + 1) * LM_EP + g + 8 * r] = v[ks][idx][1];
                }
        }
    }
    __syncwarp();
    // down: y[i] = e[i] / 2 - (-1)^i altsum(e) / 2N + sum_m G[i - m] o[m]
    float s[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        float acc0 = 0.f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int j = 8 * nt + 2 * t;
            acc0 += etile[j * LM_EP + g + 8 * r] - etile[(j + 1) * LM_EP + g + 8 * r];
        }
        acc0 += __shfl_xor_sync(0xffffffffu, acc0, 1);
        acc0 += __shfl_xor_sync(0xffffffffu, acc0, 2);
        s[r] = acc0 * (1.0f / (2 * N));
    }
#pragma unroll
    for (int nt0 = 0; nt0 < NT; nt0 += 4) {
        float acc[4][4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = 8 * (nt0 + c) + 2 * t;
            acc[c][0] = fmaf(0.5f, etile[j * LM_EP + g], -s[0]);
            acc[c][1] = fmaf(0.5f, etile[(j + 1) * LM_EP + g], s[0]);
            acc[c][2] = fmaf(0.5f, etile[j * LM_EP + g + 8], -s[1]);
            acc[c][3] = fmaf(0.5f, etile[(j + 1) * LM_EP + g + 8], s[1]);
        }
        lm_circ4<N>(ah, al, fd, lane, nt0, acc);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = 8 * (nt0 + c) + 2 * t;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                if (ok[r]) {
                    if (a.y_half) {
                        youth[r][(size_t)j * a.out_pos] = __float2half_rn(acc[c][2 * r]);
                        youth[r][(size_t)(j + 1) * a.out_pos] = __float2half_rn(acc[c][2 * r + 1]);
                    } else {
                        yout[r][(size_t)j * a.out_pos] = acc[c][2 * r];
                        yout[r][(size_t)(j + 1) * a.out_pos] = acc[c][2 * r + 1];
                    }
                }
            }
        }
    }
}

template <int N, int OP, int ACT>
int launch_lines_mma(const LineArgs& a, cudaStream_t st) {
    constexpr int smem = 2 * (N / 8) * 32 * 16 + LM_WARPS * N * LM_EP * 4;
    auto kern = line_mma_kernel<N, OP, ACT>;
    static std::atomic<unsigned long long> configured{0};      // one bit per device
    {
        cudaError_t e = set_max_dyn_smem(kern, smem, configured);
        if (e != cudaSuccess) return (int)e;
    }
    launch_k(kern, dim3(a.C / 8, ceil_div(ceil_div(a.n_lines, 2), LM_WARPS)), dim3(32 * LM_WARPS), smem, st, a);
    return 0;
}

bool large_mma_enabled() {
    static const bool on = !(getenv("AFLDM_FACT_MMA") && atoi(getenv("AFLDM_FACT_MMA")) == 0);
    return on;
}

// dispatch: tensor-core line pass (default) or the exact-FMA one
template <int N, int OP, int ACT>
int launch_pass(const LineArgs& a, cudaStream_t st) {
    if (large_mma_enabled() && a.C % 8 == 0) return launch_lines_mma<N, OP, ACT>(a, st);
    if (a.y_half) return AFLDM_E_NOKERNEL;          // fp16 stores live in the tensor-core passes
    return launch_lines<N, OP, ACT>(a, st);
}

template <int N>
int run_large(int mode, int act, const float* x, float* y, int B, int C, const float* scale, const float* shift,
              float* ws, cudaStream_t st, int y_half) {
    // mode: 0 filtered act, 1 up2, 2 down2 (the enum of resample.cu)
    const size_t plane2 = (size_t)B * N * 2 * N * C;   // [B][N][2N][C] intermediate
    int rc = 0, launches = 0;
    if (mode == 0) {
        float* t1 = ws;
        float* y1 = ws + plane2;
        LineArgs r = rows(x, t1, B, N, N, 2 * N, C);
        r.scale = scale; r.shift = shift;
        rc = launch_pass<N, OP_UP, AFLDM_ACT_IDENTITY>(r, st);
        if (rc) return rc;
        LineArgs cmid = cols(t1, y1, B, N, N, 2 * N, C);
        rc = (act == AFLDM_ACT_SILU) ? launch_pass<N, OP_UPACTDOWN, AFLDM_ACT_SILU>(cmid, st)
                                     : launch_pass<N, OP_UPACTDOWN, AFLDM_ACT_IDENTITY>(cmid, st);
        if (rc) return rc;
        LineArgs last = rows(y1, y, B, N, 2 * N, N, C);
        last.y_half = y_half;                         // only the pass that writes the result
        rc = launch_pass<N, OP_DOWN, AFLDM_ACT_IDENTITY>(last, st);
        launches = 3;
    } else if (mode == 1) {
        float* t1 = ws;
        LineArgs r = rows(x, t1, B, N, N, 2 * N, C);
        r.scale = scale; r.shift = shift;
        rc = launch_pass<N, OP_UP, AFLDM_ACT_IDENTITY>(r, st);
        if (rc) return rc;
        LineArgs last = cols(t1, y, B, N, 2 * N, 2 * N, C);
        last.y_half = y_half;
        rc = launch_pass<N, OP_UP, AFLDM_ACT_IDENTITY>(last, st);
        launches = 2;
    } else {
        float* y1 = ws;
        rc = launch_pass<N, OP_DOWN, AFLDM_ACT_IDENTITY>(cols(x, y1, B, 2 * N, N, 2 * N, C), st);
        if (rc) return rc;
        if (y_half) return AFLDM_E_NOKERNEL;
        rc = launch_pass<N, OP_DOWN, AFLDM_ACT_IDENTITY>(rows(y1, y, B, N, 2 * N, N, C), st);
        launches = 2;
    }
    if (rc) return rc;
    return launched(launches);
}

}  // namespace

size_t resample_large_workspace_floats(int mode, int B, int n, int C) {
    if (n != 64 && n != 128) return 0;
    const size_t plane2 = (size_t)B * n * 2 * n * C;
    return mode == 0 ? 2 * plane2 : plane2;
}

int resample_large(int mode, int act, const float* x, float* y, int B, int n, int C, const float* scale,
                   const float* shift, float* ws, size_t ws_floats, cudaStream_t st, int y_half) {
    if (n != 64 && n != 128) return AFLDM_E_NOKERNEL;
    if (y_half && !(large_mma_enabled() && C % 8 == 0)) return AFLDM_E_NOKERNEL;   // before any pass is launched
    if (C % 32 != 0) return AFLDM_E_SHAPE;
    if (ws == nullptr || ws_floats < resample_large_workspace_floats(mode, B, n, C)) return AFLDM_E_WORKSPACE;
    if (n == 64) return run_large<64>(mode, act, x, y, B, C, scale, shift, ws, st, y_half);
    return run_large<128>(mode, act, x, y, B, C, scale, shift, ws, st, y_half);
}

}  // namespace afldm
