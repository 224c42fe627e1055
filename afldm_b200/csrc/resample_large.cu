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
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
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

template <int N>
int run_large(int mode, int act, const float* x, float* y, int B, int C, const float* scale, const float* shift,
              float* ws, cudaStream_t st) {
    // mode: 0 filtered act, 1 up2, 2 down2 (the enum of resample.cu)
    const size_t plane2 = (size_t)B * N * 2 * N * C;   // [B][N][2N][C] intermediate
    int rc = 0, launches = 0;
    if (mode == 0) {
        float* t1 = ws;
        float* y1 = ws + plane2;
        LineArgs r = rows(x, t1, B, N, N, 2 * N, C);
        r.scale = scale; r.shift = shift;
        rc = launch_lines<N, OP_UP, AFLDM_ACT_IDENTITY>(r, st);
        if (rc) return rc;
        LineArgs cmid = cols(t1, y1, B, N, N, 2 * N, C);
        rc = (act == AFLDM_ACT_SILU) ? launch_lines<N, OP_UPACTDOWN, AFLDM_ACT_SILU>(cmid, st)
                                     : launch_lines<N, OP_UPACTDOWN, AFLDM_ACT_IDENTITY>(cmid, st);
        if (rc) return rc;
        rc = launch_lines<N, OP_DOWN, AFLDM_ACT_IDENTITY>(rows(y1, y, B, N, 2 * N, N, C), st);
        launches = 3;
    } else if (mode == 1) {
        float* t1 = ws;
        LineArgs r = rows(x, t1, B, N, N, 2 * N, C);
        r.scale = scale; r.shift = shift;
        rc = launch_lines<N, OP_UP, AFLDM_ACT_IDENTITY>(r, st);
        if (rc) return rc;
        rc = launch_lines<N, OP_UP, AFLDM_ACT_IDENTITY>(cols(t1, y, B, N, 2 * N, 2 * N, C), st);
        launches = 2;
    } else {
        float* y1 = ws;
        rc = launch_lines<N, OP_DOWN, AFLDM_ACT_IDENTITY>(cols(x, y1, B, 2 * N, N, 2 * N, C), st);
        if (rc) return rc;
        rc = launch_lines<N, OP_DOWN, AFLDM_ACT_IDENTITY>(rows(y1, y, B, N, 2 * N, N, C), st);
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
                   const float* shift, float* ws, size_t ws_floats, cudaStream_t st) {
    if (n != 64 && n != 128) return AFLDM_E_NOKERNEL;
    if (C % 32 != 0) return AFLDM_E_SHAPE;
    if (ws == nullptr || ws_floats < resample_large_workspace_floats(mode, B, n, C)) return AFLDM_E_WORKSPACE;
    if (n == 64) return run_large<64>(mode, act, x, y, B, C, scale, shift, ws, st);
    return run_large<128>(mode, act, x, y, B, C, scale, shift, ws, st);
}

}  // namespace afldm
