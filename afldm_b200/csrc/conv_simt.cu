// fp32 SIMT implicit-GEMM convolution (k = 1 or 3, stride 1, "same" zero padding) on NHWC.
//
// Exact-fp32 class path (AFLDM_CONV_SIMT_F32): used for parity, for layers whose channel
// counts do not fit the tensor-core path (conv_in 4->C, conv_out C->4/3) and as the oracle-near
// reference for the tcgen05 TF32 path.  GEMM view: M = B*H*W output pixels, N = Cout,
// K = k*k*Cin with k index = tap*Cin + ci (weights packed [Cout][tap][Cin]).
//
// CTA tile 128 x 64, K step 16, 256 threads, 8 x 4 outputs per thread, register-staged double
// buffering of the global loads.  Small-M layers (the 4x4 / 2x2 levels of the UNet, where the
// 9*Cin*Cout*4 B of weights dominate) are split along K over blockIdx.z so that the weight
// stream is spread over all SMs; partial sums go to a caller workspace and a second kernel
// reduces them in a fixed order (deterministic) and applies the epilogue.
#include <algorithm>

#include "common.cuh"
#include "conv.cuh"

namespace afldm {

constexpr int CBM = 128, CBN = 64, CBK = 16;

ConvPlan conv_simt_plan(int B, int H, int W, int Cin, int Cout, int ks) {
    ConvPlan p;
    p.M = B * H * W;
    p.K = ks * ks * Cin;
    p.mtiles = ceil_div(p.M, CBM);
    p.ntiles = ceil_div(Cout, CBN);
    p.kchunks = ceil_div(p.K, CBK);
    const int tiles = p.mtiles * p.ntiles;
    int s = 1;
    if (tiles < 148) {
        s = ceil_div(2 * 148, tiles);
        s = std::min(s, std::max(1, p.kchunks / 8));  // keep >= 128 k per split
        s = std::min(s, 64);
    }
    p.chunks_per_split = ceil_div(p.kchunks, s);
    p.splitk = ceil_div(p.kchunks, p.chunks_per_split);
    return p;
}

namespace {

constexpr int LDA = CBM + 4, LDB = CBN + 4;

struct ConvArgs {
    const float* x;
    const float* w;
    const float* bias;
    const float* row_add;
    const float* residual;
    float* y;
    float* ws;
    int x_pitch, res_pitch, y_pitch, row_add_pitch;
    int B, H, W, Cin, Cout, ks, M, K, chunks_per_split, kchunks;
};

// FAST: Cin % 16 == 0, x_pitch % 4 == 0, x and w 16-byte aligned -> float4 loads, one tap per chunk.
template <bool FAST>
__global__ void __launch_bounds__(256, 2) conv_simt_kernel(const ConvArgs a) {
    pdl_trigger();
    pdl_wait();
    __shared__ __align__(16) float As[2][CBK][LDA];
    __shared__ __align__(16) float Bs[2][CBK][LDB];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * CBM, n0 = blockIdx.y * CBN;
    const int tx = tid & 15, ty = tid >> 4;
    const int pad = a.ks >> 1;
    const int HW = a.H * a.W;

    // loader coordinates: A rows (pixels) lr, lr + 64; quad kq of the 16-wide chunk
    const int lr = tid >> 2, kq = tid & 3;
    int pb[2], ph[2], pw[2];
    bool pv[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int m = m0 + lr + 64 * r;
        pv[r] = m < a.M;
        const int mm = pv[r] ? m : 0;
        pb[r] = mm / HW;
        const int rem = mm - pb[r] * HW;
        ph[r] = rem / a.W;
        pw[r] = rem - ph[r] * a.W;
    }
    const int bn = n0 + lr;  // B row (output channel) loaded by this thread
    const bool bv = bn < a.Cout;

    const int c_beg = blockIdx.z * a.chunks_per_split;
    const int c_end = min(a.kchunks, c_beg + a.chunks_per_split);

    float4 ra[2], rb;
    auto load_chunk = [&](int chunk) {
        const int k0 = chunk * CBK;
        if constexpr (FAST) {
            const int tap = k0 / a.Cin;
            const int ci = k0 - tap * a.Cin + kq * 4;
            const int dh = tap / a.ks - pad, dw = tap % a.ks - pad;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int hs = ph[r] + dh, wsrc = pw[r] + dw;
                const bool ok = pv[r] && hs >= 0 && hs < a.H && wsrc >= 0 && wsrc < a.W;
                ra[r] = ok ? *reinterpret_cast<const float4*>(
                                 a.x + ((size_t)(pb[r] * a.H + hs) * a.W + wsrc) * a.x_pitch + ci)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            rb = bv ? *reinterpret_cast<const float4*>(a.w + (size_t)bn * a.K + k0 + kq * 4)
                    : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            float tmp[2][4], tb[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k = k0 + kq * 4 + e;
                const bool kv = k < a.K;
                const int kk = kv ? k : 0;
                const int tap = kk / a.Cin, ci = kk - tap * a.Cin;
                const int dh = tap / a.ks - pad, dw = tap % a.ks - pad;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int hs = ph[r] + dh, wsrc = pw[r] + dw;
                    const bool ok = kv && pv[r] && hs >= 0 && hs < a.H && wsrc >= 0 && wsrc < a.W;
                    tmp[r][e] = ok ? a.x[((size_t)(pb[r] * a.H + hs) * a.W + wsrc) * a.x_pitch + ci] : 0.f;
                }
                tb[e] = (kv && bv) ? a.w[(size_t)bn * a.K + k] : 0.f;
            }
            ra[0] = make_float4(tmp[0][0], tmp[0][1], tmp[0][2], tmp[0][3]);
            ra[1] = make_float4(tmp[1][0], tmp[1][1], tmp[1][2], tmp[1][3]);
            rb = make_float4(tb[0], tb[1], tb[2], tb[3]);
        }
    };
    auto store_chunk = [&](int buf) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int m = lr + 64 * r;
            As[buf][kq * 4 + 0][m] = ra[r].x;
            As[buf][kq * 4 + 1][m] = ra[r].y;
            As[buf][kq * 4 + 2][m] = ra[r].z;
            As[buf][kq * 4 + 3][m] = ra[r].w;
        }
        Bs[buf][kq * 4 + 0][lr] = rb.x;
        Bs[buf][kq * 4 + 1][lr] = rb.y;
        Bs[buf][kq * 4 + 2][lr] = rb.z;
        Bs[buf][kq * 4 + 3][lr] = rb.w;
    };

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    if (c_beg < c_end) {
        load_chunk(c_beg);
        store_chunk(0);
    }
    __syncthreads();
    for (int chunk = c_beg; chunk < c_end; ++chunk) {
        const int buf = (chunk - c_beg) & 1;
        const bool more = chunk + 1 < c_end;
        if (more) load_chunk(chunk + 1);
#pragma unroll
        for (int kk = 0; kk < CBK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bw[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bw[j], acc[i][j]);
        }
        if (more) store_chunk(buf ^ 1);
        __syncthreads();
    }

    // epilogue
    const int n = n0 + tx * 4;
    const bool split = gridDim.z > 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= a.M) continue;
        if (split) {
            float* dst = a.ws + ((size_t)blockIdx.z * a.M + m) * a.Cout + n;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n + j < a.Cout) dst[j] = acc[i][j];
        } else {
            const int b = m / HW;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (n + j >= a.Cout) continue;
                float v = acc[i][j];
                if (a.bias != nullptr) v += a.bias[n + j];
                if (a.row_add != nullptr) v += a.row_add[(size_t)b * a.row_add_pitch + n + j];
                if (a.residual != nullptr) v += a.residual[(size_t)m * a.res_pitch + n + j];
                a.y[(size_t)m * a.y_pitch + n + j] = v;
            }
        }
    }
}

// Shared with the tensor-core path: y = sum_z ws[z] + bias + row_add + residual, summed in a fixed
// order.  One thread owns one output channel of one image and walks its HW rows (consecutive
// threads = consecutive channels: coalesced), so the per-(image, channel) GroupNorm partial sums
// (sum, sum of squares) of the finished output fall out for free (gn_partial [B][1][Cout] float2 | NULL).
constexpr int RED_COLS = 32, RED_ROWL = 8;     // CTA: 32 channels x 8 row lanes, RED_ROWS rows of one image

__global__ void __launch_bounds__(RED_COLS * RED_ROWL)
splitk_reduce_kernel(const float* __restrict__ ws, int splitk, const float* __restrict__ bias,
                     const float* __restrict__ row_add, int row_add_pitch, const float* residual,
                     int res_pitch, float* y, int y_pitch, int M, int Cout, int HW, float* gn_partial) {
    pdl_trigger();
    pdl_wait();
    __shared__ float2 red[RED_ROWL][RED_COLS];
    const int slot = blockIdx.z, slots = gridDim.z;
    const int col = threadIdx.x % RED_COLS, rl = threadIdx.x / RED_COLS;
    const int n = blockIdx.x * RED_COLS + col;
    const int b = blockIdx.y;
    const bool ok = n < Cout;
    const size_t total = (size_t)M * Cout;
    float s = 0.f, q = 0.f;
    if (ok) {
        const float add = (bias != nullptr ? bias[n] : 0.f) +
                          (row_add != nullptr ? row_add[(size_t)b * row_add_pitch + n] : 0.f);
        const int m_beg = b * HW + slot * SPLITK_REDUCE_ROWS;
        const int m_end = min(M, min((b + 1) * HW, m_beg + SPLITK_REDUCE_ROWS));
        for (int m = m_beg + rl; m < m_end; m += RED_ROWL) {
            const float* p = ws + (size_t)m * Cout + n;
            float v = 0.f;
            int z = 0;
            for (; z + 4 <= splitk; z += 4) {   // 4 loads in flight, summed in z order
                const float v0 = p[(size_t)z * total], v1 = p[(size_t)(z + 1) * total];
                const float v2 = p[(size_t)(z + 2) * total], v3 = p[(size_t)(z + 3) * total];
                v = (((v + v0) + v1) + v2) + v3;
            }
            for (; z < splitk; ++z) v += p[(size_t)z * total];
            v += add;
            if (residual != nullptr) v += residual[(size_t)m * res_pitch + n];
            y[(size_t)m * y_pitch + n] = v;
            s += v;
            q = fmaf(v, v, q);
        }
    }
    if (gn_partial == nullptr) return;
    red[rl][col] = make_float2(s, q);
    __syncthreads();
    if (rl == 0 && ok) {
        float ts = 0.f, tq = 0.f;
#pragma unroll
        for (int r = 0; r < RED_ROWL; ++r) {    // fixed order
            ts += red[r][col].x;
            tq += red[r][col].y;
        }
        reinterpret_cast<float2*>(gn_partial)[((size_t)b * slots + slot) * Cout + n] = make_float2(ts, tq);
    }
}

}  // namespace

void splitk_reduce_launch(const float* ws, int splitk, const float* bias, const float* row_add,
                          int row_add_pitch, const float* residual, int res_pitch, float* y, int y_pitch, int M,
                          int Cout, int HW, float* gn_partial, cudaStream_t st) {
    const int B = ceil_div(M, HW);
    launch_k(splitk_reduce_kernel, dim3(ceil_div(Cout, RED_COLS), B, splitk_reduce_slots(HW)),
             dim3(RED_COLS * RED_ROWL), 0, st, ws, splitk, bias, row_add,
             row_add_pitch, residual, res_pitch, y, y_pitch, M, Cout, HW, gn_partial);
}

int conv_simt_launch(const float* x, int x_pitch, const float* w, const float* bias,
                     const float* row_add, int row_add_pitch, const float* residual, int res_pitch, float* y,
                     int y_pitch, int B, int H, int W, int Cin, int Cout, int ks, float* workspace,
                     size_t workspace_floats, cudaStream_t st) {
    const ConvPlan p = conv_simt_plan(B, H, W, Cin, Cout, ks);
    if (p.splitk > 1 && (workspace == nullptr || workspace_floats < (size_t)p.splitk * p.M * Cout))
        return AFLDM_E_WORKSPACE;
    ConvArgs a;
    a.x = x; a.w = w; a.bias = bias; a.row_add = row_add; a.residual = residual; a.y = y;
    a.ws = workspace;
    a.x_pitch = x_pitch; a.res_pitch = res_pitch; a.y_pitch = y_pitch; a.row_add_pitch = row_add_pitch;
    a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.ks = ks; a.M = p.M; a.K = p.K;
    a.chunks_per_split = p.chunks_per_split; a.kchunks = p.kchunks;
    const bool fast = (Cin % 16 == 0) && (x_pitch % 4 == 0) && aligned16(x) && aligned16(w);
    dim3 grid(p.mtiles, p.ntiles, p.splitk);
    if (fast)
        launch_k(conv_simt_kernel<true>, dim3(grid), dim3(256), 0, st, a);
    else
        launch_k(conv_simt_kernel<false>, dim3(grid), dim3(256), 0, st, a);
    int launches = 1;
    if (p.splitk > 1) {
        splitk_reduce_launch(workspace, p.splitk, bias, row_add, row_add_pitch, residual, res_pitch, y, y_pitch, p.M, Cout,
                             H * W, nullptr, st);
        ++launches;
    }
    return launched(launches);
}

}  // namespace afldm
