// fp32 scaled-dot-product attention (flash-style online softmax), separate K/V source.
//
// Replaces F.scaled_dot_product_attention inside diffusers AttnProcessor2_0 as used by the
// reference including the cross-frame case of
// afldm/pipelines/cross_frame_attn.py:79-97, where K/V come from a stored reference-frame map
// with a smaller batch (batch b reads K/V batch b / (B / Bkv)).
//
// One thread owns one query row: q, the output accumulator and the running max / sum live in
// registers.  K and V tiles of 64 keys are staged in shared memory and read as warp-wide
// broadcasts (all threads of a CTA walk the keys in lock-step), so shared-memory traffic is
// one 128-bit broadcast per 4 FMAs.  Scores are kept in the exp2 domain (q is pre-multiplied by
// scale * log2 e); the running max is updated once per 8 keys.
#include <cuda_fp16.h>
#include <math.h>

#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace afldm {
namespace {

constexpr int ATT_KT = 64;   // keys per shared-memory tile
constexpr int ATT_KC = 8;    // keys per online-softmax update
constexpr int ATT_THREADS = 128;

template <int D>
__global__ void __launch_bounds__(ATT_THREADS)
attention_kernel(const float* __restrict__ q, int q_pitch, const float* __restrict__ k,
                 const float* __restrict__ v, int kv_pitch, float* __restrict__ o, int o_pitch,
                 int Bkv_rep, int Nq, int Nk, float qscale) {
    pdl_trigger();
    pdl_wait();
    constexpr int D4 = D / 4;
    __shared__ float4 Ks[ATT_KT * D4];
    __shared__ float4 Vs[ATT_KT * D4];
    const int b = blockIdx.z, head = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = r < Nq;
    const int bkv = b / Bkv_rep;

    float qr[D], acc[D];
    {
        const float4* qp = reinterpret_cast<const float4*>(q + ((size_t)b * Nq + (valid ? r : 0)) * q_pitch + head * D);
#pragma unroll
        for (int t = 0; t < D4; ++t) {
            const float4 t4 = qp[t];
            qr[4 * t + 0] = t4.x * qscale;
            qr[4 * t + 1] = t4.y * qscale;
            qr[4 * t + 2] = t4.z * qscale;
            qr[4 * t + 3] = t4.w * qscale;
        }
    }
#pragma unroll
    for (int t = 0; t < D; ++t) acc[t] = 0.f;
    float mrun = -INFINITY, lrun = 0.f;

    const float* kbase = k + ((size_t)bkv * Nk) * kv_pitch + head * D;
    const float* vbase = v + ((size_t)bkv * Nk) * kv_pitch + head * D;

    for (int k0 = 0; k0 < Nk; k0 += ATT_KT) {
        const int nk = min(ATT_KT, Nk - k0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < nk * D4; idx += blockDim.x) {
            const int key = idx / D4, t = idx - key * D4;
            Ks[idx] = *reinterpret_cast<const float4*>(kbase + (size_t)(k0 + key) * kv_pitch + 4 * t);
            Vs[idx] = *reinterpret_cast<const float4*>(vbase + (size_t)(k0 + key) * kv_pitch + 4 * t);
        }
        __syncthreads();
        for (int kk = 0; kk < nk; kk += ATT_KC) {
            float s[ATT_KC];
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < ATT_KC; ++j) {
                float d = 0.f;
                if (kk + j < nk) {
                    const float4* kr = Ks + (kk + j) * D4;
#pragma unroll
                    for (int t = 0; t < D4; ++t) {
                        const float4 kv4 = kr[t];
                        d = fmaf(qr[4 * t + 0], kv4.x, d);
                        d = fmaf(qr[4 * t + 1], kv4.y, d);
                        d = fmaf(qr[4 * t + 2], kv4.z, d);
                        d = fmaf(qr[4 * t + 3], kv4.w, d);
                    }
                } else {
                    d = -INFINITY;
                }
                s[j] = d;
                mx = fmaxf(mx, d);
            }
            const float mnew = fmaxf(mrun, mx);
            const float corr = exp2f(mrun - mnew);
            lrun *= corr;
#pragma unroll
            for (int t = 0; t < D; ++t) acc[t] *= corr;
#pragma unroll
            for (int j = 0; j < ATT_KC; ++j) {
                if (kk + j < nk) {
                    const float p = exp2f(s[j] - mnew);
                    lrun += p;
                    const float4* vr = Vs + (kk + j) * D4;
#pragma unroll
                    for (int t = 0; t < D4; ++t) {
                        const float4 vv = vr[t];
                        acc[4 * t + 0] = fmaf(p, vv.x, acc[4 * t + 0]);
                        acc[4 * t + 1] = fmaf(p, vv.y, acc[4 * t + 1]);
                        acc[4 * t + 2] = fmaf(p, vv.z, acc[4 * t + 2]);
                        acc[4 * t + 3] = fmaf(p, vv.w, acc[4 * t + 3]);
                    }
                }
            }
            mrun = mnew;
        }
    }
    if (valid) {
        const float inv = 1.0f / lrun;
        float4* op = reinterpret_cast<float4*>(o + ((size_t)b * Nq + r) * o_pitch + head * D);
#pragma unroll
        for (int t = 0; t < D4; ++t)
            op[t] = make_float4(acc[4 * t] * inv, acc[4 * t + 1] * inv, acc[4 * t + 2] * inv, acc[4 * t + 3] * inv);
    }
}

// x[r][:] = softmax(scale * x[r][:]); one warp per row, row kept in registers when cols <= 1024.
__global__ void __launch_bounds__(256)
softmax_rows_kernel(float* __restrict__ x, long long rows, int cols, int pitch, float scale_log2e) {
    pdl_trigger();
    pdl_wait();
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    float* xr = x + row * pitch;
    float mx = -INFINITY;
    for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, xr[c] * scale_log2e);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int c = lane; c < cols; c += 32) {
        const float p = exp2f(xr[c] * scale_log2e - mx);
        xr[c] = p;
        sum += p;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int c = lane; c < cols; c += 32) xr[c] *= inv;
}

// ------------------------------------------------------------------------------------------
// Tensor-core variant (AFLDM_ATTN_MMA_TF32): flash attention on warp-level mma.m16n8k8 TF32 with
// fp32 accumulation and fp32 online softmax.  One warp owns 16 query rows; a CTA of 4 warps shares
// 64-key K / V tiles in shared memory (row pitch D + 4 words: every fragment load is conflict-free).
// The S accumulator fragment (cols 2t, 2t+1 per lane) is reused directly as the A fragment of P.V by
// reading V rows in the matching permuted order (k = t <- key 2t, k = t + 4 <- key 2t + 1), so no
// shuffles or shared-memory round trip are needed between the two products.
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float ex2_approx(float x) {      // 2^x, one MUFU; -inf -> 0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// MT = 16-row query tiles per warp: 2 for long sequences (every K / V fragment read from shared memory feeds
// two MMAs; ncu on the MT = 1 version showed ~700 issued instructions per 48 HMMA, profiles/r01_ncu_attention.md),
// 1 for short ones (more CTAs).  K fragments are read as 64-bit words: MMA k-index t <- head-dim column 2t,
// t + 4 <- 2t + 1 inside each 8-column chunk, the Q fragment uses the same permutation.
// Head dims above 64 (SD-1.5: 80 and 160) keep Q and O fragments of D / 8 k-steps in registers: one or two CTAs per SM.
template <int D, int MT>
__global__ void __launch_bounds__(128, D > 96 ? 1 : (D > 64 ? 2 : (MT == 2 ? 3 : 5)))
attention_mma_kernel(const float* __restrict__ q, int q_pitch, const float* __restrict__ k,
                     const float* __restrict__ v, int kv_pitch, float* __restrict__ o, int o_pitch,
                     int Bkv_rep, int Nq, int Nk, float qscale) {
    pdl_trigger();
    pdl_wait();
    constexpr int KT = 64;            // keys per tile
    constexpr int PK = (D % 32 == 8) ? D : ((D / 32) * 32 + 40);   // K row pitch (words), = 8 mod 32: conflict-free 64-bit fragment loads
    constexpr int P = D + 4;          // V row pitch (words): conflict-free 32-bit fragment loads
    constexpr int DK = D / 8;         // k-steps of Q.K^T == n-tiles of P.V
    constexpr int D4 = D / 4;
    // double-buffered K / V tiles: tile i+1 streams in with cp.async while tile i is consumed
    extern __shared__ __align__(16) float att_smem[];          // [K0 | K1 | V0 | V1]
    auto Kbuf = [&](int i) { return att_smem + i * (KT * PK); };
    auto Vbuf = [&](int i) { return att_smem + 2 * (KT * PK) + i * (KT * P); };
    const int b = blockIdx.z, head = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int bkv = b / Bkv_rep;
    const int wrow = blockIdx.x * (64 * MT) + warp * (16 * MT);     // first query row of this warp

    const float* kbase = k + ((size_t)bkv * Nk) * kv_pitch + head * D;
    const float* vbase = v + ((size_t)bkv * Nk) * kv_pitch + head * D;
    auto load_tile = [&](int buf, int k0) {
        const int nk = min(KT, Nk - k0);
        float* kb = Kbuf(buf);
        float* vb = Vbuf(buf);
#pragma unroll
        for (int it = 0; it < (KT * D4 + 127) / 128; ++it) {
            const int idx = it * 128 + threadIdx.x;
            if (idx >= KT * D4) break;
            const int key = idx / D4, c4 = idx - key * D4;
            float* kd = kb + key * PK + 4 * c4;
            float* vd = vb + key * P + 4 * c4;
            if (key < nk) {
                const uint32_t ks_ = (uint32_t)__cvta_generic_to_shared(kd), vs_ = (uint32_t)__cvta_generic_to_shared(vd);
                const size_t goff = (size_t)(k0 + key) * kv_pitch + 4 * c4;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ks_), "l"(kbase + goff) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(vs_), "l"(vbase + goff) : "memory");
            } else {
                *reinterpret_cast<float4*>(kd) = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(vd) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    load_tile(0, 0);

    // Q fragments (scaled into the exp2 domain, rounded to TF32 once); permuted k order (see above)
    uint32_t qa[MT][DK][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int r0 = wrow + mt * 16 + g, r1 = r0 + 8;
        const float* q0 = q + ((size_t)b * Nq + min(r0, Nq - 1)) * q_pitch + head * D + 2 * t;
        const float* q1 = q + ((size_t)b * Nq + min(r1, Nq - 1)) * q_pitch + head * D + 2 * t;
#pragma unroll
        for (int ks = 0; ks < DK; ++ks) {
            const float2 a0 = *reinterpret_cast<const float2*>(q0 + 8 * ks);
            const float2 a1 = *reinterpret_cast<const float2*>(q1 + 8 * ks);
            qa[mt][ks][0] = to_tf32(a0.x * qscale);
            qa[mt][ks][1] = to_tf32(a1.x * qscale);
            qa[mt][ks][2] = to_tf32(a0.y * qscale);
            qa[mt][ks][3] = to_tf32(a1.y * qscale);
        }
    }
    float oacc[MT][DK][4];
    float mrun[MT][2], lrun[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int i = 0; i < DK; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) oacc[mt][i][j] = 0.f;
        mrun[mt][0] = mrun[mt][1] = -INFINITY;
        lrun[mt][0] = lrun[mt][1] = 0.f;
    }

    int buf = 0;
    for (int k0 = 0; k0 < Nk; k0 += KT, buf ^= 1) {
        const int nk = min(KT, Nk - k0);
        if (k0 + KT < Nk) {
            load_tile(buf ^ 1, k0 + KT);                       // prefetch the next tile
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const float* Kt = Kbuf(buf);
        const float* Vt = Vbuf(buf);

        // The tile body exists twice: the hot copy for full tiles carries no masking code at all (as one body the
        // compiler if-converts the tail mask into 64 always-executed selects per 32 x 64 scores).
        auto tile_body = [&](auto masked) {
            // S = Q K^T for 64 keys: 8 n-tiles of 8 keys
            float s[MT][8][4];
    #pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
    #pragma unroll
                for (int mt = 0; mt < MT; ++mt) s[mt][nt][0] = s[mt][nt][1] = s[mt][nt][2] = s[mt][nt][3] = 0.f;
                const float* kr = &Kt[(nt * 8 + g) * PK + 2 * t];
    #pragma unroll
                for (int ks = 0; ks < DK; ++ks) {
                    const float2 kb = *reinterpret_cast<const float2*>(kr + 8 * ks);
    #pragma unroll
                    for (int mt = 0; mt < MT; ++mt) mma_tf32(s[mt][nt], qa[mt][ks], __float_as_uint(kb.x), __float_as_uint(kb.y));
                }
            }
            if constexpr (decltype(masked)::value) {    // only the last tile of a ragged sequence has keys to mask
    #pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    const int key = nt * 8 + 2 * t;
    #pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        if (key >= nk) { s[mt][nt][0] = -INFINITY; s[mt][nt][2] = -INFINITY; }
                        if (key + 1 >= nk) { s[mt][nt][1] = -INFINITY; s[mt][nt][3] = -INFINITY; }
                    }
                }
            }
            float nm[MT][2];
    #pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                float mx0 = -INFINITY, mx1 = -INFINITY;
    #pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    mx0 = fmaxf(mx0, fmaxf(s[mt][nt][0], s[mt][nt][1]));
                    mx1 = fmaxf(mx1, fmaxf(s[mt][nt][2], s[mt][nt][3]));
                }
                mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
                mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
                mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
                mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
                const float n0 = fmaxf(mrun[mt][0], mx0), n1 = fmaxf(mrun[mt][1], mx1);   // finite: key k0 is always valid
                const float c0 = ex2_approx(mrun[mt][0] - n0), c1 = ex2_approx(mrun[mt][1] - n1);
                mrun[mt][0] = n0; mrun[mt][1] = n1;
                nm[mt][0] = n0; nm[mt][1] = n1;
                lrun[mt][0] *= c0; lrun[mt][1] *= c1;
    #pragma unroll
                for (int dn = 0; dn < DK; ++dn) {
                    oacc[mt][dn][0] *= c0; oacc[mt][dn][1] *= c0;
                    oacc[mt][dn][2] *= c1; oacc[mt][dn][3] *= c1;
                }
            }
            // P = exp2(S - m); O += P V
    #pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                uint32_t pa[MT][4];
    #pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const float p00 = ex2_approx(s[mt][nt][0] - nm[mt][0]), p01 = ex2_approx(s[mt][nt][1] - nm[mt][0]);
                    const float p10 = ex2_approx(s[mt][nt][2] - nm[mt][1]), p11 = ex2_approx(s[mt][nt][3] - nm[mt][1]);
                    lrun[mt][0] += p00 + p01;
                    lrun[mt][1] += p10 + p11;
                    // A fragment: k = t <- key 2t (c0 / c2), k = t + 4 <- key 2t + 1 (c1 / c3)
                    // p in [0, 1]: round-to-nearest TF32 is one integer add of half an ulp of the 10-bit mantissa (the MMA
                    // ignores the low 13 bits); cvt.rna.tf32 costs three instructions with its NaN / range handling
                    pa[mt][0] = __float_as_uint(p00) + 0x1000u; pa[mt][1] = __float_as_uint(p10) + 0x1000u;
                    pa[mt][2] = __float_as_uint(p01) + 0x1000u; pa[mt][3] = __float_as_uint(p11) + 0x1000u;
                }
                const float* vr0 = &Vt[(nt * 8 + 2 * t) * P + g];
                const float* vr1 = vr0 + P;
    #pragma unroll
                for (int dn = 0; dn < DK; ++dn) {
                    const uint32_t b0 = __float_as_uint(vr0[8 * dn]), b1 = __float_as_uint(vr1[8 * dn]);
    #pragma unroll
                    for (int mt = 0; mt < MT; ++mt) mma_tf32(oacc[mt][dn], pa[mt], b0, b1);
                }
            }
        };
        if (nk < KT) tile_body(std::true_type{});
        else tile_body(std::false_type{});
        __syncthreads();                                       // everyone is done with `buf` before it is refilled
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        float l0 = lrun[mt][0], l1 = lrun[mt][1];
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float i0 = 1.0f / l0, i1 = 1.0f / l1;
        const int r0 = wrow + mt * 16 + g, r1 = r0 + 8;
        if (r0 < Nq) {
            float* op = o + ((size_t)b * Nq + r0) * o_pitch + head * D + 2 * t;
#pragma unroll
            for (int dn = 0; dn < DK; ++dn)
                *reinterpret_cast<float2*>(op + 8 * dn) = make_float2(oacc[mt][dn][0] * i0, oacc[mt][dn][1] * i0);
        }
        if (r1 < Nq) {
            float* op = o + ((size_t)b * Nq + r1) * o_pitch + head * D + 2 * t;
#pragma unroll
            for (int dn = 0; dn < DK; ++dn)
                *reinterpret_cast<float2*>(op + 8 * dn) = make_float2(oacc[mt][dn][2] * i1, oacc[mt][dn][3] * i1);
        }
    }
}

template <int D, int MT>
int launch_mma_mt(const float* q, int q_pitch, const float* k, const float* v, int kv_pitch, float* o, int o_pitch,
                  int B, int Bkv, int Nq, int Nk, int heads, cudaStream_t st) {
    const float qscale = (float)((1.0 / sqrt((double)D)) * 1.4426950408889634);
    constexpr int PK = (D % 32 == 8) ? D : ((D / 32) * 32 + 40);
    constexpr int smem = 2 * 64 * (PK + D + 4) * 4;
    if (smem > 48 * 1024) {         // per (function, device): set on every launch (a host-side no-op after the first)
        cudaError_t e = cudaFuncSetAttribute(attention_mma_kernel<D, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
    }
    launch_k(attention_mma_kernel<D, MT>, dim3(ceil_div(Nq, 64 * MT), heads, B), dim3(128), smem, st,
        q, q_pitch, k, v, kv_pitch, o, o_pitch, B / Bkv, Nq, Nk, qscale);
    return launched();
}

template <int D>
int launch_mma(const float* q, int q_pitch, const float* k, const float* v, int kv_pitch, float* o, int o_pitch,
               int B, int Bkv, int Nq, int Nk, int heads, cudaStream_t st) {
    // two query tiles per warp once there are enough CTAs left to fill the chip twice over
    static const int force_mt = getenv("AFLDM_ATTN_MT") ? atoi(getenv("AFLDM_ATTN_MT")) : 0;
    if (force_mt == 1) return launch_mma_mt<D, 1>(q, q_pitch, k, v, kv_pitch, o, o_pitch, B, Bkv, Nq, Nk, heads, st);
    if constexpr (D <= 32) {
        if (Nq >= 256 && (long long)ceil_div(Nq, 128) * heads * B >= 2 * 148)
            return launch_mma_mt<D, 2>(q, q_pitch, k, v, kv_pitch, o, o_pitch, B, Bkv, Nq, Nk, heads, st);
    }
    return launch_mma_mt<D, 1>(q, q_pitch, k, v, kv_pitch, o, o_pitch, B, Bkv, Nq, Nk, heads, st);
}

// ------------------------------------------------------------------------------------------
// fp16-operand variant (afldm_attention_f16): q, k, v arrive as fp16 (the QKV projection's epilogue stores them
// that way), products run on mma.m16n8k16 with fp32 accumulation, softmax in fp32.  fp16 carries the same 11
// significant bits as the TF32 operands of the kernel above (narrower exponent: |q|, |k|, |v| < 65504), so this
// is the same numeric class at half the operand bytes and half the tensor-pipe time, and - what matters, the
// TF32 kernel is issue-bound (~750 instructions per 96 HMMA) - far fewer instructions: K and V fragments come
// from ldmatrix (.trans for V) instead of scalar LDS, P is packed two scores per F2FP, the softmax scale is
// folded into the exponent FFMA.
__device__ __forceinline__ void mma_f16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t (&r)[2]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t addr, uint32_t (&r)[2]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// D: head dim (multiple of 8, <= 64); MT: 16-row query tiles per warp.  4 warps, 64 * MT queries per CTA.
template <int D, int MT>
__global__ void __launch_bounds__(128, MT == 2 ? 4 : 5)
attention_f16_kernel(const __half* __restrict__ q, int q_pitch, const __half* __restrict__ k,
                     const __half* __restrict__ v, int kv_pitch, float* __restrict__ o, int o_pitch,
                     int Bkv_rep, int Nq, int Nk, float scale_log2e, int o_half) {
    pdl_trigger();
    constexpr int KT = 64;                       // keys per tile
    constexpr int DP = (D + 15) / 16 * 16;       // head dim padded to the MMA k (zero columns)
    constexpr int KS = DP / 16;                  // k-steps of Q.K^T
    constexpr int DN = D / 8;                    // n-tiles of P.V
    constexpr int PH = DP + 8;                   // smem row pitch in halves: rows 16 B apart mod 128 B -> conflict-free ldmatrix
    constexpr int CH = D / 8;                    // 16-byte chunks per row
    extern __shared__ __align__(16) __half att_h[];             // [K0 | K1 | V0 | V1], each KT * PH halves
    auto Kbuf = [&](int i) { return att_h + i * (KT * PH); };
    auto Vbuf = [&](int i) { return att_h + (2 + i) * (KT * PH); };
    const int b = blockIdx.z, head = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int bkv = b / Bkv_rep;
    const int wrow = blockIdx.x * (64 * MT) + warp * (16 * MT);

    // zero the pad columns [D, PH) of all four buffers once (cp.async only ever writes columns < D)
    for (int idx = threadIdx.x; idx < 4 * KT * (PH - D) / 8; idx += blockDim.x) {
        const int row = idx / ((PH - D) / 8), c8 = idx - row * ((PH - D) / 8);
        *reinterpret_cast<uint4*>(att_h + row * PH + D + 8 * c8) = make_uint4(0u, 0u, 0u, 0u);
    }
    pdl_wait();                                  // q / k / v come from the previous kernel (QKV projection)

    const __half* kbase = k + ((size_t)bkv * Nk) * kv_pitch + head * D;
    const __half* vbase = v + ((size_t)bkv * Nk) * kv_pitch + head * D;
    auto load_tile = [&](int buf, int k0) {
        const int nk = min(KT, Nk - k0);
        __half* kb = Kbuf(buf);
        __half* vb = Vbuf(buf);
#pragma unroll
        for (int it = 0; it < (KT * CH + 127) / 128; ++it) {
            const int idx = it * 128 + threadIdx.x;
            if (idx >= KT * CH) break;
            const int key = idx / CH, c8 = idx - key * CH;
            __half* kd = kb + key * PH + 8 * c8;
            __half* vd = vb + key * PH + 8 * c8;
            if (key < nk) {
                const uint32_t ks_ = (uint32_t)__cvta_generic_to_shared(kd), vs_ = (uint32_t)__cvta_generic_to_shared(vd);
                const size_t goff = (size_t)(k0 + key) * kv_pitch + 8 * c8;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ks_), "l"(kbase + goff) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(vs_), "l"(vbase + goff) : "memory");
            } else {
                *reinterpret_cast<uint4*>(kd) = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(vd) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    load_tile(0, 0);

    // Q fragments: a0 = (row g, k 2t..2t+1), a1 = (row g+8, same), a2 / a3 = k + 8; columns >= D are zero
    uint32_t qa[MT][KS][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const int r0 = wrow + mt * 16 + g, r1 = r0 + 8;
        const __half* q0 = q + ((size_t)b * Nq + min(r0, Nq - 1)) * q_pitch + head * D;
        const __half* q1 = q + ((size_t)b * Nq + min(r1, Nq - 1)) * q_pitch + head * D;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int ca = 16 * ks + 2 * t, cb = ca + 8;
            qa[mt][ks][0] = ca < D ? *reinterpret_cast<const uint32_t*>(q0 + ca) : 0u;
            qa[mt][ks][1] = ca < D ? *reinterpret_cast<const uint32_t*>(q1 + ca) : 0u;
            qa[mt][ks][2] = cb < D ? *reinterpret_cast<const uint32_t*>(q0 + cb) : 0u;
            qa[mt][ks][3] = cb < D ? *reinterpret_cast<const uint32_t*>(q1 + cb) : 0u;
        }
    }
    float oacc[MT][DN][4];
    float mrun[MT][2], lrun[MT][2];      // running max of the RAW scores, running sum of exp2((s - m) * scale_log2e)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int i = 0; i < DN; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) oacc[mt][i][j] = 0.f;
        mrun[mt][0] = mrun[mt][1] = -INFINITY;
        lrun[mt][0] = lrun[mt][1] = 0.f;
    }
    // ldmatrix lane roles: matrix m = lane / 8, row r = lane % 8
    const int lm = lane >> 3, lr = lane & 7;

    int buf = 0;
    for (int k0 = 0; k0 < Nk; k0 += KT, buf ^= 1) {
        const int nk = min(KT, Nk - k0);
        if (k0 + KT < Nk) {
            load_tile(buf ^ 1, k0 + KT);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const uint32_t Kt = (uint32_t)__cvta_generic_to_shared(Kbuf(buf));
        const uint32_t Vt = (uint32_t)__cvta_generic_to_shared(Vbuf(buf));

        auto tile_body = [&](auto masked) {
            // S = Q K^T: B fragment of (n-tile nt, k-step ks) = two 8x8 matrices: keys nt*8.., columns 16 ks + {0, 8}
            float s[MT][8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) s[mt][nt][0] = s[mt][nt][1] = s[mt][nt][2] = s[mt][nt][3] = 0.f;
                if constexpr (KS % 2 == 0) {
#pragma unroll
                    for (int ks = 0; ks < KS; ks += 2) {
                        uint32_t kb[4];
                        ldsm_x4(Kt + (uint32_t)(((nt * 8 + lr) * PH + 16 * ks + 8 * lm) * 2), kb);
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                            mma_f16_16816(s[mt][nt], qa[mt][ks], kb[0], kb[1]);
                            mma_f16_16816(s[mt][nt], qa[mt][ks + 1], kb[2], kb[3]);
                        }
                    }
                } else {
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                        uint32_t kb[2];
                        ldsm_x2(Kt + (uint32_t)(((nt * 8 + lr) * PH + 16 * ks + 8 * (lm & 1)) * 2), kb);
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) mma_f16_16816(s[mt][nt], qa[mt][ks], kb[0], kb[1]);
                    }
                }
            }
            if constexpr (decltype(masked)::value) {
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    const int key = nt * 8 + 2 * t;
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        if (key >= nk) { s[mt][nt][0] = -INFINITY; s[mt][nt][2] = -INFINITY; }
                        if (key + 1 >= nk) { s[mt][nt][1] = -INFINITY; s[mt][nt][3] = -INFINITY; }
                    }
                }
            }
            float nms[MT][2];          // new max * scale_log2e
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    mx0 = fmaxf(mx0, fmaxf(s[mt][nt][0], s[mt][nt][1]));
                    mx1 = fmaxf(mx1, fmaxf(s[mt][nt][2], s[mt][nt][3]));
                }
                mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
                mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
                mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
                mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
                const float n0 = fmaxf(mrun[mt][0], mx0), n1 = fmaxf(mrun[mt][1], mx1);   // finite: key k0 is valid
                const float c0 = ex2_approx((mrun[mt][0] - n0) * scale_log2e), c1 = ex2_approx((mrun[mt][1] - n1) * scale_log2e);
                mrun[mt][0] = n0; mrun[mt][1] = n1;
                nms[mt][0] = n0 * scale_log2e; nms[mt][1] = n1 * scale_log2e;
                lrun[mt][0] *= c0; lrun[mt][1] *= c1;
#pragma unroll
                for (int dn = 0; dn < DN; ++dn) {
                    oacc[mt][dn][0] *= c0; oacc[mt][dn][1] *= c0;
                    oacc[mt][dn][2] *= c1; oacc[mt][dn][3] *= c1;
                }
            }
            // P = exp2(s * scale_log2e - m * scale_log2e); O += P V, 16 keys (two n-tiles of S) per k-step
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                uint32_t pa[MT][4];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int nt = 2 * kk + hh;
                        const float p00 = ex2_approx(fmaf(s[mt][nt][0], scale_log2e, -nms[mt][0]));
                        const float p01 = ex2_approx(fmaf(s[mt][nt][1], scale_log2e, -nms[mt][0]));
                        const float p10 = ex2_approx(fmaf(s[mt][nt][2], scale_log2e, -nms[mt][1]));
                        const float p11 = ex2_approx(fmaf(s[mt][nt][3], scale_log2e, -nms[mt][1]));
                        lrun[mt][0] += p00 + p01;
                        lrun[mt][1] += p10 + p11;
                        pa[mt][2 * hh] = pack_h2(p00, p01);          // row g
                        pa[mt][2 * hh + 1] = pack_h2(p10, p11);      // row g + 8
                    }
                }
                // V fragment of (k-step kk, n-tile dn): transposed 8x8 matrices, keys 16 kk + {0, 8}, columns 8 dn
#pragma unroll
                for (int dn = 0; dn < DN; ++dn) {
                    uint32_t vb[2];
                    ldsm_x2_trans(Vt + (uint32_t)(((16 * kk + 8 * (lm & 1) + lr) * PH + 8 * dn) * 2), vb);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) mma_f16_16816(oacc[mt][dn], pa[mt], vb[0], vb[1]);
                }
            }
        };
        if (nk < KT) tile_body(std::true_type{});
        else tile_body(std::false_type{});
        __syncthreads();
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        float l0 = lrun[mt][0], l1 = lrun[mt][1];
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float i0 = 1.0f / l0, i1 = 1.0f / l1;
        const int r0 = wrow + mt * 16 + g, r1 = r0 + 8;
        if (o_half) {
            // fp16 result (o_pitch in halves): the A operand of the to_out projection (afldm_conv2d_f16in_f32)
            __half* oh = reinterpret_cast<__half*>(o);
            if (r0 < Nq) {
                __half* op = oh + ((size_t)b * Nq + r0) * o_pitch + head * D + 2 * t;
#pragma unroll
                for (int dn = 0; dn < DN; ++dn)
                    *reinterpret_cast<__half2*>(op + 8 * dn) = __floats2half2_rn(oacc[mt][dn][0] * i0, oacc[mt][dn][1] * i0);
            }
            if (r1 < Nq) {
                __half* op = oh + ((size_t)b * Nq + r1) * o_pitch + head * D + 2 * t;
#pragma unroll
                for (int dn = 0; dn < DN; ++dn)
                    *reinterpret_cast<__half2*>(op + 8 * dn) = __floats2half2_rn(oacc[mt][dn][2] * i1, oacc[mt][dn][3] * i1);
            }
            continue;
        }
        if (r0 < Nq) {
            float* op = o + ((size_t)b * Nq + r0) * o_pitch + head * D + 2 * t;
#pragma unroll
            for (int dn = 0; dn < DN; ++dn)
                *reinterpret_cast<float2*>(op + 8 * dn) = make_float2(oacc[mt][dn][0] * i0, oacc[mt][dn][1] * i0);
        }
        if (r1 < Nq) {
            float* op = o + ((size_t)b * Nq + r1) * o_pitch + head * D + 2 * t;
#pragma unroll
            for (int dn = 0; dn < DN; ++dn)
                *reinterpret_cast<float2*>(op + 8 * dn) = make_float2(oacc[mt][dn][2] * i1, oacc[mt][dn][3] * i1);
        }
    }
}

template <int D, int MT>
int launch_f16_mt(const __half* q, int q_pitch, const __half* k, const __half* v, int kv_pitch, float* o, int o_pitch,
                  int B, int Bkv, int Nq, int Nk, int heads, cudaStream_t st, int o_half) {
    const float scale_log2e = (float)((1.0 / sqrt((double)D)) * 1.4426950408889634);
    constexpr int DP = (D + 15) / 16 * 16;
    constexpr int smem = 4 * 64 * (DP + 8) * 2;
    if (smem > 48 * 1024) {
        static std::atomic<unsigned long long> configured{0};
        cudaError_t e = set_max_dyn_smem(attention_f16_kernel<D, MT>, smem, configured);
        if (e != cudaSuccess) return (int)e;
    }
    launch_k(attention_f16_kernel<D, MT>, dim3(ceil_div(Nq, 64 * MT), heads, B), dim3(128), smem, st,
        q, q_pitch, k, v, kv_pitch, o, o_pitch, B / Bkv, Nq, Nk, scale_log2e, o_half);
    return launched();
}

template <int D>
int launch_f16(const __half* q, int q_pitch, const __half* k, const __half* v, int kv_pitch, float* o, int o_pitch,
               int B, int Bkv, int Nq, int Nk, int heads, cudaStream_t st, int o_half) {
    static const int force_mt = getenv("AFLDM_ATTN_MT") ? atoi(getenv("AFLDM_ATTN_MT")) : 0;
    if (force_mt != 1 && D <= 32 && Nq >= 256 && (long long)ceil_div(Nq, 128) * heads * B >= 2 * 148)
        return launch_f16_mt<D, 2>(q, q_pitch, k, v, kv_pitch, o, o_pitch, B, Bkv, Nq, Nk, heads, st, o_half);
    return launch_f16_mt<D, 1>(q, q_pitch, k, v, kv_pitch, o, o_pitch, B, Bkv, Nq, Nk, heads, st, o_half);
}

template <int D>
int launch(const float* q, int q_pitch, const float* k, const float* v, int kv_pitch, float* o, int o_pitch,
           int B, int Bkv, int Nq, int Nk, int heads, cudaStream_t st) {
    const int threads = Nq >= ATT_THREADS ? ATT_THREADS : ((Nq + 31) / 32) * 32;
    const float qscale = (float)((1.0 / sqrt((double)D)) * 1.4426950408889634);
    launch_k(attention_kernel<D>, dim3(ceil_div(Nq, threads), heads, B), dim3(threads), 0, st, 
        q, q_pitch, k, v, kv_pitch, o, o_pitch, B / Bkv, Nq, Nk, qscale);
    return launched();
}

}  // namespace
}  // namespace afldm

using namespace afldm;

extern "C" int afldm_softmax_rows_f32(float* x, long long rows, int cols, int pitch, float scale,
                                      afldm_stream_t stream) {
    if (x == nullptr || rows <= 0 || cols <= 0 || pitch < cols) return AFLDM_E_ARG;
    const long long blocks = (rows + 7) / 8;
    if (blocks > 0x7fffffffLL) return AFLDM_E_SHAPE;
    launch_k(softmax_rows_kernel, dim3((unsigned)blocks), dim3(256), 0, as_stream(stream), x, rows, cols, pitch,
                                                                         scale * 1.4426950408889634f);
    return launched();
}

extern "C" int afldm_attention_f32(const float* q, int q_pitch, const float* k, const float* v, int kv_pitch,
                                   float* o, int o_pitch, int B, int Bkv, int Nq, int Nk, int heads, int d,
                                   int algo, afldm_stream_t stream) {
    if (q == nullptr || k == nullptr || v == nullptr || o == nullptr) return AFLDM_E_ARG;
    if (B <= 0 || Bkv <= 0 || Nq <= 0 || Nk <= 0 || heads <= 0 || d <= 0) return AFLDM_E_ARG;
    if (B % Bkv != 0) return AFLDM_E_SHAPE;
    if (d % 4 != 0 || q_pitch % 4 != 0 || kv_pitch % 4 != 0 || o_pitch % 4 != 0) return AFLDM_E_SHAPE;
    if (q_pitch < heads * d || kv_pitch < heads * d || o_pitch < heads * d) return AFLDM_E_ARG;
    if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(o)) return AFLDM_E_ARG;
    cudaStream_t st = as_stream(stream);
    if (algo == AFLDM_ATTN_MMA_TF32) {
#define AFLDM_ATT_MMA(D) \
    case D: return launch_mma<D>(q, q_pitch, k, v, kv_pitch, o, o_pitch, B, Bkv, Nq, Nk, heads, st);
        switch (d) {
            AFLDM_ATT_MMA(8)
            AFLDM_ATT_MMA(16)
            AFLDM_ATT_MMA(24)
            AFLDM_ATT_MMA(32)
            AFLDM_ATT_MMA(40)
            AFLDM_ATT_MMA(48)
            AFLDM_ATT_MMA(64)
            AFLDM_ATT_MMA(80)
            AFLDM_ATT_MMA(160)
            default: return AFLDM_E_NOKERNEL;
        }
#undef AFLDM_ATT_MMA
    }
    if (algo != AFLDM_ATTN_SIMT_F32) return AFLDM_E_ARG;
#define AFLDM_ATT_CASE(D) \
    case D: return launch<D>(q, q_pitch, k, v, kv_pitch, o, o_pitch, B, Bkv, Nq, Nk, heads, st);
    switch (d) {
        AFLDM_ATT_CASE(8)
        AFLDM_ATT_CASE(16)
        AFLDM_ATT_CASE(24)
        AFLDM_ATT_CASE(32)
        AFLDM_ATT_CASE(40)
        AFLDM_ATT_CASE(48)
        AFLDM_ATT_CASE(64)
        default: return AFLDM_E_NOKERNEL;
    }
#undef AFLDM_ATT_CASE
}

static int attention_f16_impl(const void* q, int q_pitch, const void* k, const void* v, int kv_pitch, float* o,
                              int o_pitch, int B, int Bkv, int Nq, int Nk, int heads, int d,
                              afldm_stream_t stream, int o_half) {
    if (q == nullptr || k == nullptr || v == nullptr || o == nullptr) return AFLDM_E_ARG;
    if (B <= 0 || Bkv <= 0 || Nq <= 0 || Nk <= 0 || heads <= 0 || d <= 0) return AFLDM_E_ARG;
    if (B % Bkv != 0) return AFLDM_E_SHAPE;
    if (d % 8 != 0 || q_pitch % 8 != 0 || kv_pitch % 8 != 0 || o_pitch % 2 != 0) return AFLDM_E_SHAPE;
    if (q_pitch < heads * d || kv_pitch < heads * d || o_pitch < heads * d) return AFLDM_E_ARG;
    if (!aligned16(q) || !aligned16(k) || !aligned16(v) || (reinterpret_cast<uintptr_t>(o) & 7u) != 0) return AFLDM_E_ARG;
    cudaStream_t st = as_stream(stream);
    const __half* qh = static_cast<const __half*>(q);
    const __half* kh = static_cast<const __half*>(k);
    const __half* vh = static_cast<const __half*>(v);
#define AFLDM_ATT_F16(D) \
    case D: return launch_f16<D>(qh, q_pitch, kh, vh, kv_pitch, o, o_pitch, B, Bkv, Nq, Nk, heads, st, o_half);
    switch (d) {
        AFLDM_ATT_F16(8)
        AFLDM_ATT_F16(16)
        AFLDM_ATT_F16(24)
        AFLDM_ATT_F16(32)
        AFLDM_ATT_F16(40)
        AFLDM_ATT_F16(48)
        AFLDM_ATT_F16(64)
        default: return AFLDM_E_NOKERNEL;
    }
#undef AFLDM_ATT_F16
}

extern "C" int afldm_attention_f16(const void* q, int q_pitch, const void* k, const void* v, int kv_pitch, float* o,
                                   int o_pitch, int B, int Bkv, int Nq, int Nk, int heads, int d,
                                   afldm_stream_t stream) {
    return attention_f16_impl(q, q_pitch, k, v, kv_pitch, o, o_pitch, B, Bkv, Nq, Nk, heads, d, stream, 0);
}

extern "C" int afldm_attention_f16_f16out(const void* q, int q_pitch, const void* k, const void* v, int kv_pitch, void* o,
                                          int o_pitch, int B, int Bkv, int Nq, int Nk, int heads, int d,
                                          afldm_stream_t stream) {
    if ((reinterpret_cast<uintptr_t>(o) & 3u) != 0) return AFLDM_E_ARG;
    return attention_f16_impl(q, q_pitch, k, v, kv_pitch, static_cast<float*>(o), o_pitch, B, Bkv, Nq, Nk, heads, d, stream, 1);
}
