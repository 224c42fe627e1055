// Filtered activation (WarpedNonlinearity, /root/reference/afldm/af_modules/af_blocks.py:19-28) on tcgen05:
//      y = D act(U x U^T) D^T        per (b, c) plane of side n = 32 or 16
// (the FFT form of the reference, afldm/af_libs/ideal_lpf.py:69-158, restated as circulant matrices - SURVEY.md 8(a)).
//
// Every 1-D circular convolution is a GEMM whose M dimension is 128 LINES (one line = one row or column of one
// channel) issued as tcgen05.mma.kind::f16 (M = 128, N = n or 2n, K = 16) with the accumulator in TMEM.  fp32
// accuracy comes from the 3-term fp16 split of resample.cu (x = xh + xl, F = Fh + Fl; xl Fh + xh Fl + xh Fh, fp32
// accumulate: 22 significant bits per operand).  One CTA works on UNITS of (image b, CG channels); per unit:
//
//   L    global -> registers -> planar fp32 staging [c][i][j] (GroupNorm affine applied on the way, conflict-free)
//   P0   rows up      thread = line (c, i): splits its row into a K-major operand row [xh | xl];  D0 = odd columns
//   S1   the thread owns row i of T = x U^T (even columns = x, odd = D0) and writes it as row k = i of the
//        MN-MAJOR operand of P1 (m = (c, j') contiguous: 16-byte stores) - the row/column transpose costs nothing
//   P1   columns up   lines (c, j'), K = i, B = the full 2n x n up-sampling matrix (identity rows = even samples)
//   act  thread = line (c, j'): reads the 2n values of column j' of Z from TMEM, applies the activation
//   P2   columns down  odd samples -> K-major operand row [oh | ol], B = circulant G; the half-band even part
//        e[i] / 2 is PRE-STORED into the accumulator with tcgen05.st, the O(n) alternating-sum term added after
//   S2   thread owns column j' of Y1 = D A and writes it as row k = j' of the MN-major operand of P3
//   P3   rows down    lines (c, i), K = j' (2n), B = the full n x 2n down-sampling matrix
//   out  TMEM -> planar staging -> coalesced (CG channels contiguous per pixel) fp32 or fp16 stores
//
// Operand tiles use SWIZZLE_128B; the MN-major descriptor convention (LBO = M-atom stride, SBO = 8-k group stride) and
// thread-written tiles + fence.proxy.async were verified on a B200 with tools/ubench/umma_probe.cu
// (profiles/r02_umma_probe.txt).  Filter operands (24 KB for n = 32) are built ONCE per device from the fp32 taps
// (taps.inc) by fact_tc_filters_kernel into a global table; every persistent CTA copies the table to shared memory.
// Measured bounds on B200 (same probe): TMEM -> register loads run at ~57 B/clk/SM, which together with 2 MUFU per
// SiLU on the 4x up-sampled plane sets the floor of this formulation.
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"
#include "fact_common.cuh"
#include "fact_tc.cuh"
#include "taps.inc"
#include "tc_ptx.cuh"

namespace afldm {
namespace {

using namespace tc;

constexpr int FT_THREADS = 256;

template <int N>
struct Cfg {
    static constexpr int CG = 256 / N;             // channels per unit: 8 (n = 32) / 16 (n = 16); CG * N = 256 lines
    static constexpr int PITCH = N + 1;            // planar staging: row pitch (floats)
    static constexpr int PLANE = N == 32 ? 1060 : 274;   // plane pitch: >= N * PITCH, bank offset chosen for the L phase
    static constexpr int CPT1 = 128 / (2 * N);     // channels per P1 / P2 tile (128 lines (c, j'))
    static constexpr int CPT3 = 128 / N;           // channels per P0 / P3 tile (128 lines (c, i))
    static constexpr int KG1 = 2 * N / 8;          // 8-k groups of a P1 operand tile: [hi n | lo n]
    static constexpr int KG3 = 4 * N / 8;          // ... of a P3 operand tile: [hi 2n | lo 2n]
    static constexpr int MS1 = KG1 * 1024;         // M-atom stride (bytes) of the MN-major tiles
    static constexpr int MS3 = KG3 * 1024;
    static constexpr int TILE = 16384;             // stride of the P0 / P1 / P2 tiles
    static constexpr int TILE3 = 2 * MS3;          // 32 KB / 16 KB
    // filter tiles, all K-major with 128-byte rows, contiguous: row index r_all * 128
    static constexpr int B0 = 0;                   // [n]  [dh | dl]      odd-phase up-sampler circulant
    static constexpr int B1 = N * 128;             // [2n] [Uh | Ul]      full up-sampling matrix
    static constexpr int B2 = 3 * N * 128;         // [n]  [Gh | Gl]      odd-sample down-sampler circulant
    static constexpr int B3H = 4 * N * 128;        // [n]  Dh (2n k)      full down-sampling matrix, hi
    static constexpr int B3L = 5 * N * 128;        // [n]  Dl (2n k)
    static constexpr int FILT_BYTES = 6 * N * 128; // 24 KB / 12 KB
    static constexpr int STAGE_OFF = 32768;        // planar staging behind the two P0 operand tiles
    static constexpr int STAGE_BYTES = CG * PLANE * 4;
    static constexpr int AREG_BYTES = (STAGE_OFF + STAGE_BYTES > 65536 ? ((STAGE_OFF + STAGE_BYTES + 1023) / 1024) * 1024 : 65536);
    static constexpr int TMEM_COLS = 8 * N;        // 4 tiles x 2n accumulator columns
    static constexpr int SMEM_BYTES = FILT_BYTES + AREG_BYTES + 1024;   // + alignment slack
};

// fp32 tap of filter tile `tile` at (output row r, input k)
template <int N>
__device__ __forceinline__ float filt_value(int tile, int r, int k) {
    switch (tile) {
        case 0: return tap_d<N>((r - k) & (N - 1));
        case 1: return (r & 1) ? tap_d<N>(((r >> 1) - k) & (N - 1)) : ((r >> 1) == k ? 1.f : 0.f);
        case 2: return tap_g<N>((2 * ((r - k) & (N - 1)) - 1) & (2 * N - 1));
        default: return tap_g<N>((2 * r - k) & (2 * N - 1));
    }
}

template <int N>
__device__ __forceinline__ void build_filters(uint32_t filt) {
    for (int ch = threadIdx.x; ch < 6 * N * 8; ch += FT_THREADS) {
        const int row = ch >> 3, p = ch & 7;
        const int tile = row < N ? 0 : (row < 3 * N ? 1 : (row < 4 * N ? 2 : (row < 5 * N ? 3 : 4)));
        const int r = row - (tile == 0 ? 0 : (tile == 1 ? N : (tile == 2 ? 3 * N : (tile == 3 ? 4 * N : 5 * N))));
        uint32_t w[4];
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
            __half hv[2];
#pragma unroll
            for (int e1 = 0; e1 < 2; ++e1) {
                const int kp = 8 * p + 2 * e2 + e1;
                float v = 0.f;
                bool lo;
                if (tile < 3) {                       // [hi n | lo n]
                    lo = kp >= N;
                    if (kp < 2 * N) v = filt_value<N>(tile, r, lo ? kp - N : kp);
                } else {                              // hi tile / lo tile over 2n inputs
                    lo = tile == 4;
                    if (kp < 2 * N) v = filt_value<N>(3, r, kp);
                }
                const __half h = __float2half_rn(v);
                hv[e1] = lo ? __float2half_rn(v - __half2float(h)) : h;
            }
            const __half2 h2 = __halves2half2(hv[0], hv[1]);
            w[e2] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        sts128(filt + (uint32_t)(row * 128 + ((p ^ (row & 7)) << 4)), w[0], w[1], w[2], w[3]);
    }
}

// MN-major SWIZZLE_128B tile: byte offset of the 16-byte chunk that starts at (m, k'), m a multiple of 8
__device__ __forceinline__ uint32_t mn_off(int m, int kp, int ms) {
    return (uint32_t)((m >> 6) * ms + (kp >> 3) * 1024 + (kp & 7) * 128 + ((((m & 63) >> 3) ^ (kp & 7)) << 4));
}

// ---- MMA issue (one elected thread).  Term order: small terms first (lo x Fh, hi x Fl, hi x Fh).
template <int N>
__device__ __forceinline__ void issue_kmajor(uint32_t tmem, uint32_t areg, uint32_t btile, int tiles, int dstride, bool preloaded) {
    const uint32_t id = idesc_f16(N, false);
    for (int t = 0; t < tiles; ++t) {
        const uint32_t a = areg + (uint32_t)t * Cfg<N>::TILE, d = tmem + (uint32_t)(t * dstride);
        uint32_t acc = preloaded ? 1u : 0u;
#pragma unroll
        for (int term = 0; term < 3; ++term) {
            const int a_el = term == 0 ? N : 0, b_el = term == 1 ? N : 0;
#pragma unroll
            for (int ks = 0; ks < N / 16; ++ks) {
                umma_f16(d, desc_kmajor(a + (uint32_t)((a_el + 16 * ks) * 2)), desc_kmajor(btile + (uint32_t)((b_el + 16 * ks) * 2)), id, acc);
                acc = 1u;
            }
        }
    }
}

template <int N>
__device__ __forceinline__ void issue_p1(uint32_t tmem, uint32_t areg, uint32_t filt, int tiles) {
    using K = Cfg<N>;
    const uint32_t id = idesc_f16(2 * N, true);
    for (int mt = 0; mt < tiles; ++mt) {
        const uint32_t a = areg + (uint32_t)mt * K::TILE, d = tmem + (uint32_t)(mt * 2 * N);
        uint32_t acc = 0u;
#pragma unroll
        for (int term = 0; term < 3; ++term) {
            const int kg0 = term == 0 ? N / 8 : 0, b_el = term == 1 ? N : 0;
#pragma unroll
            for (int ks = 0; ks < N / 16; ++ks) {
                umma_f16(d, desc_sw128(a + (uint32_t)((kg0 + 2 * ks) * 1024), K::MS1, 1024),
                         desc_kmajor(filt + K::B1 + (uint32_t)((b_el + 16 * ks) * 2)), id, acc);
                acc = 1u;
            }
        }
    }
}

template <int N>
__device__ __forceinline__ void issue_p3(uint32_t tmem, uint32_t areg, uint32_t filt, int tiles) {
    using K = Cfg<N>;
    const uint32_t id = idesc_f16(N, true);
    for (int t = 0; t < tiles; ++t) {
        const uint32_t a = areg + (uint32_t)t * K::TILE3, d = tmem + (uint32_t)(t * N);
        uint32_t acc = 0u;
#pragma unroll
        for (int term = 0; term < 3; ++term) {
            const int kg0 = term == 0 ? 2 * N / 8 : 0;
            const uint32_t bt = filt + (term == 1 ? K::B3L : K::B3H);
#pragma unroll
            for (int ks = 0; ks < 2 * N / 16; ++ks) {
                umma_f16(d, desc_sw128(a + (uint32_t)((kg0 + 2 * ks) * 1024), K::MS3, 1024), desc_kmajor(bt + (uint32_t)(32 * ks)), id, acc);
                acc = 1u;
            }
        }
    }
}

// One-time (per device) build of the filter operand tiles into a global table; the kernels copy it (24 KB / 12 KB).
__device__ uint4 g_filters32[Cfg<32>::FILT_BYTES / 16];
__device__ uint4 g_filters16[Cfg<16>::FILT_BYTES / 16];

template <int N>
__global__ void __launch_bounds__(FT_THREADS) fact_tc_filters_kernel() {
    extern __shared__ __align__(1024) uint8_t fsm[];
    pdl_trigger();
    const uint32_t filt = (smem_u32(fsm) + 1023u) & ~1023u;
    build_filters<N>(filt);
    __syncthreads();
    uint4* dst = N == 32 ? g_filters32 : g_filters16;
    const uint4* src = reinterpret_cast<const uint4*>(fsm + (filt - smem_u32(fsm)));
    for (int i = threadIdx.x; i < Cfg<N>::FILT_BYTES / 16; i += FT_THREADS) dst[i] = src[i];
}

__device__ __forceinline__ void bar_wg(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory"); }

// tcgen05.wait::ld with the loaded registers as in/out operands: nothing that consumes them can be scheduled above it
__device__ __forceinline__ void tmem_wait_ld16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}

// The two warpgroups of a CTA are INDEPENDENT pipelines between the shared load and store phases: warpgroup g owns
// channels [g CG/2, (g + 1) CG/2) of the unit = P0 / P3 tile g and P1 / P2 tiles 2g, 2g + 1, its own shared-memory
// operand tiles, TMEM columns, mbarrier and MMA-issuing thread, and synchronises with a named barrier of 128 threads:
// while one warpgroup waits for its MMAs the other (and the second CTA on the SM) computes.
template <int N, int ACT>
__global__ void __launch_bounds__(FT_THREADS, 2)
fact_tc_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int C, const Affine af) {
    using K = Cfg<N>;
    constexpr int CG = K::CG, PLANE = K::PLANE, PITCH = K::PITCH;
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    __shared__ float s_sc[CG], s_sh[CG];
    __shared__ uint64_t s_bar[2];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int wg = warp >> 2;                               // warpgroup = pipeline
    const int wt = tid & 127;                               // TMEM lane of this thread in every tile of its warpgroup
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t filt = (raw + 1023u) & ~1023u;           // swizzle atoms need 1024-byte alignment
    const uint32_t areg = filt + K::FILT_BYTES;
    uint8_t* const filt_p = smem_raw + (filt - raw);
    uint8_t* const areg_p = filt_p + K::FILT_BYTES;
    float* const stage = reinterpret_cast<float*>(areg_p + K::STAGE_OFF);
    float* const ystage = reinterpret_cast<float*>(areg_p);
    const uint32_t bar = smem_u32(&s_bar[wg]);
    // Operand region of this warpgroup: [wg * 32 KB, + 32 KB) of the A area in EVERY phase (P0 tile, the two P1 / P2 tiles,
    // the P3 tile), so the de-synchronised warpgroups never touch each other's tiles.
    const uint32_t wreg = areg + (uint32_t)(wg * 2 * K::TILE);

    if (tid == 0) {
        mbar_init(smem_u32(&s_bar[0]), 1);
        mbar_init(smem_u32(&s_bar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"((uint32_t)K::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // this warpgroup's accumulator columns: two P1 tiles of 2n columns each
    const uint32_t tmem = s_tmem + (uint32_t)(wg * 4 * N);
    const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);     // + this warp's lane quarter
    pdl_wait();
    {   // filter operand tiles (built once per device by fact_tc_filters_kernel): global table -> shared memory
        const uint4* src = N == 32 ? g_filters32 : g_filters16;
        uint4* dst = reinterpret_cast<uint4*>(filt_p);
        for (int i = tid; i < K::FILT_BYTES / 16; i += FT_THREADS) dst[i] = __ldg(src + i);
        fence_async_smem();                                  // made visible to the tensor core by the first unit's barriers
    }

    const int groups_per_img = C / CG;
    const int units = B * groups_per_img;
    uint32_t ph = 0;
    // per-thread constants.  P0 / P3 line of this thread: channel lc, row li (tile wg, lane wt)
    const int lc = tid / N, li = tid % N;
    const int sw16 = (tid & 7) << 4;
    const uint32_t a0row = (wreg + (uint32_t)(wt * 128)) ^ (uint32_t)sw16;    // K-major row of P0: chunk g at ^ (g << 4)
    // S1: row k = li (hi) / n + li (lo) of the MN-major P1 tile of channel lc
    const int mt1 = lc / K::CPT1, cl1 = lc % K::CPT1;
    const uint32_t s1hi = areg + (uint32_t)(mt1 * K::TILE) + mn_off(cl1 * 2 * N, li, K::MS1);
    const uint32_t s1lo = areg + (uint32_t)(mt1 * K::TILE) + mn_off(cl1 * 2 * N, N + li, K::MS1);
    // (mn_off at a 64-aligned... chunk index of m within the atom is XOR-ed with k' % 8: chunk g of the run -> ^ (g << 4))
    constexpr int CH = CG / 4;                          // 16-byte chunks per pixel
    constexpr int NCH = N * N * CH / FT_THREADS;        // chunks per thread: 8 / 4
    const int lpix = tid / CH, lh = tid % CH;           // first chunk of this thread: pixel, channel quad; next: + 256 / CH pixels
    float* const st_dst = stage + (4 * lh) * PLANE + (lpix / N) * PITCH + (lpix % N);
    constexpr int ST_STEP = (FT_THREADS / CH / N) * PITCH;   // staging rows advanced per chunk step

    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int b = u / groups_per_img, c0 = (u - b * groups_per_img) * CG;

        // ---- L: this unit's pixels -> registers (all loads in flight), GroupNorm finalise, -> planar staging
        const XSrc xs = x_source(x, C, af, c0);
        float4 v[NCH];
        {
            const float* src = xs.p + ((size_t)b * N * N + lpix) * xs.pitch + 4 * lh;
            const size_t step = (size_t)(FT_THREADS / CH) * xs.pitch;
#pragma unroll
            for (int it = 0; it < NCH; ++it) v[it] = __ldg(reinterpret_cast<const float4*>(src + it * step));
        }
        if (af.pa != nullptr) {
            gn_prologue<CG>(af, b, c0, C, s_sc, s_sh);
        } else {
            if (tid < CG) {
                s_sc[tid] = af.scale != nullptr ? af.scale[(size_t)b * C + c0 + tid] : 1.f;
                s_sh[tid] = af.shift != nullptr ? af.shift[(size_t)b * C + c0 + tid] : 0.f;
            }
            __syncthreads();
        }
        {
            const float sc0 = s_sc[4 * lh], sc1 = s_sc[4 * lh + 1], sc2 = s_sc[4 * lh + 2], sc3 = s_sc[4 * lh + 3];
            const float sh0 = s_sh[4 * lh], sh1 = s_sh[4 * lh + 1], sh2 = s_sh[4 * lh + 2], sh3 = s_sh[4 * lh + 3];
#pragma unroll
            for (int it = 0; it < NCH; ++it) {
                float* dst = st_dst + it * ST_STEP;
                dst[0] = fmaf(v[it].x, sc0, sh0);
                dst[PLANE] = fmaf(v[it].y, sc1, sh1);
                dst[2 * PLANE] = fmaf(v[it].z, sc2, sh2);
                dst[3 * PLANE] = fmaf(v[it].w, sc3, sh3);
            }
        }
        __syncthreads();

        // ---- P0: rows up.  Line (lc, li) = TMEM lane wt of tile wg.
        float xr[N];
        {
            const float* src = stage + lc * PLANE + li * PITCH;
#pragma unroll
            for (int j = 0; j < N; ++j) xr[j] = src[j];
        }
        __syncthreads();             // staging is consumed: the (de-synchronised) warpgroups may overwrite it from S1 on
#pragma unroll
        for (int g = 0; g < N / 8; ++g) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_pack(xr[8 * g + 2 * e], xr[8 * g + 2 * e + 1], hi[e], lo[e]);
            sts128(a0row ^ (uint32_t)(g << 4), hi[0], hi[1], hi[2], hi[3]);
            sts128(a0row ^ (uint32_t)((N / 8 + g) << 4), lo[0], lo[1], lo[2], lo[3]);
        }
        fence_async_smem();
        tc_fence_before();
        bar_wg(wg);
        if ((warp & 3) == 0) {
            tc_fence_after();
            if (elect_one()) {
                issue_kmajor<N>(tmem, wreg, filt + K::B0, 1, 0, false);
                umma_commit(bar);
            }
            __syncwarp();
        }
        mbar_wait(bar, ph);
        ph ^= 1u;
        tc_fence_after();
        float od[N];
        tmem_ld<N>(trow, od);

        // ---- S1: row li of T (even columns xr, odd columns od) -> row k = li of the MN-major P1 operand of channel lc
#pragma unroll
        for (int g = 0; g < 2 * N / 8; ++g) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_pack(xr[4 * g + e], od[4 * g + e], hi[e], lo[e]);
            // 8 g more M-elements: same atom for N = 32 (one channel = 64 columns = one atom row); for N = 16 a channel
            // is half an atom row, so chunk (cl1 % 2) * 4 + g
            sts128(s1hi ^ (uint32_t)(g << 4), hi[0], hi[1], hi[2], hi[3]);
            sts128(s1lo ^ (uint32_t)(g << 4), lo[0], lo[1], lo[2], lo[3]);
        }
        fence_async_smem();
        tc_fence_before();
        bar_wg(wg);
        if ((warp & 3) == 0) {
            tc_fence_after();
            if (elect_one()) {
                issue_p1<N>(tmem, wreg, filt, 2);
                umma_commit(bar);
            }
            __syncwarp();
        }
        mbar_wait(bar, ph);
        ph ^= 1u;
        tc_fence_after();

        // ---- act + P2 operands: line (c, j') = lane wt of the warpgroup's two P1 tiles.  The TMEM load of the next
        // 16 columns is in flight while the current 16 are processed.
        float sline[2];
        {
            constexpr int QN = 2 * N / 16;                  // 16-column chunks per line
            uint32_t cur[16], nxt[16];
            tmem_ld16(trow, cur);
#pragma unroll
            for (int st = 0; st < 2 * QN; ++st) {
                const int ln = st / QN, q = st % QN;
                const uint32_t tcol = trow + (uint32_t)(ln * 2 * N);
                const uint32_t rowb = (wreg + (uint32_t)(ln * K::TILE + wt * 128)) ^ (uint32_t)sw16;
                tmem_wait_ld16(cur);
                if (st + 1 < 2 * QN) tmem_ld16(trow + (uint32_t)(((st + 1) / QN) * 2 * N + ((st + 1) % QN) * 16), nxt);
                float a[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) a[k] = apply_act<ACT>(__uint_as_float(cur[k]));
                // even samples e[m] = a[2m]: half-band identity  y[i] = e[i] / 2 - (-1)^i S / (2n) + (odd-sample convolution)
                float ev[8];
                float s = q == 0 ? 0.f : sline[ln];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    s += (e & 1) ? -a[2 * e] : a[2 * e];
                    ev[e] = 0.5f * a[2 * e];
                }
                sline[ln] = s;
                tmem_st8(tcol + (uint32_t)(8 * q), ev);       // columns already consumed by this thread: D2 starts as e / 2
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) split_pack(a[4 * e + 1], a[4 * e + 3], hi[e], lo[e]);
                sts128(rowb ^ (uint32_t)(q << 4), hi[0], hi[1], hi[2], hi[3]);
                sts128(rowb ^ (uint32_t)((N / 8 + q) << 4), lo[0], lo[1], lo[2], lo[3]);
#pragma unroll
                for (int k = 0; k < 16; ++k) cur[k] = nxt[k];
            }
            sline[0] *= 1.0f / (2 * N);
            sline[1] *= 1.0f / (2 * N);
        }
        tmem_wait_st();
        fence_async_smem();
        tc_fence_before();
        bar_wg(wg);
        if ((warp & 3) == 0) {
            tc_fence_after();
            if (elect_one()) {
                issue_kmajor<N>(tmem, wreg, filt + K::B2, 2, 2 * N, true);
                umma_commit(bar);
            }
            __syncwarp();
        }
        mbar_wait(bar, ph);
        ph ^= 1u;
        tc_fence_after();

        // ---- S2: column j' of Y1 = D A -> row k = j' of the MN-major P3 operand (tile wg)
#pragma unroll
        for (int ln = 0; ln < 2; ++ln) {
            const int cc = (2 * wg + ln) * K::CPT1 + wt / (2 * N), jp = wt % (2 * N);
            float y1[N];
            tmem_ld<N>(trow + (uint32_t)(ln * 2 * N), y1);
#pragma unroll
            for (int i2 = 0; i2 < N; ++i2) y1[i2] += (i2 & 1) ? sline[ln] : -sline[ln];
            const int cl3 = cc % K::CPT3;
            const uint32_t tb = wreg;
            const uint32_t s2hi = tb + mn_off(cl3 * N, jp, K::MS3), s2lo = tb + mn_off(cl3 * N, 2 * N + jp, K::MS3);
#pragma unroll
            for (int g = 0; g < N / 8; ++g) {
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) split_pack(y1[8 * g + 2 * e], y1[8 * g + 2 * e + 1], hi[e], lo[e]);
                sts128(s2hi ^ (uint32_t)(g << 4), hi[0], hi[1], hi[2], hi[3]);
                sts128(s2lo ^ (uint32_t)(g << 4), lo[0], lo[1], lo[2], lo[3]);
            }
        }
        fence_async_smem();
        tc_fence_before();
        bar_wg(wg);
        if ((warp & 3) == 0) {
            tc_fence_after();
            if (elect_one()) {
                issue_p3<N>(tmem, wreg, filt, 1);
                umma_commit(bar);
            }
            __syncwarp();
        }
        mbar_wait(bar, ph);
        ph ^= 1u;
        tc_fence_after();

        // ---- out: row li of y (channel lc) -> planar staging -> CG channels per pixel, coalesced
        float yr[N];
        tmem_ld<N>(trow, yr);
        tc_fence_before();
        __syncthreads();             // both pipelines are past their last MMA: the operand tiles may be overwritten
        {
            float* dst = ystage + lc * PLANE + li * PITCH;
#pragma unroll
            for (int j = 0; j < N; ++j) dst[j] = yr[j];
        }
        __syncthreads();
        {
            const size_t off0 = ((size_t)b * N * N + tid) * C + c0;
            const float* src0 = ystage + (tid / N) * PITCH + (tid % N);
#pragma unroll
            for (int it = 0; it < N * N / FT_THREADS; ++it) {
                const float* src = src0 + it * (FT_THREADS / N) * PITCH;
                float o[CG];
#pragma unroll
                for (int cc = 0; cc < CG; ++cc) o[cc] = src[cc * PLANE];
                const size_t off = off0 + (size_t)it * FT_THREADS * C;
                if (af.y_half) {
                    uint4* dp = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(y) + off);
#pragma unroll
                    for (int g = 0; g < CG / 8; ++g) {
                        uint32_t w[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const __half2 h2 = __floats2half2_rn(o[8 * g + 2 * e], o[8 * g + 2 * e + 1]);
                            w[e] = *reinterpret_cast<const uint32_t*>(&h2);
                        }
                        dp[g] = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                } else {
                    float4* dp = reinterpret_cast<float4*>(y + off);
#pragma unroll
                    for (int g = 0; g < CG / 4; ++g) dp[g] = make_float4(o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
                }
            }
        }
        __syncthreads();          // staging is rewritten by the next unit
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"((uint32_t)K::TMEM_COLS) : "memory");
}

int sm_count_of_current_device() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// The filter table of plane size N on the current device: built by one tiny kernel on `st` in front of the first
// filtered activation that needs it (stream order makes it visible; idempotent, so a racing second stream at worst
// rebuilds the same bytes).
template <int N>
int ensure_filters(cudaStream_t st) {
    static bool built[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return AFLDM_E_ARG;
    if (built[dev]) return 0;
    auto kern = fact_tc_filters_kernel<N>;
    const int smem = Cfg<N>::FILT_BYTES + 1024;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    launch_k(kern, dim3(1), dim3(FT_THREADS), smem, st);
    const int r = launched();
    if (r == 0) built[dev] = true;
    return r;
}

template <int N, int ACT>
int launch_t(const float* x, float* y, int B, int C, const Affine& af, cudaStream_t st) {
    using K = Cfg<N>;
    const int fr = ensure_filters<N>(st);
    if (fr != 0) return fr;
    auto kern = fact_tc_kernel<N, ACT>;
    // the attribute belongs to the (function, device) pair: set it on every launch (a host-side no-op after the first)
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    const int units = B * (C / K::CG);
    // AFLDM_FACT_TC_GRID=1: one CTA per unit (the hardware scheduler balances the tail); default: persistent CTAs
    static const int per_unit = getenv("AFLDM_FACT_TC_GRID") ? atoi(getenv("AFLDM_FACT_TC_GRID")) : 0;
    const int cap = per_unit ? units : 2 * sm_count_of_current_device();
    const int grid = units < cap ? units : cap;
    launch_k(kern, dim3(grid), dim3(FT_THREADS), K::SMEM_BYTES, st, x, y, B, C, af);
    return launched();
}

}  // namespace

bool fact_tc_enabled(int n) {
    // AFLDM_FACT_TC: 0 (default) = the mma.sync / FMA kernels of resample.cu - measured faster at n = 16 / 32 on B200
    // (profiles/r02_fact_tc.md); 1 = tcgen05 kernel for n = 16 and 32; 16 / 32 = only that plane size.
    static const int mode = getenv("AFLDM_FACT_TC") ? atoi(getenv("AFLDM_FACT_TC")) : 0;
    if (mode == 0) return false;
    if (mode == 32) return n == 32;
    if (mode == 16) return n == 16;
    return n == 32 || n == 16;
}

int fact_tc_launch(int n, int act, const float* x, float* y, int B, int C, const Affine& af, cudaStream_t st) {
    if (n != 32 && n != 16) return AFLDM_E_NOKERNEL;
    const int cg = 256 / n;
    if (C % cg != 0) return AFLDM_E_NOKERNEL;
    // 16-byte vector access on every source / destination: channel offsets are multiples of 8 already
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    if (!al16(x) || !al16(y)) return AFLDM_E_NOKERNEL;
    if (af.x2 != nullptr) {
        if (!al16(af.x2) || af.xCa % cg != 0 || (C - af.xCa) % cg != 0) return AFLDM_E_NOKERNEL;
    }
    if (act == AFLDM_ACT_SILU) {
        if (n == 32) return launch_t<32, AFLDM_ACT_SILU>(x, y, B, C, af, st);
        return launch_t<16, AFLDM_ACT_SILU>(x, y, B, C, af, st);
    }
    if (n == 32) return launch_t<32, AFLDM_ACT_IDENTITY>(x, y, B, C, af, st);
    return launch_t<16, AFLDM_ACT_IDENTITY>(x, y, B, C, af, st);
}

}  // namespace afldm
