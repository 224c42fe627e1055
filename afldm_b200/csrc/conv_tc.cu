// tcgen05 implicit-GEMM convolution (k = 1 / 3, stride 1, "same" padding) for sm_100a:
// TF32 or fp16 operands from shared memory, fp32 accumulators in TMEM, operands staged by TMA.
//
// GEMM view: D[M = B*H*W pixels][N = Cout] = A[M][K] * W[N][K]^T, K = taps * Cin, k = tap*Cin + ci.
//
//  * A (im2col) is never materialised.  The NHWC activation is one 4-D TMA tensor {C, W, H, B};
//    the 128 output pixels of a CTA tile are a box {32 ch, BW, BH, BB} of it, and filter tap
//    (kh, kw) is the SAME box shifted by (kw-1, kh-1) - TMA zero-fills whatever falls outside
//    the image, which is exactly the conv's zero padding.  Each box lands in shared memory as
//    128 rows x 128 B (K-major, 128B-swizzled), the canonical tcgen05 operand layout.
//  * W is a 2-D TMA tensor {K, Cout} (weights packed [Cout][tap][Cin]); box {32, BN}.
//  * One CTA = one 128 x BN output tile (BN in 32..256): warp 0 issues TMA into an
//    N-stage mbarrier ring, warp 1 issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8; four per
//    stage) and commits each stage back to the producer, warps 2-5 read the accumulator with
//    tcgen05.ld (one TMEM lane = one output pixel per thread) and apply the fused epilogue
//    (+bias +time-embedding row +residual) straight to global memory.
//  * Small-M layers (8x8 ... 2x2 levels, weight-stream bound) are split along K over
//    blockIdx.z; partial tiles go to the caller workspace and are reduced deterministically.
//
// Numeric class: TF32 (10-bit mantissa) products, fp32 accumulation - the same class as the
// reference's default cuDNN path (torch.backends.cudnn.allow_tf32 = True).
//
// F16 instantiation (x_half): activations and packed weights arrive as IEEE fp16 (the same 11 significant bits, stored
// by the producing kernel - DESIGN.md section 2); a 128-byte stage row then holds 64 channels, tcgen05.mma.kind::f16
// retires K = 16 per instruction, and everything measured in BYTES (ring, swizzle atoms, descriptors, halo boxes) is
// unchanged: the same kernel runs half the stages per layer on half the operand bytes.  Accumulators, epilogue
// and outputs stay fp32.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "conv.cuh"

namespace afldm {
namespace {

constexpr int TBM = 128;           // output pixels per tile (UMMA M)
constexpr int ROW_BYTES = 128;     // K bytes per stage and tile row = one swizzle row: 32 tf32 or 64 fp16 elements
constexpr int MMAS_PER_STAGE = 4;  // 32 B of K per tcgen05.mma: K = 8 (kind::tf32) or K = 16 (kind::f16)
constexpr int A_STAGE_BYTES = TBM * ROW_BYTES;   // 16 KB
// K elements per stage: fp16 operands (F16) pack twice the K into the same bytes, and kind::f16 retires twice the
// K per instruction - the same ring, descriptors and byte counts serve both, at half the stages per layer.
__host__ __device__ constexpr int tbk(bool f16) { return f16 ? 64 : 32; }
// All shared memory is dynamic: [stage ring | epilogue staging | mbarriers + TMEM slot].  (The kernel is
// persistent, so the epilogue of one tile overlaps the ring traffic of the next: staging cannot alias the ring.)
constexpr int EPI_STAGE_BYTES = 4 * 32 * 36 * 4;                 // 4 epilogue warps x 32 rows x (32 + 4) floats
constexpr int GN_STAGE_BYTES = 4 * 192 * 8;                      // 4 epilogue warps x 192 columns x float2
constexpr int STAGING_BYTES = EPI_STAGE_BYTES + GN_STAGE_BYTES;  // epilogue staging (aliases the ring when a CTA runs one item)
constexpr int BARRIER_BYTES = 256;                               // mbarriers + TMEM slot
constexpr int SMEM_ONE_PER_SM = 232448 - 1024 - BARRIER_BYTES;       // ring (+ staging), one CTA per SM
constexpr int SMEM_TWO_PER_SM = (232448 / 2) - 1024 - BARRIER_BYTES; // ring (+ staging), two CTAs per SM
constexpr int MAX_STAGES = 8;
constexpr int EPI_PITCH = 36;
constexpr int TC_THREADS = 192;    // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2..5 epilogue

// ------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// ---- CTA-pair (cta_group::2) variants: one MMA spans two SMs (M = 256), each CTA stages its own 128 rows
// of A and HALF of the B tile; TMA completions and MMA commits are routed to the leader CTA's barriers.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                             int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(bar), "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {     // shared::cta -> shared::cluster of `rank`
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ float4 ld_cluster_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start address >> 4 in [0,14), LBO (ignored for swizzled K-major) = 1 in [16,30),
// SBO = 1024 B (8 rows x 128 B) >> 4 in [32,46), version = 1 in [46,48), layout SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}

struct TcArgs {
    const float* bias;
    const float* row_add;
    const float* residual;
    float* y;
    float* ws;
    int row_add_pitch, res_pitch, y_pitch;
    int M, Cout, HW;
    int taps, ks, cin_chunks;       // K iteration space: taps x cin_chunks stages of 32
    int iters_per_split;
    int BN, stages, tmem_cols;
    int W, H;                       // image size
    int BW, BH;                     // box geometry (BB implied)
    int vec_ok;                     // all epilogue pointers / pitches allow float4 access
    float* gn_partial;              // [B][gn_slots][Cout] float2 GroupNorm partial sums of the output | NULL
    int gn_slots;
    int mtiles, ntiles, splitk;     // work items = mtiles * ntiles * splitk (pairs: mtiles / 2 M-tile pairs)
    int nacc;                       // TMEM accumulator buffers (2: epilogue of item i overlaps main loop of i+1)
    int alias_staging;              // one item per CTA: the epilogue staging reuses the (then idle) stage ring
    int halo;                       // 3x3 halo mode: one A box {32 ch, BW, BH + 2} per (kw, channel chunk) serves the 3 kh taps
    int y_half;                     // y is fp16 (y_pitch in halves): QKV projections feeding afldm_attention_f16
    int cin1_chunks;                // channel chunks [0, cin1_chunks) come from map_a, the rest from map_a2 (un-materialised concat)
    int csk;                        // cluster split-K: the `csk` CTAs of a cluster hold the K splits of one tile and reduce them
                                    // through distributed shared memory (no partials in global memory, no second launch)
    int csk_seg;                    // rows per GroupNorm slot in that mode: min(128 / csk, H * W)
};

// One lane of a converged warp (elect.sync); the same lane every time for the full mask.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar) {      // bar: shared::cluster address
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Persistent, warp-specialised implicit-GEMM kernel.  Every CTA (TWO: every CTA pair) walks the work items
// item = first, first + stride, ... ; an item is one 128 x BN (pair: 256 x BN) output tile of one K split.
//   warp 0      TMA producer: A box per filter tap + B box into the stage ring (runs ahead across items)
//   warp 1      MMA issuer (pair: only in the even CTA): tcgen05.mma into TMEM accumulator buffer (item % nacc)
//   warps 2..5  epilogue: tcgen05.ld -> smem transpose -> coalesced float4 stores (+bias +temb row +residual),
//               GroupNorm partial sums; releases the accumulator buffer back to the MMA warp
// With two accumulator buffers the epilogue of item i runs under the main loop of item i + 1, and the TMEM
// allocation, barrier init and tensor-map fetch are paid once per CTA instead of once per tile.
// TWO = true: tcgen05.mma.cta_group::2 (M = 256) over a cluster of two CTAs; each CTA stages its own 128 rows
// of A and BN/2 rows of B (16 KB + BN*64 B per stage instead of 16 KB + BN*128 B): the SM's operand ingest,
// the measured bound of the main loop, buys up to 2x the FLOPs.
template <bool TWO, bool HALO, bool F16>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a2,
               const __grid_constant__ CUtensorMap map_b, const TcArgs a) {
    pdl_trigger();                                           // dependents may start their own prologue right away
    extern __shared__ __align__(1024) uint8_t smem_raw[];   // no static smem: the ring starts 1024-aligned
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
    const uint32_t smem_base = smem_u32(smem_raw);
    if ((smem_base & 1023u) != 0u) __trap();                 // swizzle-128B atoms need 1024 B alignment
    constexpr int TBK = tbk(F16);                            // K elements (channels) per stage
    const int b_stage_bytes = (TWO ? a.BN / 2 : a.BN) * ROW_BYTES;
    // Halo mode (3x3, tile = BH >= 2 whole image rows): a stage holds the box of BH + 2 rows shifted by kw - 1
    // and the three weight tiles of taps (kh, kw), kh = 0..2.  Tap kh reads the SAME box BW rows further down
    // (a multiple of the 1024 B swizzle atom, so only the descriptor start address moves): the SM ingests
    // 3 (BH + 2) / (9 BH) of the activation bytes of the one-box-per-tap scheme - the measured bound of the
    // main loop is operand bytes staged per SM (profiles/r01_conv_notes.md).
    const int a_bytes = HALO ? (a.BH + 2) * a.BW * ROW_BYTES : A_STAGE_BYTES;
    constexpr int nb = HALO ? 3 : 1;
    const int stage_bytes = a_bytes + nb * b_stage_bytes;
    // layout: [stage ring][epilogue staging 4 x 32 x 36 floats][GroupNorm staging 4 x 192 float2][barriers]
    uint8_t* const ring_end = smem_raw + (size_t)a.stages * stage_bytes;
    uint8_t* const staging = a.alias_staging ? smem_raw : ring_end;
    float (*epi_stage)[32 * EPI_PITCH] = reinterpret_cast<float (*)[32 * EPI_PITCH]>(staging);
    float2 (*gn_stage)[192] = reinterpret_cast<float2 (*)[192]>(staging + EPI_STAGE_BYTES);
    uint64_t* const full_bar = reinterpret_cast<uint64_t*>(ring_end + (a.alias_staging ? 0 : STAGING_BYTES));
    uint64_t* const empty_bar = full_bar + MAX_STAGES;
    uint64_t* const acc_full = empty_bar + MAX_STAGES;       // [2]
    uint64_t* const acc_empty = acc_full + 2;                // [2]
    uint32_t* const tmem_slot_p = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* const csk_tile = reinterpret_cast<float*>(smem_raw + STAGING_BYTES);   // cluster split-K partial tile (aliases the idle ring)

    const int total_iters = a.taps * a.cin_chunks;
    const int csk = TWO ? 0 : a.csk;
    const uint32_t crank = (TWO || csk > 0) ? cluster_rank() : 0u;
    const bool mma_leader = !TWO || crank == 0u;
    // work items of this CTA (pair): first, first + stride, ...
    // cluster split-K: one tile per cluster, the CTA's cluster rank is its K split
    const int first = csk > 0 ? (int)blockIdx.x / csk : (TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x);
    const int n_items = csk > 0 ? first + 1 : (TWO ? a.mtiles / 2 : a.mtiles) * a.ntiles * a.splitk;
    const int stride = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&acc_full[i]), 1);
            mbar_init(smem_u32(&acc_empty[i]), TWO ? 8 : 4);     // one arrival per epilogue warp (of both CTAs)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (TWO) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             smem_u32(tmem_slot_p)), "r"((uint32_t)a.tmem_cols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             smem_u32(tmem_slot_p)), "r"((uint32_t)a.tmem_cols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a2)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if constexpr (TWO) cluster_sync_all();          // the peer's barriers exist before anyone signals them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_p;
    // Everything above touched only this CTA's shared / tensor memory: it may overlap the tail of the
    // previous kernel on the stream.  From here on global memory is read.
    pdl_wait();

    // item -> (M tile, N tile, K split).  M fastest: neighbouring CTAs (and the two CTAs sharing an SM) work on
    // neighbouring pixel tiles with the same weight tile, as in the 2-D grid this kernel grew out of.
    auto decode = [&](int item, int& mt, int& nt, int& z) {
        const int mrow = TWO ? a.mtiles / 2 : a.mtiles;
        const int mi = item % mrow;
        const int r = item / mrow;
        nt = r % a.ntiles;
        z = csk > 0 ? (int)crank : r / a.ntiles;
        mt = TWO ? 2 * mi + (int)crank : mi;
    };

    // The producer and the MMA issuer run their loops with the WHOLE warp (warp-uniform control flow) and elect one
    // lane only around the TMA / tcgen05 instructions.  Under `if (lane == 0)` the compiler cannot keep descriptors,
    // barrier addresses and coordinates in uniform registers and wraps every UTCHMMA / UTMALDG in an
    // ELECT + 5 x R2UR.BROADCAST + BRA.U.ANY waterfall (~15 instructions per 32..48-cycle MMA, ncu / SASS of the
    // r01 kernel); ring position and tap / chunk counters are carried incrementally instead of divided out.
    if (warp == 0) {
        // ================= TMA producer =================
        {
            const int pad = a.ks >> 1;
            int s = 0;                                      // ring slot and its phase, continue across items
            uint32_t ph = 0;
            for (int item = first; item < n_items; item += stride) {
                int mt, nt, z;
                decode(item, mt, nt, z);
                int w0, h0, b0;                             // tile origin in (w, h, b)
                if (a.W >= TBM) {
                    const int per_row = a.W / TBM;
                    const int row = mt / per_row;
                    w0 = (mt % per_row) * TBM;
                    h0 = row % a.H;
                    b0 = row / a.H;
                } else if (a.BH == a.H) {
                    w0 = 0; h0 = 0;
                    b0 = mt * (TBM / (a.W * a.H));
                } else {
                    const int per_img = a.H / a.BH;
                    w0 = 0;
                    h0 = (mt % per_img) * a.BH;
                    b0 = mt / per_img;
                }
                const int n0 = nt * a.BN;
                const int it_beg = z * a.iters_per_split;
                const int it_end = min(total_iters, it_beg + a.iters_per_split);
                int tap = it_beg / a.cin_chunks;
                int chunk = it_beg - tap * a.cin_chunks;
                for (int it = it_beg; it < it_end; ++it) {
                    mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
                    const bool second = chunk >= a.cin1_chunks;
                    const CUtensorMap* const ma = second ? &map_a2 : &map_a;
                    const int c0 = (second ? chunk - a.cin1_chunks : chunk) * TBK;     // channel inside its source
                    const uint32_t fb = smem_u32(&full_bar[s]);
                    const uint32_t sa = smem_base + s * stage_bytes;
                    if (elect_one()) {
                        if constexpr (HALO) {
                            // `tap` is kw here; weights are packed [Cout][kh * 3 + kw][Cin]
                            const int nb0 = n0 + (TWO ? (int)crank * (a.BN / 2) : 0);
                            const int kb = (tap * a.cin_chunks + chunk) * TBK;       // + kh * 3 * Cin
                            const int kstep = 3 * a.cin_chunks * TBK;
                            if constexpr (TWO) {
                                if (mma_leader) mbar_expect_tx(fb, 2u * (uint32_t)stage_bytes);
                                tma2_load_4d(sa, ma, fb, c0, w0 + tap - 1, h0 - 1, b0);
#pragma unroll
                                for (int kh = 0; kh < 3; ++kh)
                                    tma2_load_2d(sa + a_bytes + kh * b_stage_bytes, &map_b, fb, kb + kh * kstep, nb0);
                            } else {
                                mbar_expect_tx(fb, (uint32_t)stage_bytes);
                                tma_load_4d(sa, ma, fb, c0, w0 + tap - 1, h0 - 1, b0);
#pragma unroll
                                for (int kh = 0; kh < 3; ++kh)
                                    tma_load_2d(sa + a_bytes + kh * b_stage_bytes, &map_b, fb, kb + kh * kstep, nb0);
                            }
                        } else {
                            const int th = tap / a.ks;
                            const int dh = th - pad, dw = tap - th * a.ks - pad;
                            if constexpr (TWO) {
                                // both CTAs' bytes complete on the leader's barrier; only the leader arms it
                                if (mma_leader) mbar_expect_tx(fb, 2u * (uint32_t)stage_bytes);
                                tma2_load_4d(sa, ma, fb, c0, w0 + dw, h0 + dh, b0);
                                tma2_load_2d(sa + A_STAGE_BYTES, &map_b, fb, it * TBK, n0 + (int)crank * (a.BN / 2));
                            } else {
                                mbar_expect_tx(fb, (uint32_t)stage_bytes);
                                tma_load_4d(sa, ma, fb, c0, w0 + dw, h0 + dh, b0);
                                tma_load_2d(sa + A_STAGE_BYTES, &map_b, fb, it * TBK, n0);
                            }
                        }
                    }
                    __syncwarp();
                    if (++chunk == a.cin_chunks) { chunk = 0; ++tap; }
                    if (++s == a.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (mma_leader) {
            // instruction descriptor: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10) or F16 (0, 0), K-major both,
            // N >> 3 at bit 17, M >> 4 at bit 24 (M = 256 for the CTA pair)
            constexpr uint32_t ab_fmt = F16 ? 0u : ((2u << 7) | (2u << 10));
            const uint32_t idesc = (1u << 4) | ab_fmt | ((uint32_t)(a.BN >> 3) << 17) |
                                   ((uint32_t)((TWO ? 2 * TBM : TBM) >> 4) << 24);
            int s = 0;
            uint32_t ph = 0;
            int ab = 0;                                     // accumulator buffer of this item and its phase
            uint32_t aph = 0;
            for (int item = first; item < n_items; item += stride) {
                int mt, nt, z;
                decode(item, mt, nt, z);
                const int it_beg = z * a.iters_per_split;
                const int n_it = min(total_iters, it_beg + a.iters_per_split) - it_beg;
                mbar_wait(smem_u32(&acc_empty[ab]), aph ^ 1u);      // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmem_base + (uint32_t)(ab * a.BN);
                for (int i = 0; i < n_it; ++i) {
                    mbar_wait(smem_u32(&full_bar[s]), ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = smem_base + s * stage_bytes;
                    if (elect_one()) {
#pragma unroll
                        for (int kh = 0; kh < nb; ++kh) {
                            // halo mode: tap kh = the same box, BW pixel rows (BW * 128 B) further down
                            const uint64_t da = make_desc(sa + (HALO ? (uint32_t)(kh * a.BW * ROW_BYTES) : 0u));
                            const uint64_t db = make_desc(sa + (uint32_t)a_bytes + (uint32_t)(kh * b_stage_bytes));
#pragma unroll
                            for (int k = 0; k < MMAS_PER_STAGE; ++k) {
                                // advance 8 fp32 / 16 fp16 = 32 B inside the 128 B swizzle row: +2 in the (>>4) address field
                                const uint32_t acc = (i | kh | k) != 0 ? 1u : 0u;
                                if constexpr (F16) {
                                    if constexpr (TWO) umma2_f16(tacc, da + 2 * k, db + 2 * k, idesc, acc);
                                    else umma_f16(tacc, da + 2 * k, db + 2 * k, idesc, acc);
                                } else {
                                    if constexpr (TWO) umma2_tf32(tacc, da + 2 * k, db + 2 * k, idesc, acc);
                                    else umma_tf32(tacc, da + 2 * k, db + 2 * k, idesc, acc);
                                }
                            }
                        }
                        // frees the stage (in both CTAs of a pair) when these MMAs retire
                        if constexpr (TWO) umma2_commit_mc(smem_u32(&empty_bar[s]), 3);
                        else umma_commit(smem_u32(&empty_bar[s]));
                        // accumulator complete (in both CTAs of a pair); same lane as the MMAs (commit tracks the
                        // issuing thread's tcgen05 operations)
                        if (i == n_it - 1) {
                            if constexpr (TWO) umma2_commit_mc(smem_u32(&acc_full[ab]), 3);
                            else umma_commit(smem_u32(&acc_full[ab]));
                        }
                    }
                    __syncwarp();
                    if (++s == a.stages) { s = 0; ph ^= 1u; }
                }
                if (++ab == a.nacc) { ab = 0; aph ^= 1u; }
            }
        }
    } else {
        // ================= epilogue: warps 2..5, TMEM lane quarter = warp % 4 =================
        const int q = warp & 3;
        const bool split = a.splitk > 1 && csk == 0;
        float* stg = epi_stage[q];
        const int cq = lane & 7, rsub = lane >> 3;      // coalesced phase: 8 lanes x float4 per row, 4 rows per pass
        int j = 0;
        for (int item = first; item < n_items; item += stride, ++j) {
            int mt, nt, z;
            decode(item, mt, nt, z);
            const int m0 = mt * TBM, n0 = nt * a.BN;
            const int ab = j % a.nacc;
            const uint32_t aph = (uint32_t)(j / a.nacc) & 1u;
            mbar_wait(smem_u32(&acc_full[ab]), aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * a.BN);
            for (int c = 0; c < a.BN; c += 32) {
                // Everything the fused epilogue adds (bias + time-embedding row + residual) is fetched BEFORE the
                // accumulator read, with clamped addresses, so the global-load latency hides under the TMEM load and
                // the staging transpose instead of serialising the eight store passes (K-light 1x1 layers are
                // epilogue-bound: profiles/r01_epilogue_notes.md).
                const bool fast = !split && a.vec_ok && !a.y_half && csk == 0;
                float4 addv[8];
                if (fast) {
                    const int nc = min(n0 + c + cq * 4, a.Cout - 4);
                    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (a.bias != nullptr) bv = *reinterpret_cast<const float4*>(a.bias + nc);
#pragma unroll
                    for (int pass = 0; pass < 8; ++pass) {
                        const int mmc = min(m0 + q * 32 + pass * 4 + rsub, a.M - 1);
                        float4 t = bv;
                        if (a.row_add != nullptr) {
                            const float4 u = *reinterpret_cast<const float4*>(a.row_add + (size_t)(mmc / a.HW) * a.row_add_pitch + nc);
                            t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
                        }
                        if (a.residual != nullptr) {
                            const float4 u = *reinterpret_cast<const float4*>(a.residual + (size_t)mmc * a.res_pitch + nc);
                            t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
                        }
                        addv[pass] = t;
                    }
                }
                uint32_t r[32];
                tmem_ld32(tacc + (uint32_t)c, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c + 32 >= a.BN) {
                    // last read of this accumulator: hand it back to the MMA warp before the stores go out
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (TWO) mbar_arrive_cluster(smem_u32(&acc_empty[ab]) & PEER_BIT_MASK);
                        else mbar_arrive_local(smem_u32(&acc_empty[ab]));
                    }
                }
                if (csk > 0) {
                    // cluster split-K: the partial tile stays on chip, [128][BN + 4] fp32 behind the staging area
                    float4* prow = reinterpret_cast<float4*>(csk_tile + (size_t)(q * 32 + lane) * (a.BN + 4) + c);
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj)
                        if (c + 4 * jj < a.BN)
                            prow[jj] = make_float4(__uint_as_float(r[4 * jj]), __uint_as_float(r[4 * jj + 1]),
                                                   __uint_as_float(r[4 * jj + 2]), __uint_as_float(r[4 * jj + 3]));
                    continue;
                }
                // each thread owns one accumulator row: park it in the staging tile ...
                float4* srow = reinterpret_cast<float4*>(stg + lane * EPI_PITCH);
#pragma unroll
                for (int jj = 0; jj < 8; ++jj)
                    srow[jj] = make_float4(__uint_as_float(r[4 * jj]), __uint_as_float(r[4 * jj + 1]),
                                           __uint_as_float(r[4 * jj + 2]), __uint_as_float(r[4 * jj + 3]));
                __syncwarp();
                if (a.y_half) {
                    // fp16 output (no residual / GroupNorm sums on this path): 4 lanes x 8 columns per row, 8 rows per
                    // pass, one 16-byte store per lane
                    const int cq8 = lane & 3, rs8 = lane >> 2;
                    const int n8 = n0 + c + cq8 * 8;
#pragma unroll
                    for (int pass = 0; pass < 4; ++pass) {
                        const int rr = pass * 8 + rs8;
                        const int mm = m0 + q * 32 + rr;
                        if (mm >= a.M || n8 >= a.Cout || c + cq8 * 8 >= a.BN) continue;
                        float4 v0 = *reinterpret_cast<const float4*>(stg + rr * EPI_PITCH + cq8 * 8);
                        float4 v1 = *reinterpret_cast<const float4*>(stg + rr * EPI_PITCH + cq8 * 8 + 4);
                        if (a.bias != nullptr) {
                            const float4 t0 = *reinterpret_cast<const float4*>(a.bias + n8);
                            const float4 t1 = *reinterpret_cast<const float4*>(a.bias + n8 + 4);
                            v0.x += t0.x; v0.y += t0.y; v0.z += t0.z; v0.w += t0.w;
                            v1.x += t1.x; v1.y += t1.y; v1.z += t1.z; v1.w += t1.w;
                        }
                        if (a.row_add != nullptr) {
                            const float* rp = a.row_add + (size_t)(mm / a.HW) * a.row_add_pitch + n8;
                            const float4 t0 = *reinterpret_cast<const float4*>(rp);
                            const float4 t1 = *reinterpret_cast<const float4*>(rp + 4);
                            v0.x += t0.x; v0.y += t0.y; v0.z += t0.z; v0.w += t0.w;
                            v1.x += t1.x; v1.y += t1.y; v1.z += t1.z; v1.w += t1.w;
                        }
                        const __half2 h0 = __floats2half2_rn(v0.x, v0.y), h1 = __floats2half2_rn(v0.z, v0.w);
                        const __half2 h2 = __floats2half2_rn(v1.x, v1.y), h3 = __floats2half2_rn(v1.z, v1.w);
                        uint4 pk;
                        pk.x = *reinterpret_cast<const uint32_t*>(&h0);
                        pk.y = *reinterpret_cast<const uint32_t*>(&h1);
                        pk.z = *reinterpret_cast<const uint32_t*>(&h2);
                        pk.w = *reinterpret_cast<const uint32_t*>(&h3);
                        *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(a.y) + (size_t)mm * a.y_pitch + n8) = pk;
                    }
                    __syncwarp();
                    continue;
                }
                // ... and write it out 4 rows x 128 contiguous bytes per instruction
                const int n = n0 + c + cq * 4;
                float gs[4] = {0.f, 0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int pass = 0; pass < 8; ++pass) {
                    const int rr = pass * 4 + rsub;
                    const int mm = m0 + q * 32 + rr;
                    if (mm >= a.M || n >= a.Cout || c + cq * 4 >= a.BN) continue;
                    float4 v = *reinterpret_cast<const float4*>(stg + rr * EPI_PITCH + cq * 4);
                    if (split) {
                        *reinterpret_cast<float4*>(a.ws + ((size_t)z * a.M + mm) * a.Cout + n) = v;
                    } else if (a.vec_ok) {
                        const float4 t = addv[pass];
                        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
                        *reinterpret_cast<float4*>(a.y + (size_t)mm * a.y_pitch + n) = v;
                        gs[0] += v.x; gs[1] += v.y; gs[2] += v.z; gs[3] += v.w;
                        gq[0] = fmaf(v.x, v.x, gq[0]); gq[1] = fmaf(v.y, v.y, gq[1]);
                        gq[2] = fmaf(v.z, v.z, gq[2]); gq[3] = fmaf(v.w, v.w, gq[3]);
                    } else {
                        float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            if (n + jj >= a.Cout) break;
                            float t = e[jj];
                            if (a.bias != nullptr) t += a.bias[n + jj];
                            if (a.row_add != nullptr) t += a.row_add[(size_t)(mm / a.HW) * a.row_add_pitch + n + jj];
                            if (a.residual != nullptr) t += a.residual[(size_t)mm * a.res_pitch + n + jj];
                            a.y[(size_t)mm * a.y_pitch + n + jj] = t;
                        }
                    }
                }
                if (a.gn_partial != nullptr) {
                    // rows live in lanes rsub = 0..3 of the same column quad: fold them, lanes 0..7 keep the totals
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        gs[jj] += __shfl_xor_sync(0xffffffffu, gs[jj], 8);
                        gq[jj] += __shfl_xor_sync(0xffffffffu, gq[jj], 8);
                        gs[jj] += __shfl_xor_sync(0xffffffffu, gs[jj], 16);
                        gq[jj] += __shfl_xor_sync(0xffffffffu, gq[jj], 16);
                    }
                    if (rsub == 0) {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) gn_stage[q][c + cq * 4 + jj] = make_float2(gs[jj], gq[jj]);
                    }
                }
                __syncwarp();
            }
            if (a.gn_partial != nullptr && csk == 0) {
                // combine the four row quarters (fixed order) and publish one partial per (tile, channel)
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int et = threadIdx.x - 64;            // 0..127 over the epilogue warps
                const int bimg = m0 / a.HW;
                const int slot = (m0 - bimg * a.HW) / TBM;
                for (int col = et; col < a.BN; col += 128) {
                    const int n = n0 + col;
                    if (n >= a.Cout) continue;
                    const float2 p0 = gn_stage[0][col], p1 = gn_stage[1][col], p2 = gn_stage[2][col], p3 = gn_stage[3][col];
                    float2* gp = reinterpret_cast<float2*>(a.gn_partial);
                    if (a.HW >= TBM) {
                        gp[((size_t)bimg * a.gn_slots + slot) * a.Cout + n] =
                            make_float2(((p0.x + p1.x) + p2.x) + p3.x, ((p0.y + p1.y) + p2.y) + p3.y);
                    } else if (a.HW == 64) {
                        // the tile holds two whole 8x8 images (two row quarters each): one slot per image
                        gp[(size_t)bimg * a.Cout + n] = make_float2(p0.x + p1.x, p0.y + p1.y);
                        if ((bimg + 1) * a.HW < a.M) gp[(size_t)(bimg + 1) * a.Cout + n] = make_float2(p2.x + p3.x, p2.y + p3.y);
                    } else {                                     // HW == 32: one image per row quarter
                        const float2 pq[4] = {p0, p1, p2, p3};
#pragma unroll
                        for (int w4 = 0; w4 < 4; ++w4)
                            if ((bimg + w4) * a.HW < a.M) gp[(size_t)(bimg + w4) * a.Cout + n] = pq[w4];
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");   // staging is reused by the next item
            }
        }
    }

    if (csk > 0) {
        // ---- cluster split-K reduction.  Every CTA of the cluster has parked its partial 128 x BN tile in its own
        // shared memory; CTA `crank` now owns rows [crank * R, (crank + 1) * R), R = 128 / csk: it adds the csk
        // partials of those rows in rank order (deterministic) straight out of the peers' shared memory, applies the
        // fused epilogue (+bias +temb row +residual), stores y and emits the GroupNorm partial sums.
        cluster_sync_all();
        if (warp >= 2) {
            int mt, nt, z;
            decode(first, mt, nt, z);
            const int m0 = mt * TBM, n0 = nt * a.BN;
            const int et = (int)threadIdx.x - 64, tq = et >> 5, tl = et & 31;
            const int R = TBM / csk, PP = a.BN + 4, seg = a.csk_seg, nseg = R / seg;
            const uint32_t local = smem_u32(csk_tile);
            uint32_t rbase[8];
#pragma unroll
            for (int sidx = 0; sidx < 8; ++sidx) rbase[sidx] = map_to_cta(local, (uint32_t)min(sidx, csk - 1));
            float2* red = reinterpret_cast<float2*>(smem_raw);           // [4 row lanes][<= 4 segments][BN], in the staging area
            for (int c4 = tl; c4 * 4 < a.BN; c4 += 32) {
                const int n = n0 + c4 * 4;
                const bool colok = n < a.Cout;
                float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (colok && a.bias != nullptr) bv = *reinterpret_cast<const float4*>(a.bias + n);
                for (int sg = 0; sg < nseg; ++sg) {
                    float gs[4] = {0.f, 0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int r = sg * seg + tq; r < (sg + 1) * seg; r += 4) {
                        const int row = (int)crank * R + r;
                        const int mm = m0 + row;
                        const uint32_t off = (uint32_t)((row * PP + c4 * 4) * 4);
                        float4 v = ld_cluster_f4(rbase[0] + off);
#pragma unroll
                        for (int sidx = 1; sidx < 8; ++sidx) {
                            if (sidx < csk) {
                                const float4 u = ld_cluster_f4(rbase[sidx] + off);
                                v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
                            }
                        }
                        if (mm < a.M && colok) {
                            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
                            if (a.row_add != nullptr) {
                                const float4 t = *reinterpret_cast<const float4*>(a.row_add + (size_t)(mm / a.HW) * a.row_add_pitch + n);
                                v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
                            }
                            if (a.residual != nullptr) {
                                const float4 t = *reinterpret_cast<const float4*>(a.residual + (size_t)mm * a.res_pitch + n);
                                v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
                            }
                            *reinterpret_cast<float4*>(a.y + (size_t)mm * a.y_pitch + n) = v;
                            gs[0] += v.x; gs[1] += v.y; gs[2] += v.z; gs[3] += v.w;
                            gq[0] = fmaf(v.x, v.x, gq[0]); gq[1] = fmaf(v.y, v.y, gq[1]);
                            gq[2] = fmaf(v.z, v.z, gq[2]); gq[3] = fmaf(v.w, v.w, gq[3]);
                        }
                    }
                    if (a.gn_partial != nullptr) {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) red[(tq * 4 + sg) * a.BN + c4 * 4 + jj] = make_float2(gs[jj], gq[jj]);
                    }
                }
            }
            if (a.gn_partial != nullptr) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int idx = et; idx < nseg * a.BN; idx += 128) {
                    const int sg = idx / a.BN, col = idx - sg * a.BN;
                    const int n = n0 + col;
                    const int mm = m0 + (int)crank * R + sg * seg;          // first row of this GroupNorm slot
                    if (n >= a.Cout || mm >= a.M) continue;
                    const float2 p0 = red[(0 * 4 + sg) * a.BN + col], p1 = red[(1 * 4 + sg) * a.BN + col];
                    const float2 p2 = red[(2 * 4 + sg) * a.BN + col], p3 = red[(3 * 4 + sg) * a.BN + col];
                    const int bimg = mm / a.HW, slot = (mm - bimg * a.HW) / seg;
                    reinterpret_cast<float2*>(a.gn_partial)[((size_t)bimg * a.gn_slots + slot) * a.Cout + n] =
                        make_float2(((p0.x + p1.x) + p2.x) + p3.x, ((p0.y + p1.y) + p2.y) + p3.y);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (TWO || csk > 0) cluster_sync_all();         // no CTA leaves while a peer may still signal it / read its shared memory
    if (warp == 1) {
        if constexpr (TWO)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols)
                         : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols)
                         : "memory");
    }
}

// ------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

int num_sms() {
    static const int n = [] {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
            v = 148;     // B200
        return v;
    }();
    return n;
}

struct TcPlan {
    bool ok;
    int M, mtiles, ntiles, BN, stages, tmem_cols, total_iters, splitk, iters_per_split, BW, BH, BB;
    int two;                                            // CTA-pair (cta_group::2) mode
    int ctas_per_sm, grid_ctas, nacc, alias_staging;    // persistent grid and TMEM accumulator buffers
    int halo;                                           // 3x3 halo mode (one A box per kw and channel chunk)
    int csk, csk_seg;                                   // cluster split-K size (0 = off) and rows per GroupNorm slot
    size_t smem_bytes;
};

TcPlan tc_plan(int B, int H, int W, int Cin, int Cout, int ks, bool f16 = false) {
    TcPlan p{};
    p.ok = false;
    const int TBK = tbk(f16);            // channels per 128-byte stage row
    // Cout < 16 (conv_out: C -> 4 / 3) runs as one BN = 16 tile: the weight box rows past Cout are
    // zero-filled by TMA and the epilogue masks them.
    if (Cin % TBK != 0 || (Cout >= 16 && Cout % 16 != 0) || !is_pow2(W) || !is_pow2(H)) return p;
    if (W > TBM && W % TBM != 0) return p;
    p.M = B * H * W;
    p.mtiles = ceil_div(p.M, TBM);
    if (W >= TBM) { p.BW = TBM; p.BH = 1; p.BB = 1; }
    else {
        p.BW = W;
        p.BH = std::min(H, TBM / W);
        p.BB = TBM / (W * p.BH);
    }
    if (p.BB > 1 && p.BH != H) return p;
    if (p.BB > 256) return p;
    p.total_iters = ks * ks * (Cin / TBK);

    // Tile width: largest BN (multiple of 32, <= 256) whose grid still fills the chip; small problems take
    // BN = 64 and split K instead.  AFLDM_TC_BN / AFLDM_TC_SPLITK override the heuristic (tuning runs).
    static const int force_bn = getenv("AFLDM_TC_BN") ? atoi(getenv("AFLDM_TC_BN")) : 0;
    static const int force_split = getenv("AFLDM_TC_SPLITK") ? atoi(getenv("AFLDM_TC_SPLITK")) : 0;
    static const int force_stages = getenv("AFLDM_TC_STAGES") ? atoi(getenv("AFLDM_TC_STAGES")) : 0;
    // Tile width, from a sweep on B200 (profiles/r01_conv_bn_sweep.md).  The main loop is bound by
    // L2 -> SMEM operand traffic and by per-CTA prologue / epilogue latency, so the best width is the one
    // that puts two CTAs on every SM (one CTA's epilogue hides behind the other's main loop): 96 columns
    // when the grid is large, 64 for the 16x16 level and for 1x1 layers, 96 again for K-heavy small-M layers.
    int best;
    if (Cout < 64) best = Cout >= 32 ? 32 : 16;
    else if (Cout % 96 == 0 && (p.mtiles * (Cout / 96) >= 222 || (p.mtiles <= 8 && ks == 3))) best = 96;
    else if (Cout % 64 == 0) best = (Cout % 128 == 0 && p.mtiles * (Cout / 128) >= 296) ? 128 : 64;
    else if (Cout % 96 == 0) best = 96;
    else best = 32;
    if (force_bn > 0 && force_bn % 16 == 0 && force_bn <= 256 && (Cout % force_bn == 0 || force_bn <= 64)) best = force_bn;
    if (best > Cout) best = Cout;        // Cout in {16, 32, 48}: one narrow tile
    if (best < 16) best = 16;
    p.BN = best;
    p.ntiles = ceil_div(Cout, p.BN);
    const int tiles = p.mtiles * p.ntiles;
    int s = 1;
    if (tiles < 100) {
        // Split K until at most one wave of CTAs exists, but keep >= 8 stages per split, and do not split K-light
        // 1x1 layers that already have >= 32 tiles: their whole main loop is <= 24 stages, shorter than the
        // second launch (deterministic reduce) a split costs.  From a forced-split sweep over every small-M layer
        // of the step on B200 (profiles/r01_conv_splitk_sweep.md): -0.12 ms per step against "two waves, >= 4 stages".
        s = std::max(1, num_sms() / tiles);                // floor: tiles * s CTAs never spill into a second,
                                                           // nearly empty wave (160 CTAs on 148 SMs = 2x the time)
        // (fp16 operands: a stage carries twice the K, so ">= 4 stages" keeps the split sizes of the swept TF32 rule -
        // and with them the split-K reduce that emits the GroupNorm sums of the 4x4 / 2x2 levels)
        static const int f16_min_iters = getenv("AFLDM_TC_F16_MINIT") ? std::max(1, atoi(getenv("AFLDM_TC_F16_MINIT"))) : 4;
        s = std::min(s, std::max(1, p.total_iters / (f16 ? f16_min_iters : 8)));
        s = std::min(s, 64);
        if (ks == 1 && tiles >= 32) s = 1;
    }
    if (force_split > 0) s = std::min(force_split, p.total_iters);
    if (Cout % 4 != 0) s = 1;            // split-K partials are written as float4
    p.iters_per_split = ceil_div(p.total_iters, s);
    p.splitk = ceil_div(p.total_iters, p.iters_per_split);
    // CTA-pair mode (tcgen05.mma.cta_group::2): two consecutive M tiles share every B tile, each CTA stages
    // only half of it.  For the un-split layers with an even number of M tiles; wide tiles, since the point
    // is FLOPs per staged byte.
    static const int allow_two = getenv("AFLDM_TC_TWO") ? atoi(getenv("AFLDM_TC_TWO")) : 1;
    static const int force_bn2 = getenv("AFLDM_TC2_BN") ? atoi(getenv("AFLDM_TC2_BN")) : 0;
    p.two = 0;
    if (allow_two && p.splitk == 1 && p.mtiles % 2 == 0 && p.mtiles >= 32 && Cout >= 64 && Cout % 16 == 0) {
        // Sweep on B200 (profiles/r01_conv_notes.md): the pair wins on the K-heavy 3x3 layers (up to 1.4x) and
        // loses on K-light 1x1 layers, and the best width is the one whose grid just fits one wave at two
        // CTAs per SM (<= 296 CTAs); 1x1 layers only qualify with a wide (>= 128) tile.
        int bn2 = 0, best_ctas = 0;
        const int c2[] = {192, 128, 96, 64};
        for (int bn : c2) {
            if (Cout % bn != 0) continue;
            const int ctas = p.mtiles * (Cout / bn);
            if (ctas <= 296 && ctas > best_ctas) { best_ctas = ctas; bn2 = bn; }
        }
        if (bn2 == 0 && ks == 3) {
            for (int bn : c2)
                if (Cout % bn == 0) { bn2 = bn; break; }              // very large layer: widest tile
        }
        // 3x3 layers that run in halo mode (below): 96 columns.  With the activation bytes cut by the halo box the
        // weight half-tile dominates the ingest, and 96 beats both 64 (more weight bytes per FLOP) and 128 / 192
        // (60 KB stages, 3-deep ring) on every 16x16 / 32x32 layer of the step (profiles/r01_conv_knobs.md).
        if (ks == 3 && p.BB == 1 && p.BH >= 2 && Cout % 96 == 0) bn2 = 96;
        if (ks != 3 && bn2 < 128) bn2 = 0;
        if (force_bn2 > 0) bn2 = (Cout % force_bn2 == 0 && force_bn2 % 16 == 0 && force_bn2 <= 256) ? force_bn2 : 0;
        if (bn2 > 0) {
            p.two = 1;
            p.BN = bn2;
            p.ntiles = Cout / bn2;
        }
    }
    // Halo mode: un-split 3x3 layers whose 128-pixel tile is BH >= 2 whole rows of one image (W = 16 ... 64).
    static const int allow_halo = getenv("AFLDM_TC_HALO") ? atoi(getenv("AFLDM_TC_HALO")) : 1;
    // Not for the widest CTA-pair tiles: those layers already run tensor-bound (fewest bytes per FLOP) and the
    // 60 KB halo stages leave a 3-deep ring (sweep on B200, profiles/r01_conv_halo_sweep.md: 68.9 vs 63.2 us).
    p.halo = (allow_halo && ks == 3 && p.splitk == 1 && p.BB == 1 && p.BH >= 2 && p.BW >= 8 && p.BW * p.BH == TBM &&
              !(p.two && p.BN >= 192 && allow_halo < 2)) ? 1 : 0;
    const int b_bytes = (p.two ? p.BN / 2 : p.BN) * ROW_BYTES;
    const int stage_bytes = p.halo ? (p.BH + 2) * p.BW * ROW_BYTES + 3 * b_bytes : A_STAGE_BYTES + b_bytes;
    if (p.halo) {
        p.total_iters = 3 * (Cin / TBK);          // one stage per (kw, channel chunk): 12 MMAs
        p.iters_per_split = p.total_iters;
    }
    // Cluster split-K: the K splits of a tile form one thread-block cluster (2 / 4 / 8 CTAs) and are reduced through
    // distributed shared memory by the kernel itself - no partial tiles in global memory, no reduce launch.  The
    // GroupNorm slots it emits are min(128 / S, H*W) rows each; at most four per CTA slice.
    // Measured on B200 it LOSES to the separate reduce kernel (271 vs 287 steps/s: the 4x4-level layers take 25 us
    // instead of 14.5 us as clusters of 8, the 8x8 level is unchanged as clusters of 4), so it is opt-in
    // (AFLDM_TC_CSK=1) and documented as a negative result in DESIGN.md.
    static const int allow_csk = getenv("AFLDM_TC_CSK") ? atoi(getenv("AFLDM_TC_CSK")) : 0;
    p.csk = 0;
    p.csk_seg = 0;
    if (allow_csk && p.splitk > 1 && !p.two && !p.halo && (Cout & 3) == 0 && p.BN <= 192) {
        const int HW = H * W;
        for (int S = 8; S >= 2; S >>= 1) {
            if (S > p.splitk && S > 2) continue;               // never more splits than the rule asked for (except 2)
            const int R = TBM / S, seg = std::min(R, HW);
            if (R % seg != 0 || HW % seg != 0 || R / seg > 4) continue;
            const int ips = ceil_div(p.total_iters, S);
            if (ceil_div(p.total_iters, ips) != S || ips < 2) continue;
            p.csk = S;
            p.csk_seg = seg;
            p.splitk = S;
            p.iters_per_split = ips;
            break;
        }
    }
    const int tiles2 = p.mtiles * p.ntiles;
    // More CTAs than SMs and a small stage: size the ring so that two CTAs share an SM and one CTA's
    // prologue / epilogue hides behind the other's main loop.
    const int items = (p.two ? p.mtiles / 2 : p.mtiles) * p.ntiles * p.splitk;
    const int min_stages = p.halo ? 2 : 3;        // a halo stage carries three taps
    const bool two_per_sm = tiles2 * p.splitk > num_sms() && min_stages * stage_bytes + STAGING_BYTES <= SMEM_TWO_PER_SM;
    // Persistent grid: one or two CTAs per SM walk the work items; two TMEM accumulator buffers when they fit
    // (512 columns per SM) so that an item's epilogue runs under the next item's main loop.
    p.ctas_per_sm = two_per_sm ? 2 : 1;
    const int slots = num_sms() * p.ctas_per_sm;
    if (p.two) p.grid_ctas = 2 * std::min(items, slots / 2);
    else p.grid_ctas = std::min(items, slots);
    const int items_per_cta = ceil_div(items, p.two ? p.grid_ctas / 2 : p.grid_ctas);
    p.alias_staging = items_per_cta == 1;
    p.nacc = (2 * p.BN * p.ctas_per_sm <= 512 && items_per_cta > 1) ? 2 : 1;
    p.tmem_cols = 32;
    while (p.tmem_cols < p.nacc * p.BN) p.tmem_cols <<= 1;
    const int budget = (two_per_sm ? SMEM_TWO_PER_SM : SMEM_ONE_PER_SM) - (p.alias_staging ? 0 : STAGING_BYTES);
    p.stages = std::max(2, std::min(MAX_STAGES, budget / stage_bytes));
    if (force_stages > 0) p.stages = std::max(2, std::min(force_stages, p.stages));
    p.stages = std::min(p.stages, std::max(2, p.iters_per_split * items_per_cta));   // never more than there is to load
    if (p.csk) {
        // one tile per cluster; the parked partial tile [128][BN + 4] sits behind the staging area, inside the ring
        p.grid_ctas = tiles2 * p.csk;
        p.alias_staging = 1;
        p.nacc = 1;
        p.tmem_cols = 32;
        while (p.tmem_cols < p.BN) p.tmem_cols <<= 1;
        const int need = STAGING_BYTES + TBM * (p.BN + 4) * 4;
        while (p.stages * stage_bytes < need) ++p.stages;
        if (p.stages > MAX_STAGES) { p.csk = 0; p.ok = false; return p; }
    }
    p.smem_bytes = (size_t)p.stages * stage_bytes + (p.alias_staging ? 0 : STAGING_BYTES) + BARRIER_BYTES;
    // tcgen05.alloc blocks until its columns are free, so TMEM must never be over-subscribed by the conv CTAs that can
    // share an SM - two CTAs of this launch, the head of a PDL-overlapped successor, a conv_shortcut on the side
    // stream.  A CTA pair (cta_group::2) that holds its columns on one SM while its peer waits on the other can dead-lock
    // against a pair of another kernel doing the same (seen on B200: an intermittent hang of the SD-1.5 video loop).
    // Shared memory is the resource that decides co-residency, so every CTA asks for shared memory IN PROPORTION to its
    // columns: 454 B per column including the 1 KB the hardware reserves per CTA.  CTAs that fit on one SM (228 KB) then
    // hold at most 233472 / 454 = 514 columns - column counts are powers of two >= 32, so at most 512 - whatever mix of
    // launches they come from; 256 columns cost exactly the two-per-SM budget (115200 B), 512 columns 231424 B.
    {
        const size_t floor_bytes = (size_t)p.tmem_cols * 454 - 1024;
        if (p.smem_bytes < floor_bytes) p.smem_bytes = floor_bytes;
    }
    p.ok = true;
    return p;
}

}  // namespace

int conv_tc_gn_slots(int B, int H, int W, int Cin, int Cout, int ks, int x_half) {
    const TcPlan p = tc_plan(B, H, W, Cin, Cout, ks, x_half != 0);
    if (!p.ok || (Cout & 3) != 0 || p.BN > 192) return 0;
    if (p.csk) return (H * W) / p.csk_seg;               // cluster split-K: one slot per csk_seg rows
    if (p.splitk > 1) return splitk_reduce_slots(H * W);  // the split-K reduce emits one slot per 16 rows
    const int HW = H * W;
    if (HW % TBM == 0) return HW / TBM;                  // tiles inside one image: one slot per tile
    return (HW == 64 || HW == 32) ? 1 : 0;               // tiles of whole images aligned to the epilogue's row quarters
}

bool conv_tc_plan_query(int B, int H, int W, int Cin, int Cout, int ks, int x_half, int out[8]) {
    const TcPlan p = tc_plan(B, H, W, Cin, Cout, ks, x_half != 0);
    if (!p.ok) return false;
    out[0] = p.tmem_cols; out[1] = (int)p.smem_bytes; out[2] = p.grid_ctas; out[3] = p.two;
    out[4] = p.BN; out[5] = p.splitk; out[6] = p.stages; out[7] = p.halo;
    return true;
}

bool conv_tc_supported(int B, int H, int W, int Cin, int Cout, int ks, int x_half) {
    return tc_plan(B, H, W, Cin, Cout, ks, x_half != 0).ok;
}

bool conv_tc_workspace_floats(int B, int H, int W, int Cin, int Cout, int ks, size_t* floats, int x_half) {
    const TcPlan p = tc_plan(B, H, W, Cin, Cout, ks, x_half != 0);
    if (!p.ok) return false;
    *floats = (p.splitk > 1 && !p.csk) ? (size_t)p.splitk * p.M * Cout : 0;
    return true;
}

int conv_tc_launch(const float* x, int x_pitch, const float* w, const float* bias, const float* row_add,
                   int row_add_pitch, const float* residual, int res_pitch, float* y, int y_pitch, int B, int H,
                   int W, int Cin, int Cout, int ks, float* workspace, size_t workspace_floats, float* gn_partial,
                   cudaStream_t st, int y_half, const float* x2, int x2_pitch, int Cin1, int x_half) {
    // x_half: x (and x2) and w hold fp16 elements (pitches in elements): tcgen05.mma.kind::f16, 64 channels per stage
    const bool f16 = x_half != 0;
    const int TBK = tbk(f16);
    const int esz = f16 ? 2 : 4;                       // operand element size
    const int pmask = f16 ? 7 : 3;                     // pixel pitch must keep rows 16-byte aligned (TMA global strides)
    // x2 != NULL: input channels [0, Cin1) are read from x, [Cin1, Cin) from x2 (torch.cat never materialised)
    if (x2 != nullptr && (Cin1 <= 0 || Cin1 >= Cin || Cin1 % TBK != 0 || (x2_pitch & pmask) != 0 || !aligned16(x2)))
        return AFLDM_E_NOKERNEL;
    const TcPlan p = tc_plan(B, H, W, Cin, Cout, ks, f16);
    if (y_half && (!p.ok || p.splitk > 1 || residual != nullptr || gn_partial != nullptr || (Cout & 7) != 0 ||
                   (y_pitch & 7) != 0 || !aligned16(y) || (bias != nullptr && !aligned16(bias)) ||
                   (row_add != nullptr && ((row_add_pitch & 3) != 0 || !aligned16(row_add)))))
        return AFLDM_E_NOKERNEL;             // fp16 stores live in the vectorised un-split epilogue only
    const int gn_slots = gn_partial != nullptr ? conv_tc_gn_slots(B, H, W, Cin, Cout, ks, x_half) : 0;
    if (gn_partial != nullptr && gn_slots == 0) return AFLDM_E_SHAPE;
    if (!p.ok || (x_pitch & pmask) != 0 || !aligned16(x) || !aligned16(w)) return AFLDM_E_NOKERNEL;
    EncodeTiledFn enc = encode_fn();
    if (enc == nullptr) return AFLDM_E_NOKERNEL;
    if (p.splitk > 1 && !p.csk && (workspace == nullptr || workspace_floats < (size_t)p.splitk * p.M * Cout))
        return AFLDM_E_WORKSPACE;

    CUtensorMap map_a, map_a2, map_b;
    auto encode_a = [&](CUtensorMap* m, const float* src, int pitch, int channels) {
        const cuuint64_t dims[4] = {(cuuint64_t)channels, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        const cuuint64_t strides[3] = {(cuuint64_t)pitch * esz, (cuuint64_t)pitch * esz * W, (cuuint64_t)pitch * esz * W * H};
        const cuuint32_t box[4] = {(cuuint32_t)TBK, (cuuint32_t)p.BW, (cuuint32_t)(p.halo ? p.BH + 2 : p.BH), (cuuint32_t)p.BB};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        return enc(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(src),
                   dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    if (!encode_a(&map_a, x, x_pitch, x2 != nullptr ? Cin1 : Cin)) return AFLDM_E_NOKERNEL;
    if (x2 != nullptr) {
        if (!encode_a(&map_a2, x2, x2_pitch, Cin - Cin1)) return AFLDM_E_NOKERNEL;
    } else {
        map_a2 = map_a;
    }
    {
        const cuuint64_t K = (cuuint64_t)ks * ks * Cin;
        const cuuint64_t dims[2] = {K, (cuuint64_t)Cout};
        const cuuint64_t strides[1] = {K * esz};
        const cuuint32_t box[2] = {(cuuint32_t)TBK, (cuuint32_t)(p.two ? p.BN / 2 : p.BN)};
        const cuuint32_t estr[2] = {1, 1};
        if (enc(&map_b, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(w),
                dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return AFLDM_E_NOKERNEL;
    }

    static std::atomic<unsigned long long> configured{0};      // one bit per device ordinal (the attribute is per device)
    int dev_ord = 0;
    if (cudaGetDevice(&dev_ord) != cudaSuccess) return AFLDM_E_NOKERNEL;
    if (!(configured.load(std::memory_order_acquire) & (1ull << (dev_ord & 63)))) {
        cudaError_t e = cudaSuccess;
        const void* kerns[8] = {(const void*)conv_tc_kernel<false, false, false>, (const void*)conv_tc_kernel<false, true, false>,
                                (const void*)conv_tc_kernel<true, false, false>, (const void*)conv_tc_kernel<true, true, false>,
                                (const void*)conv_tc_kernel<false, false, true>, (const void*)conv_tc_kernel<false, true, true>,
                                (const void*)conv_tc_kernel<true, false, true>, (const void*)conv_tc_kernel<true, true, true>};
        for (const void* kp : kerns) {
            e = cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ONE_PER_SM + BARRIER_BYTES);
            if (e != cudaSuccess) return (int)e;
        }
        configured.fetch_or(1ull << (dev_ord & 63), std::memory_order_release);
    }
    TcArgs a;
    a.bias = bias; a.row_add = row_add; a.residual = residual; a.y = y; a.ws = workspace;
    a.row_add_pitch = row_add_pitch; a.res_pitch = res_pitch; a.y_pitch = y_pitch;
    a.M = p.M; a.Cout = Cout; a.HW = H * W;
    a.taps = p.halo ? 3 : ks * ks; a.ks = ks; a.cin_chunks = Cin / TBK;
    a.halo = p.halo;
    a.iters_per_split = p.iters_per_split;
    a.BN = p.BN; a.stages = p.stages; a.tmem_cols = p.tmem_cols;
    a.W = W; a.H = H; a.BW = p.BW; a.BH = p.BH;
    a.gn_partial = (p.splitk > 1 && !p.csk) ? nullptr : gn_partial;
    a.csk = p.csk;
    a.csk_seg = p.csk_seg;
    a.gn_slots = gn_slots;
    a.y_half = y_half;
    a.cin1_chunks = x2 != nullptr ? Cin1 / TBK : Cin / TBK;
    a.vec_ok = ((Cout & 3) == 0) && ((y_pitch & 3) == 0) && (y_half || aligned16(y)) && (bias == nullptr || aligned16(bias)) &&
               (row_add == nullptr || (((row_add_pitch & 3) == 0) && aligned16(row_add))) &&
               (residual == nullptr || (((res_pitch & 3) == 0) && aligned16(residual)));
    a.mtiles = p.mtiles; a.ntiles = p.ntiles; a.splitk = p.splitk; a.nacc = p.nacc;
    a.alias_staging = p.alias_staging;
    dim3 grid(p.grid_ctas, 1, 1);
    if (p.csk) {
        if (!a.vec_ok || y_half) return AFLDM_E_NOKERNEL;      // the DSMEM reduction is float4 throughout
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = p.smem_bytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = p.csk;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (f16) (void)cudaLaunchKernelEx(&cfg, conv_tc_kernel<false, false, true>, map_a, map_a2, map_b, a);
        else (void)cudaLaunchKernelEx(&cfg, conv_tc_kernel<false, false, false>, map_a, map_a2, map_b, a);
        return launched(1);
    }
    if (p.two) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = p.smem_bytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        // programmatic dependent launch for the CTA-pair launches too (they were the only kernels of the step launched
        // fully serialised; AFLDM_PDL_PAIRS=0 restores that)
        static const bool pdl_pairs = !(getenv("AFLDM_PDL_PAIRS") && atoi(getenv("AFLDM_PDL_PAIRS")) == 0);
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = (pdl_enabled() && pdl_pairs) ? 2 : 1;
        if (p.halo) {
            if (f16) (void)cudaLaunchKernelEx(&cfg, conv_tc_kernel<true, true, true>, map_a, map_a2, map_b, a);
            else (void)cudaLaunchKernelEx(&cfg, conv_tc_kernel<true, true, false>, map_a, map_a2, map_b, a);
        } else {
            if (f16) (void)cudaLaunchKernelEx(&cfg, conv_tc_kernel<true, false, true>, map_a, map_a2, map_b, a);
            else (void)cudaLaunchKernelEx(&cfg, conv_tc_kernel<true, false, false>, map_a, map_a2, map_b, a);
        }
    } else if (p.halo) {
        if (f16) launch_k(conv_tc_kernel<false, true, true>, dim3(grid), dim3(TC_THREADS), p.smem_bytes, st, map_a, map_a2, map_b, a);
        else launch_k(conv_tc_kernel<false, true, false>, dim3(grid), dim3(TC_THREADS), p.smem_bytes, st, map_a, map_a2, map_b, a);
    } else {
        if (f16) launch_k(conv_tc_kernel<false, false, true>, dim3(grid), dim3(TC_THREADS), p.smem_bytes, st, map_a, map_a2, map_b, a);
        else launch_k(conv_tc_kernel<false, false, false>, dim3(grid), dim3(TC_THREADS), p.smem_bytes, st, map_a, map_a2, map_b, a);
    }
    int launches = 1;
    if (p.splitk > 1) {
        splitk_reduce_launch(workspace, p.splitk, bias, row_add, row_add_pitch, residual, res_pitch, y, y_pitch,
                             p.M, Cout, H * W, gn_partial, st);
        ++launches;
    }
    return launched(launches);
}

}  // namespace afldm
