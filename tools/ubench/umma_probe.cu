// Probe of the tcgen05 conventions the filtered-activation kernel (csrc/fact_tc.cu) relies on, run once on a B200:
//   mode 0  A, B K-major SWIZZLE_128B tiles WRITTEN BY THREADS (st.shared + fence.proxy.async), M=128 N=32 K=64
//   mode 1  A MN-major SWIZZLE_128B, descriptor LBO = stride between 64-element M atoms, SBO = stride between 8-k groups
//   mode 2  A MN-major SWIZZLE_128B, the two descriptor fields swapped
//   mode 3  mode 0 on top of an accumulator pre-loaded with tcgen05.st (accumulate = 1 from the first MMA)
//   mode 4  N = 64 instruction shape (mode 0 data, B has 64 rows)
//   rate    tcgen05.ld 32x32b.x32 throughput with 4 / 8 / 16 warps (clock64)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu ; prints PASS / FAIL per mode.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
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
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

__host__ __device__ inline int aval(int m, int k) { return (m * 7 + k * 3) % 13 - 6; }
__host__ __device__ inline int bval(int n, int k) { return (n * 5 + k * 11) % 9 - 4; }
__host__ __device__ inline int cval(int m, int n) { return (m + 3 * n) % 17 - 8; }

// K-major SWIZZLE_128B: row r (128 B), 16-byte chunk c -> r * 128 + ((c ^ (r & 7)) * 16)
__device__ __forceinline__ uint32_t kmaj_off(int row, int k) {          // k in fp16 elements (< 64)
    return (uint32_t)(row * 128 + ((((k >> 3) ^ (row & 7)) << 4) | ((k & 7) << 1)));
}
// MN-major SWIZZLE_128B: element (m, k) -> (m / 64) * MS + (k / 8) * KS + (k % 8) * 128 + (((m % 64) / 8) ^ (k % 8)) * 16 + (m % 8) * 2
__device__ __forceinline__ uint32_t mnmaj_off(int m, int k, int MS, int KS) {
    return (uint32_t)((m >> 6) * MS + (k >> 3) * KS + (k & 7) * 128 + (((((m & 63) >> 3) ^ (k & 7))) << 4) + ((m & 7) << 1));
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(128, 1) probe_kernel(int mode, float* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* sA = smem;                  // 16 KB
    uint8_t* sB = smem + 16384;          // 8 KB
    const bool mn = (mode == 1 || mode == 2);
    const int N = (mode == 4) ? 64 : 32;
    const int K = mn ? 32 : 64;
    for (int i = tid; i < (16384 + 8192) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    __syncthreads();
    // A
    const int MS = 4096, KS = 1024;      // MN-major: M atoms 4 KB apart (4 k-groups of 1 KB each in between)
    for (int idx = tid; idx < 128 * K; idx += 128) {
        const int m = idx / K, k = idx % K;
        const uint32_t off = mn ? mnmaj_off(m, k, MS, KS) : kmaj_off(m, k);
        *reinterpret_cast<__half*>(sA + off) = __float2half_rn((float)aval(m, k));
    }
    for (int idx = tid; idx < N * K; idx += 128) {
        const int n = idx / K, k = idx % K;
        *reinterpret_cast<__half*>(sB + kmaj_off(n, k)) = __float2half_rn((float)bval(n, k));
    }
    if (tid == 0) {
        mbar_init(smem_u32(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    if (mode == 3) {
        for (int c0 = 0; c0 < N; c0 += 8) {
            uint32_t v[8];
            for (int j = 0; j < 8; ++j) v[j] = __float_as_uint((float)cval(tid, c0 + j));
            tmem_st8(trow + (uint32_t)c0, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (tid == 0) {
        const uint32_t a_major = mn ? 1u : 0u;
        const uint32_t idesc = (1u << 4) | (a_major << 15) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int ks = 0; ks < K / 16; ++ks) {
            uint64_t da, db;
            if (!mn) da = desc_sw128(smem_u32(sA) + ks * 32, 16, 1024);
            else if (mode == 1) da = desc_sw128(smem_u32(sA) + ks * 2 * KS, MS, KS);
            else da = desc_sw128(smem_u32(sA) + ks * 2 * KS, KS, MS);
            db = desc_sw128(smem_u32(sB) + ks * 32, 16, 1024);
            umma_f16(tmem, da, db, idesc, (ks > 0 || mode == 3) ? 1u : 0u);
        }
        umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(trow + (uint32_t)c0, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) out[tid * 64 + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
    (void)lane;
}

__global__ void ldtm_rate_kernel(long long* cycles, float* sink, int iters) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t trow = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) & 3) * 32u;
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t r[32];
        tmem_ld32(trow, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += __uint_as_float(r[i & 31]);
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(256u) : "memory");
}

int main() {
    float* d_out;
    cudaMalloc(&d_out, 128 * 64 * sizeof(float));
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 8192 + 1024);
    std::vector<float> h(128 * 64);
    const char* names[5] = {"K-major thread-written", "MN-major (LBO = M-atom stride, SBO = k-group stride)", "MN-major (fields swapped)",
                            "tcgen05.st preload + accumulate", "N = 64"};
    for (int mode = 0; mode < 5; ++mode) {
        cudaMemset(d_out, 0, 128 * 64 * sizeof(float));
        probe_kernel<<<1, 128, 16384 + 8192 + 1024>>>(mode, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("mode %d (%s): CUDA error %s\n", mode, names[mode], cudaGetErrorString(e));
            return 1;
        }
        cudaMemcpy(h.data(), d_out, 128 * 64 * sizeof(float), cudaMemcpyDeviceToHost);
        const int N = mode == 4 ? 64 : 32, K = (mode == 1 || mode == 2) ? 32 : 64;
        int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                int want = mode == 3 ? cval(m, n) : 0;
                for (int k = 0; k < K; ++k) want += aval(m, k) * bval(n, k);
                if (h[m * 64 + n] != (float)want) {
                    if (bad < 4) printf("  mode %d mismatch at (%d,%d): got %g want %d\n", mode, m, n, h[m * 64 + n], want);
                    ++bad;
                }
            }
        printf("mode %d (%s): %s (%d mismatches)\n", mode, names[mode], bad ? "FAIL" : "PASS", bad);
    }
    long long* d_cyc;
    float* d_sink;
    cudaMalloc(&d_cyc, 8 * sizeof(long long));
    cudaMalloc(&d_sink, 4);
    for (int threads : {128, 256, 512}) {
        const int iters = 2000;
        ldtm_rate_kernel<<<1, threads>>>(d_cyc, d_sink, iters);
        cudaDeviceSynchronize();
        ldtm_rate_kernel<<<1, threads>>>(d_cyc, d_sink, iters);
        cudaError_t e = cudaDeviceSynchronize();
        long long c = 0;
        cudaMemcpy(&c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
        const double bytes = (double)iters * (threads / 32) * 32 * 32 * 4;
        printf("LDTM 32x32b.x32, %d warps: %lld cycles for %d loads/warp -> %.1f cyc per load (per warp), %.1f B/cyc/SM (%s)\n",
               threads / 32, c, iters, (double)c / iters, bytes / (double)c, cudaGetErrorString(e));
    }
    return 0;
}
