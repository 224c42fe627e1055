// Microbenchmark: warp-level mma.sync throughput on sm_100a (legacy tensor path), per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int KIND>
__global__ void __launch_bounds__(512) k(float* out, int iters) {
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, threadIdx.x * 5u, threadIdx.x * 7u};
    uint32_t b0 = threadIdx.x * 11u, b1 = threadIdx.x * 13u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
            else if (KIND == 2)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a[0]), "r"(a[1]), "r"(b0));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 123.456f) out[0] = s;
}

template <int KIND>
void run(const char* name, int macs_per_mma, int threads) {
    float* out;
    cudaMalloc(&out, 4);
    const int iters = 20000, blocks = 148;
    k<KIND><<<blocks, threads>>>(out, 100);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<KIND><<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double mmas = (double)blocks * (threads / 32) * iters * 8;
    double tflops = mmas * macs_per_mma * 2 / (ms * 1e-3) / 1e12;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double mac_per_clk_sm = mmas * macs_per_mma / (ms * 1e-3) / 148 / (clk * 1e3);
    printf("%-22s warps/SM %2d: %8.1f TFLOP/s  %7.1f MAC/clk/SM (at %d MHz nominal)  %.3f ms\n", name, threads / 32, tflops, mac_per_clk_sm, clk / 1000, ms);
}

int main() {
    for (int th : {128, 256, 512}) {
        run<0>("tf32 m16n8k8", 16 * 8 * 8, th);
        run<3>("tf32 m16n8k4", 16 * 8 * 4, th);
        run<1>("bf16 m16n8k16", 16 * 8 * 16, th);
        run<2>("f16 m16n8k16", 16 * 8 * 16, th);
    }
    return 0;
}
