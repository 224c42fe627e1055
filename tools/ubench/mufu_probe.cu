// MUFU throughput probe (B200): ex2.approx.ftz.f32 vs ex2.approx.f16x2 (two results per instruction?).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_probe mufu_probe.cu ; run: ./mufu_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) probe(uint32_t* out, int iters, uint32_t seed) {
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = seed + threadIdx.x * 8 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(r[i]));
            else asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r[i]));
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= r[i];
    if (acc == 0x12345678u) out[0] = acc;
}

template <int MODE>
float run(int sms, int iters) {
    uint32_t* out;
    cudaMalloc(&out, 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    probe<MODE><<<sms, 1024>>>(out, iters, 0x3c003c00u);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    probe<MODE><<<sms, 1024>>>(out, iters, 0x3c003c00u);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    cudaFree(out);
    return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount, iters = 20000;
    const double instr_per_sm = 1024.0 * 8 * iters;   // thread-level MUFU ops per SM
    for (int rep = 0; rep < 2; ++rep) {
        float m0 = run<0>(sms, iters), m1 = run<1>(sms, iters);
        printf("f32   : %.3f ms -> %.2f ops/clk/SM at %d MHz nominal\n", m0, instr_per_sm / (m0 * 1e-3 * clk_khz * 1e3), clk_khz / 1000);
        printf("f16x2 : %.3f ms -> %.2f instr/clk/SM = %.2f results/clk/SM\n", m1, instr_per_sm / (m1 * 1e-3 * clk_khz * 1e3), 2 * instr_per_sm / (m1 * 1e-3 * clk_khz * 1e3));
    }
    return 0;
}
