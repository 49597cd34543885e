// fp64_peak.cu — measures the FP64 math-pipe peaks of this GPU (the roofline denominators that
// MEASURED_PEAKS.json does not carry): register-resident DMMA (mma.sync.m8n8k4.f64) and DFMA loops.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu ; prints one JSON line.
#include <cuda_runtime.h>
#include <cstdio>

__global__ void __launch_bounds__(256) k_dmma(double *out, int iters) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[2 * i]), "+d"(c[2 * i + 1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_dfma(double *out, int iters) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, blocks = sms * 8, threads = 256, iters = 20000;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best_dmma = 1e30f, best_dfma = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0); k_dmma<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best_dmma) best_dmma = ms;
        cudaEventRecord(e0); k_dfma<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best_dfma) best_dfma = ms;
    }
    const double warps = (double)blocks * threads / 32;
    const double dmma_flop = warps * iters * 8.0 * (8 * 8 * 4 * 2);
    const double dfma_flop = (double)blocks * threads * iters * 16.0 * 2;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f, \"dmma_ms\": %.3f, \"dfma_ms\": %.3f}\n",
           p.name, sms, dmma_flop / best_dmma / 1e9, dfma_flop / best_dfma / 1e9, best_dmma, best_dfma);
    return 0;
}
