// fp64 pipe micro-benchmarks for the roofline denominator (DFMA vs DMMA.8x8x4) on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dmma_kernel(double* out, int iters, double a, double b) {
    double c0[NACC], c1[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c0[i] = threadIdx.x * 1e-3; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sms %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int threads : {128, 256, 512, 1024}) {
        for (int bps : {1, 2}) {
            if (threads * bps > 2048) continue;
            int grid = p.multiProcessorCount * bps, iters = 20000;
            dfma_kernel<<<grid, threads>>>(out, 100, 1.0000001, 1e-9);
            CK(cudaDeviceSynchronize());
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0); dfma_kernel<<<grid, threads>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1);
                CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            double fl = 2.0 * 16 * iters * (double)grid * threads;
            printf("DFMA  threads %4d x %d/SM : %.3f ms  %.2f TFLOP/s\n", threads, bps, best, fl / best * 1e-9);
            dmma_kernel<8><<<grid, threads>>>(out, 100, 1.0000001, 1e-9);
            CK(cudaDeviceSynchronize());
            best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0); dmma_kernel<8><<<grid, threads>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1);
                CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            fl = 2.0 * 256 * 8 * iters * (double)grid * (threads / 32);
            printf("DMMA8 threads %4d x %d/SM : %.3f ms  %.2f TFLOP/s\n", threads, bps, best, fl / best * 1e-9);
        }
    }
    // sustained: 3 s of DMMA
    {
        int grid = p.multiProcessorCount * 2, threads = 512, iters = 200000;
        cudaEventRecord(e0);
        for (int i = 0; i < 10; ++i) dmma_kernel<8><<<grid, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 10 * 2.0 * 256 * 8 * iters * (double)grid * (threads / 32);
        printf("DMMA8 sustained %.0f ms: %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
        cudaEventRecord(e0);
        for (int i = 0; i < 10; ++i) dfma_kernel<<<grid, threads>>>(out, iters/4, 1.0000001, 1e-9);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        fl = 10 * 2.0 * 16 * (iters/4) * (double)grid * threads;
        printf("DFMA sustained %.0f ms: %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
    }
    return 0;
}
