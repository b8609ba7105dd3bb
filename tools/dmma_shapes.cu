// dmma_shapes.cu -- the larger fp64 mma.sync shapes (m16n8k4 / k8 / k16, sm_90+) on sm_100a: sustained throughput against
// m8n8k4, and a check of the fragment layouts used (A: row = g + 8 (i & 1), col = t + 4 (i >> 1); B: row = t + 4 i, col = g;
// C: c0,c1 = row g, cols 2t, 2t+1; c2,c3 = row g + 8) with g = lane >> 2, t = lane & 3.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_shapes tools/dmma_shapes.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

template <int KS> struct Frag;   // KS = k / 4: number of A pairs
__device__ __forceinline__ void mma_k4(double (&c)[4], const double (&a)[2], const double (&b)[1]) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
}
__device__ __forceinline__ void mma_k8(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma_k16(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void mma_884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int MODE, int NACC>
__global__ void thr_kernel(double* out, int iters, double x, double y) {
    double c[NACC][4];
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x * 1e-3 + i + j;
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = x + i * 1e-9;
    for (int i = 0; i < 4; ++i) b[i] = y + i * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (MODE == 0) { double c2[2] = {c[i][0], c[i][1]}; mma_884(c2, a[0], b[0]); c[i][0] = c2[0]; c[i][1] = c2[1];
                             double c3[2] = {c[i][2], c[i][3]}; mma_884(c3, a[1], b[0]); c[i][2] = c3[0]; c[i][3] = c3[1]; }
            if (MODE == 4) { double aa[2] = {a[0], a[1]}; double bb[1] = {b[0]}; mma_k4(c[i], aa, bb); }
            if (MODE == 8) { double aa[4] = {a[0], a[1], a[2], a[3]}; double bb[2] = {b[0], b[1]}; mma_k8(c[i], aa, bb); }
            if (MODE == 16) mma_k16(c[i], a, b);
        }
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void check_kernel(const double* A, const double* B, double* D, int mode) {   // A [16][16] row-major, B [16][8] (k x n), D [16][8]
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    double c[4] = {0, 0, 0, 0};
    if (mode == 16) {
        double a[8], b[4];
        for (int i = 0; i < 8; ++i) a[i] = A[(g + 8 * (i & 1)) * 16 + t + 4 * (i >> 1)];
        for (int i = 0; i < 4; ++i) b[i] = B[(t + 4 * i) * 8 + g];
        mma_k16(c, a, b);
    } else if (mode == 8) {
        for (int h = 0; h < 2; ++h) {
            double a[4], b[2];
            for (int i = 0; i < 4; ++i) a[i] = A[(g + 8 * (i & 1)) * 16 + 8 * h + t + 4 * (i >> 1)];
            for (int i = 0; i < 2; ++i) b[i] = B[(8 * h + t + 4 * i) * 8 + g];
            mma_k8(c, a, b);
        }
    } else {
        for (int h = 0; h < 4; ++h) {
            double a[2], b[1];
            for (int i = 0; i < 2; ++i) a[i] = A[(g + 8 * i) * 16 + 4 * h + t];
            b[0] = B[(4 * h + t) * 8 + g];
            mma_k4(c, a, b);
        }
    }
    D[g * 8 + 2 * t] = c[0]; D[g * 8 + 2 * t + 1] = c[1]; D[(g + 8) * 8 + 2 * t] = c[2]; D[(g + 8) * 8 + 2 * t + 1] = c[3];
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device %s sms %d\n", p.name, p.multiProcessorCount);
    double hA[256], hB[128], hD[128], ref[128];
    srand(1);
    for (int i = 0; i < 256; ++i) hA[i] = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < 128; ++i) hB[i] = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < 16; ++i) for (int j = 0; j < 8; ++j) { double s = 0; for (int k = 0; k < 16; ++k) s += hA[i * 16 + k] * hB[k * 8 + j]; ref[i * 8 + j] = s; }
    double *dA, *dB, *dD, *out;
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, sizeof hD); cudaMalloc(&out, sizeof(double) * 148 * 2 * 1024);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    for (int mode : {4, 8, 16}) {
        check_kernel<<<1, 32>>>(dA, dB, dD, mode);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
        double err = 0; for (int i = 0; i < 128; ++i) err = fmax(err, fabs(hD[i] - ref[i]));
        printf("layout check m16n8k%-2d: max abs error %.2e (%s)\n", mode, err, cudaGetErrorString(e));
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = p.multiProcessorCount * 2, threads = 512, iters = 20000;
    auto run = [&](auto kern, const char* name, double flops_per_inst_group) {
        kern<<<grid, threads>>>(out, 100, 1.0000001, 1e-9); cudaDeviceSynchronize();
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) { cudaEventRecord(e0); kern<<<grid, threads>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
        double fl = flops_per_inst_group * 8 * iters * (double)grid * (threads / 32);
        printf("%-28s %8.3f ms  %6.2f TFLOP/s\n", name, best, fl / best * 1e-9);
    };
    run(thr_kernel<0, 8>, "2 x m8n8k4 (8 acc groups)", 2.0 * 2 * 256);
    run(thr_kernel<4, 8>, "m16n8k4", 2.0 * 16 * 8 * 4);
    run(thr_kernel<8, 8>, "m16n8k8", 2.0 * 16 * 8 * 8);
    run(thr_kernel<16, 8>, "m16n8k16", 2.0 * 16 * 8 * 16);
    return 0;
}
