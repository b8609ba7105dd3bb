// barrier_probe.cu -- latency of grid-wide synchronisation + a 2-value all-reduce on B200 (148 CTAs x 544 threads, one CTA per SM):
//   (a) cooperative-groups grid.sync() + partials through memory (what pcg_persistent_kernel does today)
//   (b) counter barrier taken by thread 0 (top-bit flip) + partials through memory
//   (c) flag-carrying all-gather: every CTA publishes {lo32|flag, hi32|flag} words of its two partial sums, everybody polls all
//       slots: synchronisation and data exchange in ONE round trip, no atomics (the NCCL LL idea)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/barrier_probe tools/barrier_probe.cu ; run: tools/barrier_probe
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
constexpr int T = 544;

__device__ __forceinline__ unsigned ld_acq(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(T) probe(int mode, int reps, double* part, unsigned* ctr, unsigned long long* slots, double* out,
                                           unsigned base) {
    cg::grid_group grid = cg::this_grid();
    const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ double red_s[2][32];
    __shared__ double gath[2][160];
    double acc0 = 1.0 + cta * 1e-3 + tid * 1e-6, acc1 = 2.0, tot0 = 0, tot1 = 0;
    for (int rep = 0; rep < reps; ++rep) {
        double v0 = acc0, v1 = acc1;
        for (int o = 16; o > 0; o >>= 1) {
            v0 += __shfl_xor_sync(0xffffffffu, v0, o);
            v1 += __shfl_xor_sync(0xffffffffu, v1, o);
        }
        __syncthreads();
        if (lane == 0) { red_s[0][warp] = v0; red_s[1][warp] = v1; }
        __syncthreads();
        double t0 = 0, t1 = 0;
        if (tid == 0) {
            for (int w = 0; w < T / 32; ++w) { t0 += red_s[0][w]; t1 += red_s[1][w]; }
        }
        const int par = rep & 1;
        if (mode <= 1) {
            double* p = part + (size_t)par * G * 2;
            if (tid == 0) { p[2 * cta] = t0; p[2 * cta + 1] = t1; }
            if (mode == 0) {
                __syncthreads();
                if (tid == 0) __threadfence();
                grid.sync();
            } else {
                __syncthreads();
                if (tid == 0) {
                    __threadfence();
                    const unsigned add = cta == 0 ? 0x80000000u - (unsigned)(G - 1) : 1u;
                    const unsigned old = atomicAdd(ctr, add);
                    while (((old ^ ld_acq(ctr)) & 0x80000000u) == 0u) {}
                    __threadfence();
                }
                __syncthreads();
            }
            if (warp == 0) {
                double s0 = 0, s1 = 0;
                double2 v[5];
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const int g = lane + 32 * k;
                    v[k] = g < G ? __ldcg(reinterpret_cast<const double2*>(p) + g) : make_double2(0, 0);
                }
#pragma unroll
                for (int k = 0; k < 5; ++k) { s0 += v[k].x; s1 += v[k].y; }
                for (int o = 16; o > 0; o >>= 1) {
                    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                }
                if (lane == 0) { red_s[0][0] = s0; red_s[1][0] = s1; }
            }
            __syncthreads();
        } else {
            // publish: 4 words per CTA {lo32(v0)|flag, hi32(v0)|flag, lo32(v1)|flag, hi32(v1)|flag}
            const unsigned flag = base + (unsigned)rep + 1u;
            unsigned long long* sl = slots + (size_t)par * G * 4;
            if (tid == 0) {
                __threadfence();                                   // this CTA's earlier global writes first
                const unsigned long long b0 = (unsigned long long)__double_as_longlong(t0), b1 = (unsigned long long)__double_as_longlong(t1);
                st_relaxed64(sl + 4 * cta + 0, (b0 & 0xffffffffull) | ((unsigned long long)flag << 32));
                st_relaxed64(sl + 4 * cta + 1, (b0 >> 32) | ((unsigned long long)flag << 32));
                st_relaxed64(sl + 4 * cta + 2, (b1 & 0xffffffffull) | ((unsigned long long)flag << 32));
                st_relaxed64(sl + 4 * cta + 3, (b1 >> 32) | ((unsigned long long)flag << 32));
            }
            // gather: warps 0..4 poll 32 CTAs each (4 words per lane)
            if (warp < 5) {
                const int g = warp * 32 + lane;
                if (g < G) {
                    unsigned long long w0, w1, w2, w3;
                    do {
                        w0 = ld_relaxed64(sl + 4 * g + 0);
                        w1 = ld_relaxed64(sl + 4 * g + 1);
                        w2 = ld_relaxed64(sl + 4 * g + 2);
                        w3 = ld_relaxed64(sl + 4 * g + 3);
                    } while ((unsigned)(w0 >> 32) != flag || (unsigned)(w1 >> 32) != flag || (unsigned)(w2 >> 32) != flag ||
                             (unsigned)(w3 >> 32) != flag);
                    gath[0][g] = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
                    gath[1][g] = __longlong_as_double((long long)((w2 & 0xffffffffull) | (w3 << 32)));
                }
                __threadfence();                                   // acquire side: the other CTAs' data behind their flags
            }
            __syncthreads();
            if (warp == 0) {
                double s0 = 0, s1 = 0;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const int g = lane + 32 * k;
                    if (g < G) { s0 += gath[0][g]; s1 += gath[1][g]; }
                }
                for (int o = 16; o > 0; o >>= 1) {
                    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                }
                if (lane == 0) { red_s[0][0] = s0; red_s[1][0] = s1; }
            }
            __syncthreads();
        }
        tot0 += red_s[0][0];
        tot1 += red_s[1][0];
        acc0 += 1e-9 * tot0;
    }
    if (tid == 0 && cta == 0) { out[0] = tot0; out[1] = tot1; }
}

int main() {
    int dev = 0, sms = 0;
    cudaSetDevice(dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double *part, *out;
    unsigned* ctr;
    unsigned long long* slots;
    cudaMalloc(&part, 2 * sms * 2 * sizeof(double));
    cudaMalloc(&out, 2 * sizeof(double));
    cudaMalloc(&ctr, 64);
    cudaMalloc(&slots, 2 * sms * 4 * sizeof(unsigned long long));
    cudaMemset(ctr, 0, 64);
    cudaMemset(slots, 0, 2 * sms * 4 * sizeof(unsigned long long));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const char* names[3] = {"cg grid.sync + partials", "counter barrier (thread 0) + partials", "flag-carrying all-gather (no atomics)"};
    unsigned base = 0;
    for (int mode = 0; mode < 3; ++mode) {
        for (int pass = 0; pass < 2; ++pass) {
            int reps = 2000;
            void* args[] = {&mode, &reps, &part, &ctr, &slots, &out, &base};
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            cudaEventRecord(e0);
            cudaError_t err = cudaLaunchCooperativeKernel((void*)probe, dim3(sms), dim3(T), args, 200 * 1024, 0);
            cudaEventRecord(e1);
            cudaError_t err2 = cudaDeviceSynchronize();
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            double h[2];
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            base += reps + 7;
            if (pass == 1)
                printf("%-42s %7.3f us per synchronised 2-value all-reduce  (%s %s, sum %.6f)\n", names[mode], 1e3 * ms / reps,
                       cudaGetErrorString(err), cudaGetErrorString(err2), h[0]);
        }
    }
    return 0;
}
