// blas1.cuh -- deterministic single-launch reductions (dot products) shared by the Krylov and
// orthonormalisation kernels.  Partials are combined by the last CTA to finish, always in the
// same order, so results are bit-reproducible run to run for a fixed n.
#pragma once
#include "common.cuh"

// scratch layout (bytes): [0,2048) scalar slots | [2048,4096) counters | [4096,65536) dot partials
// | [65536, ...) split-K partials and solver workspaces
// Map of the scalar area [0, SKTT_SCRATCH_BULK_OFF) of the context scratch (zeroed when the scratch is allocated; every user
// below leaves its words in the state the next launch of the same kernel expects; all accesses are ordered by the
// context's stream).  Byte offsets:
//      0 ..  191  reduction results of blas1 / dotc (read back through the mailbox)
//    192 ..  511  convergence flags of the host-driven Krylov loops (krylov.cu)
//    512 ..  1023 Jacobi shift / small scalars of the SVD driver (svd.cu)
//   1024 .. 1031  info word of lu.cu / Cholesky; {fail, passes} of cholqr_kernel (CQ_STATUS_OFF, qr.cu) -- never live at once
//   1088 .. 1095  rank and sweep counter of the SVD driver (svd.cu)
//   1536 .. 2047  phase time stamps of cholqr_kernel (CQ_DEBUG_OFF, debug bit 0)
//   2048 .. 3071  arrival counters of the blas1 / Krylov reductions (SKTT_SCRATCH_COUNTER_OFF), SVD counters at + 256
//   3072 .. 3583  arrival counter of the small right-hand-side kernels (RHS_COUNTER_OFF, stacks.cu)
//   3584 .. 3591  block mask of stack_nat_kernel (cleared by the kernel itself)
//   3592 .. 3595  grid-barrier counter of stack_nat_kernel (self-resetting)
//   3600 .. 3855  phase time stamps of stack_nat_kernel / pcg_persistent_kernel / lu_fused_kernel (debug bit 0)
//   3864 .. 3867  sticky failure word of the deferred QR mode (CQ_STICKY_OFF, qr.cu)
//   4096 .. 65535 per-CTA partial sums of the reductions (SKTT_SCRATCH_PARTIAL_OFF)
#define SKTT_SCRATCH_COUNTER_OFF 2048
#define SKTT_SCRATCH_PARTIAL_OFF 4096
#define SKTT_SCRATCH_BULK_OFF 65536
#define SKTT_DOT_MAX_BLOCKS 1024

template <typename T>
__device__ __forceinline__ T warp_sum(T v);

template <>
__device__ __forceinline__ double warp_sum<double>(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <>
__device__ __forceinline__ cplx warp_sum<cplx>(cplx v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.re += __shfl_xor_sync(0xffffffffu, v.re, o);
        v.im += __shfl_xor_sync(0xffffffffu, v.im, o);
    }
    return v;
}

// broadcast of lane src's value to the whole warp
template <typename T>
__device__ __forceinline__ T lane_bcast(T v, int src);
template <>
__device__ __forceinline__ double lane_bcast<double>(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <>
__device__ __forceinline__ cplx lane_bcast<cplx>(cplx v, int src) {
    return make_cplx(__shfl_sync(0xffffffffu, v.re, src), __shfl_sync(0xffffffffu, v.im, src));
}

// L1-bypassing load (partials written by other SMs in the same launch)
template <typename T>
__device__ __forceinline__ T ld_cg(const T* p);
template <>
__device__ __forceinline__ double ld_cg<double>(const double* p) { return __ldcg(p); }
template <>
__device__ __forceinline__ cplx ld_cg<cplx>(const cplx* p) {
    double2 v = __ldcg((const double2*)p);
    return make_cplx(v.x, v.y);
}

// block-wide sum, result valid in every thread; blockDim.x multiple of 32, <= 1024
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* sh /* >= 32 entries */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum<T>(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    T t = (lane < nw) ? sh[lane] : Num<T>::zero();
    t = warp_sum<T>(t);
    return t;
}

// out[0] = sum_i conj(x[i]) * y[i]   (written as 2 doubles: re, im)
template <typename T>
static __global__ void dot_kernel(long long n, const T* __restrict__ x, const T* __restrict__ y, T* partial,
                                  unsigned* counter, double* out) {
    __shared__ T sh[32];
    __shared__ bool last;
    T acc = Num<T>::zero();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        Num<T>::fma(acc, Num<T>::conj(x[i]), y[i]);
    acc = block_sum<T>(acc, sh);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = acc;
        __threadfence();
        unsigned done = atomicAdd(counter, 1u);
        last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence();
        T s = Num<T>::zero();
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s = Num<T>::add(s, ld_cg<T>(partial + i));
        s = block_sum<T>(s, sh);
        if (threadIdx.x == 0) {
            out[0] = Num<T>::real(s);
            out[1] = Num<T>::imag(s);
            *counter = 0;
        }
    }
}

static inline int blas1_dot(sktt_ctx* ctx, int dtype, long long n, const void* x, const void* y, double* out_dev) {
    long long want = (n + 1023) / 1024;
    int blocks = (int)(want < 1 ? 1 : (want > 2LL * ctx->sm_count ? 2LL * ctx->sm_count : want));
    if (blocks > SKTT_DOT_MAX_BLOCKS) blocks = SKTT_DOT_MAX_BLOCKS;
    unsigned* counter = (unsigned*)((char*)ctx->scratch + SKTT_SCRATCH_COUNTER_OFF);
    void* partial = (char*)ctx->scratch + SKTT_SCRATCH_PARTIAL_OFF;
    if (dtype == SKTT_F64)
        dot_kernel<double><<<blocks, 256, 0, ctx->stream>>>(n, (const double*)x, (const double*)y, (double*)partial,
                                                            counter, out_dev);
    else
        dot_kernel<cplx><<<blocks, 256, 0, ctx->stream>>>(n, (const cplx*)x, (const cplx*)y, (cplx*)partial, counter,
                                                          out_dev);
    SKTT_LAUNCH_CHECK(ctx);
    return 0;
}
