// ctx.cu -- context management and level-1 helpers of the sktt_b200 library.
#include "common.cuh"
#include "blas1.cuh"

extern "C" int sktt_version(void) { return 100; }

extern "C" int sktt_ctx_create(int device, void* cuda_stream, sktt_ctx** out) {
    if (!out) return SKTT_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return SKTT_ERR_CUDA;
    int previous = -1;
    cudaGetDevice(&previous);
    if (cudaSetDevice(device) != cudaSuccess) return SKTT_ERR_CUDA;
    // the caller's current device is restored on every exit path (torch tracks it per thread)
    struct Restore {
        int dev;
        ~Restore() { if (dev >= 0) cudaSetDevice(dev); }
    } restore{previous};
    sktt_ctx* ctx = new sktt_ctx();
    ctx->device = device;
    ctx->stream = (cudaStream_t)cuda_stream;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete ctx;
        return SKTT_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    if (prop.major != 10) {
        // this library is built for sm_100a only; refuse to pretend otherwise
        delete ctx;
        return SKTT_ERR_CUDA;
    }
    if (cudaMallocHost(&ctx->mailbox, 4096) != cudaSuccess) {
        delete ctx;
        return SKTT_ERR_CUDA;
    }
    for (int i = 0; i < 2; ++i)
        if (cudaEventCreateWithFlags(&ctx->ev[i], cudaEventDisableTiming) != cudaSuccess) {
            delete ctx;
            return SKTT_ERR_CUDA;
        }
    *out = ctx;
    return sktt_scratch_reserve(ctx, 8 << 20);
}

extern "C" int sktt_ctx_destroy(sktt_ctx* ctx) {
    if (!ctx) return SKTT_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->lu_flags) cudaFree(ctx->lu_flags);
    if (ctx->mailbox) cudaFreeHost(ctx->mailbox);
    for (int i = 0; i < 2; ++i)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    delete ctx;
    return 0;
}

extern "C" int sktt_ctx_set_stream(sktt_ctx* ctx, void* cuda_stream) {
    if (!ctx) return SKTT_ERR_ARG;
    ctx->stream = (cudaStream_t)cuda_stream;
    return 0;
}

extern "C" const char* sktt_last_error(sktt_ctx* ctx) { return ctx ? ctx->err : "null context"; }
extern "C" int64_t sktt_launch_count(sktt_ctx* ctx) { return ctx ? ctx->launches : -1; }
extern "C" int sktt_ctx_set_gemm_mode(sktt_ctx* ctx, int mode) {
    if (!ctx || mode < 0 || mode > 2) return SKTT_ERR_ARG;
    ctx->gemm_mode = mode;
    return 0;
}

// diagnostics: switch kernel time stamps on/off, read a piece of the scalar scratch area (after a stream sync)
extern "C" int sktt_ctx_set_debug(sktt_ctx* ctx, int on) {
    if (!ctx) return SKTT_ERR_ARG;
    ctx->debug = on;      // bit 0: kernel time stamps, bit 1: Householder QR only, bit 2: no mode preconditioner
    return 0;
}
extern "C" int sktt_scratch_peek(sktt_ctx* ctx, int64_t byte_offset, int64_t bytes, void* out_host) {
    if (!ctx || !out_host || byte_offset < 0 || bytes < 0 || (size_t)(byte_offset + bytes) > ctx->scratch_bytes)
        return SKTT_ERR_ARG;
    SKTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    SKTT_CUDA(ctx, cudaMemcpy(out_host, (char*)ctx->scratch + byte_offset, (size_t)bytes, cudaMemcpyDeviceToHost));
    return 0;
}

// ---------------------------------------------------------------------------------- level 1 -----
template <typename T>
__global__ void axpby_kernel(long long n, T alpha, const T* __restrict__ x, T beta, const T* __restrict__ y,
                             T* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        T v = Num<T>::mul(alpha, x[i]);
        if (y) v = Num<T>::add(v, Num<T>::mul(beta, y[i]));
        out[i] = v;
    }
}

extern "C" int sktt_axpby(sktt_ctx* ctx, int dtype, int64_t n, const double* alpha, const void* x, const double* beta,
                          const void* y, void* out) {
    if (!ctx || !alpha) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (n <= 0) return 0;
    int blocks = (int)((n + 255) / 256 < 8LL * ctx->sm_count ? (n + 255) / 256 : 8LL * ctx->sm_count);
    double b0 = beta ? beta[0] : 1.0, b1 = beta ? beta[1] : 0.0;
    if (dtype == SKTT_F64)
        axpby_kernel<double><<<blocks, 256, 0, ctx->stream>>>(n, alpha[0], (const double*)x, b0, (const double*)y,
                                                              (double*)out);
    else
        axpby_kernel<cplx><<<blocks, 256, 0, ctx->stream>>>(n, make_cplx(alpha[0], alpha[1]), (const cplx*)x,
                                                            make_cplx(b0, b1), (const cplx*)y, (cplx*)out);
    SKTT_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" int sktt_dotc(sktt_ctx* ctx, int dtype, int64_t n, const void* x, const void* y, double* out_host) {
    if (!ctx || !out_host) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    double* slot = (double*)ctx->scratch;  // first 4096 bytes of scratch are reserved for scalars
    SKTT_TRY(blas1_dot(ctx, dtype, n, x, y, slot));
    SKTT_CUDA(ctx, cudaMemcpyAsync(ctx->mailbox, slot, 16, cudaMemcpyDeviceToHost, ctx->stream));
    SKTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    out_host[0] = ((double*)ctx->mailbox)[0];
    out_host[1] = ((double*)ctx->mailbox)[1];
    return 0;
}

extern "C" int sktt_nrm2(sktt_ctx* ctx, int dtype, int64_t n, const void* x, double* out_host) {
    double tmp[2];
    SKTT_TRY(sktt_dotc(ctx, dtype, n, x, x, tmp));
    *out_host = sqrt(tmp[0] > 0 ? tmp[0] : 0.0);
    return 0;
}

__global__ void widen_kernel(long long n, const double* __restrict__ x, cplx* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = make_cplx(x[i], 0.0);
}

extern "C" int sktt_widen(sktt_ctx* ctx, int64_t n, const double* x, void* out_c128) {
    if (!ctx) return SKTT_ERR_ARG;
    if (n <= 0) return 0;
    int blocks = (int)((n + 255) / 256 < 8LL * ctx->sm_count ? (n + 255) / 256 : 8LL * ctx->sm_count);
    widen_kernel<<<blocks, 256, 0, ctx->stream>>>(n, x, (cplx*)out_c128);
    SKTT_LAUNCH_CHECK(ctx);
    return 0;
}
