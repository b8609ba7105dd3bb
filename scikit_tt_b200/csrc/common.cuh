// common.cuh -- shared declarations of the sktt_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/sktt_b200.h"

struct sktt_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
    int gemm_mode = 0;
    int debug = 0;          // kernels that support it leave phase time stamps in the scalar scratch area
    int sm_count = 148;
    char err[512] = {0};
    // small persistent device scratch (scalars, flags, split-K partials)
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    // pinned host mailbox for scalar read-backs
    void* mailbox = nullptr;
    // two reusable events for the lagged convergence checks of the Krylov loops
    cudaEvent_t ev[2] = {nullptr, nullptr};
    // panel flags of the one-launch LU (lu_fused.cu): "published" means flags[k] == lu_epoch of the running launch
    void* lu_flags = nullptr;
    unsigned lu_epoch = 0;
    // deferred mode of the tall QR / RQ (qr.cu): no Householder launch behind the failure flag of the sketched CholeskyQR
    int qr_deferred = 0;
};

static inline int sktt_fail(sktt_ctx* ctx, int code, const char* fmt, const char* a = "") {
    if (ctx) snprintf(ctx->err, sizeof(ctx->err), fmt, a);
    return code;
}

#define SKTT_CUDA(ctx, call)                                                                     \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            if (ctx)                                                                             \
                snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d CUDA error: %s", __FILE__,       \
                         __LINE__, cudaGetErrorString(e__));                                     \
            return SKTT_ERR_CUDA;                                                                \
        }                                                                                        \
    } while (0)

#define SKTT_LAUNCH_CHECK(ctx)                                                                   \
    do {                                                                                         \
        (ctx)->launches++;                                                                       \
        SKTT_CUDA(ctx, cudaGetLastError());                                                      \
    } while (0)

// cudaFuncSetAttribute opt-ins (dynamic shared memory above 48 KB, cluster sizes) are per device: one flag per device
// index, so that a second GPU driven from the same process gets its own opt-in (one context = one device)
#define SKTT_MAX_DEVICES 64
#define SKTT_ONCE_PER_DEVICE(ctx)                                                                \
    static bool once_per_device_[SKTT_MAX_DEVICES] = {false};                                    \
    bool& configured = once_per_device_[(ctx)->device & (SKTT_MAX_DEVICES - 1)]

#define SKTT_TRY(expr)                                                                           \
    do {                                                                                         \
        int s__ = (expr);                                                                        \
        if (s__ != 0) return s__;                                                                \
    } while (0)

// ---------------------------------------------------------------- complex128 arithmetic -------
struct __align__(16) cplx {
    double re, im;
};

__host__ __device__ __forceinline__ cplx make_cplx(double a, double b) {
    cplx z;
    z.re = a;
    z.im = b;
    return z;
}

template <typename T>
struct Num;

template <>
struct Num<double> {
    static constexpr bool is_complex = false;
    __host__ __device__ static __forceinline__ double zero() { return 0.0; }
    __host__ __device__ static __forceinline__ double one() { return 1.0; }
    __host__ __device__ static __forceinline__ double from(double re, double) { return re; }
    __host__ __device__ static __forceinline__ double conj(double a) { return a; }
    __host__ __device__ static __forceinline__ double real(double a) { return a; }
    __host__ __device__ static __forceinline__ double imag(double) { return 0.0; }
    __host__ __device__ static __forceinline__ double abs2(double a) { return a * a; }
    __device__ static __forceinline__ void fma(double& c, double a, double b) { c = ::fma(a, b, c); }
    __host__ __device__ static __forceinline__ double mul(double a, double b) { return a * b; }
    __host__ __device__ static __forceinline__ double add(double a, double b) { return a + b; }
    __host__ __device__ static __forceinline__ double sub(double a, double b) { return a - b; }
    __host__ __device__ static __forceinline__ double scale(double a, double s) { return a * s; }
    __host__ __device__ static __forceinline__ double div(double a, double b) { return a / b; }
    __host__ __device__ static __forceinline__ double neg(double a) { return -a; }
};

template <>
struct Num<cplx> {
    static constexpr bool is_complex = true;
    __host__ __device__ static __forceinline__ cplx zero() { return make_cplx(0.0, 0.0); }
    __host__ __device__ static __forceinline__ cplx one() { return make_cplx(1.0, 0.0); }
    __host__ __device__ static __forceinline__ cplx from(double re, double im) { return make_cplx(re, im); }
    __host__ __device__ static __forceinline__ cplx conj(cplx a) { return make_cplx(a.re, -a.im); }
    __host__ __device__ static __forceinline__ double real(cplx a) { return a.re; }
    __host__ __device__ static __forceinline__ double imag(cplx a) { return a.im; }
    __host__ __device__ static __forceinline__ double abs2(cplx a) { return a.re * a.re + a.im * a.im; }
    __device__ static __forceinline__ void fma(cplx& c, cplx a, cplx b) {
        c.re = ::fma(a.re, b.re, c.re);
        c.re = ::fma(-a.im, b.im, c.re);
        c.im = ::fma(a.re, b.im, c.im);
        c.im = ::fma(a.im, b.re, c.im);
    }
    __host__ __device__ static __forceinline__ cplx mul(cplx a, cplx b) {
        return make_cplx(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
    }
    __host__ __device__ static __forceinline__ cplx add(cplx a, cplx b) { return make_cplx(a.re + b.re, a.im + b.im); }
    __host__ __device__ static __forceinline__ cplx sub(cplx a, cplx b) { return make_cplx(a.re - b.re, a.im - b.im); }
    __host__ __device__ static __forceinline__ cplx scale(cplx a, double s) { return make_cplx(a.re * s, a.im * s); }
    __host__ __device__ static __forceinline__ cplx div(cplx a, cplx b) {
        // Smith's algorithm (as LAPACK zladiv in spirit): avoids overflow of |b|^2
        if (fabs(b.re) >= fabs(b.im)) {
            double t = b.im / b.re, d = b.re + b.im * t;
            return make_cplx((a.re + a.im * t) / d, (a.im - a.re * t) / d);
        } else {
            double t = b.re / b.im, d = b.re * t + b.im;
            return make_cplx((a.re * t + a.im) / d, (a.im * t - a.re) / d);
        }
    }
    __host__ __device__ static __forceinline__ cplx neg(cplx a) { return make_cplx(-a.re, -a.im); }
};

// two-level index map, see sktt_idx2
struct Idx2 {
    long long d, s_hi, s_lo;
    __host__ __device__ __forceinline__ long long operator()(long long i) const {
        return (i / d) * s_hi + (i % d) * s_lo;
    }
};
static inline Idx2 mk_idx(long long d, long long s_hi, long long s_lo) {
    Idx2 x;
    x.d = d > 0 ? d : 1;
    x.s_hi = s_hi;
    x.s_lo = s_lo;
    return x;
}
static inline Idx2 lin_idx(long long stride) { return mk_idx(1LL << 40, 0, stride); }
static inline Idx2 from_abi(sktt_idx2 a) { return mk_idx(a.d, a.s_hi, a.s_lo); }

// ---------------------------------------------------------------- internal GEMM interface -----
// C[cm(i)+cn(j)] = alpha * sum_k opA(A[am(i)+ak(k)]) opB(B[bk(k)+bn(j)]) + beta * C[...]
struct GemmDesc {
    long long M, N, K;
    const void* A;
    Idx2 am, ak;
    int conjA;
    const void* B;
    Idx2 bk, bn;
    int conjB;
    void* C;
    Idx2 cm, cn;
    double alpha[2];
    double beta[2];
    int conjC;  // store conj(result) (used to fold conjugations of the reference into epilogues)
    long long batch;        // > 1: that many independent problems with identical extents and index maps ...
    long long sA, sB, sC;   // ... whose operands lie sA / sB / sC elements apart
    int npeer;              // > 0: the result is ALSO stored through these pointers (same offsets) -- peer-mapped copies of
    void* Cpeer[7];         // C on other GPUs, written from the epilogue over NVLink (fused all-gather, peer.cu)
};

static inline GemmDesc gemm_desc(long long M, long long N, long long K, const void* A, Idx2 am, Idx2 ak,
                                 const void* B, Idx2 bk, Idx2 bn, void* C, Idx2 cm, Idx2 cn) {
    GemmDesc g;
    g.M = M; g.N = N; g.K = K;
    g.A = A; g.am = am; g.ak = ak; g.conjA = 0;
    g.B = B; g.bk = bk; g.bn = bn; g.conjB = 0;
    g.C = C; g.cm = cm; g.cn = cn;
    g.alpha[0] = 1.0; g.alpha[1] = 0.0;
    g.beta[0] = 0.0; g.beta[1] = 0.0;
    g.conjC = 0;
    g.batch = 1; g.sA = g.sB = g.sC = 0;
    g.npeer = 0;
    for (int q = 0; q < 7; ++q) g.Cpeer[q] = nullptr;
    return g;
}

int sktt_gemm_run(sktt_ctx* ctx, int dtype, const GemmDesc& g);

static inline size_t dtype_size(int dtype) { return dtype == SKTT_C128 ? 16 : 8; }
static inline int check_dtype(sktt_ctx* ctx, int dtype) {
    if (dtype != SKTT_F64 && dtype != SKTT_C128) return sktt_fail(ctx, SKTT_ERR_DTYPE, "unknown dtype");
    return 0;
}

int sktt_scratch_reserve(sktt_ctx* ctx, size_t bytes);

// The fused matvec / stack kernels (fused.cu) work on an input-side solution rank padded to a multiple of 4 and on an
// output side padded to (r2, R2) = (64, 3); the image / tiling kernels fill the padding with zeros, so edge cores
// (r = 1 or r2 = 1) and small operator ranks run through the same kernels.
static inline long long fused_rpad(long long r) { return (r + 3) & ~3LL; }
