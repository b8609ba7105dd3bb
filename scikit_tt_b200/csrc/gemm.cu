// gemm.cu -- the contraction engine: two-level strided GEMM kernels for fp64 / complex128.
//
// Every np.tensordot on the reference's ALS path (sle.py:217-219, 274-276, 339-340, 381-383,
// 424-425, 464-466; evp.py:281-292, 323-334, 359-378) is a matrix product over composite
// indices.  numpy realises it as transpose-copy + dgemm; here the composite indices are resolved
// inside the tile loaders (off(i) = (i/d)*s_hi + (i%d)*s_lo), so no operand is ever permuted in
// HBM and intermediates stay in L2.
//
// Kernels
//   gemm_dmma_kernel : fp64, mma.sync.m8n8k4 (SASS DMMA.8x8x4), cp.async multi-stage smem ring.
//                      Measured pipe peak on B200: 37.1 TFLOP/s vs 33.9 for DFMA
//                      (profiles/r01_fp64_peaks.txt) -> used whenever the tile is worth filling.
//   gemm_simt_kernel : fp64 / complex128 register-tiled DFMA, any shape (small ranks, complex).
//   splitk_reduce    : deterministic reduction of split-K partials + alpha/beta epilogue.
#include "common.cuh"
#include "blas1.cuh"

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
struct GemmParams {
    int M, N, K;
    const void* A;
    const void* B;
    void* C;
    Idx2 am, ak, bk, bn, cm, cn;
    int conjA, conjB, conjC;
    double alpha_re, alpha_im, beta_re, beta_im;
    int splits;      // grid.z = batch * splits
    int k_chunk;     // K range handled per split (multiple of BK)
    void* partial;   // [batch][splits][M][N] when splits > 1
    int batch;       // independent problems with the same extents and index maps
    long long sA, sB, sC;   // element strides between consecutive problems of the batch
    int npeer;       // additional destinations of the result (peer GPUs), same offsets as C
    void* Cpeer[7];
};

template <typename T>
__device__ __forceinline__ T ld_elem(const T* p) { return *p; }

template <typename T>
__device__ __forceinline__ T alpha_beta(T acc, T cold, double are, double aim, double bre, double bim, bool use_beta) {
    T a = Num<T>::from(are, aim);
    T r = Num<T>::mul(a, acc);
    if (use_beta) {
        T b = Num<T>::from(bre, bim);
        r = Num<T>::add(r, Num<T>::mul(b, cold));
    }
    return r;
}

// ------------------------------------------------------------------------------------------------
// SIMT kernel (generic): BM x BN tile, BK slab, TM x TN per thread
// ------------------------------------------------------------------------------------------------
template <typename T, int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_simt_kernel(GemmParams p) {
    constexpr int NT = (BM / TM) * (BN / TN);
    __shared__ T As[BK][BM + 1];
    __shared__ T Bs[BK][BN + 1];
    const int bz = blockIdx.z / p.splits, sz = blockIdx.z % p.splits;
    const T* __restrict__ A = (const T*)p.A + (long long)bz * p.sA;
    const T* __restrict__ B = (const T*)p.B + (long long)bz * p.sB;
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = sz * p.k_chunk;
    const int kend = min(p.K, kbeg + p.k_chunk);
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);

    T acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = Num<T>::zero();

    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        // gather A slab: m fastest across threads (the layouts we build are m-contiguous)
        for (int e = tid; e < BM * BK; e += NT) {
            int m = e % BM, k = e / BM;
            T v = Num<T>::zero();
            if (m0 + m < p.M && k0 + k < kend) {
                v = A[p.am(m0 + m) + p.ak(k0 + k)];
                if (p.conjA) v = Num<T>::conj(v);
            }
            As[k][m] = v;
        }
        for (int e = tid; e < BN * BK; e += NT) {
            int n = e % BN, k = e / BN;
            T v = Num<T>::zero();
            if (n0 + n < p.N && k0 + k < kend) {
                v = B[p.bk(k0 + k) + p.bn(n0 + n)];
                if (p.conjB) v = Num<T>::conj(v);
            }
            Bs[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            T a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) Num<T>::fma(acc[i][j], a[i], b[j]);
        }
        __syncthreads();
    }

    if (p.splits > 1) {
        T* P = (T*)p.partial + (size_t)blockIdx.z * p.M * p.N;      // blockIdx.z = bz * splits + sz
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            int m = m0 + ty * TM + i;
            if (m >= p.M) continue;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                int n = n0 + tx * TN + j;
                if (n < p.N) P[(size_t)m * p.N + n] = acc[i][j];
            }
        }
        return;
    }
    T* C = (T*)p.C + (long long)bz * p.sC;
    const bool use_beta = (p.beta_re != 0.0 || p.beta_im != 0.0);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * TM + i;
        if (m >= p.M) continue;
        long long om = p.cm(m);
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = n0 + tx * TN + j;
            if (n >= p.N) continue;
            long long o = om + p.cn(n);
            T cold = use_beta ? C[o] : Num<T>::zero();
            T r = alpha_beta<T>(acc[i][j], cold, p.alpha_re, p.alpha_im, p.beta_re, p.beta_im, use_beta);
            if (p.conjC) r = Num<T>::conj(r);
            C[o] = r;
            for (int q = 0; q < p.npeer; ++q) ((T*)p.Cpeer[q] + (long long)bz * p.sC)[o] = r;
        }
    }
}

template <typename T>
__global__ void splitk_reduce_kernel(GemmParams p) {
    long long total = (long long)p.M * p.N;
    const T* P = (const T*)p.partial + (size_t)blockIdx.y * p.splits * total;
    T* C = (T*)p.C + (long long)blockIdx.y * p.sC;                  // grid.y = batch
    const bool use_beta = (p.beta_re != 0.0 || p.beta_im != 0.0);
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        T s = Num<T>::zero();
        for (int z = 0; z < p.splits; ++z) s = Num<T>::add(s, P[(size_t)z * total + e]);
        int m = (int)(e / p.N), n = (int)(e % p.N);
        long long o = p.cm(m) + p.cn(n);
        T cold = use_beta ? C[o] : Num<T>::zero();
        T r = alpha_beta<T>(s, cold, p.alpha_re, p.alpha_im, p.beta_re, p.beta_im, use_beta);
        if (p.conjC) r = Num<T>::conj(r);
        C[o] = r;
        for (int q = 0; q < p.npeer; ++q) ((T*)p.Cpeer[q] + (long long)blockIdx.y * p.sC)[o] = r;
    }
}

// ------------------------------------------------------------------------------------------------
// DMMA kernel (fp64): BM x BN x 16 tiles, cp.async ring, m8n8k4 tensor MMAs
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// incremental two-level offset for k = k_start, k_start + step, ...
struct KWalk {
    int hi, lo, d;
    long long s_hi, s_lo;
    __device__ __forceinline__ void init(const Idx2& ix, int k) {
        d = (int)min(ix.d, (long long)0x7fffffff);
        hi = k / d;
        lo = k % d;
        s_hi = ix.s_hi;
        s_lo = ix.s_lo;
    }
    __device__ __forceinline__ long long off() const { return hi * s_hi + lo * s_lo; }
    __device__ __forceinline__ void advance(int step) {
        lo += step;
        while (lo >= d) {
            lo -= d;
            ++hi;
        }
    }
};

// 16-byte asynchronous copy with zero fill: `bytes` (0, 8 or 16) are read from global memory, the rest of the 16 is zeroed
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}

// AKM / BKM: the operand is K-MAJOR in memory (unit stride along the contracted index, e.g. a row-major A or a
// column-major B).  Such an operand is staged as [m][k] rows of 16 k (pitch 20 = 4 mod 16: conflict-free fragment loads)
// by 16-byte copies with eight consecutive lanes on one 128-byte row segment; the m-fastest loader below would read it
// with every lane in a different sector (measured: 2048^3 with a row-major A at 20 TFLOP/s against 34 for cuBLAS).
template <int BM, int BN, int WM, int WN, int STAGES, bool AKM, bool BKM>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32)
gemm_dmma_kernel(GemmParams p) {
    constexpr int BK = 16;
    constexpr int NT = (BM / WM) * (BN / WN) * 32;
    constexpr int LDA = BM + 4, LDB = BN + 4;  // (LD mod 16) == 4 -> conflict-free 64-bit fragment loads
    constexpr int LDK = BK + 4;                // pitch of a k-major operand row
    constexpr int A_STAGE = AKM ? BM * LDK : BK * LDA, B_STAGE = BKM ? BN * LDK : BK * LDB;
    constexpr int TPK = NT / BK;               // threads sharing one k row (m-fastest loader)
    constexpr int A_PER = AKM ? (BM * (BK / 2)) / NT : BM / TPK, B_PER = BKM ? (BN * (BK / 2)) / NT : BN / TPK;
    static_assert(NT % BK == 0 && BM % TPK == 0 && BN % TPK == 0, "tile/thread mismatch");
    static_assert((BM * (BK / 2)) % NT == 0 && (BN * (BK / 2)) % NT == 0, "tile/thread mismatch (k-major loader)");
    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + STAGES * A_STAGE;

    const int bz = blockIdx.z / p.splits, sz = blockIdx.z % p.splits;
    const double* __restrict__ A = (const double*)p.A + (long long)bz * p.sA;
    const double* __restrict__ B = (const double*)p.B + (long long)bz * p.sB;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = sz * p.k_chunk;
    const int kend = min(p.K, kbeg + p.k_chunk);
    const int ntiles = (kend - kbeg + BK - 1) / BK;

    // loader roles.  m-fastest: fixed k row (tid / TPK), *_PER m's (n's) strided by TPK.  k-major: fixed 16-byte piece of
    // the k range (tid % 8), *_PER rows strided by NT / 8.
    const int lk = tid / TPK, lt = tid % TPK;
    const int kp = tid & 7, lr = tid >> 3;
    int aoff[A_PER], boff[B_PER];
    unsigned amask = 0, bmask = 0;
#pragma unroll
    for (int j = 0; j < A_PER; ++j) {
        int m = m0 + (AKM ? lr + j * (NT / 8) : lt + j * TPK);
        bool ok = m < p.M;
        aoff[j] = ok ? (int)p.am(m) : 0;
        amask |= (ok ? 1u : 0u) << j;
    }
#pragma unroll
    for (int j = 0; j < B_PER; ++j) {
        int n = n0 + (BKM ? lr + j * (NT / 8) : lt + j * TPK);
        bool ok = n < p.N;
        boff[j] = ok ? (int)p.bn(n) : 0;
        bmask |= (ok ? 1u : 0u) << j;
    }
    KWalk wa, wb;
    wa.init(p.ak, kbeg + (AKM ? 0 : lk));
    wb.init(p.bk, kbeg + (BKM ? 0 : lk));
    int kload = kbeg;                          // first k of the tile being requested

    auto issue = [&](int stage) {
        if (AKM) {
            const int left = kend - (kload + 2 * kp);                 // valid k from this piece on
            const int bytes = left >= 2 ? 16 : (left == 1 ? 8 : 0);
            const double* ap = A + wa.off() + 2 * kp;
            double* as = As + stage * A_STAGE + lr * LDK + 2 * kp;
#pragma unroll
            for (int j = 0; j < A_PER; ++j) {
                const bool ok = bytes > 0 && ((amask >> j) & 1u);
                cp_async16(as + j * (NT / 8) * LDK, ok ? ap + aoff[j] : A, ok ? bytes : 0);
            }
        } else {
            const bool kok = kload + lk < kend;
            const double* ap = A + (kok ? wa.off() : 0);
            double* as = As + stage * A_STAGE + lk * LDA + lt;
#pragma unroll
            for (int j = 0; j < A_PER; ++j) cp_async8(as + j * TPK, ap + aoff[j], kok && ((amask >> j) & 1u));
        }
        if (BKM) {
            const int left = kend - (kload + 2 * kp);
            const int bytes = left >= 2 ? 16 : (left == 1 ? 8 : 0);
            const double* bp = B + wb.off() + 2 * kp;
            double* bs = Bs + stage * B_STAGE + lr * LDK + 2 * kp;
#pragma unroll
            for (int j = 0; j < B_PER; ++j) {
                const bool ok = bytes > 0 && ((bmask >> j) & 1u);
                cp_async16(bs + j * (NT / 8) * LDK, ok ? bp + boff[j] : B, ok ? bytes : 0);
            }
        } else {
            const bool kok = kload + lk < kend;
            const double* bp = B + (kok ? wb.off() : 0);
            double* bs = Bs + stage * B_STAGE + lk * LDB + lt;
#pragma unroll
            for (int j = 0; j < B_PER; ++j) cp_async8(bs + j * TPK, bp + boff[j], kok && ((bmask >> j) & 1u));
        }
        kload += BK;
        if (kload < kend || !AKM) wa.advance(BK);
        if (kload < kend || !BKM) wb.advance(BK);
    };

    // accumulators
    constexpr int MI = WM / 8, NI = WN / 8;
    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int wm0 = (warp % (BM / WM)) * WM, wn0 = (warp / (BM / WM)) * WN;
    const int fr = lane >> 2, fk = lane & 3;  // fragment row/col index and k index

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < ntiles) issue(s);
        cp_async_commit();
    }
    for (int t = 0; t < ntiles; ++t) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        // prefetch tile t + STAGES - 1 into the slot freed at iteration t - 1
        if (t + STAGES - 1 < ntiles) issue((t + STAGES - 1) % STAGES);
        cp_async_commit();
        const double* as = As + (t % STAGES) * A_STAGE + (AKM ? (wm0 + fr) * LDK + fk : fk * LDA + wm0 + fr);
        const double* bs = Bs + (t % STAGES) * B_STAGE + (BKM ? (wn0 + fr) * LDK + fk : fk * LDB + wn0 + fr);
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double af[MI], bf[NI];
#pragma unroll
            for (int i = 0; i < MI; ++i) af[i] = AKM ? as[i * 8 * LDK + kk] : as[kk * LDA + i * 8];
#pragma unroll
            for (int j = 0; j < NI; ++j) bf[j] = BKM ? bs[j * 8 * LDK + kk] : bs[kk * LDB + j * 8];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: thread owns rows wm0 + 8 i + fr, columns wn0 + 8 j + 2 fk + {0,1}
    if (p.splits > 1) {
        double* P = (double*)p.partial + (size_t)blockIdx.z * p.M * p.N;
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            int m = m0 + wm0 + 8 * i + fr;
            if (m >= p.M) continue;
#pragma unroll
            for (int j = 0; j < NI; ++j) {
                int n = n0 + wn0 + 8 * j + 2 * fk;
                if (n < p.N) P[(size_t)m * p.N + n] = acc[i][j][0];
                if (n + 1 < p.N) P[(size_t)m * p.N + n + 1] = acc[i][j][1];
            }
        }
        return;
    }
    double* C = (double*)p.C + (long long)bz * p.sC;
    const bool use_beta = (p.beta_re != 0.0);
    long long cno[NI][2];
#pragma unroll
    for (int j = 0; j < NI; ++j) {
        int n = n0 + wn0 + 8 * j + 2 * fk;
        cno[j][0] = n < p.N ? p.cn(n) : -1;
        cno[j][1] = n + 1 < p.N ? p.cn(n + 1) : -1;
    }
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        int m = m0 + wm0 + 8 * i + fr;
        if (m >= p.M) continue;
        long long om = p.cm(m);
#pragma unroll
        for (int j = 0; j < NI; ++j) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (cno[j][h] < 0) continue;
                long long o = om + cno[j][h];
                double r = p.alpha_re * acc[i][j][h];
                if (use_beta) r = fma(p.beta_re, C[o], r);
                C[o] = r;
                for (int q = 0; q < p.npeer; ++q) ((double*)p.Cpeer[q] + (long long)bz * p.sC)[o] = r;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host dispatch
// ------------------------------------------------------------------------------------------------
int sktt_scratch_reserve(sktt_ctx* ctx, size_t bytes) {
    if (ctx->scratch_bytes >= bytes) return 0;
    size_t want = bytes < (size_t)(8 << 20) ? (size_t)(8 << 20) : bytes + bytes / 4;
    if (ctx->scratch) {
        SKTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        SKTT_CUDA(ctx, cudaFree(ctx->scratch));
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
    }
    SKTT_CUDA(ctx, cudaMalloc(&ctx->scratch, want));
    SKTT_CUDA(ctx, cudaMemsetAsync(ctx->scratch, 0, SKTT_SCRATCH_BULK_OFF, ctx->stream));
    ctx->scratch_bytes = want;
    return 0;
}

static inline long long max_off(const Idx2& ix, long long n) {
    // upper bound of |off(i)| for i < n
    long long hi = (n - 1) / ix.d, lo = (ix.d < n ? ix.d : n) - 1;
    long long a = hi * (ix.s_hi < 0 ? -ix.s_hi : ix.s_hi), b = lo * (ix.s_lo < 0 ? -ix.s_lo : ix.s_lo);
    return a + b;
}

template <int BM, int BN, int WM, int WN, int STAGES, bool AKM, bool BKM>
static int launch_dmma_v(sktt_ctx* ctx, GemmParams& p) {
    constexpr int NT = (BM / WM) * (BN / WN) * 32;
    constexpr size_t smem = (size_t)STAGES * ((AKM ? BM * 20 : 16 * (BM + 4)) + (BKM ? BN * 20 : 16 * (BN + 4))) * sizeof(double);
    SKTT_ONCE_PER_DEVICE(ctx);
    if (!configured) {
        SKTT_CUDA(ctx, cudaFuncSetAttribute(gemm_dmma_kernel<BM, BN, WM, WN, STAGES, AKM, BKM>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, p.splits * p.batch);
    gemm_dmma_kernel<BM, BN, WM, WN, STAGES, AKM, BKM><<<grid, NT, smem, ctx->stream>>>(p);
    SKTT_LAUNCH_CHECK(ctx);
    return 0;
}
template <int BM, int BN, int WM, int WN, int STAGES>
static int launch_dmma(sktt_ctx* ctx, GemmParams& p, bool akm, bool bkm) {
    if (akm && bkm) return launch_dmma_v<BM, BN, WM, WN, STAGES, true, true>(ctx, p);
    if (akm) return launch_dmma_v<BM, BN, WM, WN, STAGES, true, false>(ctx, p);
    if (bkm) return launch_dmma_v<BM, BN, WM, WN, STAGES, false, true>(ctx, p);
    return launch_dmma_v<BM, BN, WM, WN, STAGES, false, false>(ctx, p);
}

// An operand can take the k-major loader when its contracted index has unit stride in runs that a 16-k tile never straddles,
// every 16-byte piece is aligned (even row offsets, aligned base, even batch stride) and the other index is NOT the
// unit-stride one (then the m-fastest loader is already coalesced).
static bool kmajor_ok(const Idx2& ik, const Idx2& im, long long K, long long M, const void* ptr, long long batch_stride) {
    if (ik.s_lo != 1 || !(ik.d >= K || ik.d % 16 == 0)) return false;
    if (ik.d < K && (ik.s_hi & 1)) return false;
    if (im.s_lo == 1 && im.d >= 8) return false;
    if ((im.s_lo & 1) && M > 1 && im.d > 1) return false;
    if (im.d < M && (im.s_hi & 1)) return false;
    return ((uintptr_t)ptr & 15u) == 0 && (batch_stride & 1) == 0;
}

template <typename T, int BM, int BN, int BK, int TM, int TN>
static int launch_simt(sktt_ctx* ctx, GemmParams& p) {
    dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, p.splits * p.batch);
    gemm_simt_kernel<T, BM, BN, BK, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, ctx->stream>>>(p);
    SKTT_LAUNCH_CHECK(ctx);
    return 0;
}

static inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

int sktt_gemm_run(sktt_ctx* ctx, int dtype, const GemmDesc& g) {
    if (g.M <= 0 || g.N <= 0) return 0;
    if (g.M > 0x7fffffffLL || g.N > 0x7fffffffLL || g.K > 0x7fffffffLL)
        return sktt_fail(ctx, SKTT_ERR_ARG, "gemm extent exceeds int32");
    GemmParams p;
    p.M = (int)g.M; p.N = (int)g.N; p.K = (int)g.K;
    p.A = g.A; p.B = g.B; p.C = g.C;
    p.am = g.am; p.ak = g.ak; p.bk = g.bk; p.bn = g.bn; p.cm = g.cm; p.cn = g.cn;
    p.conjA = g.conjA; p.conjB = g.conjB; p.conjC = g.conjC;
    p.alpha_re = g.alpha[0]; p.alpha_im = g.alpha[1];
    p.beta_re = g.beta[0]; p.beta_im = g.beta[1];
    p.splits = 1;
    p.k_chunk = p.K > 0 ? p.K : 1;
    p.partial = nullptr;
    p.batch = g.batch > 1 ? (int)g.batch : 1;
    p.sA = g.sA; p.sB = g.sB; p.sC = g.sC;
    p.npeer = g.npeer < 0 ? 0 : (g.npeer > 7 ? 7 : g.npeer);
    for (int q = 0; q < 7; ++q) p.Cpeer[q] = g.Cpeer[q];
    if ((long long)p.batch > 65535) return sktt_fail(ctx, SKTT_ERR_ARG, "gemm batch exceeds 65535");

    const bool is_f64 = dtype == SKTT_F64;
    const int sms = ctx->sm_count;
    // DMMA path: real fp64, offsets representable in int32, problem large enough to fill tiles
    bool off32 = max_off(g.am, g.M) + max_off(g.ak, g.K ? g.K : 1) < 0x7fffffffLL &&
                 max_off(g.bk, g.K ? g.K : 1) + max_off(g.bn, g.N) < 0x7fffffffLL;
    bool dmma_ok = is_f64 && off32 && ctx->gemm_mode != 1 &&
                   (ctx->gemm_mode == 2 || (g.M >= 32 && g.N >= 32 && g.K >= 16 && g.M * g.N * g.K >= (1LL << 18)));

    if (dmma_ok) {
        long long t128 = cdiv(g.M, 128) * cdiv(g.N, 128);
        long long t64 = cdiv(g.M, 64) * cdiv(g.N, 64);
        bool big = t128 * p.batch >= (long long)(sms * 3) / 4;
        long long tiles = (big ? t128 : t64) * p.batch;
        // split-K when the tile grid cannot fill the machine and K is long
        if (tiles < sms && g.K >= 512) {
            long long want = cdiv(2LL * sms, tiles);
            long long maxs = g.K / 128;
            long long s = want < maxs ? want : maxs;
            if (s > 1) {
                p.k_chunk = (int)(cdiv(cdiv(g.K, s), 16) * 16);
                p.splits = (int)cdiv(g.K, p.k_chunk);
            }
        }
        if (p.splits > 1) {
            SKTT_TRY(sktt_scratch_reserve(ctx, (size_t)p.batch * p.splits * g.M * g.N * sizeof(double) + SKTT_SCRATCH_BULK_OFF));
            p.partial = (char*)ctx->scratch + SKTT_SCRATCH_BULK_OFF;
        }
        const bool akm = kmajor_ok(g.ak, g.am, g.K, g.M, g.A, p.sA), bkm = kmajor_ok(g.bk, g.bn, g.K, g.N, g.B, p.sB);
        // 128 x 128 tiles: sixteen warps of 32 x 32 (64 accumulator registers per thread: a 64 x 32 warp tile needs 216
        // registers, i.e. eight warps per SM -- too few to cover the LDS -> DMMA latency; measured 0.5 of the pipe peak)
        if (big) SKTT_TRY((launch_dmma<128, 128, 32, 32, 4>(ctx, p, akm, bkm)));
        else SKTT_TRY((launch_dmma<64, 64, 32, 32, 4>(ctx, p, akm, bkm)));
        if (p.splits > 1) {
            long long total = g.M * g.N;
            int blocks = (int)(cdiv(total, 256) < 4LL * sms ? cdiv(total, 256) : 4LL * sms);
            splitk_reduce_kernel<double><<<dim3(blocks, p.batch), 256, 0, ctx->stream>>>(p);
            SKTT_LAUNCH_CHECK(ctx);
        }
        return 0;
    }

    // SIMT path
    bool small = (g.M <= 16 || g.N <= 16);
    long long tiles = (small ? cdiv(g.M, 16) * cdiv(g.N, 16) : cdiv(g.M, 64) * cdiv(g.N, 64)) * p.batch;
    if (tiles < sms && g.K >= 256) {
        long long want = cdiv(2LL * sms, tiles);
        long long maxs = g.K / 64;
        long long s = want < maxs ? want : maxs;
        if (s > 1) {
            p.k_chunk = (int)(cdiv(cdiv(g.K, s), 16) * 16);
            p.splits = (int)cdiv(g.K, p.k_chunk);
        }
    }
    if (p.splits > 1) {
        SKTT_TRY(sktt_scratch_reserve(ctx, (size_t)p.batch * p.splits * g.M * g.N * dtype_size(dtype) + SKTT_SCRATCH_BULK_OFF));
        p.partial = (char*)ctx->scratch + SKTT_SCRATCH_BULK_OFF;
    }
    if (is_f64) {
        if (small) SKTT_TRY((launch_simt<double, 16, 16, 16, 1, 1>(ctx, p)));
        else SKTT_TRY((launch_simt<double, 64, 64, 16, 4, 4>(ctx, p)));
    } else {
        if (small) SKTT_TRY((launch_simt<cplx, 16, 16, 16, 1, 1>(ctx, p)));
        else SKTT_TRY((launch_simt<cplx, 64, 64, 8, 4, 4>(ctx, p)));
    }
    if (p.splits > 1) {
        long long total = g.M * g.N;
        int blocks = (int)(cdiv(total, 256) < 4LL * sms ? cdiv(total, 256) : 4LL * sms);
        if (is_f64) splitk_reduce_kernel<double><<<dim3(blocks, p.batch), 256, 0, ctx->stream>>>(p);
        else splitk_reduce_kernel<cplx><<<dim3(blocks, p.batch), 256, 0, ctx->stream>>>(p);
        SKTT_LAUNCH_CHECK(ctx);
    }
    return 0;
}

extern "C" int sktt_gemm2(sktt_ctx* ctx, int dtype, int64_t M, int64_t N, int64_t K, const double* alpha,
                          const void* A, sktt_idx2 am, sktt_idx2 ak, int conjA, const void* B, sktt_idx2 bk,
                          sktt_idx2 bn, int conjB, const double* beta, void* C, sktt_idx2 cm, sktt_idx2 cn) {
    if (!ctx) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (M < 0 || N < 0 || K < 0) return sktt_fail(ctx, SKTT_ERR_ARG, "negative gemm extent");
    GemmDesc g = gemm_desc(M, N, K, A, from_abi(am), from_abi(ak), B, from_abi(bk), from_abi(bn), C, from_abi(cm),
                           from_abi(cn));
    g.conjA = conjA;
    g.conjB = conjB;
    if (alpha) { g.alpha[0] = alpha[0]; g.alpha[1] = alpha[1]; }
    if (beta) { g.beta[0] = beta[0]; g.beta[1] = beta[1]; }
    return sktt_gemm_run(ctx, dtype, g);
}
