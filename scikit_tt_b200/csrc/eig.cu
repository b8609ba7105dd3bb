// eig.cu -- local eigen-solves of evp.als for general (non-Hermitian) micro matrices:
// scipy.linalg.eig + "closest to sigma" selection (evp.py:424-432) and
// scipy.sparse.linalg.eigs(sigma=..., v0=ones) (evp.py:417-422).  Both are served by shift-invert
// Arnoldi: LU of (M - sigma B) from lu.cu, Krylov basis orthogonalised with CGS2 through the
// contraction engine, and the small projected Hessenberg eigenproblem solved on the device by a
// complex single-shift QR iteration (one warp, rotations applied lane-parallel), so no part of the
// eigen-solve runs on the host.
#include "common.cuh"
#include "blas1.cuh"
#include "hess_eig.cuh"

#define EIG_MAX_NCV 1024   // ncv == N is the exact fallback (the Krylov space is the whole space) for small micro matrices
#define EIG_SMEM_NCV 64    // up to here the projected eigenproblem is solved in shared memory

// ------------------------------------------------------------------------------------------------
// small complex Hessenberg eigen-solver (single warp).  Hin: m x m upper Hessenberg, row-major
// complex.  Outputs: theta[m], Y[m][m] (column i = unit-norm eigenvector i of Hin).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
hess_eig_kernel(int m, const cplx* __restrict__ Hin, cplx* __restrict__ theta, cplx* __restrict__ Yout, int* info,
                cplx* gwork /* 3 m^2 elements in global memory, or null: shared memory */) {
    extern __shared__ unsigned char smem_raw[];
    cplx* H = gwork ? gwork : (cplx*)smem_raw;   // [m][m]
    cplx* Z = H + (size_t)m * m;                 // [m][m]
    cplx* X = Z + (size_t)m * m;                 // [m][m]  eigenvectors of T (columns)
    hess_eig_warp(m, Hin, theta, Yout, info, H, Z, X);
}

// order Ritz values by |theta| descending, keep k; out: sel[k] indices, lam[k] = sigma + 1/theta,
// conv flag, Ysel [m][k] complex (coefficients of the wanted Ritz vectors)
__global__ void ritz_select_kernel(int m, int k, double sigma, const cplx* __restrict__ theta,
                                   const cplx* __restrict__ Y, const double* hnext2 /* |h_{m+1,m}|^2 or null */,
                                   double tol, cplx* __restrict__ lam, cplx* __restrict__ Ysel, int* nconv) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    typedef Num<cplx> C;
    int conv = 0;
    const double hn = hnext2 ? sqrt(hnext2[0] > 0.0 ? hnext2[0] : 0.0) : 0.0;
    int taken[64];                      // k <= 64 (k <= ncv, and only a handful of pairs is ever asked for)
    for (int s = 0; s < k; ++s) {
        int best = -1;
        double bv = -1.0;
        for (int i = 0; i < m; ++i) {
            bool used = false;
            for (int t = 0; t < s; ++t) used |= (taken[t] == i);
            if (used) continue;
            double v = cabs_(theta[i]);
            if (v > bv) { bv = v; best = i; }
        }
        taken[s] = best;
        cplx th = theta[best];
        lam[s] = C::add(make_cplx(sigma, 0.0), C::div(C::one(), th));
        for (int r = 0; r < m; ++r) Ysel[r * k + s] = Y[r * m + best];
        double res = hn * cabs_(Y[(m - 1) * m + best]);
        // a Ritz value of exactly zero is the footprint of an exhausted Krylov space (Arnoldi breakdown zeroes the next
        // basis vector): lambda = sigma + 1/0 is not an eigenvalue, never count it as converged
        if (bv > 0.0 && res <= tol * bv) conv++;
    }
    *nconv = conv;
}

// rotate each column so that its largest component is real and positive (LAPACK geev convention)
__global__ void __launch_bounds__(256) phase_fix_kernel(long long N, int k, cplx* __restrict__ vecs) {
    __shared__ double sval[8];
    __shared__ long long sidx[8];
    __shared__ cplx rot;
    const int col = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double best = -1.0;
    long long bi = 0;
    for (long long i = threadIdx.x; i < N; i += 256) {
        double a = Num<cplx>::abs2(vecs[i * k + col]);
        if (a > best) { best = a; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        double ob = __shfl_xor_sync(0xffffffffu, best, o);
        long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { sval[warp] = best; sidx[warp] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            if (sval[w] > best || (sval[w] == best && sidx[w] < bi)) { best = sval[w]; bi = sidx[w]; }
        cplx v = vecs[bi * k + col];
        double a = hypot(v.re, v.im);
        rot = a > 0.0 ? make_cplx(v.re / a, -v.im / a) : make_cplx(1.0, 0.0);
    }
    __syncthreads();
    const cplx rr = rot;
    for (long long i = threadIdx.x; i < N; i += 256) vecs[i * k + col] = Num<cplx>::mul(vecs[i * k + col], rr);
}

template <typename T>
__global__ void fill_kernel(long long n, T* x, T v) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] = v;
}

// Mat -= sigma * B (B == null: identity)
template <typename T>
__global__ void shift_kernel(long long N, double sigma, const T* __restrict__ B, T* __restrict__ M) {
    long long total = B ? N * N : N;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        if (B) M[e] = Num<T>::sub(M[e], Num<T>::scale(B[e], sigma));
        else M[e * N + e] = Num<T>::sub(M[e * N + e], Num<T>::from(sigma, 0.0));
    }
}

template <typename T>
__global__ void to_cplx_kernel(long long n, const T* __restrict__ x, cplx* __restrict__ out);
template <>
__global__ void to_cplx_kernel<double>(long long n, const double* __restrict__ x, cplx* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = make_cplx(x[i], 0.0);
}
template <>
__global__ void to_cplx_kernel<cplx>(long long n, const cplx* __restrict__ x, cplx* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = x[i];
}

template <typename T>
__global__ void scale_unit_kernel(long long n, const T* __restrict__ w, const double* nrm2, T* __restrict__ v) {
    const double nrm = sqrt(nrm2[0] > 0.0 ? nrm2[0] : 0.0);
    const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        v[i] = Num<T>::scale(w[i], inv);
}

// H[(0..j), j] = h1 + h2 ; H[j+1, j] = sqrt(nrm2)   (row-major m x m complex, last row kept separately)
template <typename T>
__global__ void hess_store_kernel(int j, int m, const T* __restrict__ h1, const T* __restrict__ h2, const double* nrm2,
                                  cplx* __restrict__ H) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= j) {
        T v = Num<T>::add(h1[i], h2[i]);
        H[i * m + j] = make_cplx(Num<T>::real(v), Num<T>::imag(v));
    }
    if (i == j + 1 && j + 1 < m) H[(j + 1) * m + j] = make_cplx(sqrt(nrm2[0] > 0.0 ? nrm2[0] : 0.0), 0.0);
}

extern "C" int64_t sktt_eig_si_work(int64_t N, int64_t k, int64_t ncv) {
    if (ncv > EIG_MAX_NCV) ncv = EIG_MAX_NCV;
    // V [(ncv+1) N] + w [N] + t [N] (in units of complex128 elements, generous for real T)
    return (ncv + 3) * N + 8 * ncv * ncv + 4 * ncv + ncv * k + 2 * N + 512;
}

template <typename T>
static int eig_si_impl(sktt_ctx* ctx, int dtype, long long N, T* Mat, const T* Bmat, double sigma, int k, int ncv,
                       double tol, int max_restarts, cplx* lam, cplx* vecs, void* work_raw, int* nconv_host) {
    if (ncv > EIG_MAX_NCV) ncv = EIG_MAX_NCV;
    if (ncv > N) ncv = (int)N;
    if (k > ncv) return sktt_fail(ctx, SKTT_ERR_ARG, "eig_shift_invert: k exceeds the Krylov dimension");
    const int m = ncv;
    // workspace carve-up
    T* V = (T*)work_raw;                        // [(m+1)][N]
    T* w = V + (size_t)(m + 1) * N;             // [N]
    T* t = w + N;                               // [N]
    T* h1 = t + N;                              // [m+1]
    T* h2 = h1 + (m + 1);                       // [m+1]
    cplx* H = (cplx*)((((uintptr_t)(h2 + (m + 1))) + 15) & ~(uintptr_t)15);  // [m][m], 16-byte aligned
    cplx* Y = H + (size_t)m * m;                // [m][m]
    cplx* theta = Y + (size_t)m * m;            // [m]
    cplx* Ysel = theta + m;                     // [m][k]
    int* ipiv;
    // pivots + permutation live after Ysel
    ipiv = (int*)(Ysel + (size_t)m * k);
    int* flags = ipiv + 2 * N;                  // [0] nconv, [1] hess info
    cplx* hwork = (cplx*)((((uintptr_t)(flags + 8)) + 15) & ~(uintptr_t)15);   // [3][m][m], used when m > EIG_SMEM_NCV
    // scalar slots [0..1] (nrm2) at the head of the context scratch: re-derived at every use, because the LU / GEMM calls in
    // between may grow the scratch allocation (sktt_scratch_reserve frees and reallocates it)
#define slots ((double*)ctx->scratch)
    const int nbk = (int)((N + 255) / 256 < 4LL * ctx->sm_count ? (N + 255) / 256 : 4LL * ctx->sm_count);

    // S = M - sigma B, LU
    {
        long long total = Bmat ? N * N : N;
        int blocks = (int)((total + 255) / 256 < 1024 ? (total + 255) / 256 : 1024);
        shift_kernel<T><<<blocks, 256, 0, ctx->stream>>>(N, sigma, Bmat, Mat);
        SKTT_LAUNCH_CHECK(ctx);
    }
    int info = 0;
    SKTT_TRY(sktt_lu_factor(ctx, dtype, N, Mat, ipiv, &info));
    if (info != 0) return sktt_fail(ctx, SKTT_ERR_SINGULAR, "eig_shift_invert: (M - sigma B) is exactly singular");

    SKTT_ONCE_PER_DEVICE(ctx);
    if (!configured) {
        SKTT_CUDA(ctx, cudaFuncSetAttribute(hess_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    // start vector: ones (evp.py:418), normalised
    fill_kernel<T><<<nbk, 256, 0, ctx->stream>>>(N, w, Num<T>::one());
    SKTT_LAUNCH_CHECK(ctx);
    int nconv = 0;
    for (int restart = 0; restart <= max_restarts; ++restart) {
        SKTT_TRY(blas1_dot(ctx, dtype, N, w, w, slots));
        scale_unit_kernel<T><<<nbk, 256, 0, ctx->stream>>>(N, w, slots, V);
        SKTT_LAUNCH_CHECK(ctx);
        SKTT_CUDA(ctx, cudaMemsetAsync(H, 0, (size_t)m * m * sizeof(cplx), ctx->stream));
        for (int j = 0; j < m; ++j) {
            T* vj = V + (size_t)j * N;
            // w = S^{-1} (B vj)
            if (Bmat) {
                GemmDesc gb = gemm_desc(N, 1, N, Bmat, lin_idx(N), lin_idx(1), vj, lin_idx(1), lin_idx(0), w, lin_idx(1),
                                        lin_idx(0));
                SKTT_TRY(sktt_gemm_run(ctx, dtype, gb));
            } else {
                SKTT_CUDA(ctx, cudaMemcpyAsync(w, vj, (size_t)N * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
            }
            SKTT_TRY(sktt_lu_solve(ctx, dtype, N, 1, Mat, ipiv, w));
            for (int pass = 0; pass < 2; ++pass) {
                T* hd = pass == 0 ? h1 : h2;
                GemmDesc g1 = gemm_desc(j + 1, 1, N, V, lin_idx(N), lin_idx(1), w, lin_idx(1), lin_idx(0), hd, lin_idx(1),
                                        lin_idx(0));
                g1.conjA = 1;
                SKTT_TRY(sktt_gemm_run(ctx, dtype, g1));
                GemmDesc g2 = gemm_desc(N, 1, j + 1, V, lin_idx(1), lin_idx(N), hd, lin_idx(1), lin_idx(0), w, lin_idx(1),
                                        lin_idx(0));
                g2.alpha[0] = -1.0;
                g2.beta[0] = 1.0;
                SKTT_TRY(sktt_gemm_run(ctx, dtype, g2));
            }
            SKTT_TRY(blas1_dot(ctx, dtype, N, w, w, slots));
            hess_store_kernel<T><<<(j + 2 + 127) / 128, 128, 0, ctx->stream>>>(j, m, h1, h2, slots, H);
            SKTT_LAUNCH_CHECK(ctx);
            scale_unit_kernel<T><<<nbk, 256, 0, ctx->stream>>>(N, w, slots, V + (size_t)(j + 1) * N);
            SKTT_LAUNCH_CHECK(ctx);
        }
        // projected eigenproblem, selection, convergence
        if (m <= EIG_SMEM_NCV) {
            size_t hsmem = (size_t)3 * m * m * sizeof(cplx);
            hess_eig_kernel<<<1, 32, hsmem, ctx->stream>>>(m, H, theta, Y, flags + 1, (cplx*)nullptr);
        } else {
            hess_eig_kernel<<<1, 32, 0, ctx->stream>>>(m, H, theta, Y, flags + 1, hwork);
        }
        SKTT_LAUNCH_CHECK(ctx);
        ritz_select_kernel<<<1, 32, 0, ctx->stream>>>(m, k, sigma, theta, Y, m < N ? slots : (const double*)nullptr, tol,
                                                      lam, Ysel, flags);
        SKTT_LAUNCH_CHECK(ctx);
        SKTT_CUDA(ctx, cudaMemcpyAsync(ctx->mailbox, flags, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        SKTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        nconv = ((int*)ctx->mailbox)[0];
        if (((int*)ctx->mailbox)[1] != 0) return sktt_fail(ctx, SKTT_ERR_NOCONV, "eig_shift_invert: Hessenberg QR did not converge");
        // Ritz vectors: vecs (N x k complex) = V_m (N x m) Ysel (m x k)
        if (dtype == SKTT_C128) {
            GemmDesc gv = gemm_desc(N, k, m, V, lin_idx(1), lin_idx(N), Ysel, lin_idx(k), lin_idx(1), vecs, lin_idx(k),
                                    lin_idx(1));
            SKTT_TRY(sktt_gemm_run(ctx, SKTT_C128, gv));
        } else {
            for (int part = 0; part < 2; ++part) {
                GemmDesc gv = gemm_desc(N, k, m, V, lin_idx(1), lin_idx(N), (const double*)Ysel + part, lin_idx(2 * k),
                                        lin_idx(2), (double*)vecs + part, lin_idx(2 * k), lin_idx(2));
                SKTT_TRY(sktt_gemm_run(ctx, SKTT_F64, gv));
            }
        }
        if (nconv >= k || m >= N || restart == max_restarts) break;
        // explicit restart from the sum of the wanted Ritz vectors (real part for real problems)
        {
            double onev[2] = {1.0, 0.0};
            if (dtype == SKTT_C128) {
                GemmDesc gs = gemm_desc(N, 1, k, vecs, lin_idx(k), lin_idx(1), Ysel, lin_idx(0), lin_idx(0), w, lin_idx(1),
                                        lin_idx(0));
                (void)gs;
                // w = sum_s vecs[:, s]
                fill_kernel<T><<<1, 32, 0, ctx->stream>>>(k, h1, Num<T>::one());
                SKTT_LAUNCH_CHECK(ctx);
                GemmDesc gs2 = gemm_desc(N, 1, k, vecs, lin_idx(k), lin_idx(1), h1, lin_idx(1), lin_idx(0), w, lin_idx(1),
                                         lin_idx(0));
                SKTT_TRY(sktt_gemm_run(ctx, SKTT_C128, gs2));
            } else {
                fill_kernel<T><<<1, 32, 0, ctx->stream>>>(k, h1, Num<T>::one());
                SKTT_LAUNCH_CHECK(ctx);
                GemmDesc gs2 = gemm_desc(N, 1, k, (const double*)vecs, lin_idx(2 * k), lin_idx(2), h1, lin_idx(1),
                                         lin_idx(0), w, lin_idx(1), lin_idx(0));
                SKTT_TRY(sktt_gemm_run(ctx, SKTT_F64, gs2));
            }
            (void)onev;
        }
    }
    phase_fix_kernel<<<k, 256, 0, ctx->stream>>>(N, k, vecs);
    SKTT_LAUNCH_CHECK(ctx);
    if (nconv_host) *nconv_host = nconv;
    return 0;
#undef slots
}

extern "C" int sktt_eig_shift_invert(sktt_ctx* ctx, int dtype, int64_t N, void* Mat, const void* Bmat, double sigma,
                                     int64_t k, int64_t ncv, double tol, int max_restarts, void* lam, void* vecs,
                                     void* work, int* nconv_host) {
    if (!ctx || !Mat || !lam || !vecs || !work) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (N <= 0 || k <= 0 || k > N || ncv < k) return sktt_fail(ctx, SKTT_ERR_ARG, "eig_shift_invert: bad extents");
    if (dtype == SKTT_F64)
        return eig_si_impl<double>(ctx, dtype, N, (double*)Mat, (const double*)Bmat, sigma, (int)k, (int)ncv, tol,
                                   max_restarts, (cplx*)lam, (cplx*)vecs, work, nconv_host);
    return eig_si_impl<cplx>(ctx, dtype, N, (cplx*)Mat, (const cplx*)Bmat, sigma, (int)k, (int)ncv, tol, max_restarts,
                             (cplx*)lam, (cplx*)vecs, work, nconv_host);
}
