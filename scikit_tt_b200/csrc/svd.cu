// svd.cu -- truncated SVD for MALS core splitting (sle.py:603-614, :626-639), the eigenvector
// re-orthonormalisation of evp.als (evp.py:452-487) and TT.ortho_left/ortho_right
// (tensor_train.py:1162-1189, :1268-1299), plus the small Hermitian eigen-solver built on the
// same rotations (evp.py:434-439 and the projected problems of the Krylov eigen-solvers).
//
// Algorithm: Householder QR (qr.cu) of the tall orientation, then one-sided Jacobi (Hestenes) on
// the stacked columns [R; I]: rotations act on the column pair (p, q) of both blocks, the upper
// block converges to U * diag(s), the lower block accumulates V.  Pairs follow a round-robin
// tournament, one warp per pair; a single CTA with the columns in shared memory for small
// problems, a cooperative grid with one barrier per tournament round otherwise.  One-sided Jacobi
// gives singular values with high relative accuracy, which the strict rule s_i / s_0 > threshold
// of the reference needs in order to pick the same rank as LAPACK gesvd.
#include <cooperative_groups.h>

#include "common.cuh"
#include "blas1.cuh"
namespace cg = cooperative_groups;

int sktt_qr_internal(sktt_ctx* ctx, int dtype, int m, int n, const void* A, void* Q, void* R);

struct MatView {  // element (i, j) at base[off + i*si + j*sj], optionally conjugated
    long long off, si, sj;
    int conj;
};

// X[j][i] (column-major, column length L = mw + ncol): upper block = B(i, j), lower block = identity.
// shift (device scalar, may be null) is added to the diagonal of the upper block (eigh path).
template <typename T>
__global__ void jacobi_pack_kernel(const T* __restrict__ src, MatView v, int mw, int ncol, T* __restrict__ X,
                                   const double* shift) {
    const int L = mw + ncol;
    const double sh = shift ? *shift : 0.0;
    long long total = (long long)ncol * L;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        int j = (int)(e / L), i = (int)(e % L);
        T val;
        if (i < mw) {
            val = src[v.off + (long long)i * v.si + (long long)j * v.sj];
            if (v.conj) val = Num<T>::conj(val);
            if (i == j && sh != 0.0) val = Num<T>::add(val, Num<T>::from(sh, 0.0));
        } else {
            val = (i - mw == j) ? Num<T>::one() : Num<T>::zero();
        }
        X[e] = val;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
jacobi_sweep_kernel(T* __restrict__ Xg, int L, int mw, int n, int max_sweeps, double tol, int* counters,
                    int* sweeps_out, int use_smem, int cooperative) {
    extern __shared__ unsigned char smem_raw[];
    T* X = Xg;
    if (use_smem) {
        X = (T*)smem_raw;
        for (long long e = threadIdx.x; e < (long long)n * L; e += blockDim.x) X[e] = Xg[e];
        __syncthreads();
    }
    __shared__ int s_rot;
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    const int np = n + (n & 1);   // players (phantom index n if n is odd)
    const int half = np / 2;
    auto barrier = [&]() {
        if (cooperative) cg::this_grid().sync();
        else __syncthreads();
    };
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        if (threadIdx.x == 0) s_rot = 0;
        __syncthreads();
        int my_rot = 0;
        for (int step = 0; step < np - 1; ++step) {
            for (int kk = gwarp; kk < half; kk += nwarps) {
                int p, q;
                if (kk == 0) {
                    p = np - 1;
                    q = step;
                } else {
                    p = (step + kk) % (np - 1);
                    q = (step - kk + (np - 1)) % (np - 1);
                }
                if (p > q) { int t = p; p = q; q = t; }
                if (q >= n) continue;  // phantom
                T* xp = X + (size_t)p * L;
                T* xq = X + (size_t)q * L;
                double a = 0.0, b = 0.0;
                T g = Num<T>::zero();
                for (int i = lane; i < mw; i += 32) {
                    T vp = xp[i], vq = xq[i];
                    a += Num<T>::abs2(vp);
                    b += Num<T>::abs2(vq);
                    Num<T>::fma(g, Num<T>::conj(vp), vq);
                }
                a = warp_sum<double>(a);
                b = warp_sum<double>(b);
                g = warp_sum<T>(g);
                double gabs = sqrt(Num<T>::abs2(g));
                if (gabs == 0.0 || gabs <= tol * sqrt(a) * sqrt(b)) continue;
                my_rot++;
                double zeta = (b - a) / (2.0 * gabs);
                double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                T w = Num<T>::scale(g, 1.0 / gabs);  // phase of <xp, xq>
                T sw = Num<T>::scale(w, s), swc = Num<T>::conj(sw);
                for (int i = lane; i < L; i += 32) {
                    T vp = xp[i], vq = xq[i];
                    // xp' = c xp - s conj(w) xq ; xq' = s w xp + c xq
                    xp[i] = Num<T>::sub(Num<T>::scale(vp, c), Num<T>::mul(swc, vq));
                    xq[i] = Num<T>::add(Num<T>::mul(sw, vp), Num<T>::scale(vq, c));
                }
            }
            if (cooperative) { __threadfence(); }
            barrier();
        }
        if (lane == 0 && my_rot) atomicAdd(&s_rot, my_rot);
        __syncthreads();
        int total_rot;
        if (cooperative) {
            if (threadIdx.x == 0 && s_rot) atomicAdd(&counters[sweep], s_rot);
            __threadfence();
            cg::this_grid().sync();
            total_rot = __ldcg(&counters[sweep]);
        } else {
            total_rot = s_rot;
        }
        __syncthreads();
        if (total_rot == 0) { ++sweep; break; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && sweeps_out) *sweeps_out = sweep;
    if (use_smem) {
        __syncthreads();
        for (long long e = threadIdx.x; e < (long long)n * L; e += blockDim.x) Xg[e] = X[e];
    }
}

// Post-processing (one CTA): norms -> sort (descending) -> scaled upper block P and lower block V
// written through views; null columns of P are completed to an orthonormal set; rank rule.
//   mode_eig != 0: ascending order, values = norm - shift, only V written (Hermitian eigen path).
template <typename T>
__global__ void __launch_bounds__(256)
jacobi_post_kernel(const T* __restrict__ X, int L, int mw, int n, T* __restrict__ P, MatView pv, T* __restrict__ V,
                   MatView vv, double* __restrict__ S, double threshold, int max_rank, int* rank_out,
                   const double* shift, int mode_eig, int* ord_scratch) {
    extern __shared__ unsigned char smem_raw[];
    T* u = (T*)smem_raw;               // [mw] scratch vector (null-space completion)
    double* sig = (double*)(u + mw);   // [n]
    int* ord = (int*)(sig + n);        // [n]  ord[rank] = source column
    __shared__ double s_red[32];
    __shared__ double s_nrm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int j = warp; j < n; j += 8) {
        double a = 0.0;
        for (int i = lane; i < mw; i += 32) a += Num<T>::abs2(X[(size_t)j * L + i]);
        a = warp_sum<double>(a);
        if (lane == 0) sig[j] = sqrt(a);
    }
    __syncthreads();
    for (int j = tid; j < n; j += 256) {
        int rank = 0;
        double sj = sig[j];
        for (int i = 0; i < n; ++i) {
            double si = sig[i];
            bool before = mode_eig ? (si < sj || (si == sj && i < j)) : (si > sj || (si == sj && i < j));
            rank += before ? 1 : 0;
        }
        ord[rank] = j;
    }
    __syncthreads();
    const double sh = shift ? *shift : 0.0;
    const double smax = mode_eig ? sig[ord[n - 1]] : sig[ord[0]];
    for (int r = tid; r < n; r += 256) S[r] = sig[ord[r]] - sh;
    if (tid == 0 && rank_out) {
        int k = n;
        if (threshold != 0.0) {
            k = 0;
            for (int r = 0; r < n; ++r)
                if (sig[ord[r]] / smax > threshold) k++;  // strict, as sle.py:609 / tensor_train.py:1175
        }
        if (max_rank > 0 && k > max_rank) k = max_rank;
        *rank_out = k;
    }
    // V block
    if (V) {
        for (long long e = tid; e < (long long)n * n; e += 256) {
            int r = (int)(e / n), i = (int)(e % n);
            T val = X[(size_t)ord[r] * L + mw + i];
            if (vv.conj) val = Num<T>::conj(val);
            V[vv.off + (long long)i * vv.si + (long long)r * vv.sj] = val;
        }
    }
    if (!P) return;
    // P block, normalised; columns with negligible norm are completed afterwards
    const double cutoff = smax * 2.220446049250313e-16 * (mw > n ? mw : n);
    for (long long e = tid; e < (long long)n * mw; e += 256) {
        int r = (int)(e / mw), i = (int)(e % mw);
        double s = sig[ord[r]];
        T val = s > cutoff ? Num<T>::scale(X[(size_t)ord[r] * L + i], 1.0 / s) : Num<T>::zero();
        if (pv.conj) val = Num<T>::conj(val);
        P[pv.off + (long long)i * pv.si + (long long)r * pv.sj] = val;
    }
    __syncthreads();
    // completion of null columns (only reachable for rank-deficient input): Gram-Schmidt of unit
    // vectors against the columns accepted so far, twice, accept when the remainder is not tiny.
    int first_null = n;
    for (int r = 0; r < n; ++r)
        if (!(sig[ord[r]] > cutoff)) { first_null = r; break; }
    if (first_null == n || first_null >= mw) return;
    int trial = 0;
    for (int r = first_null; r < n && r < mw; ++r) {
        bool accepted = false;
        while (!accepted && trial < mw) {
            for (int i = tid; i < mw; i += 256) u[i] = (i == trial) ? Num<T>::one() : Num<T>::zero();
            __syncthreads();
            for (int pass = 0; pass < 2; ++pass) {
                for (int c = 0; c < r; ++c) {
                    // proj = <P[:,c], u>
                    T acc = Num<T>::zero();
                    for (int i = tid; i < mw; i += 256) {
                        T pc = P[pv.off + (long long)i * pv.si + (long long)c * pv.sj];
                        if (pv.conj) pc = Num<T>::conj(pc);
                        Num<T>::fma(acc, Num<T>::conj(pc), u[i]);
                    }
                    acc = block_sum<T>(acc, (T*)s_red);
                    __syncthreads();
                    for (int i = tid; i < mw; i += 256) {
                        T pc = P[pv.off + (long long)i * pv.si + (long long)c * pv.sj];
                        if (pv.conj) pc = Num<T>::conj(pc);
                        u[i] = Num<T>::sub(u[i], Num<T>::mul(pc, acc));
                    }
                    __syncthreads();
                }
            }
            double a = 0.0;
            for (int i = tid; i < mw; i += 256) a += Num<T>::abs2(u[i]);
            a = block_sum<double>(a, s_red);
            if (tid == 0) s_nrm = sqrt(a);
            __syncthreads();
            double nrm = s_nrm;
            if (nrm > 0.5) {
                for (int i = tid; i < mw; i += 256) {
                    T val = Num<T>::scale(u[i], 1.0 / nrm);
                    if (pv.conj) val = Num<T>::conj(val);
                    P[pv.off + (long long)i * pv.si + (long long)r * pv.sj] = val;
                }
                accepted = true;
            }
            trial++;
            __syncthreads();
        }
    }
    (void)ord_scratch;
}

// max_i sum_j |A_ij|  (Gershgorin radius bound) -> out[0]
template <typename T>
__global__ void __launch_bounds__(256) gershgorin_kernel(const T* __restrict__ A, int N, double* out) {
    __shared__ double sh[32];
    double best = 0.0;
    for (int i = 0; i < N; ++i) {
        double a = 0.0;
        for (int j = threadIdx.x; j < N; j += 256) a += sqrt(Num<T>::abs2(A[(size_t)i * N + j]));
        a = block_sum<double>(a, sh);
        best = a > best ? a : best;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = best;
}

// ------------------------------------------------------------------------------------------------
template <typename T>
static int jacobi_run(sktt_ctx* ctx, T* X, int L, int mw, int n, int* sweeps_dev) {
    const size_t xbytes = (size_t)n * L * sizeof(T);
    const int max_sweeps = 60;
    const double tol = 2.220446049250313e-16 * sqrt((double)(mw > 4 ? mw : 4));
    int* counters = (int*)((char*)ctx->scratch + SKTT_SCRATCH_COUNTER_OFF + 256);  // 64 ints
    SKTT_CUDA(ctx, cudaMemsetAsync(counters, 0, 64 * sizeof(int), ctx->stream));
    SKTT_ONCE_PER_DEVICE(ctx);
    if (!configured) {
        SKTT_CUDA(ctx, cudaFuncSetAttribute(jacobi_sweep_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            200 * 1024));
        configured = true;
    }
    int ms = max_sweeps;
    double tl = tol;
    if (xbytes <= 180 * 1024) {
        int use_smem = 1, coop = 0;
        void* args[] = {&X, &L, &mw, &n, &ms, &tl, &counters, &sweeps_dev, &use_smem, &coop};
        SKTT_CUDA(ctx, cudaLaunchKernel((void*)jacobi_sweep_kernel<T>, dim3(1), dim3(256), args, xbytes, ctx->stream));
    } else {
        int half = (n + 1) / 2;
        int G = (half + 7) / 8;
        if (G > 2 * ctx->sm_count) G = 2 * ctx->sm_count;
        int use_smem = 0, coop = G > 1 ? 1 : 0;
        void* args[] = {&X, &L, &mw, &n, &ms, &tl, &counters, &sweeps_dev, &use_smem, &coop};
        if (coop)
            SKTT_CUDA(ctx, cudaLaunchCooperativeKernel((void*)jacobi_sweep_kernel<T>, dim3(G), dim3(256), args, 0,
                                                       ctx->stream));
        else
            SKTT_CUDA(ctx, cudaLaunchKernel((void*)jacobi_sweep_kernel<T>, dim3(1), dim3(256), args, 0, ctx->stream));
    }
    ctx->launches++;
    return 0;
}

extern "C" int64_t sktt_svd_work(int64_t m, int64_t n) {
    int64_t k = m < n ? m : n, big = m < n ? n : m;
    return 4 * big * k + 3 * k * k + 256;
}

template <typename T>
static int svd_impl(sktt_ctx* ctx, int dtype, int m, int n, const T* A, T* U, double* S, T* Vh, double threshold,
                    int max_rank, T* work, int* new_rank_host, int* sweeps_host) {
    // requires m >= n (svd_dispatch transposes wide inputs)
    const int mb = m, nb = n;
    T* Qb = work;                          // [mb][nb]
    T* X = Qb + (size_t)mb * nb;           // [nb][2 nb]
    T* Ur = X + (size_t)2 * nb * nb;       // [nb][nb]
    int* rank_dev = (int*)((char*)ctx->scratch + 1024 + 64);
    int* sweeps_dev = rank_dev + 1;
    const bool use_qr = mb > nb;
    const int mw = nb, L = 2 * nb;
    {
        const T* packsrc = A;
        if (use_qr) {
            SKTT_TRY(sktt_qr_internal(ctx, dtype, mb, nb, A, Qb, Ur));  // R parked in Ur until packed
            packsrc = Ur;
        }
        MatView rv{0, nb, 1, 0};
        int blocks = (int)(((long long)nb * L + 255) / 256);
        if (blocks > 1024) blocks = 1024;
        jacobi_pack_kernel<T><<<blocks, 256, 0, ctx->stream>>>(packsrc, rv, mw, nb, X, nullptr);
        SKTT_LAUNCH_CHECK(ctx);
    }
    SKTT_TRY(jacobi_run<T>(ctx, X, L, mw, nb, sweeps_dev));
    if (nb > 4096) return sktt_fail(ctx, SKTT_ERR_ARG, "svd: min(m, n) > 4096 not supported");
    size_t post_smem = (size_t)nb * (sizeof(double) + sizeof(int)) + 16 + (size_t)mw * sizeof(T);
    SKTT_ONCE_PER_DEVICE(ctx);
    if (!configured) {
        SKTT_CUDA(ctx, cudaFuncSetAttribute(jacobi_post_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    // P (scaled upper block) -> Ur (row-major nb x nb) when QR was used, else straight into U
    MatView pv, vv;
    T* Pdst;
    if (use_qr) {
        Pdst = Ur;
        pv = MatView{0, nb, 1, 0};
    } else {
        Pdst = U;
        pv = MatView{0, nb, 1, 0};  // square: U is nb x nb row-major
    }
    vv = MatView{0, 1, n, 1};  // Vh[r][i] = conj(V[i][r]): offset i*1 + r*n
    jacobi_post_kernel<T><<<1, 256, post_smem, ctx->stream>>>(X, L, mw, nb, Pdst, pv, Vh, vv, S, threshold, max_rank,
                                                              rank_dev, nullptr, 0, nullptr);
    SKTT_LAUNCH_CHECK(ctx);
    if (use_qr) {
        // U (m x nb) = Qb (m x nb) * Ur (nb x nb)
        GemmDesc g = gemm_desc(mb, nb, nb, Qb, lin_idx(nb), lin_idx(1), Ur, lin_idx(nb), lin_idx(1), U, lin_idx(nb),
                               lin_idx(1));
        SKTT_TRY(sktt_gemm_run(ctx, dtype, g));
    }
    SKTT_CUDA(ctx, cudaMemcpyAsync(ctx->mailbox, rank_dev, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SKTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (new_rank_host) *new_rank_host = ((int*)ctx->mailbox)[0];
    if (sweeps_host) *sweeps_host = ((int*)ctx->mailbox)[1];
    return 0;
}

template <typename T>
__global__ void conj_transpose_kernel(int m, int n, const T* __restrict__ A, T* __restrict__ out) {
    // out (n x m) = A^H, A is m x n; tiled through shared memory
    __shared__ T tile[32][33];
    int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        int i = by + dy, j = bx + threadIdx.x;
        if (i < m && j < n) tile[dy][threadIdx.x] = Num<T>::conj(A[(size_t)i * n + j]);
    }
    __syncthreads();
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        int j = bx + dy, i = by + threadIdx.x;
        if (i < m && j < n) out[(size_t)j * m + i] = tile[threadIdx.x][dy];
    }
}

template <typename T>
static int svd_dispatch(sktt_ctx* ctx, int dtype, int m, int n, const T* A, T* U, double* S, T* Vh, double threshold,
                        int max_rank, T* work, int* new_rank_host, int* sweeps_host) {
    if (m >= n) return svd_impl<T>(ctx, dtype, m, n, A, U, S, Vh, threshold, max_rank, work, new_rank_host, sweeps_host);
    // wide: A^H = U' S V'^H  =>  A = V' S U'^H.  Use the tail of the workspace for A^H, U', V'^H.
    const int k = m;
    T* At = work + (size_t)n * k + 3 * (size_t)k * k + 64;  // [n][m]
    T* Ut = At + (size_t)n * m;                            // [n][k]
    T* Vht = Ut + (size_t)n * k;                           // [k][m]
    dim3 grid((n + 31) / 32, (m + 31) / 32), block(32, 8);
    conj_transpose_kernel<T><<<grid, block, 0, ctx->stream>>>(m, n, A, At);
    SKTT_LAUNCH_CHECK(ctx);
    SKTT_TRY(svd_impl<T>(ctx, dtype, n, m, At, Ut, S, Vht, threshold, max_rank, work, new_rank_host, sweeps_host));
    // U (m x k) = Vht^H ; Vh (k x n) = Ut^H
    dim3 g1((m + 31) / 32, (k + 31) / 32);
    conj_transpose_kernel<T><<<g1, block, 0, ctx->stream>>>(k, m, Vht, U);
    SKTT_LAUNCH_CHECK(ctx);
    dim3 g2((k + 31) / 32, (n + 31) / 32);
    conj_transpose_kernel<T><<<g2, block, 0, ctx->stream>>>(n, k, Ut, Vh);
    SKTT_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" int sktt_svd_truncate(sktt_ctx* ctx, int dtype, int64_t m, int64_t n, const void* A, void* U, double* S,
                                 void* Vh, double threshold, int64_t max_rank, void* work, int* new_rank_host,
                                 int* sweeps_host) {
    if (!ctx || !A || !U || !S || !Vh || !work) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (m <= 0 || n <= 0 || m > 0x7fffffff || n > 0x7fffffff) return sktt_fail(ctx, SKTT_ERR_ARG, "svd: bad extents");
    if (threshold < 0) return sktt_fail(ctx, SKTT_ERR_ARG, "svd: negative threshold");
    int mr = max_rank > 0x7fffffff ? 0 : (int)max_rank;
    if (dtype == SKTT_F64)
        return svd_dispatch<double>(ctx, dtype, (int)m, (int)n, (const double*)A, (double*)U, S, (double*)Vh, threshold,
                                    mr, (double*)work, new_rank_host, sweeps_host);
    return svd_dispatch<cplx>(ctx, dtype, (int)m, (int)n, (const cplx*)A, (cplx*)U, S, (cplx*)Vh, threshold, mr,
                              (cplx*)work, new_rank_host, sweeps_host);
}

// ------------------------------------------------------------------------------------------------
// Hermitian eigen-decomposition through the same rotations: S = A + c I (c = Gershgorin bound, so
// S is positive semidefinite), one-sided Jacobi on [S; I] -> V holds the eigenvectors and the
// column norms are lambda + c.  Ascending order as scipy.linalg.eigh.
// ------------------------------------------------------------------------------------------------
template <typename T>
static int eigh_impl(sktt_ctx* ctx, int N, T* Mat, double* W, T* V, int* sweeps_host) {
    size_t xelems = (size_t)2 * N * N;
    size_t off = SKTT_SCRATCH_BULK_OFF;
    SKTT_TRY(sktt_scratch_reserve(ctx, off + xelems * sizeof(T) + 256));
    T* X = (T*)((char*)ctx->scratch + off);
    double* shift = (double*)((char*)ctx->scratch + 512);
    int* sweeps_dev = (int*)((char*)ctx->scratch + 1024 + 64) + 1;
    gershgorin_kernel<T><<<1, 256, 0, ctx->stream>>>(Mat, N, shift);
    SKTT_LAUNCH_CHECK(ctx);
    MatView av{0, N, 1, 0};
    int blocks = (int)((xelems + 255) / 256 > 1024 ? 1024 : (xelems + 255) / 256);
    jacobi_pack_kernel<T><<<blocks, 256, 0, ctx->stream>>>(Mat, av, N, N, X, shift);
    SKTT_LAUNCH_CHECK(ctx);
    SKTT_TRY(jacobi_run<T>(ctx, X, 2 * N, N, N, sweeps_dev));
    size_t post_smem = (size_t)N * (sizeof(double) + sizeof(int)) + 16 + (size_t)N * sizeof(T);
    SKTT_ONCE_PER_DEVICE(ctx);
    if (!configured) {
        SKTT_CUDA(ctx, cudaFuncSetAttribute(jacobi_post_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    MatView pv{0, 0, 0, 0}, vv{0, N, 1, 0};  // V[i][r] = eigenvector r, row-major N x N
    jacobi_post_kernel<T><<<1, 256, post_smem, ctx->stream>>>(X, 2 * N, N, N, (T*)nullptr, pv, V, vv, W, 0.0, 0,
                                                              (int*)nullptr, shift, 1, nullptr);
    SKTT_LAUNCH_CHECK(ctx);
    if (sweeps_host) {
        SKTT_CUDA(ctx, cudaMemcpyAsync(ctx->mailbox, sweeps_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        SKTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *sweeps_host = *(int*)ctx->mailbox;
    }
    return 0;
}

extern "C" int sktt_eigh_jacobi(sktt_ctx* ctx, int dtype, int64_t N, void* Mat, double* W, void* V, int* sweeps_host) {
    if (!ctx || !Mat || !W || !V) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (N <= 0 || N > 4096) return sktt_fail(ctx, SKTT_ERR_ARG, "eigh_jacobi: N out of range (1..4096)");
    if (dtype == SKTT_F64) return eigh_impl<double>(ctx, (int)N, (double*)Mat, W, (double*)V, sweeps_host);
    return eigh_impl<cplx>(ctx, (int)N, (cplx*)Mat, W, (cplx*)V, sweeps_host);
}
