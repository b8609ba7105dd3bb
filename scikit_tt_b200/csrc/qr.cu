// qr.cu -- Householder QR / RQ of the solved ALS core (scipy.linalg.qr / rq with mode='economic',
// sle.py:517-525 and :533-541).
//
// One cooperative kernel: the m x n matrix is split by rows over G CTAs, each CTA keeps its rows
// in shared memory for the whole factorisation.  Per reflector ONE grid-wide barrier: every CTA
// publishes its partial sums  g[c] = sum_{i>j} conj(a_ij) a_ic  (c = j gives the tail norm, c > j
// gives v^H A up to the known row-j term) together with row j, so tau, beta and w = v^H A are
// available everywhere after that single barrier.  Q is then formed in place (LAPACK org2r order).
// The loader/storer take signed strides + conjugation so that RQ is QR of the row-reversed
// conjugate transpose without any HBM-side permutation.
#include <cooperative_groups.h>

#include "common.cuh"
#include "blas1.cuh"
namespace cg = cooperative_groups;

#define QR_THREADS 256
#define QR_MAX_CTAS 128

struct QrView {  // element (i, j) of the factored matrix lives at base[off + i*si + j*sj] (optionally conj)
    long long off, si, sj;
    int conj;
};

template <typename T>
struct QrShared {
    T tau;
    T scale_inv;  // 1 / (alpha - beta)
    double beta;
};

// partial buffers: part[parity][cta][n] ; rowbuf[parity][n]
template <typename T>
__global__ void __launch_bounds__(QR_THREADS)
qr_householder_kernel(const T* __restrict__ src, QrView lv, int m, int n, int rows_per_cta, T* __restrict__ qdst,
                      QrView qv, T* __restrict__ rdst, QrView rv, T* part, T* rowbuf, int cooperative) {
    extern __shared__ unsigned char smem_raw[];
    const int ld = n + 1;
    T* Arows = (T*)smem_raw;                 // [rows_per_cta][ld]
    T* wvec = Arows + (size_t)rows_per_cta * ld;  // [n]  (w = v^H A, or staging)
    T* rowj = wvec + n;                      // [n]
    T* taus = rowj + n;                      // [n]
    __shared__ T s_tau, s_scale;
    __shared__ double s_beta;
    const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x;
    const int row_begin = cta * rows_per_cta;
    const int nrows = max(0, min(rows_per_cta, m - row_begin));
    const int k = min(m, n);

    for (int e = tid; e < nrows * n; e += QR_THREADS) {
        int rr = e / n, cc = e % n;
        T v = src[lv.off + (long long)(row_begin + rr) * lv.si + (long long)cc * lv.sj];
        Arows[rr * ld + cc] = lv.conj ? Num<T>::conj(v) : v;
    }
    __syncthreads();

    auto barrier = [&]() {
        if (cooperative) cg::this_grid().sync();
        else __syncthreads();
    };

    // ------------------------------------------------------------------ factorisation
    for (int j = 0; j < k; ++j) {
        const int par = j & 1;
        T* mypart = part + ((size_t)par * G + cta) * n;
        // partial g[c] over local rows i > j, thread per column c >= j
        for (int c = j + tid; c < n; c += QR_THREADS) {
            T s = Num<T>::zero();
            for (int rr = 0; rr < nrows; ++rr) {
                if (row_begin + rr > j) Num<T>::fma(s, Num<T>::conj(Arows[rr * ld + j]), Arows[rr * ld + c]);
            }
            mypart[c] = s;
        }
        if (j >= row_begin && j < row_begin + nrows)
            for (int c = j + tid; c < n; c += QR_THREADS) rowbuf[(size_t)par * n + c] = Arows[(j - row_begin) * ld + c];
        __threadfence();
        barrier();
        // combine (fixed order -> deterministic)
        for (int c = j + tid; c < n; c += QR_THREADS) {
            T s = Num<T>::zero();
            for (int g = 0; g < G; ++g) s = Num<T>::add(s, ld_cg<T>(part + ((size_t)par * G + g) * n + c));
            wvec[c] = s;
            rowj[c] = ld_cg<T>(rowbuf + (size_t)par * n + c);
        }
        __syncthreads();
        if (tid == 0) {
            T alpha = rowj[j];
            double sigma = Num<T>::real(wvec[j]);
            double ar = Num<T>::real(alpha), ai = Num<T>::imag(alpha);
            if (sigma == 0.0 && ai == 0.0) {
                s_tau = Num<T>::zero();
                s_scale = Num<T>::zero();
                s_beta = ar;
            } else {
                double nrm = sqrt(ar * ar + ai * ai + sigma);
                double beta = ar >= 0.0 ? -nrm : nrm;
                s_beta = beta;
                s_tau = Num<T>::from((beta - ar) / beta, -ai / beta);
                s_scale = Num<T>::div(Num<T>::one(), Num<T>::from(ar - beta, ai));
            }
            taus[j] = s_tau;
        }
        __syncthreads();
        const T tau = s_tau, scl = s_scale;
        const T alpha = rowj[j];
        // w[c] = conj(tau) * (v^H A)[c],  v^H A[c] = a_jc + conj(scl) * g[c]     (c > j)
        for (int c = j + 1 + tid; c < n; c += QR_THREADS) {
            T vha = Num<T>::add(rowj[c], Num<T>::mul(Num<T>::conj(scl), wvec[c]));
            wvec[c] = Num<T>::mul(Num<T>::conj(tau), vha);
        }
        __syncthreads();
        // update local rows: row j: a_jc -= w_c ; rows i > j: v_i = a_ij * scl, a_ic -= v_i w_c, a_ij = v_i
        for (int e = tid; e < nrows * (n - j); e += QR_THREADS) {
            int rr = e / (n - j), c = j + e % (n - j);
            int gr = row_begin + rr;
            if (gr < j) continue;
            if (gr == j) {
                if (c == j) Arows[rr * ld + j] = Num<T>::from(s_beta, 0.0);
                else Arows[rr * ld + c] = Num<T>::sub(Arows[rr * ld + c], wvec[c]);
            } else if (c > j) {
                T vi = Num<T>::mul(Arows[rr * ld + j], scl);
                Arows[rr * ld + c] = Num<T>::sub(Arows[rr * ld + c], Num<T>::mul(vi, wvec[c]));
            }
        }
        __syncthreads();
        for (int rr = tid; rr < nrows; rr += QR_THREADS)
            if (row_begin + rr > j) Arows[rr * ld + j] = Num<T>::mul(Arows[rr * ld + j], scl);
        __syncthreads();
        (void)alpha;
    }

    // ------------------------------------------------------------------ R (k x n upper trapezoid)
    if (rdst) {
        for (int e = tid; e < nrows * n; e += QR_THREADS) {
            int rr = e / n, c = e % n, gr = row_begin + rr;
            if (gr >= k) continue;
            T v = c >= gr ? Arows[rr * ld + c] : Num<T>::zero();
            if (rv.conj) v = Num<T>::conj(v);
            rdst[rv.off + (long long)gr * rv.si + (long long)c * rv.sj] = v;
        }
    }

    // ------------------------------------------------------------------ form Q in place (m x k)
    // LAPACK org2r order: reflectors applied last-to-first; column c > j of the growing Q has rows
    // < c already zeroed, so row j contributes nothing to v^H Q[:, c] and no row exchange is needed.
    for (int j = k - 1; j >= 0; --j) {
        const int par = (j + 1) & 1;  // opposite phase to the factorisation's last use of the buffers
        const T tau = taus[j];
        T* mypart = part + ((size_t)par * G + cta) * n;
        for (int c = j + 1 + tid; c < k; c += QR_THREADS) {
            T s = Num<T>::zero();
            for (int rr = 0; rr < nrows; ++rr)
                if (row_begin + rr > j) Num<T>::fma(s, Num<T>::conj(Arows[rr * ld + j]), Arows[rr * ld + c]);
            mypart[c] = s;
        }
        __threadfence();
        barrier();
        for (int c = j + 1 + tid; c < k; c += QR_THREADS) {
            T s = Num<T>::zero();
            for (int g = 0; g < G; ++g) s = Num<T>::add(s, ld_cg<T>(part + ((size_t)par * G + g) * n + c));
            wvec[c] = Num<T>::mul(tau, s);  // H = I - tau v v^H applied from the left
        }
        __syncthreads();
        for (int e = tid; e < nrows * (k - j); e += QR_THREADS) {
            int rr = e / (k - j), c = j + e % (k - j);
            int gr = row_begin + rr;
            if (c == j) continue;  // column j itself is finalised below (v must stay intact here)
            if (gr < j) continue;
            T vi = gr == j ? Num<T>::one() : Arows[rr * ld + j];
            Arows[rr * ld + c] = Num<T>::sub(Arows[rr * ld + c], Num<T>::mul(vi, wvec[c]));
        }
        __syncthreads();
        for (int rr = tid; rr < nrows; rr += QR_THREADS) {
            int gr = row_begin + rr;
            if (gr < j) Arows[rr * ld + j] = Num<T>::zero();
            else if (gr == j) Arows[rr * ld + j] = Num<T>::sub(Num<T>::one(), tau);
            else Arows[rr * ld + j] = Num<T>::neg(Num<T>::mul(tau, Arows[rr * ld + j]));
        }
        __syncthreads();
    }
    for (int e = tid; e < nrows * k; e += QR_THREADS) {
        int rr = e / k, c = e % k;
        T v = Arows[rr * ld + c];
        if (qv.conj) v = Num<T>::conj(v);
        qdst[qv.off + (long long)(row_begin + rr) * qv.si + (long long)c * qv.sj] = v;
    }
}

template <typename T>
static int qr_launch(sktt_ctx* ctx, int m, int n, const T* src, QrView lv, T* qdst, QrView qv, T* rdst, QrView rv) {
    const size_t budget = 200 * 1024;
    const size_t per_row = (size_t)(n + 1) * sizeof(T);
    const size_t fixed = (size_t)3 * n * sizeof(T) + 256;
    if (fixed + per_row > budget) return sktt_fail(ctx, SKTT_ERR_ARG, "qr: too many columns for the shared-memory kernel");
    int rows_cap = (int)((budget - fixed) / per_row);
    int G = (m + 31) / 32;
    if (G > QR_MAX_CTAS) G = QR_MAX_CTAS;
    if (G > ctx->sm_count) G = ctx->sm_count;
    if (G < 1) G = 1;
    int rows_per_cta = (m + G - 1) / G;
    if (rows_per_cta > rows_cap) {
        rows_per_cta = rows_cap;
        G = (m + rows_per_cta - 1) / rows_per_cta;
        if (G > ctx->sm_count) return sktt_fail(ctx, SKTT_ERR_ARG, "qr: matrix too large for the cooperative kernel");
    }
    G = (m + rows_per_cta - 1) / rows_per_cta;
    size_t smem = (size_t)rows_per_cta * per_row + fixed;
    size_t part_bytes = ((size_t)2 * G * n + 2 * n) * sizeof(T);
    SKTT_TRY(sktt_scratch_reserve(ctx, SKTT_SCRATCH_BULK_OFF + part_bytes));
    T* part = (T*)((char*)ctx->scratch + SKTT_SCRATCH_BULK_OFF);
    T* rowbuf = part + (size_t)2 * G * n;
    static size_t configured = 0;
    if (smem > configured) {
        SKTT_CUDA(ctx, cudaFuncSetAttribute(qr_householder_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)budget));
        configured = budget;
    }
    int coop = G > 1 ? 1 : 0;
    void* args[] = {&src, &lv, &m, &n, &rows_per_cta, &qdst, &qv, &rdst, &rv, &part, &rowbuf, &coop};
    if (coop)
        SKTT_CUDA(ctx, cudaLaunchCooperativeKernel((void*)qr_householder_kernel<T>, dim3(G), dim3(QR_THREADS), args, smem,
                                                   ctx->stream));
    else
        SKTT_CUDA(ctx, cudaLaunchKernel((void*)qr_householder_kernel<T>, dim3(1), dim3(QR_THREADS), args, smem,
                                        ctx->stream));
    ctx->launches++;
    return 0;
}

extern "C" int64_t sktt_qr_work(int64_t m, int64_t n) {
    (void)m;
    (void)n;
    return 1;  // the kernel works out of shared memory + context scratch; kept for ABI stability
}

int sktt_qr_internal(sktt_ctx* ctx, int dtype, int m, int n, const void* A, void* Q, void* R) {
    int k = m < n ? m : n;
    QrView lv{0, n, 1, 0}, qv{0, k, 1, 0}, rv{0, n, 1, 0};
    if (dtype == SKTT_F64) return qr_launch<double>(ctx, m, n, (const double*)A, lv, (double*)Q, qv, (double*)R, rv);
    return qr_launch<cplx>(ctx, m, n, (const cplx*)A, lv, (cplx*)Q, qv, (cplx*)R, rv);
}

extern "C" int sktt_qr_left(sktt_ctx* ctx, int dtype, int64_t m, int64_t n, const void* A, void* Q_out, void* R_out,
                            void* work) {
    (void)work;
    if (!ctx || !A || !Q_out) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (m <= 0 || n <= 0 || m > 0x7fffffff || n > 0x7fffffff) return sktt_fail(ctx, SKTT_ERR_ARG, "qr: bad extents");
    return sktt_qr_internal(ctx, dtype, (int)m, (int)n, A, Q_out, R_out);
}

// A (m x n) = R Q with the sign convention of LAPACK gerqf (scipy.linalg.rq): the reflectors are generated from the
// LAST row backwards and annihilate the LEADING part of each row.  With C = (J_m A J_n)^H (n x m, both index orders
// reversed) = Qc Rc this is exactly Householder QR of C:  Q = J_k Qc^H J_n,  R = J_m Rc^H J_k.
extern "C" int sktt_rq_right(sktt_ctx* ctx, int dtype, int64_t m64, int64_t n64, const void* A, void* Q_out,
                             void* R_out, void* work) {
    (void)work;
    if (!ctx || !A || !Q_out) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (m64 <= 0 || n64 <= 0 || m64 > 0x7fffffff || n64 > 0x7fffffff)
        return sktt_fail(ctx, SKTT_ERR_ARG, "rq: bad extents");
    const int m = (int)m64, n = (int)n64, k = m < n ? m : n;
    // C[i][j] = conj(A[m-1-j][n-1-i])                   (C is n x m)
    QrView lv{(long long)(m - 1) * n + (n - 1), -1, -(long long)n, 1};
    // Q[k-1-j][n-1-i] = conj(Qc[i][j])                   (Q is k x n, row-major)
    QrView qv{(long long)(k - 1) * n + (n - 1), -1, -(long long)n, 1};
    // R[m-1-c][k-1-g] = conj(Rc[g][c])                   (Rc is k x m; R is m x k, row-major)
    QrView rv{(long long)(m - 1) * k + (k - 1), -1, -(long long)k, 1};
    if (dtype == SKTT_F64)
        return qr_launch<double>(ctx, n, m, (const double*)A, lv, (double*)Q_out, qv, (double*)R_out, rv);
    return qr_launch<cplx>(ctx, n, m, (const cplx*)A, lv, (cplx*)Q_out, qv, (cplx*)R_out, rv);
}
