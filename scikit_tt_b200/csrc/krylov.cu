// krylov.cu -- matrix-free local solves of the ALS/MALS micro systems.  The reference forms the
// dense micro matrix (sle.py:339-345, :381-388) and calls LAPACK; at the north-star sizes (N = r n r'
// = 262144 and beyond) that matrix cannot exist, so the same linear system is solved with CG
// (Hermitian positive definite operators) or restarted GMRES with CGS2 orthogonalisation, the
// operator applied through the fused contraction chain of stacks.cu.  All scalars of the recurrences
// live on the device; the host only peeks at the residual norm through a pinned mailbox one batch
// behind the GPU, so the launch queue never drains.
#include "common.cuh"
#include "blas1.cuh"

// fused.cu: TMA-staged matvec on vectors in the tiled layout [n][a][r2 + 4]
long long sktt_fused_tiled_len(long long r, long long n);
int sktt_fused_to_tiled_ex(sktt_ctx* ctx, long long r, long long n, const double* src, double* dst, int swap,
                           long long r_act, long long c_act);
int sktt_fused_from_tiled_ex(sktt_ctx* ctx, long long r, long long n, const double* src, double* dst, long long r_act,
                             long long c_act);
int sktt_fused_matvec_tiled(sktt_ctx* ctx, long long r, long long R, long long m, long long n, const double* image,
                            const double* vt, double* yt, double* T1p);
int sktt_fused_matvec_tiled_dots(sktt_ctx* ctx, long long r, long long R, long long m, long long n, const double* image,
                                 const double* vt, double* yt, double* T1p, const double* dvec, const double* dvec2,
                                 double* dot_part, unsigned* counter, double* dots_out, const int* skip);

int sktt_fused_pcg_persistent(sktt_ctx* ctx, long long r, long long R, long long m, long long n, const double* image,
                              const double* f, double* u, double* rv, double* p, double* s, double* w, double* T1p,
                              double tol, int max_iters, int max_cycles, int mode, int reps, double* part,
                              double* out_dev, const double* f_nat = nullptr, double* u_nat = nullptr, long long r_act = 0,
                              long long c_act = 0);

// The operator as the Krylov loops see it: either the generic contraction chain on natural-layout vectors, or the
// prepared fused matvec on tiled-layout vectors (all Krylov vectors then live in that layout; padding stays zero).
struct KOp {
    sktt_local_op op;
    bool tiled;
    long long N;   // vector length the solver iterates on
};

static int kop_matvec(sktt_ctx* ctx, int dtype, const KOp& k, const void* v, void* y, void* work) {
    if (k.tiled)
        return sktt_fused_matvec_tiled(ctx, fused_rpad(k.op.r), k.op.R, k.op.m, k.op.n, (const double*)k.op.image, (const double*)v,
                                       (double*)y, (double*)work);
    return sktt_local_matvec(ctx, dtype, &k.op, v, y, work);
}

static int64_t local_dim(const sktt_local_op* op) {
    return op->sites == 1 ? op->r * op->n * op->r3 : op->r * op->n * op->n2 * op->r3;
}
// bound of the vector length in either layout
static int64_t local_dim_bound(const sktt_local_op* op) {
    if (op->sites != 1) return local_dim(op);
    const int64_t tiled = fused_rpad(op->r) * op->n * 68, natural = local_dim(op);   // tiled layout: r' <= 64 padded to 68
    return tiled > natural ? tiled : natural;
}
static int64_t local_mv_work(const sktt_local_op* op) {
    if (op->sites == 1) {
        const int64_t generic = sktt_stack_op_work(op->r, op->R, op->m, op->n, op->r3, op->R2);
        const int64_t fused = op->R * fused_rpad(op->r) * op->n * 68 + 2 * op->n * fused_rpad(op->r) * 68;   // T1 (padded) + tiled pair
        return generic > fused ? generic : fused;
    }
    return sktt_micro_matvec_mals_work(op->r, op->R, op->m, op->n, op->R2, op->m2, op->n2, op->R3, op->r3);
}

// upper bound of the prepared-operator image (fused.cu) for one-site operators, 0 otherwise
static int64_t local_image_bound(const sktt_local_op* op) {
    if (op->sites != 1 || op->image) return 0;
    const int64_t rp = fused_rpad(op->r);
    return op->R * ((op->m + 31) / 32) * ((op->n + 15) / 16) * 1920 + 12 * 16 * 68 + ((op->R * rp + 95) / 96) * rp * 100 + 64 + 8;
}

// elements used by the solver proper (vectors of length Nb + small state), excluding matvec scratch
static int64_t solver_core_work(int method, int restart, int64_t Nb) {
    if (method == 0) return 5 * Nb + 128 * 128 + 64;                      // r, p, Ap (+ s, z, mode preconditioner of the fused variant)
    int64_t m = restart > 0 ? restart : 40;
    return (m + 2) * Nb + (m + 1) * (m + 4) + 4 * (m + 2) + 64;            // V, w, H, givens, g, y, h2
}

extern "C" int64_t sktt_krylov_work(const sktt_local_op* op, int method, int restart) {
    if (!op) return -1;
    int64_t Nb = local_dim_bound(op);
    // matvec scratch | solver | f and u in the tiled layout | operator image
    return local_mv_work(op) + solver_core_work(method, restart, Nb) + 2 * Nb + 64 + local_image_bound(op);
}

// ------------------------------------------------------------------------------------------------
// CG kernels.  slots (doubles): [0] rr_old/new ping-pong: rr[0], rr[1]; [2] pAp (2 doubles re,im);
// [4] breakdown flag (as double); [5] |f|^2
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void cg_update_xr_kernel(long long n, T* __restrict__ u, T* __restrict__ r, const T* __restrict__ p,
                                    const T* __restrict__ Ap, const double* rr_old, const double* pAp, double* flag,
                                    T* partial, unsigned* counter, double* rr_new) {
    __shared__ T sh[32];
    __shared__ bool last;
    const double denom = pAp[0];
    const double alpha = denom > 0.0 ? rr_old[0] / denom : 0.0;
    if (!(denom > 0.0) && blockIdx.x == 0 && threadIdx.x == 0 && rr_old[0] > 0.0) flag[0] = 1.0;  // not HPD
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        T pi = p[i], api = Ap[i];
        u[i] = Num<T>::add(u[i], Num<T>::scale(pi, alpha));
        T ri = Num<T>::sub(r[i], Num<T>::scale(api, alpha));
        r[i] = ri;
        acc += Num<T>::abs2(ri);
    }
    double* shd = (double*)sh;
    acc = block_sum<double>(acc, shd);
    double* part = (double*)partial;
    if (threadIdx.x == 0) {
        part[blockIdx.x] = acc;
        __threadfence();
        unsigned done = atomicAdd(counter, 1u);
        last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence();
        double s = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += __ldcg(part + i);
        s = block_sum<double>(s, shd);
        if (threadIdx.x == 0) {
            rr_new[0] = s;
            *counter = 0;
        }
    }
}

template <typename T>
__global__ void cg_update_p_kernel(long long n, T* __restrict__ p, const T* __restrict__ r, const double* rr_old,
                                   const double* rr_new) {
    const double beta = rr_old[0] > 0.0 ? rr_new[0] / rr_old[0] : 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p[i] = Num<T>::add(r[i], Num<T>::scale(p[i], beta));
}

// r = f - Au ; p = r
template <typename T>
__global__ void residual_init_kernel(long long n, const T* __restrict__ f, const T* __restrict__ Au, T* __restrict__ r,
                                     T* __restrict__ p) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        T v = Num<T>::sub(f[i], Au[i]);
        r[i] = v;
        if (p) p[i] = v;
    }
}

static inline int ew_blocks(sktt_ctx* ctx, long long n) {
    long long b = (n + 255) / 256;
    long long cap = 4LL * ctx->sm_count;
    if (cap > SKTT_DOT_MAX_BLOCKS) cap = SKTT_DOT_MAX_BLOCKS;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <typename T>
static int cg_impl(sktt_ctx* ctx, int dtype, const KOp& op, const T* f, T* u, double tol, int max_iters,
                   T* work, int* iters_host, double* relres_host) {
    const long long N = op.N;
    T* mvwork = work;
    T* r = work + local_mv_work(&op.op);
    T* p = r + N;
    T* Ap = p + N;
    double* slots = (double*)ctx->scratch;  // [0..1] rr ping-pong, [2..3] pAp, [4] flag, [5] |f|^2
    unsigned* counter = (unsigned*)((char*)ctx->scratch + SKTT_SCRATCH_COUNTER_OFF);
    T* partial = (T*)((char*)ctx->scratch + SKTT_SCRATCH_PARTIAL_OFF);
    double* mbox = (double*)ctx->mailbox;
    const int nb = ew_blocks(ctx, N);
    SKTT_CUDA(ctx, cudaMemsetAsync(slots, 0, 8 * sizeof(double), ctx->stream));
    SKTT_TRY(blas1_dot(ctx, dtype, N, f, f, slots + 5));
    // +6 is scratch for the imaginary part written by blas1_dot at [5]+1
    SKTT_TRY(kop_matvec(ctx, dtype, op, u, Ap, mvwork));
    residual_init_kernel<T><<<nb, 256, 0, ctx->stream>>>(N, f, Ap, r, p);
    SKTT_LAUNCH_CHECK(ctx);
    double* rr_init = (double*)((char*)ctx->scratch + 256);  // 2 doubles (re, im)
    SKTT_TRY(blas1_dot(ctx, dtype, N, r, r, rr_init));
    SKTT_CUDA(ctx, cudaMemcpyAsync(slots, rr_init, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    SKTT_CUDA(ctx, cudaMemcpyAsync(mbox, slots, 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SKTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const double fnorm2 = mbox[5];
    double rr = mbox[0];
    if (fnorm2 == 0.0) {
        // zero right-hand side: solution is zero
        SKTT_CUDA(ctx, cudaMemsetAsync(u, 0, (size_t)N * sizeof(T), ctx->stream));
        if (iters_host) *iters_host = 0;
        if (relres_host) *relres_host = 0.0;
        return 0;
    }
    const double target2 = tol * tol * fnorm2;
    const int batch = 8;
    cudaEvent_t ev[2];
    SKTT_CUDA(ctx, cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    SKTT_CUDA(ctx, cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    int status = 0, launched = 0, batches = 0;
    bool converged = rr <= target2;
    while (!converged && launched < max_iters) {
        const int slot = batches & 1;
        for (int b = 0; b < batch && launched < max_iters; ++b, ++launched) {
            double* rr_old = slots + (launched & 1);
            double* rr_new = slots + ((launched + 1) & 1);
            status = kop_matvec(ctx, dtype, op, p, Ap, mvwork);
            if (status) break;
            status = blas1_dot(ctx, dtype, N, p, Ap, slots + 2);
            if (status) break;
            cg_update_xr_kernel<T><<<nb, 256, 0, ctx->stream>>>(N, u, r, p, Ap, rr_old, slots + 2, slots + 4, partial,
                                                               counter, rr_new);
            ctx->launches++;
            cg_update_p_kernel<T><<<nb, 256, 0, ctx->stream>>>(N, p, r, rr_old, rr_new);
            ctx->launches++;
        }
        if (status) break;
        // mailbox record of this batch: latest |r|^2 and the breakdown flag
        cudaMemcpyAsync(mbox + 16 + slot * 8, slots + (launched & 1), sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(mbox + 16 + slot * 8 + 1, slots + 4, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        cudaEventRecord(ev[slot], ctx->stream);
        // inspect the batch queued before this one: the GPU still has the current batch to chew on
        if (batches >= 1) {
            const int prev = (batches - 1) & 1;
            cudaEventSynchronize(ev[prev]);
            if (mbox[16 + prev * 8 + 1] != 0.0) { status = SKTT_ERR_NOCONV; break; }
            if (mbox[16 + prev * 8] <= target2) converged = true;
        }
        ++batches;
    }
    cudaStreamSynchronize(ctx->stream);
    cudaEventDestroy(ev[0]);
    cudaEventDestroy(ev[1]);
    if (status == SKTT_ERR_NOCONV) return sktt_fail(ctx, SKTT_ERR_NOCONV, "cg: operator is not Hermitian positive definite (p^H A p <= 0)");
    if (status) return status;
    // final state
    SKTT_CUDA(ctx, cudaMemcpy(mbox, slots, 8 * sizeof(double), cudaMemcpyDeviceToHost));
    rr = mbox[launched & 1];
    if (mbox[4] != 0.0) return sktt_fail(ctx, SKTT_ERR_NOCONV, "cg: operator is not Hermitian positive definite (p^H A p <= 0)");
    if (iters_host) *iters_host = launched;
    if (relres_host) *relres_host = sqrt(rr / fnorm2);
    if (!(rr <= target2)) return sktt_fail(ctx, SKTT_ERR_NOCONV, "cg: no convergence within max_iters");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Preconditioned CG on the prepared (tiled) operator in the Chronopoulos-Gear form: ONE synchronisation point per
// iteration.  z = P^-1 r (mode preconditioner, below), the matvec w = A z leaves delta = <w, z> and gamma = <r, z>
// behind (fused into the epilogue of its second kernel), and one vector kernel derives alpha, beta from them, updates
// p, s = A p, x, r and reduces |r|^2 for the convergence test of the next iteration.  Four launches per iteration; a
// converged solve raises a device flag that turns every later launch of the batch into a no-op, so the host only looks
// at a pinned mailbox one batch behind.
//   state (doubles in ctx->scratch): [8] delta, [9] gamma (matvec output) | [12 + 2 k] gamma_old, alpha_old ping-pong
//                                    [20] |r|^2 of the current residual
//   ints at byte 192: [0] done, [1] iterations, [2] breakdown
//
// Mode preconditioner: P = I (x) Abar (x) I with Abar = sum_{b,b2} tr(L_b) tr(R_b2) A[b,:,:,b2] / (r r2), the partial
// trace of the micro operator over both rank indices (Hermitian positive definite whenever the micro operator is).  It
// removes the conditioning of the mode factor -- all of it when the interface bases are random (first sweep), the
// 64-point Laplacian's share of it later on.  On tiled vectors [n][a][r2 + 4] its inverse is ONE plain GEMM
// Z[m][:] = sum_n Abar^-1[m][n] V[n][:].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mode_precond_kernel(int r, int R, int n, int R2, int r2, const double* __restrict__ L, const double* __restrict__ A,
                    const double* __restrict__ Rt, double* __restrict__ Pinv) {
    // one CTA: traces, Abar, Gauss-Jordan inverse without pivoting (SPD), symmetrised; identity if a pivot is not positive
    extern __shared__ double sm[];
    double* M = sm;                 // [n][n + 1]
    double* X = M + n * (n + 1);    // [n][n + 1]
    double* tl = X + n * (n + 1);   // [R]
    double* tr = tl + R;            // [R2]
    __shared__ int bad;
    const int tid = threadIdx.x, ld = n + 1;
    if (tid == 0) bad = 0;
    for (int b = tid; b < R + R2; b += blockDim.x) {
        double t = 0.0;
        if (b < R) for (int c = 0; c < r; ++c) t += L[((size_t)c * R + b) * r + c];
        else for (int c = 0; c < r2; ++c) t += Rt[((size_t)c * R2 + (b - R)) * r2 + c];
        (b < R ? tl[b] : tr[b - R]) = t;
    }
    __syncthreads();
    const double scale = 1.0 / ((double)r * (double)r2);
    for (int e = tid; e < n * n; e += blockDim.x) {
        const int i = e / n, j = e % n;
        double a = 0.0;
        for (int b = 0; b < R; ++b)
            for (int q = 0; q < R2; ++q) a = fma(tl[b] * tr[q], A[(((size_t)b * n + i) * n + j) * R2 + q], a);
        M[i * ld + j] = a * scale;
        X[i * ld + j] = i == j ? 1.0 : 0.0;
    }
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        const double piv = M[k * ld + k];
        if (!(piv > 0.0) || !isfinite(piv)) {
            if (tid == 0) bad = 1;
            break;                                             // uniform
        }
        const double inv = 1.0 / piv;
        __syncthreads();
        for (int j = tid; j < n; j += blockDim.x) {
            M[k * ld + j] *= inv;
            X[k * ld + j] *= inv;
        }
        __syncthreads();
        for (int e = tid; e < n * n; e += blockDim.x) {
            const int i = e / n, j = e % n;
            if (i == k) continue;
            const double f = M[i * ld + k];
            if (j != k) M[i * ld + j] = fma(-f, M[k * ld + j], M[i * ld + j]);
            X[i * ld + j] = fma(-f, X[k * ld + j], X[i * ld + j]);
        }
        __syncthreads();
        for (int i = tid; i < n; i += blockDim.x)
            if (i != k) M[i * ld + k] = 0.0;
        __syncthreads();
    }
    __syncthreads();
    const int isbad = bad;
    for (int e = tid; e < n * n; e += blockDim.x) {
        const int i = e / n, j = e % n;
        Pinv[e] = isbad ? (i == j ? 1.0 : 0.0) : 0.5 * (X[i * ld + j] + X[j * ld + i]);
    }
}

// Z[m][j] = sum_n Pinv[m][n] V[n][j] for j < cols (tiled vectors: n rows of `cols` doubles); one CTA per 32 columns
__global__ void __launch_bounds__(256)
mode_precond_apply_kernel(int n, long long cols, const double* __restrict__ Pinv, const double* __restrict__ V,
                          double* __restrict__ Z, const int* __restrict__ skip) {
    extern __shared__ double sm[];
    if (skip && *skip) return;
    double* P = sm;                  // [n][n + 1]
    double* Vs = P + n * (n + 1);    // [n][32]
    const int tid = threadIdx.x, ld = n + 1;
    const long long j0 = (long long)blockIdx.x * 32;
    for (int e = tid; e < n * n; e += 256) P[(e / n) * ld + e % n] = Pinv[e];
    for (int e = tid; e < n * 32; e += 256) {
        const int k = e >> 5, j = e & 31;
        Vs[e] = j0 + j < cols ? V[(size_t)k * cols + j0 + j] : 0.0;
    }
    __syncthreads();
    const int j = tid & 31;
    for (int i = tid >> 5; i < n; i += 8) {
        double a0 = 0.0, a1 = 0.0;
        int k = 0;
        for (; k + 1 < n; k += 2) {
            a0 = fma(P[i * ld + k], Vs[k * 32 + j], a0);
            a1 = fma(P[i * ld + k + 1], Vs[(k + 1) * 32 + j], a1);
        }
        if (k < n) a0 = fma(P[i * ld + k], Vs[k * 32 + j], a0);
        if (j0 + j < cols) Z[(size_t)i * cols + j0 + j] = a0 + a1;
    }
}

__global__ void pcgear_update_kernel(long long n, double* __restrict__ x, double* r, double* __restrict__ p,
                                     double* __restrict__ s, const double* __restrict__ w, const double* z,
                                     const double* dots, const double* st_cur, double* st_next, int* flags, double* rr_slot,
                                     double* partial, unsigned* counter, double target2) {
    __shared__ double sh[32];
    __shared__ bool last;
    if (flags[0]) return;
    const double delta = dots[0], gamma = dots[1];
    const bool lead = blockIdx.x == 0 && threadIdx.x == 0;
    if (rr_slot[0] <= target2) {                  // x already solves the system to the requested accuracy
        if (lead) flags[0] = 1;
        return;
    }
    const double gamma_old = st_cur[0], alpha_old = st_cur[1];
    const bool first = gamma_old < 0.0;
    const double beta = first ? 0.0 : gamma / gamma_old;
    const double denom = first ? delta : delta - beta * gamma / alpha_old;
    if (!(denom > 0.0) || !(gamma > 0.0)) {       // not Hermitian positive definite (operator or preconditioner)
        if (lead) {
            flags[2] = 1;
            flags[0] = 1;
        }
        return;
    }
    const double alpha = gamma / denom;
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double pi = fma(beta, p[i], z[i]);
        const double si = fma(beta, s[i], w[i]);
        p[i] = pi;
        s[i] = si;
        x[i] = fma(alpha, pi, x[i]);
        const double ri = fma(-alpha, si, r[i]);
        r[i] = ri;
        acc = fma(ri, ri, acc);
    }
    acc = block_sum<double>(acc, sh);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = acc;
        __threadfence();
        const unsigned done = atomicAdd(counter, 1u);
        last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence();
        double t = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += __ldcg(partial + i);
        t = block_sum<double>(t, sh);
        if (threadIdx.x == 0) {
            rr_slot[0] = t;                       // read by the next iteration's update (stream order)
            *counter = 0;
            st_next[0] = gamma;
            st_next[1] = alpha;
            flags[1] += 1;
        }
    }
}

// Work layout of the tiled CG (doubles): matvec scratch | r | p | w | s | z | Pinv[n][n]
struct CgTiledBufs {
    double *mvwork, *r, *p, *w, *s, *z, *pinv;
};
static CgTiledBufs cg_tiled_bufs(const KOp& op, double* work) {
    CgTiledBufs b;
    b.mvwork = work;
    b.r = work + local_mv_work(&op.op);
    b.p = b.r + op.N;
    b.w = b.p + op.N;
    b.s = b.w + op.N;
    b.z = b.s + op.N;
    b.pinv = b.z + op.N;
    return b;
}

#define MODE_PRECOND_MAX_N 96      // both n x n buffers of the setup kernel must fit in shared memory

static int cg_tiled_precond_setup(sktt_ctx* ctx, const KOp& op, const CgTiledBufs& b) {
    const sktt_local_op& o = op.op;
    const int n = (int)o.n;
    const size_t smem = ((size_t)2 * n * (n + 1) + o.R + o.R2) * sizeof(double);
    SKTT_ONCE_PER_DEVICE(ctx);
    if (!configured) {
        SKTT_CUDA(ctx, cudaFuncSetAttribute(mode_precond_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        SKTT_CUDA(ctx, cudaFuncSetAttribute(mode_precond_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        configured = true;
    }
    mode_precond_kernel<<<1, 256, smem, ctx->stream>>>((int)o.r, (int)o.R, n, (int)o.R2, (int)o.r3, (const double*)o.Lst,
                                                        (const double*)o.A1, (const double*)o.Rst, b.pinv);
    SKTT_LAUNCH_CHECK(ctx);
    return 0;
}

// The CG loop proper: on entry b.r holds the residual of u (tiled layout) and rr its squared norm; iterates until
// <r, r> <= target2 by the recurrence, u updated in place.  iters_host accumulates.
static int cg_tiled_core(sktt_ctx* ctx, const KOp& op, const CgTiledBufs& b, double* u, double rr, double target2,
                         int max_iters, bool precond, int* iters_host, double* rr_host, bool* converged) {
    const long long N = op.N;
    double* slots = (double*)ctx->scratch;
    double* dots = slots + 8;
    double* st = slots + 12;                                   // [2][2]
    double* rr_slot = slots + 20;
    int* flags = (int*)((char*)ctx->scratch + 192);
    unsigned* counter = (unsigned*)((char*)ctx->scratch + SKTT_SCRATCH_COUNTER_OFF) + 8;
    unsigned* counter2 = counter + 1;
    double* dpart = (double*)((char*)ctx->scratch + SKTT_SCRATCH_PARTIAL_OFF) + 2 * SKTT_DOT_MAX_BLOCKS;
    double* upart = dpart + 2 * SKTT_DOT_MAX_BLOCKS;
    double* mbox = (double*)ctx->mailbox;
    const int nb = ew_blocks(ctx, N);
    const sktt_local_op& o = op.op;
    const int n = (int)o.n;
    const long long cols = N / n;
    const size_t smem_apply = ((size_t)n * (n + 1) + (size_t)n * 32) * sizeof(double);
    SKTT_CUDA(ctx, cudaMemsetAsync(b.p, 0, (size_t)3 * N * sizeof(double), ctx->stream));      // p, w, s
    const double init_state[5] = {-1.0, 0.0, -1.0, 0.0, rr};
    memcpy(mbox + 32, init_state, sizeof(init_state));
    SKTT_CUDA(ctx, cudaMemcpyAsync(st, mbox + 32, 4 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    SKTT_CUDA(ctx, cudaMemcpyAsync(rr_slot, mbox + 36, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    SKTT_CUDA(ctx, cudaMemsetAsync(flags, 0, 4 * sizeof(int), ctx->stream));
    const int batch = 4;
    int status = 0, launched = 0, batches = 0;
    bool finished = false;
    int* mflags = (int*)(mbox + 16);                           // [2][4] ints
    while (!finished && launched <= max_iters) {
        const int slot = batches & 1;
        for (int q = 0; q < batch && launched <= max_iters; ++q, ++launched) {
            const double* z = precond ? b.z : b.r;             // unpreconditioned: z aliases r, <r, z> = <r, r>
            if (precond) {
                mode_precond_apply_kernel<<<(unsigned)((cols + 31) / 32), 256, smem_apply, ctx->stream>>>(n, cols, b.pinv,
                                                                                                          b.r, b.z, flags);
                ctx->launches++;
            }
            status = sktt_fused_matvec_tiled_dots(ctx, fused_rpad(o.r), o.R, o.m, o.n, (const double*)o.image, z, b.w, b.mvwork, z,
                                                  precond ? b.r : nullptr, dpart, counter, dots, flags);
            if (status) break;
            pcgear_update_kernel<<<nb, 256, 0, ctx->stream>>>(N, u, b.r, b.p, b.s, b.w, z, dots, st + 2 * (launched & 1),
                                                              st + 2 * ((launched + 1) & 1), flags, rr_slot, upart, counter2,
                                                              target2);
            ctx->launches++;
        }
        if (status) break;
        cudaMemcpyAsync(mflags + 4 * slot, flags, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        cudaEventRecord(ctx->ev[slot], ctx->stream);
        if (batches >= 1) {                                    // look at the batch queued before this one
            const int prev = (batches - 1) & 1;
            cudaEventSynchronize(ctx->ev[prev]);
            if (mflags[4 * prev]) finished = true;
        }
        ++batches;
    }
    if (status) return status;
    SKTT_CUDA(ctx, cudaMemcpyAsync(mbox, rr_slot, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SKTT_CUDA(ctx, cudaMemcpyAsync(mflags, flags, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SKTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (iters_host) *iters_host += mflags[1];
    if (rr_host) *rr_host = mbox[0];
    if (mflags[2]) return sktt_fail(ctx, SKTT_ERR_NOCONV, "cg: operator is not Hermitian positive definite (p^H A p <= 0)");
    if (converged) *converged = mflags[0] != 0;
    else if (!mflags[0]) return sktt_fail(ctx, SKTT_ERR_NOCONV, "cg: no convergence within max_iters");
    return 0;
}

// r = f - A u and the two squared norms the drivers need, one host synchronisation: returns |f|^2 and |r|^2
static int cg_tiled_residual(sktt_ctx* ctx, const KOp& op, const CgTiledBufs& b, const double* f, const double* u,
                             double* fnorm2, double* rr) {
    const long long N = op.N;
    const sktt_local_op& o = op.op;
    double* slots = (double*)ctx->scratch;
    double* mbox = (double*)ctx->mailbox;
    const int nb = ew_blocks(ctx, N);
    if (fnorm2) SKTT_TRY(blas1_dot(ctx, SKTT_F64, N, f, f, slots + 5));
    SKTT_CUDA(ctx, cudaMemsetAsync(b.w, 0, (size_t)N * sizeof(double), ctx->stream));   // padding columns are never written
    SKTT_TRY(sktt_fused_matvec_tiled(ctx, fused_rpad(o.r), o.R, o.m, o.n, (const double*)o.image, u, b.w, b.mvwork));
    residual_init_kernel<double><<<nb, 256, 0, ctx->stream>>>(N, f, b.w, b.r, (double*)nullptr);
    SKTT_LAUNCH_CHECK(ctx);
    SKTT_TRY(blas1_dot(ctx, SKTT_F64, N, b.r, b.r, slots + 2));
    SKTT_CUDA(ctx, cudaMemcpyAsync(mbox, slots, 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SKTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (fnorm2) *fnorm2 = mbox[5];
    *rr = mbox[2];
    return 0;
}

static int cg_tiled_impl(sktt_ctx* ctx, const KOp& op, const double* f, double* u, double tol, int max_iters,
                         double* work, int* iters_host, double* relres_host) {
    const CgTiledBufs b = cg_tiled_bufs(op, work);
    double fnorm2 = 0.0, rr = 0.0;
    SKTT_TRY(cg_tiled_residual(ctx, op, b, f, u, &fnorm2, &rr));
    if (iters_host) *iters_host = 0;
    if (fnorm2 == 0.0) {
        SKTT_CUDA(ctx, cudaMemsetAsync(u, 0, (size_t)op.N * sizeof(double), ctx->stream));
        if (relres_host) *relres_host = 0.0;
        return 0;
    }
    int st = cg_tiled_core(ctx, op, b, u, rr, tol * tol * fnorm2, max_iters, false, iters_host, &rr, nullptr);
    if (relres_host) *relres_host = sqrt(rr / fnorm2);
    return st;
}

// Whole micro solve on the prepared operator: warm start (dropped when it is worse than the zero vector), CG, then the
// TRUE residual f - A u is recomputed and CG restarted from it until that residual is below tol * |f| or stops improving
// (the eps * cond floor any backward-stable solver, LU included, ends at).  A handful of host synchronisations per solve.
static int cg_tiled_refined_host(sktt_ctx* ctx, const KOp& op, const double* f, double* u, double tol, int max_iters,
                                 int max_cycles, double* work, int* iters_host, double* relres_host, int* cycles_host) {
    const CgTiledBufs b = cg_tiled_bufs(op, work);
    const long long N = op.N;
    double fnorm2 = 0.0, rr = 0.0;
    SKTT_TRY(cg_tiled_residual(ctx, op, b, f, u, &fnorm2, &rr));
    if (iters_host) *iters_host = 0;
    if (cycles_host) *cycles_host = 0;
    if (fnorm2 == 0.0) {
        SKTT_CUDA(ctx, cudaMemsetAsync(u, 0, (size_t)N * sizeof(double), ctx->stream));
        if (relres_host) *relres_host = 0.0;
        return 0;
    }
    if (!(rr < fnorm2)) {                                      // the warm start is no better than zero: drop it
        SKTT_CUDA(ctx, cudaMemsetAsync(u, 0, (size_t)N * sizeof(double), ctx->stream));
        SKTT_CUDA(ctx, cudaMemcpyAsync(b.r, f, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        rr = fnorm2;
    }
    // CG runs unpreconditioned first: with the random interface bases of a first sweep the micro operator is close to a
    // multiple of the identity and converges in a dozen iterations.  A solve that needs more than CG_PLAIN_ITERS builds
    // the mode preconditioner and restarts preconditioned from its current iterate (and stays preconditioned).
    const int CG_PLAIN_ITERS = 40;
    const bool can_precond = op.op.n <= MODE_PRECOND_MAX_N && !(ctx->debug & 4);
    bool precond = false;
    double prev = 1e300, relres = sqrt(rr / fnorm2);
    int status = 0;
    for (int cycle = 0; cycle < max_cycles; ++cycle) {
        if (relres <= tol || relres > 0.5 * prev) break;
        prev = relres;
        double rr_rec = rr;
        bool conv = false;
        const double target2 = 0.25 * tol * tol * fnorm2;
        if (!precond) {
            status = cg_tiled_core(ctx, op, b, u, rr, target2, can_precond ? CG_PLAIN_ITERS : max_iters, false, iters_host,
                                   &rr_rec, &conv);
            if (status == 0 && !conv && can_precond) {
                SKTT_TRY(cg_tiled_precond_setup(ctx, op, b));
                precond = true;
                status = cg_tiled_core(ctx, op, b, u, rr_rec, target2, max_iters, true, iters_host, &rr_rec, &conv);
            }
        } else {
            status = cg_tiled_core(ctx, op, b, u, rr, target2, max_iters, true, iters_host, &rr_rec, &conv);
        }
        if (status != 0 && status != SKTT_ERR_NOCONV) return status;
        SKTT_TRY(cg_tiled_residual(ctx, op, b, f, u, nullptr, &rr));
        relres = sqrt(rr / fnorm2);
        if (cycles_host) *cycles_host = cycle + 1;
        if (status == SKTT_ERR_NOCONV || !conv) break;          // breakdown / iteration limit: judged by the true residual
    }
    if (relres_host) *relres_host = relres;
    return 0;
}

// The persistent-kernel form of the same solve (fused.cu): one cooperative launch, one host synchronisation.  Returns 0
// with *finished = false when a CG run hit PERSISTENT_CG_ITERS without converging -- the caller then continues with
// the host-driven, preconditioned loop from the iterate left in u.
#define PERSISTENT_CG_ITERS 300
static int cg_tiled_persistent(sktt_ctx* ctx, const KOp& op, const double* f, double* u, double tol, int max_cycles,
                               double* work, int* iters_host, double* relres_host, int* cycles_host, bool* finished,
                               double* result_dev = nullptr, const double* f_nat = nullptr, double* u_nat = nullptr) {
    const CgTiledBufs b = cg_tiled_bufs(op, work);
    const sktt_local_op& o = op.op;
    double* part = (double*)((char*)ctx->scratch + SKTT_SCRATCH_BULK_OFF);
    double* outd = part + 4 * 256;
    double* mbox = (double*)ctx->mailbox;
    SKTT_TRY(sktt_scratch_reserve(ctx, SKTT_SCRATCH_BULK_OFF + (4 * 256 + 128) * sizeof(double)));
    part = (double*)((char*)ctx->scratch + SKTT_SCRATCH_BULK_OFF);
    outd = part + 4 * 256;
    // the matvec never writes the four padding columns of w; they enter r = f - w and every norm, so they must be zero
    // (with natural-layout operands the kernel clears w itself while it tiles f and u)
    if (!f_nat) SKTT_CUDA(ctx, cudaMemsetAsync(b.w, 0, (size_t)op.N * sizeof(double), ctx->stream));
    if (result_dev) {                                          // deferred form: the outcome stays on the device
        return sktt_fused_pcg_persistent(ctx, fused_rpad(o.r), o.R, o.m, o.n, (const double*)o.image, f, u, b.r, b.p, b.s, b.w,
                                         b.mvwork, tol, PERSISTENT_CG_ITERS, max_cycles, 0, 0, part, result_dev, f_nat, u_nat,
                                         o.r, o.r3);
    }
    SKTT_TRY(sktt_fused_pcg_persistent(ctx, fused_rpad(o.r), o.R, o.m, o.n, (const double*)o.image, f, u, b.r, b.p, b.s, b.w, b.mvwork,
                                       tol, PERSISTENT_CG_ITERS, max_cycles, 0, 0, part, outd));
    SKTT_CUDA(ctx, cudaMemcpyAsync(mbox + 40, outd, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SKTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (iters_host) *iters_host = (int)mbox[40];
    if (relres_host) *relres_host = mbox[41];
    if (cycles_host) *cycles_host = (int)mbox[42];
    *finished = mbox[43] == 0.0;
    return 0;
}

static int cg_tiled_refined(sktt_ctx* ctx, const KOp& op, const double* f, double* u, double tol, int max_iters,
                            int max_cycles, double* work, int* iters_host, double* relres_host, int* cycles_host) {
    if (!(ctx->debug & 16) && ctx->sm_count <= 256) {
        bool finished = false;
        int it0 = 0;
        SKTT_TRY(cg_tiled_persistent(ctx, op, f, u, tol, max_cycles, work, &it0, relres_host, cycles_host, &finished));
        if (finished) {
            if (iters_host) *iters_host = it0;
            return 0;
        }
        int it1 = 0;
        int st = cg_tiled_refined_host(ctx, op, f, u, tol, max_iters, max_cycles, work, &it1, relres_host, cycles_host);
        if (iters_host) *iters_host = it0 + it1;
        return st;
    }
    return cg_tiled_refined_host(ctx, op, f, u, tol, max_iters, max_cycles, work, iters_host, relres_host, cycles_host);
}

// Deferred form of sktt_krylov_solve_refined: the same solve as ONE cooperative launch with no host synchronisation.  The
// outcome {iterations, true relative residual, cycles, status (0 ok, 1 iteration limit, 2 breakdown)} is left in
// result_dev[0..3] (device memory) for the caller to inspect later -- a sweep queues all its micro steps and looks at the
// outcomes once, instead of draining the GPU after every solve.  There is no host-driven continuation here: status 1 / 2
// or a residual above the caller's acceptance level mean "redo this solve synchronously".
extern "C" int sktt_krylov_solve_refined_async(sktt_ctx* ctx, int dtype, const sktt_local_op* op, const void* f, void* u,
                                               double tol, int max_cycles, void* work, double* result_dev) {
    if (!ctx || !op || !f || !u || !work || !result_dev) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (dtype != SKTT_F64 || op->sites != 1 || (ctx->debug & 16) || ctx->sm_count > 256)
        return sktt_fail(ctx, SKTT_ERR_ARG, "krylov_solve_refined_async: unsupported operator");
    KOp k;
    k.op = *op;
    const int64_t Nb = local_dim_bound(op);
    int64_t off = local_mv_work(op) + solver_core_work(0, 0, Nb);
    off += off & 1;
    double* w = (double*)work;
    double* ft = w + off;
    double* ut = ft + Nb + (Nb & 1);
    double* image = ut + Nb + (Nb & 1) + 16;
    if (!k.op.image) SKTT_TRY(sktt_local_op_prepare(ctx, dtype, &k.op, (void*)image));
    if (!k.op.image) return sktt_fail(ctx, SKTT_ERR_ARG, "krylov_solve_refined_async: unsupported operator");
    k.tiled = true;
    k.N = sktt_fused_tiled_len(fused_rpad(k.op.r), k.op.n);
    // f and u stay in the caller's natural layout: the kernel tiles them, clears w and writes the solution back itself
    bool finished = false;
    return cg_tiled_persistent(ctx, k, ft, ut, tol, max_cycles, w, nullptr, nullptr, nullptr, &finished, result_dev,
                               (const double*)f, (double*)u);
}

// ------------------------------------------------------------------------------------------------
// GMRES(m) with CGS2.  Device-side small state: H [(m+1) x m] column-major by Arnoldi step,
// cs/sn Givens, g (rhs of the least-squares problem), y.
// ------------------------------------------------------------------------------------------------
template <typename T>
struct GivensState {
    T* H;      // [(m+1)*m], column j at H + j*(m+1)
    T* cs;     // [m]  (real stored in T)
    T* sn;     // [m]
    T* g;      // [m+1]
    T* y;      // [m]
};

// v = w / ||w||, h_next = ||w||; uses nrm2 = slot (re) computed by blas1_dot(w,w)
template <typename T>
__global__ void scale_to_unit_kernel(long long n, const T* __restrict__ w, const double* nrm2, T* __restrict__ v,
                                     T* hnext) {
    const double nrm = sqrt(nrm2[0] > 0.0 ? nrm2[0] : 0.0);
    const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
    if (hnext && blockIdx.x == 0 && threadIdx.x == 0) *hnext = Num<T>::from(nrm, 0.0);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        v[i] = Num<T>::scale(w[i], inv);
}

// h (column j, entries 0..j) += h2 (second CGS pass), then apply previous rotations, make a new one.
template <typename T>
__global__ void gmres_givens_kernel(int j, int m, T* H, const T* h2, T* cs, T* sn, T* g, double* resid_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    T* h = H + (size_t)j * (m + 1);
    for (int i = 0; i <= j; ++i) h[i] = Num<T>::add(h[i], h2[i]);
    for (int i = 0; i < j; ++i) {
        T t0 = Num<T>::add(Num<T>::mul(cs[i], h[i]), Num<T>::mul(sn[i], h[i + 1]));
        T t1 = Num<T>::sub(Num<T>::mul(cs[i], h[i + 1]), Num<T>::mul(Num<T>::conj(sn[i]), h[i]));
        h[i] = t0;
        h[i + 1] = t1;
    }
    // new rotation G = [c, s; -conj(s), c] (c real) with G [a; b] = [rho; 0]
    T a = h[j], b = h[j + 1];
    double na = sqrt(Num<T>::abs2(a)), nb = sqrt(Num<T>::abs2(b));
    double nrm = sqrt(na * na + nb * nb);
    T c, sv;
    if (nrm == 0.0) {
        c = Num<T>::one();
        sv = Num<T>::zero();
    } else if (na == 0.0) {
        c = Num<T>::zero();
        sv = Num<T>::scale(Num<T>::conj(b), 1.0 / nb);
    } else {
        c = Num<T>::from(na / nrm, 0.0);
        sv = Num<T>::scale(Num<T>::mul(Num<T>::scale(a, 1.0 / na), Num<T>::conj(b)), 1.0 / nrm);
    }
    cs[j] = c;
    sn[j] = sv;
    h[j] = Num<T>::add(Num<T>::mul(c, a), Num<T>::mul(sv, b));
    h[j + 1] = Num<T>::zero();
    T gj = g[j];
    g[j] = Num<T>::mul(c, gj);
    g[j + 1] = Num<T>::neg(Num<T>::mul(Num<T>::conj(sv), gj));
    resid_out[0] = sqrt(Num<T>::abs2(g[j + 1]));
}

template <typename T>
__global__ void gmres_backsolve_kernel(int k, int m, const T* H, const T* g, T* y) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int i = k - 1; i >= 0; --i) {
        T s = g[i];
        for (int c = i + 1; c < k; ++c) s = Num<T>::sub(s, Num<T>::mul(H[(size_t)c * (m + 1) + i], y[c]));
        T d = H[(size_t)i * (m + 1) + i];
        y[i] = Num<T>::abs2(d) > 0.0 ? Num<T>::div(s, d) : Num<T>::zero();
    }
}

template <typename T>
__global__ void gmres_init_g_kernel(int m, T* g, const double* beta2) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= m) g[i] = i == 0 ? Num<T>::from(sqrt(beta2[0] > 0.0 ? beta2[0] : 0.0), 0.0) : Num<T>::zero();
}

template <typename T>
static int gmres_impl(sktt_ctx* ctx, int dtype, const KOp& op, int restart, const T* f, T* u, double tol,
                      int max_iters, T* work, int* iters_host, double* relres_host) {
    const long long N = op.N;
    const int m = restart > 0 ? restart : 40;
    T* mvwork = work;
    T* V = work + local_mv_work(&op.op);     // [(m+1)][N]
    T* w = V + (size_t)(m + 1) * N;      // [N]
    T* H = w + N;                        // [(m+1)*m]
    T* cs = H + (size_t)(m + 1) * m;
    T* sn = cs + m;
    T* g = sn + m;                       // [m+1]
    T* y = g + (m + 1);                  // [m]
    T* h2 = y + m;                       // [m+1]
    double* slots = (double*)ctx->scratch;  // [0] nrm2 (2), [2] resid est, [5] |f|^2 (2)
    double* mbox = (double*)ctx->mailbox;
    const int nb = ew_blocks(ctx, N);
    SKTT_TRY(blas1_dot(ctx, dtype, N, f, f, slots + 5));
    SKTT_CUDA(ctx, cudaMemcpyAsync(mbox, slots + 5, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SKTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const double fnorm = sqrt(mbox[0]);
    if (fnorm == 0.0) {
        SKTT_CUDA(ctx, cudaMemsetAsync(u, 0, (size_t)N * sizeof(T), ctx->stream));
        if (iters_host) *iters_host = 0;
        if (relres_host) *relres_host = 0.0;
        return 0;
    }
    int total = 0;
    double relres = 1.0;
    const double one[2] = {1.0, 0.0}, minus_one[2] = {-1.0, 0.0};
    while (total < max_iters) {
        // r0 = f - A u -> V[0] (normalised), g = beta e1
        SKTT_TRY(kop_matvec(ctx, dtype, op, u, w, mvwork));
        residual_init_kernel<T><<<nb, 256, 0, ctx->stream>>>(N, f, w, w, (T*)nullptr);
        SKTT_LAUNCH_CHECK(ctx);
        SKTT_TRY(blas1_dot(ctx, dtype, N, w, w, slots));
        scale_to_unit_kernel<T><<<nb, 256, 0, ctx->stream>>>(N, w, slots, V, (T*)nullptr);
        SKTT_LAUNCH_CHECK(ctx);
        gmres_init_g_kernel<T><<<(m + 256) / 256, 256, 0, ctx->stream>>>(m, g, slots);
        SKTT_LAUNCH_CHECK(ctx);
        SKTT_CUDA(ctx, cudaMemcpyAsync(mbox, slots, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        SKTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        relres = sqrt(mbox[0] > 0 ? mbox[0] : 0.0) / fnorm;
        if (relres <= tol) break;
        int k = 0;
        bool done = false;
        for (int j = 0; j < m && total < max_iters; ++j) {
            T* vj = V + (size_t)j * N;
            T* hcol = H + (size_t)j * (m + 1);
            SKTT_TRY(kop_matvec(ctx, dtype, op, vj, w, mvwork));
            for (int pass = 0; pass < 2; ++pass) {
                T* hdst = pass == 0 ? hcol : h2;
                // h = V_{0..j}^H w
                GemmDesc g1 = gemm_desc(j + 1, 1, N, V, lin_idx(N), lin_idx(1), w, lin_idx(1), lin_idx(0), hdst,
                                        lin_idx(1), lin_idx(0));
                g1.conjA = 1;
                SKTT_TRY(sktt_gemm_run(ctx, dtype, g1));
                // w -= V h
                GemmDesc g2 = gemm_desc(N, 1, j + 1, V, lin_idx(1), lin_idx(N), hdst, lin_idx(1), lin_idx(0), w,
                                        lin_idx(1), lin_idx(0));
                g2.alpha[0] = minus_one[0];
                g2.beta[0] = one[0];
                SKTT_TRY(sktt_gemm_run(ctx, dtype, g2));
            }
            SKTT_TRY(blas1_dot(ctx, dtype, N, w, w, slots));
            scale_to_unit_kernel<T><<<nb, 256, 0, ctx->stream>>>(N, w, slots, V + (size_t)(j + 1) * N, hcol + j + 1);
            SKTT_LAUNCH_CHECK(ctx);
            gmres_givens_kernel<T><<<1, 32, 0, ctx->stream>>>(j, m, H, h2, cs, sn, g, slots + 2);
            SKTT_LAUNCH_CHECK(ctx);
            ++total;
            k = j + 1;
            // peek at the residual estimate every 4 steps and at the end of the cycle
            if ((j & 3) == 3 || j == m - 1 || total >= max_iters) {
                SKTT_CUDA(ctx, cudaMemcpyAsync(mbox, slots + 2, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
                SKTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                relres = mbox[0] / fnorm;
                if (relres <= tol) { done = true; break; }
            }
        }
        // u += V_k y
        gmres_backsolve_kernel<T><<<1, 32, 0, ctx->stream>>>(k, m, H, g, y);
        SKTT_LAUNCH_CHECK(ctx);
        GemmDesc gu = gemm_desc(N, 1, k, V, lin_idx(1), lin_idx(N), y, lin_idx(1), lin_idx(0), u, lin_idx(1), lin_idx(0));
        gu.beta[0] = 1.0;
        SKTT_TRY(sktt_gemm_run(ctx, dtype, gu));
        if (done) {
            // confirm with the true residual on the next loop entry (cheap: one matvec)
            continue;
        }
    }
    if (iters_host) *iters_host = total;
    if (relres_host) *relres_host = relres;
    if (!(relres <= tol)) return sktt_fail(ctx, SKTT_ERR_NOCONV, "gmres: no convergence within max_iters");
    return 0;
}

template <typename T>
static int krylov_dispatch(sktt_ctx* ctx, int dtype, const sktt_local_op* op_in, int method, int restart, const T* f, T* u,
                           double tol, int max_iters, T* work, int* iters_host, double* relres_host) {
    KOp k;
    k.op = *op_in;
    k.tiled = false;
    k.N = local_dim(op_in);
    const int64_t Nb = local_dim_bound(op_in);
    int64_t off = local_mv_work(op_in) + solver_core_work(method, restart, Nb);
    off += off & 1;                                          // keep 16-byte alignment of what follows
    T* ft = work + off;
    T* ut = ft + Nb + (Nb & 1);
    T* image = ut + Nb + (Nb & 1) + 16;
    if (!k.op.image) SKTT_TRY(sktt_local_op_prepare(ctx, dtype, &k.op, (void*)image));
    if (k.op.image && k.op.sites == 1 && dtype == SKTT_F64) {
        // iterate on tiled-layout vectors: convert the right-hand side and the initial guess, convert the result back
        k.tiled = true;
        k.N = sktt_fused_tiled_len(fused_rpad(k.op.r), k.op.n);
        SKTT_TRY(sktt_fused_to_tiled_ex(ctx, fused_rpad(k.op.r), k.op.n, (const double*)f, (double*)ft, 0, k.op.r, k.op.r3));
        SKTT_TRY(sktt_fused_to_tiled_ex(ctx, fused_rpad(k.op.r), k.op.n, (const double*)u, (double*)ut, 0, k.op.r, k.op.r3));
    }
    const T* fs = k.tiled ? ft : f;
    T* us = k.tiled ? ut : u;
    int st;
    if (method == 0 && k.tiled)
        st = cg_tiled_impl(ctx, k, (const double*)fs, (double*)us, tol, max_iters, (double*)work, iters_host, relres_host);
    else
        st = method == 0 ? cg_impl<T>(ctx, dtype, k, fs, us, tol, max_iters, work, iters_host, relres_host)
                         : gmres_impl<T>(ctx, dtype, k, restart, fs, us, tol, max_iters, work, iters_host, relres_host);
    if (k.tiled) {
        int st2 = sktt_fused_from_tiled_ex(ctx, fused_rpad(k.op.r), k.op.n, (const double*)us, (double*)u, k.op.r, k.op.r3);
        if (st == 0) st = st2;
    }
    return st;
}

extern "C" int sktt_krylov_solve(sktt_ctx* ctx, int dtype, const sktt_local_op* op, int method, int restart,
                                 const void* f, void* u, double tol, int max_iters, void* work, int* iters_host,
                                 double* relres_host) {
    if (!ctx || !op || !f || !u || !work) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (op->sites != 1 && op->sites != 2) return sktt_fail(ctx, SKTT_ERR_ARG, "krylov: sites must be 1 or 2");
    if (method != 0 && method != 1) return sktt_fail(ctx, SKTT_ERR_ARG, "krylov: unknown method");
    if (dtype == SKTT_F64)
        return krylov_dispatch<double>(ctx, dtype, op, method, restart, (const double*)f, (double*)u, tol, max_iters,
                                       (double*)work, iters_host, relres_host);
    return krylov_dispatch<cplx>(ctx, dtype, op, method, restart, (const cplx*)f, (cplx*)u, tol, max_iters, (cplx*)work,
                                 iters_host, relres_host);
}

// One-call micro solve for operators the fused matvec covers (real, one-site, prepared shapes): see cg_tiled_refined.
// Returns SKTT_ERR_ARG with "unsupported" when the operator is not covered; the caller then drives sktt_krylov_solve.
extern "C" int sktt_krylov_solve_refined(sktt_ctx* ctx, int dtype, const sktt_local_op* op, const void* f, void* u,
                                         double tol, int max_iters, int max_cycles, void* work, int* iters_host,
                                         double* relres_host, int* cycles_host) {
    if (!ctx || !op || !f || !u || !work) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (dtype != SKTT_F64 || op->sites != 1) return sktt_fail(ctx, SKTT_ERR_ARG, "krylov_solve_refined: unsupported operator");
    KOp k;
    k.op = *op;
    const int64_t Nb = local_dim_bound(op);
    int64_t off = local_mv_work(op) + solver_core_work(0, 0, Nb);
    off += off & 1;
    double* w = (double*)work;
    double* ft = w + off;
    double* ut = ft + Nb + (Nb & 1);
    double* image = ut + Nb + (Nb & 1) + 16;
    if (!k.op.image) SKTT_TRY(sktt_local_op_prepare(ctx, dtype, &k.op, (void*)image));
    if (!k.op.image) return sktt_fail(ctx, SKTT_ERR_ARG, "krylov_solve_refined: unsupported operator");
    k.tiled = true;
    k.N = sktt_fused_tiled_len(fused_rpad(k.op.r), k.op.n);
    SKTT_TRY(sktt_fused_to_tiled_ex(ctx, fused_rpad(k.op.r), k.op.n, (const double*)f, ft, 0, k.op.r, k.op.r3));
    SKTT_TRY(sktt_fused_to_tiled_ex(ctx, fused_rpad(k.op.r), k.op.n, (const double*)u, ut, 0, k.op.r, k.op.r3));
    int st = cg_tiled_refined(ctx, k, ft, ut, tol, max_iters, max_cycles, w, iters_host, relres_host, cycles_host);
    int st2 = sktt_fused_from_tiled_ex(ctx, fused_rpad(k.op.r), k.op.n, ut, (double*)u, k.op.r, k.op.r3);
    return st ? st : st2;
}
