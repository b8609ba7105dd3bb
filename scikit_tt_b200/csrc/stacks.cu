// stacks.cu -- interface stacks, micro matrices, matrix-free micro-matvecs and micro right-hand
// sides of the ALS/MALS sweeps, expressed as chains of two-level strided contractions (gemm.cu)
// whose intermediates stay in L2.  Index names follow SURVEY.md section 3:
//   a, c : left solution rank (column side / row side)      a2, c2 : right solution rank
//   b, b2: operator ranks                                    n, m   : column / row mode index
#include "common.cuh"
#include "blas1.cuh"

// fused.cu
bool sktt_fused_supported(const sktt_ctx* ctx, int dtype, long long r, long long R, long long m, long long n,
                          long long r2, long long R2);
long long sktt_fused_image_elems(long long r, long long R, long long m, long long n);
int sktt_fused_prepare(sktt_ctx* ctx, long long r, long long R, long long m, long long n, const double* Lst,
                       const double* A, const double* Rst, double* image);
int sktt_fused_matvec(sktt_ctx* ctx, long long r, long long R, long long m, long long n, const double* image,
                      const double* v, double* y, double* work);
long long sktt_fused_tiled_len(long long r, long long n);
int sktt_fused_matvec_tiled(sktt_ctx* ctx, long long r, long long R, long long m, long long n, const double* image,
                            const double* vt, double* yt, double* T1p);

static inline Idx2 two(long long d, long long s_hi, long long s_lo) { return mk_idx(d, s_hi, s_lo); }

int sktt_fused_prepare_ex(sktt_ctx* ctx, long long r, long long R, long long m, long long n, const double* Lst,
                          const double* A, const double* Rst, double* image, int swap, long long r_act, long long r2_act,
                          long long R2_act);
int sktt_fused_to_tiled_ex(sktt_ctx* ctx, long long r, long long n, const double* src, double* dst, int swap,
                           long long r_act, long long c_act);
int sktt_fused_matvec_ex(sktt_ctx* ctx, long long r, long long R, long long m, long long n, const double* image,
                         const double* v, double* y, double* work, long long r_act, long long r2_act);
int sktt_fused_stack_update(sktt_ctx* ctx, long long r, long long R, long long m, long long n, const double* image,
                            const double* xt, double* out, double* T1p, double* part);

// stack_nat.cu: the one-launch form that reads the operands in their natural layout (solution ranks 64, operator ranks 3)
bool sktt_stack_nat_supported(const sktt_ctx* ctx, int dtype, long long rin, long long Rin, long long m, long long n,
                              long long rout, long long Rout, const void* stack, const void* x, const void* A);
int sktt_stack_nat_update(sktt_ctx* ctx, long long m, long long n, const double* stack, const double* x, const double* A,
                          double* out, double* T1p, double* part, int mirror);

// scratch of the persistent stack-update kernel for an input side (rin, Rin): T1 (padded) | tiled core | tile partials
static int64_t fused_stack_need(int64_t rin, int64_t Rin, int64_t m, int64_t n) {
    const int64_t rp = fused_rpad(rin);
    return Rin * rp * n * 68 + n * rp * 68 + rp * ((m + 31) / 32) * 12288;
}
// scratch of the fused matvec on natural-layout vectors: T1 (padded) | two tiled vectors
static int64_t fused_matvec_need(int64_t r, int64_t R, int64_t n) {
    const int64_t rp = fused_rpad(r);
    return R * rp * n * 68 + 2 * n * rp * 68;
}

extern "C" int64_t sktt_stack_op_work(int64_t r, int64_t R, int64_t m, int64_t n, int64_t r2, int64_t R2) {
    int64_t t1 = R * r * n * r2, t2 = r * m * r2 * R2;
    int64_t need = t1 + t2;
    if (r2 == 64 && R2 == 3 && fused_stack_need(r, R, m, n) > need) need = fused_stack_need(r, R, m, n);
    if (r == 64 && R == 3 && fused_stack_need(r2, R2, m, n) > need) need = fused_stack_need(r2, R2, m, n);
    if (r2 <= 64 && R2 <= 3 && fused_matvec_need(r, R, n) > need) need = fused_matvec_need(r, R, n);
    return need;
}

// The persistent fused stack update (fused.cu) for the shapes it covers; `mirror` runs the right-stack update as the
// left-stack update of the mirrored cores.  (rin, Rin): ranks on the side of the OLD stack.
static int fused_stack(sktt_ctx* ctx, long long rin, long long Rin, long long m, long long n, const void* stack,
                       const void* x, const void* A, void* out, void* work, int mirror) {
    const long long rp = fused_rpad(rin);                    // the output side is exactly (64, 3) here
    if (sktt_stack_nat_supported(ctx, SKTT_F64, rin, Rin, m, n, 64, 3, stack, x, A)) {
        double* T1n = (double*)work;
        return sktt_stack_nat_update(ctx, m, n, (const double*)stack, (const double*)x, (const double*)A, (double*)out, T1n,
                                     T1n + Rin * rp * n * 68 + n * rp * 68, mirror);
    }
    const size_t img_bytes = (size_t)sktt_fused_image_elems(rp, Rin, m, n) * sizeof(double);
    SKTT_TRY(sktt_scratch_reserve(ctx, SKTT_SCRATCH_BULK_OFF + img_bytes));
    double* image = (double*)((char*)ctx->scratch + SKTT_SCRATCH_BULK_OFF);
    double* T1p = (double*)work;
    double* xt = T1p + Rin * rp * n * 68;
    double* part = xt + n * rp * 68;
    SKTT_TRY(sktt_fused_prepare_ex(ctx, rp, Rin, m, n, (const double*)stack, (const double*)A, nullptr, image, mirror, rin, 64,
                                   3));
    SKTT_TRY(sktt_fused_to_tiled_ex(ctx, rp, n, (const double*)x, xt, mirror, rin, 64));
    return sktt_fused_stack_update(ctx, rp, Rin, m, n, image, xt, (double*)out, T1p, part);
}

// Batched forms ("batched variants taking a leading batch dimension"): `batch` independent systems with
// identical shapes, every operand contiguous with the batch index leading; the contraction engine takes the batch as a
// grid dimension.
static inline void set_batch(GemmDesc& g, long long batch, long long sA, long long sB, long long sC) {
    g.batch = batch;
    g.sA = sA;
    g.sB = sB;
    g.sC = sC;
}

// T1[(b,c),(n,a2)] = sum_a L[a,(b,c)] X[a,(n,a2)]            (first tensordot of sle.py:217)
static int left_step1(sktt_ctx* ctx, int dtype, long long r, long long R, long long ncols, const void* Lst,
                      const void* X, int conjX, void* T1, long long batch = 1) {
    GemmDesc g = gemm_desc(R * r, ncols, r, Lst, lin_idx(1), lin_idx(R * r), X, lin_idx(ncols), lin_idx(1), T1,
                           lin_idx(ncols), lin_idx(1));
    g.conjB = conjX;
    set_batch(g, batch, r * R * r, r * ncols, R * r * ncols);
    return sktt_gemm_run(ctx, dtype, g);
}

// T2[c,m,a2,b2] = sum_{b,n} T1[b,c,n,a2] A[b,m,n,b2]           (second tensordot of sle.py:218)
static int left_step2(sktt_ctx* ctx, int dtype, long long r, long long R, long long m, long long n, long long r2,
                      long long R2, const void* T1, const void* A, void* T2, long long batch = 1) {
    GemmDesc g = gemm_desc(r * r2, m * R2, R * n, T1, two(r2, n * r2, 1), two(n, r * n * r2, r2), A,
                           two(n, m * n * R2, R2), two(R2, n * R2, 1), T2, two(r2, m * r2 * R2, R2),
                           two(R2, r2 * R2, 1));
    set_batch(g, batch, R * r * n * r2, R * m * n * R2, r * m * r2 * R2);
    return sktt_gemm_run(ctx, dtype, g);
}

static int stack_left_chain(sktt_ctx* ctx, int dtype, long long batch, int64_t r, int64_t R, int64_t m, int64_t n,
                            int64_t r2, int64_t R2, const void* Lst, const void* x, const void* A, void* out, void* work,
                            int conj_mode) {
    size_t es = dtype_size(dtype);
    char* T1 = (char*)work;
    char* T2 = T1 + (size_t)(batch * R * r * n * r2) * es;
    const int conj_col = (conj_mode == SKTT_CONJ_COL), conj_row = !conj_col;
    SKTT_TRY(left_step1(ctx, dtype, r, R, n * r2, Lst, x, conj_col, T1, batch));
    SKTT_TRY(left_step2(ctx, dtype, r, R, m, n, r2, R2, T1, A, T2, batch));
    // out[(a2,b2),c2] = sum_{(c,m)} T2[(c,m),(a2,b2)] X2[(c,m),c2]     (third tensordot, sle.py:219)
    GemmDesc g = gemm_desc(r2 * R2, r2, r * m, T2, lin_idx(1), lin_idx(r2 * R2), x, lin_idx(r2), lin_idx(1), out,
                           lin_idx(r2), lin_idx(1));
    g.conjB = conj_row;
    set_batch(g, batch, r * m * r2 * R2, r * m * r2, r2 * R2 * r2);
    return sktt_gemm_run(ctx, dtype, g);
}

extern "C" int sktt_stack_left_op(sktt_ctx* ctx, int dtype, int64_t r, int64_t R, int64_t m, int64_t n, int64_t r2,
                                  int64_t R2, const void* Lst, const void* x, const void* A, void* out, void* work,
                                  int conj_mode) {
    if (!ctx) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (m != n) return sktt_fail(ctx, SKTT_ERR_ARG, "stack_left_op: row and column mode sizes must agree");
    if (!(ctx->debug & 32) && r2 == 64 && R2 == 3 &&
        sktt_fused_supported(ctx, dtype, r, R, m, n, r2, R2))              // real: conjugation is the identity
        return fused_stack(ctx, r, R, m, n, Lst, x, A, out, work, 0);
    return stack_left_chain(ctx, dtype, 1, r, R, m, n, r2, R2, Lst, x, A, out, work, conj_mode);
}

static int stack_right_chain(sktt_ctx* ctx, int dtype, long long batch, int64_t r, int64_t R, int64_t m, int64_t n,
                             int64_t r2, int64_t R2, const void* Rst, const void* x, const void* A, void* out,
                             void* work) {
    size_t es = dtype_size(dtype);
    char* U1 = (char*)work;                                  // [c, m, a2, b2]
    char* U2 = U1 + (size_t)(batch * r * m * r2 * R2) * es;  // [n, a2, b, c]
    // U1[(c,m),(a2,b2)] = sum_c2 conj(x)[(c,m),c2] Rst[(a2,b2),c2]          (sle.py:274)
    GemmDesc g1 = gemm_desc(r * m, r2 * R2, r2, x, lin_idx(r2), lin_idx(1), Rst, lin_idx(1), lin_idx(r2), U1,
                            lin_idx(r2 * R2), lin_idx(1));
    g1.conjA = 1;
    set_batch(g1, batch, r * m * r2, r2 * R2 * r2, r * m * r2 * R2);
    SKTT_TRY(sktt_gemm_run(ctx, dtype, g1));
    // U2[n,a2,b,c] = sum_{m,b2} A[b,m,n,b2] U1[c,m,a2,b2]                   (sle.py:275)
    GemmDesc g2 = gemm_desc(R * n, r * r2, m * R2, A, two(n, m * n * R2, R2), two(R2, n * R2, 1), U1,
                            two(R2, r2 * R2, 1), two(r2, m * r2 * R2, R2), U2, two(n, r, r2 * R * r),
                            two(r2, 1, R * r));
    set_batch(g2, batch, R * m * n * R2, r * m * r2 * R2, n * r2 * R * r);
    SKTT_TRY(sktt_gemm_run(ctx, dtype, g2));
    // out[a,(b,c)] = sum_{(n,a2)} x[a,(n,a2)] U2[(n,a2),(b,c)]              (sle.py:276)
    GemmDesc g3 = gemm_desc(r, R * r, n * r2, x, lin_idx(n * r2), lin_idx(1), U2, lin_idx(R * r), lin_idx(1), out,
                            lin_idx(R * r), lin_idx(1));
    set_batch(g3, batch, r * n * r2, n * r2 * R * r, r * R * r);
    return sktt_gemm_run(ctx, dtype, g3);
}

extern "C" int sktt_stack_right_op(sktt_ctx* ctx, int dtype, int64_t r, int64_t R, int64_t m, int64_t n, int64_t r2,
                                   int64_t R2, const void* Rst, const void* x, const void* A, void* out, void* work) {
    if (!ctx) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (m != n) return sktt_fail(ctx, SKTT_ERR_ARG, "stack_right_op: row and column mode sizes must agree");
    if (!(ctx->debug & 32) && r == 64 && R == 3 && sktt_fused_supported(ctx, dtype, r2, R2, m, n, r, R))
        return fused_stack(ctx, r2, R2, m, n, Rst, x, A, out, work, 1);
    return stack_right_chain(ctx, dtype, 1, r, R, m, n, r2, R2, Rst, x, A, out, work);
}

// Batched interface-stack updates: operands [batch, ...] contiguous, work >= batch * (R r n r2 + r m r2 R2) elements.
extern "C" int sktt_batch_stack_left_op(sktt_ctx* ctx, int dtype, int64_t batch, int64_t r, int64_t R, int64_t m, int64_t n,
                                        int64_t r2, int64_t R2, const void* Lst, const void* x, const void* A, void* out,
                                        void* work, int conj_mode) {
    if (!ctx || batch < 1) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (m != n) return sktt_fail(ctx, SKTT_ERR_ARG, "batch_stack_left_op: row and column mode sizes must agree");
    return stack_left_chain(ctx, dtype, batch, r, R, m, n, r2, R2, Lst, x, A, out, work, conj_mode);
}

extern "C" int sktt_batch_stack_right_op(sktt_ctx* ctx, int dtype, int64_t batch, int64_t r, int64_t R, int64_t m,
                                         int64_t n, int64_t r2, int64_t R2, const void* Rst, const void* x, const void* A,
                                         void* out, void* work) {
    if (!ctx || batch < 1) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (m != n) return sktt_fail(ctx, SKTT_ERR_ARG, "batch_stack_right_op: row and column mode sizes must agree");
    return stack_right_chain(ctx, dtype, batch, r, R, m, n, r2, R2, Rst, x, A, out, work);
}

// ---------------------------------------------------------------------- small rhs ranks: one kernel per chain --
// Right-hand sides of the sweeps have tiny TT ranks (1 for the synthetic configurations, 4 in the cascade example): the
// two-GEMM chains below then move a few megabytes but cost two to four launches of a general GEMM each (40 us per
// call at the bench shape, 150 calls per sweep).  For p, p2 <= RHS_PMAX each chain is one elementwise / reduction kernel
// with the summation order of the reference's tensordots (over p first, then over the second index).
#define RHS_PMAX 4
#define RHS_COUNTER_OFF (SKTT_SCRATCH_COUNTER_OFF + 1024)

// f[(c,m),c2] = sum_p2 (sum_p bL[p,c] b[p,m,p2]) bR[p2,c2]
template <typename T>
__global__ void rhs_micro_small_kernel(int p, int r, int m, int p2, int r2, const T* __restrict__ bL,
                                       const T* __restrict__ b, const T* __restrict__ bR, T* __restrict__ f) {
    const long long total = (long long)r * m * r2;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c2 = (int)(idx % r2), mm = (int)((idx / r2) % m), c = (int)(idx / ((long long)r2 * m));
        T acc = Num<T>::zero();
        for (int q2 = 0; q2 < p2; ++q2) {
            T t = Num<T>::zero();
            for (int q = 0; q < p; ++q) Num<T>::fma(t, bL[q * r + c], b[((long long)q * m + mm) * p2 + q2]);
            Num<T>::fma(acc, t, bR[q2 * r2 + c2]);
        }
        f[idx] = acc;
    }
}

// out[p2,c2] = sum_{(c,m)} (sum_p bL[p,c] b[p,m,p2]) conj(x)[(c,m),c2]: CTAs own row chunks, lanes own columns; the CTA
// partials are summed in CTA order by the CTA that finishes last (deterministic whatever the arrival order).
template <typename T>
__global__ void __launch_bounds__(256)
rhs_left_small_kernel(int p, int r, int m, int p2, int r2, const T* __restrict__ bL, const T* __restrict__ b,
                      const T* __restrict__ x, T* __restrict__ out, T* part, unsigned* counter) {
    __shared__ T sh[8][RHS_PMAX][32];
    __shared__ bool last;
    const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long K = (long long)r * m, rows_per = (K + G - 1) / G;
    const long long row0 = cta * rows_per, row1 = row0 + rows_per < K ? row0 + rows_per : K;
    for (int cb = 0; cb < r2; cb += 32) {
        const int c2 = cb + lane;
        T acc[RHS_PMAX];
#pragma unroll
        for (int q2 = 0; q2 < RHS_PMAX; ++q2) acc[q2] = Num<T>::zero();
#pragma unroll 2
        for (long long row = row0 + warp; row < row1; row += 8) {
            const int c = (int)(row / m), mm = (int)(row % m);
            const T xv = c2 < r2 ? Num<T>::conj(x[row * r2 + c2]) : Num<T>::zero();
#pragma unroll
            for (int q2 = 0; q2 < RHS_PMAX; ++q2)
                if (q2 < p2) {
                    T t = Num<T>::zero();
                    for (int q = 0; q < p; ++q) Num<T>::fma(t, bL[q * r + c], b[((long long)q * m + mm) * p2 + q2]);
                    Num<T>::fma(acc[q2], t, xv);
                }
        }
#pragma unroll
        for (int q2 = 0; q2 < RHS_PMAX; ++q2) sh[warp][q2][lane] = acc[q2];
        __syncthreads();
        if (warp == 0 && c2 < r2)
            for (int q2 = 0; q2 < p2; ++q2) {
                T sacc = sh[0][q2][lane];
#pragma unroll
                for (int w = 1; w < 8; ++w) sacc = Num<T>::add(sacc, sh[w][q2][lane]);
                part[((long long)cta * p2 + q2) * r2 + c2] = sacc;
            }
        __syncthreads();
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) last = atomicAdd(counter, 1u) == (unsigned)(G - 1);
    __syncthreads();
    if (last) {
        __threadfence();
        // four threads per entry take every fourth partial (independent loads in flight), combined by two shuffles
        const int E = p2 * r2, sub = tid & 3;
        for (int e0 = 0; e0 < E; e0 += 64) {
            const int e = e0 + (tid >> 2);
            T sacc = Num<T>::zero();
            if (e < E) {
#pragma unroll 8
                for (int g = sub; g < G; g += 4) sacc = Num<T>::add(sacc, ld_cg<T>(part + (long long)g * E + e));
            }
            sacc = Num<T>::add(sacc, lane_bcast<T>(sacc, (tid & 31) ^ 1));
            sacc = Num<T>::add(sacc, lane_bcast<T>(sacc, (tid & 31) ^ 2));
            if (e < E && sub == 0) out[e] = sacc;
        }
        if (tid == 0) *counter = 0;
    }
}

// out[p,c] = sum_{(m,p2)} b[p,(m,p2)] (sum_c2 conj(x)[(c,m),c2] bR[p2,c2]): one CTA per c (a contiguous row of x)
template <typename T>
__global__ void __launch_bounds__(256)
rhs_right_small_kernel(int p, int r, int m, int p2, int r2, const T* __restrict__ bR, const T* __restrict__ b,
                       const T* __restrict__ x, T* __restrict__ out) {
    __shared__ T sh[8][RHS_PMAX];
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long len = (long long)m * r2;
    const T* xr = x + (long long)c * len;
    T acc[RHS_PMAX];
#pragma unroll
    for (int q = 0; q < RHS_PMAX; ++q) acc[q] = Num<T>::zero();
#pragma unroll 4
    for (long long e = tid; e < len; e += 256) {
        const int mm = (int)(e / r2), c2 = (int)(e % r2);
        const T xv = Num<T>::conj(xr[e]);
        for (int q2 = 0; q2 < p2; ++q2) {
            const T sv = Num<T>::mul(xv, bR[q2 * r2 + c2]);
#pragma unroll
            for (int q = 0; q < RHS_PMAX; ++q)
                if (q < p) Num<T>::fma(acc[q], b[((long long)q * m + mm) * p2 + q2], sv);
        }
    }
#pragma unroll
    for (int q = 0; q < RHS_PMAX; ++q) {
        acc[q] = warp_sum<T>(acc[q]);
        if (lane == 0) sh[warp][q] = acc[q];
    }
    __syncthreads();
    if (tid < p) {
        T sacc = sh[0][tid];
#pragma unroll
        for (int w = 1; w < 8; ++w) sacc = Num<T>::add(sacc, sh[w][tid]);
        out[(long long)tid * r + c] = sacc;
    }
}

static inline bool rhs_small(const sktt_ctx* ctx, int64_t p, int64_t p2, int64_t r, int64_t m, int64_t r2) {
    return ctx->gemm_mode != 1 && p >= 1 && p2 >= 1 && p <= RHS_PMAX && p2 <= RHS_PMAX && r * m * r2 < (1LL << 31);
}

// ---------------------------------------------------------------------------------- rhs stacks --
extern "C" int sktt_stack_left_rhs(sktt_ctx* ctx, int dtype, int64_t p, int64_t r, int64_t m, int64_t p2, int64_t r2,
                                   const void* bL, const void* b, const void* x, void* out, void* work) {
    if (!ctx) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (rhs_small(ctx, p, p2, r, m, r2)) {
        const long long K = r * m;
        int G = (int)((K + 31) / 32);
        if (G > ctx->sm_count) G = ctx->sm_count;
        SKTT_TRY(sktt_scratch_reserve(ctx, SKTT_SCRATCH_BULK_OFF + (size_t)G * p2 * r2 * dtype_size(dtype)));
        void* part = (char*)ctx->scratch + SKTT_SCRATCH_BULK_OFF;
        unsigned* counter = (unsigned*)((char*)ctx->scratch + RHS_COUNTER_OFF);
        if (dtype == SKTT_F64)
            rhs_left_small_kernel<double><<<G, 256, 0, ctx->stream>>>((int)p, (int)r, (int)m, (int)p2, (int)r2, (const double*)bL,
                                                                      (const double*)b, (const double*)x, (double*)out,
                                                                      (double*)part, counter);
        else
            rhs_left_small_kernel<cplx><<<G, 256, 0, ctx->stream>>>((int)p, (int)r, (int)m, (int)p2, (int)r2, (const cplx*)bL,
                                                                    (const cplx*)b, (const cplx*)x, (cplx*)out, (cplx*)part,
                                                                    counter);
        SKTT_LAUNCH_CHECK(ctx);
        return 0;
    }
    // Tb[c,(m,p2)] = sum_p bL[p,c] b[p,(m,p2)]                              (sle.py:246)
    GemmDesc g1 = gemm_desc(r, m * p2, p, bL, lin_idx(1), lin_idx(r), b, lin_idx(m * p2), lin_idx(1), work,
                            lin_idx(m * p2), lin_idx(1));
    SKTT_TRY(sktt_gemm_run(ctx, dtype, g1));
    // out[p2,c2] = sum_{(c,m)} Tb[(c,m),p2] conj(x)[(c,m),c2]               (sle.py:247)
    GemmDesc g2 = gemm_desc(p2, r2, r * m, work, lin_idx(1), lin_idx(p2), x, lin_idx(r2), lin_idx(1), out,
                            lin_idx(r2), lin_idx(1));
    g2.conjB = 1;
    return sktt_gemm_run(ctx, dtype, g2);
}

extern "C" int sktt_stack_right_rhs(sktt_ctx* ctx, int dtype, int64_t p, int64_t r, int64_t m, int64_t p2,
                                    int64_t r2, const void* bR, const void* b, const void* x, void* out, void* work) {
    if (!ctx) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (rhs_small(ctx, p, p2, r, m, r2)) {
        if (dtype == SKTT_F64)
            rhs_right_small_kernel<double><<<(int)r, 256, 0, ctx->stream>>>((int)p, (int)r, (int)m, (int)p2, (int)r2,
                                                                            (const double*)bR, (const double*)b,
                                                                            (const double*)x, (double*)out);
        else
            rhs_right_small_kernel<cplx><<<(int)r, 256, 0, ctx->stream>>>((int)p, (int)r, (int)m, (int)p2, (int)r2,
                                                                          (const cplx*)bR, (const cplx*)b, (const cplx*)x,
                                                                          (cplx*)out);
        SKTT_LAUNCH_CHECK(ctx);
        return 0;
    }
    // Tb[(c,m),p2] = sum_c2 conj(x)[(c,m),c2] bR[p2,c2]                     (sle.py:303)
    GemmDesc g1 = gemm_desc(r * m, p2, r2, x, lin_idx(r2), lin_idx(1), bR, lin_idx(1), lin_idx(r2), work,
                            lin_idx(p2), lin_idx(1));
    g1.conjA = 1;
    SKTT_TRY(sktt_gemm_run(ctx, dtype, g1));
    // out[p,c] = sum_{(m,p2)} b[p,(m,p2)] Tb[c,(m,p2)]                      (sle.py:304-305)
    GemmDesc g2 = gemm_desc(p, r, m * p2, b, lin_idx(m * p2), lin_idx(1), work, lin_idx(1), lin_idx(m * p2), out,
                            lin_idx(r), lin_idx(1));
    return sktt_gemm_run(ctx, dtype, g2);
}

// ---------------------------------------------------------------------------------- micro rhs ---
extern "C" int sktt_micro_rhs_als(sktt_ctx* ctx, int dtype, int64_t p, int64_t r, int64_t m, int64_t p2, int64_t r2,
                                  const void* bL, const void* b, const void* bR, void* f, void* work) {
    if (!ctx) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (rhs_small(ctx, p, p2, r, m, r2)) {
        const long long total = r * m * r2;
        long long blocks = (total + 255) / 256;
        if (blocks > 8LL * ctx->sm_count) blocks = 8LL * ctx->sm_count;
        if (dtype == SKTT_F64)
            rhs_micro_small_kernel<double><<<(int)blocks, 256, 0, ctx->stream>>>((int)p, (int)r, (int)m, (int)p2, (int)r2,
                                                                                 (const double*)bL, (const double*)b,
                                                                                 (const double*)bR, (double*)f);
        else
            rhs_micro_small_kernel<cplx><<<(int)blocks, 256, 0, ctx->stream>>>((int)p, (int)r, (int)m, (int)p2, (int)r2,
                                                                               (const cplx*)bL, (const cplx*)b,
                                                                               (const cplx*)bR, (cplx*)f);
        SKTT_LAUNCH_CHECK(ctx);
        return 0;
    }
    // Tb[c,(m,p2)] = sum_p bL[p,c] b[p,(m,p2)]                              (sle.py:424)
    GemmDesc g1 = gemm_desc(r, m * p2, p, bL, lin_idx(1), lin_idx(r), b, lin_idx(m * p2), lin_idx(1), work,
                            lin_idx(m * p2), lin_idx(1));
    SKTT_TRY(sktt_gemm_run(ctx, dtype, g1));
    // f[(c,m),c2] = sum_p2 Tb[(c,m),p2] bR[p2,c2]                           (sle.py:425)
    GemmDesc g2 = gemm_desc(r * m, r2, p2, work, lin_idx(p2), lin_idx(1), bR, lin_idx(r2), lin_idx(1), f,
                            lin_idx(r2), lin_idx(1));
    return sktt_gemm_run(ctx, dtype, g2);
}

extern "C" int sktt_micro_rhs_mals(sktt_ctx* ctx, int dtype, int64_t p, int64_t r, int64_t m, int64_t p2, int64_t m2,
                                   int64_t p3, int64_t r3, const void* bL, const void* b1, const void* b2,
                                   const void* bR, void* f, void* work) {
    if (!ctx) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    size_t es = dtype_size(dtype);
    char* Tb1 = (char*)work;                       // [c, m, p2]
    char* Tb2 = Tb1 + (size_t)(r * m * p2) * es;   // [c, m, m2, p3]
    GemmDesc g1 = gemm_desc(r, m * p2, p, bL, lin_idx(1), lin_idx(r), b1, lin_idx(m * p2), lin_idx(1), Tb1,
                            lin_idx(m * p2), lin_idx(1));                       // sle.py:464
    SKTT_TRY(sktt_gemm_run(ctx, dtype, g1));
    GemmDesc g2 = gemm_desc(r * m, m2 * p3, p2, Tb1, lin_idx(p2), lin_idx(1), b2, lin_idx(m2 * p3), lin_idx(1), Tb2,
                            lin_idx(m2 * p3), lin_idx(1));                      // sle.py:465
    SKTT_TRY(sktt_gemm_run(ctx, dtype, g2));
    GemmDesc g3 = gemm_desc(r * m * m2, r3, p3, Tb2, lin_idx(p3), lin_idx(1), bR, lin_idx(r3), lin_idx(1), f,
                            lin_idx(r3), lin_idx(1));                           // sle.py:466
    return sktt_gemm_run(ctx, dtype, g3);
}

// ---------------------------------------------------------------------------------- matvecs -----
extern "C" int sktt_micro_matvec_als(sktt_ctx* ctx, int dtype, int64_t r, int64_t R, int64_t m, int64_t n, int64_t r2,
                                     int64_t R2, const void* Lst, const void* A, const void* Rst, const void* v,
                                     void* y, void* work) {
    if (!ctx) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    size_t es = dtype_size(dtype);
    char* T1 = (char*)work;
    char* T2 = T1 + (size_t)(R * r * n * r2) * es;
    if (sktt_fused_supported(ctx, dtype, r, R, m, n, r2, R2)) {
        // stateless call: build the tile images in context scratch, then the two TMA-staged kernels (fused.cu)
        const long long rp = fused_rpad(r);
        const size_t img_bytes = (size_t)sktt_fused_image_elems(rp, R, m, n) * sizeof(double);
        SKTT_TRY(sktt_scratch_reserve(ctx, SKTT_SCRATCH_BULK_OFF + img_bytes));
        double* image = (double*)((char*)ctx->scratch + SKTT_SCRATCH_BULK_OFF);
        SKTT_TRY(sktt_fused_prepare_ex(ctx, rp, R, m, n, (const double*)Lst, (const double*)A, (const double*)Rst, image, 0, r,
                                       r2, R2));
        return sktt_fused_matvec_ex(ctx, rp, R, m, n, image, (const double*)v, (double*)y, (double*)work, r, r2);
    }
    SKTT_TRY(left_step1(ctx, dtype, r, R, n * r2, Lst, v, 0, T1));
    SKTT_TRY(left_step2(ctx, dtype, r, R, m, n, r2, R2, T1, A, T2));
    // y[(c,m),c2] = sum_{(a2,b2)} T2[(c,m),(a2,b2)] Rst[(a2,b2),c2]
    GemmDesc g = gemm_desc(r * m, r2, r2 * R2, T2, lin_idx(r2 * R2), lin_idx(1), Rst, lin_idx(r2), lin_idx(1), y,
                           lin_idx(r2), lin_idx(1));
    return sktt_gemm_run(ctx, dtype, g);
}

extern "C" int64_t sktt_micro_matvec_mals_work(int64_t r, int64_t R, int64_t m, int64_t n, int64_t R2, int64_t m2,
                                               int64_t n2, int64_t R3, int64_t r3) {
    return R * r * n * n2 * r3 + r * m * R2 * n2 * r3 + r * m * m2 * r3 * R3;
}

extern "C" int sktt_micro_matvec_mals(sktt_ctx* ctx, int dtype, int64_t r, int64_t R, int64_t m, int64_t n,
                                      int64_t R2, int64_t m2, int64_t n2, int64_t R3, int64_t r3, const void* Lst,
                                      const void* A1, const void* A2, const void* Rst, const void* v, void* y,
                                      void* work) {
    if (!ctx) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    size_t es = dtype_size(dtype);
    char* T1 = (char*)work;                                        // [b, c, n, n2, a3]
    char* T2 = T1 + (size_t)(R * r * n * n2 * r3) * es;            // [c, m, b2, n2, a3]
    char* T3 = T2 + (size_t)(r * m * R2 * n2 * r3) * es;           // [c, m, m2, a3, b3]
    SKTT_TRY(left_step1(ctx, dtype, r, R, n * n2 * r3, Lst, v, 0, T1));
    const long long q = n2 * r3;
    GemmDesc g2 = gemm_desc(r * q, m * R2, R * n, T1, two(q, n * q, 1), two(n, r * n * q, q), A1,
                            two(n, m * n * R2, R2), two(R2, n * R2, 1), T2, two(q, m * R2 * q, 1), two(R2, R2 * q, q));
    SKTT_TRY(sktt_gemm_run(ctx, dtype, g2));
    GemmDesc g3 = gemm_desc(r * m * r3, m2 * R3, R2 * n2, T2, two(r3, R2 * n2 * r3, 1), lin_idx(r3), A2,
                            two(n2, m2 * n2 * R3, R3), two(R3, n2 * R3, 1), T3, two(r3, m2 * r3 * R3, R3),
                            two(R3, r3 * R3, 1));
    SKTT_TRY(sktt_gemm_run(ctx, dtype, g3));
    GemmDesc g4 = gemm_desc(r * m * m2, r3, r3 * R3, T3, lin_idx(r3 * R3), lin_idx(1), Rst, lin_idx(r3), lin_idx(1), y,
                            lin_idx(r3), lin_idx(1));
    return sktt_gemm_run(ctx, dtype, g4);
}

// ---------------------------------------------------------------------------------- micro matrices
// Mout[(c,mu,c3),(a,nu,a3)] = sum_b L[a,b,c] W[b,mu,nu,a3,c3]; mu = (m,m2), nu = (n,n2) composite mode
// indices; W is addressed as W[b, m, n, m2, n2, a3, c3] (m2 = n2 = 1 for ALS).
template <typename T>
__global__ void micro_expand_kernel(int r, int R, int m, int n, int m2, int n2, int r3, int cc, int chunks,
                                    const T* __restrict__ Lst, const T* __restrict__ W, T* __restrict__ Mout) {
    // one CTA per (mu, nu, chunk of c3); threads sweep (c, c3, a, a3) with a3 fastest (contiguous in Mout).  The chunking
    // of c3 (cc values per CTA) bounds the shared-memory tile R * r3 * cc for large operator / solution ranks (co_oxidation
    // at r = 32: R = 21, r3 = 32)
    const int mu = blockIdx.y, nu = blockIdx.x;
    const int bz = blockIdx.z / chunks;                       // system of the batch (operands contiguous per system)
    const int c3lo = (blockIdx.z % chunks) * cc, c3n = min(cc, r3 - c3lo);
    const int mi = mu / m2, m2i = mu % m2, ni = nu / n2, n2i = nu % n2;
    const long long ncolmode = (long long)n * n2, nrowmode = (long long)m * m2;
    const long long Ncol = (long long)r * ncolmode * r3;
    const long long wstride_b = (long long)m * n * m2 * n2 * r3 * r3;
    Lst += (long long)bz * r * R * r;
    W += (long long)bz * R * wstride_b;
    Mout += (long long)bz * ((long long)r * nrowmode * r3) * Ncol;
    const long long wbase = ((((long long)mi * n + ni) * m2 + m2i) * n2 + n2i) * r3 * r3;
    extern __shared__ unsigned char smem_raw[];
    T* Ws = (T*)smem_raw;  // [R][r3 (a3)][cc (c3 - c3lo)]
    const int tile = r3 * cc;
    for (int e = threadIdx.x; e < R * r3 * c3n; e += blockDim.x) {
        int cl = e % c3n, a3 = (e / c3n) % r3, b = e / (c3n * r3);
        Ws[b * tile + a3 * cc + cl] = W[wbase + (long long)b * wstride_b + a3 * r3 + c3lo + cl];
    }
    __syncthreads();
    const long long total = (long long)r * r * r3 * c3n;
    for (long long e = threadIdx.x; e < total; e += blockDim.x) {
        int a3 = (int)(e % r3);
        long long t = e / r3;
        int a = (int)(t % r);
        t /= r;
        int cl = (int)(t % c3n);
        int c = (int)(t / c3n);
        T s = Num<T>::zero();
        for (int b = 0; b < R; ++b) Num<T>::fma(s, Lst[((long long)a * R + b) * r + c], Ws[b * tile + a3 * cc + cl]);
        long long row = ((long long)c * nrowmode + mu) * r3 + c3lo + cl;
        long long col = ((long long)a * ncolmode + nu) * r3 + a3;
        Mout[row * Ncol + col] = s;
    }
}

template <typename T>
static int launch_expand(sktt_ctx* ctx, int r, int R, int m, int n, int m2, int n2, int r3, const void* Lst,
                         const void* W, void* Mout, int batch = 1) {
    const size_t budget = 160 * 1024;
    int cc = r3;
    while (cc > 1 && (size_t)R * r3 * cc * sizeof(T) > budget) cc = (cc + 1) / 2;
    size_t smem = (size_t)R * r3 * cc * sizeof(T);
    if (smem > 200 * 1024) return sktt_fail(ctx, SKTT_ERR_ARG, "micro_matrix: R*r2 tile exceeds shared memory");
    SKTT_ONCE_PER_DEVICE(ctx);
    if (smem > 48 * 1024 && !configured) {
        SKTT_CUDA(ctx, cudaFuncSetAttribute(micro_expand_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(200 * 1024)));
        configured = true;
    }
    const int chunks = (r3 + cc - 1) / cc;
    dim3 grid(n * n2, m * m2, chunks * batch);
    long long per = (long long)r * r * r3 * cc;
    int threads = per >= 1024 ? 1024 : (per >= 256 ? 256 : 64);
    micro_expand_kernel<T><<<grid, threads, smem, ctx->stream>>>(r, R, m, n, m2, n2, r3, cc, chunks, (const T*)Lst,
                                                                  (const T*)W, (T*)Mout);
    SKTT_LAUNCH_CHECK(ctx);
    return 0;
}

// W[b,m,n,a2,c2] = sum_b2 A[(b,m,n),b2] Rst[a2,b2,c2]
static int right_fold(sktt_ctx* ctx, int dtype, long long rows, long long R2, long long r2, const void* A,
                      const void* Rst, void* W, long long batch = 1) {
    GemmDesc g = gemm_desc(rows, r2 * r2, R2, A, lin_idx(R2), lin_idx(1), Rst, lin_idx(r2), two(r2, R2 * r2, 1), W,
                           lin_idx(r2 * r2), lin_idx(1));
    set_batch(g, batch, rows * R2, r2 * R2 * r2, rows * r2 * r2);
    return sktt_gemm_run(ctx, dtype, g);
}

// Batched dense micro matrices: Lst [batch, r, R, r], A [batch, R, m, n, R2], Rst [batch, r2, R2, r2] ->
// Mout [batch, r m r2, r n r2]; work >= batch * R m n r2 r2 elements.
extern "C" int sktt_batch_micro_matrix_als(sktt_ctx* ctx, int dtype, int64_t batch, int64_t r, int64_t R, int64_t m,
                                           int64_t n, int64_t r2, int64_t R2, const void* Lst, const void* A,
                                           const void* Rst, void* Mout, void* work) {
    if (!ctx || batch < 1 || batch > 4096) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    SKTT_TRY(right_fold(ctx, dtype, R * m * n, R2, r2, A, Rst, work, batch));
    if (dtype == SKTT_F64)
        return launch_expand<double>(ctx, (int)r, (int)R, (int)m, (int)n, 1, 1, (int)r2, Lst, work, Mout, (int)batch);
    return launch_expand<cplx>(ctx, (int)r, (int)R, (int)m, (int)n, 1, 1, (int)r2, Lst, work, Mout, (int)batch);
}

extern "C" int sktt_micro_matrix_als(sktt_ctx* ctx, int dtype, int64_t r, int64_t R, int64_t m, int64_t n, int64_t r2,
                                     int64_t R2, const void* Lst, const void* A, const void* Rst, void* Mout,
                                     void* work) {
    if (!ctx) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    SKTT_TRY(right_fold(ctx, dtype, R * m * n, R2, r2, A, Rst, work));   // second tensordot of sle.py:340
    if (dtype == SKTT_F64)                                               // first tensordot + transpose, sle.py:339-345
        return launch_expand<double>(ctx, (int)r, (int)R, (int)m, (int)n, 1, 1, (int)r2, Lst, work, Mout);
    return launch_expand<cplx>(ctx, (int)r, (int)R, (int)m, (int)n, 1, 1, (int)r2, Lst, work, Mout);
}

extern "C" int64_t sktt_micro_matrix_mals_work(int64_t r, int64_t R, int64_t m, int64_t n, int64_t R2, int64_t m2,
                                               int64_t n2, int64_t R3, int64_t r3) {
    (void)r;
    return R2 * m2 * n2 * r3 * r3 + R * m * n * m2 * n2 * r3 * r3;
}

extern "C" int sktt_micro_matrix_mals(sktt_ctx* ctx, int dtype, int64_t r, int64_t R, int64_t m, int64_t n,
                                      int64_t R2, int64_t m2, int64_t n2, int64_t R3, int64_t r3, const void* Lst,
                                      const void* A1, const void* A2, const void* Rst, void* Mout, void* work) {
    if (!ctx) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    size_t es = dtype_size(dtype);
    char* W2 = (char*)work;                                          // [b2, m2, n2, a3, c3]
    char* V = W2 + (size_t)(R2 * m2 * n2 * r3 * r3) * es;            // [b, m, n, m2, n2, a3, c3]
    SKTT_TRY(right_fold(ctx, dtype, R2 * m2 * n2, R3, r3, A2, Rst, W2));          // sle.py:383
    long long ncol = m2 * n2 * r3 * r3;
    GemmDesc g = gemm_desc(R * m * n, ncol, R2, A1, lin_idx(R2), lin_idx(1), W2, lin_idx(ncol), lin_idx(1), V,
                           lin_idx(ncol), lin_idx(1));                            // sle.py:382
    SKTT_TRY(sktt_gemm_run(ctx, dtype, g));
    if (dtype == SKTT_F64)                                                        // sle.py:381, :386-388
        return launch_expand<double>(ctx, (int)r, (int)R, (int)m, (int)n, (int)m2, (int)n2, (int)r3, Lst, V, Mout);
    return launch_expand<cplx>(ctx, (int)r, (int)R, (int)m, (int)n, (int)m2, (int)n2, (int)r3, Lst, V, Mout);
}

// M += shift * t t^H  (evp.py:381)
template <typename T>
__global__ void rank1_kernel(long long N, double shift, const T* __restrict__ t, T* __restrict__ M) {
    long long total = N * N;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        long long i = e / N, j = e % N;
        T v = Num<T>::scale(Num<T>::mul(t[i], Num<T>::conj(t[j])), shift);
        M[e] = Num<T>::add(M[e], v);
    }
}

extern "C" int sktt_rank1_update(sktt_ctx* ctx, int dtype, int64_t N, double shift, const void* t, void* M) {
    if (!ctx) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    long long total = N * N;
    int blocks = (int)((total + 255) / 256 < 4LL * ctx->sm_count ? (total + 255) / 256 : 4LL * ctx->sm_count);
    if (blocks < 1) return 0;
    if (dtype == SKTT_F64) rank1_kernel<double><<<blocks, 256, 0, ctx->stream>>>(N, shift, (const double*)t, (double*)M);
    else rank1_kernel<cplx><<<blocks, 256, 0, ctx->stream>>>(N, shift, (const cplx*)t, (cplx*)M);
    SKTT_LAUNCH_CHECK(ctx);
    return 0;
}

// ---------------------------------------------------------------------------------- prepared local operators
extern "C" int64_t sktt_local_matvec_work(const sktt_local_op* op) {
    if (!op) return -1;
    if (op->sites == 1) return sktt_stack_op_work(op->r, op->R, op->m, op->n, op->r3, op->R2);
    return sktt_micro_matvec_mals_work(op->r, op->R, op->m, op->n, op->R2, op->m2, op->n2, op->R3, op->r3);
}

extern "C" int64_t sktt_local_op_image_size(sktt_ctx* ctx, int dtype, const sktt_local_op* op) {
    if (!ctx || !op) return -1;
    if (op->sites != 1) return 0;
    if (!sktt_fused_supported(ctx, dtype, op->r, op->R, op->m, op->n, op->r3, op->R2)) return 0;
    return sktt_fused_image_elems(fused_rpad(op->r), op->R, op->m, op->n);
}

extern "C" int sktt_local_op_prepare(sktt_ctx* ctx, int dtype, sktt_local_op* op, void* image) {
    if (!ctx || !op) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    op->image = nullptr;
    if (sktt_local_op_image_size(ctx, dtype, op) <= 0) return 0;      // nothing to prepare for this shape
    if (!image) return sktt_fail(ctx, SKTT_ERR_ARG, "local_op_prepare: image buffer is null");
    SKTT_TRY(sktt_fused_prepare_ex(ctx, fused_rpad(op->r), op->R, op->m, op->n, (const double*)op->Lst,
                                   (const double*)op->A1, (const double*)op->Rst, (double*)image, 0, op->r, op->r3, op->R2));
    op->image = image;
    return 0;
}

extern "C" int sktt_local_matvec(sktt_ctx* ctx, int dtype, const sktt_local_op* op, const void* v, void* y, void* work) {
    if (!ctx || !op || !v || !y || !work) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (op->sites == 1) {
        if (op->image && sktt_fused_supported(ctx, dtype, op->r, op->R, op->m, op->n, op->r3, op->R2))
            return sktt_fused_matvec_ex(ctx, fused_rpad(op->r), op->R, op->m, op->n, (const double*)op->image,
                                        (const double*)v, (double*)y, (double*)work, op->r, op->r3);
        return sktt_micro_matvec_als(ctx, dtype, op->r, op->R, op->m, op->n, op->r3, op->R2, op->Lst, op->A1, op->Rst, v,
                                     y, work);
    }
    if (op->sites == 2)
        return sktt_micro_matvec_mals(ctx, dtype, op->r, op->R, op->m, op->n, op->R2, op->m2, op->n2, op->R3, op->r3,
                                      op->Lst, op->A1, op->A2, op->Rst, v, y, work);
    return sktt_fail(ctx, SKTT_ERR_ARG, "local_matvec: sites must be 1 or 2");
}

extern "C" int64_t sktt_local_op_tiled_len(sktt_ctx* ctx, int dtype, const sktt_local_op* op) {
    if (!ctx || !op) return -1;
    if (op->sites != 1 || !op->image) return 0;
    if (!sktt_fused_supported(ctx, dtype, op->r, op->R, op->m, op->n, op->r3, op->R2)) return 0;
    return sktt_fused_tiled_len(fused_rpad(op->r), op->n);
}

extern "C" int sktt_local_matvec_tiled(sktt_ctx* ctx, int dtype, const sktt_local_op* op, const void* vt, void* yt,
                                       void* work) {
    if (!ctx || !op || !vt || !yt || !work) return SKTT_ERR_ARG;
    if (sktt_local_op_tiled_len(ctx, dtype, op) <= 0)
        return sktt_fail(ctx, SKTT_ERR_ARG, "local_matvec_tiled: operator is not prepared for the tiled path");
    return sktt_fused_matvec_tiled(ctx, fused_rpad(op->r), op->R, op->m, op->n, (const double*)op->image, (const double*)vt,
                                   (double*)yt, (double*)work);
}

int sktt_fused_pcg_persistent(sktt_ctx* ctx, long long r, long long R, long long m, long long n, const double* image,
                              const double* f, double* u, double* rv, double* p, double* s, double* w, double* T1p,
                              double tol, int max_iters, int max_cycles, int mode, int reps, double* part,
                              double* out_dev, const double* f_nat = nullptr, double* u_nat = nullptr, long long r_act = 0,
                              long long c_act = 0);

// `reps` applications yt = M vt of a prepared operator inside ONE persistent cooperative launch (the form the matvec
// takes inside the persistent CG kernel): the contraction chain without per-launch overheads, for the roofline line.
extern "C" int sktt_local_matvec_tiled_repeat(sktt_ctx* ctx, int dtype, const sktt_local_op* op, const void* vt, void* yt,
                                              void* work, int reps) {
    if (!ctx || !op || !vt || !yt || !work || reps < 1) return SKTT_ERR_ARG;
    if (sktt_local_op_tiled_len(ctx, dtype, op) <= 0)
        return sktt_fail(ctx, SKTT_ERR_ARG, "local_matvec_tiled_repeat: operator is not prepared for the tiled path");
    SKTT_TRY(sktt_scratch_reserve(ctx, SKTT_SCRATCH_BULK_OFF + (4 * 256 + 128) * sizeof(double)));
    double* part = (double*)((char*)ctx->scratch + SKTT_SCRATCH_BULK_OFF);
    return sktt_fused_pcg_persistent(ctx, fused_rpad(op->r), op->R, op->m, op->n, (const double*)op->image, (const double*)vt, nullptr,
                                     nullptr, nullptr, nullptr, (double*)yt, (double*)work, 0.0, (ctx->debug & 256) ? -1 : 0, 0, 1, reps, part,
                                     part + 4 * 256);
}

// ---------------------------------------------------------------------------------- TT algebra around the solvers
// TT.__matmul__ (scikit_tt/tensor_train.py:422-503), one core:
//   out[(p, s), m, n, (q, t)] = sum_k A[p, m, k, q] B[s, k, n, t]
// -- the operator-times-train product of the time steppers (ode.py:431-437), of tt.residual_error
// (tensor_train.py:2035-2074) and of the closeness checks (examples/co_oxidation.py:111).  One thread per output entry, the
// contraction index k (a mode size) is short; reads of B are coalesced along t.
template <typename T>
__global__ void tt_matmul_core_kernel(int P, int m, int K, int Q, int S, int n, int Tt, const T* __restrict__ A,
                                      const T* __restrict__ B, T* __restrict__ out) {
    const long long total = (long long)P * S * m * n * Q * Tt;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        long long r = e;
        const int t = (int)(r % Tt); r /= Tt;
        const int q = (int)(r % Q); r /= Q;
        const int nn = (int)(r % n); r /= n;
        const int mm = (int)(r % m); r /= m;
        const int s = (int)(r % S);
        const int p = (int)(r / S);
        T acc = Num<T>::zero();
        for (int k = 0; k < K; ++k)
            Num<T>::fma(acc, A[(((long long)p * m + mm) * K + k) * Q + q], B[(((long long)s * K + k) * n + nn) * Tt + t]);
        out[e] = acc;
    }
}

extern "C" int sktt_tt_matmul_core(sktt_ctx* ctx, int dtype, int64_t P, int64_t m, int64_t K, int64_t Q, int64_t S, int64_t n,
                                   int64_t Tt, const void* A, const void* B, void* out) {
    if (!ctx || !A || !B || !out) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    const long long total = P * S * m * n * Q * Tt;
    if (total <= 0) return 0;
    int blocks = (int)((total + 255) / 256 < 16LL * ctx->sm_count ? (total + 255) / 256 : 16LL * ctx->sm_count);
    if (dtype == SKTT_F64)
        tt_matmul_core_kernel<double><<<blocks, 256, 0, ctx->stream>>>((int)P, (int)m, (int)K, (int)Q, (int)S, (int)n, (int)Tt,
                                                                       (const double*)A, (const double*)B, (double*)out);
    else
        tt_matmul_core_kernel<cplx><<<blocks, 256, 0, ctx->stream>>>((int)P, (int)m, (int)K, (int)Q, (int)S, (int)n, (int)Tt,
                                                                     (const cplx*)A, (const cplx*)B, (cplx*)out);
    SKTT_LAUNCH_CHECK(ctx);
    return 0;
}

// ---------------------------------------------------------------------------------- alternating ridge regression
// scikit_tt/data_driven/regression.py:297-420 -- the ALS sweep skeleton with SAMPLE-indexed stacks:
//   left  : out[l, j] = sum_{a,k} L[a, j] Phi[k, j] C[a, k, l]          (regression.py:323-325)
//   right : out[a, j] = sum_{k,l} C[a, k, l] Phi[k, j] R[l, j]          (regression.py:354-356)
//   micro : M[(a, k, l), j] = L[a, j] Phi[k, j] R[l, j]                  (regression.py:388-392)
// j runs over the m samples (thousands), the other extents are ranks / basis sizes: one thread per output entry, the sample
// index fastest so that every read of L, Phi, R is coalesced.
__global__ void arr_stack_left_kernel(int r, int n, int r2, long long m, const double* __restrict__ L, const double* __restrict__ Phi,
                                      const double* __restrict__ C, double* __restrict__ out) {
    const long long total = (long long)r2 * m;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long j = e % m;
        const int l = (int)(e / m);
        double acc = 0.0;
        for (int a = 0; a < r; ++a) {
            const double la = L[a * m + j];
            double inner = 0.0;
            for (int k = 0; k < n; ++k) inner = fma(Phi[k * m + j], C[((long long)a * n + k) * r2 + l], inner);
            acc = fma(la, inner, acc);
        }
        out[e] = acc;
    }
}

__global__ void arr_stack_right_kernel(int r, int n, int r2, long long m, const double* __restrict__ C, const double* __restrict__ Phi,
                                       const double* __restrict__ R, double* __restrict__ out) {
    const long long total = (long long)r * m;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long j = e % m;
        const int a = (int)(e / m);
        double acc = 0.0;
        for (int k = 0; k < n; ++k) {
            const double pk = Phi[k * m + j];
            double inner = 0.0;
            for (int l = 0; l < r2; ++l) inner = fma(C[((long long)a * n + k) * r2 + l], R[l * m + j], inner);
            acc = fma(pk, inner, acc);
        }
        out[e] = acc;
    }
}

__global__ void arr_micro_matrix_kernel(int r, int n, int r2, long long m, const double* __restrict__ L, const double* __restrict__ Phi,
                                        const double* __restrict__ R, double* __restrict__ out) {
    const long long total = (long long)r * n * r2 * m;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long j = e % m;
        long long t = e / m;
        const int l = (int)(t % r2);
        t /= r2;
        const int k = (int)(t % n);
        const int a = (int)(t / n);
        out[e] = L[a * m + j] * Phi[k * m + j] * R[l * m + j];
    }
}

// t[i] <- t[i] / s[i] where s[i] > rcond * s[0], 0 elsewhere (the singular-value cut of lstsq(..., cond=rcond), gelss)
__global__ void pinv_scale_kernel(int k, const double* __restrict__ s, double rcond, double* __restrict__ t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) t[i] = (s[i] > rcond * s[0]) ? t[i] / s[i] : 0.0;
}

extern "C" int sktt_arr_stack(sktt_ctx* ctx, int which, int64_t r, int64_t n, int64_t r2, int64_t m, const void* stack,
                              const void* Phi, const void* core, void* out) {
    if (!ctx || !stack || !Phi || !core || !out) return SKTT_ERR_ARG;
    const long long total = (which == 0 ? r2 : r) * m;
    if (total <= 0) return 0;
    int blocks = (int)((total + 255) / 256 < 8LL * ctx->sm_count ? (total + 255) / 256 : 8LL * ctx->sm_count);
    if (which == 0)
        arr_stack_left_kernel<<<blocks, 256, 0, ctx->stream>>>((int)r, (int)n, (int)r2, m, (const double*)stack, (const double*)Phi,
                                                              (const double*)core, (double*)out);
    else
        arr_stack_right_kernel<<<blocks, 256, 0, ctx->stream>>>((int)r, (int)n, (int)r2, m, (const double*)core, (const double*)Phi,
                                                               (const double*)stack, (double*)out);
    SKTT_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" int sktt_arr_micro_matrix(sktt_ctx* ctx, int64_t r, int64_t n, int64_t r2, int64_t m, const void* L, const void* Phi,
                                     const void* R, void* out) {
    if (!ctx || !L || !Phi || !R || !out) return SKTT_ERR_ARG;
    const long long total = r * n * r2 * m;
    if (total <= 0) return 0;
    int blocks = (int)((total + 255) / 256 < 16LL * ctx->sm_count ? (total + 255) / 256 : 16LL * ctx->sm_count);
    arr_micro_matrix_kernel<<<blocks, 256, 0, ctx->stream>>>((int)r, (int)n, (int)r2, m, (const double*)L, (const double*)Phi,
                                                            (const double*)R, (double*)out);
    SKTT_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" int sktt_pinv_scale(sktt_ctx* ctx, int64_t k, const double* s, double rcond, double* t) {
    if (!ctx || !s || !t || k < 0) return SKTT_ERR_ARG;
    if (k == 0) return 0;
    pinv_scale_kernel<<<(unsigned)((k + 127) / 128), 128, 0, ctx->stream>>>((int)k, s, rcond, t);
    SKTT_LAUNCH_CHECK(ctx);
    return 0;
}
