/*
 * sktt_b200.h -- C-ABI of the B200-native ALS/MALS sweep hot path of scikit_tt.
 *
 * The reference (PGelss/scikit_tt) is pure Python and has no FFI layer of its own; the boundary
 * it offers is the Python call surface (sle.als/mals, evp.als, ode.implicit_euler, TT.ortho_*).
 * Each entry point below therefore cites the reference *Python function* (file:line under
 * /root/reference) whose arithmetic it replaces.  INTEGRATION.md shows the ctypes stubs a
 * maintainer would add to those functions.
 *
 * Conventions
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer unless the name
 *     ends in _host; tensors are dense, C-contiguous (last index fastest), exactly the layout of
 *     the reference's numpy cores.
 *   - dtype: SKTT_F64 (double) or SKTT_C128 (interleaved re,im doubles == numpy complex128).
 *   - every call is asynchronous on the context's stream unless it returns a value through a
 *     *_host pointer, in which case it synchronises that stream before returning.
 *   - return value: 0 ok; <0 argument/usage error; >0 numerical failure (singular pivot, no
 *     convergence) or CUDA runtime error.  Nothing throws across the ABI; sktt_last_error()
 *     returns a human-readable message for the last non-zero status of the context.
 *   - a context is bound to one device and one stream and is used by one host thread at a time.
 */
#ifndef SKTT_B200_H
#define SKTT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sktt_ctx sktt_ctx;

enum { SKTT_F64 = 0, SKTT_C128 = 1 };

/* conjugation mode of the op-stack updates (which copy of the solution core is conjugated) */
enum {
    SKTT_CONJ_ROW = 0, /* sle.py:217-219 / :274-276 and evp.py:323-325: conj on the row-side (m) core   */
    SKTT_CONJ_COL = 1  /* evp.py:281-283 (left stacks of evp.als): conj on the column-side (n) core     */
};

/* status codes */
enum {
    SKTT_OK = 0,
    SKTT_ERR_ARG = -1,
    SKTT_ERR_DTYPE = -2,
    SKTT_ERR_WORKSPACE = -3,
    SKTT_ERR_SINGULAR = 1,
    SKTT_ERR_NOCONV = 2,
    SKTT_ERR_CUDA = 3
};

/* ------------------------------------------------------------------ context ------------------ */
int sktt_version(void);
int sktt_ctx_create(int device, void* cuda_stream, sktt_ctx** out);
int sktt_ctx_destroy(sktt_ctx* ctx);
int sktt_ctx_set_stream(sktt_ctx* ctx, void* cuda_stream);
const char* sktt_last_error(sktt_ctx* ctx);
/* number of kernel launches issued through this context since creation (bench.py: gpu_launches) */
int64_t sktt_launch_count(sktt_ctx* ctx);
/* force a kernel family: 0 auto, 1 SIMT DFMA tiles only, 2 DMMA tensor tiles where legal */
int sktt_ctx_set_gemm_mode(sktt_ctx* ctx, int mode);
/* diagnostics (tools/ only): kernels that support it leave per-phase %globaltimer stamps in the context's scalar
 * scratch area; sktt_scratch_peek synchronises the stream and copies a piece of that area to the host           */
int sktt_ctx_set_debug(sktt_ctx* ctx, int on);
/* Deferred mode of sktt_qr_left / sktt_rq_right for tall matrices (the QR / RQ steps of the sweeps, scikit_tt/solvers/sle.py:517-525,
 * :533-541): the sketched CholeskyQR kernel runs alone, the Householder kernel that normally stands behind its failure flag is
 * not launched; a failure sets a sticky word that sktt_qr_deferred_failures reads (synchronising) and clears.  A caller that
 * switches the mode on inspects the word before it trusts the factors (the sweeps do so once per half sweep, together with
 * the outcomes of the deferred micro solves) and redoes its work with the mode off when it is set. */
int sktt_ctx_set_qr_deferred(sktt_ctx* ctx, int on);
int sktt_qr_deferred_failures(sktt_ctx* ctx, int32_t* out_host);
int sktt_scratch_peek(sktt_ctx* ctx, int64_t byte_offset, int64_t bytes, void* out_host);

/* ------------------------------------------------------------------ generic contraction ------
 * C[cm(i) + cn(j)] = alpha * sum_k opA(A[am(i) + ak(k)]) * opB(B[bk(k) + bn(j)]) + beta * C[..]
 * with two-level index maps off(i) = (i / d) * s_hi + (i % d) * s_lo (element units).  Every
 * tensordot of the reference (np.tensordot call sites listed in SURVEY.md 8c) is one such call,
 * executed without the transpose copies numpy makes.                                            */
typedef struct {
    int64_t d;
    int64_t s_hi;
    int64_t s_lo;
} sktt_idx2;

int sktt_gemm2(sktt_ctx* ctx, int dtype, int64_t M, int64_t N, int64_t K,
               const double* alpha, /* 2 doubles (re, im); host */
               const void* A, sktt_idx2 am, sktt_idx2 ak, int conjA,
               const void* B, sktt_idx2 bk, sktt_idx2 bn, int conjB,
               const double* beta, /* 2 doubles; host */
               void* C, sktt_idx2 cm, sktt_idx2 cn);

/* ------------------------------------------------------------------ interface stacks ---------
 * Shapes: Lst [r, R, r]; x [r, n, r2] (solution core, col_dims==1 squeezed); A [R, m, n, R2];
 *         out [r2, R2, r2].  m == n is required by the reference's use of one solution core on
 *         both sides.  work: device scratch of sktt_stack_op_work(...) elements of dtype.       */
int64_t sktt_stack_op_work(int64_t r, int64_t R, int64_t m, int64_t n, int64_t r2, int64_t R2);

/* sle.__construct_stack_left_op  (scikit_tt/solvers/sle.py:194-219), conj_mode SKTT_CONJ_ROW
 * evp.__construct_left_stacks    (scikit_tt/solvers/evp.py:253-288), conj_mode SKTT_CONJ_COL   */
int sktt_stack_left_op(sktt_ctx* ctx, int dtype, int64_t r, int64_t R, int64_t m, int64_t n,
                       int64_t r2, int64_t R2, const void* Lst, const void* x, const void* A,
                       void* out, void* work, int conj_mode);

/* sle.__construct_stack_right_op (scikit_tt/solvers/sle.py:250-276)
 * evp.__construct_right_stacks   (scikit_tt/solvers/evp.py:295-330)
 * Shapes: Rst [r2, R2, r2]; out [r, R, r].                                                      */
int sktt_stack_right_op(sktt_ctx* ctx, int dtype, int64_t r, int64_t R, int64_t m, int64_t n,
                        int64_t r2, int64_t R2, const void* Rst, const void* x, const void* A,
                        void* out, void* work);

/* sle.__construct_stack_left_rhs / _right_rhs (scikit_tt/solvers/sle.py:222-247, :279-305) and
 * the deflation stacks of evp (scikit_tt/solvers/evp.py:290-292, :332-334).
 * Shapes: bL [p, r]; b [p, m, p2] (rhs core); x [r, m, r2]; out_left [p2, r2];
 *         bR [p2, r2]; out_right [p, r].  work: p*m*r2 (left) / p2*m*r (right) elements.         */
int sktt_stack_left_rhs(sktt_ctx* ctx, int dtype, int64_t p, int64_t r, int64_t m, int64_t p2,
                        int64_t r2, const void* bL, const void* b, const void* x, void* out,
                        void* work);
int sktt_stack_right_rhs(sktt_ctx* ctx, int dtype, int64_t p, int64_t r, int64_t m, int64_t p2,
                         int64_t r2, const void* bR, const void* b, const void* x, void* out,
                         void* work);

/* ------------------------------------------------------------------ micro systems ------------
 * sle.__construct_micro_matrix_als (scikit_tt/solvers/sle.py:308-347), evp.py:359-365:
 *   Mout[(c,m,c2),(a,n,a2)] = sum_{b,b2} Lst[a,b,c] A[b,m,n,b2] Rst[a2,b2,c2]   (row-major N x N)
 *   work: r*r*m*n*R2 elements.                                                                  */
int sktt_micro_matrix_als(sktt_ctx* ctx, int dtype, int64_t r, int64_t R, int64_t m, int64_t n,
                          int64_t r2, int64_t R2, const void* Lst, const void* A, const void* Rst,
                          void* Mout, void* work);

/* matrix-free product with the same micro matrix (SURVEY.md row a4'):
 *   y[c,m,c2] = sum Lst[a,b,c] v[a,n,a2] A[b,m,n,b2] Rst[a2,b2,c2];  work: sktt_stack_op_work   */
int sktt_micro_matvec_als(sktt_ctx* ctx, int dtype, int64_t r, int64_t R, int64_t m, int64_t n,
                          int64_t r2, int64_t R2, const void* Lst, const void* A, const void* Rst,
                          const void* v, void* y, void* work);

/* sle.__construct_micro_matrix_mals (scikit_tt/solvers/sle.py:350-390): two-site micro matrix
 *   Mout[(c,m,m2,c3),(a,n,n2,a3)] = sum Lst[a,b,c] A1[b,m,n,b2] A2[b2,m2,n2,b3] Rst[a3,b3,c3]
 *   work: sktt_micro_matrix_mals_work elements.                                                 */
int64_t sktt_micro_matrix_mals_work(int64_t r, int64_t R, int64_t m, int64_t n, int64_t R2,
                                    int64_t m2, int64_t n2, int64_t R3, int64_t r3);
int sktt_micro_matrix_mals(sktt_ctx* ctx, int dtype, int64_t r, int64_t R, int64_t m, int64_t n,
                           int64_t R2, int64_t m2, int64_t n2, int64_t R3, int64_t r3,
                           const void* Lst, const void* A1, const void* A2, const void* Rst,
                           void* Mout, void* work);

/* two-site matrix-free product (matrix-free form); v, y: [r, n, n2, r3]       */
int64_t sktt_micro_matvec_mals_work(int64_t r, int64_t R, int64_t m, int64_t n, int64_t R2,
                                    int64_t m2, int64_t n2, int64_t R3, int64_t r3);
int sktt_micro_matvec_mals(sktt_ctx* ctx, int dtype, int64_t r, int64_t R, int64_t m, int64_t n,
                           int64_t R2, int64_t m2, int64_t n2, int64_t R3, int64_t r3,
                           const void* Lst, const void* A1, const void* A2, const void* Rst,
                           const void* v, void* y, void* work);

/* sle.__construct_micro_rhs_als (scikit_tt/solvers/sle.py:393-430), evp.py:377-379:
 *   f[c,m,c2] = sum bL[p,c] b[p,m,p2] bR[p2,c2];  work: r*m*p2 elements                          */
int sktt_micro_rhs_als(sktt_ctx* ctx, int dtype, int64_t p, int64_t r, int64_t m, int64_t p2,
                       int64_t r2, const void* bL, const void* b, const void* bR, void* f,
                       void* work);
/* sle.__construct_micro_rhs_mals (scikit_tt/solvers/sle.py:433-472):
 *   f[c,m,m2,c3] = sum bL[p,c] b1[p,m,p2] b2[p2,m2,p3] bR[p3,c3]; work: r*m*p2 + r*m*m2*p3       */
int sktt_micro_rhs_mals(sktt_ctx* ctx, int dtype, int64_t p, int64_t r, int64_t m, int64_t p2,
                        int64_t m2, int64_t p3, int64_t r3, const void* bL, const void* b1,
                        const void* b2, const void* bR, void* f, void* work);

/* evp.__construct_micro_matrices deflation term (scikit_tt/solvers/evp.py:376-381):
 *   M += shift * t t^H   (t: N vector)                                                          */
int sktt_rank1_update(sktt_ctx* ctx, int dtype, int64_t N, double shift, const void* t, void* M);

/* ------------------------------------------------------------------ local linear solves ------
 * np.linalg.solve / scipy.linalg.lu_factor+lu_solve in sle.__update_core_als/_mals
 * (scikit_tt/solvers/sle.py:505-509, :588-594): blocked right-looking LU with partial (row)
 * pivoting, row-major N x N matrix overwritten by L\U; ipiv is device int32[2 N]: the LAPACK-style
 * pivot rows (0-based) in [0, N), the accumulated row permutation in [N, 2 N);
 * info_host receives 0 or the 1-based index of an exactly-zero pivot.                           */
int sktt_lu_factor(sktt_ctx* ctx, int dtype, int64_t N, void* Mat, int32_t* ipiv, int* info_host);
int sktt_lu_solve(sktt_ctx* ctx, int dtype, int64_t N, int64_t nrhs, const void* LU,
                  const int32_t* ipiv, void* B /* [N, nrhs] row-major, overwritten */);

/* SPD fast path: lower Cholesky, row-major; info_host = 0 or index of first non-positive pivot  */
int sktt_chol_factor(sktt_ctx* ctx, int dtype, int64_t N, void* Mat, int* info_host);
int sktt_chol_solve(sktt_ctx* ctx, int dtype, int64_t N, int64_t nrhs, const void* Lfac, void* B);
/* one triangular half of the above: backward == 0 solves L Y = B, backward != 0 solves L^H X = B
 * (in place, B [N, nrhs] row-major).  Used to reduce the Hermitian-definite pencil of
 * scipy.linalg.eigh(M, b=B) in evp.__update_core (scikit_tt/solvers/evp.py:434-437) to standard form:
 * C = L^-1 M L^-H, v = L^-H y.                                                                   */
int sktt_chol_trsm(sktt_ctx* ctx, int dtype, int64_t N, int64_t nrhs, const void* Lfac, void* B,
                   int backward);

/* matrix-free Krylov solves of the ALS micro system M u = f (M as in sktt_micro_matvec_als), used
 * where the dense N x N micro matrix of sle.py:343-345 cannot exist (C3/C4).
 * u holds the initial guess on entry.  tol is relative to ||f||_2.  iters_host/relres_host are
 * written on return.  sites == 1: A2 NULL (one-site); sites == 2: two-site (MALS) operator.
 * work: sktt_krylov_work(...) elements of dtype.                                                 */
typedef struct {
    int sites;              /* 1 (ALS) or 2 (MALS) */
    int64_t r, R, m, n, R2; /* left rank, operator ranks, mode sizes of site 1                   */
    int64_t m2, n2, R3;     /* site 2 (ignored for sites == 1; then r3 is the right rank r2)      */
    int64_t r3;
    const void* Lst;
    const void* A1;
    const void* A2;
    const void* Rst;
    const void* image;      /* NULL, or the buffer filled by sktt_local_op_prepare for this (A1, Rst)        */
} sktt_local_op;

/* Prepared form of a one-site local operator.  For the shapes the TMA-staged two-kernel matvec covers (fp64,
 * R2 == 3, r3 == 64, mode sizes multiple of 32) the operator core and the right stack are re-laid once into the
 * padded shared-memory tile images that kernel copies in bulk; sktt_local_op_image_size returns the number of
 * elements of that buffer (0: shape not covered, the generic contraction chain is used and no image is needed).
 * sktt_local_op_prepare fills `image` and stores the pointer in op->image; it must be called again whenever Lst, A1
 * or Rst change.  sktt_local_matvec applies the operator (prepared or not): y = M v, work as sktt_local_matvec_work.  */
int64_t sktt_local_op_image_size(sktt_ctx* ctx, int dtype, const sktt_local_op* op);
int sktt_local_op_prepare(sktt_ctx* ctx, int dtype, sktt_local_op* op, void* image);
int64_t sktt_local_matvec_work(const sktt_local_op* op);
int sktt_local_matvec(sktt_ctx* ctx, int dtype, const sktt_local_op* op, const void* v, void* y, void* work);
/* The Krylov solvers keep their vectors in the layout the prepared matvec consumes and produces directly -- "tiled":
 * [n][a][r3 + 4] with zero padding, sktt_local_op_tiled_len elements (0 if the operator is not prepared / covered) --
 * and call this entry point once per iteration (two kernel launches).  Exposed for benchmarking that inner step.    */
int64_t sktt_local_op_tiled_len(sktt_ctx* ctx, int dtype, const sktt_local_op* op);
int sktt_local_matvec_tiled(sktt_ctx* ctx, int dtype, const sktt_local_op* op, const void* vt, void* yt, void* work);
/* reps applications of the same prepared operator inside one persistent cooperative launch (timing of the contraction
 * chain as it runs inside the persistent CG kernel; yt holds the last result)                                      */
int sktt_local_matvec_tiled_repeat(sktt_ctx* ctx, int dtype, const sktt_local_op* op, const void* vt, void* yt,
                                   void* work, int reps);

int64_t sktt_krylov_work(const sktt_local_op* op, int method, int restart);
/* method 0: CG (Hermitian positive definite), 1: restarted GMRES(restart)                       */
int sktt_krylov_solve(sktt_ctx* ctx, int dtype, const sktt_local_op* op, int method, int restart,
                      const void* f, void* u, double tol, int max_iters, void* work,
                      int* iters_host, double* relres_host);
/* The whole micro solve of sle.__update_core_als (scikit_tt/solvers/sle.py:505-509) in one call for operators the
 * fused matvec covers (f64, one site; SKTT_ERR_ARG otherwise): warm start u (dropped if worse than zero), CG, true
 * residual f - M u recomputed and CG restarted from it until |f - M u| <= tol |f| or it stops improving.
 * relres_host receives the TRUE relative residual; work as for sktt_krylov_solve (method 0).                      */
int sktt_krylov_solve_refined(sktt_ctx* ctx, int dtype, const sktt_local_op* op, const void* f, void* u,
                              double tol, int max_iters, int max_cycles, void* work, int* iters_host,
                              double* relres_host, int* cycles_host);
/* Deferred form of the same solve: one cooperative launch, NO host synchronisation.  result_dev (device memory, 4
 * doubles) receives {iterations, true relative residual, cycles, status: 0 ok / 1 iteration limit / 2 breakdown};
 * the caller queues the rest of the sweep (sle.py:65-90) behind it and inspects the outcomes once.  SKTT_ERR_ARG
 * when the operator is not covered by the persistent kernel (use sktt_krylov_solve_refined then).                    */
int sktt_krylov_solve_refined_async(sktt_ctx* ctx, int dtype, const sktt_local_op* op, const void* f, void* u,
                                    double tol, int max_cycles, void* work, double* result_dev);

/* ------------------------------------------------------------------ gauge products ------------
 * The triangular factor of the QR / RQ in sle.__update_core_als is discarded by the reference
 * (scikit_tt/solvers/sle.py:525, :541); the matrix-free micro solves use it as a warm start (the sweep's current
 * iterate expressed in the unknowns of the next micro system).  f64, one long index l < L, short extents <= 64, every
 * operand addressed as base[l * s_long + i * s_short]:
 *   factor: C(i, j)   = sum_l A(l, i) B(l, j)          (Q^H u, resp. u Q^H)
 *   push  : out(l, i) = sum_b Rm(i, b) X(l, b)         (R x_next, resp. x_prev R')                                  */
int sktt_gauge_factor(sktt_ctx* ctx, int dtype, int64_t L, int64_t ni, int64_t nj, const void* A, int64_t sla,
                      int64_t sia, const void* B, int64_t slb, int64_t sjb, void* C, int64_t sci, int64_t scj);
int sktt_gauge_push(sktt_ctx* ctx, int dtype, int64_t L, int64_t ni, int64_t nb, const void* Rm, int64_t sri,
                    int64_t srb, const void* X, int64_t slx, int64_t sbx, void* out, int64_t slo, int64_t sio);

/* ------------------------------------------------------------------ orthonormalisation -------
 * scipy.linalg.qr(mode='economic') in sle.__update_core_als (scikit_tt/solvers/sle.py:517-525):
 * Householder QR of the row-major m x n matrix A; Q (m x min(m,n), row-major) overwrites the
 * leading part of Q_out; R (min(m,n) x n) goes to R_out if non-NULL.
 * work: sktt_qr_work(m, n) elements.                                                            */
int64_t sktt_qr_work(int64_t m, int64_t n);
int sktt_qr_left(sktt_ctx* ctx, int dtype, int64_t m, int64_t n, const void* A, void* Q_out,
                 void* R_out, void* work);
/* scipy.linalg.rq(mode='economic') in sle.__update_core_als (scikit_tt/solvers/sle.py:533-541):
 * A (m x n) = R Q with Q (min(m,n) x n) having orthonormal rows.                                 */
int sktt_rq_right(sktt_ctx* ctx, int dtype, int64_t m, int64_t n, const void* A, void* Q_out,
                  void* R_out, void* work);

/* scipy.linalg.svd + the truncation rule of sle.__update_core_mals (scikit_tt/solvers/sle.py:
 * 603-614, :626-639), TT.ortho_left/ortho_right (scikit_tt/tensor_train.py:1162-1182, :1268-1288)
 * and utils.truncated_svd (scikit_tt/utils.py:111-158): one-sided Jacobi SVD of the row-major
 * m x n matrix A; U (m x k), S (k doubles), Vh (k x n), k = min(m,n), all row-major.
 * Rank rule: keep s_i / s_0 > threshold (strict; skipped when threshold == 0), then at most
 * max_rank (<= 0 means unbounded).  new_rank_host receives the kept rank.
 * work: sktt_svd_work(m, n) elements.                                                           */
int64_t sktt_svd_work(int64_t m, int64_t n);
int sktt_svd_truncate(sktt_ctx* ctx, int dtype, int64_t m, int64_t n, const void* A, void* U,
                      double* S, void* Vh, double threshold, int64_t max_rank, void* work,
                      int* new_rank_host, int* sweeps_host);

/* ------------------------------------------------------------------ local eigen solves -------
 * scipy.linalg.eigh(subset_by_index=largest k) in evp.__update_core (scikit_tt/solvers/evp.py:
 * 434-439): cyclic Jacobi on the dense Hermitian N x N micro matrix; eigenvalues ascending in
 * W (N doubles), eigenvectors in the columns of V (row-major N x N).  Mat is destroyed.          */
int sktt_eigh_jacobi(sktt_ctx* ctx, int dtype, int64_t N, void* Mat, double* W, void* V,
                     int* sweeps_host);

/* scipy.linalg.eig / scipy.sparse.linalg.eigs(sigma=...) in evp.__update_core (scikit_tt/solvers/
 * evp.py:417-432): k eigenpairs closest to sigma of the general N x N matrix Mat (optionally of
 * the pencil (Mat, Bmat)) by shift-invert Arnoldi on the LU factors of (Mat - sigma Bmat).
 * dtype is the storage type of Mat; results are always complex128: lam (k complex), vecs
 * (row-major N x k complex).  v0 == NULL starts from ones (as evp.py:418).
 * work: sktt_eig_si_work(N, k, ncv) complex128 elements.                                        */
int64_t sktt_eig_si_work(int64_t N, int64_t k, int64_t ncv);
int sktt_eig_shift_invert(sktt_ctx* ctx, int dtype, int64_t N, void* Mat, const void* Bmat,
                          double sigma, int64_t k, int64_t ncv, double tol, int max_restarts,
                          void* lam, void* vecs, void* work, int* nconv_host);

/* np.linalg.solve (sle.py:505-506) of one dense fp64 system in ONE cooperative launch (N <=
 * sktt_lu_fused_max_n()): panel factorisation in shared memory, DMMA trailing updates, the
 * right-hand side carried as one more column block, blocked back substitution.  Mat is overwritten
 * by L \ U (LAPACK layout), b by the solution (b == NULL: factorisation only), ipiv_dev [N] gets the
 * 0-based pivot rows, info_dev 0 or 1 + index of the first zero pivot.  Nothing is read back.   */
int64_t sktt_lu_fused_max_n(void);
int sktt_lu_solve_fused(sktt_ctx* ctx, int64_t N, void* Mat, void* b, int32_t* ipiv_dev, int32_t* info_dev);

/* ------------------------------------------------------------------ batched small systems ----
 * a batch of independent systems with identical shapes (BASELINE config 5: the
 * CO-pressure sweep of examples/co_oxidation.py:100-104 loops evp.als over the pressures; here the
 * loop over the systems is the grid).  Every operand is contiguous with the batch index leading:
 * Lst [batch, r, R, r], x [batch, r, n, r2], A [batch, R, m, n, R2], ...                        */

/* sle.__construct_stack_left_op / evp.__construct_left_stacks (sle.py:194-219, evp.py:253-288) for
 * `batch` systems; work: batch * (R r n r2 + r m r2 R2) elements.                               */
int sktt_batch_stack_left_op(sktt_ctx* ctx, int dtype, int64_t batch, int64_t r, int64_t R, int64_t m,
                             int64_t n, int64_t r2, int64_t R2, const void* Lst, const void* x,
                             const void* A, void* out, void* work, int conj_mode);
/* sle.__construct_stack_right_op / evp.__construct_right_stacks (sle.py:250-276, evp.py:295-330)  */
int sktt_batch_stack_right_op(sktt_ctx* ctx, int dtype, int64_t batch, int64_t r, int64_t R, int64_t m,
                              int64_t n, int64_t r2, int64_t R2, const void* Rst, const void* x,
                              const void* A, void* out, void* work);
/* evp.__construct_micro_matrices (evp.py:337-365) for `batch` systems: Mout [batch, N, N];
 * work: batch * R m n r2 r2 elements.                                                            */
int sktt_batch_micro_matrix_als(sktt_ctx* ctx, int dtype, int64_t batch, int64_t r, int64_t R, int64_t m,
                                int64_t n, int64_t r2, int64_t R2, const void* Lst, const void* A,
                                const void* Rst, void* Mout, void* work);
/* evp.__update_core, local eigen-solve (evp.py:417-432) for `batch` dense N x N micro matrices
 * (N <= 1024, k <= 8), ONE CTA per system, one launch for the batch: shift, LU with partial
 * pivoting, shift-invert Arnoldi (CGS2), projected Hessenberg eigenproblem, Ritz selection,
 * explicit restarts, Ritz vectors with LAPACK's phase convention.  lam [batch, k], vecs
 * [batch, N, k] complex128; status_dev [batch][2] int32 ON THE DEVICE: converged pairs, info
 * (bit 0 Hessenberg QR failed, bit 1 zero pivot).  Nothing is read back: the caller inspects
 * status_dev at its next synchronisation.  work: sktt_batch_eig_work(...) complex128 elements.  */
int64_t sktt_batch_eig_work(int64_t batch, int64_t N, int64_t k, int64_t ncv);
int sktt_batch_eig_shift_invert(sktt_ctx* ctx, int dtype, int64_t batch, int64_t N, const void* Mat,
                                double sigma, int64_t k, int64_t ncv, double tol, int max_restarts,
                                void* lam, void* vecs, void* work, int32_t* status_dev,
                                double* relres_dev /* [batch]: worst |h_{m+1,m} y_m| / |theta| of the returned pairs */);
/* evp.__update_core, re-orthonormalisation (evp.py:452-464, :472-487): the first `keep` left
 * singular vectors of `batch` small complex128 blocks F (P x Q), F(i, j) =
 * op(in[sys * in_stride + fi(i) + fj(j)]) with two-level index maps, written as
 * out[sys * out_stride + i * so_i + t * so_t] (optionally conjugated).  One-sided Jacobi in
 * shared memory, ONE CTA per system; null columns are completed to an orthonormal set.          */
int sktt_batch_svd_left(sktt_ctx* ctx, int64_t batch, int64_t P, int64_t Q, int64_t keep, const void* in,
                        int64_t in_stride, sktt_idx2 fi, sktt_idx2 fj, int conj_in, void* out,
                        int64_t out_stride, int64_t so_i, int64_t so_t, int conj_out);

/* ------------------------------------------------------------------ rank-sharded micro-matvec -
 * SURVEY.md 8e (BASELINE config 4): the one exchange step of the path.  y = M v of the one-site
 * local operator (sle.py:339-345 applied matrix-free) with the OUTPUT solution-rank index sharded
 * over the GPUs of one NVSwitch domain, one process per GPU.  sktt_sharded_matvec computes the
 * rows [lo, hi) and its last contraction stores them into y_local and, from the same epilogue,
 * into the peer-mapped y buffers of the other ranks (fused all-gather over NVLink);
 * sktt_peer_barrier (flag exchange in peer memory) orders those stores before the readers of y.
 * Buffers written by peers come from sktt_peer_alloc and travel as 64-byte CUDA IPC handles
 * (sktt_peer_export / sktt_peer_open); the cross-device entry points take the peer pointers where
 * an MPI-style library would take a communicator.                                               */
int sktt_peer_alloc(sktt_ctx* ctx, int64_t bytes, void** out);
int sktt_peer_free(sktt_ctx* ctx, void* ptr);
int sktt_peer_export(sktt_ctx* ctx, void* ptr, uint8_t* handle_out /* 64 bytes */);
int sktt_peer_open(sktt_ctx* ctx, const uint8_t* handle /* 64 bytes */, void** out);
int sktt_peer_close(sktt_ctx* ctx, void* ptr);
/* flags[g]: rank g's array of `world` uint64 slots (peer-mapped for g != rank); every rank calls
 * with the same increasing epoch.  timeout_flag_dev is set to 1 if a peer did not arrive.        */
int sktt_peer_barrier(sktt_ctx* ctx, int world, int rank, uint64_t epoch, void* const* flags,
                      int32_t* timeout_flag_dev);
int64_t sktt_sharded_matvec_work(int64_t r, int64_t R, int64_t m, int64_t n, int64_t r2, int64_t R2,
                                 int64_t rows);
int sktt_sharded_matvec(sktt_ctx* ctx, int dtype, int64_t r, int64_t R, int64_t m, int64_t n,
                        int64_t r2, int64_t R2, const void* Lst, const void* A, const void* Rst,
                        const void* v, int64_t lo, int64_t hi, void* y_local, int npeer,
                        void* const* y_peers, void* work);

/* ------------------------------------------------------------------ exponential integrators --
 * ode.__update_core_tdvp / __update_core_tdvp2site (scikit_tt/solvers/ode.py:1398-1614) apply
 * exp(-i h M) to a core through scipy's expm_multiply or the fixed-dimension local_krylov
 * (ode.py:1689-1757).  Here the action is a Krylov projection on the device (matvecs and
 * orthogonalisation through the contraction engine) and this entry point exponentiates the small
 * projected matrix: E = exp((c_re + i c_im) H), H and E complex128 row-major, m <= 64, one CTA,
 * scaling and squaring around a degree-18 Taylor polynomial.                                     */
int sktt_expm_small(sktt_ctx* ctx, int64_t m, const void* H, double c_re, double c_im, void* E);

/* ------------------------------------------------------------------ TT algebra around the path
 * TT.__matmul__ (scikit_tt/tensor_train.py:422-503), one core:
 *   out[(p,s), m, n, (q,t)] = sum_k A[p,m,k,q] B[s,k,n,t]
 * used by the operator-times-train products of the time steppers (ode.py:431-437) and by
 * tt.residual_error (tensor_train.py:2035-2074).                                                 */
int sktt_tt_matmul_core(sktt_ctx* ctx, int dtype, int64_t P, int64_t m, int64_t K, int64_t Q, int64_t S,
                        int64_t n, int64_t T, const void* A, const void* B, void* out);

/* ------------------------------------------------------------------ alternating ridge regression
 * scikit_tt/data_driven/regression.py:297-420 (fp64): the ALS skeleton with sample-indexed stacks.
 * which = 0: __arr_construct_stack_left  (:323-325)  out[l,j] = sum L[a,j] Phi[k,j] C[a,k,l]
 * which = 1: __arr_construct_stack_right (:354-356)  out[a,j] = sum C[a,k,l] Phi[k,j] R[l,j]
 * stack [r or r2, m], Phi [n, m] (basis functions of the mode evaluated on the m samples), core
 * [r, n, r2].  sktt_arr_micro_matrix (:388-392): out[(a,k,l), j] = L[a,j] Phi[k,j] R[l,j].
 * sktt_pinv_scale: the singular-value cut of lstsq(..., cond=rcond) (:419): t_i <- t_i / s_i where
 * s_i > rcond * s_0, else 0.                                                                     */
int sktt_arr_stack(sktt_ctx* ctx, int which, int64_t r, int64_t n, int64_t r2, int64_t m, const void* stack,
                   const void* Phi, const void* core, void* out);
int sktt_arr_micro_matrix(sktt_ctx* ctx, int64_t r, int64_t n, int64_t r2, int64_t m, const void* L,
                          const void* Phi, const void* R, void* out);
int sktt_pinv_scale(sktt_ctx* ctx, int64_t k, const double* s, double rcond, double* t);

/* ------------------------------------------------------------------ small helpers ------------ */
/* out[i] = alpha * x[i] (+ y[i] if y != NULL), n elements */
int sktt_axpby(sktt_ctx* ctx, int dtype, int64_t n, const double* alpha, const void* x,
               const double* beta, const void* y, void* out);
/* ||x||_2 and <x,y> (conjugating x), synchronising */
int sktt_nrm2(sktt_ctx* ctx, int dtype, int64_t n, const void* x, double* out_host);
int sktt_dotc(sktt_ctx* ctx, int dtype, int64_t n, const void* x, const void* y,
              double* out_host /* 2 doubles */);
/* f64 -> c128 widening and back (real part) */
int sktt_widen(sktt_ctx* ctx, int64_t n, const double* x, void* out_c128);

#ifdef __cplusplus
}
#endif
#endif /* SKTT_B200_H */
