// lu.cu -- dense local solves of the ALS/MALS micro systems (np.linalg.solve, sle.py:506;
// scipy.linalg.solve / lu_factor / lu_solve, sle.py:508-509, :589-594; the LU inside
// scipy.sparse.linalg.eigs(sigma=...), evp.py:419): blocked right-looking LU with partial row
// pivoting on the row-major micro matrix, fp64 / complex128.
//
//   panel  : cooperative multi-CTA kernel; each CTA keeps its rows of the N x 32 panel in shared
//            memory, one grid-wide barrier per column (candidates carry their row, so the pivot
//            row reaches every CTA with that single barrier).
//   update : row interchanges + unit-lower TRSM fused in one launch, then the rank-32 trailing
//            update through the DMMA contraction engine (gemm.cu).
//   solve  : gather by the accumulated permutation, then dependency-flag ("sync-free") blocked
//            triangular solves, one launch each for L and U.
#include <cooperative_groups.h>

#include "common.cuh"
#include "blas1.cuh"
namespace cg = cooperative_groups;

#define LU_NB 32
#define LU_ROWS_PER_CTA 256
#define LU_LD (LU_NB + 1)
#define LU_MAX_CTAS 1024

template <typename T>
struct PanelCand {
    double absval;
    int row;
    int pad;
    T vals[LU_NB];
};

template <typename T>
__device__ __forceinline__ double pivot_abs(T v);
template <>
__device__ __forceinline__ double pivot_abs<double>(double v) { return fabs(v); }
template <>
__device__ __forceinline__ double pivot_abs<cplx>(cplx v) { return fabs(v.re) + fabs(v.im); }  // izamax's cabs1

template <typename T>
__global__ void __launch_bounds__(256)
lu_panel_kernel(T* __restrict__ A, int N, int j0, int nb, int* __restrict__ ipiv, int* __restrict__ perm,
                PanelCand<T>* cand /* [2][G] */, T* rowj /* [2][NB] */, int* info, int cooperative) {
    extern __shared__ unsigned char smem_raw[];
    T* chunk = (T*)smem_raw;  // [LU_ROWS_PER_CTA][LU_LD]
    __shared__ T urow[LU_NB];
    __shared__ double red_val[8];
    __shared__ int red_row[8];
    __shared__ int s_piv, s_wc;
    const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row_begin = j0 + cta * LU_ROWS_PER_CTA;
    const int nrows = min(LU_ROWS_PER_CTA, N - row_begin);

    for (int e = tid; e < nrows * LU_NB; e += 256) {
        int rr = e / LU_NB, cc = e % LU_NB;
        chunk[rr * LU_LD + cc] = cc < nb ? A[(size_t)(row_begin + rr) * N + j0 + cc] : Num<T>::zero();
    }
    __syncthreads();

    for (int jj = 0; jj < nb; ++jj) {
        const int j = j0 + jj, par = jj & 1;
        // (a) local pivot candidate among rows >= j
        double best = -1.0;
        int brow = 0x7fffffff;
        for (int rr = tid; rr < nrows; rr += 256) {
            int gr = row_begin + rr;
            if (gr >= j) {
                double a = pivot_abs<T>(chunk[rr * LU_LD + jj]);
                if (a > best) { best = a; brow = gr; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ob = __shfl_xor_sync(0xffffffffu, best, o);
            int orow = __shfl_xor_sync(0xffffffffu, brow, o);
            if (ob > best || (ob == best && orow < brow)) { best = ob; brow = orow; }
        }
        if (lane == 0) { red_val[warp] = best; red_row[warp] = brow; }
        __syncthreads();
        if (warp == 0) {
            best = lane < 8 ? red_val[lane] : -1.0;
            brow = lane < 8 ? red_row[lane] : 0x7fffffff;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                double ob = __shfl_xor_sync(0xffffffffu, best, o);
                int orow = __shfl_xor_sync(0xffffffffu, brow, o);
                if (ob > best || (ob == best && orow < brow)) { best = ob; brow = orow; }
            }
            best = __shfl_sync(0xffffffffu, best, 0);
            brow = __shfl_sync(0xffffffffu, brow, 0);
            // (b) publish candidate (value, row index, row contents) and, if owned, row j itself
            PanelCand<T>* mine = cand + (size_t)par * G + cta;
            if (lane == 0) { mine->absval = best; mine->row = brow; }
            if (best >= 0.0) mine->vals[lane] = chunk[(brow - row_begin) * LU_LD + lane];
            if (j >= row_begin && j < row_begin + nrows) rowj[par * LU_NB + lane] = chunk[(j - row_begin) * LU_LD + lane];
            __threadfence();
        }
        // (c) one barrier per column
        if (cooperative) cg::this_grid().sync();
        else __syncthreads();
        // (d) global winner (ties -> smallest row index, deterministic)
        if (warp == 0) {
            double gb = -1.0;
            int grow = 0x7fffffff, gc = 0;
            for (int c = lane; c < G; c += 32) {
                const PanelCand<T>* pc = cand + (size_t)par * G + c;
                double v = __ldcg(&pc->absval);
                int rw = __ldcg(&pc->row);
                if (v > gb || (v == gb && rw < grow)) { gb = v; grow = rw; gc = c; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double ob = __shfl_xor_sync(0xffffffffu, gb, o);
                int orow = __shfl_xor_sync(0xffffffffu, grow, o);
                int oc = __shfl_xor_sync(0xffffffffu, gc, o);
                if (ob > gb || (ob == gb && orow < grow)) { gb = ob; grow = orow; gc = oc; }
            }
            if (lane == 0) { s_piv = grow; s_wc = gc; }
            gc = __shfl_sync(0xffffffffu, gc, 0);
            urow[lane] = ld_cg<T>(&cand[(size_t)par * G + gc].vals[lane]);
        }
        __syncthreads();
        const int piv = s_piv;
        const T pivot = urow[jj];
        const bool singular = (pivot_abs<T>(pivot) == 0.0);
        // (e) interchange rows j <-> piv inside the panel
        if (warp == 0) {
            if (piv != j && piv >= row_begin && piv < row_begin + nrows)
                chunk[(piv - row_begin) * LU_LD + lane] = ld_cg<T>(&rowj[par * LU_NB + lane]);
            if (j >= row_begin && j < row_begin + nrows) chunk[(j - row_begin) * LU_LD + lane] = urow[lane];
            if (cta == 0 && lane == 0) {
                ipiv[j] = piv;                                 // the accumulated permutation follows from ipiv afterwards
                if (singular && *info == 0) *info = j + 1;     // (perm_from_ipiv_kernel): no dependent global loads here
            }
        }
        __syncthreads();
        // (f) scale the column and rank-1 update of the remaining panel columns
        if (!singular) {
            // multipliers by the reciprocal of the pivot, as LAPACK's getf2 / getrf2 scale the column (one division per
            // column instead of one per row); the row's entry of column jj reaches the other lanes by a shuffle, so four
            // rows per warp are in flight without a hazard on shared memory
            const T pinv = Num<T>::div(Num<T>::one(), pivot);
            const T uj = urow[lane];
            for (int rr0 = warp; rr0 < nrows; rr0 += 32) {
                T cur[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int rr = rr0 + 8 * k;
                    cur[k] = rr < nrows ? chunk[rr * LU_LD + lane] : Num<T>::zero();
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int rr = rr0 + 8 * k;
                    const T l = Num<T>::mul(lane_bcast<T>(cur[k], jj), pinv);
                    const T upd = Num<T>::sub(cur[k], Num<T>::mul(l, uj));
                    const T nv = lane == jj ? l : (lane > jj ? upd : cur[k]);       // selects: no divergence inside the warp
                    if (rr < nrows && row_begin + rr > j) chunk[rr * LU_LD + lane] = nv;
                }
            }
        }
        __syncthreads();
    }
    for (int e = tid; e < nrows * LU_NB; e += 256) {
        int rr = e / LU_NB, cc = e % LU_NB;
        if (cc < nb) A[(size_t)(row_begin + rr) * N + j0 + cc] = chunk[rr * LU_LD + cc];
    }
}

// The same panel factorisation for panels of at most LU_CLUSTER_MAX CTAs (N - j0 <= 2048 rows: the portable cluster size), launched as ONE thread-block
// cluster: every CTA pushes its pivot candidate (value, row, row contents) and -- if it owns it -- row j straight into the
// shared memory of all CTAs of the cluster (distributed shared memory), one cluster barrier per column publishes them, and
// the winner is picked from local shared memory.  No global-memory round trip and no grid-wide barrier in the column loop
// (the cooperative kernel above pays ~5 us per column for them; a 1024 x 1024 factorisation spent 32 x 270 us there).
#define LU_CLUSTER_MAX 8

template <typename T>
__global__ void __launch_bounds__(256)
lu_panel_cluster_kernel(T* __restrict__ A, int N, int j0, int nb, int* __restrict__ ipiv, int* __restrict__ perm, int* info) {
    extern __shared__ unsigned char smem_raw[];
    T* chunk = (T*)smem_raw;  // [LU_ROWS_PER_CTA][LU_LD]
    __shared__ PanelCand<T> cands[2][LU_CLUSTER_MAX];
    __shared__ T rowj_s[2][LU_NB];
    __shared__ T urow[LU_NB];
    __shared__ double red_val[8];
    __shared__ int red_row[8];
    __shared__ int s_piv;
    cg::cluster_group cluster = cg::this_cluster();
    const int G = (int)cluster.num_blocks(), cta = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row_begin = j0 + cta * LU_ROWS_PER_CTA;
    const int nrows = max(0, min(LU_ROWS_PER_CTA, N - row_begin));

    for (int e = tid; e < nrows * LU_NB; e += 256) {
        int rr = e / LU_NB, cc = e % LU_NB;
        chunk[rr * LU_LD + cc] = cc < nb ? A[(size_t)(row_begin + rr) * N + j0 + cc] : Num<T>::zero();
    }
    cluster.sync();                                            // every CTA of the cluster is resident before the first push

    for (int jj = 0; jj < nb; ++jj) {
        const int j = j0 + jj, par = jj & 1;
        // (a) local pivot candidate among rows >= j
        double best = -1.0;
        int brow = 0x7fffffff;
        for (int rr = tid; rr < nrows; rr += 256) {
            int gr = row_begin + rr;
            if (gr >= j) {
                double a = pivot_abs<T>(chunk[rr * LU_LD + jj]);
                if (a > best) { best = a; brow = gr; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ob = __shfl_xor_sync(0xffffffffu, best, o);
            int orow = __shfl_xor_sync(0xffffffffu, brow, o);
            if (ob > best || (ob == best && orow < brow)) { best = ob; brow = orow; }
        }
        if (lane == 0) { red_val[warp] = best; red_row[warp] = brow; }
        __syncthreads();
        if (warp == 0) {
            best = lane < 8 ? red_val[lane] : -1.0;
            brow = lane < 8 ? red_row[lane] : 0x7fffffff;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                double ob = __shfl_xor_sync(0xffffffffu, best, o);
                int orow = __shfl_xor_sync(0xffffffffu, brow, o);
                if (ob > best || (ob == best && orow < brow)) { best = ob; brow = orow; }
            }
            best = __shfl_sync(0xffffffffu, best, 0);
            brow = __shfl_sync(0xffffffffu, brow, 0);
            // (b) push the candidate (and row j, if owned) into the shared memory of every CTA of the cluster
            const T cval = best >= 0.0 ? chunk[(brow - row_begin) * LU_LD + lane] : Num<T>::zero();
            const bool own_j = j >= row_begin && j < row_begin + nrows;
            const T jval = own_j ? chunk[(j - row_begin) * LU_LD + lane] : Num<T>::zero();
            for (int r = 0; r < G; ++r) {
                PanelCand<T>* dst = cluster.map_shared_rank(&cands[par][cta], r);
                if (lane == 0) { dst->absval = best; dst->row = brow; }
                dst->vals[lane] = cval;
                if (own_j) cluster.map_shared_rank(&rowj_s[par][0], r)[lane] = jval;
            }
        }
        // (c) one cluster barrier per column (release / acquire: the pushes are visible behind it)
        cluster.sync();
        // (d) winner (ties -> smallest row index, deterministic), from local shared memory
        if (warp == 0) {
            double gb = -1.0;
            int grow = 0x7fffffff, gc = 0;
            if (lane < G) { gb = cands[par][lane].absval; grow = cands[par][lane].row; gc = lane; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double ob = __shfl_xor_sync(0xffffffffu, gb, o);
                int orow = __shfl_xor_sync(0xffffffffu, grow, o);
                int oc = __shfl_xor_sync(0xffffffffu, gc, o);
                if (ob > gb || (ob == gb && orow < grow)) { gb = ob; grow = orow; gc = oc; }
            }
            if (lane == 0) s_piv = grow;
            urow[lane] = cands[par][gc].vals[lane];
        }
        __syncthreads();
        const int piv = s_piv;
        const T pivot = urow[jj];
        const bool singular = (pivot_abs<T>(pivot) == 0.0);
        // (e) interchange rows j <-> piv inside the panel
        if (warp == 0) {
            if (piv != j && piv >= row_begin && piv < row_begin + nrows) chunk[(piv - row_begin) * LU_LD + lane] = rowj_s[par][lane];
            if (j >= row_begin && j < row_begin + nrows) chunk[(j - row_begin) * LU_LD + lane] = urow[lane];
            if (cta == 0 && lane == 0) {
                ipiv[j] = piv;                                 // the accumulated permutation follows from ipiv afterwards
                if (singular && *info == 0) *info = j + 1;     // (perm_from_ipiv_kernel): no dependent global loads here
            }
        }
        __syncthreads();
        // (f) scale the column and rank-1 update of the remaining panel columns
        if (!singular) {
            const T pinv = Num<T>::div(Num<T>::one(), pivot);
            const T uj = urow[lane];
            for (int rr0 = warp; rr0 < nrows; rr0 += 32) {
                T cur[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int rr = rr0 + 8 * k;
                    cur[k] = rr < nrows ? chunk[rr * LU_LD + lane] : Num<T>::zero();
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int rr = rr0 + 8 * k;
                    const T l = Num<T>::mul(lane_bcast<T>(cur[k], jj), pinv);
                    const T upd = Num<T>::sub(cur[k], Num<T>::mul(l, uj));
                    const T nv = lane == jj ? l : (lane > jj ? upd : cur[k]);       // selects: no divergence inside the warp
                    if (rr < nrows && row_begin + rr > j) chunk[rr * LU_LD + lane] = nv;
                }
            }
        }
        __syncthreads();
    }
    for (int e = tid; e < nrows * LU_NB; e += 256) {
        int rr = e / LU_NB, cc = e % LU_NB;
        if (cc < nb) A[(size_t)(row_begin + rr) * N + j0 + cc] = chunk[rr * LU_LD + cc];
    }
    cluster.sync();                                            // no CTA leaves while another may still push into it
}

// row interchanges outside the panel + U12 = L11^{-1} A12 (thread per column).  The nb interchanges of the panel touch at
// most 2 nb rows; their net effect is worked out once per CTA on an index array (cur[q] = which of those rows ends up at
// position q), so a thread gathers its column's <= 2 nb entries with independent loads into shared memory and scatters them
// once, instead of walking through nb dependent load-load-store-store swaps in global memory.
#define LU_TRSM_THREADS 128

template <typename T>
__global__ void __launch_bounds__(LU_TRSM_THREADS)
lu_swap_trsm_kernel(T* __restrict__ A, int N, int j0, int nb, const int* __restrict__ ipiv) {
    extern __shared__ unsigned char trsm_raw[];
    T* vals = (T*)trsm_raw;                                    // [2 * LU_NB][LU_TRSM_THREADS]
    __shared__ T L11[LU_NB][LU_NB + 1];
    __shared__ int rows_s[2 * LU_NB], cur_s[2 * LU_NB], ns_s;
    const int tid = threadIdx.x;
    for (int e = tid; e < nb * nb; e += blockDim.x) L11[e / nb][e % nb] = A[(size_t)(j0 + e / nb) * N + j0 + e % nb];
    if (tid == 0) {
        int ns = nb;
        for (int k = 0; k < nb; ++k) {
            rows_s[k] = j0 + k;
            cur_s[k] = k;
        }
        for (int k = 0; k < nb; ++k) {
            const int p = ipiv[j0 + k];
            int q;
            if (p < j0 + nb) q = p - j0;                       // pivots are rows >= j0 + k
            else {
                for (q = nb; q < ns; ++q)
                    if (rows_s[q] == p) break;
                if (q == ns) {
                    rows_s[ns] = p;
                    cur_s[ns] = ns;
                    ++ns;
                }
            }
            const int t = cur_s[k];
            cur_s[k] = cur_s[q];
            cur_s[q] = t;
        }
        ns_s = ns;
    }
    __syncthreads();
    const int ns = ns_s;
    const int j1 = j0 + nb;
    const int ncols = N - nb;
    for (int t = blockIdx.x * blockDim.x + tid; t < ncols; t += gridDim.x * blockDim.x) {
        const int c = t < j0 ? t : t + nb;
        for (int q = 0; q < ns; ++q) vals[q * LU_TRSM_THREADS + tid] = A[(size_t)rows_s[q] * N + c];
        for (int q = nb; q < ns; ++q)
            if (cur_s[q] != q) A[(size_t)rows_s[q] * N + c] = vals[cur_s[q] * LU_TRSM_THREADS + tid];
        if (c >= j1) {
            T x[LU_NB];
#pragma unroll
            for (int i = 0; i < LU_NB; ++i) {
                if (i < nb) {
                    T v = vals[cur_s[i] * LU_TRSM_THREADS + tid];
#pragma unroll
                    for (int k = 0; k < LU_NB; ++k)
                        if (k < i) v = Num<T>::sub(v, Num<T>::mul(L11[i][k], x[k]));
                    x[i] = v;
                    A[(size_t)(j0 + i) * N + c] = v;
                }
            }
        } else {
            for (int i = 0; i < nb; ++i)
                if (cur_s[i] != i) A[(size_t)(j0 + i) * N + c] = vals[cur_s[i] * LU_TRSM_THREADS + tid];
        }
    }
}


// perm = the row permutation accumulated by the interchanges ipiv[0..N-1] (applied in order to the identity), built
// once after the factorisation: in shared memory while N ints fit, in global memory otherwise; one thread, N swaps.
__global__ void perm_from_ipiv_kernel(int N, const int* __restrict__ ipiv, int* __restrict__ perm, int in_smem) {
    extern __shared__ int perm_s[];
    int* pp = in_smem ? perm_s : perm;
    for (int i = threadIdx.x; i < N; i += blockDim.x) pp[i] = i;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int j = 0; j < N; ++j) {
            const int piv = ipiv[j];
            if (piv != j) {
                const int t = pp[j];
                pp[j] = pp[piv];
                pp[piv] = t;
            }
        }
    }
    __syncthreads();
    if (in_smem)
        for (int i = threadIdx.x; i < N; i += blockDim.x) perm[i] = perm_s[i];
}

template <typename T>
static int lu_factor_impl(sktt_ctx* ctx, int dtype, int N, T* A, int* ipiv, int* info_host) {
    int* perm = ipiv + N;
    const int G_max = (N + LU_ROWS_PER_CTA - 1) / LU_ROWS_PER_CTA;
    if (G_max > LU_MAX_CTAS) return sktt_fail(ctx, SKTT_ERR_ARG, "lu_factor: N too large for the panel kernel");
    size_t cand_bytes = 2 * (size_t)G_max * sizeof(PanelCand<T>) + 2 * LU_NB * sizeof(T) + 64;
    SKTT_TRY(sktt_scratch_reserve(ctx, SKTT_SCRATCH_BULK_OFF + cand_bytes));
    int* info_dev = (int*)((char*)ctx->scratch + 1024);
    PanelCand<T>* cand = (PanelCand<T>*)((char*)ctx->scratch + SKTT_SCRATCH_BULK_OFF);
    T* rowj = (T*)((char*)cand + 2 * (size_t)G_max * sizeof(PanelCand<T>));
    SKTT_CUDA(ctx, cudaMemsetAsync(info_dev, 0, sizeof(int), ctx->stream));
    const size_t smem = (size_t)LU_ROWS_PER_CTA * LU_LD * sizeof(T);
    const size_t trsm_smem = (size_t)2 * LU_NB * LU_TRSM_THREADS * sizeof(T);
    SKTT_ONCE_PER_DEVICE(ctx);
    if (!configured) {
        SKTT_CUDA(ctx, cudaFuncSetAttribute(lu_panel_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SKTT_CUDA(ctx, cudaFuncSetAttribute(lu_panel_cluster_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SKTT_CUDA(ctx, cudaFuncSetAttribute(lu_panel_cluster_kernel<T>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        SKTT_CUDA(ctx, cudaFuncSetAttribute(lu_swap_trsm_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trsm_smem));
        configured = true;
    }
    for (int j0 = 0; j0 < N; j0 += LU_NB) {
        int nb = N - j0 < LU_NB ? N - j0 : LU_NB;
        int G = (N - j0 + LU_ROWS_PER_CTA - 1) / LU_ROWS_PER_CTA;
        int coop = G > 1 ? 1 : 0;
        void* args[] = {&A, &N, &j0, &nb, &ipiv, &perm, &cand, &rowj, &info_dev, &coop};
        if (G <= LU_CLUSTER_MAX && !(ctx->debug & 64)) {       // one thread-block cluster (debug bit 64: cooperative kernel)
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(G);
            cfg.blockDim = dim3(256);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = ctx->stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = G;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            SKTT_CUDA(ctx, cudaLaunchKernelEx(&cfg, lu_panel_cluster_kernel<T>, A, N, j0, nb, ipiv, perm, info_dev));
        } else if (coop) {
            SKTT_CUDA(ctx, cudaLaunchCooperativeKernel((void*)lu_panel_kernel<T>, dim3(G), dim3(256), args, smem,
                                                       ctx->stream));
        } else {
            SKTT_CUDA(ctx, cudaLaunchKernel((void*)lu_panel_kernel<T>, dim3(1), dim3(256), args, smem, ctx->stream));
        }
        ctx->launches++;
        if (N - nb > 0) {
            int blocks = (N - nb + LU_TRSM_THREADS - 1) / LU_TRSM_THREADS;
            lu_swap_trsm_kernel<T><<<blocks, LU_TRSM_THREADS, trsm_smem, ctx->stream>>>(A, N, j0, nb, ipiv);
            SKTT_LAUNCH_CHECK(ctx);
        }
        int j1 = j0 + nb;
        if (j1 < N) {
            long long Mn = N - j1;
            GemmDesc g = gemm_desc(Mn, Mn, nb, A + (size_t)j1 * N + j0, lin_idx(N), lin_idx(1), A + (size_t)j0 * N + j1,
                                   lin_idx(N), lin_idx(1), A + (size_t)j1 * N + j1, lin_idx(N), lin_idx(1));
            g.alpha[0] = -1.0;
            g.beta[0] = 1.0;
            SKTT_TRY(sktt_gemm_run(ctx, dtype, g));
        }
    }
    {
        const int in_smem = (size_t)N * sizeof(int) <= 40 * 1024 ? 1 : 0;
        perm_from_ipiv_kernel<<<1, 256, in_smem ? (size_t)N * sizeof(int) : 0, ctx->stream>>>(N, ipiv, perm, in_smem);
        SKTT_LAUNCH_CHECK(ctx);
    }
    if (info_host) {
        SKTT_CUDA(ctx, cudaMemcpyAsync(ctx->mailbox, info_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        SKTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *info_host = *(int*)ctx->mailbox;
    }
    return 0;
}

extern "C" int sktt_lu_factor(sktt_ctx* ctx, int dtype, int64_t N, void* Mat, int32_t* ipiv, int* info_host) {
    if (!ctx || !Mat || !ipiv) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (N <= 0 || N > 0x7fffffff) return sktt_fail(ctx, SKTT_ERR_ARG, "lu_factor: bad N");
    if (dtype == SKTT_F64) return lu_factor_impl<double>(ctx, dtype, (int)N, (double*)Mat, ipiv, info_host);
    return lu_factor_impl<cplx>(ctx, dtype, (int)N, (cplx*)Mat, ipiv, info_host);
}

// ------------------------------------------------------------------------------------------------
// triangular solves with dependency flags: CTA (ticket-ordered) owns one 32-row block, consumes the
// solved blocks it depends on as their flags appear, solves its diagonal block, raises its flag.
// ------------------------------------------------------------------------------------------------
#define TRSV_TB 32

template <typename T, bool UPPER>
__global__ void __launch_bounds__(256)
trsv_flags_kernel(const T* __restrict__ LU, int N, T* __restrict__ x /* rhs in, solution out */, int xstride,
                  int* flags, unsigned* ticket) {
    __shared__ int s_blk;
    __shared__ T part[8][TRSV_TB];
    __shared__ T D[TRSV_TB][TRSV_TB + 1];
    const int nblk = (N + TRSV_TB - 1) / TRSV_TB;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_blk = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int t = s_blk;
    const int I = UPPER ? nblk - 1 - t : t;
    const int r0 = I * TRSV_TB;
    // diagonal block to smem (zero padded)
    for (int e = tid; e < TRSV_TB * TRSV_TB; e += 256) {
        int rr = e / TRSV_TB, cc = e % TRSV_TB;
        int gr = r0 + rr, gc = r0 + cc;
        D[rr][cc] = (gr < N && gc < N) ? LU[(size_t)gr * N + gc] : Num<T>::zero();
    }
    // off-diagonal blocks: lane = column inside block J, acc[rr] partial for row rr of block I
    T acc[TRSV_TB];
#pragma unroll
    for (int rr = 0; rr < TRSV_TB; ++rr) acc[rr] = Num<T>::zero();
    const int ndep = t;  // number of blocks this one depends on
    for (int q = warp; q < ndep; q += 8) {
        const int J = UPPER ? nblk - 1 - q : q;
        if (lane == 0) {
            while (atomicAdd(&flags[J], 0) == 0) __nanosleep(40);
        }
        __syncwarp();
        __threadfence();
        const int gc = J * TRSV_TB + lane;
        T xv = gc < N ? ld_cg<T>(&x[(size_t)gc * xstride]) : Num<T>::zero();
#pragma unroll
        for (int rr = 0; rr < TRSV_TB; ++rr) {
            int gr = r0 + rr;
            T a = (gr < N && gc < N) ? LU[(size_t)gr * N + gc] : Num<T>::zero();
            Num<T>::fma(acc[rr], a, xv);
        }
    }
#pragma unroll
    for (int rr = 0; rr < TRSV_TB; ++rr) {
        T s = warp_sum<T>(acc[rr]);
        if (lane == rr) part[warp][rr] = s;
    }
    __syncthreads();
    if (warp == 0) {
        const int gr = r0 + lane;
        T rhs = gr < N ? x[(size_t)gr * xstride] : Num<T>::zero();
#pragma unroll
        for (int w = 0; w < 8; ++w) rhs = Num<T>::sub(rhs, part[w][lane]);
        // diagonal solve by substitution across lanes
        if (!UPPER) {
            for (int k = 0; k < TRSV_TB; ++k) {
                T bk = lane_bcast<T>(rhs, k);  // unit diagonal: x_k = rhs_k
                if (lane > k) rhs = Num<T>::sub(rhs, Num<T>::mul(D[lane][k], bk));
            }
        } else {
            for (int k = TRSV_TB - 1; k >= 0; --k) {
                bool valid = (r0 + k) < N;
                T dk = valid ? D[k][k] : Num<T>::one();
                T xk = Num<T>::div(rhs, dk);
                T bk = lane_bcast<T>(xk, k);
                if (lane == k) rhs = bk;
                if (lane < k) rhs = Num<T>::sub(rhs, Num<T>::mul(D[lane][k], bk));
            }
        }
        if (gr < N) x[(size_t)gr * xstride] = rhs;
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicExch(&flags[I], 1);
    }
}

template <typename T>
__global__ void gather_rows_kernel(int N, int nrhs, const int* __restrict__ perm, const T* __restrict__ B,
                                   T* __restrict__ out) {
    long long total = (long long)N * nrhs;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        int i = (int)(e / nrhs), c = (int)(e % nrhs);
        out[e] = B[(size_t)perm[i] * nrhs + c];
    }
}

template <typename T>
static int trsv_run(sktt_ctx* ctx, const T* LU, int N, T* x, int xstride, bool upper) {
    const int nblk = (N + TRSV_TB - 1) / TRSV_TB;
    SKTT_TRY(sktt_scratch_reserve(ctx, SKTT_SCRATCH_BULK_OFF + (size_t)(nblk + 16) * sizeof(int)));
    int* flags = (int*)((char*)ctx->scratch + SKTT_SCRATCH_BULK_OFF);
    unsigned* ticket = (unsigned*)(flags + nblk);
    SKTT_CUDA(ctx, cudaMemsetAsync(flags, 0, (size_t)(nblk + 1) * sizeof(int), ctx->stream));
    if (upper) trsv_flags_kernel<T, true><<<nblk, 256, 0, ctx->stream>>>(LU, N, x, xstride, flags, ticket);
    else trsv_flags_kernel<T, false><<<nblk, 256, 0, ctx->stream>>>(LU, N, x, xstride, flags, ticket);
    SKTT_LAUNCH_CHECK(ctx);
    return 0;
}

template <typename T>
static int lu_solve_impl(sktt_ctx* ctx, int N, int nrhs, const T* LU, const int* ipiv, T* B) {
    const int* perm = ipiv + N;
    // gather into a temporary, then copy back (B[i] <- B[perm[i]])
    size_t bytes = (size_t)N * nrhs * sizeof(T);
    const int nblk = (N + TRSV_TB - 1) / TRSV_TB;
    size_t off = SKTT_SCRATCH_BULK_OFF + (size_t)(nblk + 16) * sizeof(int);
    off = (off + 255) / 256 * 256;
    SKTT_TRY(sktt_scratch_reserve(ctx, off + bytes));
    T* tmp = (T*)((char*)ctx->scratch + off);
    long long total = (long long)N * nrhs;
    int blocks = (int)((total + 255) / 256 < 1024 ? (total + 255) / 256 : 1024);
    gather_rows_kernel<T><<<blocks, 256, 0, ctx->stream>>>(N, nrhs, perm, B, tmp);
    SKTT_LAUNCH_CHECK(ctx);
    SKTT_CUDA(ctx, cudaMemcpyAsync(B, tmp, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    for (int c = 0; c < nrhs; ++c) {
        SKTT_TRY(trsv_run<T>(ctx, LU, N, B + c, nrhs, false));
        SKTT_TRY(trsv_run<T>(ctx, LU, N, B + c, nrhs, true));
    }
    return 0;
}

extern "C" int sktt_lu_solve(sktt_ctx* ctx, int dtype, int64_t N, int64_t nrhs, const void* LU, const int32_t* ipiv,
                             void* B) {
    if (!ctx || !LU || !ipiv || !B) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (N <= 0 || nrhs <= 0) return sktt_fail(ctx, SKTT_ERR_ARG, "lu_solve: bad extents");
    if (dtype == SKTT_F64) return lu_solve_impl<double>(ctx, (int)N, (int)nrhs, (const double*)LU, ipiv, (double*)B);
    return lu_solve_impl<cplx>(ctx, (int)N, (int)nrhs, (const cplx*)LU, ipiv, (cplx*)B);
}

// ------------------------------------------------------------------------------------------------
// Cholesky (SPD / HPD fast path): lower factor, row-major, blocked right-looking
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024)
chol_diag_kernel(T* __restrict__ A, int N, int j0, int nb, int* info) {
    // factor the nb x nb diagonal block in shared memory (one CTA, thread (i,j) per entry)
    __shared__ T D[LU_NB][LU_NB + 1];
    const int i = threadIdx.y, j = threadIdx.x;
    D[i][j] = (i < nb && j < nb) ? A[(size_t)(j0 + i) * N + j0 + j] : Num<T>::zero();
    __syncthreads();
    for (int k = 0; k < nb; ++k) {
        double dkk = Num<T>::real(D[k][k]);
        if (!(dkk > 0.0)) {
            if (i == 0 && j == 0 && *info == 0) *info = j0 + k + 1;
            dkk = 1.0;
        }
        double lkk = sqrt(dkk);
        __syncthreads();
        if (j == k && i == k) D[k][k] = Num<T>::from(lkk, 0.0);
        if (j == k && i > k && i < nb) D[i][k] = Num<T>::scale(D[i][k], 1.0 / lkk);
        __syncthreads();
        if (i > k && j > k && j <= i && i < nb) D[i][j] = Num<T>::sub(D[i][j], Num<T>::mul(D[i][k], Num<T>::conj(D[j][k])));
        __syncthreads();
    }
    if (i < nb && j < nb) A[(size_t)(j0 + i) * N + j0 + j] = (j <= i) ? D[i][j] : Num<T>::zero();
}

// L21 = A21 L11^{-H}: thread per row of A21
template <typename T>
__global__ void chol_trsm_kernel(T* __restrict__ A, int N, int j0, int nb) {
    __shared__ T L11[LU_NB][LU_NB + 1];
    for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) L11[e / nb][e % nb] = A[(size_t)(j0 + e / nb) * N + j0 + e % nb];
    __syncthreads();
    const int j1 = j0 + nb;
    for (int rr = j1 + blockIdx.x * blockDim.x + threadIdx.x; rr < N; rr += gridDim.x * blockDim.x) {
        T x[LU_NB];
#pragma unroll
        for (int c = 0; c < LU_NB; ++c) {
            if (c < nb) {
                T v = A[(size_t)rr * N + j0 + c];
#pragma unroll
                for (int k = 0; k < LU_NB; ++k)
                    if (k < c) v = Num<T>::sub(v, Num<T>::mul(x[k], Num<T>::conj(L11[c][k])));
                v = Num<T>::scale(v, 1.0 / Num<T>::real(L11[c][c]));
                x[c] = v;
                A[(size_t)rr * N + j0 + c] = v;
            }
        }
    }
}

template <typename T>
static int chol_factor_impl(sktt_ctx* ctx, int dtype, int N, T* A, int* info_host) {
    int* info_dev = (int*)((char*)ctx->scratch + 1024);
    SKTT_CUDA(ctx, cudaMemsetAsync(info_dev, 0, sizeof(int), ctx->stream));
    for (int j0 = 0; j0 < N; j0 += LU_NB) {
        int nb = N - j0 < LU_NB ? N - j0 : LU_NB;
        chol_diag_kernel<T><<<1, dim3(LU_NB, LU_NB), 0, ctx->stream>>>(A, N, j0, nb, info_dev);
        SKTT_LAUNCH_CHECK(ctx);
        int j1 = j0 + nb;
        if (j1 < N) {
            int blocks = (N - j1 + 127) / 128;
            chol_trsm_kernel<T><<<blocks, 128, 0, ctx->stream>>>(A, N, j0, nb);
            SKTT_LAUNCH_CHECK(ctx);
            long long Mn = N - j1;
            // A22 -= L21 L21^H (full square update; only the lower triangle is referenced later)
            GemmDesc g = gemm_desc(Mn, Mn, nb, A + (size_t)j1 * N + j0, lin_idx(N), lin_idx(1), A + (size_t)j1 * N + j0,
                                   lin_idx(1), lin_idx(N), A + (size_t)j1 * N + j1, lin_idx(N), lin_idx(1));
            g.conjB = 1;
            g.alpha[0] = -1.0;
            g.beta[0] = 1.0;
            SKTT_TRY(sktt_gemm_run(ctx, dtype, g));
        }
    }
    if (info_host) {
        SKTT_CUDA(ctx, cudaMemcpyAsync(ctx->mailbox, info_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        SKTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *info_host = *(int*)ctx->mailbox;
    }
    return 0;
}

extern "C" int sktt_chol_factor(sktt_ctx* ctx, int dtype, int64_t N, void* Mat, int* info_host) {
    if (!ctx || !Mat) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (N <= 0 || N > 0x7fffffff) return sktt_fail(ctx, SKTT_ERR_ARG, "chol_factor: bad N");
    if (dtype == SKTT_F64) return chol_factor_impl<double>(ctx, dtype, (int)N, (double*)Mat, info_host);
    return chol_factor_impl<cplx>(ctx, dtype, (int)N, (cplx*)Mat, info_host);
}

// L L^H x = b: forward with L (non-unit), backward with L^H.  The flag kernel handles row-major
// lower (forward, unit) and upper (backward, non-unit); Cholesky reuses a simple two-kernel blocked
// substitution instead because both of its factors have non-unit diagonals and L^H is a transpose.
template <typename T, bool BACKWARD>
__global__ void chol_subst_diag_kernel(const T* __restrict__ L, int N, int j0, int nb, T* __restrict__ x, int nrhs) {
    // one warp per rhs column: solve the nb x nb diagonal system by lane substitution
    __shared__ T D[LU_NB][LU_NB + 1];
    for (int e = threadIdx.x; e < LU_NB * LU_NB; e += blockDim.x) {
        int rr = e / LU_NB, cc = e % LU_NB;
        D[rr][cc] = (rr < nb && cc < nb) ? L[(size_t)(j0 + rr) * N + j0 + cc] : Num<T>::zero();
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int c = warp; c < nrhs; c += nw) {
        T rhs = lane < nb ? x[(size_t)(j0 + lane) * nrhs + c] : Num<T>::zero();
        if (!BACKWARD) {
            for (int k = 0; k < nb; ++k) {
                T xk = Num<T>::scale(rhs, 1.0 / Num<T>::real(D[k][k]));
                T bk = lane_bcast<T>(xk, k);
                if (lane == k) rhs = bk;
                if (lane > k) rhs = Num<T>::sub(rhs, Num<T>::mul(D[lane][k], bk));
            }
        } else {
            for (int k = nb - 1; k >= 0; --k) {
                T xk = Num<T>::scale(rhs, 1.0 / Num<T>::real(D[k][k]));
                T bk = lane_bcast<T>(xk, k);
                if (lane == k) rhs = bk;
                if (lane < k) rhs = Num<T>::sub(rhs, Num<T>::mul(Num<T>::conj(D[k][lane]), bk));
            }
        }
        if (lane < nb) x[(size_t)(j0 + lane) * nrhs + c] = rhs;
    }
}

// which: 1 forward half (L y = b), 2 backward half (L^H x = y), 3 both
template <typename T>
static int chol_solve_impl(sktt_ctx* ctx, int dtype, int N, int nrhs, const T* L, T* B, int which = 3) {
    // forward: for each block, solve diagonal then B[j1:,:] -= L[j1:, j0:j1] x_blk
    for (int j0 = 0; (which & 1) && j0 < N; j0 += LU_NB) {
        int nb = N - j0 < LU_NB ? N - j0 : LU_NB;
        chol_subst_diag_kernel<T, false><<<1, 128, 0, ctx->stream>>>(L, N, j0, nb, B, nrhs);
        SKTT_LAUNCH_CHECK(ctx);
        int j1 = j0 + nb;
        if (j1 < N) {
            GemmDesc g = gemm_desc(N - j1, nrhs, nb, L + (size_t)j1 * N + j0, lin_idx(N), lin_idx(1),
                                   B + (size_t)j0 * nrhs, lin_idx(nrhs), lin_idx(1), B + (size_t)j1 * nrhs,
                                   lin_idx(nrhs), lin_idx(1));
            g.alpha[0] = -1.0;
            g.beta[0] = 1.0;
            SKTT_TRY(sktt_gemm_run(ctx, dtype, g));
        }
    }
    // backward with L^H: for blocks from the end, solve diagonal then B[:j0,:] -= L[j0:j1, :j0]^H x_blk
    int last = ((N - 1) / LU_NB) * LU_NB;
    for (int j0 = last; (which & 2) && j0 >= 0; j0 -= LU_NB) {
        int nb = N - j0 < LU_NB ? N - j0 : LU_NB;
        chol_subst_diag_kernel<T, true><<<1, 128, 0, ctx->stream>>>(L, N, j0, nb, B, nrhs);
        SKTT_LAUNCH_CHECK(ctx);
        if (j0 > 0) {
            GemmDesc g = gemm_desc(j0, nrhs, nb, L + (size_t)j0 * N, lin_idx(1), lin_idx(N), B + (size_t)j0 * nrhs,
                                   lin_idx(nrhs), lin_idx(1), B, lin_idx(nrhs), lin_idx(1));
            g.conjA = 1;
            g.alpha[0] = -1.0;
            g.beta[0] = 1.0;
            SKTT_TRY(sktt_gemm_run(ctx, dtype, g));
        }
    }
    return 0;
}

extern "C" int sktt_chol_solve(sktt_ctx* ctx, int dtype, int64_t N, int64_t nrhs, const void* Lfac, void* B) {
    if (!ctx || !Lfac || !B) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (N <= 0 || nrhs <= 0) return sktt_fail(ctx, SKTT_ERR_ARG, "chol_solve: bad extents");
    if (dtype == SKTT_F64) return chol_solve_impl<double>(ctx, dtype, (int)N, (int)nrhs, (const double*)Lfac, (double*)B);
    return chol_solve_impl<cplx>(ctx, dtype, (int)N, (int)nrhs, (const cplx*)Lfac, (cplx*)B);
}

extern "C" int sktt_chol_trsm(sktt_ctx* ctx, int dtype, int64_t N, int64_t nrhs, const void* Lfac, void* B,
                              int backward) {
    if (!ctx || !Lfac || !B) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (N <= 0 || nrhs <= 0) return sktt_fail(ctx, SKTT_ERR_ARG, "chol_trsm: bad extents");
    const int which = backward ? 2 : 1;
    if (dtype == SKTT_F64)
        return chol_solve_impl<double>(ctx, dtype, (int)N, (int)nrhs, (const double*)Lfac, (double*)B, which);
    return chol_solve_impl<cplx>(ctx, dtype, (int)N, (int)nrhs, (const cplx*)Lfac, (cplx*)B, which);
}
