// lu_fused.cu -- np.linalg.solve of one dense fp64 micro system (sle.py:505-509: gesv = LU with partial pivoting + two
// triangular solves) as ONE cooperative launch, for the sizes of BASELINE config 1 (signaling cascade, solution rank 4:
// 1024 unknowns) and of the same-configuration leg of the bench.  Two forms live here: the DATAFLOW form further down (one
// CTA per block of 16 columns, a flag per panel instead of grid barriers, the panel in registers: 2.1 ms for 1024 unknowns)
// is the one that runs; the first form described next (3.5 ms) is kept behind ctx debug bit 10 for A/B timing.  The host-driven factorisation of lu.cu issues ~100
// launches per 1024^2 system and is bound by their latency and by one cluster barrier per column (4.0 ms + 0.34 ms for the
// solve); here the loop over the panels lives on the device:
//
//   per panel of 16 columns
//     CTA 0        : panel factorisation in shared memory ((N - j0) x 16 doubles): per column a block-wide arg-max (LAPACK's
//                    first-maximum rule), the row interchange, multipliers and the rank-1 update of the panel -- three CTA
//                    barriers per column;
//     grid barrier
//     all CTAs     : one work item per block of 8 trailing columns (the right-hand side is one more item): the rows the
//                    panel's interchanges touch are gathered once, permuted in shared memory and scattered back; the top 16
//                    rows are solved against the unit-lower triangle (U12); every remaining row block of 8 gets its rank-16
//                    update as four DMMA.8x8x4 per warp (C = A[i:i+8, J] -= L21[i:i+8, :] U12[:, J]);
//     grid barrier
//   afterwards     : the interchanges of later panels are applied to the L columns of earlier ones (net permutation per
//                    column block, staged through shared memory), and CTA 0 runs the blocked back substitution with U on the
//                    right-hand side, which has received every L update on the way as if it were a column of the matrix.
//
// Pivot sequence and arithmetic order of the elimination are those of the unblocked right-looking algorithm, so the pivots
// equal LAPACK's (tests/test_gpu_kernels.py) and the result is bit-reproducible run to run.
#include <cooperative_groups.h>

#include "common.cuh"
#include "blas1.cuh"
namespace cg = cooperative_groups;

#define LF_NB 16
#define LF_CB 8
#define LF_THREADS 512
#define LF_MAX_N 2048
#define LF_LDP (LF_NB + 1)          // pitch of the panel in shared memory

struct LfArgs {
    int N;
    double* A;        // [N][N] row-major, overwritten by L \ U
    double* b;        // [N] right-hand side, overwritten by the solution (may be null: factorisation only)
    int* ipiv;        // [N] 0-based pivot rows (row j was exchanged with row ipiv[j]), as LAPACK's getrf minus one
    int* info;        // 0, or 1 + index of the first exactly-zero pivot
};

__device__ __forceinline__ void lf_dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// element (row, col) of a column block that is either part of the matrix or the right-hand side
struct LfCols {
    double* base;
    long long ld;
    __device__ __forceinline__ double& at(int row, int col) const { return base[(long long)row * ld + col]; }
};

// First form (kept for A/B timing, ctx debug bit 10): panels by CTA 0 in shared memory, two grid barriers per panel.
__global__ void __launch_bounds__(LF_THREADS) lu_fused_kernel_v1(LfArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::grid_group grid = cg::this_grid();
    const int N = a.N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = LF_THREADS / 32;
    double* sm = reinterpret_cast<double*>(smem_raw);
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    __shared__ int s_piv[LF_NB];
    __shared__ int s_pos[2 * LF_NB], s_content[2 * LF_NB], s_np;
    const int npan = (N + LF_NB - 1) / LF_NB;

    for (int k = 0; k < npan; ++k) {
        const int j0 = k * LF_NB, nb = min(LF_NB, N - j0), rows = N - j0;
        // ------------------------------------------------------------------ panel factorisation (CTA 0)
        if (blockIdx.x == 0) {
            double* P = sm;                                   // [rows][LF_LDP]
            for (int e = tid; e < rows * nb; e += LF_THREADS) {
                const int r = e / nb, c = e - r * nb;
                P[r * LF_LDP + c] = a.A[(long long)(j0 + r) * N + j0 + c];
            }
            __syncthreads();
            for (int c = 0; c < nb; ++c) {
                double best = -1.0;
                int bi = c;
                for (int r = c + tid; r < rows; r += LF_THREADS) {
                    const double v = fabs(P[r * LF_LDP + c]);
                    if (v > best) { best = v; bi = r; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                }
                if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
                __syncthreads();
                best = s_val[lane & (LF_THREADS / 32 - 1)];   // 16 warp results, combined by a 4-step butterfly in every warp
                bi = s_idx[lane & (LF_THREADS / 32 - 1)];
#pragma unroll
                for (int o = LF_THREADS / 64; o > 0; o >>= 1) {
                    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                }
                // every thread holds the same (best, bi): LAPACK's idamax picks the first maximum
                if (tid < nb && bi != c) {                    // row interchange inside the panel
                    const double t = P[c * LF_LDP + tid];
                    P[c * LF_LDP + tid] = P[bi * LF_LDP + tid];
                    P[bi * LF_LDP + tid] = t;
                }
                if (tid == 0) {
                    s_piv[c] = j0 + bi;
                    if (!(best > 0.0) && atomicCAS(a.info, 0, j0 + c + 1) == 0) {}
                }
                __syncthreads();
                const double piv = P[c * LF_LDP + c];
                const double inv = piv != 0.0 ? 1.0 / piv : 0.0;
                for (int r = c + 1 + tid; r < rows; r += LF_THREADS) {
                    double* row = P + r * LF_LDP;
                    const double l = row[c] * inv;            // multiplier by the reciprocal of the pivot, as LAPACK's getf2
                    row[c] = l;
#pragma unroll
                    for (int cc = 0; cc < LF_NB; ++cc)
                        if (cc > c && cc < nb) row[cc] = fma(-l, P[c * LF_LDP + cc], row[cc]);
                }
                __syncthreads();
            }
            for (int e = tid; e < rows * nb; e += LF_THREADS) {
                const int r = e / nb, c = e - r * nb;
                a.A[(long long)(j0 + r) * N + j0 + c] = P[r * LF_LDP + c];
            }
            if (tid < nb) a.ipiv[j0 + tid] = s_piv[tid];
            __threadfence();
        }
        grid.sync();
        // ------------------------------------------------------------------ trailing update (all CTAs)
        const int ctrail = N - j0 - nb;                       // trailing columns
        const int nmat = (ctrail + LF_CB - 1) / LF_CB;
        const int nitems = nmat + (a.b ? 1 : 0);
        if (nitems > 0) {
            double* L11 = sm;                                 // [nb][LF_LDP] unit-lower triangle of the panel
            double* top = L11 + LF_NB * LF_LDP;               // [2 nb][LF_CB] gathered rows, then U12 in its first nb rows
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                LfCols cols;
                int cw;
                if (item < nmat) {
                    const int c0 = j0 + nb + item * LF_CB;
                    cols.base = a.A + c0;
                    cols.ld = N;
                    cw = min(LF_CB, N - c0);
                } else {
                    cols.base = a.b;
                    cols.ld = 1;
                    cw = 1;
                }
                __syncthreads();
                for (int e = tid; e < nb * nb; e += LF_THREADS) {
                    const int r = e / nb, c = e - r * nb;
                    L11[r * LF_LDP + c] = a.A[(long long)(j0 + r) * N + j0 + c];
                }
                if (tid < nb) s_piv[tid] = a.ipiv[j0 + tid];
                __syncthreads();
                if (tid == 0) {
                    // positions the interchanges touch: the nb top rows, then every pivot row not yet listed; `content`
                    // follows the rows through the sequence of interchanges
                    int np = nb;
                    for (int c = 0; c < nb; ++c) { s_pos[c] = j0 + c; s_content[c] = c; }
                    for (int c = 0; c < nb; ++c) {
                        const int p = s_piv[c];
                        int ip = -1;
                        for (int q = 0; q < np; ++q)
                            if (s_pos[q] == p) ip = q;
                        if (ip < 0) { ip = np; s_pos[np] = p; s_content[np] = np; ++np; }
                        const int t = s_content[c];
                        s_content[c] = s_content[ip];
                        s_content[ip] = t;
                    }
                    s_np = np;
                }
                __syncthreads();
                const int np = s_np;
                for (int e = tid; e < np * cw; e += LF_THREADS) {
                    const int q = e / cw, c = e - q * cw;
                    top[q * LF_CB + c] = cols.at(s_pos[s_content[q]], c);
                }
                __syncthreads();
                // U12 = L11^-1 top (unit lower): column-oriented elimination, one thread per (row, column)
                // (row-oriented inside one warp: lane = (row, column); row r needs rows < r only, so the rows are walked in
                // order with a warp barrier between them instead of a CTA barrier per column)
                if (warp == 0) {
                    for (int r = 1; r < nb; ++r) {
                        if (lane < cw) {
                            double acc = top[r * LF_CB + lane];
                            for (int c = 0; c < r; ++c) acc = fma(-L11[r * LF_LDP + c], top[c * LF_CB + lane], acc);
                            top[r * LF_CB + lane] = acc;
                        }
                        __syncwarp();
                    }
                }
                __syncthreads();
                for (int e = tid; e < np * cw; e += LF_THREADS) {
                    const int q = e / cw, c = e - q * cw;
                    cols.at(s_pos[q], c) = top[q * LF_CB + c];
                }
                __syncthreads();
                // rank-nb update of the rows below the panel: one 8-row block per warp and step
                const int fr = lane >> 2, fk = lane & 3;
                const int rbeg = j0 + nb, nblk8 = (N - rbeg + 7) / 8;
                double bfrag[4];                              // B fragments: k = 4 s + fk, n = fr (U12, shared memory)
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const int kk = 4 * s + fk;
                    bfrag[s] = (kk < nb && fr < cw) ? top[kk * LF_CB + fr] : 0.0;
                }
                const int cc0 = 2 * fk;
                for (int blk0 = warp; blk0 < nblk8; blk0 += 4 * nwarps) {
                    // four row blocks per warp and pass: every load is issued before the first store
                    double af[4][4], cv[4][2];
                    int irow[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int blk = blk0 + u * nwarps;
                        const int i = rbeg + 8 * blk + fr;
                        const bool rok = blk < nblk8 && i < N;
                        irow[u] = rok ? i : -1;
                        const double* lrow = a.A + (long long)(rok ? i : rbeg) * N + j0;
#pragma unroll
                        for (int s = 0; s < 4; ++s) {
                            const int kk = 4 * s + fk;
                            af[u][s] = (rok && kk < nb) ? lrow[kk] : 0.0;
                        }
                        cv[u][0] = (rok && cc0 < cw) ? cols.at(i, cc0) : 0.0;
                        cv[u][1] = (rok && cc0 + 1 < cw) ? cols.at(i, cc0 + 1) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
                        for (int s = 0; s < 4; ++s) lf_dmma(acc0, acc1, af[u][s], bfrag[s]);
                        if (irow[u] >= 0) {
                            if (cc0 < cw) cols.at(irow[u], cc0) = cv[u][0] - acc0;
                            if (cc0 + 1 < cw) cols.at(irow[u], cc0 + 1) = cv[u][1] - acc1;
                        }
                    }
                }
            }
            __threadfence();
        }
        grid.sync();
    }
    // ---------------------------------------------------------------------- interchanges of later panels on earlier L columns
    {
        int* perm = reinterpret_cast<int*>(sm);               // [N]
        double* stage = sm + ((N + 1) / 2 + 1);               // [N][LF_CB] (doubles; perm occupies N ints in front)
        const int ncb = (N + LF_CB - 1) / LF_CB;
        for (int cb = blockIdx.x; cb < ncb; cb += gridDim.x) {
            const int c0 = cb * LF_CB, cw = min(LF_CB, N - c0);
            const int kp = c0 / LF_NB;                        // panel these columns belong to (LF_NB is a multiple of LF_CB)
            const int rfirst = (kp + 1) * LF_NB;              // interchanges of panels > kp act on rows >= rfirst
            if (rfirst >= N) continue;
            __syncthreads();
            for (int i = rfirst + tid; i < N; i += LF_THREADS) perm[i] = i;
            __syncthreads();
            if (tid == 0)
                for (int j = rfirst; j < N; ++j) {
                    const int p = a.ipiv[j];
                    if (p != j) {
                        const int t = perm[j];
                        perm[j] = perm[p];
                        perm[p] = t;
                    }
                }
            __syncthreads();
            for (int e = tid; e < (N - rfirst) * cw; e += LF_THREADS) {
                const int i = rfirst + e / cw, c = e % cw;
                stage[(size_t)(i - rfirst) * LF_CB + c] = a.A[(long long)perm[i] * N + c0 + c];
            }
            __syncthreads();
            for (int e = tid; e < (N - rfirst) * cw; e += LF_THREADS) {
                const int i = rfirst + e / cw, c = e % cw;
                a.A[(long long)i * N + c0 + c] = stage[(size_t)(i - rfirst) * LF_CB + c];
            }
        }
    }
    if (!a.b) return;
    // ---------------------------------------------------------------------- back substitution with U (CTA 0); the L part is done
    if (blockIdx.x != 0) return;
    {
        double* x = sm;                                       // [N]
        double* D = x + N;                                    // [LF_NB][LF_LDP] diagonal block of U
        for (int i = tid; i < N; i += LF_THREADS) x[i] = a.b[i];
        __syncthreads();
        for (int k = npan - 1; k >= 0; --k) {
            const int j0 = k * LF_NB, nb = min(LF_NB, N - j0);
            for (int e = tid; e < nb * nb; e += LF_THREADS) {
                const int r = e / nb, c = e - r * nb;
                D[r * LF_LDP + c] = a.A[(long long)(j0 + r) * N + j0 + c];
            }
            __syncthreads();
            if (warp == 0) {
                double val = lane < nb ? x[j0 + lane] : 0.0;
                for (int c = nb - 1; c >= 0; --c) {
                    if (lane == c) val = val / D[c * LF_LDP + c];
                    const double xc = __shfl_sync(0xffffffffu, val, c);
                    if (lane < c) val = fma(-D[lane * LF_LDP + c], xc, val);
                }
                if (lane < nb) x[j0 + lane] = val;
            }
            __syncthreads();
            for (int i = tid; i < j0; i += LF_THREADS) {      // rows above: x_i -= U[i, block] x_block
                const double* u = a.A + (long long)i * N + j0;
                double acc = x[i];
#pragma unroll
                for (int c = 0; c < LF_NB; ++c)
                    if (c < nb) acc = fma(-u[c], x[j0 + c], acc);
                x[i] = acc;
            }
            __syncthreads();
        }
        for (int i = tid; i < N; i += LF_THREADS) a.b[i] = x[i];
    }
}


// ================================================================================================ dataflow form
// Second form: no grid barrier at all.  The matrix is cut into blocks of 16 columns, block p belongs to CTA p (the
// right-hand side is one more block with its own CTA).  CTA p applies the updates of panels 0 .. p-1 to ITS columns as soon
// as each panel is published (one flag per panel, release / acquire), then factorises its own block as panel p and
// publishes it.  The critical path per panel is therefore [flag, update of ONE block, panel factorisation] -- the updates of
// all other blocks run beside it (look-ahead for free) -- instead of [panel, grid barrier, slowest trailing item, grid
// barrier].
//
// Panel factorisation: every thread keeps its rows of the (N - j0) x 16 panel in REGISTERS (row r of the panel lives in
// thread r mod 512), so the rank-1 update of a column touches no shared memory; per column there are two CTA barriers
// (pivot candidates of the 16 warps, then the pivot row and the row it is exchanged with), the candidates of column
// c + 1 are taken from the registers the update of column c has just written.  Arithmetic and pivot rule (first maximum)
// are those of the unblocked right-looking elimination, multipliers by the reciprocal of the pivot (LAPACK's getf2).
struct LfArgs2 {
    int N;
    double* A;
    double* b;
    int* ipiv;
    int* info;
    unsigned* flags;      // [npan] panel k is published when flags[k] == epoch
    unsigned epoch;
    unsigned long long* stamps;   // optional %globaltimer stamps of the CTA of the middle block (ctx debug bit 0)
};

__device__ __forceinline__ unsigned lf_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void lf_st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

// (best, bi) <- the larger value, the smaller row on ties (idamax: first maximum)
__device__ __forceinline__ void lf_take(double& best, int& bi, double ob, int oi) {
    const bool t = ob > best || (ob == best && oi < bi);
    best = t ? ob : best;
    bi = t ? oi : bi;
}
// The same rule over a warp with three redux.sync instead of five shuffle rounds (measured: 120 cycles per round of
// 64-bit shuffles + compares): the bit pattern of a non-negative double orders like an unsigned integer, so the maximum is
// found word by word, then the smallest row among the lanes that hold it.  Lanes without a candidate pass (0, 0, INT_MAX).
__device__ __forceinline__ void lf_warp_argmax(unsigned& hi, unsigned& lo, int& idx) {
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const bool c1 = hi == mh;
    const unsigned ml = __reduce_max_sync(0xffffffffu, c1 ? lo : 0u);
    const bool c2 = c1 && lo == ml;
    idx = (int)__reduce_min_sync(0xffffffffu, c2 ? (unsigned)idx : 0x7fffffffu);
    hi = mh;
    lo = ml;
}

// stage: shared memory for 512 rows x 17 doubles -- the panel enters and leaves the registers through it, so that global
// memory sees whole 128-byte rows read by eight lanes each (every thread fetching its own row with sixteen 8-byte loads
// cost 4 us per direction: one line per lane and instruction).
// s_rows: [16 warps][16] candidate rows, s_crow: row c before the interchange.
template <int RPT>
__device__ void lf2_panel(const LfArgs2& a, int j0, int nb, double* stage, double* s_rows, double* s_crow, unsigned* s_key,
                          int* s_idx, int* s_piv) {
    const int N = a.N, rows = N - j0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double pr[RPT][LF_NB];
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
        const int r0 = u * LF_THREADS, cnt = min(LF_THREADS, rows - r0);       // rows of this chunk (block-uniform)
        __syncthreads();
        for (int e = tid; e < cnt * LF_NB; e += LF_THREADS) {
            const int rr = e >> 4, c = e & 15;
            stage[rr * LF_LDP + c] = c < nb ? __ldcg(a.A + (long long)(j0 + r0 + rr) * N + j0 + c) : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < LF_NB; ++c) pr[u][c] = tid < cnt ? stage[tid * LF_LDP + c] : 0.0;
    }
#pragma unroll
    for (int c = 0; c < LF_NB; ++c) {
        if (c < nb) {                                         // uniform
            // the exchange buffers alternate between columns: what column c + 1 publishes is not what a slow thread may
            // still be reading for column c, so ONE barrier per column is enough
            double* const x_rows = s_rows + (c & 1) * (16 * LF_NB);
            double* const x_crow = s_crow + (c & 1) * LF_NB;
            unsigned* const x_key = s_key + (c & 1) * 32;
            int* const x_idx = s_idx + (c & 1) * 16;
            double best = -1.0;
            int bi = 0x7fffffff;
#pragma unroll
            for (int u = 0; u < RPT; ++u) {
                const int r = tid + u * LF_THREADS;
                if (r >= c && r < rows) lf_take(best, bi, fabs(pr[u][c]), r);
            }
            unsigned hi = best >= 0.0 ? (unsigned)__double2hiint(best) : 0u, lo = best >= 0.0 ? (unsigned)__double2loint(best) : 0u;
            lf_warp_argmax(hi, lo, bi);
            // the warp's candidate row goes to shared memory together with its key: ONE barrier per column
#pragma unroll
            for (int u = 0; u < RPT; ++u) {
                const int r = tid + u * LF_THREADS;
                if (r == bi) {
#pragma unroll
                    for (int cc = 0; cc < LF_NB; ++cc) x_rows[warp * LF_NB + cc] = pr[u][cc];
                }
                if (r == c) {
#pragma unroll
                    for (int cc = 0; cc < LF_NB; ++cc) x_crow[cc] = pr[u][cc];
                }
            }
            if (lane == 0) {
                x_key[2 * warp] = hi;
                x_key[2 * warp + 1] = lo;
                x_idx[warp] = bi;
            }
            __syncthreads();
            hi = x_key[2 * (lane & 15)];
            lo = x_key[2 * (lane & 15) + 1];
            bi = x_idx[lane & 15];
            lf_warp_argmax(hi, lo, bi);
            // every thread holds the same key and pivot row bi; the warp that offered it: the first whose candidate is bi
            const int wwin = __ffs(__ballot_sync(0xffffffffu, lane < 16 && x_idx[lane & 15] == bi)) - 1;
            const double* prow = x_rows + wwin * LF_NB;
            if (tid == 0) {
                s_piv[c] = j0 + bi;
                if (hi == 0u && lo == 0u) atomicCAS(a.info, 0, j0 + c + 1);
            }
            const double piv = prow[c];
            const double inv = piv != 0.0 ? 1.0 / piv : 0.0;
#pragma unroll
            for (int u = 0; u < RPT; ++u) {
                const int r = tid + u * LF_THREADS;
                if (bi != c) {                                // the interchange, in registers
                    if (r == c) {
#pragma unroll
                        for (int cc = 0; cc < LF_NB; ++cc) pr[u][cc] = prow[cc];
                    } else if (r == bi) {
#pragma unroll
                        for (int cc = 0; cc < LF_NB; ++cc) pr[u][cc] = x_crow[cc];
                    }
                }
                if (r > c && r < rows) {
                    const double l = pr[u][c] * inv;          // multiplier by the reciprocal of the pivot, as LAPACK's getf2
                    pr[u][c] = l;
#pragma unroll
                    for (int cc = 0; cc < LF_NB; ++cc)
                        if (cc > c) pr[u][cc] = fma(-l, prow[cc], pr[u][cc]);
                }
            }
        }
    }
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
        const int r0 = u * LF_THREADS, cnt = min(LF_THREADS, rows - r0);
        __syncthreads();
        if (tid < cnt) {
#pragma unroll
            for (int c = 0; c < LF_NB; ++c) stage[tid * LF_LDP + c] = pr[u][c];
        }
        __syncthreads();
        for (int e = tid; e < cnt * LF_NB; e += LF_THREADS) {
            const int rr = e >> 4, c = e & 15;
            if (c < nb) a.A[(long long)(j0 + r0 + rr) * N + j0 + c] = stage[rr * LF_LDP + c];
        }
    }
    __syncthreads();
    if (tid < nb) a.ipiv[j0 + tid] = s_piv[tid];
}

__global__ void __launch_bounds__(LF_THREADS) lu_fused_kernel(LfArgs2 a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = LF_THREADS / 32;
    double* sm = reinterpret_cast<double*>(smem_raw);
    __shared__ double s_rows[2 * 16 * LF_NB], s_crow[2 * LF_NB];
    __shared__ unsigned s_key[2 * 32];
    __shared__ int s_idx[2 * 16], s_piv[LF_NB], s_pos[2 * LF_NB], s_content[2 * LF_NB], s_np;
    const int npan = (N + LF_NB - 1) / LF_NB;
    const int p = blockIdx.x;                                 // my block: columns 16 p ..; p == npan: the right-hand side
    const bool is_rhs = p == npan;
    LfCols cols;
    int cw;
    if (is_rhs) {
        cols.base = a.b;
        cols.ld = 1;
        cw = 1;
    } else {
        cols.base = a.A + (long long)p * LF_NB;
        cols.ld = N;
        cw = min(LF_NB, N - p * LF_NB);
    }
    double* L11 = sm;                                         // [nb][LF_LDP]
    double* top = L11 + LF_NB * LF_LDP;                       // [<= 2 nb][LF_NB]: rows the interchanges touch, then U12 on top

    const int nupd = is_rhs ? npan : p;
    unsigned long long* st = (a.stamps && p == npan / 2 && tid == 0) ? a.stamps : nullptr;
    auto tick = [&](int i) {
        if (st) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(st[i]));
    };
    for (int k = 0; k < nupd; ++k) {
        const int j0 = k * LF_NB, nb = min(LF_NB, N - j0);
        if (tid == 0)
            while (lf_ld_acquire(a.flags + k) != a.epoch) {}
        __syncthreads();
        if (k == nupd - 1) tick(0);
        // ---- unit-lower triangle and pivots of panel k
        for (int e = tid; e < nb * nb; e += LF_THREADS) {
            const int r = e / nb, c = e - r * nb;
            L11[r * LF_LDP + c] = __ldcg(a.A + (long long)(j0 + r) * N + j0 + c);
        }
        if (warp == 0) {
            // rows the interchanges touch: the nb top rows, then every pivot row not yet listed (slot = lane); `content`
            // follows the rows through the sequence of interchanges
            int pos = lane < nb ? j0 + lane : -1, content = lane, np = nb;
            const int mypiv = lane < nb ? __ldcg(a.ipiv + j0 + lane) : -1;
            for (int c = 0; c < nb; ++c) {
                const int pv = __shfl_sync(0xffffffffu, mypiv, c);
                const unsigned found = __ballot_sync(0xffffffffu, pos == pv);
                int ip;
                if (found) {
                    ip = __ffs(found) - 1;
                } else {
                    ip = np;
                    if (lane == np) pos = pv;
                    ++np;
                }
                const int ca = __shfl_sync(0xffffffffu, content, c), cb = __shfl_sync(0xffffffffu, content, ip);
                if (lane == c) content = cb;
                if (lane == ip) content = ca;
            }
            s_pos[lane] = pos;
            s_content[lane] = content;
            if (lane == 0) s_np = np;
        }
        __syncthreads();
        if (k == nupd - 1) tick(1);
        const int np = s_np;
        for (int e = tid; e < np * cw; e += LF_THREADS) {
            const int q = e / cw, c = e - q * cw;
            top[q * LF_NB + c] = __ldcg(&cols.at(s_pos[s_content[q]], c));
        }
        __syncthreads();
        if (k == nupd - 1) tick(2);
        // ---- U12 = L11^-1 top: warp w takes column w, lane r holds row r; right-looking elimination through shuffles
        if (warp < cw) {
            double v = lane < nb ? top[lane * LF_NB + warp] : 0.0;
            double lrow[LF_NB];                               // my row of the unit-lower triangle (zero from the diagonal on)
#pragma unroll
            for (int c = 0; c < LF_NB; ++c) lrow[c] = (lane < nb && c < lane) ? L11[lane * LF_LDP + c] : 0.0;
#pragma unroll
            for (int c = 0; c + 1 < LF_NB; ++c) {
                const double vc = __shfl_sync(0xffffffffu, v, c);
                v = fma(-lrow[c], vc, v);
            }
            if (lane < nb) top[lane * LF_NB + warp] = v;
        }
        __syncthreads();
        if (k == nupd - 1) tick(3);
        for (int e = tid; e < np * cw; e += LF_THREADS) {
            const int q = e / cw, c = e - q * cw;
            cols.at(s_pos[q], c) = top[q * LF_NB + c];
        }
        __syncthreads();
        if (k == nupd - 1) tick(4);
        // ---- rank-nb update of the rows below the panel: C[i, :] -= L21[i, :] U12, one 8-row block per warp and step, four
        // blocks in flight
        if (nb == LF_NB && cw == LF_NB && (N & 1) == 0) {
            // full block, 16-byte aligned rows.  The fragment maps are permuted so that a lane's four k (its four columns of C)
            // are CONTIGUOUS: k = 4 fk + s in step s, local column j of tile h <-> column 4 (j >> 1) + 2 h + (j & 1).  A row of
            // L21 and a row of C are then one 128-byte line read by four lanes with two 16-byte loads each (the textbook
            // map k = 4 s + fk costs four 8-byte loads per lane and twice the load instructions -- the update was bound by
            // the LSU, one line per row and instruction: 4.4 us for 512 rows)
            const int fr = lane >> 2, fk = lane & 3;
            const int rbeg = j0 + LF_NB, nblk8 = (N - rbeg + 7) / 8;
            double bfrag[2][4];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int s2 = 0; s2 < 4; ++s2) bfrag[h][s2] = top[(4 * fk + s2) * LF_NB + 4 * (fr >> 1) + 2 * h + (fr & 1)];
            for (int blk0 = warp; blk0 < nblk8; blk0 += 4 * nwarps) {
                double2 af[4][2], cv[4][2];
                int irow[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int blk = blk0 + u * nwarps;
                    const int i = rbeg + 8 * blk + fr;
                    const bool rok = blk < nblk8 && i < N;
                    irow[u] = rok ? i : -1;
                    const double2* lrow = reinterpret_cast<const double2*>(a.A + (long long)(rok ? i : rbeg) * N + j0 + 4 * fk);
                    const double2* crow = reinterpret_cast<const double2*>(&cols.at(rok ? i : rbeg, 4 * fk));
                    af[u][0] = __ldcg(lrow);
                    af[u][1] = __ldcg(lrow + 1);
                    cv[u][0] = crow[0];
                    cv[u][1] = crow[1];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        lf_dmma(acc[h][0], acc[h][1], af[u][0].x, bfrag[h][0]);
                        lf_dmma(acc[h][0], acc[h][1], af[u][0].y, bfrag[h][1]);
                        lf_dmma(acc[h][0], acc[h][1], af[u][1].x, bfrag[h][2]);
                        lf_dmma(acc[h][0], acc[h][1], af[u][1].y, bfrag[h][3]);
                    }
                    if (irow[u] >= 0) {
                        double2* crow = reinterpret_cast<double2*>(&cols.at(irow[u], 4 * fk));
                        crow[0] = make_double2(cv[u][0].x - acc[0][0], cv[u][0].y - acc[0][1]);
                        crow[1] = make_double2(cv[u][1].x - acc[1][0], cv[u][1].y - acc[1][1]);
                    }
                }
            }
        } else {
            const int fr = lane >> 2, fk = lane & 3;
            const int rbeg = j0 + nb, nblk8 = (N - rbeg + 7) / 8;
            double bfrag[2][4];                               // B fragments: k = 4 s + fk, n = 8 h + fr (U12, shared memory)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const int kk = 4 * s + fk, n = 8 * h + fr;
                    bfrag[h][s] = (kk < nb && n < cw) ? top[kk * LF_NB + n] : 0.0;
                }
            const int cc0 = 2 * fk;
            for (int blk0 = warp; blk0 < nblk8; blk0 += 4 * nwarps) {
                double af[4][4], cv[4][2][2];
                int irow[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int blk = blk0 + u * nwarps;
                    const int i = rbeg + 8 * blk + fr;
                    const bool rok = blk < nblk8 && i < N;
                    irow[u] = rok ? i : -1;
                    const double* lrow = a.A + (long long)(rok ? i : rbeg) * N + j0;
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const int kk = 4 * s + fk;
                        af[u][s] = (rok && kk < nb) ? __ldcg(lrow + kk) : 0.0;
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int cc = 8 * h + cc0;
                        cv[u][h][0] = (rok && cc < cw) ? cols.at(i, cc) : 0.0;
                        cv[u][h][1] = (rok && cc + 1 < cw) ? cols.at(i, cc + 1) : 0.0;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (8 * h < cw) {                     // uniform
                            double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
                            for (int s = 0; s < 4; ++s) lf_dmma(acc0, acc1, af[u][s], bfrag[h][s]);
                            const int cc = 8 * h + cc0;
                            if (irow[u] >= 0) {
                                if (cc < cw) cols.at(irow[u], cc) = cv[u][h][0] - acc0;
                                if (cc + 1 < cw) cols.at(irow[u], cc + 1) = cv[u][h][1] - acc1;
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
    }

    tick(5);
    if (!is_rhs) {
        // ---- my block is panel p
        const int j0 = p * LF_NB, rows = N - j0;
        if (rows <= LF_THREADS) lf2_panel<1>(a, j0, cw, sm, s_rows, s_crow, s_key, s_idx, s_piv);
        else if (rows <= 2 * LF_THREADS) lf2_panel<2>(a, j0, cw, sm, s_rows, s_crow, s_key, s_idx, s_piv);
        else lf2_panel<3>(a, j0, cw, sm, s_rows, s_crow, s_key, s_idx, s_piv);
        __syncthreads();
        tick(6);
        if (tid == 0) {
            __threadfence();
            lf_st_release(a.flags + p, a.epoch);
        }
        tick(7);
        // ---- interchanges of later panels on my L columns (LAPACK's layout of the factors)
        const int rfirst = (p + 1) * LF_NB;
        if (rfirst >= N) return;
        if (tid == 0)
            while (lf_ld_acquire(a.flags + npan - 1) != a.epoch) {}
        __syncthreads();
        int* perm = reinterpret_cast<int*>(sm);               // [N]
        double* stage = sm + ((N + 1) / 2 + 1);               // [N - rfirst][LF_NB]
        for (int i = rfirst + tid; i < N; i += LF_THREADS) perm[i] = i;
        __syncthreads();
        if (tid == 0)
            for (int j = rfirst; j < N; ++j) {
                const int pv = __ldcg(a.ipiv + j);
                if (pv != j) {
                    const int t = perm[j];
                    perm[j] = perm[pv];
                    perm[pv] = t;
                }
            }
        __syncthreads();
        for (int e = tid; e < (N - rfirst) * cw; e += LF_THREADS) {
            const int i = rfirst + e / cw, c = e % cw;
            stage[(size_t)(i - rfirst) * LF_NB + c] = cols.at(perm[i], c);
        }
        __syncthreads();
        for (int e = tid; e < (N - rfirst) * cw; e += LF_THREADS) {
            const int i = rfirst + e / cw, c = e % cw;
            cols.at(i, c) = stage[(size_t)(i - rfirst) * LF_NB + c];
        }
        return;
    }
    // ---- right-hand side: every L update has been applied; back substitution with U (all panels are published)
    {
        double* x = sm;                                       // [N]
        double* D = x + N;                                    // [LF_NB][LF_LDP] diagonal block of U
        for (int i = tid; i < N; i += LF_THREADS) x[i] = a.b[i];
        __syncthreads();
        for (int k = npan - 1; k >= 0; --k) {
            const int j0 = k * LF_NB, nb = min(LF_NB, N - j0);
            for (int e = tid; e < nb * nb; e += LF_THREADS) {
                const int r = e / nb, c = e - r * nb;
                D[r * LF_LDP + c] = __ldcg(a.A + (long long)(j0 + r) * N + j0 + c);
            }
            __syncthreads();
            if (warp == 0) {
                double val = lane < nb ? x[j0 + lane] : 0.0;
                for (int c = nb - 1; c >= 0; --c) {
                    if (lane == c) val = val / D[c * LF_LDP + c];
                    const double xc = __shfl_sync(0xffffffffu, val, c);
                    if (lane < c) val = fma(-D[lane * LF_LDP + c], xc, val);
                }
                if (lane < nb) x[j0 + lane] = val;
            }
            __syncthreads();
            for (int i = tid; i < j0; i += LF_THREADS) {      // rows above: x_i -= U[i, block] x_block
                const double* u = a.A + (long long)i * N + j0;
                double acc = x[i];
#pragma unroll
                for (int c = 0; c < LF_NB; ++c)
                    if (c < nb) acc = fma(-__ldcg(u + c), x[j0 + c], acc);
                x[i] = acc;
            }
            __syncthreads();
        }
        for (int i = tid; i < N; i += LF_THREADS) a.b[i] = x[i];
    }
}

static size_t lf_smem(int N) {
    size_t panel = (size_t)N * LF_LDP * sizeof(double);
    size_t upd = ((size_t)LF_NB * LF_LDP + 2 * LF_NB * LF_NB) * sizeof(double);
    size_t fix = ((size_t)(N + 1) / 2 + 1 + (size_t)N * LF_NB) * sizeof(double);
    size_t back = ((size_t)N + LF_NB * LF_LDP) * sizeof(double);
    size_t m = panel;
    size_t stage = (size_t)LF_THREADS * LF_LDP * sizeof(double);
    if (stage > m) m = stage;
    if (upd > m) m = upd;
    if (fix > m) m = fix;
    if (back > m) m = back;
    return m + 64;
}

// Largest N the one-launch form takes (shared-memory panel of N x 17 doubles).
extern "C" int64_t sktt_lu_fused_max_n(void) { return 1536; }

// Solves A x = b for a dense fp64 system in ONE launch: A [N][N] row-major is overwritten by its LU factors (unit lower L,
// LAPACK layout), b [N] by the solution (b == NULL: factorisation only), ipiv_dev [N] receives the 0-based pivot rows,
// info_dev (int32, device) 0 or 1 + index of the first exactly-zero pivot.  Nothing is read back.
extern "C" int sktt_lu_solve_fused(sktt_ctx* ctx, int64_t N, void* A, void* b, int32_t* ipiv_dev, int32_t* info_dev) {
    if (!ctx || !A || !ipiv_dev || !info_dev) return SKTT_ERR_ARG;
    if (N < 1 || N > sktt_lu_fused_max_n()) return sktt_fail(ctx, SKTT_ERR_ARG, "lu_solve_fused: N out of range");
    const size_t smem = lf_smem((int)N);
    if (smem > 220 * 1024) return sktt_fail(ctx, SKTT_ERR_ARG, "lu_solve_fused: shared memory budget exceeded");
    SKTT_ONCE_PER_DEVICE(ctx);
    if (!configured) {
        SKTT_CUDA(ctx, cudaFuncSetAttribute(lu_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        SKTT_CUDA(ctx, cudaFuncSetAttribute(lu_fused_kernel_v1, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        configured = true;
    }
    SKTT_CUDA(ctx, cudaMemsetAsync(info_dev, 0, sizeof(int32_t), ctx->stream));
    const int npan = (int)((N + LF_NB - 1) / LF_NB), want2 = npan + (b ? 1 : 0);
    if (!(ctx->debug & 1024) && want2 <= ctx->sm_count) {
        // dataflow form: one CTA per block of 16 columns (+ one for the right-hand side), flags instead of grid barriers
        if (!ctx->lu_flags) {
            SKTT_CUDA(ctx, cudaMalloc(&ctx->lu_flags, 256 * sizeof(unsigned)));
            SKTT_CUDA(ctx, cudaMemsetAsync(ctx->lu_flags, 0, 256 * sizeof(unsigned), ctx->stream));
            ctx->lu_epoch = 0;
        }
        LfArgs2 a2;
        a2.N = (int)N;
        a2.A = (double*)A;
        a2.b = (double*)b;
        a2.ipiv = ipiv_dev;
        a2.info = info_dev;
        a2.flags = (unsigned*)ctx->lu_flags;
        a2.stamps = (ctx->debug & 1) ? (unsigned long long*)((char*)ctx->scratch + 3600) : nullptr;
        a2.epoch = ++ctx->lu_epoch;
        if (a2.epoch == 0) a2.epoch = ++ctx->lu_epoch;        // 0 is the cleared state
        void* args2[] = {&a2};
        SKTT_CUDA(ctx, cudaLaunchCooperativeKernel((void*)lu_fused_kernel, dim3(want2), dim3(LF_THREADS), args2, smem, ctx->stream));
        ctx->launches++;
        return 0;
    }
    LfArgs a;
    a.N = (int)N;
    a.A = (double*)A;
    a.b = (double*)b;
    a.ipiv = ipiv_dev;
    a.info = info_dev;
    int per_sm = 0;
    SKTT_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lu_fused_kernel_v1, LF_THREADS, smem));
    if (per_sm < 1) return sktt_fail(ctx, SKTT_ERR_ARG, "lu_solve_fused: kernel does not fit an SM");
    int grid = ctx->sm_count;
    const int want = (int)((N / LF_CB) + 1);
    if (grid > want) grid = want;
    void* args[] = {&a};
    SKTT_CUDA(ctx, cudaLaunchCooperativeKernel((void*)lu_fused_kernel_v1, dim3(grid), dim3(LF_THREADS), args, smem, ctx->stream));
    ctx->launches++;
    return 0;
}
