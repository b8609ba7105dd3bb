// batch.cu -- small independent micro problems, ONE CTA PER SYSTEM, the batch as the grid (BASELINE config 5: a sweep over
// 64 CO pressures of evp.als on co_oxidation(20); SURVEY.md 8b "batched variants", 8e "each GPU runs batched kernels over
// its share").  At these sizes (micro matrices of 9 ... 1024 unknowns) a host-driven eigen-solve is launch-bound: the
// single-system path issues ~235 launches per micro step.  Here the whole local eigen-solve of evp.py:417-432 -- shift,
// LU with partial pivoting, shift-invert Arnoldi with CGS2, the projected Hessenberg eigenproblem, Ritz selection,
// convergence test, explicit restarts, Ritz vectors, phase convention -- is one kernel launch for the whole batch, and the
// SVD re-orthonormalisation of the eigenvector block (evp.py:452-487) another.
//
//   beig_kernel : per system  S = M - sigma I  ->  P S = L U  ->  Arnoldi on S^-1  ->  (lambda, vectors) closest to sigma
//   bsvd_kernel : per system  left singular vectors of a small block (one-sided Jacobi in shared memory), null columns
//                 completed to an orthonormal set
#include "common.cuh"
#include "blas1.cuh"
#include "hess_eig.cuh"

#define BE_THREADS 512
#define BE_MAX_NCV 32
#define BE_MAX_K 8
#define BE_MAX_N 1024
#define BE_DBUF 2          // diagonal-block buffers (17 KB each): that many blocks are inverted concurrently

typedef Num<cplx> CX;

__device__ __forceinline__ cplx to_cplx(double v) { return make_cplx(v, 0.0); }
__device__ __forceinline__ cplx to_cplx(cplx v) { return v; }

struct BeigArgs {
    int N, k, m, max_restarts;
    double sigma, tol;
    const void* Min;   // [batch][N][N] row-major, double or cplx
    cplx* S;           // [batch][N][N] work (may alias Min for complex input)
    cplx* V;           // [batch][m + 1][N]
    cplx* lam;         // [batch][k]
    cplx* vecs;        // [batch][N][k]
    int v_in_smem;     // the Krylov basis fits into shared memory next to everything else
    int* status;       // [batch][2]: converged Ritz pairs, info (bit 0: Hessenberg QR failed, bit 1: zero pivot)
    double* relres;    // [batch]: largest residual estimate |h_{m+1,m} y_m| / |theta| among the k returned pairs
};


template <typename TIN>
__global__ void __launch_bounds__(BE_THREADS) beig_kernel(BeigArgs a) {
    extern __shared__ unsigned char smem_raw[];
    const int N = a.N, m = a.m, k = a.k;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = BE_THREADS / 32;
    const long long sys = blockIdx.x;
    cplx* S = a.S + sys * (long long)N * N;
    cplx* V = a.V + sys * (long long)(m + 1) * N;            // Krylov basis: global memory (L2) unless it fits below
    // shared-memory carve-up
    cplx* ws = (cplx*)smem_raw;                    // [N] current vector
    cplx* H = ws + N;                              // [(m + 1)][m]
    cplx* Hw = H + (m + 1) * m;                    // [3][m][m] work of the Hessenberg solver
    cplx* Y = Hw + 3 * m * m;                      // [m][m]
    cplx* theta = Y + m * m;                       // [m]
    cplx* Ysel = theta + m;                        // [m][k]
    cplx* h1 = Ysel + m * k;                       // [m + 1]
    cplx* hs = h1 + (m + 1);                       // [m + 1]
    cplx* D = hs + (m + 1);                        // [BE_DBUF][32][33] diagonal blocks (inversion: one per warp; solves: the first)
    cplx* tmp = D + BE_DBUF * 32 * 33;             // [32]
    cplx* red = tmp + 32;                          // [32]
    double* dred = (double*)(red + 32);            // [32]
    int* ired = (int*)(dred + 32);                 // [32]
    int* perm = ired + 32;                         // [N]
    int* sflag = perm + N;                         // [4]: pivot row, nconv, hess info, lu info
    cplx* invd = (cplx*)(((uintptr_t)(sflag + 4) + 15) & ~(uintptr_t)15);   // [N] reciprocals of the pivots (diagonal of U)
    cplx* urow = invd + N;                         // [N] pivot row of the current LU column
    cplx* lcol = urow + N;                         // [N] multipliers of the current LU column
    if (a.v_in_smem) V = lcol + N;                 // [(m + 1)][N] Krylov basis in shared memory

    // ---- S = M - sigma I
    {
        const TIN* Min = (const TIN*)a.Min + sys * (long long)N * N;
        for (long long e = tid; e < (long long)N * N; e += BE_THREADS) {
            cplx v = to_cplx(Min[e]);
            if (e / N == e % N) v.re -= a.sigma;
            S[e] = v;
        }
        for (int i = tid; i < N; i += BE_THREADS) perm[i] = i;
        if (tid == 0) { sflag[1] = 0; sflag[2] = 0; sflag[3] = 0; }
    }
    __syncthreads();

    // ---- P S = L U, right-looking, partial pivoting on |re| + |im| (LAPACK izamax).  Three CTA barriers per column:
    //   1. row interchange j <-> p; the pivot row (columns > j) is staged in shared memory on the way
    //   2. multipliers l_i = S[i,j] / pivot: written back and staged in shared memory
    //   3. rank-1 update of the trailing block from the two staged vectors; the threads that produce column j + 1 also
    //      vote for the next pivot (atomicMax on a key made of the truncated magnitude and the row index)
    // Blocked form (used whenever the Krylov basis lives in shared memory, whose space is idle during the factorisation):
    // panels of 8 columns are factorised in shared memory, the interchanges reach the rest of the matrix as ONE gather /
    // scatter per column (net permutation of the <= 16 touched rows), U12 by forward substitution in registers, and the
    // trailing block is read and written once per PANEL instead of once per column (N = 192: 9 MB instead of 75 MB through
    // one SM's L2 port, 24 instead of 192 rounds of dependent L2 round trips).  Same elimination order, same fma sequence
    // per entry as the column form below: the factors are bit-identical; the pivot is the exact first maximum of
    // |re| + |im| (LAPACK's izamax).
    constexpr int PB = 8;
    const bool blocked = a.v_in_smem && m + 1 >= 2 * PB;
    if (blocked) {
        cplx* Pn = V;                                   // [N][PB] panel rows (row i = matrix row j0 + i)
        cplx* Ust = Pn + (size_t)N * PB;                // [PB][N] rows of U12 (absolute column index)
        __shared__ int b_pos[2 * PB], b_content[2 * PB], b_piv[PB], b_np;
        for (int j0 = 0; j0 < N; j0 += PB) {
            const int nb = min(PB, N - j0), R = N - j0;
            for (int e = tid; e < R * PB; e += BE_THREADS) {
                const int i = e / PB, c = e % PB;
                Pn[e] = c < nb ? S[(long long)(j0 + i) * N + j0 + c] : CX::zero();
            }
            __syncthreads();
            for (int c = 0; c < nb; ++c) {
                double best = -1.0;
                int bi = c;
                for (int i = c + tid; i < R; i += BE_THREADS) {
                    const cplx v = Pn[i * PB + c];
                    const double c1 = fabs(v.re) + fabs(v.im);
                    if (c1 > best) { best = c1; bi = i; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                }
                if (lane == 0) { tmp[warp].re = best; ired[warp] = bi; }
                __syncthreads();
                best = tmp[0].re;
                bi = ired[0];
                for (int w = 1; w < nwarps; ++w)
                    if (tmp[w].re > best || (tmp[w].re == best && ired[w] < bi)) { best = tmp[w].re; bi = ired[w]; }
                const int p = bi;                             // the same in every thread
                cplx mine = CX::zero(), theirs = CX::zero();
                if (tid < PB) { mine = Pn[c * PB + tid]; theirs = Pn[p * PB + tid]; }
                __syncthreads();                              // everybody has read the candidates and the two rows
                if (tid < PB && p != c) {
                    Pn[c * PB + tid] = theirs;
                    Pn[p * PB + tid] = mine;
                }
                if (tid == 0) {
                    b_piv[c] = j0 + p;
                    const int t = perm[j0 + c];
                    perm[j0 + c] = perm[j0 + p];
                    perm[j0 + p] = t;
                }
                __syncthreads();
                const cplx piv = Pn[c * PB + c];
                const bool okp = (fabs(piv.re) + fabs(piv.im)) > 0.0;
                const cplx inv = okp ? CX::div(CX::one(), piv) : CX::zero();
                for (int i = c + 1 + tid; i < R; i += BE_THREADS) {
                    cplx* row = Pn + i * PB;
                    const cplx l = CX::mul(row[c], inv);
                    row[c] = l;
#pragma unroll
                    for (int cc = 1; cc < PB; ++cc)
                        if (cc > c && cc < nb) {
                            const cplx uu = Pn[c * PB + cc];
                            cplx w = row[cc];
                            w.re = fma(-l.re, uu.re, w.re);
                            w.re = fma(l.im, uu.im, w.re);
                            w.im = fma(-l.re, uu.im, w.im);
                            w.im = fma(-l.im, uu.re, w.im);
                            row[cc] = w;
                        }
                }
                if (tid == 0) {
                    invd[j0 + c] = inv;
                    if (!okp) sflag[3] |= 2;
                }
                __syncthreads();
            }
            for (int e = tid; e < R * PB; e += BE_THREADS) {
                const int i = e / PB, c = e % PB;
                if (c < nb) S[(long long)(j0 + i) * N + j0 + c] = Pn[e];
            }
            if (warp == 0) {
                // rows the interchanges touch: the nb top rows, then every pivot row not yet listed (slot = lane);
                // `content` follows the rows through the sequence of interchanges
                int pos = lane < nb ? j0 + lane : -1, content = lane, np = nb;
                const int mypiv = lane < nb ? b_piv[lane] : -1;
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
                if (lane < 2 * PB) {
                    b_pos[lane] = pos;
                    b_content[lane] = content;
                }
                if (lane == 0) b_np = np;
            }
            __syncthreads();
            const int np = b_np;
            // the interchanges on the columns outside the panel (a thread per column, all its loads in flight), and U12 on the
            // columns to the right
            for (int col = tid; col < N; col += BE_THREADS) {
                if (col >= j0 && col < j0 + nb) continue;
                cplx v[2 * PB];
#pragma unroll
                for (int q = 0; q < 2 * PB; ++q)
                    v[q] = q < np ? S[(long long)b_pos[b_content[q]] * N + col] : CX::zero();
                if (col >= j0 + nb) {
#pragma unroll
                    for (int r = 1; r < PB; ++r)
                        if (r < nb) {
#pragma unroll
                            for (int c = 0; c < PB; ++c)
                                if (c < r) {
                                    const cplx l = Pn[r * PB + c], uu = v[c];
                                    cplx w = v[r];
                                    w.re = fma(-l.re, uu.re, w.re);
                                    w.re = fma(l.im, uu.im, w.re);
                                    w.im = fma(-l.re, uu.im, w.im);
                                    w.im = fma(-l.im, uu.re, w.im);
                                    v[r] = w;
                                }
                        }
#pragma unroll
                    for (int r = 0; r < PB; ++r)
                        if (r < nb) Ust[r * N + col] = v[r];
                }
#pragma unroll
                for (int q = 0; q < 2 * PB; ++q)
                    if (q < np) S[(long long)b_pos[q] * N + col] = v[q];
            }
            __syncthreads();
            // trailing block: S[i, col] -= sum_c L21[i, c] U12[c, col], a warp per pair of rows, lanes along the row
            for (int i = j0 + nb + 2 * warp; i < N; i += 2 * nwarps) {
                const bool two = i + 1 < N;
                cplx l0[PB], l1[PB];
#pragma unroll
                for (int c = 0; c < PB; ++c) {
                    l0[c] = c < nb ? Pn[(i - j0) * PB + c] : CX::zero();
                    l1[c] = (two && c < nb) ? Pn[(i + 1 - j0) * PB + c] : CX::zero();
                }
                cplx* row0 = S + (long long)i * N;
                cplx* row1 = row0 + N;
                for (int c0 = j0 + nb + lane; c0 < N; c0 += 32 * 2) {
                    cplx v0[2], v1[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int c = c0 + 32 * u;
                        v0[u] = c < N ? row0[c] : CX::zero();
                        v1[u] = (c < N && two) ? row1[c] : CX::zero();
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int c = c0 + 32 * u;
                        if (c < N) {
                            cplx w0 = v0[u], w1 = v1[u];
#pragma unroll
                            for (int cc = 0; cc < PB; ++cc)
                                if (cc < nb) {
                                    const cplx uu = Ust[cc * N + c];
                                    w0.re = fma(-l0[cc].re, uu.re, w0.re);
                                    w0.re = fma(l0[cc].im, uu.im, w0.re);
                                    w0.im = fma(-l0[cc].re, uu.im, w0.im);
                                    w0.im = fma(-l0[cc].im, uu.re, w0.im);
                                    w1.re = fma(-l1[cc].re, uu.re, w1.re);
                                    w1.re = fma(l1[cc].im, uu.im, w1.re);
                                    w1.im = fma(-l1[cc].re, uu.im, w1.im);
                                    w1.im = fma(-l1[cc].im, uu.re, w1.im);
                                }
                            row0[c] = w0;
                            if (two) row1[c] = w1;
                        }
                    }
                }
            }
            __syncthreads();
        }
    } else {
        unsigned long long* pkey = reinterpret_cast<unsigned long long*>(dred);        // [0]: next pivot key (dred is free here)
        {
            // pivot of column 0
            double best = -1.0;
            int bi = 0;
            for (int i = tid; i < N; i += BE_THREADS) {
                cplx v = S[(long long)i * N];
                double c1 = fabs(v.re) + fabs(v.im);
                if (c1 > best) { best = c1; bi = i; }
            }
    #pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double ob = __shfl_xor_sync(0xffffffffu, best, o);
                int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (lane == 0) { tmp[warp].re = best; ired[warp] = bi; }
            __syncthreads();
            if (tid == 0) {
                for (int w = 1; w < nwarps; ++w)
                    if (tmp[w].re > best || (tmp[w].re == best && ired[w] < bi)) { best = tmp[w].re; bi = ired[w]; }
                sflag[0] = bi;
            }
            __syncthreads();
        }
        for (int j = 0; j < N; ++j) {
            const int p = sflag[0];
            cplx* rowj = S + (long long)j * N;
            cplx* rowp = S + (long long)p * N;
            for (int c = tid; c < N; c += BE_THREADS) {                 // (1)
                cplx aj = rowj[c], ap = rowp[c];
                if (p != j) {
                    rowj[c] = ap;
                    rowp[c] = aj;
                }
                if (c >= j) urow[c] = ap;
            }
            if (tid == 0) {
                int t = perm[j];
                perm[j] = perm[p];
                perm[p] = t;
                pkey[0] = 0ull;
            }
            __syncthreads();
            const cplx piv = urow[j];
            const bool okp = (fabs(piv.re) + fabs(piv.im)) > 0.0;
            const cplx inv = okp ? CX::div(CX::one(), piv) : CX::zero();
            for (int i = j + 1 + tid; i < N; i += BE_THREADS) {         // (2)
                cplx l = CX::mul(S[(long long)i * N + j], inv);
                S[(long long)i * N + j] = l;
                lcol[i] = l;
            }
            if (tid == 0) {
                invd[j] = inv;
                if (!okp) sflag[3] |= 2;
            }
            __syncthreads();
            // (3): a warp per row, lanes along the row (coalesced), four independent 16-byte loads in flight per thread -- the
            // trailing block lives in L2, so the update is bound by how many loads a CTA keeps outstanding
            for (int i = j + 1 + 2 * warp; i < N; i += 2 * nwarps) {
                const bool two = i + 1 < N;
                const cplx l0 = lcol[i], l1 = two ? lcol[i + 1] : CX::zero();
                cplx* row0 = S + (long long)i * N;
                cplx* row1 = row0 + N;
                for (int c0 = j + 1 + lane; c0 < N; c0 += 32 * 4) {
                    cplx v0[4], v1[4];
    #pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int c = c0 + 32 * u;
                        if (c < N) {
                            v0[u] = row0[c];
                            if (two) v1[u] = row1[c];
                        }
                    }
    #pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int c = c0 + 32 * u;
                        if (c < N) {
                            const cplx uu = urow[c];
                            cplx w = v0[u];
                            w.re = fma(-l0.re, uu.re, w.re);
                            w.re = fma(l0.im, uu.im, w.re);
                            w.im = fma(-l0.re, uu.im, w.im);
                            w.im = fma(-l0.im, uu.re, w.im);
                            row0[c] = w;
                            cplx w1 = v1[u];
                            if (two) {
                                w1.re = fma(-l1.re, uu.re, w1.re);
                                w1.re = fma(l1.im, uu.im, w1.re);
                                w1.im = fma(-l1.re, uu.im, w1.im);
                                w1.im = fma(-l1.im, uu.re, w1.im);
                                row1[c] = w1;
                            }
                            if (c == j + 1) {
                                // key: magnitude with its 10 lowest mantissa bits replaced by (1023 - row): the largest |.|
                                // wins, ties (to 2^-42 relative) go to the smaller row; positive doubles order like their bits
                                const double c1 = fabs(w.re) + fabs(w.im);
                                atomicMax(pkey, ((unsigned long long)__double_as_longlong(c1) & ~0x3FFull) | (unsigned long long)(1023 - i));
                                if (two) {
                                    const double c2 = fabs(w1.re) + fabs(w1.im);
                                    atomicMax(pkey, ((unsigned long long)__double_as_longlong(c2) & ~0x3FFull) |
                                                        (unsigned long long)(1023 - (i + 1)));
                                }
                            }
                        }
                    }
                }
            }
            __syncthreads();
            if (tid == 0 && j + 1 < N) sflag[0] = 1023 - (int)(pkey[0] & 0x3FFull);
            __syncthreads();
        }
    }
    // ---- the 32 x 32 diagonal blocks of L and U are replaced by their inverses (in place: strictly lower part L_kk^-1,
    // upper part U_kk^-1), so that the triangular solves below apply them as small dense products instead of walking a
    // dependent chain of 32 substitution steps
    {
        const int nblk = (N + 31) / 32;
        for (int k0 = 0; k0 < nblk; k0 += BE_DBUF) {
            const int kb = warp < BE_DBUF ? k0 + warp : nblk;       // warp w < BE_DBUF inverts block k0 + w in its own buffer
            cplx* Dw = D + (size_t)(warp < BE_DBUF ? warp : 0) * (32 * 33);
            if (kb < nblk) {
                const int i0 = kb * 32, nb = min(32, N - i0);
                for (int e = lane; e < nb * nb; e += 32) Dw[(e / nb) * 33 + e % nb] = S[(long long)(i0 + e / nb) * N + i0 + e % nb];
                __syncwarp();
                if (lane < nb) {
                    const int c = lane;                             // this lane produces column c of both inverses
                    {
                        // L^-1 (unit lower): x_c = 1, x_i = - sum_{t=c}^{i-1} L[i][t] x_t
                        cplx x[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) x[i] = CX::zero();
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (i == c) x[i] = CX::one();
                            else if (i > c && i < nb) {
                                cplx acc = CX::zero();
#pragma unroll
                                for (int t = 0; t < 32; ++t)
                                    if (t >= c && t < i) CX::fma(acc, Dw[i * 33 + t], x[t]);
                                x[i] = CX::neg(acc);
                                S[(long long)(i0 + i) * N + i0 + c] = x[i];      // Dw is a copy: the block in S is free to take it
                            }
                        }
                    }
                    {
                        // U^-1 (upper): y_c = 1 / U[c][c], y_i = - (sum_{t=i+1}^{c} U[i][t] y_t) / U[i][i]
                        cplx y[32];
#pragma unroll
                        for (int i = 31; i >= 0; --i) {
                            y[i] = CX::zero();
                            if (i == c) y[i] = invd[i0 + c];
                            else if (i < c) {
                                cplx acc = CX::zero();
#pragma unroll
                                for (int t = 0; t < 32; ++t)
                                    if (t > i && t <= c) CX::fma(acc, Dw[i * 33 + t], y[t]);
                                y[i] = CX::neg(CX::mul(acc, invd[i0 + i]));
                            }
                            if (i <= c) S[(long long)(i0 + i) * N + i0 + c] = y[i];
                        }
                    }
                }
            }
            __syncthreads();
        }
    }

    // ws <- S^-1 (P applied by the caller): forward substitution with unit-lower L, backward with U, 32 rows at a time:
    // all warps form the block's dot products with the part already solved, warp 0 applies the inverted diagonal block
    auto solve = [&]() {
        const int nblk = (N + 31) / 32;
        for (int I = 0; I < nblk; ++I) {
            const int i0 = I * 32, nb = min(32, N - i0);
            for (int row = warp; row < nb; row += nwarps) {
                const cplx* Li = S + (long long)(i0 + row) * N;
                cplx acc = CX::zero();
#pragma unroll 8
                for (int c = lane; c < i0; c += 32) CX::fma(acc, Li[c], ws[c]);
                acc = warp_sum<cplx>(acc);
                if (lane == 0) tmp[row] = acc;
            }
            for (int e = tid; e < nb * nb; e += BE_THREADS) D[(e / nb) * 33 + e % nb] = S[(long long)(i0 + e / nb) * N + i0 + e % nb];
            __syncthreads();
            if (warp == 0) {
                const cplx t = lane < nb ? CX::sub(ws[i0 + lane], tmp[lane]) : CX::zero();
                cplx val = t;                                       // unit diagonal
                for (int c = 0; c < nb; ++c) {
                    const cplx tc = lane_bcast<cplx>(t, c);
                    if (lane > c && lane < nb) CX::fma(val, D[lane * 33 + c], tc);
                }
                if (lane < nb) ws[i0 + lane] = val;
            }
            __syncthreads();
        }
        for (int I = nblk - 1; I >= 0; --I) {
            const int i0 = I * 32, nb = min(32, N - i0);
            for (int row = warp; row < nb; row += nwarps) {
                const cplx* Ui = S + (long long)(i0 + row) * N;
                cplx acc = CX::zero();
#pragma unroll 8
                for (int c = i0 + nb + lane; c < N; c += 32) CX::fma(acc, Ui[c], ws[c]);
                acc = warp_sum<cplx>(acc);
                if (lane == 0) tmp[row] = acc;
            }
            for (int e = tid; e < nb * nb; e += BE_THREADS) D[(e / nb) * 33 + e % nb] = S[(long long)(i0 + e / nb) * N + i0 + e % nb];
            __syncthreads();
            if (warp == 0) {
                const cplx t = lane < nb ? CX::sub(ws[i0 + lane], tmp[lane]) : CX::zero();
                cplx val = CX::zero();
                for (int c = 0; c < nb; ++c) {
                    const cplx tc = lane_bcast<cplx>(t, c);
                    if (lane <= c && lane < nb) CX::fma(val, D[lane * 33 + c], tc);
                }
                if (lane < nb) ws[i0 + lane] = val;
            }
            __syncthreads();
        }
    };

    // ---- shift-invert Arnoldi with explicit restarts
    cplx* vstart = V + (long long)m * N;            // the start vector of a cycle lives in the (m+1)-th slot until used
    for (int i = tid; i < N; i += BE_THREADS) vstart[i] = CX::one();      // evp.py:418: v0 = ones
    __syncthreads();
    int nconv = 0;
    double worst_rel = 1e300;
    for (int restart = 0; restart <= a.max_restarts; ++restart) {
        {   // V[0] = start / ||start||
            double s2 = 0.0;
            for (int i = tid; i < N; i += BE_THREADS) s2 += CX::abs2(vstart[i]);
            s2 = block_sum<double>(s2, dred);
            const double inv = s2 > 0.0 ? 1.0 / sqrt(s2) : 0.0;
            for (int i = tid; i < N; i += BE_THREADS) V[i] = CX::scale(vstart[i], inv);
            for (int e = tid; e < (m + 1) * m; e += BE_THREADS) H[e] = CX::zero();
        }
        __syncthreads();
        double hlast = 0.0;
        for (int j = 0; j < m; ++j) {
            const cplx* vj = V + (long long)j * N;
            for (int i = tid; i < N; i += BE_THREADS) ws[i] = vj[perm[i]];
            for (int i = tid; i <= j; i += BE_THREADS) hs[i] = CX::zero();
            __syncthreads();
            solve();
            for (int pass = 0; pass < 2; ++pass) {                       // classical Gram-Schmidt, twice
                for (int i = warp; i <= j; i += nwarps) {
                    const cplx* vi = V + (long long)i * N;
                    cplx acc = CX::zero();
                    for (int n = lane; n < N; n += 32) CX::fma(acc, CX::conj(vi[n]), ws[n]);
                    acc = warp_sum<cplx>(acc);
                    if (lane == 0) { h1[i] = acc; hs[i] = CX::add(hs[i], acc); }
                }
                __syncthreads();
                for (int n = tid; n < N; n += BE_THREADS) {
                    cplx v = ws[n];
                    for (int i = 0; i <= j; ++i) v = CX::sub(v, CX::mul(V[(long long)i * N + n], h1[i]));
                    ws[n] = v;
                }
                __syncthreads();
            }
            double s2 = 0.0;
            for (int i = tid; i < N; i += BE_THREADS) s2 += CX::abs2(ws[i]);
            s2 = block_sum<double>(s2, dred);
            const double nrm = sqrt(s2 > 0.0 ? s2 : 0.0);
            const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
            for (int i = tid; i <= j; i += BE_THREADS) H[i * m + j] = hs[i];
            if (tid == 0 && j + 1 < m) H[(j + 1) * m + j] = make_cplx(nrm, 0.0);
            cplx* vn = V + (long long)(j + 1) * N;
            for (int i = tid; i < N; i += BE_THREADS) vn[i] = CX::scale(ws[i], inv);
            hlast = nrm;
            __syncthreads();
        }
        // projected eigenproblem (one warp), selection of the k Ritz values of largest modulus, convergence
        if (warp == 0) hess_eig_warp(m, H, theta, Y, sflag + 2, Hw, Hw + m * m, Hw + 2 * m * m);
        __syncthreads();
        if (tid == 0) {
            int conv = 0;
            double worst = 0.0;
            unsigned used = 0u;
            const double hn = m < N ? hlast : 0.0;
            for (int s = 0; s < k; ++s) {
                int bestj = -1;
                double bv = -1.0;
                for (int i = 0; i < m; ++i) {
                    if ((used >> i) & 1u) continue;
                    double v = cabs_(theta[i]);
                    if (v > bv) { bv = v; bestj = i; }
                }
                used |= 1u << bestj;
                a.lam[sys * k + s] = CX::add(make_cplx(a.sigma, 0.0), CX::div(CX::one(), theta[bestj]));
                for (int r = 0; r < m; ++r) Ysel[r * k + s] = Y[r * m + bestj];
                double res = hn * cabs_(Y[(m - 1) * m + bestj]);
                if (bv > 0.0 && res <= a.tol * bv) conv++;
                const double rel = bv > 0.0 ? res / bv : 1e300;
                worst = rel > worst ? rel : worst;
            }
            sflag[1] = conv;
            dred[0] = worst;
        }
        __syncthreads();
        nconv = sflag[1];
        const double prev_rel = worst_rel;
        worst_rel = dred[0];
        // a cycle that does not gain a factor of ten on the residual estimate will not reach the tolerance in any
        // reasonable number of restarts (eigenvalues at nearly equal distance from sigma): stop, report unconverged
        const bool stagnated = restart >= 1 && worst_rel > 0.1 * prev_rel;
        __syncthreads();
        cplx* vecs = a.vecs + sys * (long long)N * k;
        for (int e = tid; e < N * k; e += BE_THREADS) {
            int n = e / k, s = e % k;
            cplx acc = CX::zero();
            for (int i = 0; i < m; ++i) CX::fma(acc, V[(long long)i * N + n], Ysel[i * k + s]);
            vecs[e] = acc;
        }
        __syncthreads();
        if (nconv >= k || m >= N || restart == a.max_restarts || sflag[2] != 0 || stagnated) break;
        for (int n = tid; n < N; n += BE_THREADS) {                       // restart from the sum of the wanted Ritz vectors
            cplx acc = CX::zero();
            for (int s = 0; s < k; ++s) acc = CX::add(acc, vecs[n * k + s]);
            vstart[n] = acc;
        }
        __syncthreads();
    }
    // ---- phase convention of LAPACK's geev: the largest component of every vector is real and positive
    {
        cplx* vecs = a.vecs + sys * (long long)N * k;
        for (int s = 0; s < k; ++s) {
            double best = -1.0;
            int bi = 0;
            for (int i = tid; i < N; i += BE_THREADS) {
                double v = CX::abs2(vecs[i * k + s]);
                if (v > best) { best = v; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double ob = __shfl_xor_sync(0xffffffffu, best, o);
                int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            __syncthreads();
            if (lane == 0) { dred[warp] = best; ired[warp] = bi; }
            __syncthreads();
            if (tid == 0) {
                for (int w = 1; w < nwarps; ++w)
                    if (dred[w] > best || (dred[w] == best && ired[w] < bi)) { best = dred[w]; bi = ired[w]; }
                cplx v = vecs[bi * k + s];
                double ab = hypot(v.re, v.im);
                tmp[0] = ab > 0.0 ? make_cplx(v.re / ab, -v.im / ab) : CX::one();
            }
            __syncthreads();
            const cplx rot = tmp[0];
            for (int i = tid; i < N; i += BE_THREADS) vecs[i * k + s] = CX::mul(vecs[i * k + s], rot);
            __syncthreads();
        }
    }
    if (tid == 0) {
        a.status[sys * 2 + 0] = nconv;
        a.status[sys * 2 + 1] = (sflag[2] ? 1 : 0) | sflag[3];
        a.relres[sys] = worst_rel;
    }
}

static size_t beig_smem(int N, int m, int k) {
    size_t c = (size_t)N + (size_t)(m + 1) * m + 3 * (size_t)m * m + (size_t)m * m + m + (size_t)m * k + 2 * (m + 1) +
               (size_t)BE_DBUF * 32 * 33 + 32 + 32;
    return c * sizeof(cplx) + 32 * sizeof(double) + 32 * sizeof(int) + (size_t)N * sizeof(int) + 4 * sizeof(int) + 64 +
           3 * (size_t)N * sizeof(cplx);
}

extern "C" int64_t sktt_batch_eig_work(int64_t batch, int64_t N, int64_t k, int64_t ncv) {
    if (ncv > BE_MAX_NCV) ncv = BE_MAX_NCV;
    (void)k;
    return batch * (N * N + (ncv + 1) * N) + 64;        // complex128 elements: S and the Krylov basis
}

// k eigenpairs closest to sigma of each of `batch` dense N x N matrices (double or complex128, row-major, contiguous);
// lam [batch][k], vecs [batch][N][k] complex128, status [batch][2] int32 (converged pairs, info) and relres [batch] (largest
// relative residual estimate of the returned pairs) on the DEVICE -- nothing is read back here, the caller inspects them
// when it synchronises next.
extern "C" int sktt_batch_eig_shift_invert(sktt_ctx* ctx, int dtype, int64_t batch, int64_t N, const void* Mat, double sigma,
                                           int64_t k, int64_t ncv, double tol, int max_restarts, void* lam, void* vecs,
                                           void* work, int32_t* status_dev, double* relres_dev) {
    if (!ctx || !Mat || !lam || !vecs || !work || !status_dev || !relres_dev) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (batch < 1 || N < 1 || N > BE_MAX_N || k < 1 || k > BE_MAX_K || k > N)
        return sktt_fail(ctx, SKTT_ERR_ARG, "batch_eig_shift_invert: extents out of range (N <= 1024, k <= 8)");
    int m = (int)(ncv > BE_MAX_NCV ? BE_MAX_NCV : ncv);
    if (m > N) m = (int)N;
    if (m < k) return sktt_fail(ctx, SKTT_ERR_ARG, "batch_eig_shift_invert: Krylov dimension below k");
    BeigArgs a;
    a.N = (int)N; a.k = (int)k; a.m = m; a.max_restarts = max_restarts;
    a.sigma = sigma; a.tol = tol;
    a.Min = Mat;
    a.S = (cplx*)work;
    a.V = a.S + batch * N * N;
    a.lam = (cplx*)lam;
    a.vecs = (cplx*)vecs;
    a.status = status_dev;
    a.relres = relres_dev;
    size_t smem = beig_smem((int)N, m, (int)k);
    if (smem > 200 * 1024) return sktt_fail(ctx, SKTT_ERR_ARG, "batch_eig_shift_invert: shared memory budget exceeded");
    const size_t vbytes = (size_t)(m + 1) * N * sizeof(cplx);
    a.v_in_smem = smem + vbytes <= 200 * 1024 ? 1 : 0;
    if (a.v_in_smem) smem += vbytes;
    SKTT_ONCE_PER_DEVICE(ctx);
    if (!configured) {
        SKTT_CUDA(ctx, cudaFuncSetAttribute(beig_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        SKTT_CUDA(ctx, cudaFuncSetAttribute(beig_kernel<cplx>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    if (dtype == SKTT_F64) beig_kernel<double><<<(unsigned)batch, BE_THREADS, smem, ctx->stream>>>(a);
    else beig_kernel<cplx><<<(unsigned)batch, BE_THREADS, smem, ctx->stream>>>(a);
    SKTT_LAUNCH_CHECK(ctx);
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// Left singular vectors of small blocks.  F is P x Q with F(i, j) = op(in[fi(i) + fj(j)]) (two-level index maps, optional
// conjugation -- so that F can be the eigenvector block of evp.py:452 or the conjugate transpose of the one of evp.py:475
// without a copy); the first `keep` left singular vectors (descending singular values) are written as
// out[i * so_i + t * so_t] (optionally conjugated).
//   P >= Q: one-sided Jacobi on the columns of F; U = normalised columns, null columns completed (Gram-Schmidt of unit
//           vectors) as LAPACK's full_matrices=True would supply some orthonormal completion.
//   P <  Q: one-sided Jacobi on the columns of [F^H; I_P]; the lower block accumulates the rotations = U.
struct BsvdArgs {
    int P, Q, keep, conj_in, conj_out;
    const cplx* in;
    long long in_stride;      // elements between systems
    Idx2 fi, fj;
    cplx* out;
    long long out_stride, so_i, so_t;
    int* sweeps;              // [batch] (may be null)
};

#define BS_THREADS 512

__global__ void __launch_bounds__(BS_THREADS) bsvd_kernel(BsvdArgs a) {
    extern __shared__ unsigned char smem_raw[];
    const int P = a.P, Q = a.Q;
    const bool wide = P < Q;
    const int nc = wide ? P : Q;                 // columns that are rotated
    const int mw = wide ? Q : P;                 // length of the part that defines the rotations
    const int L = wide ? Q + P : P;              // stored column length
    cplx* X = (cplx*)smem_raw;                   // [nc][L]
    double* sig = (double*)(X + (size_t)nc * L); // [nc]
    int* ord = (int*)(sig + nc);                 // [nc]
    cplx* u = (cplx*)(((uintptr_t)(ord + nc) + 15) & ~(uintptr_t)15);   // [P] completion scratch
    __shared__ cplx red[32];
    __shared__ double dred[32];
    __shared__ int s_rot;
    __shared__ double s_nrm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = BS_THREADS / 32;
    const cplx* in = a.in + (long long)blockIdx.x * a.in_stride;
    cplx* out = a.out + (long long)blockIdx.x * a.out_stride;

    for (int e = tid; e < nc * L; e += BS_THREADS) {
        int j = e / L, i = e % L;
        cplx v;
        if (!wide) {
            v = in[a.fi(i) + a.fj(j)];
            if (a.conj_in) v = CX::conj(v);
        } else if (i < Q) {
            v = in[a.fi(j) + a.fj(i)];           // (F^H)(i, j) = conj(F(j, i))
            if (!a.conj_in) v = CX::conj(v);
        } else {
            v = (i - Q == j) ? CX::one() : CX::zero();
        }
        X[e] = v;
    }
    __syncthreads();
    const double tol = 2.220446049250313e-16 * sqrt((double)(mw > 4 ? mw : 4));
    const int np = nc + (nc & 1), half = np / 2;
    int sweep = 0;
    for (; sweep < 60; ++sweep) {
        if (tid == 0) s_rot = 0;
        __syncthreads();
        int my_rot = 0;
        for (int step = 0; step < np - 1; ++step) {
            for (int kk = warp; kk < half; kk += nwarps) {
                int p, q;
                if (kk == 0) { p = np - 1; q = step; }
                else { p = (step + kk) % (np - 1); q = (step - kk + (np - 1)) % (np - 1); }
                if (p > q) { int t = p; p = q; q = t; }
                if (q >= nc) continue;
                cplx* xp = X + (size_t)p * L;
                cplx* xq = X + (size_t)q * L;
                double aa = 0.0, bb = 0.0;
                cplx g = CX::zero();
                for (int i = lane; i < mw; i += 32) {
                    cplx vp = xp[i], vq = xq[i];
                    aa += CX::abs2(vp);
                    bb += CX::abs2(vq);
                    CX::fma(g, CX::conj(vp), vq);
                }
                aa = warp_sum<double>(aa);
                bb = warp_sum<double>(bb);
                g = warp_sum<cplx>(g);
                double gabs = sqrt(CX::abs2(g));
                if (gabs == 0.0 || gabs <= tol * sqrt(aa) * sqrt(bb)) continue;
                my_rot++;
                double zeta = (bb - aa) / (2.0 * gabs);
                double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                cplx w = CX::scale(g, 1.0 / gabs);
                cplx sw = CX::scale(w, s), swc = CX::conj(sw);
                for (int i = lane; i < L; i += 32) {
                    cplx vp = xp[i], vq = xq[i];
                    xp[i] = CX::sub(CX::scale(vp, c), CX::mul(swc, vq));
                    xq[i] = CX::add(CX::mul(sw, vp), CX::scale(vq, c));
                }
            }
            __syncthreads();
        }
        if (lane == 0 && my_rot) atomicAdd(&s_rot, my_rot);
        __syncthreads();
        const int total = s_rot;
        __syncthreads();
        if (total == 0) { ++sweep; break; }
    }
    if (tid == 0 && a.sweeps) a.sweeps[blockIdx.x] = sweep;
    // singular values = norms of the rotated part; descending order (stable)
    for (int j = warp; j < nc; j += nwarps) {
        double s2 = 0.0;
        for (int i = lane; i < mw; i += 32) s2 += CX::abs2(X[(size_t)j * L + i]);
        s2 = warp_sum<double>(s2);
        if (lane == 0) sig[j] = sqrt(s2);
    }
    __syncthreads();
    for (int j = tid; j < nc; j += BS_THREADS) {
        int rank = 0;
        double sj = sig[j];
        for (int i = 0; i < nc; ++i) rank += (sig[i] > sj || (sig[i] == sj && i < j)) ? 1 : 0;
        ord[rank] = j;
    }
    __syncthreads();
    const int keep = a.keep;
    auto store = [&](int i, int t, cplx v) {
        if (a.conj_out) v = CX::conj(v);
        out[(long long)i * a.so_i + (long long)t * a.so_t] = v;
    };
    if (wide) {
        for (int e = tid; e < keep * P; e += BS_THREADS) {
            int t = e / P, i = e % P;
            store(i, t, X[(size_t)ord[t] * L + Q + i]);
        }
        return;
    }
    const double smax = sig[ord[0]];
    const double cutoff = smax * 2.220446049250313e-16 * (P > Q ? P : Q);
    for (int e = tid; e < keep * P; e += BS_THREADS) {
        int t = e / P, i = e % P;
        double s = sig[ord[t]];
        store(i, t, s > cutoff ? CX::scale(X[(size_t)ord[t] * L + i], 1.0 / s) : CX::zero());
    }
    __syncthreads();
    int first_null = keep;
    for (int t = 0; t < keep; ++t)
        if (!(sig[ord[t]] > cutoff)) { first_null = t; break; }
    if (first_null >= keep) return;
    // completion: Gram-Schmidt (twice) of unit vectors against the columns written so far
    auto load_out = [&](int i, int t) {
        cplx v = out[(long long)i * a.so_i + (long long)t * a.so_t];
        return a.conj_out ? CX::conj(v) : v;
    };
    int trial = 0;
    for (int t = first_null; t < keep; ++t) {
        bool accepted = false;
        while (!accepted && trial < P) {
            for (int i = tid; i < P; i += BS_THREADS) u[i] = (i == trial) ? CX::one() : CX::zero();
            __syncthreads();
            for (int pass = 0; pass < 2; ++pass)
                for (int c = 0; c < t; ++c) {
                    cplx acc = CX::zero();
                    for (int i = tid; i < P; i += BS_THREADS) CX::fma(acc, CX::conj(load_out(i, c)), u[i]);
                    acc = block_sum<cplx>(acc, red);
                    __syncthreads();
                    for (int i = tid; i < P; i += BS_THREADS) u[i] = CX::sub(u[i], CX::mul(load_out(i, c), acc));
                    __syncthreads();
                }
            double s2 = 0.0;
            for (int i = tid; i < P; i += BS_THREADS) s2 += CX::abs2(u[i]);
            s2 = block_sum<double>(s2, dred);
            if (tid == 0) s_nrm = sqrt(s2);
            __syncthreads();
            const double nrm = s_nrm;
            if (nrm > 0.5) {
                for (int i = tid; i < P; i += BS_THREADS) store(i, t, CX::scale(u[i], 1.0 / nrm));
                accepted = true;
            }
            trial++;
            __syncthreads();
        }
    }
}

extern "C" int sktt_batch_svd_left(sktt_ctx* ctx, int64_t batch, int64_t P, int64_t Q, int64_t keep, const void* in,
                                   int64_t in_stride, sktt_idx2 fi, sktt_idx2 fj, int conj_in, void* out, int64_t out_stride,
                                   int64_t so_i, int64_t so_t, int conj_out) {
    if (!ctx || !in || !out) return SKTT_ERR_ARG;
    if (batch < 1 || P < 1 || Q < 1 || keep < 1 || keep > (P < Q ? P : Q))
        return sktt_fail(ctx, SKTT_ERR_ARG, "batch_svd_left: bad extents");
    const bool wide = P < Q;
    const long long nc = wide ? P : Q, L = wide ? Q + P : P;
    size_t smem = (size_t)nc * L * sizeof(cplx) + (size_t)nc * (sizeof(double) + sizeof(int)) + (size_t)P * sizeof(cplx) + 64;
    if (smem > 200 * 1024) return sktt_fail(ctx, SKTT_ERR_ARG, "batch_svd_left: block exceeds shared memory");
    SKTT_ONCE_PER_DEVICE(ctx);
    if (!configured) {
        SKTT_CUDA(ctx, cudaFuncSetAttribute(bsvd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    BsvdArgs a;
    a.P = (int)P; a.Q = (int)Q; a.keep = (int)keep; a.conj_in = conj_in; a.conj_out = conj_out;
    a.in = (const cplx*)in; a.in_stride = in_stride;
    a.fi = from_abi(fi); a.fj = from_abi(fj);
    a.out = (cplx*)out; a.out_stride = out_stride; a.so_i = so_i; a.so_t = so_t;
    a.sweeps = nullptr;
    bsvd_kernel<<<(unsigned)batch, BS_THREADS, smem, ctx->stream>>>(a);
    SKTT_LAUNCH_CHECK(ctx);
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// E = exp(c * H) for one small dense complex matrix (m <= 64), one CTA: scaling and squaring around a Taylor polynomial of
// degree 18 (Horner form) on ||c H||_1 / 2^s <= 1/2 (truncation error ~ 0.5^19 / 19! relative, far below fp64 rounding).
// The projected problems of the exponential integrators (ode.tdvp*: scipy.sparse.linalg.expm_multiply of the micro matrix,
// ode.py:1437-1508; local_krylov, ode.py:1689-1757) are of this size.
#define EX_THREADS 256
#define EX_MAX_M 64
__global__ void __launch_bounds__(EX_THREADS) expm_small_kernel(int m, const cplx* __restrict__ H, cplx cc, cplx* __restrict__ E) {
    extern __shared__ unsigned char smem_raw[];
    cplx* A = (cplx*)smem_raw;        // [m][m]  c H / 2^s
    cplx* X = A + m * m;              // Horner accumulator / squaring result
    cplx* W = X + m * m;              // product scratch
    __shared__ double colsum[EX_MAX_M];
    __shared__ int s_scale;
    const int tid = threadIdx.x, mm = m * m;
    for (int e = tid; e < mm; e += EX_THREADS) A[e] = CX::mul(cc, H[e]);
    __syncthreads();
    for (int j = tid; j < m; j += EX_THREADS) {
        double a = 0.0;
        for (int i = 0; i < m; ++i) a += sqrt(CX::abs2(A[i * m + j]));
        colsum[j] = a;
    }
    __syncthreads();
    if (tid == 0) {
        double nrm = 0.0;
        for (int j = 0; j < m; ++j) nrm = colsum[j] > nrm ? colsum[j] : nrm;
        int s = 0;
        while (nrm > 0.5 && s < 60) { nrm *= 0.5; ++s; }
        s_scale = s;
    }
    __syncthreads();
    const int s = s_scale;
    const double sc = ldexp(1.0, -s);
    for (int e = tid; e < mm; e += EX_THREADS) {
        A[e] = CX::scale(A[e], sc);
        X[e] = (e / m == e % m) ? CX::one() : CX::zero();
    }
    __syncthreads();
    for (int k = 18; k >= 1; --k) {                           // X <- I + A X / k
        const double invk = 1.0 / k;
        for (int e = tid; e < mm; e += EX_THREADS) {
            const int i = e / m, j = e % m;
            cplx acc = CX::zero();
            for (int t = 0; t < m; ++t) CX::fma(acc, A[i * m + t], X[t * m + j]);
            acc = CX::scale(acc, invk);
            if (i == j) acc.re += 1.0;
            W[e] = acc;
        }
        __syncthreads();
        for (int e = tid; e < mm; e += EX_THREADS) X[e] = W[e];
        __syncthreads();
    }
    for (int q = 0; q < s; ++q) {
        for (int e = tid; e < mm; e += EX_THREADS) {
            const int i = e / m, j = e % m;
            cplx acc = CX::zero();
            for (int t = 0; t < m; ++t) CX::fma(acc, X[i * m + t], X[t * m + j]);
            W[e] = acc;
        }
        __syncthreads();
        for (int e = tid; e < mm; e += EX_THREADS) X[e] = W[e];
        __syncthreads();
    }
    for (int e = tid; e < mm; e += EX_THREADS) E[e] = X[e];
}

// E (m x m complex128, row-major) = exp((c_re + i c_im) * H), H complex128 row-major, m <= 64.
extern "C" int sktt_expm_small(sktt_ctx* ctx, int64_t m, const void* H, double c_re, double c_im, void* E) {
    if (!ctx || !H || !E) return SKTT_ERR_ARG;
    if (m < 1 || m > EX_MAX_M) return sktt_fail(ctx, SKTT_ERR_ARG, "expm_small: m out of range (1..64)");
    const size_t smem = (size_t)3 * m * m * sizeof(cplx);
    SKTT_ONCE_PER_DEVICE(ctx);
    if (!configured) {
        SKTT_CUDA(ctx, cudaFuncSetAttribute(expm_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    expm_small_kernel<<<1, EX_THREADS, smem, ctx->stream>>>((int)m, (const cplx*)H, make_cplx(c_re, c_im), (cplx*)E);
    SKTT_LAUNCH_CHECK(ctx);
    return 0;
}
