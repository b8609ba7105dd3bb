// hess_eig.cuh -- eigen-decomposition of a small complex upper Hessenberg matrix by ONE warp (single-shift QR iteration
// with Wilkinson shifts, rotations applied lane-parallel, eigenvectors by back substitution).  Shared by the host-driven
// shift-invert Arnoldi (eig.cu) and the fused batched one (batch.cu).
#pragma once
#include "common.cuh"
#include "blas1.cuh"

__device__ __forceinline__ double cabs_(cplx a) { return hypot(a.re, a.im); }

__device__ __forceinline__ cplx csqrt_(cplx z) {
    double r = hypot(z.re, z.im);
    if (r == 0.0) return make_cplx(0.0, 0.0);
    double sr = sqrt(0.5 * (r + fabs(z.re)));
    double si = 0.5 * z.im / sr;
    if (z.re >= 0.0) return make_cplx(sr, si);
    return make_cplx(fabs(si), z.im >= 0.0 ? sr : -sr);
}

// Hin: m x m upper Hessenberg, row-major complex (global or shared).  Outputs: theta[m], Yout[m][m] (column i = unit-norm
// eigenvector i of Hin).  H, Z, X: three m x m complex work arrays (shared memory).  Must be called by all 32 lanes of one warp.
static __device__ __noinline__ void hess_eig_warp(int m, const cplx* __restrict__ Hin, cplx* __restrict__ theta,
                                           cplx* __restrict__ Yout, int* info, cplx* H, cplx* Z, cplx* X) {
    const int lane = threadIdx.x & 31;
    typedef Num<cplx> C;
    for (int e = lane; e < m * m; e += 32) {
        int r = e / m, c = e % m;
        H[e] = (r <= c + 1) ? Hin[e] : C::zero();
        Z[e] = (r == c) ? C::one() : C::zero();
    }
    __syncwarp();
    double hnorm = 0.0;
    for (int e = lane; e < m * m; e += 32) hnorm += C::abs2(H[e]);
    hnorm = sqrt(warp_sum<double>(hnorm));
    const double eps = 2.220446049250313e-16;
    const double tiny = hnorm > 0.0 ? hnorm * eps : eps;
    int hi = m - 1, iter = 0, total_iter = 0, fail = 0;
    while (hi > 0) {
        // deflation scan (uniform across lanes: every lane evaluates the same scalars)
        int l = hi;
        while (l > 0) {
            double sub = cabs_(H[l * m + l - 1]);
            double dsum = cabs_(H[(l - 1) * m + l - 1]) + cabs_(H[l * m + l]);
            if (dsum == 0.0) dsum = hnorm;
            if (sub <= eps * dsum || sub <= tiny * 1e-3) break;
            --l;
        }
        if (l > 0 && lane == 0) H[l * m + l - 1] = C::zero();
        __syncwarp();
        if (l == hi) {
            --hi;
            iter = 0;
            continue;
        }
        if (++total_iter > 60 * m) { fail = 1; break; }
        ++iter;
        // Wilkinson shift from the trailing 2x2 of the active block
        cplx a = H[(hi - 1) * m + hi - 1], b = H[(hi - 1) * m + hi], c = H[hi * m + hi - 1], d = H[hi * m + hi];
        cplx mu;
        if (iter % 11 == 10) {
            mu = C::add(d, make_cplx(cabs_(c) * 0.75, cabs_(c) * -0.4375));  // exceptional shift
        } else {
            cplx tr2 = C::scale(C::sub(a, d), 0.5);
            cplx disc = csqrt_(C::add(C::mul(tr2, tr2), C::mul(b, c)));
            cplx e1 = C::add(C::add(d, tr2), disc), e2 = C::sub(C::add(d, tr2), disc);
            mu = cabs_(C::sub(e1, d)) <= cabs_(C::sub(e2, d)) ? e1 : e2;
        }
        cplx x = C::sub(H[l * m + l], mu), y = H[(l + 1) * m + l];
        for (int k = l; k < hi; ++k) {
            // G = [c s; -conj(s) c], G [x; y] = [rho; 0]
            double nx = cabs_(x), ny = cabs_(y), nrm = hypot(nx, ny);
            double cg;
            cplx sg;
            if (nrm == 0.0) {
                cg = 1.0;
                sg = C::zero();
            } else if (nx == 0.0) {
                cg = 0.0;
                sg = C::scale(C::conj(y), 1.0 / ny);
            } else {
                cg = nx / nrm;
                sg = C::scale(C::mul(C::scale(x, 1.0 / nx), C::conj(y)), 1.0 / nrm);
            }
            const cplx sgc = C::conj(sg);
            // rows k, k+1 of H (columns from max(l, k-1) to m-1)
            const int c0 = k > l ? k - 1 : l;
            for (int cc = c0 + lane; cc < m; cc += 32) {
                cplx h0 = H[k * m + cc], h1 = H[(k + 1) * m + cc];
                H[k * m + cc] = C::add(C::scale(h0, cg), C::mul(sg, h1));
                H[(k + 1) * m + cc] = C::sub(C::scale(h1, cg), C::mul(sgc, h0));
            }
            __syncwarp();
            // columns k, k+1 of H (rows 0 .. min(k+2, hi)) times G^H
            const int r1 = k + 2 < hi ? k + 2 : hi;
            for (int rr = lane; rr <= r1; rr += 32) {
                cplx h0 = H[rr * m + k], h1 = H[rr * m + k + 1];
                H[rr * m + k] = C::add(C::scale(h0, cg), C::mul(sgc, h1));
                H[rr * m + k + 1] = C::sub(C::scale(h1, cg), C::mul(sg, h0));
            }
            for (int rr = lane; rr < m; rr += 32) {
                cplx z0 = Z[rr * m + k], z1 = Z[rr * m + k + 1];
                Z[rr * m + k] = C::add(C::scale(z0, cg), C::mul(sgc, z1));
                Z[rr * m + k + 1] = C::sub(C::scale(z1, cg), C::mul(sg, z0));
            }
            __syncwarp();
            if (k < hi - 1) {
                x = H[(k + 1) * m + k];
                y = H[(k + 2) * m + k];
            }
        }
    }
    __syncwarp();
    // eigenvectors of the triangular factor: lane i solves (T - t_ii I) x = 0 with x_i = 1
    for (int i = lane; i < m; i += 32) {
        cplx tii = H[i * m + i];
        for (int r = m - 1; r > i; --r) X[r * m + i] = C::zero();
        X[i * m + i] = C::one();
        for (int r = i - 1; r >= 0; --r) {
            cplx s = C::zero();
            for (int c = r + 1; c <= i; ++c) C::fma(s, H[r * m + c], X[c * m + i]);
            cplx den = C::sub(H[r * m + r], tii);
            if (cabs_(den) < tiny) den = make_cplx(tiny, 0.0);
            X[r * m + i] = C::neg(C::div(s, den));
        }
        theta[i] = tii;
    }
    __syncwarp();
    // Y = Z X, unit-norm columns
    for (int i = lane; i < m; i += 32) {
        double nrm2 = 0.0;
        for (int r = 0; r < m; ++r) {
            cplx s = C::zero();
            for (int c = 0; c <= i; ++c) C::fma(s, Z[r * m + c], X[c * m + i]);
            Yout[r * m + i] = s;
            nrm2 += C::abs2(s);
        }
        double inv = nrm2 > 0.0 ? 1.0 / sqrt(nrm2) : 0.0;
        for (int r = 0; r < m; ++r) Yout[r * m + i] = C::scale(Yout[r * m + i], inv);
    }
    if (lane == 0 && info) *info = fail;
}


