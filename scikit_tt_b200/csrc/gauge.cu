// gauge.cu -- the two small products that move the gauge of the sweep by one core (f64).
//
// The reference discards the triangular factor of the QR / RQ of a solved core (scikit_tt/solvers/sle.py:525, :541).  The
// matrix-free micro solves use it as a warm start: with u = Q R (resp. u = R' Q') the current iterate of the sweep,
// expressed in the unknowns of the NEXT micro system, is R x_next (resp. x_prev R').  Both products have one long index
// (r n or n r' = 4096 at the bench shape) and two short ones (<= 64), a shape on which the general GEMM spends two to three
// launches of ~17 us; here each is one launch of a few microseconds.
//
//   factor:  C(i, j)   = sum_l A(l, i) B(l, j)       l < L (long),  i < ni, j < nj (short)
//   push  :  out(l, i) = sum_b Rm(i, b) X(l, b)      l < L,         i < ni, b < nb
//
// Every operand is addressed through (long stride, short stride), so the same kernels serve the forward half sweep (Q, u
// stored [l][i]) and the backward one (stored [i][l]).  CTA partials of `factor` are summed in CTA order by a second small
// kernel: bit-reproducible.
#include "common.cuh"
#include "blas1.cuh"

#define GAUGE_TL 32          // long indices per smem tile
#define GAUGE_LD 65

namespace {

__global__ void __launch_bounds__(256)
gauge_factor_kernel(int L, int ni, int nj, int chunk, const double* __restrict__ A, long long sla, long long sia,
                    const double* __restrict__ B, long long slb, long long sjb, double* __restrict__ C, long long sci,
                    long long scj, double* part) {
    __shared__ double As[GAUGE_TL][GAUGE_LD], Bs[GAUGE_TL][GAUGE_LD];
    const int cta = blockIdx.x, tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int l_begin = cta * chunk, l_end = min(L, l_begin + chunk);
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    for (int l0 = l_begin; l0 < l_end; l0 += GAUGE_TL) {
        for (int e = tid; e < GAUGE_TL * 64; e += 256) {
            int l, i;
            if (sla == 1) { l = e % GAUGE_TL; i = e / GAUGE_TL; } else { i = e % 64; l = e / 64; }
            As[l][i] = (l0 + l < l_end && i < ni) ? A[(long long)(l0 + l) * sla + (long long)i * sia] : 0.0;
            if (slb == 1) { l = e % GAUGE_TL; i = e / GAUGE_TL; } else { i = e % 64; l = e / 64; }
            Bs[l][i] = (l0 + l < l_end && i < nj) ? B[(long long)(l0 + l) * slb + (long long)i * sjb] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int l = 0; l < GAUGE_TL; ++l) {
            double av[4], bv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) av[a] = As[l][ty + 16 * a];
#pragma unroll
            for (int b = 0; b < 4; ++b) bv[b] = Bs[l][tx + 16 * b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
        }
        __syncthreads();
    }
    double* mine = part + (size_t)cta * 4096;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) mine[(ty + 16 * a) * 64 + tx + 16 * b] = acc[a][b];
}

// C(i, j) = sum over the CTA partials in CTA order: four threads per entry take every fourth partial (independent loads),
// combined by two shuffles -- the same order on every run
__global__ void __launch_bounds__(256)
gauge_reduce_kernel(int G, int ni, int nj, const double* __restrict__ part, double* __restrict__ C, long long sci,
                    long long scj) {
    const int e = blockIdx.x * 64 + (threadIdx.x >> 2), sub = threadIdx.x & 3;
    double s = 0.0;
#pragma unroll 8
    for (int g = sub; g < G; g += 4) s += __ldcg(part + (size_t)g * 4096 + e);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    const int i = e >> 6, j = e & 63;
    if (sub == 0 && i < ni && j < nj) C[(long long)i * sci + (long long)j * scj] = s;
}

#define GAUGE_PL 16          // long indices per CTA of the push kernel

__global__ void __launch_bounds__(256)
gauge_push_kernel(int L, int ni, int nb, const double* __restrict__ Rm, long long sri, long long srb,
                  const double* __restrict__ X, long long slx, long long sbx, double* __restrict__ out, long long slo,
                  long long sio) {
    __shared__ double Rs[64][GAUGE_LD];          // Rs[b][i]
    __shared__ double Xs[GAUGE_PL][GAUGE_LD];    // Xs[l][b], reused as the output tile Os[l][i]
    const int tid = threadIdx.x, tl = tid >> 4, ti = tid & 15;
    const int l0 = blockIdx.x * GAUGE_PL;
    for (int e = tid; e < 64 * 64; e += 256) {
        int i, b;
        if (sri == 1) { i = e % 64; b = e / 64; } else { b = e % 64; i = e / 64; }
        Rs[b][i] = (i < ni && b < nb) ? Rm[(long long)i * sri + (long long)b * srb] : 0.0;
    }
    for (int e = tid; e < GAUGE_PL * 64; e += 256) {
        int l, b;
        if (slx == 1) { l = e % GAUGE_PL; b = e / GAUGE_PL; } else { b = e % 64; l = e / 64; }
        Xs[l][b] = (l0 + l < L && b < nb) ? X[(long long)(l0 + l) * slx + (long long)b * sbx] : 0.0;
    }
    __syncthreads();
    double acc[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) acc[a] = 0.0;
#pragma unroll 8
    for (int b = 0; b < 64; ++b) {
        const double xv = Xs[tl][b];
#pragma unroll
        for (int a = 0; a < 4; ++a) acc[a] = fma(xv, Rs[b][ti + 16 * a], acc[a]);
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; ++a) Xs[tl][ti + 16 * a] = acc[a];
    __syncthreads();
    for (int e = tid; e < GAUGE_PL * 64; e += 256) {
        int l, i;
        if (slo == 1) { l = e % GAUGE_PL; i = e / GAUGE_PL; } else { i = e % 64; l = e / 64; }
        if (l0 + l < L && i < ni) out[(long long)(l0 + l) * slo + (long long)i * sio] = Xs[l][i];
    }
}

}  // namespace

extern "C" int sktt_gauge_factor(sktt_ctx* ctx, int dtype, int64_t L, int64_t ni, int64_t nj, const void* A, int64_t sla,
                                 int64_t sia, const void* B, int64_t slb, int64_t sjb, void* C, int64_t sci, int64_t scj) {
    if (!ctx || !A || !B || !C) return SKTT_ERR_ARG;
    if (dtype != SKTT_F64 || L < 1 || ni < 1 || nj < 1 || ni > 64 || nj > 64 || L > 0x7fffffff)
        return sktt_fail(ctx, SKTT_ERR_ARG, "gauge_factor: f64, short extents <= 64");
    int chunk = GAUGE_TL;                                      // one tile per CTA while the CTAs fit the SMs: shortest serial chain
    int G = (int)((L + chunk - 1) / chunk);
    if (G > ctx->sm_count) {
        chunk = (int)(((L + ctx->sm_count - 1) / ctx->sm_count + GAUGE_TL - 1) / GAUGE_TL * GAUGE_TL);
        G = (int)((L + chunk - 1) / chunk);
    }
    SKTT_TRY(sktt_scratch_reserve(ctx, SKTT_SCRATCH_BULK_OFF + (size_t)G * 4096 * sizeof(double)));
    double* part = (double*)((char*)ctx->scratch + SKTT_SCRATCH_BULK_OFF);
    gauge_factor_kernel<<<G, 256, 0, ctx->stream>>>((int)L, (int)ni, (int)nj, chunk, (const double*)A, sla, sia, (const double*)B,
                                                    slb, sjb, (double*)C, sci, scj, part);
    SKTT_LAUNCH_CHECK(ctx);
    gauge_reduce_kernel<<<64, 256, 0, ctx->stream>>>(G, (int)ni, (int)nj, part, (double*)C, sci, scj);
    SKTT_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" int sktt_gauge_push(sktt_ctx* ctx, int dtype, int64_t L, int64_t ni, int64_t nb, const void* Rm, int64_t sri,
                               int64_t srb, const void* X, int64_t slx, int64_t sbx, void* out, int64_t slo, int64_t sio) {
    if (!ctx || !Rm || !X || !out) return SKTT_ERR_ARG;
    if (dtype != SKTT_F64 || L < 1 || ni < 1 || nb < 1 || ni > 64 || nb > 64 || L > 0x7fffffff)
        return sktt_fail(ctx, SKTT_ERR_ARG, "gauge_push: f64, short extents <= 64");
    const int G = (int)((L + GAUGE_PL - 1) / GAUGE_PL);
    gauge_push_kernel<<<G, 256, 0, ctx->stream>>>((int)L, (int)ni, (int)nb, (const double*)Rm, sri, srb, (const double*)X, slx,
                                                  sbx, (double*)out, slo, sio);
    SKTT_LAUNCH_CHECK(ctx);
    return 0;
}
