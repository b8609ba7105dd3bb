// stack_nat.cu -- the interface-stack update of the ALS sweep (sle.py:217-219 left, :274-276 right through the mirrored
// cores) at the bench shape (solution ranks 64, operator ranks 3, fp64) as ONE cooperative launch that reads its three
// operands in the layout the caller holds them in -- no image build, no tiling pass, no memset in front of it:
//
//   new[a2,b2,c2] = sum_{a,b,c,n,m} L[a,b,c] x[a,n,a2] A[b,m,n,b2] x[c,m,c2]
//
//   phase 1 (one tile per CTA: column-mode index n x 96 rows (b,c)):  T1[(b,c),n,a2] = sum_a L[a,(b,c)] x[a,n,a2]
//   grid barrier
//   phase 2 (one tile per CTA: row-side rank c x 32 row-mode indices m):
//              T2[m,(a2,b2)] = sum_{b,n} A[b,m,n,b2] T1[b,c,n,a2]       (stays in shared memory)
//              P_tile[(a2,b2),c2] = sum_{m in tile} T2[m,(a2,b2)] x[c,m,c2]
//   grid barrier
//   phase 3: new = sum of the tile partials in a fixed order (bit-reproducible).
//
// Operand staging without re-laid images: every operand tile is a set of ROWS that are contiguous in the natural layout
// (96 doubles of a row of L, 64 doubles of a row of x, 8 column-mode indices x 3 operator ranks = 24 doubles of a row of
// A).  Rows that short are moved by the producer warp with 16-byte asynchronous copies (cp.async.cg, SASS LDGSTS.128: 8
// cycles per warp-wide instruction; one TMA bulk copy per row measured 20 ns EACH at the copy engine -- 792 of them per
// phase-2 tile made that phase 30 us) that complete on the same mbarriers as the TMA bulk copies of the T1 blocks (which
// are contiguous: one copy per block), into padded shared-memory rows whose pitch keeps the 64-bit fragment loads of a
// half-warp on 16 distinct bank pairs:
// pitch = 4 (mod 16) doubles for unit-stride operands, pitch = 12 (mod 16) for the operator rows whose entries of one
// rank index lie 3 doubles apart.  The mirrored update (right stack) reads the same memory with the roles of the two rank
// indices of A exchanged and the solution core transposed; both are addressing modes of the fragment loads, not copies.
// Zero (b, b2) blocks of the operator core (SLIM / MPO operators are block-sparse in their rank indices) are found by the
// kernel itself while phase 1 runs (one OR per warp into a mask word that the last phase clears again) and skipped in
// the second contraction.
#include <cooperative_groups.h>

#include "common.cuh"
#include "fused_common.cuh"
namespace cg = cooperative_groups;

namespace {

constexpr int NR = 64;                    // solution ranks on both sides
constexpr int NB = 3;                     // operator ranks on both sides
constexpr int N1_BM = 96, N1_LDA = N1_BM + 4, N1_LDB = NR + 4, N1_GROUPS = 4;
constexpr int N2_MB = 32;                 // row-mode indices per phase-2 tile
constexpr int N2_KC = 8;                  // column-mode indices per ring slot
constexpr int N2_AP = N2_KC * NB + 4;     // operator row pitch: 28 = 12 (mod 16)
constexpr int N2_ACH = NB * N2_MB * N2_AP;          // operator rows of one slot: [natural b][mm][AP]
constexpr int N2_BP = NR + 4;
constexpr int N2_BCH = NB * N2_KC * N2_BP;          // T1 rows of one slot:       [b][k][BP]
constexpr int N2_SLOT = N2_ACH + N2_BCH;
constexpr int N2_STAGES = 4;
constexpr int N2_LDT = NR * NB + 4;       // T2 in shared memory [MB][LDT]
constexpr int N2_XP = N2_MB + 4;          // transposed conj-side rows (mirror): [c2][XP]
constexpr size_t NAT_PHASE_BYTES = ((size_t)N2_STAGES * N2_SLOT + (size_t)N2_MB * N2_LDT) * sizeof(double);
static_assert((size_t)NR * (N1_LDA + N1_LDB) * sizeof(double) <= NAT_PHASE_BYTES, "phase 1 fits the phase-2 plan");
static_assert(N2_AP % 16 == 12 && N2_BP % 16 == 4 && N2_LDT % 16 == 4 && N2_XP % 16 == 4 && N1_LDA % 16 == 4, "pitches");
static_assert((N2_AP * 8) % 16 == 0 && (N2_BP * 8) % 16 == 0 && (N1_LDA * 8) % 16 == 0, "bulk copy destinations");

// 16-byte asynchronous copy global -> shared (L2 only) and the arrival of a thread's earlier copies on an mbarrier
__device__ __forceinline__ void ldgsts16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void ldgsts_arrive(unsigned long long* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

struct StackNatParams {
    const double* L;          // old stack [64][3][64]
    const double* x;          // solution core [64][n][64] (mirror: read transposed)
    const double* A;          // operator core [3][m][n][3] (mirror: rank indices exchanged)
    int mtot, ntot;
    double* T1p;              // [192][n][68]
    double* part;             // [tiles][192][64]
    double* out;              // [64][3][64]
    unsigned long long* mask; // one word, zero between launches
    unsigned long long* stamps;   // optional %globaltimer stamps of CTA 0
};

// ---------------------------------------------------------------------------------------------------- phase 1
template <bool SWAP>
__device__ void n_s1_tile(unsigned char* smem_raw, unsigned long long* full, const double* __restrict__ L,
                          const double* __restrict__ x, double* __restrict__ T1p, int ntot, int nn, int mt, bool first) {
    double* As = reinterpret_cast<double*>(smem_raw);     // [a][N1_LDA]: L[a, mt * 96 + .]
    double* Bs = As + (size_t)NR * N1_LDA;                 // [a][N1_LDB]: x[a, nn, .]   (mirror: [a2][N1_LDB]: x[a2, nn, .])
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __syncthreads();
    if (tid == 0) {
        for (int g = 0; g < N1_GROUPS; ++g) {
            if (!first) mbar_inval(full + g);
            mbar_init(full + g, 32);                       // one arrival per producer lane, after its copies of the group
        }
        mbar_fence_init();
    }
    __syncthreads();
    if (warp == CONSUMER_WARPS) {
        // rows of 96 (L) and 64 (x) doubles as 16-byte pieces: four lanes share a row of x, eight lanes a half row of L
        const double* lsrc = L + mt * N1_BM + 2 * (lane & 7);
        const double* xsrc = x + (size_t)nn * NR + 2 * (lane & 7);
        if (SWAP) {                                        // every row of the transposed core is needed from the first k on
#pragma unroll 4
            for (int i = 0; i < 16; ++i) {
                const int row = 4 * i + (lane >> 3);
#pragma unroll
                for (int h = 0; h < 4; ++h)
                    ldgsts16(Bs + (size_t)row * N1_LDB + 2 * (lane & 7) + 16 * h, xsrc + (size_t)row * ntot * NR + 16 * h);
            }
        }
        for (int g = 0; g < N1_GROUPS; ++g) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int a = 16 * g + 4 * i + (lane >> 3);
#pragma unroll
                for (int h = 0; h < 6; ++h)
                    ldgsts16(As + (size_t)a * N1_LDA + 2 * (lane & 7) + 16 * h, lsrc + (size_t)a * (NB * NR) + 16 * h);
                if (!SWAP) {
#pragma unroll
                    for (int h = 0; h < 4; ++h)
                        ldgsts16(Bs + (size_t)a * N1_LDB + 2 * (lane & 7) + 16 * h, xsrc + (size_t)a * ntot * NR + 16 * h);
                }
            }
            ldgsts_arrive(full + g);
        }
        return;
    }
    const int tile = warp & 7, khalf = warp >> 3;
    const int wm0 = (tile & 1) * 48, wn0 = (tile >> 1) * 16;
    const int fr = lane >> 2, fk = lane & 3;
    double acc[6][2][2];
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int g = 0; g < N1_GROUPS; ++g) {
        mbar_wait(full + g, 0);
#pragma unroll
        for (int st = 0; st < 2; ++st) {
            const int kk = 16 * g + 8 * khalf + 4 * st;
            const double* as = As + (kk + fk) * N1_LDA + wm0 + fr;
            double af[6], bf[2];
#pragma unroll
            for (int i = 0; i < 6; ++i) af[i] = as[8 * i];
#pragma unroll
            for (int j = 0; j < 2; ++j)
                bf[j] = SWAP ? Bs[(wn0 + fr + 8 * j) * N1_LDB + kk + fk] : Bs[(kk + fk) * N1_LDB + wn0 + fr + 8 * j];
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    consumer_bar_sync();
    double* red = As + (size_t)tile * (48 * 16);
    if (khalf == 1) {
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
                *reinterpret_cast<double2*>(red + (8 * i + fr) * 16 + 8 * j + 2 * fk) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
    consumer_bar_sync();
    if (khalf == 0) {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int m = mt * N1_BM + wm0 + 8 * i + fr;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const double2 o = *reinterpret_cast<const double2*>(red + (8 * i + fr) * 16 + 8 * j + 2 * fk);
                double* dst = T1p + ((size_t)m * ntot + nn) * N2_BP + wn0 + 8 * j + 2 * fk;
                *reinterpret_cast<double2*>(dst) = make_double2(acc[i][j][0] + o.x, acc[i][j][1] + o.y);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------- phase 2
// One contraction block b of one ring slot for a COMPILE-TIME mask QM of the non-zero output blocks q (zero blocks cost
// no issue slots).  The entry A~[b, m, n, q] of the (possibly mirrored) operator core sits at row (blk * 32 + m), column
// 3 n + il of the slot with (blk, il) = (b, q), mirrored: (q, b).
template <bool SWAP, unsigned QM>
__device__ __forceinline__ void n_s2_block(double (&acc2)[NB][2][2][2], const double* __restrict__ ach,
                                           const double* __restrict__ bch, int b, int row0, int kcol, int col0) {
    double bf[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) bf[j] = bch[(b * N2_KC + kcol) * N2_BP + col0 + 8 * j];
#pragma unroll
    for (int q = 0; q < NB; ++q) {
        if ((QM >> q) & 1u) {
            const int blk = SWAP ? q : b, il = SWAP ? b : q;
            double af[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) af[i] = ach[(blk * N2_MB + row0 + 8 * i) * N2_AP + kcol * NB + il];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma(acc2[q][i][j][0], acc2[q][i][j][1], af[i], bf[j]);
        }
    }
}

template <bool SWAP>
__device__ void n_s2x_tile(unsigned char* smem_raw, unsigned long long* full, const double* __restrict__ T1p,
                           const double* __restrict__ A, const double* __restrict__ x, double* __restrict__ part, int mtot,
                           int ntot, unsigned blockmask, int c, int mblk, bool first) {
    double* ring = reinterpret_cast<double*>(smem_raw);
    double* T2s = ring + (size_t)N2_STAGES * N2_SLOT;
    unsigned long long* empty = full + N2_STAGES;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = mblk * N2_MB;
    const int nslots = ntot / N2_KC;
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < N2_STAGES; ++s) {
            if (!first) {
                mbar_inval(full + s);
                mbar_inval(empty + s);
            }
            mbar_init(full + s, 33);                       // 32 producer lanes (operator rows) + the expect_tx of the T1 blocks
            mbar_init(empty + s, CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();
    if (warp == CONSUMER_WARPS) {
        fence_proxy_async();
        // operator rows of a slot: 96 rows (natural b, mm) of 24 doubles = 12 pieces of 16 bytes; four lanes share a row
        // (three pieces each), the warp covers eight rows per step
        const double* asrc = A + ((size_t)(m0 + (lane >> 2)) * ntot) * NB + 6 * (lane & 3);
        const size_t bstride = (size_t)mtot * ntot * NB, rstride = (size_t)8 * ntot * NB;
        for (int t = 0; t < nslots; ++t) {
            const int s = t % N2_STAGES;
            if (t >= N2_STAGES) mbar_wait(empty + s, ((t / N2_STAGES) & 1) ^ 1);
            double* slot = ring + (size_t)s * N2_SLOT;
            if (lane == 0) mbar_expect_tx(full + s, (unsigned)(NB * N2_KC * N2_BP * sizeof(double)));
            if (lane < NB)
                bulk_g2s(slot + N2_ACH + (size_t)lane * N2_KC * N2_BP,
                         T1p + (((size_t)lane * NR + c) * ntot + (size_t)t * N2_KC) * N2_BP, N2_KC * N2_BP * 8, full + s);
            double* adst = slot + (size_t)(lane >> 2) * N2_AP + 6 * (lane & 3);
            const double* at = asrc + (size_t)t * N2_KC * NB;
#pragma unroll
            for (int k = 0; k < NB; ++k)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double* src = at + k * bstride + i * rstride;
                    double* dst = adst + (size_t)(k * N2_MB + 8 * i) * N2_AP;
                    ldgsts16(dst, src);
                    ldgsts16(dst + 2, src + 2);
                    ldgsts16(dst + 4, src + 4);
                }
            ldgsts_arrive(full + s);
        }
        return;
    }
    const int tile = warp & 7, khalf = warp >> 3;
    const int wm0 = (tile & 1) * 16, wn0 = (tile >> 1) * 16;
    const int fr = lane >> 2, fk = lane & 3;
    // rows of the conj-side core this tile contracts with at the end: x[c, m0 + mm, c2] (mirror: x[c2, m0 + mm, c]); the
    // loads are issued now and land in shared memory once the ring is free
    double xr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int e = tid + 512 * k;
        xr[k] = SWAP ? x[((size_t)(e >> 5) * mtot + m0 + (e & 31)) * NR + c] : x[((size_t)c * mtot + m0 + (e >> 6)) * NR + (e & 63)];
    }
    double acc2[NB][2][2][2];
#pragma unroll
    for (int q = 0; q < NB; ++q)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) acc2[q][i][j][0] = acc2[q][i][j][1] = 0.0;
    const int kcol = 4 * khalf + fk;
    for (int t = 0; t < nslots; ++t) {
        const int s = t % N2_STAGES;
        mbar_wait(full + s, (t / N2_STAGES) & 1);
        const double* ach = ring + (size_t)s * N2_SLOT;
        const double* bch = ach + N2_ACH;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            switch ((blockmask >> (b * NB)) & 7u) {                    // warp-uniform
                case 1: n_s2_block<SWAP, 1>(acc2, ach, bch, b, wm0 + fr, kcol, wn0 + fr); break;
                case 2: n_s2_block<SWAP, 2>(acc2, ach, bch, b, wm0 + fr, kcol, wn0 + fr); break;
                case 3: n_s2_block<SWAP, 3>(acc2, ach, bch, b, wm0 + fr, kcol, wn0 + fr); break;
                case 4: n_s2_block<SWAP, 4>(acc2, ach, bch, b, wm0 + fr, kcol, wn0 + fr); break;
                case 5: n_s2_block<SWAP, 5>(acc2, ach, bch, b, wm0 + fr, kcol, wn0 + fr); break;
                case 6: n_s2_block<SWAP, 6>(acc2, ach, bch, b, wm0 + fr, kcol, wn0 + fr); break;
                case 7: n_s2_block<SWAP, 7>(acc2, ach, bch, b, wm0 + fr, kcol, wn0 + fr); break;
                default: break;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
    }
    consumer_bar_sync();                                   // every warp is done with the ring
    double* Xs = ring;                                     // [mm][N2_BP]   (mirror: [c2][N2_XP])
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int e = tid + 512 * k;
        if (SWAP) Xs[(e >> 5) * N2_XP + (e & 31)] = xr[k];
        else Xs[(e >> 6) * N2_BP + (e & 63)] = xr[k];
    }
    // T2[m, a2, b2] = sum of the two K-halves -> T2s[m][a2 * 3 + b2]
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (khalf == h) {
#pragma unroll
            for (int q = 0; q < NB; ++q)
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int mm = wm0 + 8 * i + fr, a2 = wn0 + 8 * j + 2 * fk;
                        double* d0 = T2s + (size_t)mm * N2_LDT + (size_t)a2 * NB + q;
                        double* d1 = d0 + NB;
                        if (h == 0) {
                            *d0 = acc2[q][i][j][0];
                            *d1 = acc2[q][i][j][1];
                        } else {
                            *d0 += acc2[q][i][j][0];
                            *d1 += acc2[q][i][j][1];
                        }
                    }
        }
        consumer_bar_sync();
    }
    // P[(a2,b2), c2] = sum_m T2s[m][(a2,b2)] x[c, m, c2]: warp (wr, wc) owns rows 24 wr .. +24, columns 32 wc .. +32
    const int wr = warp >> 1, wc = warp & 1;
    double acc[3][4][2];
#pragma unroll
    for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[t][u][0] = acc[t][u][1] = 0.0;
#pragma unroll
    for (int k0 = 0; k0 < N2_MB; k0 += 4) {
        double af[3], bf[4];
#pragma unroll
        for (int t = 0; t < 3; ++t) af[t] = T2s[(size_t)(k0 + fk) * N2_LDT + 24 * wr + 8 * t + fr];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            bf[u] = SWAP ? Xs[(32 * wc + 8 * u + fr) * N2_XP + k0 + fk] : Xs[(k0 + fk) * N2_BP + 32 * wc + 8 * u + fr];
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
            for (int u = 0; u < 4; ++u) dmma(acc[t][u][0], acc[t][u][1], af[t], bf[u]);
    }
    double* dst = part + (size_t)(mblk * NR + c) * (NR * NB) * NR;
#pragma unroll
    for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int u = 0; u < 4; ++u)
            *reinterpret_cast<double2*>(dst + (size_t)(24 * wr + 8 * t + fr) * NR + 32 * wc + 8 * u + 2 * fk) =
                make_double2(acc[t][u][0], acc[t][u][1]);
}

// ---------------------------------------------------------------------------------------------------- the kernel
template <bool SWAP>
__global__ void __launch_bounds__(THREADS) stack_nat_kernel(StackNatParams a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cg::grid_group grid = cg::this_grid();
    const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long* bars1 = reinterpret_cast<unsigned long long*>(smem_raw + NAT_PHASE_BYTES);
    unsigned long long* bars2 = bars1 + N1_GROUPS;
    const int tiles1 = a.ntot * 2, tiles2 = NR * (a.mtot / N2_MB);
    int nstamp = 0;
    auto stamp = [&]() {
        if (a.stamps && cta == 0 && tid == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            a.stamps[1 + nstamp++] = t;
            a.stamps[0] = (unsigned long long)nstamp;
        }
    };
    auto gsync = [&]() {
        __syncthreads();
        if (tid == 0) {
            fence_proxy_async();
            __threadfence();
        }
        grid.sync();
    };
    stamp();
    // which (b, b2) blocks of the (possibly mirrored) operator core hold a non-zero entry: every thread looks at its share
    // of the core while phase 1 runs; the bits meet in one word behind the grid barrier
    unsigned bits = 0;
    {
        const long long per_b = (long long)a.mtot * a.ntot * NB, total = NB * per_b;
        for (long long e = (long long)cta * THREADS + tid; e < total; e += (long long)G * THREADS) {
            const int bn = (int)(e / per_b), qn = (int)(e % NB);
            if (a.A[e] != 0.0) bits |= 1u << (SWAP ? qn * NB + bn : bn * NB + qn);
        }
    }
    bool first1 = true, first2 = true;
    for (int t = cta; t < tiles1; t += G) {
        n_s1_tile<SWAP>(smem_raw, bars1, a.L, a.x, a.T1p, a.ntot, t % a.ntot, t / a.ntot, first1);
        first1 = false;
    }
    __syncwarp();
    bits = __reduce_or_sync(0xffffffffu, bits);
    if (lane == 0 && bits) atomicOr(a.mask, (unsigned long long)bits);
    stamp();
    gsync();
    stamp();
    const unsigned blockmask = (unsigned)__ldcg(a.mask);
    for (int t = cta; t < tiles2; t += G) {
        n_s2x_tile<SWAP>(smem_raw, bars2, a.T1p, a.A, a.x, a.part, a.mtot, a.ntot, blockmask, t % NR, t / NR, first2);
        first2 = false;
    }
    stamp();
    gsync();
    stamp();
    if (cta == 0 && tid == 0) *a.mask = 0ull;              // everybody has read it: the next launch finds it clear
    // out[e] = sum over the tile partials in a fixed order: a CTA takes 32 consecutive entries at a time, lanes run along the
    // entries (coalesced rows of the partials), sixteen warps split the tiles with their loads in flight at once, warp 0
    // adds the sixteen sums in warp order
    const int E = NR * NB * NR;
    __shared__ double red[16][32];
    for (int c = cta; c < E / 32; c += G) {
        const int e = c * 32 + lane;
        if (warp < 16) {
            double v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int t = warp + 16 * k;
                v[k] = t < tiles2 ? __ldcg(a.part + (size_t)t * E + e) : 0.0;
            }
            double sacc = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
            for (int t = warp + 128; t < tiles2; t += 16) sacc += __ldcg(a.part + (size_t)t * E + e);
            red[warp][lane] = sacc;
        }
        __syncthreads();
        if (warp == 0) {
            double sacc = red[0][lane];
#pragma unroll
            for (int w = 1; w < 16; ++w) sacc += red[w][lane];
            a.out[e] = sacc;
        }
        __syncthreads();
    }
    stamp();
}

}  // namespace

// Shapes of the natural-layout kernel: both solution ranks 64, both operator ranks 3, square mode of a multiple of 32,
// operands 16-byte aligned (TMA bulk copies).  Everything else runs the image-based kernel of fused.cu or the generic chain.
bool sktt_stack_nat_supported(const sktt_ctx* ctx, int dtype, long long rin, long long Rin, long long m, long long n,
                              long long rout, long long Rout, const void* stack, const void* x, const void* A) {
    if (dtype != SKTT_F64 || ctx->gemm_mode == 1 || (ctx->debug & 64)) return false;
    if (rin != NR || rout != NR || Rin != NB || Rout != NB) return false;
    if (m != n || m % N2_MB != 0 || m > 4096) return false;
    return (((uintptr_t)stack | (uintptr_t)x | (uintptr_t)A) & 15u) == 0;
}

#define STACK_NAT_MASK_OFF 3584          // byte offsets in the scalar area of the context scratch (zeroed when it is allocated)
#define STACK_NAT_STAMP_OFF 3600

// out = new stack [64][3][64].  T1p: 192 * n * 68 doubles, part: 64 * (m / 32) * 12288 doubles.
int sktt_stack_nat_update(sktt_ctx* ctx, long long m, long long n, const double* stack, const double* x, const double* A,
                          double* out, double* T1p, double* part, int mirror) {
    StackNatParams a;
    a.L = stack;
    a.x = x;
    a.A = A;
    a.mtot = (int)m;
    a.ntot = (int)n;
    a.T1p = T1p;
    a.part = part;
    a.out = out;
    a.mask = (unsigned long long*)((char*)ctx->scratch + STACK_NAT_MASK_OFF);
    a.stamps = (ctx->debug & 1) ? (unsigned long long*)((char*)ctx->scratch + STACK_NAT_STAMP_OFF) : nullptr;
    const size_t smem = NAT_PHASE_BYTES + 128;
    SKTT_ONCE_PER_DEVICE(ctx);
    if (!configured) {
        SKTT_CUDA(ctx, cudaFuncSetAttribute(stack_nat_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SKTT_CUDA(ctx, cudaFuncSetAttribute(stack_nat_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    void* args[] = {&a};
    const void* fn = mirror ? (const void*)stack_nat_kernel<true> : (const void*)stack_nat_kernel<false>;
    SKTT_CUDA(ctx, cudaLaunchCooperativeKernel(fn, dim3(ctx->sm_count), dim3(THREADS), args, smem, ctx->stream));
    ctx->launches++;
    return 0;
}
