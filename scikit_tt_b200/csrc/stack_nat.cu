// stack_nat.cu -- the interface-stack update of the ALS sweep (sle.py:217-219 left, :274-276 right through the mirrored
// cores) at the bench shape (solution ranks 64, operator ranks 3, mode size <= 64, fp64) as ONE cooperative launch that
// reads its three operands in the layout the caller holds them in -- no image build, no tiling pass, no memset:
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
// Roles: 16 consumer warps (LDS.64 + DMMA.8x8x4 only) and one producer warp per CTA, one CTA per SM.
//
// What the producer moves, and when (measured on the way here, profiles/README.md round 2):
//   * The 96 operator rows A[b, m, :, :] of a phase-2 tile (192 doubles each, contiguous in the natural layout) do not
//     depend on phase 1.  They are RESIDENT for the whole tile: the rows of b = 0, 1 are requested when the kernel starts
//     (one TMA bulk copy per row, cp.async.bulk / SASS UBLKCP), i.e. they arrive while phase 1 computes; the rows of b = 2
//     go into the shared memory phase 1 used, as soon as its consumers are done with it -- while they store T1 and wait at
//     the grid barrier.  Behind the barrier only T1 is still to come.  (Chunking these rows along n into a ring made 792
//     copies of 192 bytes per tile: the copy engine takes ~20 ns per copy whatever its size, phase 2 took 30 us; 16-byte
//     cp.async copies of the same pieces from 64 CTAs walking the same rows took 6-8 us per 74 KB.)
//   * T1 blocks (8 column-mode indices x 64, contiguous, three per ring slot): TMA bulk copies through a 4-slot ring.
//   * Phase-1 operands (rows of L and of x): 16-byte asynchronous copies (cp.async.cg / LDGSTS.128) that complete on
//     mbarriers, one per group of 16 k-rows; the 64 CTAs that share a tile of L start at different groups.
// Shared-memory pitches keep the 64-bit fragment loads of a half-warp on 16 distinct bank pairs: pitch = 4 (mod 16)
// doubles for unit-stride operands, pitch = 12 (mod 16) for the operator rows whose entries of one rank index lie 3
// doubles apart.  The mirrored update (right stack) reads the same memory with the roles of the two rank indices of A
// exchanged and the solution core transposed; both are addressing modes of the fragment loads, not copies.
// Zero (b, b2) blocks of the operator core (SLIM / MPO operators are block-sparse in their rank indices) are found by the
// CTAs that have no phase-1 tile while the others compute (one OR per warp into a mask word that the last phase clears
// again) and are skipped in the second contraction.
// The grid barrier is taken by the consumer warps only, so the producer's copies for the next phase never hold it up.
#include "common.cuh"
#include "fused_common.cuh"

namespace {

constexpr int NR = 64;                    // solution ranks on both sides
constexpr int NB = 3;                     // operator ranks on both sides
constexpr int N1_BM = 96, N1_LDA = N1_BM + 4, N1_LDB = NR + 4, N1_GROUPS = 4;
constexpr int N1_ELEMS = NR * (N1_LDA + N1_LDB);    // doubles of the phase-1 operand tiles
constexpr int N2_MB = 32;                 // row-mode indices per phase-2 tile
constexpr int N2_KC = 8;                  // column-mode indices per ring slot
constexpr int N2_BP = NR + 4;
constexpr int N2_SLOT = NB * N2_KC * N2_BP;         // T1 rows of one slot: [b][k][BP]
constexpr int N2_STAGES = 4;
constexpr int N2_LDT = NR * NB + 4;       // T2 in shared memory [MB][LDT]
constexpr int N2_XP = N2_MB + 4;          // transposed conj-side rows (mirror): [c2][XP]
static_assert(N2_BP % 16 == 4 && N2_LDT % 16 == 4 && N2_XP % 16 == 4 && N1_LDA % 16 == 4, "pitches");
static_assert(N1_GROUPS == 4, "group rotation uses a mask");

// operator row pitch for mode size n: 3 n + 12 = 12 (mod 16) needs 3 n = 0 (mod 16); the kernel takes n = 32 and n = 64
__host__ __device__ inline int nat_ap(int ntot) { return ntot * NB + 12; }
// shared-memory plan (doubles): [operator rows b = 0, 1 | operator rows b = 2 | T1 ring]; phase 1 lives behind the rows of
// b = 0, 1 (which arrive while it runs); T2 reuses the front once the second contraction is done
__host__ __device__ inline size_t nat_smem_doubles(int ntot) {
    const size_t ap = nat_ap(ntot), plan2 = 3 * N2_MB * ap + (size_t)N2_STAGES * N2_SLOT, plan1 = 2 * N2_MB * ap + N1_ELEMS;
    return plan2 > plan1 ? plan2 : plan1;
}

// 16-byte asynchronous copy global -> shared (L2 only) and the arrival of a thread's earlier copies on an mbarrier
__device__ __forceinline__ void ldgsts16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void ldgsts_arrive(unsigned long long* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// named CTA barriers between the producer warp and the consumers (0: __syncthreads, 1: the consumers among themselves)
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;\n" ::"r"(id), "r"(count) : "memory"); }
constexpr int BAR_SMEM_FREE = 2, BAR_GRID_PASSED = 3;

// Grid barrier of the CONSUMER warps (co-residency comes from the cooperative launch): one arrival per CTA on a counter whose
// top bit flips when the last CTA arrives (CTA 0 adds what is missing to 2^31), so the counter needs no reset between
// barriers or launches.
__device__ __forceinline__ void grid_barrier_consumers(unsigned* ctr, int G) {
    consumer_bar_sync();
    if (threadIdx.x == 0) {
        fence_proxy_async();
        __threadfence();
        const unsigned add = blockIdx.x == 0 ? 0x80000000u - (unsigned)(G - 1) : 1u;
        const unsigned old = atomicAdd(ctr, add);
        unsigned now;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(now) : "l"(ctr) : "memory");
        } while (((old ^ now) & 0x80000000u) == 0u);
        __threadfence();
    }
    consumer_bar_sync();
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
    unsigned* gbar;           // grid barrier counter (self-resetting)
    unsigned long long* stamps;   // optional %globaltimer stamps of CTA 0
};

// One contraction block b of one ring slot for a COMPILE-TIME mask QM of the non-zero output blocks q (zero blocks cost
// no issue slots).  The entry A~[b, m, n, q] of the (possibly mirrored) operator core sits in row (blk * 32 + m) of the
// resident rows at column 3 n + il with (blk, il) = (b, q), mirrored: (q, b).
template <bool SWAP, unsigned QM>
__device__ __forceinline__ void n_s2_block(double (&acc2)[NB][2][2][2], const double* __restrict__ arow,
                                           const double* __restrict__ bch, int ap, int b) {
    double bf[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) bf[j] = bch[b * N2_KC * N2_BP + 8 * j];
#pragma unroll
    for (int q = 0; q < NB; ++q) {
        if ((QM >> q) & 1u) {
            const int blk = SWAP ? q : b, il = SWAP ? b : q;
            double af[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) af[i] = arow[(size_t)(blk * N2_MB + 8 * i) * ap + il];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma(acc2[q][i][j][0], acc2[q][i][j][1], af[i], bf[j]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------- the kernel
template <bool SWAP>
__global__ void __launch_bounds__(THREADS) stack_nat_kernel(StackNatParams a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[N1_GROUPS + 2 + 2 * N2_STAGES];
    const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ntot = a.ntot, mtot = a.mtot, ap = nat_ap(ntot), nslots = ntot / N2_KC;
    const int tiles1 = ntot * 2, tiles2 = NR * (mtot / N2_MB);           // both <= gridDim.x (checked by the host)
    double* smem = reinterpret_cast<double*>(smem_raw);
    double* Arows = smem;                                  // [natural b * 32 + mm][ap]
    double* As = smem + (size_t)2 * N2_MB * ap;            // phase 1: [a][N1_LDA] = L[a, mt * 96 + .]
    double* Bs = As + (size_t)NR * N1_LDA;                 //          [a][N1_LDB] = x[a, nn, .]  (mirror: [a2][.] = x[a2, nn, .])
    double* ring = smem + (size_t)3 * N2_MB * ap;          // phase 2: [slot][b][k][N2_BP]
    double* T2s = smem;                                    //          [mm][N2_LDT], once the operator rows are dead
    unsigned long long* full1 = bars;                      // phase-1 K groups
    unsigned long long* abar = bars + N1_GROUPS;           // operator rows: [0] b = 0, 1   [1] b = 2
    unsigned long long* full2 = abar + 2;
    unsigned long long* empty2 = full2 + N2_STAGES;
    const bool has1 = cta < tiles1, has2 = cta < tiles2;
    const int nn = cta % ntot, mt = cta / ntot;            // phase-1 tile
    const int c = cta % NR, m0 = (cta / NR) * N2_MB;       // phase-2 tile
    const int rot = c % nslots;                            // the CTAs that share T1 / operator rows start at different chunks
    unsigned long long* st = (a.stamps && cta == 0 && tid == 0) ? a.stamps : nullptr;
    auto tick = [&](int k) {                               // diagnostics: %globaltimer of thread 0 of CTA 0
        if (st) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(st[k]));
    };
    tick(0);
    if (tid == 0) {
        for (int g = 0; g < N1_GROUPS; ++g) mbar_init(full1 + g, 32);    // one arrival per producer lane, behind its copies
        mbar_init(abar + 0, 1);
        mbar_init(abar + 1, 1);
        for (int s = 0; s < N2_STAGES; ++s) {
            mbar_init(full2 + s, 1);
            mbar_init(empty2 + s, CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == CONSUMER_WARPS) {
        // ============================================================================================== PRODUCER
        const size_t rowbytes = (size_t)ntot * NB * sizeof(double);
        auto issue_rows = [&](int blk, unsigned long long* bar) {       // rows (blk, lane) of the phase-2 tile
            const int mm = (lane + c) & 31;                             // 64 CTAs read these rows: not all the same one first
            bulk_g2s(Arows + (size_t)(blk * N2_MB + mm) * ap, a.A + ((size_t)blk * mtot + m0 + mm) * ntot * NB, (unsigned)rowbytes,
                     bar);
        };
        if (has1) {
            // rows of 96 (L) and 64 (x) doubles as 16-byte pieces: eight lanes share a row of x (four pieces each) and a row of
            // the L tile (six pieces each), the warp covers four rows per step
            const double* lsrc = a.L + mt * N1_BM + 2 * (lane & 7);
            const double* xsrc = a.x + (size_t)nn * NR + 2 * (lane & 7);
            if (SWAP) {                                    // every row of the transposed core is needed from the first k on
#pragma unroll 4
                for (int i = 0; i < 16; ++i) {
                    const int row = 4 * i + (lane >> 3);
#pragma unroll
                    for (int h = 0; h < 4; ++h)
                        ldgsts16(Bs + (size_t)row * N1_LDB + 2 * (lane & 7) + 16 * h, xsrc + (size_t)row * ntot * NR + 16 * h);
                }
            }
            for (int g0 = 0; g0 < N1_GROUPS; ++g0) {
                const int g = (g0 + nn) & (N1_GROUPS - 1); // the CTAs that share this tile of L start at different groups
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = 16 * g + 4 * i + (lane >> 3);
#pragma unroll
                    for (int h = 0; h < 6; ++h)
                        ldgsts16(As + (size_t)k * N1_LDA + 2 * (lane & 7) + 16 * h, lsrc + (size_t)k * (NB * NR) + 16 * h);
                    if (!SWAP) {
#pragma unroll
                        for (int h = 0; h < 4; ++h)
                            ldgsts16(Bs + (size_t)k * N1_LDB + 2 * (lane & 7) + 16 * h, xsrc + (size_t)k * ntot * NR + 16 * h);
                    }
                }
                ldgsts_arrive(full1 + g);
            }
        }
        if (has2) {                                        // behind the phase-1 operands: these have all of phase 1 to arrive
            if (lane == 0) mbar_expect_tx(abar + 0, (unsigned)(2 * N2_MB * rowbytes));
            __syncwarp();
            issue_rows(0, abar + 0);
            issue_rows(1, abar + 0);
        }
        if (has1) {
            named_sync(BAR_SMEM_FREE, THREADS);            // the consumers are done with the shared memory of phase 1
        }
        if (has2) {
            fence_proxy_async();
            if (lane == 0) mbar_expect_tx(abar + 1, (unsigned)(N2_MB * rowbytes));
            __syncwarp();
            issue_rows(2, abar + 1);
        }
        named_sync(BAR_GRID_PASSED, THREADS);              // T1 is complete
        if (has2) {
            fence_proxy_async();
            for (int t = 0; t < nslots; ++t) {
                const int s = t % N2_STAGES;
                if (t >= N2_STAGES) mbar_wait(empty2 + s, ((t / N2_STAGES) & 1) ^ 1);
                const int tn = t + rot < nslots ? t + rot : t + rot - nslots;
                if (lane == 0) mbar_expect_tx(full2 + s, (unsigned)(N2_SLOT * sizeof(double)));
                __syncwarp();
                if (lane < NB)
                    bulk_g2s(ring + (size_t)s * N2_SLOT + (size_t)lane * N2_KC * N2_BP,
                             a.T1p + (((size_t)lane * NR + c) * ntot + (size_t)tn * N2_KC) * N2_BP, N2_KC * N2_BP * 8, full2 + s);
            }
        }
        return;
    }

    // ================================================================================================== CONSUMERS
    const int tile = warp & 7, khalf = warp >> 3;
    const int fr = lane >> 2, fk = lane & 3;
    if (has1) {
        // ---- phase 1: 8 warp tiles of 48 x 16, two warps per tile splitting every K group
        const int wm0 = (tile & 1) * 48, wn0 = (tile >> 1) * 16;
        double acc[6][2][2];
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        tick(20);
        for (int g0 = 0; g0 < N1_GROUPS; ++g0) {
            const int g = (g0 + nn) & (N1_GROUPS - 1);
            mbar_wait(full1 + g, 0);
            if (g0 == 0) tick(21);
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
                const int kk = 16 * g + 8 * khalf + 4 * s2;
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
        tick(22);
        // the two K-halves meet in shared memory (the operand tiles are dead once every consumer got here)
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
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const double2 o = *reinterpret_cast<const double2*>(red + (8 * i + fr) * 16 + 8 * j + 2 * fk);
                    acc[i][j][0] += o.x;
                    acc[i][j][1] += o.y;
                }
        }
        named_arrive(BAR_SMEM_FREE, THREADS);              // the producer may fill this memory for phase 2
        if (khalf == 0) {
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const int m = mt * N1_BM + wm0 + 8 * i + fr;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    double* dst = a.T1p + ((size_t)m * ntot + nn) * N2_BP + wn0 + 8 * j + 2 * fk;
                    *reinterpret_cast<double2*>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
                }
            }
        }
        tick(23);
    }
    // which (b, b2) blocks of the (possibly mirrored) operator core hold a non-zero entry: the CTAs without a phase-1 tile look
    // through the core while the others compute (everybody shares the work when every CTA has a tile); the bits meet in
    // one word behind the grid barrier
    {
        const int idle = G > tiles1 ? G - tiles1 : 0, scanners = idle ? idle : G, me = idle ? cta - tiles1 : cta;
        if (me >= 0) {
            const int per_b = mtot * ntot * NB, total = NB * per_b;
            unsigned bits = 0;
            for (int e = me * 512 + tid; e < total; e += scanners * 512)
                if (a.A[e] != 0.0) bits |= 1u << (SWAP ? (e % NB) * NB + e / per_b : (e / per_b) * NB + e % NB);
            __syncwarp();
            bits = __reduce_or_sync(0xffffffffu, bits);
            if (lane == 0 && bits) atomicOr(a.mask, (unsigned long long)bits);
        }
    }
    // rows of the conj-side core the phase-2 tile contracts with at the end: x[c, m0 + mm, c2] (mirror: x[c2, m0 + mm, c]), four
    // entries per consumer thread, requested before the barrier (they do not depend on phase 1 either)
    double xr[4] = {0.0, 0.0, 0.0, 0.0};
    if (has2) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int e = tid + 512 * k;
            xr[k] = SWAP ? a.x[((size_t)(e >> 5) * mtot + m0 + (e & 31)) * NR + c] : a.x[((size_t)c * mtot + m0 + (e >> 6)) * NR + (e & 63)];
        }
    }
    tick(1);
    grid_barrier_consumers(a.gbar, G);
    named_arrive(BAR_GRID_PASSED, THREADS);
    tick(2);
    if (has2) {
        // ---- phase 2: 8 warp tiles of 16 (m) x 16 (a2) x 3 (b2), two warps per tile splitting the 8 k of every ring slot
        const unsigned blockmask = (unsigned)__ldcg(a.mask);
        const int wm0 = (tile & 1) * 16, wn0 = (tile >> 1) * 16;
        double acc2[NB][2][2][2];
#pragma unroll
        for (int q = 0; q < NB; ++q)
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) acc2[q][i][j][0] = acc2[q][i][j][1] = 0.0;
        const int kcol = 4 * khalf + fk;
        mbar_wait(abar + 0, 0);
        mbar_wait(abar + 1, 0);
        tick(12);
        for (int t = 0; t < nslots; ++t) {
            const int s = t % N2_STAGES;
            const int tn = t + rot < nslots ? t + rot : t + rot - nslots;
            mbar_wait(full2 + s, (t / N2_STAGES) & 1);
            if (t == 0) tick(13);
            const double* arow = Arows + (size_t)(wm0 + fr) * ap + (tn * N2_KC + kcol) * NB;
            const double* bch = ring + (size_t)s * N2_SLOT + kcol * N2_BP + wn0 + fr;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                switch ((blockmask >> (b * NB)) & 7u) {    // warp-uniform
                    case 1: n_s2_block<SWAP, 1>(acc2, arow, bch, ap, b); break;
                    case 2: n_s2_block<SWAP, 2>(acc2, arow, bch, ap, b); break;
                    case 3: n_s2_block<SWAP, 3>(acc2, arow, bch, ap, b); break;
                    case 4: n_s2_block<SWAP, 4>(acc2, arow, bch, ap, b); break;
                    case 5: n_s2_block<SWAP, 5>(acc2, arow, bch, ap, b); break;
                    case 6: n_s2_block<SWAP, 6>(acc2, arow, bch, ap, b); break;
                    case 7: n_s2_block<SWAP, 7>(acc2, arow, bch, ap, b); break;
                    default: break;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty2 + s);
        }
        tick(14);
        consumer_bar_sync();                               // every warp is done with the ring and the operator rows
        double* Xs = ring;                                 // [mm][N2_BP]   (mirror: [c2][N2_XP])
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
        tick(15);
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
        tick(16);
        double* dst = a.part + (size_t)cta * (NR * NB) * NR;
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
            for (int u = 0; u < 4; ++u)
                *reinterpret_cast<double2*>(dst + (size_t)(24 * wr + 8 * t + fr) * NR + 32 * wc + 8 * u + 2 * fk) =
                    make_double2(acc[t][u][0], acc[t][u][1]);
        tick(17);
    }
    tick(3);
    grid_barrier_consumers(a.gbar, G);
    tick(4);
    if (cta == 0 && tid == 0) *a.mask = 0ull;              // everybody has read it: the next launch finds it clear
    // out[e] = sum over the tile partials in a fixed order.  A CTA takes 84 consecutive entries (148 CTAs cover the 12288 in one
    // round); lanes run along the entries (coalesced rows of the partials), six thread groups split the tiles (group g
    // adds tiles g, g + 6, ... in that order with all its loads in flight at once), the first group adds the six sums in
    // group order
    constexpr int E = NR * NB * NR, RW = 84, RG = 6, RT = 22;
    __shared__ double red[RG][RW];
    const int ent = tid % RW, grp = tid / RW;
    for (int ch = cta; ch * RW < E; ch += G) {
        const int e = ch * RW + ent;
        if (grp < RG && e < E) {
            double v[RT];
#pragma unroll
            for (int k = 0; k < RT; ++k) {
                const int t = grp + RG * k;
                v[k] = t < tiles2 ? __ldcg(a.part + (size_t)t * E + e) : 0.0;
            }
            double sacc = v[0];
#pragma unroll
            for (int k = 1; k < RT; ++k) sacc += v[k];
            for (int t = grp + RG * RT; t < tiles2; t += RG) sacc += __ldcg(a.part + (size_t)t * E + e);
            red[grp][ent] = sacc;
        }
        consumer_bar_sync();
        if (grp == 0 && e < E) {
            double sacc = red[0][ent];
#pragma unroll
            for (int g = 1; g < RG; ++g) sacc += red[g][ent];
            a.out[e] = sacc;
        }
        consumer_bar_sync();
    }
    tick(5);
}

}  // namespace

// Shapes of the natural-layout kernel: both solution ranks 64, both operator ranks 3, square mode of 32 or 64 (one tile per
// CTA in both phases, operator rows resident in shared memory), operands 16-byte aligned.  Everything else runs the
// image-based kernel of fused.cu or the generic chain.
bool sktt_stack_nat_supported(const sktt_ctx* ctx, int dtype, long long rin, long long Rin, long long m, long long n,
                              long long rout, long long Rout, const void* stack, const void* x, const void* A) {
    if (dtype != SKTT_F64 || ctx->gemm_mode == 1 || (ctx->debug & 64)) return false;
    if (rin != NR || rout != NR || Rin != NB || Rout != NB) return false;
    if (m != n || (m != 32 && m != 64)) return false;
    if (2 * n > ctx->sm_count || NR * (m / N2_MB) > ctx->sm_count) return false;
    return (((uintptr_t)stack | (uintptr_t)x | (uintptr_t)A) & 15u) == 0;
}

#define STACK_NAT_MASK_OFF 3584          // byte offsets in the scalar area of the context scratch (zeroed when it is allocated)
#define STACK_NAT_GBAR_OFF 3592
#define STACK_NAT_STAMP_OFF 3600         // 32 stamps (debug bit 0)

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
    a.gbar = (unsigned*)((char*)ctx->scratch + STACK_NAT_GBAR_OFF);
    a.stamps = (ctx->debug & 1) ? (unsigned long long*)((char*)ctx->scratch + STACK_NAT_STAMP_OFF) : nullptr;
    const size_t smem = nat_smem_doubles((int)n) * sizeof(double);
    SKTT_ONCE_PER_DEVICE(ctx);
    if (!configured) {
        const int cap = (int)(nat_smem_doubles(64) * sizeof(double));
        SKTT_CUDA(ctx, cudaFuncSetAttribute(stack_nat_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        SKTT_CUDA(ctx, cudaFuncSetAttribute(stack_nat_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        configured = true;
    }
    void* args[] = {&a};
    const void* fn = mirror ? (const void*)stack_nat_kernel<true> : (const void*)stack_nat_kernel<false>;
    if (ctx->debug & 128) {                                // diagnostics: plain launch (no co-residency guarantee)
        if (mirror) stack_nat_kernel<true><<<ctx->sm_count, THREADS, smem, ctx->stream>>>(a);
        else stack_nat_kernel<false><<<ctx->sm_count, THREADS, smem, ctx->stream>>>(a);
        SKTT_LAUNCH_CHECK(ctx);
        return 0;
    }
    SKTT_CUDA(ctx, cudaLaunchCooperativeKernel(fn, dim3(ctx->sm_count), dim3(THREADS), args, smem, ctx->stream));
    ctx->launches++;
    return 0;
}
