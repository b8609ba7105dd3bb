// peer.cu -- the one exchange step of the path (BASELINE config 4): the micro-matvec
//     y[c,m,c2] = sum L[a,b,c] v[a,n,a2] A[b,m,n,b2] Rt[a2,b2,c2]
// sharded over the OUTPUT solution-rank index c across the GPUs of one NVSwitch domain, one process per GPU.  Rank g runs
// the three strided contractions for its rows c in [lo, hi) -- the shard is an offset and an extent of the left stack, never
// a copy -- and the LAST contraction's epilogue stores its tile of y straight into the y buffer of EVERY rank through
// peer-mapped pointers (CUDA IPC over NVLink): the all-gather is fused into the producing kernel, tile by tile, no separate
// collective and no staging copy.  A flag barrier across the GPUs (one tiny kernel: system-scope release of an epoch
// counter into every peer, acquire-spin on the own slots) orders the remote stores before the consumers of y.
//
// Buffers that peers write into must come from sktt_peer_alloc (cudaMalloc, exportable as an IPC handle); the host layer
// exchanges the 64-byte handles through torch.distributed and opens them once per solver call.
#include "common.cuh"
#include "blas1.cuh"

extern "C" int sktt_peer_alloc(sktt_ctx* ctx, int64_t bytes, void** out) {
    if (!ctx || !out || bytes <= 0) return SKTT_ERR_ARG;
    SKTT_CUDA(ctx, cudaMalloc(out, (size_t)bytes));
    SKTT_CUDA(ctx, cudaMemsetAsync(*out, 0, (size_t)bytes, ctx->stream));
    SKTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int sktt_peer_free(sktt_ctx* ctx, void* ptr) {
    if (!ctx) return SKTT_ERR_ARG;
    if (ptr) {
        SKTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        SKTT_CUDA(ctx, cudaFree(ptr));
    }
    return 0;
}

// handle_out: 64 bytes (cudaIpcMemHandle_t)
extern "C" int sktt_peer_export(sktt_ctx* ctx, void* ptr, uint8_t* handle_out) {
    if (!ctx || !ptr || !handle_out) return SKTT_ERR_ARG;
    cudaIpcMemHandle_t h;
    SKTT_CUDA(ctx, cudaIpcGetMemHandle(&h, ptr));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle_out, &h, 64);
    return 0;
}

extern "C" int sktt_peer_open(sktt_ctx* ctx, const uint8_t* handle, void** out) {
    if (!ctx || !handle || !out) return SKTT_ERR_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    SKTT_CUDA(ctx, cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int sktt_peer_close(sktt_ctx* ctx, void* ptr) {
    if (!ctx) return SKTT_ERR_ARG;
    if (ptr) {
        SKTT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        SKTT_CUDA(ctx, cudaIpcCloseMemHandle(ptr));
    }
    return 0;
}

// flags layout on every rank: unsigned long long [world]; slot s of rank g holds the last epoch rank s has announced to g.
struct PeerBarrierArgs {
    int world, rank;
    unsigned long long epoch;
    unsigned long long* flags[8];     // flags[g] = rank g's slot array (peer-mapped for g != rank)
    unsigned long long spin_limit;    // clock64 ticks before the kernel gives up (a lost peer must not hang the GPU)
    int* timeout_flag;                // local, set to 1 when the spin gave up
};

__global__ void peer_barrier_kernel(PeerBarrierArgs a) {
    const int t = threadIdx.x;
    if (t < a.world) {
        // everything this GPU stored before (the epilogue's remote tiles included) becomes visible system-wide first
        __threadfence_system();
        volatile unsigned long long* dst = a.flags[t] + a.rank;
        *dst = a.epoch;
        __threadfence_system();
        volatile unsigned long long* mine = a.flags[a.rank] + t;
        const long long t0 = clock64();
        while (*mine < a.epoch) {
            if ((unsigned long long)(clock64() - t0) > a.spin_limit) {
                *a.timeout_flag = 1;
                break;
            }
        }
        __threadfence_system();
    }
}

// All ranks must call this the same number of times with the same epoch sequence (1, 2, 3, ...).
extern "C" int sktt_peer_barrier(sktt_ctx* ctx, int world, int rank, uint64_t epoch, void* const* flags /* [world] */,
                                 int32_t* timeout_flag_dev) {
    if (!ctx || !flags || world < 1 || world > 8 || rank < 0 || rank >= world || !timeout_flag_dev) return SKTT_ERR_ARG;
    PeerBarrierArgs a;
    a.world = world;
    a.rank = rank;
    a.epoch = epoch;
    for (int g = 0; g < 8; ++g) a.flags[g] = g < world ? (unsigned long long*)flags[g] : nullptr;
    a.spin_limit = 20ull * 1000ull * 1000ull * 1000ull;       // ~10 s at 2 GHz
    a.timeout_flag = timeout_flag_dev;
    peer_barrier_kernel<<<1, 32, 0, ctx->stream>>>(a);
    SKTT_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" int64_t sktt_sharded_matvec_work(int64_t r, int64_t R, int64_t m, int64_t n, int64_t r2, int64_t R2, int64_t rows) {
    (void)r;
    return R * rows * n * r2 + rows * m * r2 * R2;
}

// Rows [lo, hi) of y = M v for the one-site local operator (Lst [r, R, r], A [R, m, n, R2], Rst [r2, R2, r2]), written into
// y_local [r, m, r2] AND into the same rows of y_peers[0 .. npeer) -- peer-mapped y buffers of the other ranks.
//   T1[b,cs,n,a2]  = sum_a  L[a,b,lo+cs] v[a,n,a2]
//   T2[cs,m,a2,b2] = sum_bn A[b,m,n,b2] T1[b,cs,n,a2]
//   y[lo+cs,m,c2]  = sum    T2[cs,m,a2,b2] Rt[a2,b2,c2]        <- epilogue stores to every rank
extern "C" int sktt_sharded_matvec(sktt_ctx* ctx, int dtype, int64_t r, int64_t R, int64_t m, int64_t n, int64_t r2,
                                   int64_t R2, const void* Lst, const void* A, const void* Rst, const void* v, int64_t lo,
                                   int64_t hi, void* y_local, int npeer, void* const* y_peers, void* work) {
    if (!ctx || !Lst || !A || !Rst || !v || !y_local || !work) return SKTT_ERR_ARG;
    SKTT_TRY(check_dtype(ctx, dtype));
    if (lo < 0 || hi > r || npeer < 0 || npeer > 7 || (npeer > 0 && !y_peers))
        return sktt_fail(ctx, SKTT_ERR_ARG, "sharded_matvec: bad shard or peer list");
    const long long rc = hi - lo;
    if (rc <= 0) return 0;
    const size_t es = dtype_size(dtype);
    char* T1 = (char*)work;
    char* T2 = T1 + (size_t)(R * rc * n * r2) * es;
    const long long big = 1LL << 40;
    GemmDesc g1 = gemm_desc(R * rc, n * r2, r, (const char*)Lst + (size_t)lo * es, mk_idx(rc, r, 1), mk_idx(big, 0, R * r), v,
                            mk_idx(big, 0, n * r2), mk_idx(big, 0, 1), T1, mk_idx(big, 0, n * r2), mk_idx(big, 0, 1));
    SKTT_TRY(sktt_gemm_run(ctx, dtype, g1));
    GemmDesc g2 = gemm_desc(rc * r2, m * R2, R * n, T1, mk_idx(r2, n * r2, 1), mk_idx(n, rc * n * r2, r2), A,
                            mk_idx(n, m * n * R2, R2), mk_idx(R2, n * R2, 1), T2, mk_idx(r2, m * r2 * R2, R2),
                            mk_idx(R2, r2 * R2, 1));
    SKTT_TRY(sktt_gemm_run(ctx, dtype, g2));
    const size_t row_off = (size_t)lo * m * r2 * es;
    GemmDesc g3 = gemm_desc(rc * m, r2, r2 * R2, T2, mk_idx(big, 0, r2 * R2), mk_idx(big, 0, 1), Rst, mk_idx(big, 0, r2),
                            mk_idx(big, 0, 1), (char*)y_local + row_off, mk_idx(big, 0, r2), mk_idx(big, 0, 1));
    g3.npeer = npeer;
    for (int q = 0; q < npeer; ++q) g3.Cpeer[q] = (char*)y_peers[q] + row_off;
    return sktt_gemm_run(ctx, dtype, g3);
}
