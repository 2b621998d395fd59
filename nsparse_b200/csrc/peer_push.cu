// nsparse-b200: the pusher of the multi-GPU allgatherv (see PeerOut in common.cuh).
//
// A persistent kernel on a few SMs of its own, launched on a side stream BEFORE the numeric kernels of a
// product.  One warp per CTA, and of that warp one thread: it drives the TMA.  The thread draws tickets; ticket
// k is the k-th tile of this rank's block of C that the numeric kernels complete (tiles_done).  As soon as a
// ticket's tile is published the thread issues two bulk loads global -> shared (cp.async.bulk ... mbarrier::
// complete_tx: C.col and C.val of the tile, written moments ago by the computing SMs: L2 hits) into the next
// free stage of a ring in shared memory; when a stage's mbarrier says the bytes have landed it issues one bulk
// store shared -> global per peer and array (2 * npeers stores of 16 / 32 KiB over NVLink) and commits them as
// a bulk group; a stage is reused when cp.async.bulk.wait_group.read says the stores have read it.  With a ring
// of kStages the SM has kStages tiles in flight in each direction and no thread ever touches the data: the
// first version (threads load, then hand to the TMA, two stages) was latency bound at ~15 GB/s per SM
// (profiles/r2_bench_g2_pusher_v1_*.json).  The computing SMs never execute a remote store.
//
// Tiles are aligned in the full arrays (all bases are 256-byte aligned allocations), so every tile but the first
// and the last of the block is two aligned bulk copies; the at most six entries that a ragged block end leaves
// outside 16-byte granules are copied by the thread itself.
#include "context.h"

namespace nsp {

constexpr int kPushStagesBytes = 192 * 1024;      // ring of the pusher CTA (one CTA per SM)

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_load(void *sdst, const void *gsrc, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
}

template <int VB>   // bytes per value
__global__ void __launch_bounds__(32, 1)
push_tiles_kernel(const int *__restrict__ c_col, const unsigned char *__restrict__ c_val, long long spin_limit,
                  const __grid_constant__ PeerOut peer)
{
    // c_col / c_val: this rank's FULL local arrays (the block starts at element peer.off)
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int T = 1 << kTileLog;
    constexpr int kStageBytes = T * (4 + VB);
    constexpr int kStages = kPushStagesBytes / kStageBytes;
    static_assert(kStages >= 2 && kStages <= 8, "ring depth");
    __shared__ __align__(8) unsigned long long s_full[kStages];
    if (threadIdx.x != 0) return;
    for (int s = 0; s < kStages; ++s) mbar_init(&s_full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");

    // per stage: what was loaded (entries [ca, cb) of col as granules, [va, vb) of val)
    long long st_ca[kStages], st_va[kStages];
    int st_ncg[kStages], st_nvg[kStages];
    unsigned phase = 0;                          // bit s: parity the next wait on stage s expects
    int loaded = 0, stored = 0;                  // tiles whose loads were issued / whose stores were issued
    int ticket = -1;                             // a ticket drawn but not yet published by the producers
    bool drained = false;                        // no tickets left
    long long t_wait = 0;
    constexpr int VG = 16 / VB;
    while (true) {
        bool progress = false;
        // ---- issue loads: next ticket into the next free stage ----
        if (!drained && loaded - stored < kStages) {
            if (ticket < 0) {
                ticket = atomicAdd(peer.q_ctl + 1, 1);
                if (ticket >= peer.ntiles) {
                    drained = true;
                    ticket = -1;
                } else {
                    t_wait = clock64();
                }
            }
            if (ticket >= 0) {
                int tile;
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(tile) : "l"(peer.queue + ticket) : "memory");
                if (tile >= 0) {
                    const int s = loaded % kStages;
                    // the stage was handed to the bulk stores kStages tiles ago: that group must have read it, i.e. at
                    // most kStages - 1 - (loaded - stored) of the groups committed since may still be reading
                    switch (kStages - 1 - (loaded - stored)) {
                        case 0: asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); break;
                        case 1: asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); break;
                        case 2: asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); break;
                        case 3: asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); break;
                        case 4: asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); break;
                        case 5: asm volatile("cp.async.bulk.wait_group.read 5;" ::: "memory"); break;
                        case 6: asm volatile("cp.async.bulk.wait_group.read 6;" ::: "memory"); break;
                        default: asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory"); break;
                    }
                    const long long lo = (peer.tile0 + tile) << kTileLog;
                    const long long a = lo > peer.off ? lo : peer.off;
                    const long long b = (lo + T) < (peer.off + peer.nnz) ? (lo + T) : (peer.off + peer.nnz);
                    const long long ca = (a + 3) & ~3ll, cb = b & ~3ll;
                    const long long va = (a + VG - 1) & ~(long long)(VG - 1), vb = b & ~(long long)(VG - 1);
                    const int ncg = cb > ca ? (int)((cb - ca) >> 2) : 0;
                    const int nvg = vb > va ? (int)((vb - va) / VG) : 0;
                    st_ca[s] = ca;
                    st_va[s] = va;
                    st_ncg[s] = ncg;
                    st_nvg[s] = nvg;
                    // the tile was written through the generic proxy by other SMs (made visible by the acquire above);
                    // the TMA reads through the async proxy
                    asm volatile("fence.proxy.async.global;" ::: "memory");
                    unsigned char *s_col = smem + (size_t)s * kStageBytes;
                    unsigned char *s_val = s_col + (size_t)T * 4;
                    mbar_expect_tx(&s_full[s], (unsigned)(ncg + nvg) * 16u);
                    if (ncg) bulk_load(s_col, c_col + ca, (unsigned)ncg * 16u, &s_full[s]);
                    if (nvg) bulk_load(s_val, c_val + (size_t)va * VB, (unsigned)nvg * 16u, &s_full[s]);
                    // ragged ends (first / last tile of the block): at most six entries, copied right here
                    if (ca != a || cb != b || va != a || vb != b || ncg == 0 || nvg == 0) {
                        for (long long k = a; k < b; ++k) {
                            if (ncg == 0 || k < ca || k >= cb) {
                                const int c = __ldcg(c_col + k);
                                for (int p = 0; p < peer.n; ++p) peer.col[p][k] = c;
                            }
                            if (nvg == 0 || k < va || k >= vb) {
                                if (VB == 4) {
                                    const unsigned v = __ldcg(reinterpret_cast<const unsigned *>(c_val) + k);
                                    for (int p = 0; p < peer.n; ++p) static_cast<unsigned *>(peer.val[p])[k] = v;
                                } else {
                                    const unsigned long long v = __ldcg(reinterpret_cast<const unsigned long long *>(c_val) + k);
                                    for (int p = 0; p < peer.n; ++p) static_cast<unsigned long long *>(peer.val[p])[k] = v;
                                }
                            }
                        }
                    }
                    ++loaded;
                    ticket = -1;
                    progress = true;
                } else if (clock64() - t_wait > spin_limit) {     // a producer never finished: report, do not hang
                    atomicExch(peer.q_ctl + 2, 1);
                    drained = true;
                    ticket = -1;
                }
            }
        }
        // ---- issue stores: the oldest loaded stage, once its bytes have landed ----
        if (stored < loaded) {
            const int s = stored % kStages;
            if (mbar_try_wait(&s_full[s], (phase >> s) & 1u)) {
                phase ^= 1u << s;
                unsigned char *s_col = smem + (size_t)s * kStageBytes;
                unsigned char *s_val = s_col + (size_t)T * 4;
                for (int p = 0; p < peer.n; ++p) {
                    if (st_ncg[s]) bulk_store(peer.col[p] + st_ca[s], s_col, (unsigned)st_ncg[s] * 16u);
                    if (st_nvg[s]) bulk_store(static_cast<unsigned char *>(peer.val[p]) + (size_t)st_va[s] * VB, s_val, (unsigned)st_nvg[s] * 16u);
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                ++stored;
                progress = true;
            }
        }
        if (drained && stored == loaded) break;
        if (!progress) __nanosleep(128);
    }
    // all bulk stores of this CTA complete (not just read) before the kernel ends
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void push_init_kernel(int *tile_cnt, int *queue, int *q_ctl, int ntiles)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ntiles) {
        tile_cnt[i] = 0;
        queue[i] = -1;
    }
    if (i < 2) q_ctl[i] = 0;       // tail, head; the error flag [2] is sticky until read
}

// Everything of the hand-off that ALLOCATES (workspace for `ntiles` tiles, side stream, events) or loads code:
// peer_push_begin calls it, and the single-process multi-GPU entry calls it on every GPU BEFORE any pusher starts
// (cudaMalloc on one GPU of a process with peer access enabled synchronises with the peers, i.e. would wait for
// their spinning pushers).
int peer_push_reserve(nsp_context *ctx, long long ntiles)
{
    if ((size_t)ntiles > ctx->push_cap) {
        NSP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->push_stream) NSP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->push_stream));
        cudaFree(ctx->d_push_ws);
        ctx->d_push_ws = nullptr;
        ctx->push_cap = 0;
        const size_t cap = (size_t)ntiles + (size_t)ntiles / 8 + 64;
        NSP_CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_push_ws, sizeof(int) * (2 * cap + 8)));
        NSP_CUDA_TRY(ctx, cudaMemset(ctx->d_push_ws, 0, sizeof(int) * (2 * cap + 8)));
        ctx->push_cap = cap;
    }
    if (!ctx->push_stream) {
        NSP_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->push_stream, cudaStreamNonBlocking));
        NSP_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_push_fork, cudaEventDisableTiming));
        NSP_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_push_join, cudaEventDisableTiming));
        cudaFuncAttributes at;
        NSP_CUDA_TRY(ctx, cudaFuncGetAttributes(&at, push_tiles_kernel<4>));
        NSP_CUDA_TRY(ctx, cudaFuncGetAttributes(&at, push_tiles_kernel<8>));
        NSP_CUDA_TRY(ctx, cudaFuncGetAttributes(&at, push_init_kernel));
        NSP_CUDA_TRY(ctx, cudaFuncSetAttribute(push_tiles_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPushStagesBytes));
        NSP_CUDA_TRY(ctx, cudaFuncSetAttribute(push_tiles_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPushStagesBytes));
    }
    return 0;
}

// PeerOut's tile fields and the pusher launch on the push stream.  Called by spgemm_numeric right before its
// kernels when peers are set.
int peer_push_begin(nsp_context *ctx, const int *c_col_full, const void *c_val_full, int val_bytes, long long nnz_block)
{
    PeerOut &po = ctx->peer_out;
    po.nnz = nnz_block;
    po.tile_log = kTileLog;
    po.done = nullptr;
    po.tile0 = po.off >> kTileLog;
    po.ntiles = nnz_block > 0 ? (int)(((po.off + nnz_block - 1) >> kTileLog) - po.tile0 + 1) : 0;
    ctx->push_active = false;
    if (po.n <= 0 || po.ntiles == 0) return 0;
    if (peer_push_reserve(ctx, po.ntiles) != 0) return -1;
    po.q_ctl = ctx->d_push_ws;
    po.tile_cnt = ctx->d_push_ws + 8;
    po.queue = po.tile_cnt + ctx->push_cap;
    push_init_kernel<<<(po.ntiles + 255) / 256, 256, 0, ctx->stream>>>(po.tile_cnt, po.queue, po.q_ctl, po.ntiles);
    ctx->launches += 1;
    NSP_CUDA_TRY(ctx, cudaGetLastError());
    NSP_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_push_fork, ctx->stream));
    NSP_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->push_stream, ctx->ev_push_fork, 0));
    int ctas = ctx->opt_push_sms > 0 ? (int)ctx->opt_push_sms : 16;
    if (ctas > ctx->sm_count / 2) ctas = ctx->sm_count / 2;
    if (ctas > po.ntiles) ctas = po.ntiles;
    const long long spin_limit = 20ll * 1000 * 1000 * 1000;      // ~10 s of SM clocks: a hang guard, not a schedule
    const unsigned char *cv = static_cast<const unsigned char *>(c_val_full);
    if (val_bytes == 4)
        push_tiles_kernel<4><<<ctas, 32, kPushStagesBytes, ctx->push_stream>>>(c_col_full, cv, spin_limit, po);
    else
        push_tiles_kernel<8><<<ctas, 32, kPushStagesBytes, ctx->push_stream>>>(c_col_full, cv, spin_limit, po);
    ctx->launches += 1;
    NSP_CUDA_TRY(ctx, cudaGetLastError());
    NSP_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_push_join, ctx->push_stream));
    ctx->push_active = true;
    ctx->push_ctas = ctas;
    ctx->last_push = po;
    return 0;
}

// joins the pusher into the context's stream (every tile of the block has left for the peers when the stream
// reaches this point)
int peer_push_end(nsp_context *ctx)
{
    if (!ctx->push_active) return 0;
    ctx->push_active = false;
    NSP_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_push_join, 0));
    return 0;
}

}  // namespace nsp
