// nsparse-b200: the pusher of the multi-GPU allgatherv (see PeerOut in common.cuh).
//
// A persistent kernel on a few SMs of its own, launched on a side stream BEFORE the numeric kernels of a
// product.  Its CTAs draw tickets; ticket k waits until the k-th completed tile of this rank's block of C
// has been published by the numeric kernels (tiles_done), stages the tile's C.col and C.val in shared memory
// with 128-bit loads (they were written moments ago: L2 hits) and hands the two buffers to the TMA:
// one cp.async.bulk shared -> global per peer and array, i.e. 2 * npeers bulk stores of 32 / 64 KiB that
// the copy engine of the SM carries over NVLink while the CTA already waits for and loads its next tile
// (two stages; a stage is reused when cp.async.bulk.wait_group.read says its bulk stores have left shared
// memory).  The computing SMs never execute a remote store.
//
// Tiles are aligned in the full arrays (all bases are 256-byte aligned allocations), so every tile but the
// first and the last of the block is one aligned bulk copy; the up to three entries that a ragged block end
// leaves outside 16-byte granules are stored with plain 4 / 8-byte stores.
#include "context.h"

namespace nsp {

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
}

template <int VB>   // bytes per value
__global__ void __launch_bounds__(256, 1)
push_tiles_kernel(const int *__restrict__ c_col, const unsigned char *__restrict__ c_val, long long spin_limit,
                  const __grid_constant__ PeerOut peer)
{
    // c_col / c_val: this rank's FULL local arrays (the block starts at element peer.off)
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int T = 1 << kTileLog;
    constexpr int kStageBytes = T * (4 + VB);
    __shared__ int s_tile;
    const int t = threadIdx.x;
    int stage = 0;
    while (true) {
        if (t == 0) {
            int tile = -2;
            const int ticket = atomicAdd(peer.q_ctl + 1, 1);
            if (ticket < peer.ntiles) {
                const long long t0 = clock64();
                while (true) {
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(tile) : "l"(peer.queue + ticket) : "memory");
                    if (tile >= 0) break;
                    if (clock64() - t0 > spin_limit) {          // a producer never finished: report, do not hang
                        atomicExch(peer.q_ctl + 2, 1);
                        tile = -2;
                        break;
                    }
                    __nanosleep(256);
                }
            }
            s_tile = tile;
            // the stage about to be overwritten was handed to the TMA two tiles ago
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        __syncthreads();
        const int tile = s_tile;
        if (tile < 0) break;
        // entries [a, b) of the full arrays
        const long long lo = (peer.tile0 + tile) << kTileLog;
        const long long a = lo > peer.off ? lo : peer.off;
        const long long b = (lo + T) < (peer.off + peer.nnz) ? (lo + T) : (peer.off + peer.nnz);
        unsigned char *s_col = smem + (size_t)stage * kStageBytes;
        unsigned char *s_val = s_col + (size_t)T * 4;
        // 16-byte granules of the tile that lie completely inside [a, b), per array
        const long long ca = (a + 3) & ~3ll, cb = b & ~3ll;                                  // col: 4 entries per granule
        constexpr int VG = 16 / VB;
        const long long va = (a + VG - 1) & ~(long long)(VG - 1), vb = b & ~(long long)(VG - 1);
        const int ncg = cb > ca ? (int)((cb - ca) >> 2) : 0;
        const int nvg = vb > va ? (int)((vb - va) / VG) : 0;
        {
            const uint4 *g = reinterpret_cast<const uint4 *>(c_col + ca);
            uint4 *s = reinterpret_cast<uint4 *>(s_col);
            for (int i = t; i < ncg; i += 256) s[i] = __ldcg(g + i);
            const uint4 *gv = reinterpret_cast<const uint4 *>(c_val + (size_t)va * VB);
            uint4 *sv = reinterpret_cast<uint4 *>(s_val);
            for (int i = t; i < nvg; i += 256) sv[i] = __ldcg(gv + i);
        }
        // ragged ends (first / last tile of the block only): plain stores straight to the peers
        if (ncg == 0 || ca != a || cb != b || va != a || vb != b) {
            for (long long k = a + t; k < b; k += 256) {
                const bool col_edge = ncg == 0 || k < ca || k >= cb;
                const bool val_edge = nvg == 0 || k < va || k >= vb;
                if (col_edge) {
                    const int c = __ldcg(c_col + k);
                    for (int p = 0; p < peer.n; ++p) peer.col[p][k] = c;
                }
                if (val_edge) {
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
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> async-proxy reads
        __syncthreads();
        if (t == 0) {
            for (int p = 0; p < peer.n; ++p) {
                if (ncg) bulk_store(peer.col[p] + ca, s_col, (unsigned)ncg * 16u);
                if (nvg) bulk_store(static_cast<unsigned char *>(peer.val[p]) + (size_t)va * VB, s_val, (unsigned)nvg * 16u);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        stage ^= 1;
    }
    // all bulk stores of this CTA complete (not just read) before the kernel ends
    if (t == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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

// workspace of the tile hand-off (grow-only), PeerOut's tile fields, and the pusher launch on the push stream.
// Called by spgemm_numeric right before its kernels when peers are set.
int peer_push_begin(nsp_context *ctx, const int *c_col_full, const void *c_val_full, int val_bytes, long long nnz_block)
{
    PeerOut &po = ctx->peer_out;
    po.nnz = nnz_block;
    po.tile0 = po.off >> kTileLog;
    po.ntiles = nnz_block > 0 ? (int)(((po.off + nnz_block - 1) >> kTileLog) - po.tile0 + 1) : 0;
    ctx->push_active = false;
    if (po.n <= 0 || po.ntiles == 0) return 0;
    if ((size_t)po.ntiles > ctx->push_cap) {
        NSP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->push_stream) NSP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->push_stream));
        cudaFree(ctx->d_push_ws);
        ctx->d_push_ws = nullptr;
        ctx->push_cap = 0;
        const size_t cap = (size_t)po.ntiles + (size_t)po.ntiles / 8 + 64;
        NSP_CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_push_ws, sizeof(int) * (2 * cap + 8)));
        NSP_CUDA_TRY(ctx, cudaMemset(ctx->d_push_ws, 0, sizeof(int) * (2 * cap + 8)));
        ctx->push_cap = cap;
    }
    if (!ctx->push_stream) {
        NSP_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->push_stream, cudaStreamNonBlocking));
        NSP_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_push_fork, cudaEventDisableTiming));
        NSP_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_push_join, cudaEventDisableTiming));
    }
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
    const size_t smem = (size_t)2 * (1u << kTileLog) * (4 + val_bytes);
    const long long spin_limit = 20ll * 1000 * 1000 * 1000;      // ~10 s of SM clocks: a hang guard, not a schedule
    const unsigned char *cv = static_cast<const unsigned char *>(c_val_full);
    if (val_bytes == 4) {
        NSP_CUDA_TRY(ctx, cudaFuncSetAttribute(push_tiles_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        push_tiles_kernel<4><<<ctas, 256, smem, ctx->push_stream>>>(c_col_full, cv, spin_limit, po);
    } else {
        NSP_CUDA_TRY(ctx, cudaFuncSetAttribute(push_tiles_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        push_tiles_kernel<8><<<ctas, 256, smem, ctx->push_stream>>>(c_col_full, cv, spin_limit, po);
    }
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
