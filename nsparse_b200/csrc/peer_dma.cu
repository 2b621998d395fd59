// nsparse-b200: multi-GPU allgatherv of C by the COPY ENGINES, overlapped with the numeric phase (the default
// of nsp_spgemm_set_peers; see PeerOut in common.cuh).
//
// Why not SMs: measured on 2 x B200 (profiles/r2_bench_g2_pusher_*.json), a dedicated pusher SM sustains 8-15 GB/s
// of remote stores, with plain stores or with an all-TMA ring alike, so feeding NVLink from "a few SMs of their own"
// does not work, and remote stores from the computing CTAs stall them (round 1: the transfer and the compute
// added up).  The copy engines move 64-256 MB pieces at NVLink speed without any SM.  What they need is (1) big
// contiguous pieces that are FINISHED while the rest of the block is still being computed and (2) someone to
// start them:
//   (1) tiles of 2^24 .. 2^26 entries, and the rows of the heavy class are processed tile by tile (by the tile of
//       their first entry, heaviest first inside a tile) instead of heaviest first over the whole block
//       (order_rows_by_tile); the light classes, a few per cent of the time, run before the heavy launch;
//   (2) the numeric kernels count finished entries per tile; the thread that completes a tile raises its flag in
//       host-mapped memory; the host thread that called the numeric phase polls the flags while the kernels run
//       and hands every finished tile to cudaMemcpyAsync, one stream per peer (peer_dma_drive).  The call returns
//       when every tile has been issued; the context's stream then waits for the copy streams.
#include <stdlib.h>

#include <chrono>
#include <thread>

#include <cub/cub.cuh>

#include "context.h"
#include "spgemm_plan.h"

namespace nsp {

__global__ void dma_init_kernel(int *tile_cnt, int ntiles)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ntiles) tile_cnt[i] = 0;
}

__global__ void __launch_bounds__(256)
tile_keys_kernel(const int *__restrict__ row_perm, int n, const long long *__restrict__ c_rpt, long long off, int tile_log,
                 unsigned long long *__restrict__ keys)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int rid = row_perm[i];
    const long long s = c_rpt[rid], e = c_rpt[rid + 1];
    const unsigned long long tile = (unsigned long long)((off + s) >> tile_log);
    const unsigned long long cnt = (unsigned long long)(e - s);
    keys[i] = (tile << 32) | (0xffffffffull - (cnt > 0xffffffffull ? 0xffffffffull : cnt));
}

static int sort_reserve(nsp_context *ctx, int n, size_t *temp_out)
{
    nsp_dma_push &dp = ctx->dma;
    size_t temp = 0;
    unsigned long long *k0 = nullptr, *k1 = nullptr;
    int *v0 = nullptr, *v1 = nullptr;
    NSP_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, temp, k0, k1, v0, v1, n, 0, 64, ctx->stream));
    const size_t need = (size_t)n * (8 + 8 + 4) + temp + 1024;
    if (need > dp.sort_bytes) {
        NSP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(dp.d_sort);
        dp.d_sort = nullptr;
        dp.sort_bytes = 0;
        NSP_CUDA_TRY(ctx, cudaMalloc((void **)&dp.d_sort, need + need / 4));
        dp.sort_bytes = need + need / 4;
    }
    if (temp_out) *temp_out = temp;
    return 0;
}

// row_perm[0 .. n): the rows of the heavy class, re-ordered by (tile of the row's first entry, entries descending)
int order_rows_by_tile(nsp_context *ctx, int *row_perm, int n, const long long *c_rpt, long long off, int tile_log)
{
    if (n <= 1) return 0;
    nsp_dma_push &dp = ctx->dma;
    size_t temp = 0;
    if (sort_reserve(ctx, n, &temp) != 0) return -1;
    unsigned long long *k0 = nullptr, *k1 = nullptr;
    int *v1 = nullptr;
    char *p = dp.d_sort;
    k0 = reinterpret_cast<unsigned long long *>(p);
    k1 = k0 + n;
    v1 = reinterpret_cast<int *>(k1 + n);
    void *tmp = reinterpret_cast<void *>((reinterpret_cast<uintptr_t>(v1 + n) + 255) & ~uintptr_t(255));
    tile_keys_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(row_perm, n, c_rpt, off, tile_log, k0);
    NSP_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(tmp, temp, k0, k1, row_perm, v1, n, 0, 64, ctx->stream));
    NSP_CUDA_TRY(ctx, cudaMemcpyAsync(row_perm, v1, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->launches += 3;
    NSP_CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}

int peer_dma_reserve(nsp_context *ctx, long long ntiles, int npeers, int max_rows)
{
    nsp_dma_push &dp = ctx->dma;
    if (max_rows > 1 && sort_reserve(ctx, max_rows, nullptr) != 0) return -1;
    if ((size_t)ntiles > dp.cap) {
        NSP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(dp.d_tile_cnt);
        if (dp.h_done) cudaFreeHost(dp.h_done);
        dp.d_tile_cnt = nullptr;
        dp.h_done = nullptr;
        dp.cap = 0;
        const size_t cap = (size_t)ntiles + 64;
        NSP_CUDA_TRY(ctx, cudaMalloc((void **)&dp.d_tile_cnt, sizeof(int) * cap));
        NSP_CUDA_TRY(ctx, cudaHostAlloc((void **)&dp.h_done, sizeof(int) * cap, cudaHostAllocMapped | cudaHostAllocPortable));
        NSP_CUDA_TRY(ctx, cudaHostGetDevicePointer((void **)&dp.d_done, dp.h_done, 0));
        dp.cap = cap;
    }
    for (int p = 0; p < npeers; ++p) {
        if (!dp.copy_st[p]) {
            NSP_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&dp.copy_st[p], cudaStreamNonBlocking));
            NSP_CUDA_TRY(ctx, cudaEventCreateWithFlags(&dp.ev_copy[p], cudaEventDisableTiming));
        }
    }
    return 0;
}

// tile size of a block of `nnz` entries: ~64 .. 128 tiles, at least 2^20 and at most 2^26 entries each
int dma_tile_log(long long nnz)
{
    int lg = 20;
    while (lg < 26 && (nnz >> lg) > 128) ++lg;
    return lg;
}

// sets the tile fields of ctx->peer_out for the copy-engine gather and clears the counters / flags
int peer_dma_begin(nsp_context *ctx, long long nnz_block)
{
    PeerOut &po = ctx->peer_out;
    nsp_dma_push &dp = ctx->dma;
    po.nnz = nnz_block;
    po.tile_log = ctx->opt_dma_tile_log > 0 ? (int)ctx->opt_dma_tile_log : dma_tile_log(nnz_block);
    po.tile0 = po.off >> po.tile_log;
    po.ntiles = nnz_block > 0 ? (int)(((po.off + nnz_block - 1) >> po.tile_log) - po.tile0 + 1) : 0;
    po.queue = nullptr;
    po.q_ctl = nullptr;
    dp.active = false;
    if (po.n <= 0 || po.ntiles == 0) return 0;
    if (peer_dma_reserve(ctx, po.ntiles, po.n, 0) != 0) return -1;
    po.tile_cnt = dp.d_tile_cnt;
    po.done = dp.d_done;
    for (int t = 0; t < po.ntiles; ++t) dp.h_done[t] = 0;
    dma_init_kernel<<<(po.ntiles + 255) / 256, 256, 0, ctx->stream>>>(po.tile_cnt, po.ntiles);
    ctx->launches += 1;
    NSP_CUDA_TRY(ctx, cudaGetLastError());
    dp.active = true;
    ctx->last_push = po;
    return 0;
}

// Polls the tile flags while the numeric kernels run and gives every finished tile to the copy engines; returns
// when all tiles are issued (the context's stream then waits for the copies) or nothing moved for ten seconds.
int peer_dma_drive(nsp_context *ctx, const int *c_col_full, const void *c_val_full, int val_bytes)
{
    nsp_dma_push &dp = ctx->dma;
    if (!dp.active) return 0;
    dp.active = false;
    const PeerOut &po = ctx->peer_out;
    const int nt = po.ntiles;
    std::vector<char> sent((size_t)nt, 0);
    int nsent = 0, first = 0;
    auto last = std::chrono::steady_clock::now();
    // NSP_DMA_TRACE=1: when (ms after the start of the polling) the tiles were seen finished and when the kernels /
    // the copies ended -- the timeline of the overlap without a profiler
    const bool trace = getenv("NSP_DMA_TRACE") != nullptr;
    const auto t_begin = last;
    std::vector<float> seen_ms;
    double kernels_done_ms = -1;
    auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(t - t_begin).count(); };
    volatile int *done = dp.h_done;
    const char *cv = static_cast<const char *>(c_val_full);
    while (nsent < nt) {
        bool any = false;
        for (int t = first; t < nt; ++t) {
            if (sent[t] || !done[t]) continue;
            // a run of adjacent finished tiles goes out as one copy per peer and array
            int t1 = t;
            while (t1 + 1 < nt && !sent[t1 + 1] && done[t1 + 1]) ++t1;
            const long long lo = (po.tile0 + t) << po.tile_log, hi = (po.tile0 + t1 + 1) << po.tile_log;
            const long long a = lo > po.off ? lo : po.off, b = hi < po.off + po.nnz ? hi : po.off + po.nnz;
            for (int p = 0; p < po.n; ++p) {
                NSP_CUDA_TRY(ctx, cudaMemcpyAsync(po.col[p] + a, c_col_full + a, sizeof(int) * (size_t)(b - a), cudaMemcpyDeviceToDevice,
                                                  dp.copy_st[p]));
                NSP_CUDA_TRY(ctx, cudaMemcpyAsync(static_cast<char *>(po.val[p]) + (size_t)a * val_bytes, cv + (size_t)a * val_bytes,
                                                  (size_t)val_bytes * (size_t)(b - a), cudaMemcpyDeviceToDevice, dp.copy_st[p]));
            }
            for (int k = t; k <= t1; ++k) sent[k] = 1;
            nsent += t1 - t + 1;
            if (trace)
                for (int k = t; k <= t1; ++k) seen_ms.push_back((float)ms_since(std::chrono::steady_clock::now()));
            any = true;
            t = t1;
        }
        while (first < nt && sent[first]) ++first;
        if (trace && kernels_done_ms < 0 && cudaStreamQuery(ctx->stream) == cudaSuccess) kernels_done_ms = ms_since(std::chrono::steady_clock::now());
        if (any) {
            last = std::chrono::steady_clock::now();
        } else {
            if (cudaStreamQuery(ctx->stream) == cudaSuccess && std::chrono::steady_clock::now() - last > std::chrono::seconds(1)) {
                // every kernel of the product has finished and flags are still missing: an accounting error
                bool missing = false;
                for (int t = first; t < nt; ++t) missing = missing || (!sent[t] && !done[t]);
                if (missing) return ctx->fail(-1, "multi-GPU allgatherv: " + std::to_string(nt - nsent) + " of " + std::to_string(nt) +
                                                      " tiles of C were never completed by the numeric kernels");
            }
            if (std::chrono::steady_clock::now() - last > std::chrono::seconds(20))
                return ctx->fail(-1, "multi-GPU allgatherv: no tile of C finished for 20 s");
            std::this_thread::yield();
        }
    }
    for (int p = 0; p < po.n; ++p) {
        NSP_CUDA_TRY(ctx, cudaEventRecord(dp.ev_copy[p], dp.copy_st[p]));
        NSP_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, dp.ev_copy[p], 0));
    }
    if (trace && !seen_ms.empty()) {
        const double issued_ms = ms_since(std::chrono::steady_clock::now());
        if (kernels_done_ms < 0) {
            // (the stream now also waits for the copies: time the kernels by the last flag instead)
            kernels_done_ms = seen_ms.back();
        }
        for (int p = 0; p < po.n; ++p) cudaStreamSynchronize(dp.copy_st[p]);
        const double copies_ms = ms_since(std::chrono::steady_clock::now());
        const size_t n = seen_ms.size();
        fprintf(stderr, "[nsp dma] dev %d: %d tiles of 2^%d entries; flags seen at %.1f / %.1f / %.1f / %.1f / %.1f ms (first, 25%%, 50%%, 75%%, "
                        "last); kernels done %.1f ms; all copies issued %.1f ms, landed %.1f ms\n",
                ctx->device, nt, po.tile_log, seen_ms[0], seen_ms[n / 4], seen_ms[n / 2], seen_ms[3 * n / 4], seen_ms[n - 1], kernels_done_ms,
                issued_ms, copies_ms);
    }
    return 0;
}

void peer_dma_destroy(nsp_context *ctx)
{
    nsp_dma_push &dp = ctx->dma;
    for (int p = 0; p < kMaxPeerOut; ++p) {
        if (dp.copy_st[p]) {
            cudaStreamSynchronize(dp.copy_st[p]);
            cudaStreamDestroy(dp.copy_st[p]);
            cudaEventDestroy(dp.ev_copy[p]);
        }
    }
    cudaFree(dp.d_tile_cnt);
    if (dp.h_done) cudaFreeHost(dp.h_done);
    cudaFree(dp.d_sort);
    dp = nsp_dma_push();
}

}  // namespace nsp
