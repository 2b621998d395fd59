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
//   (3) With many peers the block has to leave N-1 times and the copy engines of one GPU move ~0.4 TB/s in total
//       (8 GPUs: profiles/r2_bench_c2_gpus8_compute_balanced.json), the whole GPU's SMs ~0.7 TB/s
//       (profiles/r1_probe_nvlink_multicast_2gpu.txt) -- ~5 GB/s per SM, which is why a FEW SMs cannot do it.  So the
//       copy-engine queue is kept short (kSlots batches per peer), and from the moment the numeric kernels have
//       ended -- all SMs idle -- what is finished and not yet queued is stored to the peers by push_tiles_sm_kernel,
//       which reads every piece once for all peers, while the copy engines keep draining their queue.
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <thread>
#include <vector>

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
            for (int k = 0; k < nsp_dma_push::kSlots; ++k)
                NSP_CUDA_TRY(ctx, cudaEventCreateWithFlags(&dp.ev_slot[p][k], cudaEventDisableTiming));
        }
    }
    for (int k = 0; k < nsp_dma_push::kSmSlots; ++k)
        if (!dp.ev_sm[k]) NSP_CUDA_TRY(ctx, cudaEventCreateWithFlags(&dp.ev_sm[k], cudaEventDisableTiming));
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

// [a, b) of the block's col and val arrays to every peer by SM stores: each 16-byte piece is loaded once and stored
// to all peers (the copy engines read it once per peer).  Source and destinations are the same offsets of buffers
// whose bases are 256-byte aligned, so one head / body / tail split serves all of them.
struct PushPtrs {
    char *col[kMaxPeerOut];
    char *val[kMaxPeerOut];
};

template <int U>
__device__ __forceinline__ void push_span(const char *__restrict__ src, char *const *dst, int np, size_t byte_lo, size_t nbytes,
                                          size_t tid, size_t nth)
{
    const size_t mis = (16 - (byte_lo & 15)) & 15;
    const size_t head = mis < nbytes ? mis : nbytes;
    const size_t body = (nbytes - head) & ~size_t(15);
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src + byte_lo + head);
    const size_t n16 = body / 16;
    size_t i = tid;
    // U loads in flight per thread before the first store leaves
    for (; i + (U - 1) * nth < n16; i += U * nth) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w)
                         : "l"(s4 + i + u * nth));
        for (int p = 0; p < np; ++p) {
            uint4 *d4 = reinterpret_cast<uint4 *>(dst[p] + byte_lo + head);
#pragma unroll
            for (int u = 0; u < U; ++u) d4[i + u * nth] = v[u];
        }
    }
    for (; i < n16; i += nth) {
        const uint4 v = s4[i];
        for (int p = 0; p < np; ++p) reinterpret_cast<uint4 *>(dst[p] + byte_lo + head)[i] = v;
    }
    const size_t tail0 = head + body;
    const size_t nsmall = head / 4 + (nbytes - tail0) / 4;
    for (size_t k = tid; k < nsmall; k += nth) {
        const size_t off = byte_lo + (k < head / 4 ? k * 4 : tail0 + (k - head / 4) * 4);
        const unsigned v = *reinterpret_cast<const unsigned *>(src + off);
        for (int p = 0; p < np; ++p) *reinterpret_cast<unsigned *>(dst[p] + off) = v;
    }
}

__global__ void __launch_bounds__(512)
push_tiles_sm_kernel(PushPtrs pp, int np, const char *__restrict__ col, const char *__restrict__ val, long long a, long long b,
                     int val_bytes)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nth = (size_t)gridDim.x * blockDim.x;
    push_span<4>(col, pp.col, np, (size_t)a * 4, (size_t)(b - a) * 4, tid, nth);
    push_span<4>(val, pp.val, np, (size_t)a * val_bytes, (size_t)(b - a) * val_bytes, tid, nth);
}

// Polls the tile flags while the numeric kernels run and gives every finished tile to the copy engines, at most
// kSlots batches queued per peer; once the kernels have ended, what is finished and not yet queued leaves through
// push_tiles_sm_kernel on the context's stream, next to the copy engines.  (Measured on 8 x B200: seven peer streams
// of copy-engine traffic leave one GPU at ~0.4 TB/s, SM stores at ~0.7 TB/s -- profiles/r1_probe_nvlink_multicast_2gpu.txt;
// while the kernels run the SMs are not available, afterwards they are idle.)  Returns when all tiles are issued
// (the context's stream then waits for the copies) or nothing moved for 20 s.
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
    std::vector<float> seen_ms((size_t)nt, -1.f);
    double kernels_done_ms = -1;
    auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(t - t_begin).count(); };
    volatile int *done = dp.h_done;
    const char *cv = static_cast<const char *>(c_val_full);
    const size_t tile_bytes = ((size_t)1 << po.tile_log) * (size_t)(4 + val_bytes);
    // a batch: up to 256 MB (copy engines) / 512 MB (SM stores) per peer
    const int ce_run = (int)std::max<size_t>(1, ((size_t)256 << 20) / tile_bytes);
    const int sm_run = (int)std::max<size_t>(1, ((size_t)512 << 20) / tile_bytes);
    const int ce_slots = ctx->opt_gather_sm == 2 ? 0 : nsp_dma_push::kSlots;
    const bool sm_ok = ctx->opt_gather_sm != 0;
    int ce_head = 0, ce_inflight = 0, sm_head = 0, sm_inflight = 0;
    bool kernels_done = false;
    long long ce_tiles = 0, sm_tiles = 0;
    PushPtrs pp;
    for (int p = 0; p < po.n; ++p) {
        pp.col[p] = reinterpret_cast<char *>(po.col[p]);
        pp.val[p] = static_cast<char *>(po.val[p]);
    }
    while (nsent < nt) {
        // (before any SM launch of this call is queued behind them)
        if (!kernels_done && cudaStreamQuery(ctx->stream) == cudaSuccess) {
            kernels_done = true;
            const auto now = std::chrono::steady_clock::now();
            kernels_done_ms = ms_since(now);
            dp.last_kernel_ms = std::chrono::duration<double, std::milli>(now - dp.t_numeric).count();
        }
        // retire finished batches
        while (ce_inflight > 0) {
            const int slot = (ce_head - ce_inflight + 2 * nsp_dma_push::kSlots) % nsp_dma_push::kSlots;
            bool fin = true;
            for (int p = 0; p < po.n && fin; ++p) fin = cudaEventQuery(dp.ev_slot[p][slot]) == cudaSuccess;
            if (!fin) break;
            --ce_inflight;
        }
        while (sm_inflight > 0) {
            const int slot = (sm_head - sm_inflight + 2 * nsp_dma_push::kSmSlots) % nsp_dma_push::kSmSlots;
            if (cudaEventQuery(dp.ev_sm[slot]) != cudaSuccess) break;
            --sm_inflight;
        }
        bool any = false;
        for (int t = first; t < nt; ++t) {
            if (sent[t] || !done[t]) continue;
            const bool by_ce = ce_inflight < ce_slots;
            const bool by_sm = !by_ce && kernels_done && (sm_ok || ce_slots == 0) && sm_inflight < nsp_dma_push::kSmSlots;
            if (!by_ce && !by_sm) break;
            // a run of adjacent finished tiles goes out as one batch
            const int run = by_ce ? ce_run : sm_run;
            int t1 = t;
            while (t1 + 1 < nt && t1 + 1 - t < run && !sent[t1 + 1] && done[t1 + 1]) ++t1;
            const long long lo = (po.tile0 + t) << po.tile_log, hi = (po.tile0 + t1 + 1) << po.tile_log;
            const long long a = lo > po.off ? lo : po.off, b = hi < po.off + po.nnz ? hi : po.off + po.nnz;
            if (by_ce) {
                const int slot = ce_head % nsp_dma_push::kSlots;
                for (int p = 0; p < po.n; ++p) {
                    NSP_CUDA_TRY(ctx, cudaMemcpyAsync(po.col[p] + a, c_col_full + a, sizeof(int) * (size_t)(b - a), cudaMemcpyDeviceToDevice,
                                                      dp.copy_st[p]));
                    NSP_CUDA_TRY(ctx, cudaMemcpyAsync(static_cast<char *>(po.val[p]) + (size_t)a * val_bytes, cv + (size_t)a * val_bytes,
                                                      (size_t)val_bytes * (size_t)(b - a), cudaMemcpyDeviceToDevice, dp.copy_st[p]));
                    NSP_CUDA_TRY(ctx, cudaEventRecord(dp.ev_slot[p][slot], dp.copy_st[p]));
                }
                ce_head = (ce_head + 1) % nsp_dma_push::kSlots;
                ++ce_inflight;
                ce_tiles += t1 - t + 1;
            } else {
                const int slot = sm_head % nsp_dma_push::kSmSlots;
                push_tiles_sm_kernel<<<ctx->sm_count * 4, 512, 0, ctx->stream>>>(pp, po.n, reinterpret_cast<const char *>(c_col_full), cv, a, b,
                                                                                  val_bytes);
                {
                    // (the queries above may have left cudaErrorNotReady behind: not an error)
                    const cudaError_t e = cudaGetLastError();
                    if (e != cudaSuccess && e != cudaErrorNotReady) NSP_CUDA_TRY(ctx, e);
                }
                NSP_CUDA_TRY(ctx, cudaEventRecord(dp.ev_sm[slot], ctx->stream));
                ctx->launches += 1;
                sm_head = (sm_head + 1) % nsp_dma_push::kSmSlots;
                ++sm_inflight;
                sm_tiles += t1 - t + 1;
            }
            for (int k = t; k <= t1; ++k) sent[k] = 1;
            nsent += t1 - t + 1;
            any = true;
            t = t1;
        }
        if (trace) {
            float now = -1.f;
            for (int t = first; t < nt; ++t)
                if (seen_ms[t] < 0 && done[t]) {
                    if (now < 0) now = (float)ms_since(std::chrono::steady_clock::now());
                    seen_ms[t] = now;
                }
        }
        while (first < nt && sent[first]) ++first;
        if (any) {
            last = std::chrono::steady_clock::now();
        } else {
            if (kernels_done && ce_inflight == 0 && sm_inflight == 0 && std::chrono::steady_clock::now() - last > std::chrono::seconds(1)) {
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
    dp.last_ce_tiles = ce_tiles;
    dp.last_sm_tiles = sm_tiles;
    // (the last tile went to the copy engines the moment its flag came up: the kernels end about now)
    if (!kernels_done) dp.last_kernel_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - dp.t_numeric).count();
    {
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess && e != cudaErrorNotReady) NSP_CUDA_TRY(ctx, e);
    }
    for (int p = 0; p < po.n; ++p) {
        NSP_CUDA_TRY(ctx, cudaEventRecord(dp.ev_copy[p], dp.copy_st[p]));
        NSP_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, dp.ev_copy[p], 0));
    }
    if (trace && nt > 0) {
        const double issued_ms = ms_since(std::chrono::steady_clock::now());
        std::vector<float> sm(seen_ms);
        std::sort(sm.begin(), sm.end());
        if (kernels_done_ms < 0) kernels_done_ms = sm.back();   // (the stream now also waits for the copies)
        for (int p = 0; p < po.n; ++p) cudaStreamSynchronize(dp.copy_st[p]);
        const double copies_ms = ms_since(std::chrono::steady_clock::now());
        cudaStreamSynchronize(ctx->stream);
        const double all_ms = ms_since(std::chrono::steady_clock::now());
        const size_t n = sm.size();
        fprintf(stderr, "[nsp dma] dev %d: %d tiles of 2^%d entries (%lld by copy engines, %lld by SM stores); flags seen at %.1f / %.1f / "
                        "%.1f / %.1f / %.1f ms (first, 25%%, 50%%, 75%%, last); kernels done %.1f ms; all issued %.1f ms, copy engines "
                        "done %.1f ms, everything %.1f ms\n",
                ctx->device, nt, po.tile_log, ce_tiles, sm_tiles, sm[0], sm[n / 4], sm[n / 2], sm[3 * n / 4], sm[n - 1], kernels_done_ms,
                issued_ms, copies_ms, all_ms);
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
            for (int k = 0; k < nsp_dma_push::kSlots; ++k) cudaEventDestroy(dp.ev_slot[p][k]);
        }
    }
    for (int k = 0; k < nsp_dma_push::kSmSlots; ++k)
        if (dp.ev_sm[k]) cudaEventDestroy(dp.ev_sm[k]);
    cudaFree(dp.d_tile_cnt);
    if (dp.h_done) cudaFreeHost(dp.h_done);
    cudaFree(dp.d_sort);
    dp = nsp_dma_push();
}

}  // namespace nsp
