// nsparse-b200: the extern "C" boundary (include/nsparse_b200.h).
#include "../../include/nsparse_b200.h"

#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "context.h"

using nsp::context_create;
using nsp::context_destroy;

#define NSP_REQUIRE_CTX(ctx) \
    if (!(ctx)) return NSP_ERR_ARG; \
    cudaSetDevice((ctx)->device)

// ---- fold of a device CSR (parity checks at sizes no host copy is wanted for) ---------------------------
__device__ __forceinline__ unsigned long long fold_mix(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

template <typename real>
__global__ void __launch_bounds__(256)
fold_csr_kernel(const long long *__restrict__ rpt, const int *__restrict__ col, const real *__restrict__ val, int M,
                long long nnz, unsigned long long *__restrict__ out_h, double *__restrict__ out_f)
{
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    unsigned long long hr = 0, hc = 0;
    double sv = 0.0, sw = 0.0;
    for (long long i = tid; i <= M; i += nth) hr += fold_mix((unsigned long long)i * 0x100000001B3ull ^ (unsigned long long)rpt[i]);
    for (long long i = tid; i < nnz; i += nth) {
        const int c = col[i];
        hc += fold_mix((unsigned long long)i * 0x100000001B3ull ^ (unsigned long long)(unsigned)c);
        const double v = (double)val[i];
        sv += v;
        sw += v * (double)((c % 1021) + 1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        hr += __shfl_xor_sync(0xffffffffu, hr, o);
        hc += __shfl_xor_sync(0xffffffffu, hc, o);
        sv += __shfl_xor_sync(0xffffffffu, sv, o);
        sw += __shfl_xor_sync(0xffffffffu, sw, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out_h, hr);
        atomicAdd(out_h + 1, hc);
        atomicAdd(out_f, sv);
        atomicAdd(out_f + 1, sw);
    }
}

template <typename real>
static int fold_csr(nsp_context *ctx, int M, long long nnz, const long long *d_rpt64, const int *d_col, const real *d_val,
                    unsigned long long *h_hash2, double *h_sum2)
{
    if (M < 0 || nnz < 0 || !d_rpt64 || !h_hash2 || !h_sum2) return ctx->fail(NSP_ERR_ARG, "nsp_csr_fold: bad argument");
    unsigned long long *d = nullptr;
    NSP_CUDA_TRY(ctx, cudaMalloc((void **)&d, 32));
    NSP_CUDA_TRY(ctx, cudaMemsetAsync(d, 0, 32, ctx->stream));
    fold_csr_kernel<real><<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d_rpt64, d_col, d_val, M, nnz, d, reinterpret_cast<double *>(d + 2));
    ctx->launches += 1;
    unsigned long long h[4];
    cudaError_t e = cudaMemcpyAsync(h, d, 32, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    NSP_CUDA_TRY(ctx, e);
    h_hash2[0] = h[0];
    h_hash2[1] = h[1];
    memcpy(h_sum2, h + 2, 16);
    return 0;
}

extern "C" {

int nsp_create(nsp_context **ctx, int device) { return context_create(ctx, device); }
int nsp_destroy(nsp_context *ctx) { return context_destroy(ctx); }
const char *nsp_last_error(nsp_context *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int nsp_set_stream(nsp_context *ctx, void *cuda_stream)
{
    NSP_REQUIRE_CTX(ctx);
    if (ctx->stream == (cudaStream_t)cuda_stream) return 0;      // bindings call this before every op: no sync then
    NSP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = (cudaStream_t)cuda_stream;
    return 0;
}

int nsp_sync(nsp_context *ctx)
{
    NSP_REQUIRE_CTX(ctx);
    NSP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int nsp_set_option(nsp_context *ctx, const char *name, long long value)
{
    if (!ctx || !name) return NSP_ERR_ARG;
    if (!strcmp(name, "sym_bitmap_min")) ctx->opt_sym_bitmap_min = value;
    else if (!strcmp(name, "num_bitmap_min")) ctx->opt_num_bitmap_min = value;
    else if (!strcmp(name, "profile")) ctx->profile = value != 0;
    else if (!strcmp(name, "debug")) ctx->opt_debug = value;
    else if (!strcmp(name, "no_vec")) ctx->opt_no_vec = value;
    else if (!strcmp(name, "no_fork")) ctx->opt_no_fork = value;
    else if (!strcmp(name, "num_cap")) ctx->opt_num_cap = value;
    else if (!strcmp(name, "push_sms")) ctx->opt_push_sms = value;
    else if (!strcmp(name, "no_seg")) ctx->opt_no_seg = value;
    else if (!strcmp(name, "no_flat")) ctx->opt_no_flat = value;
    else if (!strcmp(name, "gather_tma")) ctx->opt_gather_tma = value;
    else if (!strcmp(name, "dma_tile_log")) ctx->opt_dma_tile_log = value;
    else if (!strcmp(name, "gather_sm")) ctx->opt_gather_sm = value;
    else if (!strcmp(name, "hash_order")) ctx->opt_hash_order = value;
    else if (!strcmp(name, "no_ranges")) ctx->opt_no_ranges = value;
    else if (!strcmp(name, "sort")) ctx->opt_unsorted = value == 0;
    else if (!strcmp(name, "phase_timing")) {
        // value 1: start accumulating; value 2: print the totals (cycles summed over CTAs) and reset
        if (value == 2 && ctx->d_phase) {
            long long h[12];
            cudaDeviceSynchronize();
            cudaMemcpy(h, ctx->d_phase, sizeof(h), cudaMemcpyDeviceToHost);
            cudaMemset(ctx->d_phase, 0, sizeof(h));
            static const char *nm[12] = {"row+clear", "mark stage", "mark run", "rank", "scan+bounds", "emit", "copy cols",
                                         "chunk stage", "value run", "copy vals", "red mode", "-"};
            long long tot = 0;
            for (int i = 0; i < 12; ++i) tot += h[i];
            for (int i = 0; i < 11; ++i)
                printf("   phase %-12s %8.3f Mcyc/CTA  %5.1f%%\n", nm[i], h[i] / 1e6 / ctx->sm_count, 100.0 * h[i] / (tot ? tot : 1));
            fflush(stdout);
        }
        ctx->opt_phase_timing = value != 0;
    }
    else if (!strcmp(name, "sym_window_shift")) ctx->opt_sym_window_shift = value;
    else if (!strcmp(name, "num_window_shift")) ctx->opt_num_window_shift = value;
    else
        return ctx->fail(NSP_ERR_ARG, std::string("unknown option ") + name);
    return 0;
}

// ---- allgatherv over peer memory ------------------------------------------------------------------
constexpr int kMaxPeers = 16;
struct PeerList {
    char *base[kMaxPeers];
};

// Source and destinations share the byte offset modulo 16 (all bases are 256-byte aligned allocations),
// so one head / 16-byte body / tail split serves every copy.  Stores to a peer travel over NVLink as
// posted writes; the loads are the only dependent step.
__global__ void __launch_bounds__(256)
push_to_peers_kernel(PeerList peers, int npeers, size_t byte_offset, const char *__restrict__ src, size_t nbytes)
{
    const size_t mis = (16 - (byte_offset & 15)) & 15;
    const size_t head = mis < nbytes ? mis : nbytes;
    const size_t body = (nbytes - head) & ~size_t(15);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nth = (size_t)gridDim.x * blockDim.x;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src + head);
    for (size_t i = tid; i < body / 16; i += nth) {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(s4 + i));
        for (int p = 0; p < npeers; ++p)
            reinterpret_cast<uint4 *>(peers.base[p] + byte_offset + head)[i] = v;
    }
    // head and tail in 4-byte units
    const size_t tail0 = head + body;
    const size_t nsmall = head / 4 + (nbytes - tail0) / 4;
    for (size_t i = tid; i < nsmall; i += nth) {
        const size_t off = i < head / 4 ? i * 4 : tail0 + (i - head / 4) * 4;
        const unsigned v = *reinterpret_cast<const unsigned *>(src + off);
        for (int p = 0; p < npeers; ++p) *reinterpret_cast<unsigned *>(peers.base[p] + byte_offset + off) = v;
    }
}

int nsp_peer_alloc(nsp_context *ctx, size_t bytes, void **d_ptr, unsigned char *handle64)
{
    NSP_REQUIRE_CTX(ctx);
    if (!d_ptr || !handle64) return ctx->fail(NSP_ERR_ARG, "nsp_peer_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    NSP_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    NSP_CUDA_TRY(ctx, cudaMalloc(d_ptr, bytes ? bytes : 16));
    cudaIpcMemHandle_t h;
    NSP_CUDA_TRY(ctx, cudaIpcGetMemHandle(&h, *d_ptr));
    memcpy(handle64, &h, 64);
    return 0;
}

int nsp_peer_open(nsp_context *ctx, const unsigned char *handle64, void **d_ptr)
{
    NSP_REQUIRE_CTX(ctx);
    if (!d_ptr || !handle64) return ctx->fail(NSP_ERR_ARG, "nsp_peer_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    NSP_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    NSP_CUDA_TRY(ctx, cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int nsp_peer_close(nsp_context *ctx, void *d_ptr)
{
    NSP_REQUIRE_CTX(ctx);
    if (d_ptr) NSP_CUDA_TRY(ctx, cudaIpcCloseMemHandle(d_ptr));
    return 0;
}

int nsp_peer_free(nsp_context *ctx, void *d_ptr)
{
    NSP_REQUIRE_CTX(ctx);
    if (d_ptr) NSP_CUDA_TRY(ctx, cudaFree(d_ptr));
    return 0;
}

int nsp_copy_async(nsp_context *ctx, void *d_dst, const void *d_src, size_t bytes, void *cuda_stream)
{
    NSP_REQUIRE_CTX(ctx);
    if (bytes == 0) return 0;
    if (!d_dst || !d_src) return ctx->fail(NSP_ERR_ARG, "nsp_copy_async: bad argument");
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    NSP_CUDA_TRY(ctx, cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDefault, st));
    return 0;
}

int nsp_spgemm_set_peers(nsp_context *ctx, int npeers, void *const *d_peer_col, void *const *d_peer_val,
                         long long elem_offset)
{
    NSP_REQUIRE_CTX(ctx);
    if (npeers < 0 || npeers > nsp::kMaxPeerOut || (npeers > 0 && (!d_peer_col || !d_peer_val)) || elem_offset < 0)
        return ctx->fail(NSP_ERR_ARG, "nsp_spgemm_set_peers: bad argument (at most 7 peers)");
    nsp::PeerOut po;
    po.n = npeers;
    po.off = elem_offset;
    for (int p = 0; p < npeers; ++p) {
        po.col[p] = static_cast<int *>(d_peer_col[p]);
        po.val[p] = d_peer_val[p];
    }
    ctx->peer_out = po;
    return 0;
}

int nsp_spgemm_peers_stats(nsp_context *ctx, long long *h_copy_engine_tiles, long long *h_sm_tiles, double *h_kernel_ms)
{
    NSP_REQUIRE_CTX(ctx);
    if (h_copy_engine_tiles) *h_copy_engine_tiles = ctx->dma.last_ce_tiles;
    if (h_sm_tiles) *h_sm_tiles = ctx->dma.last_sm_tiles;
    if (h_kernel_ms) *h_kernel_ms = ctx->dma.last_kernel_ms;
    return 0;
}

int nsp_spgemm_peers_status(nsp_context *ctx, int *h_error)
{
    NSP_REQUIRE_CTX(ctx);
    if (!h_error) return ctx->fail(NSP_ERR_ARG, "nsp_spgemm_peers_status: bad argument");
    *h_error = 0;
    if (!ctx->d_push_ws) return 0;
    NSP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    int ctl[3] = {0, 0, 0};
    NSP_CUDA_TRY(ctx, cudaMemcpy(ctl, ctx->d_push_ws, sizeof(ctl), cudaMemcpyDeviceToHost));
    if (ctl[2]) {
        *h_error = 1;
        NSP_CUDA_TRY(ctx, cudaMemset(ctx->d_push_ws + 2, 0, sizeof(int)));
        // which tiles never filled up (diagnosis: a kernel that did not count its entries)
        const nsp::PeerOut &po = ctx->last_push;
        std::string detail;
        if (po.ntiles > 0 && po.tile_cnt) {
            std::vector<int> cnt((size_t)po.ntiles);
            cudaMemcpy(cnt.data(), po.tile_cnt, sizeof(int) * (size_t)po.ntiles, cudaMemcpyDeviceToHost);
            int shown = 0, bad = 0;
            for (int t = 0; t < po.ntiles; ++t) {
                const int want = nsp::tile_len(po, t);
                if (cnt[t] != want) {
                    ++bad;
                    if (shown < 6) {
                        detail += " tile " + std::to_string(t) + ": " + std::to_string(cnt[t]) + "/" + std::to_string(want);
                        ++shown;
                    }
                }
            }
            detail = "; " + std::to_string(bad) + " of " + std::to_string(po.ntiles) + " tiles incomplete:" + detail;
        }
        return ctx->fail(NSP_ERR_CUDA, "multi-GPU allgatherv: the pusher kernel gave up waiting for a tile of C (" +
                                           std::to_string(ctl[0]) + " tiles were published" + detail + ")");
    }
    return 0;
}

// The same through an NVSwitch multicast address: ONE multimem.st per 16 bytes, the switch replicates the
// write into the buffer of every GPU bound to the multicast object (NVLS).
__global__ void __launch_bounds__(256)
push_multicast_kernel(char *mc_base, size_t byte_offset, const char *__restrict__ src, size_t nbytes)
{
    const size_t mis = (16 - (byte_offset & 15)) & 15;
    const size_t head = mis < nbytes ? mis : nbytes;
    const size_t body = (nbytes - head) & ~size_t(15);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nth = (size_t)gridDim.x * blockDim.x;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src + head);
    char *d = mc_base + byte_offset + head;
    for (size_t i = tid; i < body / 16; i += nth) {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(s4 + i));
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(d + i * 16),
                     "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)),
                     "f"(__uint_as_float(v.w))
                     : "memory");
    }
    const size_t tail0 = head + body;
    const size_t nsmall = head / 4 + (nbytes - tail0) / 4;
    for (size_t i = tid; i < nsmall; i += nth) {
        const size_t off = i < head / 4 ? i * 4 : tail0 + (i - head / 4) * 4;
        const unsigned v = *reinterpret_cast<const unsigned *>(src + off);
        asm volatile("multimem.st.relaxed.sys.global.b32 [%0], %1;" ::"l"(mc_base + byte_offset + off), "r"(v) : "memory");
    }
}

int nsp_push_multicast(nsp_context *ctx, void *d_multicast_base, size_t byte_offset, const void *d_src, size_t nbytes)
{
    NSP_REQUIRE_CTX(ctx);
    if (!d_multicast_base || (nbytes > 0 && !d_src)) return ctx->fail(NSP_ERR_ARG, "nsp_push_multicast: bad argument");
    if (((byte_offset | nbytes | reinterpret_cast<uintptr_t>(d_src)) & 3) != 0 ||
        (reinterpret_cast<uintptr_t>(d_multicast_base) & 15) != 0 ||
        ((reinterpret_cast<uintptr_t>(d_src) - byte_offset) & 15) != 0)
        return ctx->fail(NSP_ERR_ARG, "nsp_push_multicast: alignment (see nsp_push_to_peers)");
    if (nbytes == 0) return 0;
    size_t want = (nbytes / 16 + 255) / 256;
    int grid = (int)(want < (size_t)ctx->sm_count * 8 ? (want ? want : 1) : (size_t)ctx->sm_count * 8);
    push_multicast_kernel<<<grid, 256, 0, ctx->stream>>>(static_cast<char *>(d_multicast_base), byte_offset,
                                                         static_cast<const char *>(d_src), nbytes);
    ctx->launches += 1;
    NSP_CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}

int nsp_push_to_peers(nsp_context *ctx, int npeers, void *const *d_peer_bases, size_t byte_offset,
                      const void *d_src, size_t nbytes)
{
    NSP_REQUIRE_CTX(ctx);
    if (npeers < 0 || npeers > kMaxPeers || (npeers > 0 && !d_peer_bases) || (nbytes > 0 && !d_src))
        return ctx->fail(NSP_ERR_ARG, "nsp_push_to_peers: bad argument");
    if (((byte_offset | nbytes | reinterpret_cast<uintptr_t>(d_src)) & 3) != 0)
        return ctx->fail(NSP_ERR_ARG, "nsp_push_to_peers: offsets and sizes must be multiples of 4 bytes");
    if (npeers == 0 || nbytes == 0) return 0;
    PeerList pl;
    for (int p = 0; p < npeers; ++p) {
        if ((reinterpret_cast<uintptr_t>(d_peer_bases[p]) & 15) != 0)
            return ctx->fail(NSP_ERR_ARG, "nsp_push_to_peers: destination bases must be 16-byte aligned");
        pl.base[p] = static_cast<char *>(d_peer_bases[p]);
    }
    if (((reinterpret_cast<uintptr_t>(d_src) - byte_offset) & 15) != 0)
        return ctx->fail(NSP_ERR_ARG, "nsp_push_to_peers: source and destinations must share the offset modulo 16");
    size_t want = (nbytes / 16 + 255) / 256;
    int grid = (int)(want < (size_t)ctx->sm_count * 8 ? (want ? want : 1) : (size_t)ctx->sm_count * 8);
    push_to_peers_kernel<<<grid, 256, 0, ctx->stream>>>(pl, npeers, byte_offset, static_cast<const char *>(d_src), nbytes);
    ctx->launches += 1;
    NSP_CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}

int nsp_csr_fold_s(nsp_context *ctx, int M, long long nnz, const long long *d_rpt64, const int *d_col, const float *d_val,
                   unsigned long long *h_hash2, double *h_sum2)
{
    NSP_REQUIRE_CTX(ctx);
    return fold_csr<float>(ctx, M, nnz, d_rpt64, d_col, d_val, h_hash2, h_sum2);
}

int nsp_csr_fold_d(nsp_context *ctx, int M, long long nnz, const long long *d_rpt64, const int *d_col, const double *d_val,
                   unsigned long long *h_hash2, double *h_sum2)
{
    NSP_REQUIRE_CTX(ctx);
    return fold_csr<double>(ctx, M, nnz, d_rpt64, d_col, d_val, h_hash2, h_sum2);
}

long long nsp_launch_count(nsp_context *ctx) { return ctx ? ctx->launches : 0; }

int nsp_profile_dump(nsp_context *ctx, char *buf, size_t buflen)
{
    NSP_REQUIRE_CTX(ctx);
    NSP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    NSP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->aux_stream));
    std::string out;
    for (size_t i = 0; i < ctx->prof.size(); ++i) {
        auto &r = ctx->prof[i];
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        // "<class>_long" runs on the side stream next to "<class>", which is launched right after it: the
        // time of the class is the span from the first start to the later end, reported for "<class>"
        if (i > 0) {
            auto &q = ctx->prof[i - 1];
            if (q.name == r.name + "_long") {
                float a = 0.f, b = 0.f;
                cudaEventElapsedTime(&a, q.e0, q.e1);
                cudaEventElapsedTime(&b, q.e0, r.e1);
                ms = a > b ? a : b;
            }
        }
        char line[256];
        if (i > 0 && ctx->prof[i - 1].name == r.name + "_long") {
            // the main launch's own duration next to the span of the class
            float own = 0.f;
            cudaEventElapsedTime(&own, r.e0, r.e1);
            snprintf(line, sizeof(line), "%s_own %.6f %lld %lld %lld %lld\n", r.name.c_str(), own, r.rows, r.ip, r.alen, r.out);
            out += line;
        }
        snprintf(line, sizeof(line), "%s %.6f %lld %lld %lld %lld\n", r.name.c_str(), ms, r.rows, r.ip, r.alen, r.out);
        out += line;
    }
    for (auto &r : ctx->prof) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    ctx->prof.clear();
    if (buf && buflen) {
        const size_t n = out.size() < buflen - 1 ? out.size() : buflen - 1;
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return 0;
}

// ---- SpGEMM, device pointers ---------------------------------------------------------------
int nsp_spgemm_flop(nsp_context *ctx, int M, const int *d_a_rpt, const int *d_a_col, const int *d_b_rpt,
                    long long *h_flop)
{
    NSP_REQUIRE_CTX(ctx);
    return nsp::spgemm_flop(ctx, M, d_a_rpt, d_a_col, d_b_rpt, h_flop);
}

int nsp_spgemm_symbolic(nsp_context *ctx, int M, int K, int N, const int *d_a_rpt, const int *d_a_col,
                        const int *d_b_rpt, const int *d_b_col, long long *d_c_rpt64, long long *h_nnz_c,
                        long long *h_intprod)
{
    NSP_REQUIRE_CTX(ctx);
    return nsp::spgemm_symbolic(ctx, M, K, N, d_a_rpt, d_a_col, d_b_rpt, d_b_col, d_c_rpt64, h_nnz_c,
                                h_intprod);
}

int nsp_spgemm_numeric_s(nsp_context *ctx, int M, int K, int N, const int *d_a_rpt, const int *d_a_col,
                         const float *d_a_val, const int *d_b_rpt, const int *d_b_col, const float *d_b_val,
                         const long long *d_c_rpt64, int *d_c_col, float *d_c_val)
{
    NSP_REQUIRE_CTX(ctx);
    return nsp::spgemm_numeric<float>(ctx, M, K, N, d_a_rpt, d_a_col, d_a_val, d_b_rpt, d_b_col, d_b_val,
                                      d_c_rpt64, d_c_col, d_c_val);
}

int nsp_spgemm_numeric_d(nsp_context *ctx, int M, int K, int N, const int *d_a_rpt, const int *d_a_col,
                         const double *d_a_val, const int *d_b_rpt, const int *d_b_col,
                         const double *d_b_val, const long long *d_c_rpt64, int *d_c_col, double *d_c_val)
{
    NSP_REQUIRE_CTX(ctx);
    return nsp::spgemm_numeric<double>(ctx, M, K, N, d_a_rpt, d_a_col, d_a_val, d_b_rpt, d_b_col, d_b_val,
                                       d_c_rpt64, d_c_col, d_c_val);
}

int nsp_spgemm_numeric_rows_s(nsp_context *ctx, int M, int K, int N, int row0, int nrows, const int *d_a_rpt,
                              const int *d_a_col, const float *d_a_val, const int *d_b_rpt, const int *d_b_col,
                              const float *d_b_val, const long long *d_c_rpt64, int *d_c_col, float *d_c_val)
{
    NSP_REQUIRE_CTX(ctx);
    return nsp::spgemm_numeric<float>(ctx, M, K, N, d_a_rpt, d_a_col, d_a_val, d_b_rpt, d_b_col, d_b_val,
                                      d_c_rpt64, d_c_col, d_c_val, row0, nrows);
}

int nsp_spgemm_numeric_rows_d(nsp_context *ctx, int M, int K, int N, int row0, int nrows, const int *d_a_rpt,
                              const int *d_a_col, const double *d_a_val, const int *d_b_rpt, const int *d_b_col,
                              const double *d_b_val, const long long *d_c_rpt64, int *d_c_col, double *d_c_val)
{
    NSP_REQUIRE_CTX(ctx);
    return nsp::spgemm_numeric<double>(ctx, M, K, N, d_a_rpt, d_a_col, d_a_val, d_b_rpt, d_b_col, d_b_val,
                                       d_c_rpt64, d_c_col, d_c_val, row0, nrows);
}

int nsp_rpt64_to_rpt32(nsp_context *ctx, int M, const long long *d_rpt64, long long nnz, int *d_rpt32)
{
    NSP_REQUIRE_CTX(ctx);
    return nsp::rpt64_to_rpt32(ctx, M, d_rpt64, nnz, d_rpt32);
}

}  // extern "C"

// ---- SpGEMM, host buffers ------------------------------------------------------------------
template <typename real>
static int upload_csr(nsp_context *ctx, int rows, const int *h_rpt, const int *h_col, const real *h_val,
                      int *&d_rpt, int *&d_col, void *&d_val, int &row_cap, size_t &nnz_cap, int &nnz_out)
{
    const int nnz = h_rpt[rows];
    nnz_out = nnz;
    if (rows > row_cap || !d_rpt) {
        cudaFree(d_rpt);
        d_rpt = nullptr;
        NSP_CUDA_TRY(ctx, cudaMalloc((void **)&d_rpt, sizeof(int) * ((size_t)rows + 1)));
        row_cap = rows;
    }
    if ((size_t)nnz > nnz_cap || !d_col || ctx->host.in_val_bytes != (int)sizeof(real)) {
        cudaFree(d_col);
        cudaFree(d_val);
        d_col = nullptr;
        d_val = nullptr;
        const size_t cap = (size_t)nnz + 1;
        NSP_CUDA_TRY(ctx, cudaMalloc((void **)&d_col, sizeof(int) * cap));
        NSP_CUDA_TRY(ctx, cudaMalloc((void **)&d_val, sizeof(real) * cap));
        nnz_cap = cap;
    }
    cudaStream_t st = ctx->stream;
    NSP_CUDA_TRY(ctx, cudaMemcpyAsync(d_rpt, h_rpt, sizeof(int) * ((size_t)rows + 1), cudaMemcpyHostToDevice, st));
    NSP_CUDA_TRY(ctx, cudaMemcpyAsync(d_col, h_col, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, st));
    NSP_CUDA_TRY(ctx, cudaMemcpyAsync(d_val, h_val, sizeof(real) * (size_t)nnz, cudaMemcpyHostToDevice, st));
    return 0;
}

// uploads A and B, symbolic phase, (re)allocation of C on the device; returns the device views of B
template <typename real>
static int spgemm_host_prepare(nsp_context *ctx, int M, int K, int N, const int *h_a_rpt, const int *h_a_col,
                               const real *h_a_val, const int *h_b_rpt, const int *h_b_col, const real *h_b_val,
                               const int *&b_rpt, const int *&b_col, const real *&b_val, long long &nnz)
{
    if (M < 0 || K < 0 || N < 0 || !h_a_rpt || !h_b_rpt) return ctx->fail(NSP_ERR_ARG, "nsp_spgemm_host: bad argument");
    nsp_host_result &h = ctx->host;
    if (h.in_val_bytes != (int)sizeof(real)) {
        // precision switch: drop the cached input buffers
        cudaFree(h.d_a_col); cudaFree(h.d_a_val); cudaFree(h.d_b_col); cudaFree(h.d_b_val);
        h.d_a_col = h.d_b_col = nullptr;
        h.d_a_val = h.d_b_val = nullptr;
        h.a_nnz_cap = h.b_nnz_cap = 0;
    }
    int a_nnz = 0, b_nnz = 0;
    const bool same = (h_a_rpt == h_b_rpt && h_a_col == h_b_col && (const void *)h_a_val == (const void *)h_b_val && M == K);
    if (upload_csr<real>(ctx, M, h_a_rpt, h_a_col, h_a_val, h.d_a_rpt, h.d_a_col, h.d_a_val, h.a_m_cap,
                         h.a_nnz_cap, a_nnz) != 0)
        return -1;
    h.in_val_bytes = (int)sizeof(real);
    b_rpt = h.d_a_rpt;
    b_col = h.d_a_col;
    b_val = (const real *)h.d_a_val;
    if (!same) {
        if (upload_csr<real>(ctx, K, h_b_rpt, h_b_col, h_b_val, h.d_b_rpt, h.d_b_col, h.d_b_val, h.b_m_cap,
                             h.b_nnz_cap, b_nnz) != 0)
            return -1;
        b_rpt = h.d_b_rpt;
        b_col = h.d_b_col;
        b_val = (const real *)h.d_b_val;
    }
    if (h.M < M || !h.d_rpt64) {
        cudaFree(h.d_rpt64);
        h.d_rpt64 = nullptr;
        NSP_CUDA_TRY(ctx, cudaMalloc((void **)&h.d_rpt64, sizeof(long long) * ((size_t)M + 1)));
    }
    h.M = M;
    long long ip = 0;
    nnz = 0;
    if (nsp::spgemm_symbolic(ctx, M, K, N, h.d_a_rpt, h.d_a_col, b_rpt, b_col, h.d_rpt64, &nnz, &ip) != 0)
        return -1;
    if (nnz > h.nnz || h.val_bytes != (int)sizeof(real) || !h.d_col) {
        cudaFree(h.d_col);
        cudaFree(h.d_val);
        h.d_col = nullptr;
        h.d_val = nullptr;
        NSP_CUDA_TRY(ctx, cudaMalloc((void **)&h.d_col, sizeof(int) * (size_t)(nnz + 1)));
        NSP_CUDA_TRY(ctx, cudaMalloc((void **)&h.d_val, sizeof(real) * (size_t)(nnz + 1)));
    }
    h.nnz = nnz;
    h.val_bytes = (int)sizeof(real);
    return 0;
}

template <typename real>
static int spgemm_host(nsp_context *ctx, int M, int K, int N, const int *h_a_rpt, const int *h_a_col,
                       const real *h_a_val, const int *h_b_rpt, const int *h_b_col, const real *h_b_val,
                       long long *h_nnz_c)
{
    const int *b_rpt, *b_col;
    const real *b_val;
    long long nnz = 0;
    if (spgemm_host_prepare<real>(ctx, M, K, N, h_a_rpt, h_a_col, h_a_val, h_b_rpt, h_b_col, h_b_val, b_rpt, b_col,
                                  b_val, nnz) != 0)
        return -1;
    nsp_host_result &h = ctx->host;
    if (nsp::spgemm_numeric<real>(ctx, M, K, N, h.d_a_rpt, h.d_a_col, (const real *)h.d_a_val, b_rpt, b_col,
                                  b_val, h.d_rpt64, h.d_col, (real *)h.d_val) != 0)
        return -1;
    if (h_nnz_c) *h_nnz_c = nnz;
    return 0;
}

// Double-buffered device -> pinned staging transfer of [src, src + len) on `st`, folded into `sum`.
struct StagedDrain {
    unsigned char *stage;
    size_t half;
    cudaStream_t st;
    cudaEvent_t ev[2];
    bool pending[2] = {false, false};
    size_t pend_len[2] = {0, 0};
    int slot = 0;
    unsigned long long sum = 0;
    long long total = 0;
    void consume(int s)
    {
        // fold the first and last word of the chunk: proves the bytes arrived without a host pass
        // over tens of GB
        cudaEventSynchronize(ev[s]);
        const unsigned char *p = stage + (size_t)s * half;
        unsigned long long a = 0, b = 0;
        memcpy(&a, p, pend_len[s] >= 8 ? 8 : pend_len[s]);
        if (pend_len[s] >= 8) memcpy(&b, p + pend_len[s] - 8, 8);
        sum = sum * 1099511628211ull + (a ^ (b << 1));
        pending[s] = false;
    }
    cudaError_t range(const void *src, size_t len)
    {
        for (size_t off = 0; off < len; off += half) {
            const size_t n = len - off < half ? len - off : half;
            if (pending[slot]) consume(slot);
            cudaError_t e = cudaMemcpyAsync(stage + (size_t)slot * half, (const unsigned char *)src + off, n,
                                            cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) return e;
            e = cudaEventRecord(ev[slot], st);
            if (e != cudaSuccess) return e;
            pending[slot] = true;
            pend_len[slot] = n;
            total += (long long)n;
            slot ^= 1;
        }
        return cudaSuccess;
    }
    void finish()
    {
        if (pending[slot]) consume(slot);
        if (pending[slot ^ 1]) consume(slot ^ 1);
    }
};

__global__ void find_row_cuts_kernel(const long long *__restrict__ rpt, int M, long long nnz, int pieces,
                                     int *__restrict__ rows, long long *__restrict__ offs)
{
    const int k = threadIdx.x;
    if (k > pieces) return;
    int r = M;
    if (k < pieces) {
        const long long target = nnz / pieces * k;
        int lo = 0, hi = M;                       // first row whose start is >= target
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (rpt[mid] < target) lo = mid + 1; else hi = mid;
        }
        r = lo;
    }
    rows[k] = r;
    offs[k] = rpt[r];
}

// spgemm_kernel_hash with HOST buffers END TO END: the numeric phase runs in `pieces` contiguous row ranges
// of ~equal output and a second host thread streams every finished range to the host through the pinned
// staging buffer on its own stream, so the PCIe transfer of C (78 GB at R-MAT scale 20, 87 % of the
// un-overlapped call) hides the compute of the following ranges.
template <typename real>
static int spgemm_host_stream(nsp_context *ctx, int M, int K, int N, const int *h_a_rpt, const int *h_a_col,
                              const real *h_a_val, const int *h_b_rpt, const int *h_b_col, const real *h_b_val,
                              void *h_stage, size_t stage_bytes, int pieces, long long *h_nnz_c,
                              unsigned long long *h_checksum, long long *h_bytes)
{
    if (!h_stage || stage_bytes < 4096) return ctx->fail(NSP_ERR_ARG, "nsp_spgemm_host_stream: bad staging buffer");
    if (pieces < 1) pieces = 1;
    if (pieces > 64) pieces = 64;
    const int *b_rpt, *b_col;
    const real *b_val;
    long long nnz = 0;
    if (spgemm_host_prepare<real>(ctx, M, K, N, h_a_rpt, h_a_col, h_a_val, h_b_rpt, h_b_col, h_b_val, b_rpt, b_col,
                                  b_val, nnz) != 0)
        return -1;
    nsp_host_result &h = ctx->host;
    if (M < pieces) pieces = M > 0 ? M : 1;
    // row cuts of equal output (the 65-entry scratch lives in the context's host-call state)
    int rows[65];
    long long offs[65];
    {
        if (!h.d_cut_rows) {
            NSP_CUDA_TRY(ctx, cudaMalloc((void **)&h.d_cut_rows, sizeof(int) * 65));
            NSP_CUDA_TRY(ctx, cudaMalloc((void **)&h.d_cut_offs, sizeof(long long) * 65));
        }
        find_row_cuts_kernel<<<1, 128, 0, ctx->stream>>>(h.d_rpt64, M, nnz, pieces, h.d_cut_rows, h.d_cut_offs);
        ctx->launches += 1;
        NSP_CUDA_TRY(ctx, cudaMemcpyAsync(rows, h.d_cut_rows, sizeof(int) * (pieces + 1), cudaMemcpyDeviceToHost, ctx->stream));
        NSP_CUDA_TRY(ctx, cudaMemcpyAsync(offs, h.d_cut_offs, sizeof(long long) * (pieces + 1), cudaMemcpyDeviceToHost, ctx->stream));
        NSP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    // stream + events of the drain, released on every exit path
    struct DrainRes {
        cudaStream_t st = nullptr;
        std::vector<cudaEvent_t> done;
        cudaEvent_t ev[2] = {nullptr, nullptr};
        ~DrainRes()
        {
            if (st) cudaStreamSynchronize(st);
            for (auto e : done)
                if (e) cudaEventDestroy(e);
            for (auto e : ev)
                if (e) cudaEventDestroy(e);
            if (st) cudaStreamDestroy(st);
        }
    } res;
    NSP_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&res.st, cudaStreamNonBlocking));
    res.done.assign((size_t)pieces, nullptr);
    for (auto &e : res.done) NSP_CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    NSP_CUDA_TRY(ctx, cudaEventCreateWithFlags(&res.ev[0], cudaEventDisableTiming));
    NSP_CUDA_TRY(ctx, cudaEventCreateWithFlags(&res.ev[1], cudaEventDisableTiming));
    cudaStream_t copy_st = res.st;
    std::vector<cudaEvent_t> &done = res.done;
    StagedDrain dr;
    dr.stage = (unsigned char *)h_stage;
    dr.half = (stage_bytes / 2) & ~size_t(255);
    dr.st = copy_st;
    dr.ev[0] = res.ev[0];
    dr.ev[1] = res.ev[1];
    std::atomic<int> launched(0);
    std::atomic<int> failed(0);
    cudaError_t drain_err = cudaSuccess;
    const int device = ctx->device;
    std::thread drainer([&]() {
        cudaSetDevice(device);
        // the row pointer is final after the symbolic phase
        drain_err = dr.range(h.d_rpt64, sizeof(long long) * ((size_t)M + 1));
        for (int k = 0; k < pieces && drain_err == cudaSuccess; ++k) {
            while (launched.load(std::memory_order_acquire) <= k) {
                if (failed.load()) return;
                std::this_thread::yield();
            }
            drain_err = cudaStreamWaitEvent(copy_st, done[k], 0);
            const size_t e0 = (size_t)offs[k], e1 = (size_t)offs[k + 1];
            if (drain_err == cudaSuccess) drain_err = dr.range(h.d_col + e0, sizeof(int) * (e1 - e0));
            if (drain_err == cudaSuccess) drain_err = dr.range((const real *)h.d_val + e0, sizeof(real) * (e1 - e0));
        }
        dr.finish();
    });
    int rc = 0;
    for (int k = 0; k < pieces; ++k) {
        if (rows[k + 1] > rows[k])
            rc = nsp::spgemm_numeric<real>(ctx, M, K, N, h.d_a_rpt, h.d_a_col, (const real *)h.d_a_val, b_rpt, b_col, b_val,
                                           h.d_rpt64, h.d_col, (real *)h.d_val, rows[k], rows[k + 1] - rows[k]);
        if (rc != 0 || cudaEventRecord(done[k], ctx->stream) != cudaSuccess) {
            failed.store(1);
            rc = rc ? rc : -1;
            break;
        }
        launched.store(k + 1, std::memory_order_release);
    }
    drainer.join();
    cudaStreamSynchronize(copy_st);
    if (rc != 0) return rc;
    if (drain_err != cudaSuccess) return ctx->fail(NSP_ERR_CUDA, std::string("nsp_spgemm_host_stream: ") + cudaGetErrorString(drain_err));
    if (h_nnz_c) *h_nnz_c = nnz;
    if (h_checksum) *h_checksum = dr.sum;
    if (h_bytes) *h_bytes = dr.total;
    return 0;
}

extern "C" {

int nsp_spgemm_host_s(nsp_context *ctx, int M, int K, int N, const int *h_a_rpt, const int *h_a_col,
                      const float *h_a_val, const int *h_b_rpt, const int *h_b_col, const float *h_b_val,
                      long long *h_nnz_c)
{
    NSP_REQUIRE_CTX(ctx);
    return spgemm_host<float>(ctx, M, K, N, h_a_rpt, h_a_col, h_a_val, h_b_rpt, h_b_col, h_b_val, h_nnz_c);
}

int nsp_spgemm_host_d(nsp_context *ctx, int M, int K, int N, const int *h_a_rpt, const int *h_a_col,
                      const double *h_a_val, const int *h_b_rpt, const int *h_b_col, const double *h_b_val,
                      long long *h_nnz_c)
{
    NSP_REQUIRE_CTX(ctx);
    return spgemm_host<double>(ctx, M, K, N, h_a_rpt, h_a_col, h_a_val, h_b_rpt, h_b_col, h_b_val, h_nnz_c);
}

int nsp_spgemm_host_stream_s(nsp_context *ctx, int M, int K, int N, const int *h_a_rpt, const int *h_a_col,
                             const float *h_a_val, const int *h_b_rpt, const int *h_b_col, const float *h_b_val,
                             void *h_stage, size_t stage_bytes, int pieces, long long *h_nnz_c,
                             unsigned long long *h_checksum, long long *h_bytes)
{
    NSP_REQUIRE_CTX(ctx);
    return spgemm_host_stream<float>(ctx, M, K, N, h_a_rpt, h_a_col, h_a_val, h_b_rpt, h_b_col, h_b_val, h_stage,
                                     stage_bytes, pieces, h_nnz_c, h_checksum, h_bytes);
}

int nsp_spgemm_host_stream_d(nsp_context *ctx, int M, int K, int N, const int *h_a_rpt, const int *h_a_col,
                             const double *h_a_val, const int *h_b_rpt, const int *h_b_col, const double *h_b_val,
                             void *h_stage, size_t stage_bytes, int pieces, long long *h_nnz_c,
                             unsigned long long *h_checksum, long long *h_bytes)
{
    NSP_REQUIRE_CTX(ctx);
    return spgemm_host_stream<double>(ctx, M, K, N, h_a_rpt, h_a_col, h_a_val, h_b_rpt, h_b_col, h_b_val, h_stage,
                                      stage_bytes, pieces, h_nnz_c, h_checksum, h_bytes);
}

static int host_fetch(nsp_context *ctx, int val_bytes, long long *h_c_rpt64, int *h_c_col, void *h_c_val)
{
    nsp_host_result &h = ctx->host;
    if (!h.d_rpt64 || h.val_bytes != val_bytes) return ctx->fail(NSP_ERR_ARG, "nsp_spgemm_host_fetch: no result of this precision on the context");
    cudaStream_t st = ctx->stream;
    if (h_c_rpt64)
        NSP_CUDA_TRY(ctx, cudaMemcpyAsync(h_c_rpt64, h.d_rpt64, sizeof(long long) * ((size_t)h.M + 1), cudaMemcpyDeviceToHost, st));
    if (h_c_col && h.nnz)
        NSP_CUDA_TRY(ctx, cudaMemcpyAsync(h_c_col, h.d_col, sizeof(int) * (size_t)h.nnz, cudaMemcpyDeviceToHost, st));
    if (h_c_val && h.nnz)
        NSP_CUDA_TRY(ctx, cudaMemcpyAsync(h_c_val, h.d_val, (size_t)val_bytes * (size_t)h.nnz, cudaMemcpyDeviceToHost, st));
    NSP_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return 0;
}

int nsp_spgemm_host_fetch_s(nsp_context *ctx, long long *h_c_rpt64, int *h_c_col, float *h_c_val)
{
    NSP_REQUIRE_CTX(ctx);
    return host_fetch(ctx, 4, h_c_rpt64, h_c_col, h_c_val);
}

int nsp_spgemm_host_fetch_d(nsp_context *ctx, long long *h_c_rpt64, int *h_c_col, double *h_c_val)
{
    NSP_REQUIRE_CTX(ctx);
    return host_fetch(ctx, 8, h_c_rpt64, h_c_col, h_c_val);
}

int nsp_spgemm_host_drain(nsp_context *ctx, void *h_stage, size_t stage_bytes, unsigned long long *h_checksum,
                          long long *h_bytes)
{
    NSP_REQUIRE_CTX(ctx);
    nsp_host_result &h = ctx->host;
    if (!h.d_rpt64 || !h_stage || stage_bytes < 4096) return ctx->fail(NSP_ERR_ARG, "nsp_spgemm_host_drain: bad argument");
    StagedDrain dr;
    dr.stage = (unsigned char *)h_stage;
    dr.half = (stage_bytes / 2) & ~size_t(255);
    dr.st = ctx->stream;
    NSP_CUDA_TRY(ctx, cudaEventCreateWithFlags(&dr.ev[0], cudaEventDisableTiming));
    NSP_CUDA_TRY(ctx, cudaEventCreateWithFlags(&dr.ev[1], cudaEventDisableTiming));
    const void *src[3] = {h.d_rpt64, h.d_col, h.d_val};
    const size_t len[3] = {sizeof(long long) * ((size_t)h.M + 1), sizeof(int) * (size_t)h.nnz,
                           (size_t)h.val_bytes * (size_t)h.nnz};
    cudaError_t e = cudaSuccess;
    for (int a = 0; a < 3 && e == cudaSuccess; ++a) e = dr.range(src[a], len[a]);
    dr.finish();
    cudaEventDestroy(dr.ev[0]);
    cudaEventDestroy(dr.ev[1]);
    NSP_CUDA_TRY(ctx, e);
    const unsigned long long sum = dr.sum;
    const long long total = dr.total;
    if (h_checksum) *h_checksum = sum;
    if (h_bytes) *h_bytes = total;
    return 0;
}

int nsp_spgemm_host_release(nsp_context *ctx)
{
    NSP_REQUIRE_CTX(ctx);
    nsp_host_result &h = ctx->host;
    cudaStreamSynchronize(ctx->stream);
    cudaFree(h.d_rpt64); cudaFree(h.d_col); cudaFree(h.d_val);
    cudaFree(h.d_a_rpt); cudaFree(h.d_a_col); cudaFree(h.d_a_val);
    cudaFree(h.d_b_rpt); cudaFree(h.d_b_col); cudaFree(h.d_b_val);
    cudaFree(h.d_cut_rows); cudaFree(h.d_cut_offs);
    h = nsp_host_result();
    return 0;
}

}  // extern "C"
