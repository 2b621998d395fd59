// nsparse-b200: device-side planning for the hash SpGEMM.
//
// Replaces set_max_bin / set_min_bin of the reference (kernel_spgemm_hash_d.cu:156-246):
//   * set_intprod_num (:70-86)   -> count_ip_kernel   (warp-cooperative on long rows, 64-bit total,
//                                                     block-aggregated histogram instead of one
//                                                     same-address atomic per row)
//   * set_bin (:88-112)          -> folded into count_ip_kernel / hist_kernel
//   * host prefix + 3 memcpys    -> bin_offsets_kernel (stays on the device, no round trip)
//   * set_row_perm (:125-154)    -> scatter_rows_kernel (one atomic per bin per block)
//   * thrust::exclusive_scan (:1183) -> scan_* kernels, int32 counts -> int64 row pointer
//
// Bins are logarithmic (bin 0: v <= 2^s, bin b: 2^(s+b-1) < v <= 2^(s+b)) and laid out in
// row_perm from the heaviest bin to the lightest, so a kernel class that covers bins [lo, hi]
// owns one contiguous slice and meets its most expensive rows first (LPT order for the
// dynamic row queue).
#include "context.h"
#include "spgemm_plan.h"

namespace nsp {

__device__ __forceinline__ long long warp_sum_ll(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One thread per row for short rows, the whole warp for rows longer than 64 entries.
__global__ void __launch_bounds__(256)
count_ip_kernel(const int *__restrict__ a_rpt, const int *__restrict__ a_col,
                const int *__restrict__ b_rpt, int M, int cap, int shift, int *__restrict__ row_ip,
                int *__restrict__ hist, unsigned long long *__restrict__ binsum,
                unsigned long long *__restrict__ total_ip, unsigned long long *__restrict__ max_len = nullptr)
{
    __shared__ int s_hist[kNumBins];
    __shared__ unsigned long long s_ipsum[kNumBins], s_lensum[kNumBins];
    __shared__ unsigned long long s_total;
    if (threadIdx.x < kNumBins) {
        s_hist[threadIdx.x] = 0;
        s_ipsum[threadIdx.x] = 0ull;
        s_lensum[threadIdx.x] = 0ull;
    }
    if (threadIdx.x == 0) s_total = 0ull;
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int beg = 0, end = 0;
    if (row < M) {
        beg = a_rpt[row];
        end = a_rpt[row + 1];
    }
    long long ip = 0;
    const int len = end - beg;
    if (len <= 64) {
        for (int j = beg; j < end; ++j) {
            const int c = a_col[j];
            ip += ld_nc(b_rpt + c + 1) - ld_nc(b_rpt + c);
        }
    }
    unsigned longmask = __ballot_sync(0xffffffffu, len > 64);
    while (longmask) {
        const int src = __ffs(longmask) - 1;
        longmask &= longmask - 1;
        const int b = __shfl_sync(0xffffffffu, beg, src);
        const int e = __shfl_sync(0xffffffffu, end, src);
        long long s = 0;
        for (int j = b + lane; j < e; j += 32) {
            const int c = a_col[j];
            s += ld_nc(b_rpt + c + 1) - ld_nc(b_rpt + c);
        }
        s = warp_sum_ll(s);
        if (lane == src) ip = s;
    }
    if (row < M) {
        // row_ip keeps the true product count (saturated); bins and table sizes use min(ip, cap)
        row_ip[row] = (int)(ip < 0x7fffffffll ? ip : 0x7fffffffll);
        const int v = (int)(ip < (long long)cap ? ip : (long long)cap);
        const int b = log_bin(v, shift);
        atomicAdd(&s_hist[b], 1);
        if (binsum && len > 0) {
            atomicAdd(&s_ipsum[b], (unsigned long long)ip);
            atomicAdd(&s_lensum[b], (unsigned long long)len);
        }
    }
    const long long wsum = warp_sum_ll(ip);
    if (lane == 0 && wsum) atomicAdd(&s_total, (unsigned long long)wsum);
    if (max_len) {
        int ml = len;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ml = max(ml, __shfl_xor_sync(0xffffffffu, ml, o));
        if (lane == 0 && ml > 1024) atomicMax(max_len, (unsigned long long)ml);
    }
    __syncthreads();
    if (threadIdx.x < kNumBins && s_hist[threadIdx.x]) {
        atomicAdd(&hist[threadIdx.x], s_hist[threadIdx.x]);
        if (binsum) {
            atomicAdd(&binsum[kSumIp + threadIdx.x], s_ipsum[threadIdx.x]);
            atomicAdd(&binsum[kSumLen + threadIdx.x], s_lensum[threadIdx.x]);
        }
    }
    if (threadIdx.x == 0 && s_total) atomicAdd(total_ip, s_total);
}

__global__ void __launch_bounds__(256)
hist_kernel(const int *__restrict__ values, const int *__restrict__ row_ip,
            const int *__restrict__ a_rpt, int M, int shift, int *__restrict__ hist,
            unsigned long long *__restrict__ binsum)
{
    __shared__ int s_hist[kNumBins];
    __shared__ unsigned long long s_ipsum[kNumBins], s_lensum[kNumBins], s_cntsum[kNumBins];
    if (threadIdx.x < kNumBins) {
        s_hist[threadIdx.x] = 0;
        s_ipsum[threadIdx.x] = 0ull;
        s_lensum[threadIdx.x] = 0ull;
        s_cntsum[threadIdx.x] = 0ull;
    }
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M;
         i += (long long)gridDim.x * blockDim.x) {
        const int b = log_bin(values[i], shift);
        atomicAdd(&s_hist[b], 1);
        const int len = a_rpt[i + 1] - a_rpt[i];
        if (len > 0) {
            atomicAdd(&s_ipsum[b], (unsigned long long)row_ip[i]);
            atomicAdd(&s_lensum[b], (unsigned long long)len);
            atomicAdd(&s_cntsum[b], (unsigned long long)values[i]);
        }
    }
    __syncthreads();
    if (threadIdx.x < kNumBins && s_hist[threadIdx.x]) {
        atomicAdd(&hist[threadIdx.x], s_hist[threadIdx.x]);
        atomicAdd(&binsum[kSumIp + threadIdx.x], s_ipsum[threadIdx.x]);
        atomicAdd(&binsum[kSumLen + threadIdx.x], s_lensum[threadIdx.x]);
        atomicAdd(&binsum[kSumCnt + threadIdx.x], s_cntsum[threadIdx.x]);
    }
}

// Are the rows of B column-sorted (non-decreasing)?  unsorted = descents over the whole column array
// minus the descents that sit on a row boundary.  The heavy kernels cut B rows by column range with
// binary searches when they are, and fall back to filtering when not.
__global__ void __launch_bounds__(256)
check_b_sorted_kernel(const int *__restrict__ b_rpt, const int *__restrict__ b_col, int K,
                      unsigned long long *__restrict__ unsorted)
{
    const long long nnz = b_rpt[K];
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    long long d = 0;
    for (long long k = tid + 1; k < nnz; k += nth) d += b_col[k] < b_col[k - 1];
    for (long long i = tid + 1; i < K; i += nth) {
        const int s = b_rpt[i];
        if (s > 0 && s < b_rpt[i + 1]) d -= b_col[s] < b_col[s - 1];
    }
    d = warp_sum_ll(d);
    if ((threadIdx.x & 31) == 0 && d) atomicAdd(unsorted, (unsigned long long)d);
    if (tid == 0) unsorted[kScalarNnzB - kScalarUnsorted] = (unsigned long long)nnz;
}

// Window cuts of every B row an entry of A refers to, for the segment mode of the heavy numeric kernel
// (spgemm_device.cuh stage_window_seg): seg[w * stride + j] = first product of entry j's B row at or beyond
// column w << wshift, w = 0 .. nwin (w = 0: the row's start, w = nwin: its end).  One thread per entry, each cut
// found by a binary search that starts at the previous one.  Needs column-sorted rows of B.
__global__ void __launch_bounds__(256)
entry_segments_kernel(const int *__restrict__ a_col, long long count, const int *__restrict__ b_rpt,
                      const int *__restrict__ b_col, int nwin, int wshift, long long stride, int *__restrict__ seg)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    const int ac = ld_stream(a_col + j);
    int lo = ld_nc(b_rpt + ac);
    const int ke = ld_nc(b_rpt + ac + 1);
    seg[j] = lo;
    for (int w = 1; w < nwin; ++w) {
        const int key = (int)((unsigned)w << wshift);
        int hi = ke;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (ld_nc(b_col + mid) < key) lo = mid + 1; else hi = mid;
        }
        seg[(long long)w * stride + j] = lo;
    }
    seg[(long long)nwin * stride + j] = ke;
}

// (nwin + 1) * count ints in the context's grow-only segment buffer
int reserve_entry_segments(nsp_context *ctx, long long count, int nwin)
{
    const size_t want = (size_t)(nwin + 1) * (size_t)(count > 0 ? count : 1);
    if (want > ctx->seg_cap) {
        NSP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->aux_stream) NSP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->aux_stream));
        cudaFree(ctx->d_seg);
        ctx->d_seg = nullptr;
        ctx->seg_cap = 0;
        const size_t cap = want + want / 8;
        NSP_CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_seg, sizeof(int) * cap));
        ctx->seg_cap = cap;
    }
    return 0;
}

int build_entry_segments(nsp_context *ctx, const int *a_col, long long count, const int *b_rpt, const int *b_col,
                         int nwin, int wshift)
{
    if (reserve_entry_segments(ctx, count, nwin) != 0) return -1;
    if (count > 0) {
        entry_segments_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(a_col, count, b_rpt, b_col, nwin, wshift,
                                                                                      count, ctx->d_seg);
        ctx->launches += 1;
        NSP_CUDA_TRY(ctx, cudaGetLastError());
    }
    return 0;
}

// start[b] = number of rows in bins heavier than b; also clears cursors and queue heads.
__global__ void bin_offsets_kernel(int *bins)
{
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int b = kNumBins - 1; b >= 0; --b) {
            bins[kBinStart + b] = acc;
            acc += bins[kBinHist + b];
        }
    }
    if (threadIdx.x < kNumBins) bins[kBinCursor + threadIdx.x] = 0;
    if (threadIdx.x < kNumQueues) bins[kBinQueue + threadIdx.x] = 0;
}

__global__ void __launch_bounds__(256)
scatter_rows_kernel(const int *__restrict__ values, int M, int cap, int shift, int *__restrict__ bins,
                    int *__restrict__ row_perm)
{
    __shared__ int s_cnt[kNumBins];
    __shared__ int s_base[kNumBins];
    if (threadIdx.x < kNumBins) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int b = 0, local = 0;
    if (row < M) {
        b = log_bin(min(values[row], cap), shift);
        local = atomicAdd(&s_cnt[b], 1);
    }
    __syncthreads();
    if (threadIdx.x < kNumBins && s_cnt[threadIdx.x])
        s_base[threadIdx.x] = bins[kBinStart + threadIdx.x] +
                              atomicAdd(&bins[kBinCursor + threadIdx.x], s_cnt[threadIdx.x]);
    __syncthreads();
    if (row < M) row_perm[s_base[b] + local] = (int)row;
}

// ---- exclusive scan int32 -> int64, reduce-then-scan, 4096 items per block --------------------
constexpr int kScanBlock = 512;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanBlock * kScanItems;

__device__ __forceinline__ long long block_exclusive_scan_ll(long long v, long long *total)
{
    // kScanBlock threads; returns exclusive prefix of v, *total = block sum
    __shared__ long long s_warp[kScanBlock / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        long long w = lane < kScanBlock / 32 ? s_warp[lane] : 0;
        long long winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < kScanBlock / 32) s_warp[lane] = winc - w;   // exclusive warp offsets
        if (lane == kScanBlock / 32 - 1) *total = winc;
    }
    __syncthreads();
    const long long r = s_warp[wid] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kScanBlock)
scan_reduce_kernel(const int *__restrict__ in, long long n, long long *__restrict__ block_sums)
{
    __shared__ long long s_total;
    const long long base = (long long)blockIdx.x * kScanTile;
    long long s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const long long i = base + (long long)k * kScanBlock + threadIdx.x;
        if (i < n) s += in[i];
    }
    (void)block_exclusive_scan_ll(s, &s_total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = s_total;
}

// single block; scans up to any number of block sums serially in chunks of kScanBlock
__global__ void __launch_bounds__(kScanBlock)
scan_spine_kernel(long long *block_sums, int nblocks, long long *total_out)
{
    __shared__ long long s_total;
    long long carry = 0;
    for (int base = 0; base < nblocks; base += kScanBlock) {
        const int i = base + threadIdx.x;
        const long long v = i < nblocks ? block_sums[i] : 0;
        const long long ex = block_exclusive_scan_ll(v, &s_total);
        if (i < nblocks) block_sums[i] = carry + ex;
        carry += s_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(kScanBlock)
scan_apply_kernel(const int *__restrict__ in, long long n, const long long *__restrict__ block_sums,
                  long long *__restrict__ out /* n + 1 */, const long long *__restrict__ total)
{
    __shared__ long long s_total;
    const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
    int v[kScanItems];
    long long s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    long long ex = block_exclusive_scan_ll(s, &s_total) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = *total;
}

__global__ void narrow_rpt_kernel(const long long *__restrict__ in, long long n, int *__restrict__ out)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        out[i] = (int)in[i];
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int plan_reserve(nsp_context *ctx, int M)
{
    // [row_cnt M+1][row_ip M+1][row_perm M][bins][bin sums][scalars][scan tmp]
    const size_t nblk = (size_t)(M / kScanTile + 2);
    size_t bytes = 0;
    auto pad = [](size_t b) { return (b + 255) & ~size_t(255); };
    bytes += 3 * pad(sizeof(int) * ((size_t)M + 1));
    bytes += pad(sizeof(int) * kBinInts);
    bytes += pad(sizeof(unsigned long long) * kSumInts);
    bytes += pad(sizeof(long long) * 8);
    bytes += pad(sizeof(long long) * nblk);
    if (ctx->arena_reserve(bytes) != 0) return -1;
    ctx->arena_reset();
    nsp_spgemm_state &sp = ctx->sp;
    sp.d_row_cnt = ctx->arena_take<int>((size_t)M + 1);
    sp.d_row_ip = ctx->arena_take<int>((size_t)M + 1);
    sp.d_row_perm = ctx->arena_take<int>((size_t)M + 1);
    sp.d_bins = ctx->arena_take<int>(kBinInts);
    sp.d_binsum = ctx->arena_take<unsigned long long>(kSumInts);
    sp.d_scalars = ctx->arena_take<long long>(8);
    sp.d_scan_tmp = ctx->arena_take<long long>(nblk);
    if (!sp.h_scalars) {
        NSP_CUDA_TRY(ctx, cudaMallocHost((void **)&sp.h_scalars, sizeof(long long) * 8));
        NSP_CUDA_TRY(ctx, cudaMallocHost((void **)&sp.h_bins, sizeof(int) * kBinInts));
        NSP_CUDA_TRY(ctx, cudaMallocHost((void **)&sp.h_binsum, sizeof(unsigned long long) * kSumInts));
    }
    return 0;
}

// Bring the 28-bin histogram (+ per-bin product / A-entry sums) to the host: ~700 bytes, one
// sync.  The host uses it to skip empty kernel classes and to pick the lanes-per-B-row of each
// class; the reference does 3 blocking memcpys here plus a host prefix (:173-185).
static int plan_fetch(nsp_context *ctx)
{
    nsp_spgemm_state &sp = ctx->sp;
    cudaStream_t st = ctx->stream;
    NSP_CUDA_TRY(ctx, cudaMemcpyAsync(sp.h_bins, sp.d_bins, sizeof(int) * kBinInts, cudaMemcpyDeviceToHost, st));
    NSP_CUDA_TRY(ctx, cudaMemcpyAsync(sp.h_binsum, sp.d_binsum, sizeof(unsigned long long) * kSumInts,
                                      cudaMemcpyDeviceToHost, st));
    NSP_CUDA_TRY(ctx, cudaMemcpyAsync(sp.h_scalars, sp.d_scalars, sizeof(long long) * 8, cudaMemcpyDeviceToHost, st));
    NSP_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return 0;
}

// ip per row (capped at `cap`), histogram, offsets, permutation.
int plan_by_intprod(nsp_context *ctx, int M, int K, int cap, const int *a_rpt, const int *a_col,
                    const int *b_rpt, const int *b_col)
{
    nsp_spgemm_state &sp = ctx->sp;
    cudaStream_t st = ctx->stream;
    NSP_CUDA_TRY(ctx, cudaMemsetAsync(sp.d_bins, 0, sizeof(int) * kBinInts, st));
    NSP_CUDA_TRY(ctx, cudaMemsetAsync(sp.d_binsum, 0, sizeof(unsigned long long) * kSumInts, st));
    NSP_CUDA_TRY(ctx, cudaMemsetAsync(sp.d_scalars, 0, sizeof(long long) * 8, st));
    if (M > 0) {
        const int grid = (M + 255) / 256;
        count_ip_kernel<<<grid, 256, 0, st>>>(a_rpt, a_col, b_rpt, M, cap, kSymShift, sp.d_row_ip,
                                              sp.d_bins + kBinHist, sp.d_binsum,
                                              (unsigned long long *)(sp.d_scalars + kScalarIp),
                                              (unsigned long long *)(sp.d_scalars + kScalarMaxLen));
        bin_offsets_kernel<<<1, 32, 0, st>>>(sp.d_bins);
        scatter_rows_kernel<<<grid, 256, 0, st>>>(sp.d_row_ip, M, cap, kSymShift, sp.d_bins, sp.d_row_perm);
        ctx->launches += 3;
        if (K > 0) {
            check_b_sorted_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(
                b_rpt, b_col, K, (unsigned long long *)(sp.d_scalars + kScalarUnsorted));
            ctx->launches += 1;
        }
    }
    NSP_CUDA_TRY(ctx, cudaGetLastError());
    if (plan_fetch(ctx) != 0) return -1;
    sp.b_sorted = sp.h_scalars[kScalarUnsorted] == 0;
    sp.b_nnz = K > 0 && M > 0 ? sp.h_scalars[kScalarNnzB] : 0;
    sp.has_multi_slab = sp.h_scalars[kScalarMaxLen] > 1024;
    sp.a_nnz = 0;
    for (int b = 0; b < kNumBins; ++b) sp.a_nnz += (long long)sp.h_binsum[kSumLen + b];
    return 0;
}

// a_rpt is already entered at row0; the per-row arrays of the plan are entered here
int plan_by_count(nsp_context *ctx, int M, int shift, const int *a_rpt, int row0)
{
    nsp_spgemm_state &sp = ctx->sp;
    cudaStream_t st = ctx->stream;
    const int *row_cnt = sp.d_row_cnt + row0;
    const int *row_ip = sp.d_row_ip + row0;
    if (sp.join_pending) {
        // a previous numeric phase left through an error path before joining its side-stream launch, which
        // still reads d_bins / d_row_perm: order it before they are rewritten
        NSP_CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->ev_join, 0));
        sp.join_pending = false;
    }
    NSP_CUDA_TRY(ctx, cudaMemsetAsync(sp.d_bins, 0, sizeof(int) * kBinInts, st));
    NSP_CUDA_TRY(ctx, cudaMemsetAsync(sp.d_binsum, 0, sizeof(unsigned long long) * kSumInts, st));
    if (M > 0) {
        const int grid = (M + 255) / 256;
        int hgrid = grid < ctx->sm_count * 8 ? grid : ctx->sm_count * 8;
        hist_kernel<<<hgrid, 256, 0, st>>>(row_cnt, row_ip, a_rpt, M, shift, sp.d_bins + kBinHist, sp.d_binsum);
        bin_offsets_kernel<<<1, 32, 0, st>>>(sp.d_bins);
        scatter_rows_kernel<<<grid, 256, 0, st>>>(row_cnt, M, 0x7fffffff, shift, sp.d_bins, sp.d_row_perm);
        ctx->launches += 3;
    }
    NSP_CUDA_TRY(ctx, cudaGetLastError());
    return plan_fetch(ctx);
}

// row_cnt[0..M) -> rpt64[0..M], total -> d_scalars[kScalarNnz]
int scan_row_counts(nsp_context *ctx, int M, long long *rpt64)
{
    nsp_spgemm_state &sp = ctx->sp;
    cudaStream_t st = ctx->stream;
    const int nblk = M / kScanTile + 1;
    scan_reduce_kernel<<<nblk, kScanBlock, 0, st>>>(sp.d_row_cnt, M, sp.d_scan_tmp);
    scan_spine_kernel<<<1, kScanBlock, 0, st>>>(sp.d_scan_tmp, nblk, sp.d_scalars + kScalarNnz);
    scan_apply_kernel<<<nblk, kScanBlock, 0, st>>>(sp.d_row_cnt, M, sp.d_scan_tmp, rpt64,
                                                  sp.d_scalars + kScalarNnz);
    ctx->launches += 3;
    NSP_CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}

int spgemm_flop(nsp_context *ctx, int M, const int *a_rpt, const int *a_col, const int *b_rpt,
                long long *h_flop)
{
    if (M < 0 || !h_flop) return ctx->fail(-2, "nsp_spgemm_flop: bad argument");
    // flop counting must not disturb a pending symbolic->numeric pair, so it uses its own scratch
    int *d_cnt = nullptr, *d_hist = nullptr;
    unsigned long long *d_total = nullptr;
    cudaStream_t st = ctx->stream;
    NSP_CUDA_TRY(ctx, cudaMalloc((void **)&d_cnt, sizeof(int) * ((size_t)M + 1)));
    NSP_CUDA_TRY(ctx, cudaMalloc((void **)&d_hist, sizeof(int) * kNumBins));
    NSP_CUDA_TRY(ctx, cudaMalloc((void **)&d_total, sizeof(unsigned long long)));
    NSP_CUDA_TRY(ctx, cudaMemsetAsync(d_hist, 0, sizeof(int) * kNumBins, st));
    NSP_CUDA_TRY(ctx, cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), st));
    if (M > 0) {
        count_ip_kernel<<<(M + 255) / 256, 256, 0, st>>>(a_rpt, a_col, b_rpt, M, 0x7fffffff, kSymShift,
                                                       d_cnt, d_hist, nullptr, d_total);
        ctx->launches += 1;
    }
    unsigned long long total = 0;
    NSP_CUDA_TRY(ctx, cudaMemcpyAsync(&total, d_total, sizeof(total), cudaMemcpyDeviceToHost, st));
    NSP_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    cudaFree(d_cnt);
    cudaFree(d_hist);
    cudaFree(d_total);
    *h_flop = 2ll * (long long)total;
    return 0;
}

int rpt64_to_rpt32(nsp_context *ctx, int M, const long long *rpt64, long long nnz, int *rpt32)
{
    if (nnz > 0x7fffffffll)
        return ctx->fail(-3, "nnz(C) = " + std::to_string(nnz) +
                                 " does not fit the int32 row pointer of sfCSR");
    const long long n = (long long)M + 1;
    int grid = (int)((n + 255) / 256);
    if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
    narrow_rpt_kernel<<<grid, 256, 0, ctx->stream>>>(rpt64, n, rpt32);
    ctx->launches += 1;
    NSP_CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}

}  // namespace nsp
