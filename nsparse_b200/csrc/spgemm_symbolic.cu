// nsparse-b200: SYMBOLIC phase of the hash SpGEMM -- exact nnz of every row of C = A*B.
//
// Reference: set_row_nnz + set_row_nz_bin_* (kernel_spgemm_hash_d.cu:266-622, 1077-1185).
// What is kept: row-wise Gustavson, rows binned by an upper bound of their size, one hash set
// per row probed linearly, table in shared memory (the hash itself is Fibonacci hashing, common.cuh
// hash_slot, not the reference's (col * 107) & (size - 1): DESIGN.md 4.1).
// What is re-designed for B200:
//   * ladder derived from 227 KiB of shared memory: hash sets up to 32768 keys (reference: 8192);
//     every table is sized per ROW (next pow2 of 4/3 * bound), so clearing it costs what the row
//     needs, not what the bin allows, and the load factor never exceeds 3/4 (reference: up to 1).
//   * rows above the hash ladder use a shared-memory BITMAP over a column tile instead of the
//     try-then-redo pair each_tb_large / each_gl (:474-622) whose global table needs
//     fail_count * max_intprod ints.  Bits are tested before the atomicOr, so the atomic count is
//     nnz(C_i), not the number of products.  Columns beyond one tile (N > ~1.7 M) take one pass
//     per tile; memory use is fixed and there is no failure path.
//   * products of a row are spread evenly over the threads (spgemm_device.cuh), whatever the
//     B-row lengths; persistent CTAs pull rows from a queue ordered heaviest-first; no
//     warp-synchronous assumptions: every phase boundary is a barrier.
#include "context.h"
#include "spgemm_device.cuh"
#include "spgemm_plan.h"

namespace nsp {

// ---- bin 0: at most 32 products; 4 threads per row, 64-slot table (ref: set_row_nz_bin_pwarp) ----
constexpr int kPw = 4;
constexpr int kPwSymSlots = 64;

__global__ void __launch_bounds__(256)
sym_pwarp_kernel(const int *__restrict__ a_rpt, const int *__restrict__ a_col,
                 const int *__restrict__ b_rpt, const int *__restrict__ b_col,
                 const int *__restrict__ row_perm, int *__restrict__ row_cnt,
                 const int *__restrict__ bins)
{
    __shared__ int tab[(256 / kPw) * kPwSymSlots];
    int lo, hi;
    class_range(bins, 0, 0, lo, hi);
    const int n = hi - lo;
    const int lr = threadIdx.x / kPw, t = threadIdx.x % kPw;
    int *my = tab + lr * kPwSymSlots;
    for (int base = blockIdx.x * (256 / kPw); base < n; base += gridDim.x * (256 / kPw)) {
        for (int i = t; i < kPwSymSlots; i += kPw) my[i] = kEmptyKey;
        __syncwarp();
        const int r = base + lr;
        int cnt = 0, rid = 0;
        if (r < n) {
            rid = row_perm[lo + r];
            const int a_end = a_rpt[rid + 1];
            for (int j = a_rpt[rid] + t; j < a_end; j += kPw) {
                const int ac = ld_stream(a_col + j);
                const int ke = ld_nc(b_rpt + ac + 1);
                for (int k = ld_nc(b_rpt + ac); k < ke; ++k)
                    cnt += hash_insert_key(my, kPwSymSlots - 1, ld_nc(b_col + k));
            }
        }
        cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
        if (r < n && t == 0) row_cnt[rid] = cnt;
        __syncwarp();
    }
}

// ---- hash classes: a group (warp or CTA) per row, table of `tmax` keys per group -----------------
template <int GROUP, int BS>
__global__ void __launch_bounds__(BS, (BS >= 1024 ? 1 : 2))
sym_hash_kernel(const int *__restrict__ a_rpt, const int *__restrict__ a_col,
                const int *__restrict__ b_rpt, const int *__restrict__ b_col,
                const int *__restrict__ row_perm, const int *__restrict__ row_ip,
                int *__restrict__ row_cnt, int *__restrict__ bins, int bin_lo, int bin_hi, int queue,
                int tmax)
{
    extern __shared__ int smem_i[];
    constexpr int NG = BS / GROUP;
    __shared__ FlatScratch<GROUP, float> s_flat[NG];
    __shared__ int s_row, s_cnt;
    const int g = threadIdx.x / GROUP, t = threadIdx.x % GROUP;
    int *tab = smem_i + g * tmax;
    int lo, hi;
    class_range(bins, bin_lo, bin_hi, lo, hi);
    const int n = hi - lo;
    while (true) {
        int r;
        if (GROUP == 32) {
            r = 0;
            if (t == 0) r = atomicAdd(&bins[kBinQueue + queue], 1);
            r = __shfl_sync(0xffffffffu, r, 0);
        } else {
            if (t == 0) {
                s_row = atomicAdd(&bins[kBinQueue + queue], 1);
                s_cnt = 0;
            }
            __syncthreads();
            r = s_row;
        }
        if (r >= n) break;
        const int rid = row_perm[lo + r];
        const int tsize = table_size_for(row_ip[rid], tmax);
        const unsigned mask = (unsigned)tsize - 1u;
        for (int i = t; i < tsize; i += GROUP) tab[i] = kEmptyKey;
        group_sync<GROUP>();
        int cnt = 0;
        for_each_product<GROUP, false, float>(
            t, a_rpt[rid], a_rpt[rid + 1], a_col, (const float *)nullptr, b_rpt, b_col,
            (const float *)nullptr, s_flat[g], [&](int c, float) { cnt += hash_insert_key(tab, mask, c); });
        cnt = warp_sum(cnt);
        if (GROUP == 32) {
            if (t == 0) row_cnt[rid] = cnt;
            __syncwarp();
        } else {
            if ((t & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
            __syncthreads();
            if (t == 0) row_cnt[rid] = s_cnt;
            __syncthreads();
        }
    }
}

// ---- bitmap class: one CTA per row, one pass over the row's products per column window -----------
// Window of W = 2^wshift columns (W/8 bytes of shared memory, up to 2^20).  Every product sets its bit
// with one shared-memory atomicOr per distinct 32-bit word a lane touches (plain bit order, word pairs XOR-
// swizzled inside their batch: spgemm_device.cuh bitmap_word32), the row's count
// is the popcount of the window, taken by the sweep that also clears it for the next row.  With more
// than one window the sub-range of each B row is found by binary search, so a product is still read
// exactly once.
template <int BS, bool kSorted, bool kFlat>
__global__ void __launch_bounds__(BS, 1)
sym_bitmap_kernel(const int *__restrict__ a_rpt, const int *__restrict__ a_col,
                  const int *__restrict__ b_rpt, const int *__restrict__ b_col,
                  const int *__restrict__ row_perm, int *__restrict__ row_cnt, int *__restrict__ bins,
                  int bin_lo, int bin_hi, int queue, int N, int wshift, int b_vec_end)
{
    extern __shared__ __align__(16) int smem_i[];
    unsigned *bm = reinterpret_cast<unsigned *>(smem_i);
    uint4 *bm4 = reinterpret_cast<uint4 *>(smem_i);
    __shared__ PartScratch<BS, float> s_part;
    __shared__ int s_row, s_cnt;
    const int t = threadIdx.x;
    const unsigned W = 1u << wshift;
    const int nw4 = (int)(W >> 7);
    const int nwin = (int)(((unsigned)N + W - 1u) >> wshift);
    int lo, hi;
    class_range(bins, bin_lo, bin_hi, lo, hi);
    const int n = hi - lo;
    for (int i = t; i < nw4; i += BS) bm4[i] = make_uint4(0u, 0u, 0u, 0u);
    while (true) {
        if (t == 0) {
            s_row = atomicAdd(&bins[kBinQueue + queue], 1);
            s_cnt = 0;
        }
        __syncthreads();
        const int r = s_row;
        if (r >= n) break;
        const int rid = row_perm[lo + r];
        const int a_beg = a_rpt[rid], a_end = a_rpt[rid + 1];
        int cnt = 0;
        for (int win = 0; win < nwin; ++win) {
            const int c0 = (int)((unsigned)win << wshift);
            const int c1 = (int)min((unsigned)N, (unsigned)c0 + W);
            for (int base = a_beg; base < a_end; base += BS) {
                const int total = stage_parts_range<BS, false, float, kFlat>(t, base, a_end, a_col, (const float *)nullptr,
                                                                             b_rpt, b_col, c0, c1, kSorted && win > 0,
                                                                             kSorted && win < nwin - 1, s_part);
                if (kFlat) {
                    // short B rows (products per entry of A below 48 on average): flat traversal, one bit per product
                    run_flat<BS, false, float>(t, total, b_col, (const float *)nullptr, s_part, [&](int c, float) {
                        const unsigned cc = (unsigned)(c - c0);
                        atomicOr(bm + bitmap_word32(cc >> 5), 1u << (cc & 31u));
                    });
                } else {
                    run_parts_mark<BS, !kSorted, float>(t, total, b_col, b_vec_end, s_part, bm, c0, (unsigned)(c1 - c0));
                }
            }
            // count and clear in one sweep (run_parts ended with a barrier)
            for (int i = t; i < nw4; i += BS) {
                const uint4 v = bm4[i];
                cnt += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
                bm4[i] = make_uint4(0u, 0u, 0u, 0u);
            }
            // the next marks are ordered after this sweep by the barriers of stage_parts_range
        }
        cnt = warp_sum(cnt);
        if ((t & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
        __syncthreads();
        if (t == 0) row_cnt[rid] = s_cnt;
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static void prof_class(nsp_context *ctx, const char *name, int bin_lo, int bin_hi)
{
    if (!ctx->profile) return;
    long long rows = 0, ip = 0, len = 0;
    for (int b = bin_lo; b <= bin_hi; ++b) {
        rows += ctx->sp.h_bins[kBinHist + b];
        ip += (long long)ctx->sp.h_binsum[kSumIp + b];
        len += (long long)ctx->sp.h_binsum[kSumLen + b];
    }
    ctx->prof_begin(name, rows, ip, len);
}

static long long rows_in(const nsp_spgemm_state &sp, int bin_lo, int bin_hi)
{
    long long n = 0;
    for (int b = bin_lo; b <= bin_hi; ++b) n += sp.h_bins[kBinHist + b];
    return n;
}

template <int GROUP, int BS>
static int launch_sym_hash(nsp_context *ctx, const char *name, int grid, size_t smem, const int *a_rpt,
                           const int *a_col, const int *b_rpt, const int *b_col, int bin_lo, int bin_hi,
                           int queue, int tmax)
{
    nsp_spgemm_state &sp = ctx->sp;
    auto kern = sym_hash_kernel<GROUP, BS>;
    NSP_CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prof_class(ctx, name, bin_lo, bin_hi);
    kern<<<grid, BS, smem, ctx->stream>>>(a_rpt, a_col, b_rpt, b_col, sp.d_row_perm, sp.d_row_ip, sp.d_row_cnt,
                                          sp.d_bins, bin_lo, bin_hi, queue, tmax);
    ctx->prof_end();
    ctx->launches += 1;
    NSP_CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}

static inline int imin(long long a, long long b) { return (int)(a < b ? a : b); }

int spgemm_symbolic(nsp_context *ctx, int M, int K, int N, const int *a_rpt, const int *a_col,
                    const int *b_rpt, const int *b_col, long long *c_rpt64, long long *h_nnz,
                    long long *h_ip)
{
    if (M < 0 || K < 0 || N < 0 || !c_rpt64) return ctx->fail(-2, "nsp_spgemm_symbolic: bad argument");
    nsp_spgemm_state &sp = ctx->sp;
    sp.symbolic_done = false;
    sp.M = M;
    sp.K = K;
    sp.N = N;
    if (plan_reserve(ctx, M) != 0) return -1;
    if (plan_by_intprod(ctx, M, K, N > 0 ? N : 1, a_rpt, a_col, b_rpt, b_col) != 0) return -1;

    // ---- class ladder (symbolic shift 5: bin b holds 2^(4+b) < v <= 2^(5+b)) ----
    //   bin 0          <= 32        4 threads / row, 64 slots
    //   bins 1..4      <= 512       warp / row, <= 1024 slots, 8 rows per CTA
    //   bins 5..7      <= 4096      CTA(256) / row, <= 8192 slots (32 KiB)
    //   bins 8..9      <= 16384     CTA(1024) / row, <= 32768 slots (128 KiB)
    //   bins >= bm_bin              CTA(1024) / row, bitmap over column tiles
    // bitmap window: power of two >= N in [2^16, 2^20] columns (8 KiB .. 128 KiB of shared memory)
    int ws_max = ctx->opt_sym_window_shift > 0 ? (int)ctx->opt_sym_window_shift : 20;
    if (ws_max < 16) ws_max = 16;
    if (ws_max > 20) ws_max = 20;
    int wshift = 16;
    while (wshift < ws_max && (1ll << wshift) < (long long)N) ++wshift;
    // Class boundary: a bitmap row costs a clear + count sweep per window (~4 us at 2^20 columns), the
    // hash set costs per product; measured on R-MAT scale 20 the bitmap wins above ~4096 products per row.
    // With many windows (very wide C) the hash ladder keeps everything it can hold (16384 products).
    const long long nwin_host = ((long long)N + (1ll << wshift) - 1) >> wshift;
    int bm_bin = nwin_host <= 4 ? 8 : 10;
    if (ctx->opt_sym_bitmap_min >= 0) {
        bm_bin = log_bin(imin(ctx->opt_sym_bitmap_min, 0x7fffffff), kSymShift) + 1;
        if (bm_bin < 1) bm_bin = 1;
        if (bm_bin > 10) bm_bin = 10;
    }
    const int sms = ctx->sm_count;
    if (M > 0) {
        // heaviest first
        if (rows_in(sp, bm_bin, kNumBins - 1) > 0) {
            const size_t smem = (size_t)1 << (wshift - 3);
            const int grid = imin(rows_in(sp, bm_bin, kNumBins - 1), (long long)sms);
            long long cls_ip = 0, cls_len = 0;
            for (int b = bm_bin; b < kNumBins; ++b) {
                cls_ip += (long long)sp.h_binsum[kSumIp + b];
                cls_len += (long long)sp.h_binsum[kSumLen + b];
            }
            // (flat traversal: every pass touches every product, so only with few windows -- measured on the full-size
            // configs, profiles/r2_ab_flat_traversal_windows.txt: C4, 8 windows, 54.9 ms flat / 65.0 ms searched sub-ranges;
            // C5, 16 windows, 157.9 / 81.4 ms)
            const bool flat = sp.b_sorted && ((cls_len > 0 && cls_ip < 48 * cls_len && nwin_host <= 8 && !ctx->opt_no_flat) || ctx->opt_no_flat < 0);
            auto kern = !sp.b_sorted ? sym_bitmap_kernel<1024, false, false>
                                     : (flat ? sym_bitmap_kernel<1024, true, true> : sym_bitmap_kernel<1024, true, false>);
            const int b_vec_end = ((reinterpret_cast<uintptr_t>(b_col) & 15u) != 0 || ctx->opt_no_vec)
                                      ? 0 : (int)(sp.b_nnz & ~3ll);
            NSP_CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            prof_class(ctx, "sym_bitmap", bm_bin, kNumBins - 1);
            kern<<<grid, 1024, smem, ctx->stream>>>(a_rpt, a_col, b_rpt, b_col, sp.d_row_perm, sp.d_row_cnt,
                                                    sp.d_bins, bm_bin, kNumBins - 1, 4, N, wshift,
                                                    b_vec_end);
            ctx->prof_end();
            ctx->launches += 1;
            NSP_CUDA_TRY(ctx, cudaGetLastError());
        }
        if (bm_bin > 8 && rows_in(sp, 8, imin(9, bm_bin - 1)) > 0) {
            const int hi = imin(9, bm_bin - 1);
            const int tmax = 32768;
            const int grid = imin(rows_in(sp, 8, hi), sms);
            if (launch_sym_hash<1024, 1024>(ctx, "sym_hash_cta1024", grid, (size_t)tmax * 4, a_rpt, a_col, b_rpt,
                                            b_col, 8, hi, 3, tmax) != 0)
                return -1;
        }
        if (bm_bin > 5 && rows_in(sp, 5, imin(7, bm_bin - 1)) > 0) {
            const int hi = imin(7, bm_bin - 1);
            const int tmax = 8192;
            const int grid = imin(rows_in(sp, 5, hi), (long long)sms * 6);
            if (launch_sym_hash<256, 256>(ctx, "sym_hash_cta256", grid, (size_t)tmax * 4, a_rpt, a_col, b_rpt,
                                          b_col, 5, hi, 2, tmax) != 0)
                return -1;
        }
        if (rows_in(sp, 1, imin(4, bm_bin - 1)) > 0) {
            const int hi = imin(4, bm_bin - 1);
            const int tmax = 1024;
            const int grid = imin((rows_in(sp, 1, hi) + 7) / 8, (long long)sms * 6);
            if (launch_sym_hash<32, 256>(ctx, "sym_hash_warp", grid, (size_t)tmax * 4 * 8, a_rpt, a_col, b_rpt,
                                         b_col, 1, hi, 1, tmax) != 0)
                return -1;
        }
        if (rows_in(sp, 0, 0) > 0) {
            const int grid = imin((rows_in(sp, 0, 0) + 63) / 64, (long long)sms * 8);
            prof_class(ctx, "sym_pwarp", 0, 0);
            sym_pwarp_kernel<<<grid, 256, 0, ctx->stream>>>(a_rpt, a_col, b_rpt, b_col, sp.d_row_perm,
                                                          sp.d_row_cnt, sp.d_bins);
            ctx->prof_end();
            ctx->launches += 1;
            NSP_CUDA_TRY(ctx, cudaGetLastError());
        }
    }
    if (scan_row_counts(ctx, M, c_rpt64) != 0) return -1;
    NSP_CUDA_TRY(ctx, cudaMemcpyAsync(sp.h_scalars, sp.d_scalars, sizeof(long long) * 8,
                                      cudaMemcpyDeviceToHost, ctx->stream));
    NSP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (h_nnz) *h_nnz = sp.h_scalars[kScalarNnz];
    if (h_ip) *h_ip = sp.h_scalars[kScalarIp];
    sp.symbolic_done = true;
    return 0;
}

}  // namespace nsp
