// nsparse-b200: NUMERIC phase of the hash SpGEMM -- column indices and values of C = A*B.
//
// Reference: set_min_bin + calculate_value_col_bin_* (kernel_spgemm_hash_d.cu:201-246, 631-1027,
// 1187-1288).  Kept: rows re-binned by their exact nnz from the symbolic phase, per-row hash table
// of (column, value) in shared memory, CAS on the key + floating-point atomic add on the value,
// output rows sorted by ascending column, numerical zeros kept.  Re-designed for B200:
//   * tables up to 16384 (key,value) slots (192 KiB in fp64; reference: 4096), sized per row and
//     never above 3/4 load;
//   * the whole table is sorted bitonically with the free slots (key 0xffffffff) sinking to the
//     end, which replaces the global-atomic compaction (:904-912) AND the O(nnz^2) counting sort
//     (:917-925) with O(n log^2 n) shared-memory work and no global atomics;
//   * rows above the hash ladder (reference: each_gl with 2*max_nz global slots PER ROW and an
//     O(nnz^2) sort in global memory, :929-1027) use a column-tile BITMAP + RANK scheme: pass 1
//     marks the row's columns in a shared-memory bitmap, a CTA-wide scan turns it into ranks,
//     C.col is emitted straight from the bitmap (already sorted), pass 2 adds every product into
//     C.val[rpt + rank(col)] with red.global.add.  No table, no compaction, no sort, no workspace;
//   * native fp64 atomics everywhere (reference SpMV/SpGEMM fall back to CAS loops on fp64).
#pragma once

#include "context.h"
#include "spgemm_device.cuh"
#include "spgemm_plan.h"

namespace nsp {

// ---- bin 0: nnz(C_i) <= 16; 4 threads per row, 32 slots (ref: calculate_value_col_bin_pwarp) ----
constexpr int kPwNum = 4;
constexpr int kPwNumSlots = 32;

template <typename real>
__global__ void __launch_bounds__(256)
num_pwarp_kernel(const int *__restrict__ a_rpt, const int *__restrict__ a_col,
                 const real *__restrict__ a_val, const int *__restrict__ b_rpt,
                 const int *__restrict__ b_col, const real *__restrict__ b_val,
                 const long long *__restrict__ c_rpt, int *__restrict__ c_col, real *__restrict__ c_val,
                 const int *__restrict__ row_perm, const int *__restrict__ bins)
{
    __shared__ int keys[(256 / kPwNum) * kPwNumSlots];
    __shared__ real vals[(256 / kPwNum) * kPwNumSlots];
    int lo, hi;
    class_range(bins, 0, 0, lo, hi);
    const int n = hi - lo;
    const int lr = threadIdx.x / kPwNum, t = threadIdx.x % kPwNum;
    int *mk = keys + lr * kPwNumSlots;
    real *mv = vals + lr * kPwNumSlots;
    for (int base = blockIdx.x * (256 / kPwNum); base < n; base += gridDim.x * (256 / kPwNum)) {
        for (int i = t; i < kPwNumSlots; i += kPwNum) {
            mk[i] = kEmptyKey;
            mv[i] = real(0);
        }
        __syncwarp();
        const int r = base + lr;
        int rid = 0;
        if (r < n) {
            rid = row_perm[lo + r];
            const int a_end = a_rpt[rid + 1];
            for (int j = a_rpt[rid] + t; j < a_end; j += kPwNum) {
                const int ac = ld_stream(a_col + j);
                const real av = ld_stream(a_val + j);
                const int ke = ld_nc(b_rpt + ac + 1);
                for (int k = ld_nc(b_rpt + ac); k < ke; ++k)
                    hash_accumulate(mk, mv, kPwNumSlots - 1, ld_nc(b_col + k), av * ld_nc(b_val + k));
            }
        }
        __syncwarp();
        if (r < n) {
            // rank of every occupied slot among the row's keys = its position in the sorted row
            const long long off = c_rpt[rid];
            for (int i = t; i < kPwNumSlots; i += kPwNum) {
                const int key = mk[i];
                if (key == kEmptyKey) continue;
                int rank = 0;
#pragma unroll
                for (int q = 0; q < kPwNumSlots; ++q) rank += (unsigned)mk[q] < (unsigned)key;
                c_col[off + rank] = key;
                c_val[off + rank] = mv[i];
            }
        }
        __syncwarp();
    }
}

// ---- hash classes -------------------------------------------------------------------------------
template <typename real, int GROUP, int BS>
__global__ void __launch_bounds__(BS, (BS >= 1024 ? 1 : 2))
num_hash_kernel(const int *__restrict__ a_rpt, const int *__restrict__ a_col,
                const real *__restrict__ a_val, const int *__restrict__ b_rpt,
                const int *__restrict__ b_col, const real *__restrict__ b_val,
                const long long *__restrict__ c_rpt, int *__restrict__ c_col, real *__restrict__ c_val,
                const int *__restrict__ row_perm, int *__restrict__ bins, int bin_lo, int bin_hi,
                int queue, int tmax)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NG = BS / GROUP;
    __shared__ FlatScratch<GROUP, real> s_flat[NG];
    __shared__ int s_row;
    const int g = threadIdx.x / GROUP, t = threadIdx.x % GROUP;
    // values first (8-byte aligned for fp64), then keys
    real *vals = reinterpret_cast<real *>(smem_raw) + (size_t)g * tmax;
    int *keys = reinterpret_cast<int *>(smem_raw + sizeof(real) * (size_t)NG * tmax) + (size_t)g * tmax;
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
            if (t == 0) s_row = atomicAdd(&bins[kBinQueue + queue], 1);
            __syncthreads();
            r = s_row;
        }
        if (r >= n) break;
        const int rid = row_perm[lo + r];
        const long long off = c_rpt[rid];
        const int nnz = (int)(c_rpt[rid + 1] - off);
        const int tsize = table_size_for(nnz, tmax);
        const unsigned mask = (unsigned)tsize - 1u;
        for (int i = t; i < tsize; i += GROUP) {
            keys[i] = kEmptyKey;
            vals[i] = real(0);
        }
        group_sync<GROUP>();
        for_each_product<GROUP, true, real>(t, a_rpt[rid], a_rpt[rid + 1], a_col, a_val, b_rpt, b_col, b_val,
                                            s_flat[g],
                                            [&](int c, real v) { hash_accumulate(keys, vals, mask, c, v); });
        group_sync<GROUP>();
        bitonic_sort_slots<GROUP, real>(keys, vals, tsize, t);
        for (int i = t; i < nnz; i += GROUP) {
            c_col[off + i] = keys[i];
            c_val[off + i] = vals[i];
        }
        group_sync<GROUP>();
    }
}

// ---- bitmap + rank class ------------------------------------------------------------------------
// shared memory: bm[nw] 64-bit bitmap words of the column tile, pre[nw] exclusive popcount prefix.
// Every warp owns a contiguous segment of the words (lanes interleaved, so the 8-byte reads are
// bank-conflict free): the prefix needs two barriers per tile whatever N is, and a lane emits the
// columns of its words as one contiguous run of C.col.
template <typename real, int BS, bool kSingle>
__global__ void __launch_bounds__(BS, 1)
num_bitmap_kernel(const int *__restrict__ a_rpt, const int *__restrict__ a_col,
                  const real *__restrict__ a_val, const int *__restrict__ b_rpt,
                  const int *__restrict__ b_col, const real *__restrict__ b_val,
                  const long long *__restrict__ c_rpt, int *__restrict__ c_col, real *__restrict__ c_val,
                  const int *__restrict__ row_perm, int *__restrict__ bins, int bin_lo, int bin_hi,
                  int queue, int N, int tile_cols, int dbg)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = BS / 32;
    __shared__ PartScratch<BS, real> s_part;
    __shared__ int s_row;
    __shared__ int s_wsum[NW];
    const int tile_words = (tile_cols + 63) >> 6;
    unsigned long long *bm = reinterpret_cast<unsigned long long *>(smem_raw);
    unsigned *bm32 = reinterpret_cast<unsigned *>(smem_raw);
    int *pre = reinterpret_cast<int *>(smem_raw + sizeof(unsigned long long) * (size_t)((tile_words + 1) & ~1));
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    int lo, hi;
    class_range(bins, bin_lo, bin_hi, lo, hi);
    const int n = hi - lo;
    while (true) {
        if (t == 0) s_row = atomicAdd(&bins[kBinQueue + queue], 1);
        __syncthreads();
        const int r = s_row;
        if (r >= n) break;
        const int rid = row_perm[lo + r];
        const int a_beg = a_rpt[rid], a_end = a_rpt[rid + 1];
        long long out = c_rpt[rid];
        for (int t0 = 0; t0 < N; t0 += tile_cols) {
            const int ncols = min(tile_cols, N - t0);
            const int nw = (ncols + 63) >> 6;
            {
                uint4 *bm4 = reinterpret_cast<uint4 *>(bm);
                const int nw4 = (nw + 1) >> 1;
                for (int i = t; i < nw4; i += BS) bm4[i] = make_uint4(0u, 0u, 0u, 0u);
            }
            __syncthreads();
            // pass 1: structure of the tile.  A row whose A entries fit one slab (nearly all) is
            // staged once, with its values, and pass 2 walks the same staged parts again.
            const bool one_slab = a_end - a_beg <= BS;
            int staged_total = 0;
            auto mark = [&](int c, real) {
                const unsigned cc = (unsigned)(c - t0);
                if (kSingle || cc < (unsigned)ncols) {
                    const unsigned bit = 1u << (cc & 31);
                    unsigned *w = bm32 + (cc >> 5);
                    if (!(*((volatile unsigned *)w) & bit)) atomicOr(w, bit);
                }
            };
            if (one_slab) {
                if (t == 0) s_part.next = NW;
                staged_total = stage_parts<BS, true, real>(t, a_beg, a_end, a_col, a_val, b_rpt, s_part);
                run_parts<BS, false, real>(t, staged_total, b_col, b_val, s_part, mark);
            } else {
                for_each_product_parts<BS, false, real>(t, a_beg, a_end, a_col, a_val, b_rpt, b_col, b_val, s_part,
                                                        mark);
            }
            // exclusive prefix of the per-word popcounts: warp totals, then per-warp running scan
            const int seg = ((nw + BS - 1) / BS) * 32;         // words per warp (multiple of 32)
            const int w0 = wid * seg, w1 = min(nw, w0 + seg);
            int tot = 0;
            for (int j = w0 + lane; j < w1; j += 32) tot += __popcll(bm[j]);
            tot = warp_sum(tot);
            if (lane == 0) s_wsum[wid] = tot;
            __syncthreads();
            int carry, tile_nnz;
            {
                const int ws = lane < NW ? s_wsum[lane] : 0;
                carry = warp_sum(lane < wid ? ws : 0);
                tile_nnz = warp_sum(ws);
            }
            for (int jb = w0; jb < w1; jb += 32) {
                const int j = jb + lane;
                const int c = j < w1 ? __popcll(bm[j]) : 0;
                int inc = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                if (j < w1) pre[j] = carry + inc - c;
                carry += __shfl_sync(0xffffffffu, inc, 31);
            }
            // values start at zero (coalesced), columns come straight from the bitmap (sorted)
            if (!(dbg & 4)) for (int i = t; i < tile_nnz; i += BS) c_val[out + i] = real(0);
            if (!(dbg & 1)) {
                int *crow = c_col + out;
                for (int jb = w0; jb < w1; jb += 32) {
                    const int j = jb + lane;
                    if (j < w1) {
                        // 32-bit halves, 32-bit output index: ~9 instructions per emitted column
                        unsigned lo32 = bm32[2 * j], hi32 = bm32[2 * j + 1];
                        int idx = pre[j];
                        const int cbase = t0 + (j << 6);
                        while (lo32) {
                            crow[idx++] = cbase + __ffs((int)lo32) - 1;
                            lo32 &= lo32 - 1;
                        }
                        while (hi32) {
                            crow[idx++] = cbase + 31 + __ffs((int)hi32);
                            hi32 &= hi32 - 1;
                        }
                    }
                }
            }
            __syncthreads();
            // pass 2: values
            real *cv = c_val + out;
            auto add = [&](int c, real v) {
                const unsigned cc = (unsigned)(c - t0);
                if (kSingle || cc < (unsigned)ncols) {
                    // rank = prefix of the 64-bit word + set bits below cc, in 32-bit operations
                    const unsigned w = cc >> 6;
                    const uint2 word = *reinterpret_cast<const uint2 *>(bm + w);
                    const unsigned below = (1u << (cc & 31)) - 1u;
                    const bool upper = (cc & 32) != 0;
                    const int rank = pre[w] + __popc(word.x & (upper ? 0xffffffffu : below)) +
                                     __popc(word.y & (upper ? below : 0u));
                    atomicAdd(cv + rank, v);
                }
            };
            if (dbg & 2) {
            } else if (one_slab)
                run_parts<BS, true, real>(t, staged_total, b_col, b_val, s_part, add);
            else
                for_each_product_parts<BS, true, real>(t, a_beg, a_end, a_col, a_val, b_rpt, b_col, b_val, s_part,
                                                       add);
            out += tile_nnz;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static inline long long num_rows_in(const nsp_spgemm_state &sp, int bin_lo, int bin_hi)
{
    long long n = 0;
    for (int b = bin_lo; b <= bin_hi; ++b) n += sp.h_bins[kBinHist + b];
    return n;
}

static inline int num_imin(long long a, long long b) { return (int)(a < b ? a : b); }

static inline void num_prof_class(nsp_context *ctx, const char *name, int bin_lo, int bin_hi)
{
    if (!ctx->profile) return;
    long long rows = 0, ip = 0, len = 0, out = 0;
    for (int b = bin_lo; b <= bin_hi; ++b) {
        rows += ctx->sp.h_bins[kBinHist + b];
        ip += (long long)ctx->sp.h_binsum[kSumIp + b];
        len += (long long)ctx->sp.h_binsum[kSumLen + b];
        out += (long long)ctx->sp.h_binsum[kSumCnt + b];
    }
    ctx->prof_begin(name, rows, ip, len, out);
}

#define NSP_NUM_ARGS                                                                               \
    a_rpt, a_col, a_val, b_rpt, b_col, b_val, c_rpt64, c_col, c_val, sp.d_row_perm, sp.d_bins

template <typename real, int GROUP, int BS>
static int launch_num_hash(nsp_context *ctx, const char *name, int grid, size_t smem, const int *a_rpt,
                           const int *a_col, const real *a_val, const int *b_rpt, const int *b_col,
                           const real *b_val, const long long *c_rpt64, int *c_col, real *c_val,
                           int bin_lo, int bin_hi, int queue, int tmax)
{
    nsp_spgemm_state &sp = ctx->sp;
    auto kern = num_hash_kernel<real, GROUP, BS>;
    NSP_CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    num_prof_class(ctx, name, bin_lo, bin_hi);
    kern<<<grid, BS, smem, ctx->stream>>>(NSP_NUM_ARGS, bin_lo, bin_hi, queue, tmax);
    ctx->prof_end();
    ctx->launches += 1;
    NSP_CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}

template <typename real>
int spgemm_numeric(nsp_context *ctx, int M, int K, int N, const int *a_rpt, const int *a_col,
                   const real *a_val, const int *b_rpt, const int *b_col, const real *b_val,
                   const long long *c_rpt64, int *c_col, real *c_val)
{
    nsp_spgemm_state &sp = ctx->sp;
    if (!sp.symbolic_done || sp.M != M || sp.K != K || sp.N != N)
        return ctx->fail(-2, "nsp_spgemm_numeric: call nsp_spgemm_symbolic on the same context and shapes first");
    if (M == 0) return 0;
    if (plan_by_count(ctx, M, kNumShift, a_rpt) != 0) return -1;

    // ---- class ladder (numeric shift 4: bin b holds 2^(3+b) < nnz <= 2^(4+b)) ----
    //   bin 0          <= 16        4 threads / row, 32 slots
    //   bins 1..4      <= 256       warp / row, <= 512 slots, 8 rows per CTA
    //   bins 5..7      <= 2048      CTA(256) / row, <= 4096 slots
    //   bins 8..9      <= 8192      CTA(1024) / row, <= 16384 slots (128 KiB fp32 / 192 KiB fp64)
    //   bins >= bm_bin              CTA(1024) / row, bitmap + rank over column tiles
    const int smem_cap = ctx->max_smem_optin - kStaticSmemReserve;
    const int tile_max = ((smem_cap - 64) / 24) * 128;
    const int tile_cols = N < tile_max ? ((N + 127) / 128) * 128 : tile_max;
    const int slot_bytes = 4 + (int)sizeof(real);
    int bm_bin = 10;
    if (N <= tile_max) {
        // single tile: bitmap + rank needs no sort, the hash path pays an O(n log^2 n) bitonic sort
        // per row; measured crossover on R-MAT ~N/1024 entries per row
        const int v = N / 1024 + 1;
        bm_bin = log_bin(v, kNumShift) + 1;
        if (bm_bin < 5) bm_bin = 5;
        if (bm_bin > 10) bm_bin = 10;
    }
    if (ctx->opt_num_bitmap_min >= 0) {
        bm_bin = log_bin(num_imin(ctx->opt_num_bitmap_min, 0x7fffffff), kNumShift) + 1;
        if (bm_bin < 1) bm_bin = 1;
        if (bm_bin > 10) bm_bin = 10;
    }
    const int sms = ctx->sm_count;
    if (num_rows_in(sp, bm_bin, kNumBins - 1) > 0) {
        const size_t tw = (size_t)(tile_cols + 63) / 64;
        const size_t smem = ((tw + 1) & ~size_t(1)) * 8 + tw * 4 + 16;
        const int grid = num_imin(num_rows_in(sp, bm_bin, kNumBins - 1), (long long)sms);
        auto kern = N <= tile_max ? num_bitmap_kernel<real, 1024, true> : num_bitmap_kernel<real, 1024, false>;
        NSP_CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        num_prof_class(ctx, "num_bitmap", bm_bin, kNumBins - 1);
        kern<<<grid, 1024, smem, ctx->stream>>>(NSP_NUM_ARGS, bm_bin, kNumBins - 1, 4, N, tile_cols, (int)ctx->opt_debug);
        ctx->prof_end();
        ctx->launches += 1;
        NSP_CUDA_TRY(ctx, cudaGetLastError());
    }
    if (bm_bin > 8 && num_rows_in(sp, 8, num_imin(9, bm_bin - 1)) > 0) {
        const int hi = num_imin(9, bm_bin - 1);
        const int tmax = 16384;
        const int grid = num_imin(num_rows_in(sp, 8, hi), sms);
        if (launch_num_hash<real, 1024, 1024>(ctx, "num_hash_cta1024", grid, (size_t)tmax * slot_bytes, a_rpt,
                                              a_col, a_val, b_rpt, b_col, b_val, c_rpt64, c_col, c_val, 8, hi, 3,
                                              tmax) != 0)
            return -1;
    }
    if (bm_bin > 5 && num_rows_in(sp, 5, num_imin(7, bm_bin - 1)) > 0) {
        const int hi = num_imin(7, bm_bin - 1);
        const int tmax = 4096;
        const int grid = num_imin(num_rows_in(sp, 5, hi), (long long)sms * 4);
        if (launch_num_hash<real, 256, 256>(ctx, "num_hash_cta256", grid, (size_t)tmax * slot_bytes, a_rpt, a_col,
                                            a_val, b_rpt, b_col, b_val, c_rpt64, c_col, c_val, 5, hi, 2, tmax) != 0)
            return -1;
    }
    if (num_rows_in(sp, 1, num_imin(4, bm_bin - 1)) > 0) {
        const int hi = num_imin(4, bm_bin - 1);
        const int tmax = 512;
        const int grid = num_imin((num_rows_in(sp, 1, hi) + 7) / 8, (long long)sms * 4);
        if (launch_num_hash<real, 32, 256>(ctx, "num_hash_warp", grid, (size_t)tmax * slot_bytes * 8, a_rpt, a_col,
                                           a_val, b_rpt, b_col, b_val, c_rpt64, c_col, c_val, 1, hi, 1, tmax) != 0)
            return -1;
    }
    if (num_rows_in(sp, 0, 0) > 0) {
        const int grid = num_imin((num_rows_in(sp, 0, 0) + 63) / 64, (long long)sms * 8);
        num_prof_class(ctx, "num_pwarp", 0, 0);
        num_pwarp_kernel<real><<<grid, 256, 0, ctx->stream>>>(a_rpt, a_col, a_val, b_rpt, b_col, b_val,
                                                              c_rpt64, c_col, c_val, sp.d_row_perm, sp.d_bins);
        ctx->prof_end();
        ctx->launches += 1;
        NSP_CUDA_TRY(ctx, cudaGetLastError());
    }
    return 0;
}

#undef NSP_NUM_ARGS

}  // namespace nsp
