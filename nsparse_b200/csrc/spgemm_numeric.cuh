// nsparse-b200: NUMERIC phase of the hash SpGEMM -- column indices and values of C = A*B.
//
// Reference: set_min_bin + calculate_value_col_bin_* (kernel_spgemm_hash_d.cu:201-246, 631-1027,
// 1187-1288).  Kept: rows re-binned by their exact nnz from the symbolic phase, per-row hash table
// of (column, value) in shared memory, CAS on the key + floating-point atomic add on the value,
// output rows sorted by ascending column, numerical zeros kept.  Re-designed for B200:
//   * tables up to 16384 (key,value) slots (192 KiB in fp64; reference: 4096), sized per row and
//     never above 3/4 load;
//   * the row is brought into column order by BUCKETS of the column range (count, scan, scatter; small buckets are
//     ordered inside shared memory, clustered ones ranked by counting, see num_hash_kernel), with a bitonic sort of
//     the whole table -- free slots (key 0xffffffff) sinking to the end -- as the fallback: this replaces the
//     global-atomic compaction (:904-912) AND the O(nnz^2) counting sort (:917-925);
//   * rows above ~2048 entries (reference: each_gl with 2*max_nz global slots PER ROW and an O(nnz^2) sort
//     in global memory, :929-1027) use a column-window BITMAP + RANK scheme with accumulators in shared
//     memory: the row's columns are marked in a bitmap, a sweep turns it into ranks, the products are added
//     at acc[rank] with shared-memory atomics chunk by chunk and written out with coalesced stores; no
//     table, no compaction, no sort, no workspace (num_bitmap_kernel below);
//   * where C is many windows wide and the rows of B are short, those rows take hash passes over column ranges
//     sized by their products instead (num_hash_ranges_kernel below);
//   * native fp64 atomics everywhere (reference SpMV/SpGEMM fall back to CAS loops on fp64 in global memory).
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
                 const int *__restrict__ row_perm, const int *__restrict__ bins, const __grid_constant__ PeerOut peer)
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
        if (peer.n > 0 && r < n && t == 0) tiles_done(peer, c_rpt[rid], c_rpt[rid + 1] - c_rpt[rid]);
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
                int queue, int tmax, int sorted, int nb_max, int n_cols, const __grid_constant__ PeerOut peer)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NG = BS / GROUP;
    __shared__ FlatScratch<GROUP, real> s_flat[NG];
    __shared__ int s_row;
    __shared__ int s_fill[NG];
    __shared__ int s_maxb[NG];
    const int g = threadIdx.x / GROUP, t = threadIdx.x % GROUP;
    // values first (8-byte aligned for fp64), then keys, then the bucket counters of the output ordering
    real *vals = reinterpret_cast<real *>(smem_raw) + (size_t)g * tmax;
    int *keys = reinterpret_cast<int *>(smem_raw + sizeof(real) * (size_t)NG * tmax) + (size_t)g * tmax;
    int *cnt = reinterpret_cast<int *>(smem_raw + (sizeof(real) + sizeof(int)) * (size_t)NG * tmax) + (size_t)g * nb_max;
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
        if (t == 0) s_fill[g] = 0;
        group_sync<GROUP>();
        if (sorted) {
            // Output order by BUCKETS instead of sorting the table (round 1: bitonic sort of all tsize slots, ~100
            // barrier-separated stages at 16384 slots -- 99 ms of the 165 ms of config C5 on 8 GPUs, 10.8 ms of C2):
            // the column range is cut into nb <= tsize / 2 equal buckets, the row's keys are counted per bucket
            // (shared-memory atomics), the counts are scanned, every (key, value) goes straight to C at its bucket's
            // cursor, and the keys that share a bucket are ranked among themselves by counting.  Rows whose columns
            // cluster into one bucket (more than kMaxBucket keys) fall back to the bitonic sort.
            constexpr int kMaxBucket = 768;     // beyond this the O(bucket^2) ranking costs more than the bitonic sort
            // Buckets of at most kSmallBucket keys (columns spread evenly: B with scattered columns, configs C4 / C5)
            // are ordered WITHOUT leaving shared memory: every thread takes the keys of its <= 16 table slots into
            // registers, the key table is then free and receives, at the bucket cursors, one word per entry --
            // (column bits below the bucket, table slot of the value) -- one thread orders the few words of a bucket
            // and writes the bucket's entries to C, neighbouring threads neighbouring positions.  (Scattering the
            // entries to C and ranking them there, the path below, costs four random L2 transactions per entry:
            // 238 ms instead of the bitonic sort's 99 ms on C5's 5.3e8 entries per GPU, profiles/r2_bench_c5_gpus8*.json.)
            constexpr int kSmallBucket = 24;
            int nb = tsize >> 1;
            if (nb > nb_max) nb = nb_max;
            int shift = 0;
            while (((unsigned)(n_cols - 1) >> shift) >= (unsigned)nb) ++shift;
            for (int i = t; i < nb; i += GROUP) cnt[i] = 0;
            if (t == 0) s_maxb[g] = 0;
            group_sync<GROUP>();
            for (int i = t; i < tsize; i += GROUP) {
                const int key = keys[i];
                if (key != kEmptyKey) atomicAdd(&cnt[(unsigned)key >> shift], 1);
            }
            group_sync<GROUP>();
            // exclusive scan of the counters: thread t owns `per` consecutive buckets
            const int per = nb >= GROUP ? nb / GROUP : 1;
            const int b0 = t * per;
            int sum = 0, mx = 0;
            if (b0 < nb)
                for (int k = 0; k < per; ++k) {
                    const int c = cnt[b0 + k];
                    sum += c;
                    mx = c > mx ? c : mx;
                }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if ((t & 31) == 0 && mx > 0) atomicMax(&s_maxb[g], mx);
            const int inc = group_inclusive_scan<GROUP>(sum, t, s_flat[g].wtot);
            group_sync<GROUP>();
            const int tbits = 31 - __clz(tsize);
            const int mxb = s_maxb[g];
            if (sorted == 2 || mxb > kMaxBucket) {
                bitonic_sort_slots<GROUP, real>(keys, vals, tsize, t);
                for (int i = t; i < nnz; i += GROUP) {
                    c_col[off + i] = keys[i];
                    c_val[off + i] = vals[i];
                }
            } else if (sorted != 3 && mxb <= kSmallBucket && shift + tbits <= 32 && tsize <= 16 * GROUP) {
                int run = inc - sum;
                if (b0 < nb)
                    for (int k = 0; k < per; ++k) {
                        const int c = cnt[b0 + k];
                        cnt[b0 + k] = run;          // cursor of the bucket
                        run += c;
                    }
                int rk[16];                         // tmax / GROUP = 16 slots per thread in every class
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int i = t + j * GROUP;
                    rk[j] = i < tsize ? keys[i] : kEmptyKey;
                }
                group_sync<GROUP>();                // cursors complete, every key is in a register
                const unsigned lowmask = shift >= 32 ? 0xffffffffu : ((1u << shift) - 1u);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (rk[j] != kEmptyKey) {
                        const int pos = atomicAdd(&cnt[(unsigned)rk[j] >> shift], 1);
                        keys[pos] = (int)((((unsigned)rk[j] & lowmask) << tbits) | (unsigned)(t + j * GROUP));
                    }
                }
                group_sync<GROUP>();
                // cnt[b] is now the END of bucket b
                for (int b = t; b < nb; b += GROUP) {
                    const int s0 = b > 0 ? cnt[b - 1] : 0;
                    const int e0 = cnt[b];
                    for (int q = s0 + 1; q < e0; ++q) {
                        const unsigned w = (unsigned)keys[q];
                        int r = q - 1;
                        while (r >= s0 && (unsigned)keys[r] > w) {
                            keys[r + 1] = keys[r];
                            --r;
                        }
                        keys[r + 1] = (int)w;
                    }
                    for (int q = s0; q < e0; ++q) {
                        const unsigned w = (unsigned)keys[q];
                        c_col[off + q] = (int)(((unsigned)b << shift) | (w >> tbits));
                        c_val[off + q] = vals[w & mask];
                    }
                }
            } else {
                int run = inc - sum;
                if (b0 < nb)
                    for (int k = 0; k < per; ++k) {
                        const int c = cnt[b0 + k];
                        cnt[b0 + k] = run;          // cursor of the bucket
                        run += c;
                    }
                group_sync<GROUP>();
                for (int i = t; i < tsize; i += GROUP) {
                    const int key = keys[i];
                    if (key != kEmptyKey) {
                        const int pos = atomicAdd(&cnt[(unsigned)key >> shift], 1);
                        c_col[off + pos] = key;
                        c_val[off + pos] = vals[i];
                    }
                }
                group_sync<GROUP>();                // (the group's global writes are visible to the group)
                // cnt[b] is now the END of bucket b.  Every entry finds its place inside its bucket by counting the
                // smaller keys of the bucket (rank by counting, read back from C through L1 / L2) and goes to the -- now
                // free -- table at that position; the ordered row is then copied out with coalesced stores.
                for (int p = t; p < nnz; p += GROUP) {
                    const int key = c_col[off + p];
                    const real v = c_val[off + p];
                    const unsigned bk = (unsigned)key >> shift;
                    const int s0 = bk > 0 ? cnt[bk - 1] : 0;
                    const int e0 = cnt[bk];
                    int r = 0;
                    for (int q = s0; q < e0; ++q) r += c_col[off + q] < key;
                    keys[s0 + r] = key;
                    vals[s0 + r] = v;
                }
                group_sync<GROUP>();
                for (int i = t; i < nnz; i += GROUP) {
                    c_col[off + i] = keys[i];
                    c_val[off + i] = vals[i];
                }
            }
        } else {
            // sort = false (cuda-cpp/inc/HashSpGEMM_volta.hpp:508-605, 1018-1031): the occupied slots are compacted in
            // table order, columns of a row come out unsorted
            for (int i = t; i < tsize; i += GROUP) {
                const int key = keys[i];
                if (key != kEmptyKey) {
                    const int pos = atomicAdd(&s_fill[g], 1);
                    c_col[off + pos] = key;
                    c_val[off + pos] = vals[i];
                }
            }
        }
        group_sync<GROUP>();
        if (peer.n > 0 && t == 0) tiles_done(peer, off, nnz);
    }
}

// ---- wide C, short B rows: hash passes over COLUMN RANGES -----------------------------------------
// The rows above the hash ladder (more than 8192 entries) of a product whose C is many bitmap windows wide and whose
// B rows are short (configs C4 / C5: B with 4 entries per row, C 2^23 / 2^24 columns wide).  The bitmap kernel
// walks all entries of the row of A once per 2^19-column window -- 16 / 32 passes whatever the row holds
// (profiles/r2_ab_flat_traversal_windows.txt).  Here the number of passes follows the OUTPUT instead: one pass counts the row's
// products per column bin (2048 bins of <= 8192 columns), the bins are grouped into ranges of at most 3/4 * tmax
// products -- so a range can never overflow the table; a single bin cannot either, it has fewer columns than that --
// and every range is one hash pass over the row's products (those outside the range are skipped), ordered by the
// bucket scheme of num_hash_kernel and appended to the row.  Ranges ascend, so the row comes out sorted.
// This is the bounded table north_star asks for where the reference falls back to global memory
// (kernel_spgemm_hash_d.cu:929-1033): shared memory, bounded by construction, no retry.
constexpr int kRangeBins = 2048;

template <typename real>
__global__ void __launch_bounds__(1024, 1)
num_hash_ranges_kernel(const int *__restrict__ a_rpt, const int *__restrict__ a_col,
                       const real *__restrict__ a_val, const int *__restrict__ b_rpt,
                       const int *__restrict__ b_col, const real *__restrict__ b_val,
                       const long long *__restrict__ c_rpt, int *__restrict__ c_col, real *__restrict__ c_val,
                       const int *__restrict__ row_perm, int *__restrict__ bins, int bin_lo, int bin_hi,
                       int queue, int tmax, int nb_max, int n_cols, int bshift, const __grid_constant__ PeerOut peer)
{
    constexpr int GROUP = 1024;
    constexpr int kSmallBucket = 24;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ FlatScratch<GROUP, real> s_flat;
    __shared__ int s_row, s_maxb, s_tot;
    __shared__ int s_rng[3];
    real *vals = reinterpret_cast<real *>(smem_raw);
    int *keys = reinterpret_cast<int *>(smem_raw + sizeof(real) * (size_t)tmax);
    int *cnt = keys + tmax;
    int *hist = cnt + nb_max;              // products per column bin, then their inclusive prefix
    const int t = threadIdx.x;
    const int cap = tmax / 4 * 3;
    int lo, hi;
    class_range(bins, bin_lo, bin_hi, lo, hi);
    const int n = hi - lo;
    while (true) {
        __syncthreads();                   // (s_row of the previous row has been read by everyone)
        if (t == 0) s_row = atomicAdd(&bins[kBinQueue + queue], 1);
        __syncthreads();
        const int r = s_row;
        if (r >= n) break;
        const int rid = row_perm[lo + r];
        const long long off = c_rpt[rid];
        const int nnz = (int)(c_rpt[rid + 1] - off);
        const int a_lo = a_rpt[rid], a_hi = a_rpt[rid + 1];
        for (int i = t; i < kRangeBins; i += GROUP) hist[i] = 0;
        __syncthreads();
        for_each_product<GROUP, false, real>(t, a_lo, a_hi, a_col, a_val, b_rpt, b_col, b_val, s_flat,
                                             [&](int c, real) { atomicAdd(&hist[(unsigned)c >> bshift], 1); });
        __syncthreads();
        {
            // inclusive prefix over the bins: two consecutive bins per thread
            const int v0 = hist[2 * t], v1 = hist[2 * t + 1];
            const int inc = group_inclusive_scan<GROUP>(v0 + v1, t, s_flat.wtot);
            hist[2 * t] = inc - v1;
            hist[2 * t + 1] = inc;
        }
        __syncthreads();
        const int row_products = hist[kRangeBins - 1];
        int written = 0;
        int bin = 0;                       // first bin not yet processed (uniform)
        while (true) {
            if (t == 0) {
                const int base = bin > 0 ? hist[bin - 1] : 0;
                if (base >= row_products) {
                    s_rng[0] = -1;
                } else {
                    // first bin with products, then as many bins as fit `cap` products (at least one)
                    int l = bin, h = kRangeBins - 1;
                    while (l < h) {
                        const int m = (l + h) >> 1;
                        if (hist[m] > base) h = m; else l = m + 1;
                    }
                    const int first = l;
                    l = first;
                    h = kRangeBins;
                    while (l < h) {
                        const int m = (l + h) >> 1;
                        if (hist[m] - base > cap) h = m; else l = m + 1;
                    }
                    const int end = l > first + 1 ? l : first + 1;
                    s_rng[0] = first;
                    s_rng[1] = end;
                    s_rng[2] = hist[end - 1] - base;
                }
            }
            __syncthreads();
            const int rb0 = s_rng[0];
            if (rb0 < 0) break;
            const int rb1 = s_rng[1];
            const int prod = s_rng[2];
            bin = rb1;
            const int c_lo = rb0 << bshift;
            const long long c_hi = (long long)rb1 << bshift;
            const int width = (int)((c_hi < (long long)n_cols ? c_hi : (long long)n_cols) - c_lo);
            const int tsize = table_size_for(prod < width ? prod : width, tmax);
            const unsigned mask = (unsigned)tsize - 1u;
            for (int i = t; i < tsize; i += GROUP) {
                keys[i] = kEmptyKey;
                vals[i] = real(0);
            }
            __syncthreads();               // (also: everyone has read s_rng)
            const unsigned uw = (unsigned)width;
            for_each_product<GROUP, true, real>(t, a_lo, a_hi, a_col, a_val, b_rpt, b_col, b_val, s_flat,
                                                [&](int c, real v) {
                                                    if ((unsigned)(c - c_lo) < uw) hash_accumulate(keys, vals, mask, c, v);
                                                });
            __syncthreads();
            // order the range: buckets over [c_lo, c_lo + width), see num_hash_kernel
            int nb = tsize >> 1;
            if (nb > nb_max) nb = nb_max;
            int shift = 0;
            while (((unsigned)(width - 1) >> shift) >= (unsigned)nb) ++shift;
            for (int i = t; i < nb; i += GROUP) cnt[i] = 0;
            if (t == 0) s_maxb = 0;
            __syncthreads();
            for (int i = t; i < tsize; i += GROUP) {
                const int key = keys[i];
                if (key != kEmptyKey) atomicAdd(&cnt[(unsigned)(key - c_lo) >> shift], 1);
            }
            __syncthreads();
            const int per = nb >= GROUP ? nb / GROUP : 1;
            const int b0 = t * per;
            int sum = 0, mx = 0;
            if (b0 < nb)
                for (int k = 0; k < per; ++k) {
                    const int c = cnt[b0 + k];
                    sum += c;
                    mx = c > mx ? c : mx;
                }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if ((t & 31) == 0 && mx > 0) atomicMax(&s_maxb, mx);
            const int inc = group_inclusive_scan<GROUP>(sum, t, s_flat.wtot);
            if (t == GROUP - 1) s_tot = inc;
            __syncthreads();
            const int total = s_tot;       // entries of the range
            const int tbits = 31 - __clz(tsize);
            const long long out = off + written;
            if (s_maxb > kSmallBucket || shift + tbits > 32) {
                // clustered columns: bitonic sort of the table (free slots sort behind the keys)
                bitonic_sort_slots<GROUP, real>(keys, vals, tsize, t);
                for (int i = t; i < total; i += GROUP) {
                    c_col[out + i] = keys[i];
                    c_val[out + i] = vals[i];
                }
            } else {
                int run = inc - sum;
                if (b0 < nb)
                    for (int k = 0; k < per; ++k) {
                        const int c = cnt[b0 + k];
                        cnt[b0 + k] = run;          // cursor of the bucket
                        run += c;
                    }
                int rk[16];                         // tmax / GROUP <= 16 slots per thread
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int i = t + j * GROUP;
                    rk[j] = i < tsize ? keys[i] : kEmptyKey;
                }
                __syncthreads();                    // cursors complete, every key is in a register
                const unsigned lowmask = (1u << shift) - 1u;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (rk[j] != kEmptyKey) {
                        const unsigned rel = (unsigned)(rk[j] - c_lo);
                        const int pos = atomicAdd(&cnt[rel >> shift], 1);
                        keys[pos] = (int)(((rel & lowmask) << tbits) | (unsigned)(t + j * GROUP));
                    }
                }
                __syncthreads();
                // cnt[b] is now the END of bucket b
                for (int b = t; b < nb; b += GROUP) {
                    const int s0 = b > 0 ? cnt[b - 1] : 0;
                    const int e0 = cnt[b];
                    for (int q = s0 + 1; q < e0; ++q) {
                        const unsigned w = (unsigned)keys[q];
                        int rr = q - 1;
                        while (rr >= s0 && (unsigned)keys[rr] > w) {
                            keys[rr + 1] = keys[rr];
                            --rr;
                        }
                        keys[rr + 1] = (int)w;
                    }
                    for (int q = s0; q < e0; ++q) {
                        const unsigned w = (unsigned)keys[q];
                        c_col[out + q] = c_lo + (int)(((unsigned)b << shift) | (w >> tbits));
                        c_val[out + q] = vals[w & mask];
                    }
                }
            }
            written += total;
            __syncthreads();               // the table is cleared for the next range
        }
        if (peer.n > 0 && t == 0) tiles_done(peer, off, nnz);
    }
}

// ---- bitmap + rank + shared-memory accumulator class ---------------------------------------------
// A row of C is produced in ascending column WINDOWS of W = 2^wshift columns and, inside a window,
// in CHUNKS of at most `cap` output entries:
//   mark     every product of the window sets its column's bit (run_parts_mark: 128-bit loads, bits
//            merged per word in registers, one shared-memory atomicOr per distinct word);
//   rank     one sweep over the bitmap, a batch of 32 word pairs (2048 columns) per warp step: 16-bit
//            exclusive popcount prefix of every 32-bit word inside its batch, outputs per batch, then a
//            CTA-wide exclusive scan of the batch totals.  rank(col) = batch base + word prefix +
//            popcount of the bits below;
//   per chunk
//     columns  every warp claims batches, writes the columns of their set bits into a staging
//              buffer at rank - chunk base, and the CTA copies the buffer to C.col with coalesced stores;
//     values   acc[0..cap) (the same shared memory) is zeroed, every product whose column lies in the
//              chunk's column range is added at acc[rank - chunk base] with a shared-memory atomic
//              (measured on B200: 1.0 T adds/s fp32, 0.55 T/s fp64, against 0.22 T/s for the
//              red.global the first version used, scripts/micro/atomics_bench.cu), and acc is copied
//              to C.val with coalesced stores.
// Every product is read once per pass whatever the number of windows and chunks: the sub-range of each
// B row that falls into a column range is found by searching the sorted B row (stage_window /
// stage_chunk keep a cursor per entry; rows with more than 1024 entries search from scratch).
// Shared memory: W/8 bitmap + W/16 prefixes + cap * (sizeof(real) + 4) accumulators and columns.
constexpr int kMaxChunks = 512;
constexpr int kMaxBatches = 256;            // W <= 2^19: 2048 columns per batch

// n-th (0-based) set bit of x
__device__ __forceinline__ int select32(unsigned x, int n)
{
    for (int i = 0; i < n; ++i) x &= x - 1;
    return __ffs((int)x) - 1;
}

// kMode: 0 -- the rows of B are NOT column-sorted (the reference reader leaves the rows of symmetric files
// unsorted, nsparse.cu:115-123; checked once per call on the device): every pass walks the whole B rows and
// filters by column range; 1 -- sorted, window / chunk cuts found by searching with a cursor per entry;
// 2 -- sorted and at most 4 windows: the window cuts are precomputed per entry of A (seg, see
// stage_window_seg) and the chunk cuts of a window are searched all at once (build_chunk_table).
// kPeers: the fused allgatherv variant (multi-GPU, see PeerOut); a separate instantiation so that the
// single-GPU kernel carries none of its code (with a run-time test only, the mere presence of the peer
// loops cost the single-GPU kernel 25 % -- register allocation of the hot loops).
// kMulti: the instantiation for the rows with more than BS entries of A (red.global mode); the other one
// skips them and vice versa -- two launches over the same class, so that neither carries the other's code.
template <typename real, int BS, int kMode, bool kPeers, bool kMulti, bool kFlat>
__global__ void __launch_bounds__(BS, 1)
num_bitmap_kernel(const int *__restrict__ a_rpt, const int *__restrict__ a_col,
                  const real *__restrict__ a_val, const int *__restrict__ b_rpt,
                  const int *__restrict__ b_col, const real *__restrict__ b_val,
                  const long long *__restrict__ c_rpt, int *__restrict__ c_col, real *__restrict__ c_val,
                  const int *__restrict__ row_perm, int *__restrict__ bins, int bin_lo, int bin_hi,
                  int queue, int N, int wshift, int cap, int b_vec_end, int dbg, long long *phase_cycles,
                  const int *__restrict__ seg, long long seg_stride, const __grid_constant__ PeerOut peer)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = BS / 32;
    constexpr bool kSorted = kMode != 0;
    constexpr bool kSeg = kMode == 2;
    static_assert(NW == 32, "the batch scan and the entry search assume 32 warps");
    // development aid (build with -DNSP_PHASE_TIMING, nsp_set_option "phase_timing"): thread 0 charges the
    // cycles between barriers to twelve phases and adds them to phase_cycles[] at the end
#ifdef NSP_PHASE_TIMING
    long long ph_last = 0;
    long long ph_acc[12];
    if (phase_cycles) {
#pragma unroll
        for (int i = 0; i < 12; ++i) ph_acc[i] = 0;
        ph_last = clock64();
    }
#define PH(i)                                              \
    if (phase_cycles && threadIdx.x == 0) {                \
        const long long now = clock64();                   \
        ph_acc[i] += now - ph_last;                        \
        ph_last = now;                                     \
    }
#else
#define PH(i)
#endif
    __shared__ PartScratch<BS, real> s_part;
    __shared__ int s_batch[kMaxBatches + 1];    // outputs before the batch (exclusive), window relative
    __shared__ int s_bound[kMaxChunks + 1];
    const unsigned W = 1u << wshift;
    const int nbatch = (int)(W >> 11);          // a batch = 32 word pairs = 2048 columns
    uint2 *bm64 = reinterpret_cast<uint2 *>(smem_raw);
    unsigned *bm32 = reinterpret_cast<unsigned *>(smem_raw);
    unsigned *pre2 = reinterpret_cast<unsigned *>(smem_raw + (W >> 3));            // per pair: two 16-bit prefixes
    const unsigned short *pre16 = reinterpret_cast<const unsigned short *>(pre2);  // the same per 32-bit word
    real *acc = reinterpret_cast<real *>(smem_raw + (W >> 3) + (W >> 4));
    int *cols = reinterpret_cast<int *>(acc + cap);
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    int lo, hi;
    class_range(bins, bin_lo, bin_hi, lo, hi);
    const int n = hi - lo;
    const int nwin = (int)(((unsigned)N + W - 1u) >> wshift);
    // Row queue with a look-ahead of one row: while the CTA works on a row, thread 0 claims the next one and walks
    // the dependent chain queue -> row_perm -> A.rpt / C.rpt one step per phase (every step consumes a value whose
    // load was issued a phase earlier), so that a row starts from shared memory instead of four global round trips.
    __shared__ int s_nrid, s_nab, s_nae;
    __shared__ long long s_nout;
    int nx_rid = -1, nx_ab = 0, nx_ae = 0;
    long long nx_out = 0;
    if (t == 0) {
        const int r = atomicAdd(&bins[kBinQueue + queue], 1);
        if (r < n) {
            nx_rid = row_perm[lo + r];
            nx_ab = a_rpt[nx_rid];
            nx_ae = a_rpt[nx_rid + 1];
            nx_out = c_rpt[nx_rid];
        }
        s_nrid = nx_rid;
        s_nab = nx_ab;
        s_nae = nx_ae;
        s_nout = nx_out;
    }
    while (true) {
        __syncthreads();
        const int rid = s_nrid;
        if (rid < 0) break;
        const int a_beg = s_nab, a_end = s_nae;
        const int E = a_end - a_beg;
        long long out = s_nout;
        __syncthreads();                           // everyone has read the slot before thread 0 refills it
        int nx_r = 0;
        if ((E > BS) != kMulti) {                  // the other launch's row (CTA uniform): claim the next one at once
            if (t == 0) {
                nx_r = atomicAdd(&bins[kBinQueue + queue], 1);
                nx_rid = -1;
                if (nx_r < n) {
                    nx_rid = row_perm[lo + nx_r];
                    s_nab = a_rpt[nx_rid];
                    s_nae = a_rpt[nx_rid + 1];
                    s_nout = c_rpt[nx_rid];
                }
                s_nrid = nx_rid;
            }
            continue;
        }
        if (t == 0) nx_r = atomicAdd(&bins[kBinQueue + queue], 1);      // step A: claim
        constexpr bool one_slab = !kMulti;
        const int glog = entry_group_log(E, BS);
        for (int win = 0; win < nwin; ++win) {
            const int c0 = (int)((unsigned)win << wshift);
            const int c1 = (int)min((unsigned)N, (unsigned)c0 + W);
            const bool cut_lo = kSorted && win > 0, cut_hi = kSorted && win < nwin - 1;
            {
                uint4 *bm4 = reinterpret_cast<uint4 *>(smem_raw);
                for (int i = t; i < (int)(W >> 7); i += BS) bm4[i] = make_uint4(0u, 0u, 0u, 0u);
            }
            __syncthreads();
            PH(0);
            // ---- mark ----
            int staged_total = 0;
            const unsigned ncols = (unsigned)(c1 - c0);
            auto mark_one = [&](int c, real) {                 // flat traversal: one bit per product
                const unsigned cc = (unsigned)(c - c0);
                atomicOr(bm32 + bitmap_word32(cc >> 5), 1u << (cc & 31u));
            };
            if (one_slab) {
                if (kSeg)
                    staged_total = stage_window_seg<BS, true, real, kFlat>(t, a_beg, E, a_val, seg, seg_stride, win, win == 0, s_part);
                else
                    staged_total = stage_window<BS, true, real, kFlat>(t, a_beg, E, glog, a_col, a_val, b_rpt, b_col, c1,
                                                                win == 0 || !kSorted, cut_hi, s_part);
                PH(1);
                if (kFlat)
                    run_flat<BS, false, real>(t, staged_total, b_col, b_val, s_part, mark_one);
                else
                    run_parts_mark<BS, !kSorted, real>(t, staged_total, b_col, b_vec_end, s_part, bm32, c0, ncols);
                PH(2);
            } else {
                for (int base = a_beg; base < a_end; base += BS) {
                    const int total = kSeg ? stage_slab_seg<BS, false, real, kFlat>(t, base, a_end, a_val, seg, seg_stride, win, s_part)
                                           : stage_parts_range<BS, false, real, kFlat>(t, base, a_end, a_col, a_val, b_rpt, b_col, c0,
                                                                                c1, cut_lo, cut_hi, s_part);
                    if (kFlat)
                        run_flat<BS, false, real>(t, total, b_col, b_val, s_part, mark_one);
                    else
                        run_parts_mark<BS, !kSorted, real>(t, total, b_col, b_vec_end, s_part, bm32, c0, ncols);
                    PH(2);
                }
            }
            if (t == 0 && win == 0) nx_rid = nx_r < n ? row_perm[lo + nx_r] : -1;          // step B: which row
            // ---- rank: per-word prefixes inside every batch (four batches in flight per warp: the scan is
            //      a chain of five dependent shuffles), batch totals, exclusive scan of the totals by warp 0 ----
            for (int b0 = wid * 4; b0 < nbatch; b0 += NW * 4) {
                unsigned pj[4];
                int cl[4], c[4], inc[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    pj[u] = (unsigned)((b0 + u) << 5) | ((unsigned)lane ^ bitmap_swz((unsigned)(b0 + u)));
                    const uint2 wd = bm64[pj[u]];
                    cl[u] = __popc(wd.x);
                    c[u] = cl[u] + __popc(wd.y);
                    inc[u] = c[u];
                }
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int v = __shfl_up_sync(0xffffffffu, inc[u], o);
                        if (lane >= o) inc[u] += v;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const unsigned ex = (unsigned)(inc[u] - c[u]);
                    pre2[pj[u]] = ex | ((ex + (unsigned)cl[u]) << 16);
                    if (lane == 31) s_batch[b0 + u] = inc[u];
                }
            }
            __syncthreads();
            PH(3);
            if (t == 0 && win == 0 && nx_rid >= 0) {                                     // step C: its extent
                nx_ab = a_rpt[nx_rid];
                nx_ae = a_rpt[nx_rid + 1];
                nx_out = c_rpt[nx_rid];
            }
            if (wid == 0) {
                // nbatch <= 256: lane l owns the batches 8l .. 8l+7 (nbatch is a multiple of 32)
                const int per = nbatch >> 5;
                int v[8], sum = 0;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    v[u] = u < per ? s_batch[lane * per + u] : 0;
                    sum += v[u];
                }
                int inc = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int x = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += x;
                }
                int run = inc - sum;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (u < per) s_batch[lane * per + u] = run;
                    run += v[u];
                }
                if (lane == 31) s_batch[nbatch] = inc;
            }
            __syncthreads();
            const int tile_nnz = s_batch[nbatch];
            const int nch = (tile_nnz + cap - 1) / cap;
            // chunk boundary columns: chunk k starts at the column of rank k * cap (one warp per boundary)
            for (int k = 1 + wid; k < nch; k += NW) {
                const int R = k * cap;
                // batch: largest b with s_batch[b] <= R  (nbatch <= 256: 8 entries per lane)
                int cntb = 0;
                for (int i = lane; i < nbatch; i += 32) cntb += s_batch[i] <= R;
                const int b = warp_sum(cntb) - 1;
                const unsigned pj = (unsigned)(b << 5) | ((unsigned)lane ^ bitmap_swz((unsigned)b));
                const uint2 wd = bm64[pj];
                const unsigned pp = pre2[pj];
                const int rb = R - s_batch[b];                      // rank inside the batch
                const int p_lo = (int)(pp & 0xffffu), p_hi = (int)(pp >> 16);
                const int p_end = p_hi + __popc(wd.y);
                if (rb >= p_lo && rb < p_end) {                     // exactly one lane
                    const int colw = rb < p_hi ? select32(wd.x, rb - p_lo) : 32 + select32(wd.y, rb - p_hi);
                    s_bound[k] = c0 + (((b << 5) + lane) << 6) + colw;
                }
            }
            __syncthreads();
            // segment mode: every chunk cut of the window in one round of searches
            const bool tab_ok = kSeg && one_slab && nch > 1 && E * (nch + 1) <= 2 * BS;
            if (tab_ok) build_chunk_table<BS, real>(t, E, nch, b_col, s_bound, s_part);
            PH(4);
            // ---- chunk by chunk ----
            // Rows whose A entries fit one slab: acc[0..cnt) is zeroed, every product of the chunk's column
            // range is added at acc[rank - chunk base] with a shared-memory atomic and stores its column at
            // cols[rank - chunk base] (a plain store: every writer of a slot writes the same value), then both
            // arrays go to C.val / C.col with coalesced stores.  Emitting the columns here costs two
            // instructions per product with all lanes busy; walking the set bits of the bitmap instead cost
            // 24 % of the kernel's instructions at 3-4 active lanes (profiles/r1_ncu_bitmap_s20_v4.txt).
            for (int k = 0; k < nch && one_slab && !(dbg & 2); ++k) {
                const int r0 = k * cap;
                const int cnt = min(tile_nnz - r0, cap);
                const int col_lo = k > 0 ? s_bound[k] : c0;
                const int col_hi = k < nch - 1 ? s_bound[k + 1] : c1;
                for (int i = t; i < ((cnt + 31) & ~31); i += BS) acc[i] = real(0);   // whole swizzle groups
                // (the barriers of the staging below order the zeroes before the adds)
                auto add = [&](int c, real v) {
                    if (!kSorted && (unsigned)(c - col_lo) >= (unsigned)(col_hi - col_lo)) return;
                    const unsigned cc = (unsigned)(c - c0);
                    const unsigned w32 = cc >> 5;
                    const unsigned p32 = bitmap_word32(w32);
                    const int rank = s_batch[w32 >> 6] + (int)pre16[p32] +
                                     __popc(bm32[p32] & ((1u << (cc & 31u)) - 1u));
                    const int idx = acc_swz(rank - r0);
                    cols[idx] = c;
                    atomicAdd(acc + idx, v);
                };
                int total = staged_total;                  // one chunk (or unsorted B): the mark pass staged it
                if (kSeg && nch > 1)
                    total = tab_ok ? stage_chunk_tab<BS, real, kFlat>(t, E, k, s_part)
                                   : stage_chunk_seg<BS, real, kFlat>(t, a_beg, E, glog, b_col, seg, seg_stride, win, col_lo, col_hi,
                                                               k == 0, k == nch - 1, s_part);
                else if (kSorted && nch > 1)
                    total = stage_chunk<BS, real, kFlat>(t, E, glog, b_col, col_hi, k == nch - 1, s_part);
                else
                    __syncthreads();
                PH(7);
                // (a variant that issued the rank lookups, accumulator reads and compare-and-swaps of four products
                // side by side instead of one LDS / FADD / ATOMS.CAST.SPIN chain per product measured 8 % SLOWER:
                // profiles/r2_ab_value_pass_batched_cas_vs_spin_s20.txt)
                if (kFlat)
                    run_flat<BS, true, real>(t, total, b_col, b_val, s_part, add);
                else
                    run_parts<BS, true, real>(t, total, b_col, b_val, s_part, add);
                PH(8);
                real *cv = c_val + out + r0;
                int *cc = c_col + out + r0;
                for (int i = t; i < cnt; i += BS) {
                    const int j = acc_swz(i);
                    cv[i] = acc[j];
                    cc[i] = cols[j];
                }
                __syncthreads();                       // the next chunk reuses the buffers
                // multi-GPU: the chunk is final in the local C; count it into its tiles (peer_push.cu sends them)
                if (kPeers && t == 0) tiles_done(peer, out + r0, cnt);
                PH(9);
            }
            // Rows with more than BS entries of A (0.1 % of the rows, a seventh of the products on R-MAT)
            // would re-search every slab for every chunk: they emit their columns from the bitmap, chunk by
            // chunk through the cols buffer, and add their values straight into C.val with red.global in
            // ONE pass over the products.
            if (!one_slab) {
                for (int k = 0; k < nch && !(dbg & 1); ++k) {
                    const int r0 = k * cap;
                    const int cnt = min(tile_nnz - r0, cap);
                    for (int step = 0; step < (nbatch >> 5); ++step) {
                        const int b = (step << 5) | ((wid + 11 * step) & 31);
                        const int base = s_batch[b] - r0;
                        const int bend = s_batch[b + 1] - r0;
                        if (bend <= 0 || base >= cnt || bend == base) continue;      // no output of this chunk
                        const unsigned pj = (unsigned)(b << 5) | ((unsigned)lane ^ bitmap_swz((unsigned)b));
                        uint2 wd = bm64[pj];
                        int pos = base + (int)(pre2[pj] & 0xffffu);
                        const int cbase = c0 + (((b << 5) + lane) << 6);
                        while (wd.x) {
                            if ((unsigned)pos < (unsigned)cnt) cols[pos] = cbase + __ffs((int)wd.x) - 1;
                            ++pos;
                            wd.x &= wd.x - 1;
                        }
                        while (wd.y) {
                            if ((unsigned)pos < (unsigned)cnt) cols[pos] = cbase + 31 + __ffs((int)wd.y);
                            ++pos;
                            wd.y &= wd.y - 1;
                        }
                    }
                    __syncthreads();
                    int *cc = c_col + out + r0;
                    for (int i = t; i < cnt; i += BS) cc[i] = cols[i];
                    __syncthreads();
                }
                PH(6);
            }
            if (!one_slab && !(dbg & 2)) {
                real *cv = c_val + out;
                for (int i = t; i < tile_nnz; i += BS) cv[i] = real(0);
                auto add_red = [&](int c, real v) {
                    const unsigned cc = (unsigned)(c - c0);
                    if (!kSorted && cc >= ncols) return;
                    const unsigned w32 = cc >> 5;
                    const unsigned p32 = bitmap_word32(w32);
                    const int rank = s_batch[w32 >> 6] + (int)pre16[p32] +
                                     __popc(bm32[p32] & ((1u << (cc & 31u)) - 1u));
                    atomicAdd(cv + rank, v);
                };
                for (int base = a_beg; base < a_end; base += BS) {
                    // (the barriers of the staging order the zero fill before the adds)
                    const int total = kSeg ? stage_slab_seg<BS, true, real, kFlat>(t, base, a_end, a_val, seg, seg_stride, win, s_part)
                                           : stage_parts_range<BS, true, real, kFlat>(t, base, a_end, a_col, a_val, b_rpt, b_col, c0, c1,
                                                                               cut_lo, cut_hi, s_part);
                    if (kFlat)
                        run_flat<BS, true, real>(t, total, b_col, b_val, s_part, add_red);
                    else
                        run_parts<BS, true, real>(t, total, b_col, b_val, s_part, add_red);
                    PH(10);
                }
                if (kPeers) {
                    // the window's entries are final once every thread's reds are performed
                    __threadfence();
                    __syncthreads();
                    if (t == 0) tiles_done(peer, out, tile_nnz);
                }
            }
            out += tile_nnz;
        }
        if (t == 0) {                                                                     // step D: hand over
            s_nrid = nx_rid;
            s_nab = nx_ab;
            s_nae = nx_ae;
            s_nout = nx_out;
        }
    }
#ifdef NSP_PHASE_TIMING
    if (phase_cycles && threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < 12; ++i) atomicAdd((unsigned long long *)phase_cycles + i, (unsigned long long)ph_acc[i]);
    }
#endif
#undef PH
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

// see run_parts_mark: last index from which B.col may be read with a 128-bit load (0: never)
static inline int b_vec_end_of(const nsp_context *ctx, const int *b_col)
{
    if ((reinterpret_cast<uintptr_t>(b_col) & 15u) != 0 || ctx->opt_no_vec) return 0;
    return (int)(ctx->sp.b_nnz & ~3ll);
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
static int launch_num_hash(nsp_context *ctx, const char *name, int grid, const int *a_rpt,
                           const int *a_col, const real *a_val, const int *b_rpt, const int *b_col,
                           const real *b_val, const long long *c_rpt64, int *c_col, real *c_val,
                           int bin_lo, int bin_hi, int queue, int tmax, int n_cols)
{
    nsp_spgemm_state &sp = ctx->sp;
    auto kern = num_hash_kernel<real, GROUP, BS>;
    constexpr int NG = BS / GROUP;
    // table (values + keys) of every group, plus tmax / 2 bucket counters per group where the shared memory allows
    // (the 1024-thread class in fp64 fills it: tmax / 4 there)
    const size_t table = (size_t)tmax * (sizeof(real) + sizeof(int)) * NG;
    const size_t limit = (size_t)ctx->max_smem_optin - (sizeof(FlatScratch<GROUP, real>) * NG + 256);
    int nb_max = tmax / 2;
    while (nb_max > 16 && table + (size_t)nb_max * sizeof(int) * NG > limit) nb_max >>= 1;
    const size_t smem = table + (size_t)nb_max * sizeof(int) * NG;
    NSP_CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // 0: unsorted rows; 1: buckets (in shared memory where they are small, through C otherwise, bitonic sort when the
    // columns cluster); 2 / 3 (option "hash_order" 1 / 2, measurements): bitonic sort always / no shared-memory buckets
    const int order = ctx->opt_unsorted ? 0 : (ctx->opt_hash_order == 1 ? 2 : (ctx->opt_hash_order == 2 ? 3 : 1));
    num_prof_class(ctx, name, bin_lo, bin_hi);
    kern<<<grid, BS, smem, ctx->stream>>>(NSP_NUM_ARGS, bin_lo, bin_hi, queue, tmax, order, nb_max, n_cols,
                                          ctx->peer_out);
    ctx->prof_end();
    ctx->launches += 1;
    NSP_CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}

// CUDA loads kernels lazily, on their first launch, and the load may need a context-wide synchronisation: it then
// waits for every running kernel -- also for the persistent pusher kernel of the multi-GPU path, which itself waits
// for the kernels being loaded.  Every function the numeric phase can launch while the pusher runs is therefore
// loaded BEFORE the pusher starts (cudaFuncGetAttributes forces the load); once per context and precision.
template <typename real>
static int preload_numeric_kernels(nsp_context *ctx)
{
    const int which = sizeof(real) == 8 ? 1 : 0;
    if (ctx->peers_preloaded[which]) return 0;
    cudaFuncAttributes at;
    NSP_CUDA_TRY(ctx, cudaFuncGetAttributes(&at, num_pwarp_kernel<real>));
    NSP_CUDA_TRY(ctx, cudaFuncGetAttributes(&at, num_hash_kernel<real, 32, 256>));
    NSP_CUDA_TRY(ctx, cudaFuncGetAttributes(&at, num_hash_kernel<real, 256, 256>));
    NSP_CUDA_TRY(ctx, cudaFuncGetAttributes(&at, num_hash_kernel<real, 1024, 1024>));
    NSP_CUDA_TRY(ctx, cudaFuncGetAttributes(&at, num_hash_ranges_kernel<real>));
    NSP_CUDA_TRY(ctx, cudaFuncGetAttributes(&at, num_bitmap_kernel<real, 1024, 0, true, false, false>));
    NSP_CUDA_TRY(ctx, cudaFuncGetAttributes(&at, num_bitmap_kernel<real, 1024, 1, true, false, false>));
    NSP_CUDA_TRY(ctx, cudaFuncGetAttributes(&at, num_bitmap_kernel<real, 1024, 2, true, false, false>));
    NSP_CUDA_TRY(ctx, cudaFuncGetAttributes(&at, num_bitmap_kernel<real, 1024, 0, true, true, false>));
    NSP_CUDA_TRY(ctx, cudaFuncGetAttributes(&at, num_bitmap_kernel<real, 1024, 1, true, true, false>));
    NSP_CUDA_TRY(ctx, cudaFuncGetAttributes(&at, num_bitmap_kernel<real, 1024, 2, true, true, false>));
    NSP_CUDA_TRY(ctx, cudaFuncGetAttributes(&at, num_bitmap_kernel<real, 1024, 1, true, false, true>));
    NSP_CUDA_TRY(ctx, cudaFuncGetAttributes(&at, num_bitmap_kernel<real, 1024, 2, true, false, true>));
    NSP_CUDA_TRY(ctx, cudaFuncGetAttributes(&at, num_bitmap_kernel<real, 1024, 1, true, true, true>));
    NSP_CUDA_TRY(ctx, cudaFuncGetAttributes(&at, num_bitmap_kernel<real, 1024, 2, true, true, true>));
    ctx->peers_preloaded[which] = true;
    return 0;
}

template <typename real>
int spgemm_numeric_reserve(nsp_context *ctx, int N, long long a_nnz, long long nnz_block, int rows, int npeers)
{
    if (preload_numeric_kernels<real>(ctx) != 0) return -1;
    int ws_max = ctx->opt_num_window_shift > 0 ? (int)ctx->opt_num_window_shift : 19;
    ws_max = ws_max < 16 ? 16 : (ws_max > 19 ? 19 : ws_max);
    int wshift = 16;
    while (wshift < ws_max && (1ll << wshift) < (long long)N) ++wshift;
    const long long nwin = ((long long)N + (1ll << wshift) - 1) >> wshift;
    if (nwin <= 4 && !ctx->opt_no_seg && reserve_entry_segments(ctx, a_nnz, (int)nwin) != 0) return -1;
    if (ctx->opt_gather_tma) return peer_push_reserve(ctx, (nnz_block >> kTileLog) + 2);
    return peer_dma_reserve(ctx, (nnz_block >> dma_tile_log(nnz_block)) + 2, npeers, rows);
}

template <typename real>
int spgemm_numeric(nsp_context *ctx, int M, int K, int N, const int *a_rpt, const int *a_col,
                   const real *a_val, const int *b_rpt, const int *b_col, const real *b_val,
                   const long long *c_rpt64, int *c_col, real *c_val, int row0, int nrows)
{
    nsp_spgemm_state &sp = ctx->sp;
    ctx->dma.t_numeric = std::chrono::steady_clock::now();
    ctx->dma.last_kernel_ms = 0;
    if (!sp.symbolic_done || sp.M != M || sp.K != K || sp.N != N)
        return ctx->fail(-2, "nsp_spgemm_numeric: call nsp_spgemm_symbolic on the same context and shapes first");
    if (nrows < 0) nrows = M - row0;
    if (row0 < 0 || nrows < 0 || row0 + nrows > M) return ctx->fail(-2, "nsp_spgemm_numeric: bad row range");
    if (ctx->peer_out.n > 0 && (row0 != 0 || nrows != M))
        return ctx->fail(-2, "nsp_spgemm_numeric_rows: not available while peers are set (nsp_spgemm_set_peers)");
    // Rows [row0, row0 + nrows) only (the multi-GPU pipeline computes a block in pieces so that finished
    // pieces travel to the peers while the next one is computed): every per-row array is entered at row0,
    // row ids are then relative to it, and the row pointers keep indexing the full col / val arrays.
    a_rpt += row0;
    c_rpt64 += row0;
    M = nrows;
    if (M == 0) return 0;
    // Segment mode of the heavy class (spgemm_device.cuh stage_window_seg): B column-sorted, at most 4 windows,
    // whole-matrix call (the cuts are indexed like A.col, which needs A.rpt[0] == 0).  Built BEFORE the plan's host
    // sync, so that the side-stream launch of the long rows below does not wait behind it and still takes its SMs
    // ahead of the main launch.
    int seg_wshift = 16;
    {
        int ws_max = ctx->opt_num_window_shift > 0 ? (int)ctx->opt_num_window_shift : 19;
        ws_max = ws_max < 16 ? 16 : (ws_max > 19 ? 19 : ws_max);
        while (seg_wshift < ws_max && (1ll << seg_wshift) < (long long)N) ++seg_wshift;
    }
    const long long seg_nwin = ((long long)N + (1ll << seg_wshift) - 1) >> seg_wshift;
    const bool use_seg = sp.b_sorted && seg_nwin <= 4 && row0 == 0 && nrows == sp.M && !ctx->opt_no_seg && sp.a_nnz > 0;
    if (use_seg && build_entry_segments(ctx, a_col, sp.a_nnz, b_rpt, b_col, (int)seg_nwin, seg_wshift) != 0) return -1;
    if (plan_by_count(ctx, M, kNumShift, a_rpt, row0) != 0) return -1;

    // ---- class ladder (numeric shift 4: bin b holds 2^(3+b) < nnz <= 2^(4+b)) ----
    //   bin 0          <= 16        4 threads / row, 32 slots
    //   bins 1..4      <= 256       warp / row, <= 512 slots, 8 rows per CTA
    //   bins 5..7      <= 2048      CTA(256) / row, <= 4096 slots
    //   bins 8..9      <= 8192      CTA(1024) / row, <= 16384 slots (128 KiB fp32 / 192 KiB fp64)
    //   bins >= bm_bin              CTA(1024) / row, bitmap + rank over column tiles
    const int smem_cap = ctx->max_smem_optin - kStaticSmemReserve;
    // window: power of two >= N, at least 2^16 (one prefix segment of >= 32 words per warp), at most
    // 2^ws_max; 3/16 byte per column for bitmap + prefixes, the rest of the CTA's shared memory holds
    // the accumulators
    int ws_max = ctx->opt_num_window_shift > 0 ? (int)ctx->opt_num_window_shift : 19;
    if (ws_max < 16) ws_max = 16;
    if (ws_max > 19) ws_max = 19;
    int wshift = 16;
    while (wshift < ws_max && (1ll << wshift) < (long long)N) ++wshift;
    const size_t fixed = ((size_t)3 << wshift) / 16;
    int cap = (int)((smem_cap - (long long)fixed) / (long long)(sizeof(real) + sizeof(int)));
    if (ctx->opt_num_cap > 0 && ctx->opt_num_cap < cap) cap = (int)ctx->opt_num_cap;   // tests: force many chunks
    cap &= ~127;
    // Class boundary.  A bitmap row costs a sweep per window whatever its size (~10 us), a hash row a
    // bitonic sort of its table (n log^2 n): measured on R-MAT scale 20 (two windows) the bitmap wins above
    // ~2048 entries.  With many windows (very wide C) the hash ladder keeps everything it can hold.
    const long long nwin_host = ((long long)N + (1ll << wshift) - 1) >> wshift;
    int bm_bin = nwin_host <= 4 ? 8 : 10;
    if (ctx->opt_num_bitmap_min >= 0) {
        bm_bin = log_bin(num_imin(ctx->opt_num_bitmap_min, 0x7fffffff), kNumShift) + 1;
        if (bm_bin < 1) bm_bin = 1;
        if (bm_bin > 10) bm_bin = 10;
    }
    int bm_lo = bm_bin;      // first bin of the bitmap kernels (above bm_bin when num_hash_ranges_kernel takes the lower bins)
    const int sms = ctx->sm_count;
    // the side-stream launch of the long rows is joined on EVERY exit path, also the early error returns of the
    // launches below (the aux kernel reads d_bins / d_row_perm, which the next call rewrites)
    struct JoinGuard {
        nsp_context *ctx;
        ~JoinGuard()
        {
            if (ctx->sp.join_pending) {
                cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0);
                ctx->sp.join_pending = false;
            }
        }
    } join_guard{ctx};
    const bool peers = ctx->peer_out.n > 0;
    long long nnz_block = 0;
    for (int b = 0; b < kNumBins; ++b) nnz_block += (long long)sp.h_binsum[kSumCnt + b];
    // multi-GPU: the pusher kernel takes its SMs first (peer_push.cu); the computing kernels only count tiles
    const bool tma = peers && ctx->opt_gather_tma;
    if (tma) {
        if (preload_numeric_kernels<real>(ctx) != 0) return -1;
        if (peer_push_begin(ctx, c_col - ctx->peer_out.off, c_val - ctx->peer_out.off, (int)sizeof(real), nnz_block) != 0) return -1;
    } else if (peers) {
        // copy-engine gather: big tiles, and the heavy rows in tile order so that tiles finish one after another
        if (peer_dma_begin(ctx, nnz_block) != 0) return -1;
        if (order_rows_by_tile(ctx, sp.d_row_perm, num_imin(num_rows_in(sp, bm_bin, kNumBins - 1), 0x7fffffff), c_rpt64,
                               ctx->peer_out.off, ctx->peer_out.tile_log) != 0)
            return -1;
    }
    const int push_sms = ctx->push_active ? ctx->push_ctas : 0;

    auto launch_bitmap = [&](int multi) -> int {
        if (num_rows_in(sp, bm_lo, kNumBins - 1) == 0) return 0;
        if (multi && !sp.has_multi_slab) return 0;
        if (cap < 128 || ((1ll << wshift) + cap - 1) / cap > kMaxChunks)
            return ctx->fail(-4, "nsp_spgemm_numeric: shared memory too small for the bitmap kernel");
        const size_t smem = fixed + (size_t)cap * (sizeof(real) + sizeof(int));
        const int grid = num_imin(num_rows_in(sp, bm_lo, kNumBins - 1), (long long)(sms - push_sms));
        const long long a_entries = sp.a_nnz;
        const int mode = !sp.b_sorted ? 0 : (use_seg ? 2 : 1);
        // flat traversal when the B rows of the class are short (spgemm_device.cuh run_flat): products per entry of A
        // below 48 on average, and at most 4 windows (every pass touches every product: with the 16 / 32 windows of the
        // full-size configs C4 / C5 it was 1.6-2.2x slower than the searched sub-ranges,
        // profiles/r2_ab_flat_traversal_windows.txt)
        long long cls_ip = 0, cls_len = 0;
        for (int b = bm_lo; b < kNumBins; ++b) {
            cls_ip += (long long)sp.h_binsum[kSumIp + b];
            cls_len += (long long)sp.h_binsum[kSumLen + b];
        }
        const int flat = (mode != 0 && cls_len > 0 && cls_ip < 48 * cls_len && nwin_host <= 4 && !ctx->opt_no_flat) || (mode != 0 && ctx->opt_no_flat < 0);
        // [multi][peers][mode][flat]
        using kern_t = void (*)(const int *, const int *, const real *, const int *, const int *, const real *, const long long *,
                                int *, real *, const int *, int *, int, int, int, int, int, int, int, int, long long *,
                                const int *, long long, const PeerOut);
#define NSP_BM(mode_, peers_, multi_, flat_) num_bitmap_kernel<real, 1024, mode_, peers_, multi_, flat_>
        static const kern_t kerns[2][2][3][2] = {
            {{{NSP_BM(0, false, false, false), NSP_BM(0, false, false, false)},
              {NSP_BM(1, false, false, false), NSP_BM(1, false, false, true)},
              {NSP_BM(2, false, false, false), NSP_BM(2, false, false, true)}},
             {{NSP_BM(0, true, false, false), NSP_BM(0, true, false, false)},
              {NSP_BM(1, true, false, false), NSP_BM(1, true, false, true)},
              {NSP_BM(2, true, false, false), NSP_BM(2, true, false, true)}}},
            {{{NSP_BM(0, false, true, false), NSP_BM(0, false, true, false)},
              {NSP_BM(1, false, true, false), NSP_BM(1, false, true, true)},
              {NSP_BM(2, false, true, false), NSP_BM(2, false, true, true)}},
             {{NSP_BM(0, true, true, false), NSP_BM(0, true, true, false)},
              {NSP_BM(1, true, true, false), NSP_BM(1, true, true, true)},
              {NSP_BM(2, true, true, false), NSP_BM(2, true, true, true)}}}};
#undef NSP_BM
        // The long rows (few, each with millions of products: a tail of a handful of CTAs) go to a side stream
        // and start first; as their CTAs retire, the SMs pick up the CTAs of the main launch, whose dynamic row
        // queue balances whatever number of them is running.  Joined at the end of the phase.
        auto kern = kerns[multi][peers ? 1 : 0][mode][flat];
        NSP_CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaStream_t st = ctx->stream;
        if (multi && !ctx->opt_no_fork) {
            NSP_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
            NSP_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
            st = ctx->aux_stream;
        }
        ctx->prof_on_aux = st != ctx->stream;
        num_prof_class(ctx, multi ? "num_bitmap_long" : "num_bitmap", bm_lo, kNumBins - 1);
        kern<<<grid, 1024, smem, st>>>(NSP_NUM_ARGS, bm_lo, kNumBins - 1, multi ? 5 : 4, N, wshift, cap,
                                       b_vec_end_of(ctx, b_col), (int)ctx->opt_debug,
                                       ctx->opt_phase_timing ? ctx->phase_cycles() : nullptr, use_seg ? ctx->d_seg : nullptr,
                                       a_entries, ctx->peer_out);
        ctx->prof_end();
        ctx->prof_on_aux = false;
        ctx->launches += 1;
        NSP_CUDA_TRY(ctx, cudaGetLastError());
        if (st != ctx->stream) {
            NSP_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_join, ctx->aux_stream));
            sp.join_pending = true;
        }
        return 0;
    };
    auto launch_light = [&]() -> int {
        if (bm_bin > 8 && num_rows_in(sp, 8, num_imin(9, bm_bin - 1)) > 0) {
            const int hi = num_imin(9, bm_bin - 1);
            const int tmax = 16384;
            const int grid = num_imin(num_rows_in(sp, 8, hi), sms - push_sms);
            if (launch_num_hash<real, 1024, 1024>(ctx, "num_hash_cta1024", grid, a_rpt, a_col, a_val, b_rpt, b_col, b_val,
                                                  c_rpt64, c_col, c_val, 8, hi, 3, tmax, N) != 0)
                return -1;
        }
        if (bm_bin > 5 && num_rows_in(sp, 5, num_imin(7, bm_bin - 1)) > 0) {
            const int hi = num_imin(7, bm_bin - 1);
            const int tmax = 4096;
            const int grid = num_imin(num_rows_in(sp, 5, hi), (long long)sms * 4);
            if (launch_num_hash<real, 256, 256>(ctx, "num_hash_cta256", grid, a_rpt, a_col, a_val, b_rpt, b_col, b_val,
                                                c_rpt64, c_col, c_val, 5, hi, 2, tmax, N) != 0)
                return -1;
        }
        if (num_rows_in(sp, 1, num_imin(4, bm_bin - 1)) > 0) {
            const int hi = num_imin(4, bm_bin - 1);
            const int tmax = 512;
            const int grid = num_imin((num_rows_in(sp, 1, hi) + 7) / 8, (long long)sms * 4);
            if (launch_num_hash<real, 32, 256>(ctx, "num_hash_warp", grid, a_rpt, a_col, a_val, b_rpt, b_col, b_val, c_rpt64,
                                               c_col, c_val, 1, hi, 1, tmax, N) != 0)
                return -1;
        }
        if (num_rows_in(sp, 0, 0) > 0) {
            const int grid = num_imin((num_rows_in(sp, 0, 0) + 63) / 64, (long long)sms * 8);
            num_prof_class(ctx, "num_pwarp", 0, 0);
            num_pwarp_kernel<real><<<grid, 256, 0, ctx->stream>>>(a_rpt, a_col, a_val, b_rpt, b_col, b_val, c_rpt64, c_col,
                                                                  c_val, sp.d_row_perm, sp.d_bins, ctx->peer_out);
            ctx->prof_end();
            ctx->launches += 1;
            NSP_CUDA_TRY(ctx, cudaGetLastError());
        }
        return 0;
    };
    // Wide C with short B rows (C4 / C5 at full size): the rows above the hash ladder take hash passes over column
    // ranges (num_hash_ranges_kernel) instead of the bitmap windows.  Needs at most 8192 columns per bin.
    int rbshift = 0;
    while (((long long)N - 1) >> rbshift >= (long long)kRangeBins) ++rbshift;
    bool use_ranges = false;
    {
        long long cls_ip = 0, cls_len = 0;
        for (int b = bm_bin; b < kNumBins; ++b) {
            cls_ip += (long long)sp.h_binsum[kSumIp + b];
            cls_len += (long long)sp.h_binsum[kSumLen + b];
        }
        use_ranges = rbshift <= 13 && ((nwin_host > 4 && cls_len > 0 && cls_ip < 48 * cls_len && !ctx->opt_no_ranges) || ctx->opt_no_ranges < 0);
    }
    // A row of P products costs P / 12288 + 1 walks over its products there and nwin walks over its entries of A in the
    // bitmap kernel: the bins below nwin * 12288 entries go to the ranges, the rows above stay with the bitmap kernels
    // (C4: the hub rows of the R-MAT factor, up to millions of products).  no_ranges = -1 (tests): everything.
    int rg_hi = kNumBins - 1;
    if (use_ranges && ctx->opt_no_ranges >= 0) {
        const long long lim = nwin_host * 12288ll;
        rg_hi = log_bin((int)(lim < 0x7fffffffll ? lim : 0x7fffffffll), kNumShift) - 1;
        if (rg_hi > kNumBins - 1) rg_hi = kNumBins - 1;
        if (rg_hi < bm_bin) use_ranges = false;
    }
    bm_lo = use_ranges ? rg_hi + 1 : bm_bin;
    auto launch_ranges = [&]() -> int {
        const long long rows = num_rows_in(sp, bm_bin, rg_hi);
        if (rows == 0) return 0;
        const int tmax = 16384;
        const size_t table = (size_t)tmax * (sizeof(real) + sizeof(int));
        const size_t limit = (size_t)ctx->max_smem_optin - (sizeof(FlatScratch<1024, real>) + 256) - sizeof(int) * kRangeBins;
        int nb_max = tmax / 2;
        while (nb_max > 16 && table + (size_t)nb_max * sizeof(int) > limit) nb_max >>= 1;
        const size_t smem = table + (size_t)nb_max * sizeof(int) + sizeof(int) * kRangeBins;
        auto kern = num_hash_ranges_kernel<real>;
        NSP_CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = num_imin(rows, (long long)(sms - push_sms));
        num_prof_class(ctx, "num_hash_ranges", bm_bin, rg_hi);
        kern<<<grid, 1024, smem, ctx->stream>>>(NSP_NUM_ARGS, bm_bin, rg_hi, 6, tmax, nb_max, N, rbshift, ctx->peer_out);
        ctx->prof_end();
        ctx->launches += 1;
        NSP_CUDA_TRY(ctx, cudaGetLastError());
        return 0;
    };
    // One GPU: heaviest class first (the long rows on their side stream, then the main launch), the light classes fill
    // the tail.  Multi-GPU: a tile of C can leave for the peers once ALL rows that overlap it are done, so the long
    // rows start at once (nothing is queued ahead of them: if they had to wait for the light classes next to the main
    // launch, whichever of the two the hardware released first took every SM -- measured: the long rows then ran last
    // and no tile finished before the end), the (short) light classes follow and the main launch, which completes
    // tiles steadily, comes last.
    if (use_ranges) {
        if (peers ? (launch_bitmap(1) != 0 || launch_light() != 0 || launch_ranges() != 0 || launch_bitmap(0) != 0)
                  : (launch_bitmap(1) != 0 || launch_bitmap(0) != 0 || launch_ranges() != 0 || launch_light() != 0))
            return -1;
    } else if (peers) {
        if (launch_bitmap(1) != 0 || launch_light() != 0 || launch_bitmap(0) != 0) return -1;
    } else {
        if (launch_bitmap(1) != 0 || launch_bitmap(0) != 0 || launch_light() != 0) return -1;
    }
    if (sp.join_pending) {
        NSP_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
        sp.join_pending = false;
    }
    if (tma && peer_push_end(ctx) != 0) return -1;
    if (peers && !tma && peer_dma_drive(ctx, c_col - ctx->peer_out.off, c_val - ctx->peer_out.off, (int)sizeof(real)) != 0) return -1;
    return 0;
}

#undef NSP_NUM_ARGS

}  // namespace nsp
