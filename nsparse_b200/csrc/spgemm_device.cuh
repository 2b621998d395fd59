// nsparse-b200: device building blocks shared by the symbolic and numeric SpGEMM kernels.
#pragma once

#include "common.cuh"
#include "spgemm_plan.h"

namespace nsp {

// group-wide barrier: a warp (GROUP == 32, several rows per CTA) or the whole CTA
template <int GROUP>
__device__ __forceinline__ void group_sync()
{
    if (GROUP == 32)
        __syncwarp();
    else
        __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Row traversal -- load-balanced over PRODUCTS, not over A entries.
//
// A "group" of GROUP threads (a warp or a CTA) owns one row i of C.  The reference gives every
// warp one entry a_ij and lets its lanes stride the B row j (kernel_spgemm_hash_d.cu:427-430):
// on heavy-tailed inputs one warp then walks a 40 000-entry B row alone while 31 warps idle
// (measured here: 1.1 G products/s for the 1024-thread classes on R-MAT scale 20).  Instead the
// group stages a slab of up to GROUP entries of the A row in shared memory as
//     s_kb[e]  = B.rpt[a_col[e]]            first product of entry e in B.col / B.val
//     s_pre[e] = sum_{e' < e} nnz(B_{a_col[e']})   (exclusive scan; s_pre[GROUP] = slab total)
//     s_av[e]  = a_val[e]                   (numeric only)
// and then walks the slab's products p = 0 .. total-1 with stride GROUP: thread t finds its entry
// by a branch-free binary search over s_pre (log2(GROUP) shared-memory reads, mostly broadcasts)
// and reads B.col[s_kb[e] + p - s_pre[e]].  Consecutive lanes hit consecutive addresses inside a
// B row, every lane has work whatever the B-row length distribution, and four products per thread
// are in flight before the first table update.  f(col, value) is called once per product.
// ---------------------------------------------------------------------------------------------
template <int GROUP, typename real>
struct FlatScratch {
    int pre[GROUP + 1];
    int kb[GROUP];
    real av[GROUP];
    int wtot[GROUP / 32 + 1];
};

template <int GROUP>
__device__ __forceinline__ int group_inclusive_scan(int v, int t, int *wtot)
{
    const int lane = t & 31;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int x = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += x;
    }
    if (GROUP > 32) {
        const int wid = t >> 5;
        if (lane == 31) wtot[wid] = inc;
        __syncthreads();
        int add = 0;
        for (int w = 0; w < wid; ++w) add += wtot[w];
        inc += add;
    }
    return inc;
}

template <int GROUP, bool kNumeric, typename real, typename F>
__device__ __forceinline__ void for_each_product(int t, int a_beg, int a_end,
                                                 const int *__restrict__ a_col,
                                                 const real *__restrict__ a_val,
                                                 const int *__restrict__ b_rpt,
                                                 const int *__restrict__ b_col,
                                                 const real *__restrict__ b_val,
                                                 FlatScratch<GROUP, real> &s, F &&f)
{
    constexpr int U = 4;
    for (int base = a_beg; base < a_end; base += GROUP) {
        int len = 0, kb = 0;
        if (base + t < a_end) {
            const int ac = ld_stream(a_col + base + t);
            kb = ld_nc(b_rpt + ac);
            len = ld_nc(b_rpt + ac + 1) - kb;
            if (kNumeric) s.av[t] = ld_stream(a_val + base + t);
        }
        s.kb[t] = kb;
        const int inc = group_inclusive_scan<GROUP>(len, t, s.wtot);
        s.pre[t + 1] = inc;
        if (t == 0) s.pre[0] = 0;
        group_sync<GROUP>();
        const int total = s.pre[GROUP];
        // Every warp takes contiguous chunks of 32*U products.  The entry owning the first product
        // of a chunk is found ONCE per chunk by a cooperative 32-ary search (two ballots); the lanes
        // then only advance monotonically, keeping the current entry's bounds in registers.  This
        // costs ~1 compare per product on long B rows (a per-product binary search made the
        // kernels issue-bound at ~100 warp instructions per product: profiles/r1_*).
        constexpr int S1 = GROUP / 32;       // level-1 stride of the 32-ary search
        const int lane = t & 31;
        for (int base_p = (t >> 5) * (32 * U); base_p < total; base_p += GROUP * U) {
            int e = 0;
            {
                const unsigned q1 = __ballot_sync(0xffffffffu, s.pre[lane * S1] <= base_p);
                e = (__popc(q1) - 1) * S1;
                if (S1 > 1) {
                    const unsigned q2 = __ballot_sync(0xffffffffu, lane < S1 && s.pre[e + lane] <= base_p);
                    e += __popc(q2) - 1;
                }
            }
            int next = s.pre[e + 1];
            int koff = s.kb[e] - s.pre[e];
            real av = kNumeric ? s.av[e] : real(0);
            int c[U];
            real v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int p = base_p + u * 32 + lane;
                c[u] = -1;
                if (p < total) {
                    if (p >= next) {
                        do {
                            ++e;
                            next = s.pre[e + 1];
                        } while (p >= next);
                        koff = s.kb[e] - s.pre[e];
                        if (kNumeric) av = s.av[e];
                    }
                    c[u] = ld_nc(b_col + koff + p);
                    if (kNumeric) v[u] = av * ld_nc(b_val + koff + p);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (c[u] >= 0) f(c[u], kNumeric ? v[u] : real(0));
        }
        group_sync<GROUP>();
    }
}

// ---------------------------------------------------------------------------------------------
// Row traversal for the CTA-per-row kernels of the heavy classes -- WARP-ALIGNED PARTS.
//
// The flat traversal above balances perfectly but pays for it per product: every lane tracks its
// own A entry (compare + divergent advance loop), which made the bitmap kernels issue-bound at
// 1.6 (symbolic) / 2.4 (numeric, per pass) warp instructions per product with 27 % of the stall
// samples on barriers (profiles/r1_ncu_full_bitmap_scale18.txt).  Heavy rows of C are unions of
// LONG B rows (R-MAT scale 20: 1300 entries on average), so here the unit of work is a PART: up to
// kPartLen consecutive products of ONE B row.  A warp claims parts from a shared-memory counter
// (dynamic balance at 256-product granularity), finds the part's entry with two ballots, and then
// all 32 lanes stream the same B row: one coalesced 128-byte load of B.col (and B.val) per 32
// products, a warp-uniform a_ij, no per-lane bookkeeping.
// ---------------------------------------------------------------------------------------------
#ifndef NSP_PART_LEN
#define NSP_PART_LEN 256
#endif
constexpr int kPartLen = NSP_PART_LEN;

template <int BS, typename real>
struct PartScratch {
    int pre[BS + 1];   // exclusive prefix of the parts of the slab's entries
    int kb[BS];        // first product of the entry in B.col / B.val
    int len[BS];       // products of the entry
    real av[BS];       // a_ij (numeric only)
    int wtot[BS / 32 + 1];
    int next;          // next unclaimed part
};

// Stage the slab [base, base + BS) of the A row: B-row starts / lengths / a_ij and the part prefix.
// Ends with a barrier; returns the number of parts.
template <int BS, bool kLoadVal, typename real>
__device__ __forceinline__ int stage_parts(int t, int base, int a_end, const int *__restrict__ a_col,
                                           const real *__restrict__ a_val, const int *__restrict__ b_rpt,
                                           PartScratch<BS, real> &s)
{
    int len = 0, kb = 0;
    if (base + t < a_end) {
        const int ac = ld_stream(a_col + base + t);
        kb = ld_nc(b_rpt + ac);
        len = ld_nc(b_rpt + ac + 1) - kb;
        if (kLoadVal) s.av[t] = ld_stream(a_val + base + t);
    }
    s.kb[t] = kb;
    s.len[t] = len;
    const int inc = group_inclusive_scan<BS>((len + kPartLen - 1) / kPartLen, t, s.wtot);
    s.pre[t + 1] = inc;
    if (t == 0) s.pre[0] = 0;
    __syncthreads();
    return s.pre[BS];
}

// Walk the staged slab's parts.  The caller guarantees a barrier between stage_parts (or the
// previous run_parts) and this call, and s.next == BS / 32 on entry; ends with a barrier that
// also re-arms s.next.
template <int BS, bool kNumeric, typename real, typename F>
__device__ __forceinline__ void run_parts(int t, int total, const int *__restrict__ b_col,
                                          const real *__restrict__ b_val, PartScratch<BS, real> &s, F &&f)
{
    constexpr int NW = BS / 32;
    constexpr int S1 = BS / 32;   // level-1 stride of the 32-ary entry search
    const int lane = t & 31;
    int q = t >> 5;
    while (q < total) {
        // entry of part q = largest e with pre[e] <= q
        const unsigned q1 = __ballot_sync(0xffffffffu, s.pre[lane * S1] <= q);
        int e = (__popc(q1) - 1) * S1;
        if (S1 > 1) {
            const unsigned q2 = __ballot_sync(0xffffffffu, lane < S1 && s.pre[e + lane] <= q);
            e += __popc(q2) - 1;
        }
        const int k0 = s.kb[e] + (q - s.pre[e]) * kPartLen;
        const int k1 = min(s.kb[e] + s.len[e], k0 + kPartLen);
        const real av = kNumeric ? s.av[e] : real(0);
#pragma unroll 1
        for (int k = k0 + lane; k < k1; k += 128) {
            int c[4];
            real v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int kk = k + 32 * u;
                c[u] = -1;
                if (kk < k1) {
                    c[u] = ld_nc(b_col + kk);
                    if (kNumeric) v[u] = av * ld_nc(b_val + kk);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (c[u] >= 0) f(c[u], kNumeric ? v[u] : real(0));
        }
        if (lane == 0) q = atomicAdd(&s.next, 1);
        q = __shfl_sync(0xffffffffu, q, 0);
    }
    __syncthreads();
    if (t == 0) s.next = NW;
}

template <int BS, bool kNumeric, typename real, typename F>
__device__ __forceinline__ void for_each_product_parts(int t, int a_beg, int a_end,
                                                       const int *__restrict__ a_col,
                                                       const real *__restrict__ a_val,
                                                       const int *__restrict__ b_rpt,
                                                       const int *__restrict__ b_col,
                                                       const real *__restrict__ b_val,
                                                       PartScratch<BS, real> &s, F &&f)
{
    for (int base = a_beg; base < a_end; base += BS) {
        if (t == 0) s.next = BS / 32;      // ordered before the claims by stage_parts' barriers
        const int total = stage_parts<BS, kNumeric, real>(t, base, a_end, a_col, a_val, b_rpt, s);
        run_parts<BS, kNumeric, real>(t, total, b_col, b_val, s, f);
    }
}

// ---------------------------------------------------------------------------------------------
// Open-addressing insert, linear probing, key-only (symbolic).  Returns 1 if the key was new.
// Same scheme as the reference probe loop (kernel_spgemm_hash_d.cu:299-317) but the table is
// sized per row (mask) and never runs above 3/4 load.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int hash_insert_key(int *tab, unsigned mask, int col)
{
    unsigned h = hash_col(col) & mask;
    while (true) {
        const int k = *((volatile int *)(tab + h));
        if (k == col) return 0;
        if (k == kEmptyKey) {
            const int old = atomicCAS(tab + h, kEmptyKey, col);
            if (old == kEmptyKey) return 1;
            if (old == col) return 0;
        }
        h = (h + 1) & mask;
    }
}

// key + value accumulate (numeric).  ref: kernel_spgemm_hash_d.cu:871-888
template <typename real>
__device__ __forceinline__ void hash_accumulate(int *keys, real *vals, unsigned mask, int col, real v)
{
    unsigned h = hash_col(col) & mask;
    while (true) {
        const int k = *((volatile int *)(keys + h));
        if (k == col) break;
        if (k == kEmptyKey) {
            const int old = atomicCAS(keys + h, kEmptyKey, col);
            if (old == kEmptyKey || old == col) break;
        }
        h = (h + 1) & mask;
    }
    atomicAdd(vals + h, v);
}

// Bitonic sort of n (power of two) (key, value) slots by unsigned key, so that the free slots
// (key == -1 == 0xffffffff) end up behind the occupied ones.  Replaces the O(n^2) rank-by-counting
// sort of the reference (kernel_spgemm_hash_d.cu:917-925).
template <int GROUP, typename real>
__device__ __forceinline__ void bitonic_sort_slots(int *keys, real *vals, int n, int t)
{
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = t; i < (n >> 1); i += GROUP) {
                const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                const int hi = lo + j;
                const unsigned a = (unsigned)keys[lo], b = (unsigned)keys[hi];
                const bool up = (lo & k) == 0;
                if ((a > b) == up) {
                    keys[lo] = (int)b;
                    keys[hi] = (int)a;
                    const real va = vals[lo];
                    vals[lo] = vals[hi];
                    vals[hi] = va;
                }
            }
            group_sync<GROUP>();
        }
    }
}

// table size for a row that holds at most `cnt` distinct keys: next pow2 of 4/3*cnt, >= 32
__device__ __forceinline__ int table_size_for(int cnt, int tmax)
{
    if (cnt > tmax) cnt = tmax;
    int want = cnt + cnt / 3 + 1;
    if (want < 32) want = 32;
    int ts = 1 << (32 - __clz(want - 1));
    return ts < tmax ? ts : tmax;
}

}  // namespace nsp
