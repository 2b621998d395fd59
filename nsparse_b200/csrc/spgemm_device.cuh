// nsparse-b200: device building blocks shared by the symbolic and numeric SpGEMM kernels.
#pragma once

#include "common.cuh"
#include "spgemm_plan.h"

namespace nsp {

// ---------------------------------------------------------------------------------------------
// Row traversal.  A "group" of GT threads owns one row of C.  It is split into sub-groups of LB
// lanes; each sub-group takes one entry a_ij of the A row at a time and strides the B row j with
// its LB lanes, so loads of B.col / B.val are coalesced runs of LB elements.  LB is picked on the
// host from the mean B-row length seen by the class (long rows: 32, ER-like 4-nnz rows: 4).
// The (column, rpt pair) of the NEXT A entry is fetched before the current B row is walked so
// the dependent a_col -> b_rpt -> b_col chain of the reference (kernel_spgemm_hash_d.cu:427-430)
// is overlapped with the probe loop, and the B row is read four strides at a time.
// f(col, value) is called once per intermediate product.
// ---------------------------------------------------------------------------------------------
template <int GT, int LB, bool kNumeric, typename real, typename F>
__device__ __forceinline__ void for_each_product(int t, int a_beg, int a_end,
                                                 const int *__restrict__ a_col,
                                                 const real *__restrict__ a_val,
                                                 const int *__restrict__ b_rpt,
                                                 const int *__restrict__ b_col,
                                                 const real *__restrict__ b_val, F &&f)
{
    constexpr int NSG = GT / LB;
    constexpr int U = 4;
    const int sg = t / LB, sl = t % LB;
    int j = a_beg + sg;
    int kb = 0, ke = 0;
    real av = real(0);
    if (j < a_end) {
        const int ac = ld_stream(a_col + j);
        if (kNumeric) av = ld_stream(a_val + j);
        kb = ld_nc(b_rpt + ac);
        ke = ld_nc(b_rpt + ac + 1);
    }
    while (j < a_end) {
        const int jn = j + NSG;
        int kbn = 0, ken = 0;
        real avn = real(0);
        if (jn < a_end) {
            const int ac = ld_stream(a_col + jn);
            if (kNumeric) avn = ld_stream(a_val + jn);
            kbn = ld_nc(b_rpt + ac);
            ken = ld_nc(b_rpt + ac + 1);
        }
        for (int k = kb + sl; k < ke; k += U * LB) {
            int c[U];
            real v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int kk = k + u * LB;
                c[u] = kk < ke ? ld_nc(b_col + kk) : -1;
                if (kNumeric) v[u] = kk < ke ? ld_nc(b_val + kk) : real(0);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (c[u] >= 0) f(c[u], kNumeric ? av * v[u] : real(0));
            }
        }
        j = jn;
        kb = kbn;
        ke = ken;
        av = avn;
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

// group-wide barrier: a warp (GROUP == 32, several rows per CTA) or the whole CTA
template <int GROUP>
__device__ __forceinline__ void group_sync()
{
    if (GROUP == 32)
        __syncwarp();
    else
        __syncthreads();
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
    int want = cnt + cnt / 3 + 1;
    if (want < 32) want = 32;
    int ts = 1 << (32 - __clz(want - 1));
    return ts < tmax ? ts : tmax;
}

}  // namespace nsp
