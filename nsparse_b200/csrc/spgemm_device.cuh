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
        // every warp scans the (at most 32) warp totals itself: no serial loop, no second barrier
        int w = lane < GROUP / 32 ? wtot[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += x;
        }
        const int before = __shfl_sync(0xffffffffu, w, (wid + 31) & 31);
        inc += wid > 0 ? before : 0;
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
// own A entry (compare + divergent advance loop).  Heavy rows of C are unions of LONG B rows (R-MAT
// scale 20: 1300 entries on average, 95 % of the products come from B rows longer than 256), so here
// the unit of work is a PART: up to kPartLen consecutive products of ONE B row.  Warp w takes the
// parts w, w + 32, ... of the staged slab (static round robin: a shared-memory claim counter cost
// 10 % of the stall samples of the round-1 kernel), finds the part's entry with two ballots, issues
// ALL loads of the part before the first use (8 x 128 B of B.col, and of B.val, in flight per
// warp), and then every lane handles its products with a warp-uniform a_ij.
// ---------------------------------------------------------------------------------------------
#ifndef NSP_PART_LEN
#define NSP_PART_LEN 256
#endif
constexpr int kPartLen = NSP_PART_LEN;

template <int BS, typename real>
struct PartScratch {
    int pre[BS + 1];   // exclusive prefix of the parts of the slab's entries
    int pre1[BS / 32]; // pre[32 * i]: bank-conflict-free first level of the entry search
    int kb[BS];        // first product of the entry's current segment in B.col / B.val
    int len[BS];       // products of the segment
    union {
        struct {
            int cur[BS];   // cursor: first product of the entry not yet consumed by the value chunks
            int end[BS];   // end of the entry's segment in the current column window
        };
        int tab[2 * BS];   // segment mode: tab[k * E + e] = first product of entry e at or beyond chunk boundary k
    };
    real av[BS];       // a_ij (numeric only)
    int wtot[BS / 32 + 1];
};

// Parts of a segment [kb, kb + len) of B.col / B.val: consecutive kPartLen-element pieces of the global
// arrays counted from the 16-byte aligned index kb & ~3, so that the mark pass can read a part with
// 128-bit loads; elements before kb in the first part are masked off.
__device__ __forceinline__ int parts_of(int kb, int len)
{
    return len > 0 ? ((kb & 3) + len + kPartLen - 1) / kPartLen : 0;
}

// parts prefix of the staged segments; two barriers; returns the number of parts.  kFlat (classes whose B rows are
// short, see run_flat): the prefix counts PRODUCTS instead of parts.
template <int BS, typename real, bool kFlat = false>
__device__ __forceinline__ int scan_parts(int t, int kb, int len, PartScratch<BS, real> &s)
{
    const int np = kFlat ? len : parts_of(kb, len);
    const int inc = group_inclusive_scan<BS>(np, t, s.wtot);
    s.pre[t + 1] = inc;
    if ((t & 31) == 0) s.pre1[t >> 5] = inc - np;
    if (t == 0) s.pre[0] = 0;
    __syncthreads();
    return s.pre[BS];
}

// stage_parts restricted to the products whose column lies in [col_lo, col_hi): the heavy kernels
// walk a row of C in ascending column ranges (bitmap windows, accumulator chunks) and every product
// is visited once per pass because the sub-range of each B row is found by binary search instead of
// filtering.  use_lo / use_hi say which bounds actually cut (both false: the whole B row).
// One thread per entry, both bounds from scratch: the path of rows with more than BS entries.
template <int BS, bool kLoadVal, typename real, bool kFlat = false>
__device__ __forceinline__ int stage_parts_range(int t, int base, int a_end, const int *__restrict__ a_col,
                                                 const real *__restrict__ a_val, const int *__restrict__ b_rpt,
                                                 const int *__restrict__ b_col, int col_lo, int col_hi,
                                                 bool use_lo, bool use_hi, PartScratch<BS, real> &s)
{
    int len = 0, kb = 0;
    if (base + t < a_end) {
        const int ac = ld_stream(a_col + base + t);
        kb = ld_nc(b_rpt + ac);
        int ke = ld_nc(b_rpt + ac + 1);
        if (use_lo || use_hi) {
            // the two searches run interleaved (independent load chains)
            int l0 = kb, h0 = use_lo ? ke : kb, l1 = kb, h1 = use_hi ? ke : kb;
            while ((l0 < h0) | (l1 < h1)) {
                const int m0 = (l0 + h0) >> 1, m1 = (l1 + h1) >> 1;
                const bool a0 = l0 < h0, a1 = l1 < h1;
                const int c0 = a0 ? ld_nc(b_col + m0) : 0;
                const int c1 = a1 ? ld_nc(b_col + m1) : 0;
                if (a0) {
                    if (c0 < col_lo) l0 = m0 + 1; else h0 = m0;
                }
                if (a1) {
                    if (c1 < col_hi) l1 = m1 + 1; else h1 = m1;
                }
            }
            if (use_hi) ke = l1;
            if (use_lo) kb = l0;
            if (ke < kb) ke = kb;
        }
        len = ke - kb;
        if (kLoadVal) s.av[t] = ld_stream(a_val + base + t);
    }
    s.kb[t] = kb;
    s.len[t] = len;
    return scan_parts<BS, real, kFlat>(t, kb, len, s);
}

// Lower bound by a GROUP of g = 2^glog lanes (g-ary search: ceil(log_{g+1} n) dependent loads instead
// of log_2 n).  All 32 lanes of the warp call; lo / hi / key are uniform within a group; a group
// with lo >= hi takes no part.  Returns the first k in [lo, hi) with b_col[k] >= key.
__device__ __forceinline__ int group_lower_bound(const int *__restrict__ b_col, int lo, int hi, int key, int glog,
                                                 int lane)
{
    const int g = 1 << glog;
    const int gl = lane & (g - 1);
    const int gfirst = lane & ~(g - 1);
    const unsigned gmask = g == 32 ? 0xffffffffu : ((1u << g) - 1u);
    while (__any_sync(0xffffffffu, lo < hi)) {
        const int n = hi - lo;
        const int step = n > 0 ? (int)(((unsigned)n + (unsigned)g) / (unsigned)(g + 1)) : 1;   // ceil(n / (g + 1)) >= 1
        const int p = lo + gl * step + (step - 1);                                           // probe i = gl
        const bool below = n > 0 && p < hi && ld_nc(b_col + p) < key;
        const int m = __popc((__ballot_sync(0xffffffffu, below) >> gfirst) & gmask);
        if (n > 0) {
            // probes 0 .. m-1 are below the key, probe m (if inside) is not
            const int nlo = lo + m * step;
            const int nhi = m < g ? min(hi, nlo + step - 1) : hi;
            lo = nlo;
            hi = nhi;
        }
    }
    return lo;
}

// ---- cursor staging for rows whose A entries fit one slab (E <= BS; 99.9 % of the heavy rows) ------
// glog: log2 of the lanes that share an entry, min(5, log2(BS / pow2ceil(E))).
__device__ __forceinline__ int entry_group_log(int E, int BS)
{
    int glog = 0;
    while (glog < 5 && (E << (glog + 1)) <= BS) ++glog;
    return glog;
}

// Window stage: segment of every B row inside the window [., c1).  first: the window is the first of
// the row (segments start at the B row's start), otherwise they start where the previous window ended
// (s.end).  cut_hi: the window does not reach N, so the end is searched.  Leaves kb/len = the window
// segment, cur = its start, end = its end, and the part prefix.
template <int BS, bool kLoadVal, typename real, bool kFlat = false>
__device__ __forceinline__ int stage_window(int t, int a_beg, int E, int glog, const int *__restrict__ a_col,
                                            const real *__restrict__ a_val, const int *__restrict__ b_rpt,
                                            const int *__restrict__ b_col, int c1, bool first, bool cut_hi,
                                            PartScratch<BS, real> &s)
{
    const int e = t >> glog;
    if (((t & ~31) >> glog) < E) {          // warp-uniform: this warp owns at least one entry
        int lo = 0, hi = 0;
        if (e < E) {
            const int ac = ld_stream(a_col + a_beg + e);
            lo = first ? ld_nc(b_rpt + ac) : s.end[e];
            hi = ld_nc(b_rpt + ac + 1);
        }
        int res = hi;
        if (cut_hi) res = group_lower_bound(b_col, lo, hi, c1, glog, t & 31);
        __syncwarp();      // every lane of the entry's group has read s.end[e] before its leader overwrites it (racecheck)
        if (e < E && (t & ((1 << glog) - 1)) == 0) {
            s.kb[e] = lo;
            s.len[e] = res - lo;
            s.cur[e] = lo;
            s.end[e] = res;
            if (kLoadVal && first) s.av[e] = ld_stream(a_val + a_beg + e);
        }
    }
    __syncthreads();
    return scan_parts<BS, real, kFlat>(t, t < E ? s.kb[t] : 0, t < E ? s.len[t] : 0, s);
}

// Chunk stage: segment [cur, first column >= col_hi) of every entry (last: up to the window end), and
// the cursor moves on.
template <int BS, typename real, bool kFlat = false>
__device__ __forceinline__ int stage_chunk(int t, int E, int glog, const int *__restrict__ b_col, int col_hi,
                                           bool last, PartScratch<BS, real> &s)
{
    const int e = t >> glog;
    if (((t & ~31) >> glog) < E) {
        int lo = 0, hi = 0;
        if (e < E) {
            lo = s.cur[e];
            hi = s.end[e];
        }
        int res = hi;
        if (!last) res = group_lower_bound(b_col, lo, hi, col_hi, glog, t & 31);
        __syncwarp();      // every lane of the entry's group has read s.cur[e] before its leader moves it (racecheck)
        if (e < E && (t & ((1 << glog) - 1)) == 0) {
            s.kb[e] = lo;
            s.len[e] = res - lo;
            s.cur[e] = res;
        }
    }
    __syncthreads();
    return scan_parts<BS, real, kFlat>(t, t < E ? s.kb[t] : 0, t < E ? s.len[t] : 0, s);
}

// ---- segment mode (B column-sorted, at most 4 windows): the window cuts of every B row are PRECOMPUTED per
// entry of A by entry_segments_kernel (spgemm_plan.cu): seg[w * stride + j] = first product of entry j's B row at
// or beyond column w << wshift (w = 0: the row's start, w = nwin: its end).  A window stage is then two coalesced
// loads per entry instead of the dependent chain a_col -> B.rpt -> search probes of stage_window, which was pure
// exposed latency at the start of every window (profiles/r2_phase_cycles_num_bitmap_s20_baseline.txt: "mark
// stage" 7 % of the kernel).
template <int BS, bool kLoadVal, typename real, bool kFlat = false>
__device__ __forceinline__ int stage_window_seg(int t, int a_beg, int E, const real *__restrict__ a_val,
                                                const int *__restrict__ seg, long long stride, int win, bool first,
                                                PartScratch<BS, real> &s)
{
    int lo = 0, hi = 0;
    if (t < E) {
        lo = ld_nc(seg + (long long)win * stride + a_beg + t);
        hi = ld_nc(seg + (long long)(win + 1) * stride + a_beg + t);
        s.kb[t] = lo;
        s.len[t] = hi - lo;
        if (kLoadVal && first) s.av[t] = ld_stream(a_val + a_beg + t);
    }
    return scan_parts<BS, real, kFlat>(t, lo, hi - lo, s);
}

// the same for one slab of a row with more than BS entries
template <int BS, bool kLoadVal, typename real, bool kFlat = false>
__device__ __forceinline__ int stage_slab_seg(int t, int base, int a_end, const real *__restrict__ a_val,
                                              const int *__restrict__ seg, long long stride, int win,
                                              PartScratch<BS, real> &s)
{
    int lo = 0, hi = 0;
    if (base + t < a_end) {
        lo = ld_nc(seg + (long long)win * stride + base + t);
        hi = ld_nc(seg + (long long)(win + 1) * stride + base + t);
        if (kLoadVal) s.av[t] = ld_stream(a_val + base + t);
    }
    s.kb[t] = lo;
    s.len[t] = hi - lo;
    return scan_parts<BS, real, kFlat>(t, lo, hi - lo, s);
}

// All chunk boundaries of a window at once: tab[k * E + e], k = 0 .. nch, from the staged window segments
// (s.kb / s.len) and the boundary columns bound[1 .. nch-1].  E * (nch - 1) independent searches run side by
// side (2^q lanes each), so the window pays ONE chain of dependent probes instead of one per chunk
// (stage_chunk: "chunk stage" 10 % of the kernel).  Needs E * (nch + 1) <= 2 * BS.  Ends with a barrier.
template <int BS, typename real>
__device__ __forceinline__ void build_chunk_table(int t, int E, int nch, const int *__restrict__ b_col,
                                                  const int *bound, PartScratch<BS, real> &s)
{
    const int lane = t & 31;
    if (t < E) {
        s.tab[t] = s.kb[t];
        s.tab[nch * E + t] = s.kb[t] + s.len[t];
    }
    const int S = E * (nch - 1);
    int q = 0;
    while (q < 5 && (S << (q + 1)) <= BS) ++q;
    for (int base = 0; base < S; base += BS >> q) {
        const int sidx = base + (t >> q);
        int lo = 0, hi = 0, key = 0, e = 0, k = 0;
        if (sidx < S) {
            k = sidx / E + 1;
            e = sidx - (k - 1) * E;
            lo = s.kb[e];
            hi = lo + s.len[e];
            key = bound[k];
        }
        const int res = group_lower_bound(b_col, lo, hi, key, q, lane);
        if (sidx < S && (t & ((1 << q) - 1)) == 0) s.tab[k * E + e] = res;
    }
    __syncthreads();
}

// chunk k of the window from the table
template <int BS, typename real, bool kFlat = false>
__device__ __forceinline__ int stage_chunk_tab(int t, int E, int k, PartScratch<BS, real> &s)
{
    int lo = 0, hi = 0;
    if (t < E) {
        lo = s.tab[k * E + t];
        hi = s.tab[(k + 1) * E + t];
        s.kb[t] = lo;
        s.len[t] = hi - lo;
    }
    return scan_parts<BS, real, kFlat>(t, lo, hi - lo, s);
}

// chunk [col_lo, col_hi) of the window when the table does not fit: both ends searched in the window segment,
// which is re-read from the precomputed cuts (first / last: that end is the window's)
template <int BS, typename real, bool kFlat = false>
__device__ __forceinline__ int stage_chunk_seg(int t, int a_beg, int E, int glog, const int *__restrict__ b_col,
                                               const int *__restrict__ seg, long long stride, int win, int col_lo,
                                               int col_hi, bool first, bool last, PartScratch<BS, real> &s)
{
    const int e = t >> glog;
    if (((t & ~31) >> glog) < E) {
        int wl = 0, wh = 0;
        if (e < E) {
            wl = ld_nc(seg + (long long)win * stride + a_beg + e);
            wh = ld_nc(seg + (long long)(win + 1) * stride + a_beg + e);
        }
        int lo = wl, hi = wh;
        if (!first) lo = group_lower_bound(b_col, wl, wh, col_lo, glog, t & 31);
        if (!last) hi = group_lower_bound(b_col, lo, wh, col_hi, glog, t & 31);
        if (e < E && (t & ((1 << glog) - 1)) == 0) {
            s.kb[e] = lo;
            s.len[e] = hi - lo;
        }
    }
    __syncthreads();
    return scan_parts<BS, real, kFlat>(t, t < E ? s.kb[t] : 0, t < E ? s.len[t] : 0, s);
}

// entry of part q = largest e with pre[e] <= q (two ballots: 32-ary search over pre1, then over pre)
template <int BS, typename real>
__device__ __forceinline__ int part_entry(const PartScratch<BS, real> &s, int q, int lane)
{
    static_assert(BS == 1024, "two-level 32-ary entry search");
    const unsigned q1 = __ballot_sync(0xffffffffu, s.pre1[lane] <= q);
    const int e = (__popc(q1) - 1) * 32;
    const unsigned q2 = __ballot_sync(0xffffffffu, s.pre[e + lane] <= q);
    return e + __popc(q2) - 1;
}

// Walk the staged slab's parts (value pass: lane-strided, so that the 32 products of one instruction
// have 32 different ranks); the caller guarantees a barrier between the staging and this call.
// Ends with a barrier.  Both kernels are ISSUE-bound once the shared-memory conflicts are gone
// (profiles/r1_ncu_bitmap_s20_v3.txt: 63 % issue-slot utilisation, 8 warp instructions per product),
// so full parts -- four out of five -- run without any per-element predicate.
template <int BS, bool kNumeric, typename real, typename F>
__device__ __forceinline__ void run_parts(int t, int total, const int *__restrict__ b_col,
                                          const real *__restrict__ b_val, PartScratch<BS, real> &s, F &&f)
{
    constexpr int NW = BS / 32;
    constexpr int U = kPartLen / 32;
    const int lane = t & 31;
    for (int q = t >> 5; q < total; q += NW) {
        const int e = part_entry<BS, real>(s, q, lane);
        const int kb = s.kb[e];
        const int ke = kb + s.len[e];
        const int p0 = (kb & ~3) + (q - s.pre[e]) * kPartLen;     // first element of the part
        const real av = kNumeric ? s.av[e] : real(0);
        const int *pc = b_col + p0 + lane;
        const real *pv = b_val + p0 + lane;
        int c[U];
        real v[U];
        if (p0 >= kb && p0 + kPartLen <= ke) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                c[u] = ld_nc(pc + 32 * u);
                if (kNumeric) v[u] = ld_nc(pv + 32 * u);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) f(c[u], kNumeric ? av * v[u] : real(0));
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int kk = p0 + lane + 32 * u;
                c[u] = -1;
                if (kk >= kb && kk < ke) {
                    c[u] = ld_nc(pc + 32 * u);
                    if (kNumeric) v[u] = ld_nc(pv + 32 * u);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (c[u] >= 0) f(c[u], kNumeric ? av * v[u] : real(0));
        }
    }
    __syncthreads();
}

// Walk of the staged segments for classes whose B rows are SHORT (C = A * B with a few entries per row of B:
// configs C4 / C5): a part -- up to 256 products of ONE B row -- would hold 4 products there and a warp would spend
// its ~100 instructions of part bookkeeping on them.  Here the prefix of the staged segments counts products
// (scan_parts<kFlat>), a warp takes 128 consecutive products of the slab, finds the entry of the first one with the
// same two-ballot search and every lane advances through the entries on its own (the traversal of the hash classes,
// for_each_product, on pre-cut segments).  Ends with a barrier.
template <int BS, bool kNumeric, typename real, typename F>
__device__ __forceinline__ void run_flat(int t, int total, const int *__restrict__ b_col, const real *__restrict__ b_val,
                                         PartScratch<BS, real> &s, F &&f)
{
    constexpr int U = 4;
    const int lane = t & 31;
    for (int base_p = (t >> 5) * (32 * U); base_p < total; base_p += BS * U) {
        int e = part_entry<BS, real>(s, base_p, lane);
        int next = s.pre[e + 1];
        int koff = s.kb[e] - s.pre[e];
        real av = kNumeric ? s.av[e] : real(0);
        int c[U];
        real v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = base_p + u * 32 + lane;
            c[u] = -1;
            v[u] = real(0);
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
            if (c[u] >= 0) f(c[u], v[u]);
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Column bitmap of the heavy kernels: plain bit order (column cc of the window is bit cc & 31 of
// 32-bit word cc >> 5); the 64-bit word PAIRS are swizzled inside every batch of 32 pairs (2048
// columns): pair j of batch b lives at j ^ (b & 31).  Graph generators such as
// R-MAT make every bit of a column index 0 with probability ~0.76, so a quarter of ALL products share
// any given 5-bit field of the column index and would hit one bank (measured: 14 wavefronts per
// shared-memory atomic instead of ~3); after the swizzle the bank depends on the column bits 6..15
// (one fold: two instructions per address -- the kernels are issue-bound).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned bitmap_swz(unsigned batch) { return batch & 31u; }

// physical index of logical 32-bit word w32 (= cc >> 5)
__device__ __forceinline__ unsigned bitmap_word32(unsigned w32)
{
    return w32 ^ ((w32 >> 5) & 0x3eu);
}

// Mark pass over the staged parts.  Every lane reads FOUR CONSECUTIVE products per 128-bit load (two
// loads per part, all in flight together), merges the bits that fall into the same 32-bit word in
// registers, and issues one shared-memory atomicOr per distinct word: a dense B row (hub vertex)
// costs one atomic per four products and an 8-way instead of a 32-way same-word conflict, a sparse one
// costs what it did.  b_vec_end: largest index k (multiple of 4) such that b_col[k .. k+3] may be read
// with one load (0 disables the vector path, e.g. for a misaligned B.col).  c0 / ncols: the window;
// kFilter drops columns outside it (unsorted B).
template <bool kChecked, bool kFilter>
__device__ __forceinline__ void mark4(unsigned *bm32, const int4 c, int k, int kb, int ke, int c0, unsigned ncols)
{
    const int col[4] = {c.x, c.y, c.z, c.w};
    unsigned w[4], m[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const unsigned cc = (unsigned)(col[j] - c0);
        bool ok = true;
        if (kChecked) ok = k + j >= kb && k + j < ke;
        if (kFilter) ok = ok && cc < ncols;
        w[j] = (kChecked || kFilter) ? (ok ? cc >> 5 : 0xffffffffu) : cc >> 5;
        m[j] = 1u << (cc & 31u);
    }
#pragma unroll
    for (int j = 1; j < 4; ++j) {
        if (w[j] == w[j - 1]) {          // same word as the predecessor: hand its bits on
            m[j] |= m[j - 1];
            w[j - 1] = 0xffffffffu;
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (w[j] != 0xffffffffu) atomicOr(bm32 + bitmap_word32(w[j]), m[j]);
}

template <int BS, bool kFilter, typename real>
__device__ __forceinline__ void run_parts_mark(int t, int total, const int *__restrict__ b_col, int b_vec_end,
                                               PartScratch<BS, real> &s, unsigned *bm32, int c0, unsigned ncols)
{
    constexpr int NW = BS / 32;
    constexpr int U = kPartLen / 128;
    const int lane = t & 31;
    for (int q = t >> 5; q < total; q += NW) {
        const int e = part_entry<BS, real>(s, q, lane);
        const int kb = s.kb[e];
        const int ke = kb + s.len[e];
        const int p0 = (kb & ~3) + (q - s.pre[e]) * kPartLen;
        const int k0 = p0 + 4 * lane;
        int4 c[U];
        if (p0 >= kb && p0 + kPartLen <= ke && p0 + kPartLen <= b_vec_end) {
#pragma unroll
            for (int u = 0; u < U; ++u) c[u] = __ldg(reinterpret_cast<const int4 *>(b_col + k0 + 128 * u));
#pragma unroll
            for (int u = 0; u < U; ++u) mark4<false, kFilter>(bm32, c[u], 0, 0, 0, c0, ncols);
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = k0 + 128 * u;
                c[u] = make_int4(0, 0, 0, 0);
                if (k < ke) {
                    if (k + 4 <= b_vec_end) {
                        c[u] = __ldg(reinterpret_cast<const int4 *>(b_col + k));
                    } else {
                        c[u].x = ld_nc(b_col + k);
                        if (k + 1 < ke) c[u].y = ld_nc(b_col + k + 1);
                        if (k + 2 < ke) c[u].z = ld_nc(b_col + k + 2);
                        if (k + 3 < ke) c[u].w = ld_nc(b_col + k + 3);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) mark4<true, kFilter>(bm32, c[u], k0 + 128 * u, kb, ke, c0, ncols);
        }
    }
    __syncthreads();
}

// Accumulator index swizzle.  On a dense row of C the rank of a column is (nearly) the column itself,
// so the ranks inherit the bit skew of the column indices: a quarter of the shared-memory CAS of a warp
// hit one bank (measured 16-22 wavefronts per ATOMS.CAST instead of ~5).  XOR-ing the next five index
// bits into the bank bits is a permutation inside every aligned group of 32 accumulators, so the
// copy-out stays conflict free.
__device__ __forceinline__ int acc_swz(int i) { return i ^ ((i >> 5) & 31); }

// ---------------------------------------------------------------------------------------------
// Open-addressing insert, linear probing, key-only (symbolic).  Returns 1 if the key was new.
// Same scheme as the reference probe loop (kernel_spgemm_hash_d.cu:299-317) but the table is
// sized per row (mask) and never runs above 3/4 load.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int hash_insert_key(int *tab, unsigned mask, int col)
{
    unsigned h = hash_slot(col, mask);
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
    unsigned h = hash_slot(col, mask);
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
