// nsparse-b200: CSR -> AMB conversion (sf_csr2amb, cuda-c/src/conversion/convert_amb.cu:604-929).
//
// The output arrays are the ones the reference produces for the same (seg_size, block_size) on
// rows whose columns are ascending and distinct (checked array by array against the CPU oracle,
// oracle/amb.py, which is pinned to the reference's own convert_amb.cu).  What is re-designed:
//
//   * SPARSE virtual rows.  The reference materialises pad_M x seg_num dense counters, a dense
//     permutation and runs one Thrust sort per (segment, 32768-row window) from a host loop
//     (:646-696); at the 4096^2 Laplacian (M = 2^24, 256 segments) that is 2^32 virtual rows and
//     `int total_pad_row_num` overflows to 0.  Here only the NON-EMPTY virtual rows exist (1.13 per
//     row for that matrix): they are found by run-length encoding the entries' (segment,row) keys,
//     and ordered by two device-wide radix sorts -- (segment,row), then stable by
//     (segment, window, descending count) -- which is exactly what the per-window stable sorts
//     produce.  Chunks are 32 consecutive sorted virtual rows of a window; the rows that pad a
//     window's last chunk are the window's lowest-numbered rows without entries in that segment
//     (then the rows >= M), found by a k-th-missing binary search instead of a dense sort.
//   * the unblocked SELL arrays (:104-136, :313-346) are never written: the blocking kernels read
//     the virtual rows straight from the CSR arrays (fast path: rows with ascending segments) or
//     from one key-sorted copy (rows with columns in arbitrary order).
//   * block size / segment size: the reference's footprint model (:785-797), evaluated for all 20
//     block sizes in ONE pass over the lanes, instead of 100 rebuilds + 1000 timed SpMVs.
//   * no host round trip per step: three scalar read-backs in total (virtual rows, chunks, nnz).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <vector>

#include "../../include/nsparse_b200.h"
#include <stdlib.h>

#include "amb.h"
#include "context.h"

namespace nsp {

namespace {

struct DevPool {
    // temporaries of one conversion; freed together.  Stream-ordered allocations from the context's own
    // memory pool, which keeps what it is given back: a conversion makes ~25 allocations of up to a GB, and
    // with cudaMalloc / cudaFree the mapping and unmapping cost 70-180 ms of the 110-220 ms a conversion of
    // the 4096^2 Laplacian took (the kernels are ~15 ms).
    std::vector<void *> bufs;
    nsp_context *ctx;
    explicit DevPool(nsp_context *c) : ctx(c) {}
    ~DevPool()
    {
        for (void *p : bufs) cudaFreeAsync(p, ctx->stream);
    }
    template <typename T>
    T *take(size_t n)
    {
        void *p = nullptr;
        const cudaError_t e = ctx->mem_pool ? cudaMallocFromPoolAsync(&p, sizeof(T) * (n ? n : 1), ctx->mem_pool, ctx->stream)
                                            : cudaMallocAsync(&p, sizeof(T) * (n ? n : 1), ctx->stream);
        if (e != cudaSuccess) {
            ctx->fail(-4, "amb conversion: cudaMalloc failed");
            return nullptr;
        }
        bufs.push_back(p);
        return reinterpret_cast<T *>(p);
    }
};

#define AMB_TAKE(var, type, n)            \
    type *var = pool.take<type>(n);       \
    if (!var) return -4

typedef unsigned long long u64;

// ---- K1: (segment,row) key of every entry + "some row is not strictly ascending" flag ---------------
__global__ void __launch_bounds__(256)
amb_keys_kernel(const int *__restrict__ rpt, const int *__restrict__ col, int M, u64 pad_M, unsigned seg_size,
                u64 *__restrict__ keys, unsigned char *__restrict__ head, int *__restrict__ unsorted)
{
    __shared__ int s_rpt[257];
    const int r0 = blockIdx.x * 256;
    const int nr = min(256, M - r0);
    for (int i = threadIdx.x; i <= nr; i += 256) s_rpt[i] = rpt[r0 + i];
    __syncthreads();
    const int e0 = s_rpt[0], e1 = s_rpt[nr];
    for (int e = e0 + threadIdx.x; e < e1; e += 256) {
        // row of entry e: last r with s_rpt[r] <= e
        int lo = 0, hi = nr;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_rpt[mid] <= e) lo = mid; else hi = mid;
        }
        const unsigned g = (unsigned)col[e] / seg_size;
        keys[e] = (u64)g * pad_M + (u64)(r0 + lo);
        unsigned char h = 1;
        if (e > s_rpt[lo]) {
            const unsigned gp = (unsigned)col[e - 1] / seg_size;
            h = g != gp;
            if (col[e] <= col[e - 1]) *unsorted = 1;   // not strictly ascending: take the sorted-copy path
        }
        head[e] = h;
    }
}

// sorted-copy path: sort key = (segment,row) key << 16 | column inside the segment
__global__ void amb_sortkey_kernel(const u64 *__restrict__ keys, const int *__restrict__ col, int n, unsigned seg_size,
                                   u64 *__restrict__ skeys)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) skeys[e] = (keys[e] << 16) | (u64)((unsigned)col[e] % seg_size);
}

__global__ void amb_head_from_sorted_kernel(const u64 *__restrict__ skeys, int n, u64 *__restrict__ keys,
                                            unsigned char *__restrict__ head)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) {
        const u64 k = skeys[e] >> 16;
        keys[e] = k;
        head[e] = e == 0 || k != (skeys[e - 1] >> 16);
    }
}

template <typename real>
__global__ void amb_gather_kernel(const unsigned *__restrict__ perm, int n, const int *__restrict__ col,
                                  const real *__restrict__ val, int *__restrict__ col_o, real *__restrict__ val_o)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) {
        const unsigned s = perm[e];
        col_o[e] = col[s];
        val_o[e] = val[s];
    }
}

// ---- K2: per virtual row: key from its first entry, count, and the window-sort key ----------------
// key2 = (segment * W + window) << 32 | (0xffffffff - count)
__global__ void amb_vrow_kernel(const int *__restrict__ vr_start, int nvr, int nnz, const u64 *__restrict__ ekeys,
                                u64 pad_M, unsigned S, unsigned W, u64 *__restrict__ vr_key)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nvr) vr_key[r] = ekeys[vr_start[r]];
}

__global__ void amb_key2_kernel(const u64 *__restrict__ vr_key, const int *__restrict__ vr_start_sorted,
                                const int *__restrict__ vr_cnt_sorted, int nvr, u64 pad_M, unsigned S, unsigned W,
                                u64 *__restrict__ key2, unsigned *__restrict__ group_ne)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nvr) return;
    const u64 k = vr_key[r];
    const u64 g = k / pad_M;
    const unsigned row = (unsigned)(k - g * pad_M);
    const u64 gw = g * W + row / S;
    key2[r] = (gw << 32) | (u64)(0xffffffffu - (unsigned)vr_cnt_sorted[r]);
    atomicAdd(&group_ne[gw], 1u);
}

__global__ void amb_cnt_kernel(const int *__restrict__ vr_start, int nvr, int nnz, int *__restrict__ vr_cnt)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nvr) vr_cnt[r] = (r + 1 < nvr ? vr_start[r + 1] : nnz) - vr_start[r];
}

template <typename T>
__global__ void amb_gather_idx_kernel(const unsigned *__restrict__ idx, int n, const T *__restrict__ in,
                                      T *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[idx[i]];
}

__global__ void amb_iota_kernel(unsigned *p, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (unsigned)i;
}

__global__ void amb_group_chunks_kernel(const unsigned *__restrict__ group_ne, long long ngroups,
                                        unsigned *__restrict__ group_chunks)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < ngroups) group_chunks[q] = (group_ne[q] + 31u) >> 5;
}

// ---- K4: lanes.  Sorted virtual row s of group gw sits at lane cb[gw]*32 + (s - gs[gw]) ------------
__global__ void amb_lanes_kernel(const u64 *__restrict__ key2_sorted, const unsigned *__restrict__ order2,
                                 const u64 *__restrict__ vr_key, const int *__restrict__ vr_start,
                                 const int *__restrict__ vr_cnt, int nvr, u64 pad_M, unsigned W,
                                 const unsigned *__restrict__ gs, const unsigned *__restrict__ cb,
                                 int *__restrict__ lane_start, int *__restrict__ lane_cnt,
                                 int *__restrict__ write_perm, unsigned *__restrict__ chunk_seg)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nvr) return;
    const u64 gw = key2_sorted[s] >> 32;
    const unsigned r = order2[s];
    const unsigned rank = (unsigned)s - gs[gw];
    const long long lane = (long long)cb[gw] * 32 + rank;
    const u64 k = vr_key[r];
    const u64 g = k / pad_M;
    lane_start[lane] = vr_start[r];
    lane_cnt[lane] = vr_cnt[r];
    write_perm[lane] = (int)(k - g * pad_M);
    if ((rank & 31u) == 0) chunk_seg[lane >> 5] = (unsigned)g;
}

// Rows that pad the last chunk of a window: the j-th row (ascending) of the window that has no
// entry in this segment; past the window's end the numbering simply continues (rows >= M).
// One warp per group; rows of the group in ascending order = vr_key[gs .. gs+ne) of the
// (segment,row)-ordered list.
__global__ void amb_fillers_kernel(const unsigned *__restrict__ group_ne, const unsigned *__restrict__ gs,
                                   const unsigned *__restrict__ cb, long long ngroups, const u64 *__restrict__ vr_key,
                                   u64 pad_M, unsigned S, unsigned W, int *__restrict__ lane_start,
                                   int *__restrict__ lane_cnt, int *__restrict__ write_perm)
{
    const long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned t = threadIdx.x & 31u;
    if (q >= ngroups) return;
    const unsigned ne = group_ne[q];
    const unsigned rem = ne & 31u;
    if (ne == 0 || rem == 0 || t < rem) return;
    const unsigned j = t - rem;
    const unsigned w = (unsigned)(q % W);
    const u64 g = (u64)q / W;
    const unsigned win_start = w * S;
    const u64 *rows = vr_key + gs[q];
    const u64 gbase = g * pad_M;
    // k = number of present rows i with (row_i - win_start - i) <= j
    unsigned lo = 0, hi = ne;
    while (lo < hi) {
        const unsigned mid = (lo + hi) >> 1;
        const unsigned row = (unsigned)(rows[mid] - gbase);
        if (row - win_start - mid <= j) lo = mid + 1; else hi = mid;
    }
    const long long lane = ((long long)cb[q] + (ne >> 5)) * 32 + t;
    lane_start[lane] = 0;
    lane_cnt[lane] = 0;
    write_perm[lane] = (int)(win_start + j + lo);
}

// ---- the unblocked SELL row of a lane, read in place ------------------------------------------------
// entry j of the lane: its own j-th entry, or -- padding -- the column of the chunk's first row
// (convert_amb.cu:121-133); columns are taken modulo the segment size (:329).
struct LaneView {
    const int *col;
    int start, cnt, fstart, width;
    unsigned seg_size;
    __device__ __forceinline__ int colmod(int j) const
    {
        const int c = col[(j < cnt ? start : fstart) + j];
        return (int)((unsigned)c % seg_size);
    }
};

__device__ __forceinline__ LaneView lane_view(const int *col, const int *lane_start, const int *lane_cnt,
                                              long long lane, unsigned seg_size)
{
    LaneView v;
    v.col = col;
    v.start = lane_start[lane];
    v.cnt = lane_cnt[lane];
    const long long first = lane & ~31ll;
    v.fstart = lane_start[first];
    v.width = lane_cnt[first];
    v.seg_size = seg_size;
    return v;
}

// ---- K6: blocks needed per chunk for every block size 1..20 (set_blocked_cl, :388-429) -------------
// bs_only == 0: accumulate the totals of all 20 block sizes (planner)
// bs_only >  0: store blocks[chunk] for that block size
__global__ void __launch_bounds__(256)
amb_count_blocks_kernel(const int *__restrict__ col, const int *__restrict__ lane_start,
                        const int *__restrict__ lane_cnt, long long lanes, unsigned seg_size, int bs_only,
                        u64 *__restrict__ totals, int *__restrict__ blocks)
{
    const long long lane = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ u64 s_tot[kAmbMaxBlock];
    if (bs_only == 0) {
        if (threadIdx.x < kAmbMaxBlock) s_tot[threadIdx.x] = 0;
        __syncthreads();
    }
    if (lane < lanes) {   // lanes is a multiple of 32: whole warps in or out
        const LaneView v = lane_view(col, lane_start, lane_cnt, lane, seg_size);
        if (bs_only > 0) {
            // new block when the column leaves the block, or (never on ascending distinct columns,
            // see oracle/amb.py) when an own entry repeats a column
            int base = v.colmod(0), nb = 1, prev = base;
            for (int j = 1; j < v.width; ++j) {
                const int c = v.colmod(j);
                if (c - base >= bs_only || (j < v.cnt && c <= prev)) {
                    base = c;
                    ++nb;
                }
                prev = c;
            }
            nb = __reduce_max_sync(0xffffffffu, nb);
            if ((threadIdx.x & 31) == 0) blocks[lane >> 5] = nb;
        } else {
            int base[kAmbMaxBlock], nb[kAmbMaxBlock];
            const int c0 = v.colmod(0);
#pragma unroll
            for (int b = 0; b < kAmbMaxBlock; ++b) {
                base[b] = c0;
                nb[b] = 1;
            }
            int prev = c0;
            for (int j = 1; j < v.width; ++j) {
                const int c = v.colmod(j);
                const bool repeat = j < v.cnt && c <= prev;
#pragma unroll
                for (int b = 0; b < kAmbMaxBlock; ++b) {
                    if (c - base[b] >= b + 1 || repeat) {
                        base[b] = c;
                        ++nb[b];
                    }
                }
                prev = c;
            }
#pragma unroll
            for (int b = 0; b < kAmbMaxBlock; ++b) {
                const int m = __reduce_max_sync(0xffffffffu, nb[b]);
                if ((threadIdx.x & 31) == 0) atomicAdd(&s_tot[b], (u64)m);
            }
        }
    }
    if (bs_only == 0) {
        __syncthreads();
        if (threadIdx.x < kAmbMaxBlock && s_tot[threadIdx.x]) atomicAdd(&totals[threadIdx.x], s_tot[threadIdx.x]);
    }
}

__global__ void amb_chunk_sizes_kernel(const int *__restrict__ blocks, int c_size, int bs,
                                       long long *__restrict__ sizes)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < c_size) sizes[c] = (long long)blocks[c] * 32 * bs;
}

__global__ void amb_cl_cs_kernel(const int *__restrict__ blocks, const unsigned *__restrict__ chunk_seg,
                                 const long long *__restrict__ offs, int c_size, unsigned *__restrict__ cl,
                                 int *__restrict__ cs)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < c_size) {
        cl[c] = (unsigned)(blocks[c] - 1) | (chunk_seg[c] << 16);
        cs[c] = (int)offs[c];
    }
}

// ---- K8: blocked columns and values (set_blocked_col_val, :473-525) ---------------------------------
template <typename real>
__global__ void __launch_bounds__(256)
amb_fill_kernel(const int *__restrict__ col, const real *__restrict__ val, const int *__restrict__ lane_start,
                const int *__restrict__ lane_cnt, long long lanes, unsigned seg_size, int bs,
                const unsigned *__restrict__ cl, const int *__restrict__ cs, unsigned short *__restrict__ b_col,
                real *__restrict__ b_val)
{
    const long long lane = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= lanes) return;
    const LaneView v = lane_view(col, lane_start, lane_cnt, lane, seg_size);
    const int chunk = (int)(lane >> 5), tid = (int)(lane & 31);
    const int nblocks = (int)(cl[chunk] & 0xffffu) + 1;
    const long long vbase = (long long)cs[chunk] + tid;
    const long long cbase = (long long)cs[chunk] / bs + tid;
    int it = 0;
    for (int k = 0; k < nblocks; ++k) {
        if (it < v.width) {
            const int base = v.colmod(it);
            b_col[cbase + (long long)k * 32] = (unsigned short)base;
            b_val[vbase + (long long)(k * bs) * 32] = it < v.cnt ? val[v.start + it] : real(0);
            ++it;
            for (int h = 1; h < bs; ++h) {
                real x = real(0);
                if (it < v.width && v.colmod(it) - base == h) {
                    if (it < v.cnt) x = val[v.start + it];
                    ++it;
                }
                b_val[vbase + (long long)(k * bs + h) * 32] = x;
            }
        } else {
            const int last = v.colmod(v.width - 1);
            b_col[cbase + (long long)k * 32] = (unsigned short)((last / bs) * bs);
            for (int h = 0; h < bs; ++h) b_val[vbase + (long long)(k * bs + h) * 32] = real(0);
        }
    }
}

// ---- K9: 16-bit write permutation (compress_s_write_permutation, :282-299) ---------------------------
__global__ void amb_perm16_kernel(const int *__restrict__ write_perm, long long lanes,
                                  unsigned short *__restrict__ s_perm, unsigned short *__restrict__ s_off)
{
    const long long lane = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= lanes) return;
    const int p = write_perm[lane];
    s_perm[lane] = (unsigned short)(p % 65536);
    if ((lane & 31) == 0) s_off[lane >> 5] = (unsigned short)(p / 65536);
}

inline int blocks_for(long long n, int bs) { return (int)((n + bs - 1) / bs); }

inline int bits_for(u64 v)
{
    int b = 1;
    while (b < 64 && (v >> b)) ++b;
    return b;
}

template <typename K, typename V>
int radix_sort_pairs(nsp_context *ctx, DevPool &pool, const K *kin, K *kout, const V *vin, V *vout, int n,
                     int begin_bit, int end_bit)
{
    size_t bytes = 0;
    NSP_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, n, begin_bit, end_bit,
                                                      ctx->stream));
    void *tmp = pool.take<char>(bytes);
    if (!tmp) return -4;
    NSP_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(tmp, bytes, kin, kout, vin, vout, n, begin_bit, end_bit,
                                                      ctx->stream));
    return 0;
}

template <typename T>
int exclusive_sum(nsp_context *ctx, DevPool &pool, const T *in, T *out, long long n)
{
    size_t bytes = 0;
    NSP_CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, ctx->stream));
    void *tmp = pool.take<char>(bytes);
    if (!tmp) return -4;
    NSP_CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, n, ctx->stream));
    return 0;
}

// ---- write plan (nsp_amb_plan, context.h): how many virtual rows a row has, two mode bits per lane ----------
__global__ void amb_row_vrows_kernel(const int *__restrict__ lane_cnt, const int *__restrict__ write_perm, long long lanes,
                                     int M, int *__restrict__ row_nvr)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= lanes) return;
    const int row = write_perm[i];
    if (lane_cnt[i] > 0 && row < M) atomicAdd(row_nvr + row, 1);
}

__global__ void amb_mode_kernel(const int *__restrict__ lane_cnt, const int *__restrict__ write_perm, long long lanes, int M,
                                const int *__restrict__ row_nvr, unsigned long long *__restrict__ mode)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // lanes is a multiple of 32
    if (i >= lanes) return;
    const int row = write_perm[i];
    unsigned long long m = 0;
    if (lane_cnt[i] > 0 && row < M) m = row_nvr[row] == 1 ? 2ull : 1ull;
    m <<= 2 * (threadIdx.x & 31);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
    if ((threadIdx.x & 31) == 0) mode[i >> 5] = m;
}

struct AmbNotOne {
    const int *nvr;
    __device__ bool operator()(int i) const { return nvr[i] != 1; }
};

// Everything up to the lane table for one segment size.
struct AmbLanes {
    const int *col = nullptr;      // CSR arrays the lanes index (original or key-sorted copy)
    const void *val = nullptr;
    int *lane_start = nullptr, *lane_cnt = nullptr, *write_perm = nullptr;
    unsigned *chunk_seg = nullptr;
    int c_size = 0;
    long long lanes = 0;
    long long seg_num = 0;
};

template <typename real>
int build_lanes(nsp_context *ctx, DevPool &pool, int M, int N, int nnz, const int *rpt, const int *col,
                const real *val, long long seg_size, AmbLanes &L)
{
    cudaStream_t st = ctx->stream;
    const u64 pad_M = 32ull * ((u64)(M + 31) / 32);
    const u64 G = ((u64)N + (u64)seg_size - 1) / (u64)seg_size;
    const unsigned S = (unsigned)(M < kAmbSigma ? (M > 0 ? M : 1) : kAmbSigma);
    const unsigned W = (unsigned)(((u64)M + S - 1) / S);
    const u64 ngroups = G * (u64)(W ? W : 1);
    L.seg_num = (long long)G;
    if (G > 65536ull) return ctx->fail(-2, "amb: more than 65536 column segments (N too large for this seg_size)");
    if (ngroups > (1ull << 31)) return ctx->fail(-2, "amb: segment x window table too large");
    if (nnz == 0 || M == 0) {
        L.c_size = 0;
        L.lanes = 0;
        return 0;
    }

    // entry keys, virtual-row heads
    AMB_TAKE(ekeys, u64, nnz);
    AMB_TAKE(head, unsigned char, nnz);
    AMB_TAKE(d_flag, int, 4);
    NSP_CUDA_TRY(ctx, cudaMemsetAsync(d_flag, 0, sizeof(int) * 4, st));
    amb_keys_kernel<<<(M + 255) / 256, 256, 0, st>>>(rpt, col, M, pad_M, (unsigned)seg_size, ekeys, head, d_flag);
    ctx->launches++;
    int h_flag = 0;
    NSP_CUDA_TRY(ctx, cudaMemcpyAsync(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    NSP_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    const int key_bits = bits_for(G * pad_M);
    const u64 *keys_for_vr = ekeys;
    L.col = col;
    L.val = val;
    bool grouped_by_seg_row = false;     // virtual-row list already in (segment,row) order?
    if (h_flag) {
        // rows whose columns are not strictly ascending: one stable sort of the entries by
        // (segment, row, column); duplicates keep their CSR order
        AMB_TAKE(skeys, u64, nnz);
        AMB_TAKE(skeys2, u64, nnz);
        AMB_TAKE(eidx, unsigned, nnz);
        AMB_TAKE(eidx2, unsigned, nnz);
        amb_iota_kernel<<<blocks_for(nnz, 256), 256, 0, st>>>(eidx, nnz);
        amb_sortkey_kernel<<<blocks_for(nnz, 256), 256, 0, st>>>(ekeys, col, nnz, (unsigned)seg_size, skeys);
        if (radix_sort_pairs(ctx, pool, skeys, skeys2, eidx, eidx2, nnz, 0, key_bits + 16) != 0) return -1;
        AMB_TAKE(col2, int, nnz);
        AMB_TAKE(val2, real, nnz);
        amb_gather_kernel<real><<<blocks_for(nnz, 256), 256, 0, st>>>(eidx2, nnz, col, val, col2, val2);
        amb_head_from_sorted_kernel<<<blocks_for(nnz, 256), 256, 0, st>>>(skeys2, nnz, ekeys, head);
        ctx->launches += 4;
        keys_for_vr = ekeys;
        L.col = col2;
        L.val = val2;
        grouped_by_seg_row = true;
    }

    // virtual rows = flagged entries
    AMB_TAKE(vr_start0, int, nnz);
    AMB_TAKE(d_nvr, int, 1);
    {
        size_t bytes = 0;
        thrust::counting_iterator<int> it(0);
        NSP_CUDA_TRY(ctx, cub::DeviceSelect::Flagged(nullptr, bytes, it, head, vr_start0, d_nvr, nnz, st));
        void *tmp = pool.take<char>(bytes);
        if (!tmp) return -4;
        NSP_CUDA_TRY(ctx, cub::DeviceSelect::Flagged(tmp, bytes, it, head, vr_start0, d_nvr, nnz, st));
    }
    int nvr = 0;
    NSP_CUDA_TRY(ctx, cudaMemcpyAsync(&nvr, d_nvr, sizeof(int), cudaMemcpyDeviceToHost, st));
    NSP_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    AMB_TAKE(vr_key0, u64, nvr);
    AMB_TAKE(vr_cnt0, int, nvr);
    amb_vrow_kernel<<<blocks_for(nvr, 256), 256, 0, st>>>(vr_start0, nvr, nnz, keys_for_vr, pad_M, S, W, vr_key0);
    amb_cnt_kernel<<<blocks_for(nvr, 256), 256, 0, st>>>(vr_start0, nvr, nnz, vr_cnt0);
    ctx->launches += 2;

    // (segment,row) order
    u64 *vr_key = vr_key0;
    int *vr_start = vr_start0, *vr_cnt = vr_cnt0;
    if (!grouped_by_seg_row && G > 1) {
        AMB_TAKE(k1, u64, nvr);
        AMB_TAKE(i0, unsigned, nvr);
        AMB_TAKE(i1, unsigned, nvr);
        amb_iota_kernel<<<blocks_for(nvr, 256), 256, 0, st>>>(i0, nvr);
        if (radix_sort_pairs(ctx, pool, vr_key0, k1, i0, i1, nvr, 0, key_bits) != 0) return -1;
        AMB_TAKE(s1, int, nvr);
        AMB_TAKE(c1, int, nvr);
        amb_gather_idx_kernel<int><<<blocks_for(nvr, 256), 256, 0, st>>>(i1, nvr, vr_start0, s1);
        amb_gather_idx_kernel<int><<<blocks_for(nvr, 256), 256, 0, st>>>(i1, nvr, vr_cnt0, c1);
        ctx->launches += 3;
        vr_key = k1;
        vr_start = s1;
        vr_cnt = c1;
    }

    // window sort: stable by (segment*W + window, descending count)
    AMB_TAKE(group_ne, unsigned, ngroups);
    NSP_CUDA_TRY(ctx, cudaMemsetAsync(group_ne, 0, sizeof(unsigned) * ngroups, st));
    AMB_TAKE(key2, u64, nvr);
    AMB_TAKE(key2s, u64, nvr);
    AMB_TAKE(o0, unsigned, nvr);
    AMB_TAKE(o1, unsigned, nvr);
    amb_key2_kernel<<<blocks_for(nvr, 256), 256, 0, st>>>(vr_key, vr_start, vr_cnt, nvr, pad_M, S, W, key2, group_ne);
    amb_iota_kernel<<<blocks_for(nvr, 256), 256, 0, st>>>(o0, nvr);
    ctx->launches += 2;
    if (radix_sort_pairs(ctx, pool, key2, key2s, o0, o1, nvr, 0, 32 + bits_for(ngroups)) != 0) return -1;

    // group offsets: gs = start in the sorted list, cb = first chunk
    AMB_TAKE(gs, unsigned, ngroups + 1);
    AMB_TAKE(gchunks, unsigned, ngroups + 1);
    AMB_TAKE(cb, unsigned, ngroups + 1);
    amb_group_chunks_kernel<<<blocks_for((long long)ngroups, 256), 256, 0, st>>>(group_ne, (long long)ngroups, gchunks);
    ctx->launches++;
    if (exclusive_sum(ctx, pool, group_ne, gs, (long long)ngroups) != 0) return -1;
    if (exclusive_sum(ctx, pool, gchunks, cb, (long long)ngroups) != 0) return -1;
    unsigned last_cb = 0, last_ch = 0;
    NSP_CUDA_TRY(ctx, cudaMemcpyAsync(&last_cb, cb + (ngroups - 1), sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    NSP_CUDA_TRY(ctx, cudaMemcpyAsync(&last_ch, gchunks + (ngroups - 1), sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    NSP_CUDA_TRY(ctx, cudaStreamSynchronize(st));
    const long long c_size = (long long)last_cb + last_ch;
    if (c_size * 32 > 0x7fffffffll) return ctx->fail(-3, "amb: more than 2^31 lanes");
    L.c_size = (int)c_size;
    L.lanes = c_size * 32;

    L.lane_start = pool.take<int>(L.lanes);
    L.lane_cnt = pool.take<int>(L.lanes);
    L.write_perm = pool.take<int>(L.lanes);
    L.chunk_seg = pool.take<unsigned>(c_size);
    if (!L.lane_start || !L.lane_cnt || !L.write_perm || !L.chunk_seg) return -4;
    amb_lanes_kernel<<<blocks_for(nvr, 256), 256, 0, st>>>(key2s, o1, vr_key, vr_start, vr_cnt, nvr, pad_M, W, gs, cb,
                                                         L.lane_start, L.lane_cnt, L.write_perm, L.chunk_seg);
    amb_fillers_kernel<<<blocks_for((long long)ngroups * 32, 256), 256, 0, st>>>(
        group_ne, gs, cb, (long long)ngroups, vr_key, pad_M, S, W, L.lane_start, L.lane_cnt, L.write_perm);
    ctx->launches += 2;
    NSP_CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}

// footprint model of the reference (convert_amb.cu:785-791), in its int arithmetic
long long amb_footprint(long long c_nnz, int bs, long long c_size, int M, int vbytes)
{
    long long f = 0;
    f += (c_nnz / bs) * 2;
    f += c_nnz * vbytes;
    f += c_size * 4 * 2;
    f += c_size * 32 * 2 + c_size * 2;
    f += c_size * 32 * vbytes * 2;
    f += (long long)M * vbytes * 2;
    return f;
}

}  // namespace

template <typename real>
int amb_convert(nsp_context *ctx, int M, int N, int nnz, const int *rpt, const int *col, const real *val,
                long long seg_size, int block_size, nsp_amb *out)
{
    if (!out || M < 0 || N < 0 || nnz < 0) return ctx->fail(-2, "nsp_csr2amb: bad argument");
    if (seg_size < 0 || seg_size > 65536) return ctx->fail(-2, "nsp_csr2amb: seg_size must be in [1, 65536] (0 = plan)");
    if (block_size < 0 || block_size > kAmbMaxBlock) return ctx->fail(-2, "nsp_csr2amb: block_size must be in [1, 20] (0 = plan)");
    cudaStream_t st = ctx->stream;
    memset(out, 0, sizeof(*out));

    // candidate segment sizes (sf_csr2amb, :879-892)
    std::vector<long long> segs;
    if (seg_size > 0)
        segs.push_back(seg_size);
    else {
        segs.push_back(65536);
        if (N < 128 * 1024) {
            for (int i = 1; i < 5; ++i) segs.push_back(N < 100 ? i : i * 1024);
        }
    }

    long long best_f = -1, best_seg = segs[0];
    int best_bs = block_size > 0 ? block_size : 1;
    const bool need_plan = segs.size() > 1 || block_size == 0;
    if (need_plan && nnz > 0 && M > 0) {
        for (long long seg : segs) {
            DevPool pool(ctx);
            AmbLanes L;
            if (build_lanes<real>(ctx, pool, M, N, nnz, rpt, col, val, seg, L) != 0) return -1;
            u64 *d_tot = pool.take<u64>(kAmbMaxBlock);
            if (!d_tot) return -4;
            NSP_CUDA_TRY(ctx, cudaMemsetAsync(d_tot, 0, sizeof(u64) * kAmbMaxBlock, st));
            if (L.lanes > 0) {
                amb_count_blocks_kernel<<<blocks_for(L.lanes, 256), 256, 0, st>>>(L.col, L.lane_start, L.lane_cnt, L.lanes,
                                                                                (unsigned)seg, 0, d_tot, nullptr);
                ctx->launches++;
            }
            u64 h_tot[kAmbMaxBlock];
            NSP_CUDA_TRY(ctx, cudaMemcpyAsync(h_tot, d_tot, sizeof(h_tot), cudaMemcpyDeviceToHost, st));
            NSP_CUDA_TRY(ctx, cudaStreamSynchronize(st));
            const int b_lo = block_size > 0 ? block_size : 1, b_hi = block_size > 0 ? block_size : kAmbMaxBlock;
            for (int bs = b_lo; bs <= b_hi; ++bs) {
                const long long c_nnz = (long long)h_tot[bs - 1] * 32 * bs;
                const long long f = amb_footprint(c_nnz, bs, L.c_size, M, (int)sizeof(real));
                if (best_f < 0 || best_f > f) {
                    best_f = f;
                    best_seg = seg;
                    best_bs = bs;
                }
            }
        }
    }

    // build with the chosen parameters
    DevPool pool(ctx);
    AmbLanes L;
    if (build_lanes<real>(ctx, pool, M, N, nnz, rpt, col, val, best_seg, L) != 0) return -1;
    const int bs = best_bs;
    out->block_size = bs;
    out->M = M;
    out->N = N;
    out->pad_M = 32 * ((M + 31) / 32);
    out->chunk = 32;
    out->SIGMA = kAmbSigma;
    out->seg_size = best_seg;
    out->seg_num = L.seg_num;
    out->c_size = L.c_size;
    out->thread_block = 256;
    out->thread_grid = (L.lanes + 255) / 256;
    const int c_size = L.c_size;
    const long long lanes = L.lanes;
    long long c_nnz = 0;
    int *blocks = nullptr;
    long long *sizes = nullptr, *offs = nullptr;
    if (c_size > 0) {
        blocks = pool.take<int>(c_size);
        sizes = pool.take<long long>(c_size + 1);
        offs = pool.take<long long>(c_size + 1);
        if (!blocks || !sizes || !offs) return -4;
        amb_count_blocks_kernel<<<blocks_for(lanes, 256), 256, 0, st>>>(L.col, L.lane_start, L.lane_cnt, lanes,
                                                                      (unsigned)best_seg, bs, nullptr, blocks);
        amb_chunk_sizes_kernel<<<blocks_for(c_size, 256), 256, 0, st>>>(blocks, c_size, bs, sizes);
        ctx->launches += 2;
        if (exclusive_sum(ctx, pool, sizes, offs, (long long)c_size) != 0) return -1;
        long long last_off = 0, last_size = 0;
        NSP_CUDA_TRY(ctx, cudaMemcpyAsync(&last_off, offs + (c_size - 1), sizeof(long long), cudaMemcpyDeviceToHost, st));
        NSP_CUDA_TRY(ctx, cudaMemcpyAsync(&last_size, sizes + (c_size - 1), sizeof(long long), cudaMemcpyDeviceToHost, st));
        NSP_CUDA_TRY(ctx, cudaStreamSynchronize(st));
        c_nnz = last_off + last_size;
        if (c_nnz > 0x7fffffffll)
            return ctx->fail(-3, "amb: padded nnz " + std::to_string(c_nnz) + " does not fit sfAMB's int nnz / int cs[]");
    }
    out->nnz = (int)c_nnz;

    // the seven arrays of sfAMB; plain cudaMalloc because release_amb frees them with cudaFree
    const size_t nc = c_size ? c_size : 1, nl = lanes ? lanes : 1, nv = c_nnz ? c_nnz : 1;
    NSP_CUDA_TRY(ctx, cudaMalloc((void **)&out->d_cs, sizeof(int) * nc));
    NSP_CUDA_TRY(ctx, cudaMalloc((void **)&out->d_cl, sizeof(unsigned) * nc));
    NSP_CUDA_TRY(ctx, cudaMalloc((void **)&out->d_sellcs_col, sizeof(unsigned short) * (nv / bs + 1)));
    NSP_CUDA_TRY(ctx, cudaMalloc((void **)&out->d_sellcs_val, sizeof(real) * nv));
    NSP_CUDA_TRY(ctx, cudaMalloc((void **)&out->d_s_write_permutation, sizeof(unsigned short) * nl));
    NSP_CUDA_TRY(ctx, cudaMalloc((void **)&out->d_s_write_permutation_offset, sizeof(unsigned short) * nc));
    NSP_CUDA_TRY(ctx, cudaMalloc((void **)&out->d_write_permutation, sizeof(int) * nl));
    if (c_size > 0) {
        amb_cl_cs_kernel<<<blocks_for(c_size, 256), 256, 0, st>>>(blocks, L.chunk_seg, offs, c_size, out->d_cl, out->d_cs);
        amb_fill_kernel<real><<<blocks_for(lanes, 256), 256, 0, st>>>(
            L.col, (const real *)L.val, L.lane_start, L.lane_cnt, lanes, (unsigned)best_seg, bs, out->d_cl, out->d_cs,
            out->d_sellcs_col, (real *)out->d_sellcs_val);
        amb_perm16_kernel<<<blocks_for(lanes, 256), 256, 0, st>>>(L.write_perm, lanes, out->d_s_write_permutation,
                                                                out->d_s_write_permutation_offset);
        NSP_CUDA_TRY(ctx, cudaMemcpyAsync(out->d_write_permutation, L.write_perm, sizeof(int) * lanes,
                                          cudaMemcpyDeviceToDevice, st));
        ctx->launches += 3;
    }
    // write plan of the SpMV (see nsp_amb_plan)
    if (c_size > 0 && !getenv("NSPARSE_AMB_NO_PLAN")) {
        nsp_amb_plan wp;
        int *row_nvr = pool.take<int>((size_t)M + 1);
        int *zr = pool.take<int>((size_t)M + 1);
        int *d_nz = pool.take<int>(1);
        if (!row_nvr || !zr || !d_nz) return -4;
        NSP_CUDA_TRY(ctx, cudaMemsetAsync(row_nvr, 0, sizeof(int) * ((size_t)M + 1), st));
        NSP_CUDA_TRY(ctx, cudaMalloc((void **)&wp.d_mode, sizeof(unsigned long long) * (size_t)c_size));
        amb_row_vrows_kernel<<<blocks_for(lanes, 256), 256, 0, st>>>(L.lane_cnt, L.write_perm, lanes, M, row_nvr);
        amb_mode_kernel<<<blocks_for(lanes, 256), 256, 0, st>>>(L.lane_cnt, L.write_perm, lanes, M, row_nvr, wp.d_mode);
        ctx->launches += 2;
        {
            size_t bytes = 0;
            thrust::counting_iterator<int> it(0);
            AmbNotOne pred{row_nvr};
            NSP_CUDA_TRY(ctx, cub::DeviceSelect::If(nullptr, bytes, it, zr, d_nz, M, pred, st));
            void *tmp = pool.take<char>(bytes);
            if (!tmp) return -4;
            NSP_CUDA_TRY(ctx, cub::DeviceSelect::If(tmp, bytes, it, zr, d_nz, M, pred, st));
        }
        int nz = 0;
        NSP_CUDA_TRY(ctx, cudaMemcpyAsync(&nz, d_nz, sizeof(int), cudaMemcpyDeviceToHost, st));
        NSP_CUDA_TRY(ctx, cudaStreamSynchronize(st));
        wp.n_zero_rows = nz;
        if (nz > 0) {
            NSP_CUDA_TRY(ctx, cudaMalloc((void **)&wp.d_zero_rows, sizeof(int) * (size_t)nz));
            NSP_CUDA_TRY(ctx, cudaMemcpyAsync(wp.d_zero_rows, zr, sizeof(int) * (size_t)nz, cudaMemcpyDeviceToDevice, st));
        }
        wp.M = M;
        wp.c_size = c_size;
        ctx->amb_plans[out->d_cs] = wp;
    }
    NSP_CUDA_TRY(ctx, cudaGetLastError());
    NSP_CUDA_TRY(ctx, cudaStreamSynchronize(st));   // the pool's temporaries are freed on return
    return 0;
}

template int amb_convert<float>(nsp_context *, int, int, int, const int *, const int *, const float *, long long, int,
                                nsp_amb *);
template int amb_convert<double>(nsp_context *, int, int, int, const int *, const int *, const double *, long long, int,
                                 nsp_amb *);

}  // namespace nsp
