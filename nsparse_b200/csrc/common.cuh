// nsparse-b200: shared device/host helpers.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace nsp {

// ---------------------------------------------------------------------------------------------
// hardware parameters of the B200 this library is written for.  The reference derives its bin
// ladder from 48 KiB smem/block, 64 KiB/SM (spgemm_hash_kernel_gen.c:40-44); here the ladder is
// derived from the 227 KiB opt-in limit and re-checked against cudaDeviceProp at context creation.
// ---------------------------------------------------------------------------------------------
constexpr int kWarp = 32;
constexpr int kMaxSmemOptin = 227 * 1024;   // bytes per CTA on sm_100
constexpr int kStaticSmemReserve = 34 * 1024;   // static __shared__ of the 1024-thread kernels (row slab)
constexpr int kEmptyKey = -1;              // columns are >= 0, so -1 marks a free slot (ref: init_check)
// The reference hashes with (col * 107) & (size - 1) (HASH_SCAL, kernel_spgemm_hash_d.cu:30, :299): the
// low bits of the product depend only on the low bits of the column, so inputs whose column indices
// have skewed low bits (every Kronecker / R-MAT graph: a quarter of all entries end in five zero
// bits) pile into a few runs of slots and linear probing degenerates -- measured here 81 us per
// 2000-product row.  Fibonacci hashing takes the HIGH bits of col * 2^32/phi instead; the result of
// the SpGEMM does not depend on the hash.
constexpr unsigned kHashMul = 0x9E3779B1u;

// Multi-GPU allgatherv of C, overlapped with the numeric phase.  Every GPU holds the FULL C.col / C.val arrays; the
// bases of the other GPUs' copies are mapped into this process (CUDA IPC, or plain peer access inside one process)
// and `off` is the element displacement of this rank's row block in them.  The block is cut into TILES of
// 2^tile_log consecutive entries, aligned in the full arrays.  The numeric kernels only COUNT: whoever has written
// entries [s, e) of the block adds the overlap to the counter of every tile it touches (tiles_done below); the
// thread that completes a tile hands it on, and something that is NOT a computing SM moves it to the peers:
//   * default (peer_dma.cu): tiles of 2^24 .. 2^26 entries, heavy rows processed tile by tile; the completing thread
//     raises a flag in host-mapped memory, the host thread that issued the product polls the flags and gives every
//     finished tile to the copy engines (cudaMemcpyAsync to each peer on its own stream);
//   * option "gather_tma" (peer_push.cu): tiles of 4096 entries in a ready queue, a persistent pusher kernel on a few
//     SMs of its own moves them with the TMA alone (cp.async.bulk global -> shared -> peer global).  Measured: one SM
//     sustains only ~8-15 GB/s of remote stores, so saturating NVLink needs a third of the GPU
//     (profiles/r2_bench_g2_pusher_*.json).
// n == 0: single GPU.
constexpr int kMaxPeerOut = 7;
constexpr int kTileLog = 12;               // pusher tiles: 4096 entries, 16 KiB of C.col + 16 / 32 KiB of C.val
struct PeerOut {
    int n = 0;
    long long off = 0;
    int *col[kMaxPeerOut] = {};
    void *val[kMaxPeerOut] = {};
    int tile_log = kTileLog;
    long long tile0 = 0;        // index of the first tile of the block (off >> tile_log)
    int ntiles = 0;
    long long nnz = 0;          // entries of the block
    int *tile_cnt = nullptr;    // [ntiles] entries written so far
    int *queue = nullptr;       // pusher: [ntiles] ready tiles (block relative), -1 = not yet published
    int *q_ctl = nullptr;       // pusher: [0] tail (producers), [1] head (pusher tickets), [2] error flag
    int *done = nullptr;        // copy engines: [ntiles] flags in host-mapped memory
};

// entries the block [off, off + nnz) has in its tile t (block relative)
__host__ __device__ __forceinline__ int tile_len(const PeerOut &p, int t)
{
    const long long lo = (p.tile0 + t) << p.tile_log, hi = lo + (1ll << p.tile_log);
    const long long a = lo > p.off ? lo : p.off, b = hi < p.off + p.nnz ? hi : p.off + p.nnz;
    return (int)(b - a);
}

#ifdef __CUDACC__
// Called by ONE thread after the entries [first, first + count) of this rank's block (indices relative to the
// block) have been written to the local C.col AND C.val by its group (CTA or warp) and a group barrier has
// ordered those writes before this call; the fence makes them visible at GPU scope before the counters move.
__device__ __forceinline__ void tiles_done(const PeerOut &p, long long first, long long count)
{
    if (count <= 0) return;
    __threadfence();
    const long long s = p.off + first, e = s + count;
    for (long long t = s >> p.tile_log; t <= (e - 1) >> p.tile_log; ++t) {
        const long long lo = t << p.tile_log, hi = lo + (1ll << p.tile_log);
        const int add = (int)((e < hi ? e : hi) - (s > lo ? s : lo));
        const int ti = (int)(t - p.tile0);
        const int old = atomicAdd(p.tile_cnt + ti, add);
        if (old + add == tile_len(p, ti)) {
            if (p.done) {
                __threadfence_system();                       // every writer's entries, before the host sees the flag
                *reinterpret_cast<volatile int *>(p.done + ti) = 1;
            } else {
                __threadfence();                              // the other writers' entries (seen through the counter)
                const int slot = atomicAdd(p.q_ctl, 1);
                asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.queue + slot), "r"(ti) : "memory");
            }
        }
    }
}
#endif

struct Error {
    int code;
    std::string msg;
};

#define NSP_CUDA_TRY(ctx, expr)                                                                   \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            (ctx)->fail(-1, std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + \
                                ":" + std::to_string(__LINE__));                                  \
            return -1;                                                                            \
        }                                                                                         \
    } while (0)

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_nc(const int *p) { return __ldg(p); }
__device__ __forceinline__ float ld_nc(const float *p) { return __ldg(p); }
__device__ __forceinline__ double ld_nc(const double *p) { return __ldg(p); }
__device__ __forceinline__ long long ld_nc(const long long *p) { return __ldg(p); }

// streamed (read-once) data: bypass L1 allocation so the gathered B rows keep the cache
__device__ __forceinline__ int ld_stream(const int *p)
{
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_stream(const float *p)
{
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_stream(const double *p)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// slot of `col` in a table of mask + 1 = 2^k slots
__device__ __forceinline__ unsigned hash_slot(int col, unsigned mask)
{
    return ((unsigned)col * kHashMul) >> (32 - __popc(mask));
}

__host__ __device__ __forceinline__ int next_pow2_int(int v)
{
    // smallest power of two >= v, v >= 1
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

__device__ __forceinline__ int warp_sum(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// log2-style bin of a per-row count:  v <= (1<<s) -> 0 ; else ceil(log2(v)) - s
__host__ __device__ __forceinline__ int log_bin(int v, int s)
{
    if (v <= (1 << s)) return 0;
#ifdef __CUDA_ARCH__
    return 32 - __clz(v - 1) - s;
#else
    int b = 0;
    unsigned u = (unsigned)(v - 1);
    while (u) {
        ++b;
        u >>= 1;
    }
    return b - s;
#endif
}

constexpr int kNumBins = 28;   // log_bin(INT_MAX, 4) = 27

}  // namespace nsp
