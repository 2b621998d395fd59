// nsparse-b200: AMB SpMV y = A x (sf_spmv_amb, cuda-c/src/kernel/kernel_spmv_amb.cu:10-104).
//
// The format is pinned by sfAMB (nsparse.h:78-107), so the thread <-> virtual-row mapping of the
// reference kernel is kept: lane i of chunk i>>5 walks (cl & 0xffff) + 1 column blocks of
// block_size values each, values interleaved by 32 so a warp reads 32 consecutive reals per step.
// What changes on B200:
//   * y is cleared with a stream-ordered memset (copy-engine rate) instead of a kernel (:10-19);
//   * fp64 partial sums go to y with the native red.global.add.f64, not the CAS loop (:70-76);
//   * the matrix stream (values, 16-bit columns, 16-bit permutation) is read with
//     ld.global.nc.L1::no_allocate so the 227 KB L1 keeps x, which is the only reused operand;
//     the reference reads it with ld.global.cv (uncached) and x through the texture path;
//   * the block loop is unrolled so every lane has up to 4 column blocks (4..20 independent value
//     loads) in flight before the first FMA retires;
//   * x is never read beyond N: block columns are clamped to N-1 (the stored value there is 0),
//     because the sample driver's pad x[N..N+20) is uninitialised (spmv_amb.cu:32-34);
//   * lanes whose row is >= M (rows that pad the last chunk) do not touch y, so y needs M entries,
//     not M + 32.
#include "amb.h"
#include "context.h"

namespace nsp {

__device__ __forceinline__ unsigned ld_stream_u16(const unsigned short *p)
{
    unsigned short v;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return v;
}

template <typename real, int BS>
__global__ void __launch_bounds__(256)
amb_spmv_kernel(real *__restrict__ y, const real *__restrict__ value, const unsigned short *__restrict__ col,
                const unsigned *__restrict__ cl, const int *__restrict__ cs, const real *__restrict__ x,
                const unsigned short *__restrict__ perm, const unsigned short *__restrict__ perm_off, int lanes,
                int seg_size, int M, int N, const unsigned long long *__restrict__ mode)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= lanes) return;
    const int chunk = i >> 5, lane = threadIdx.x & 31;
    const int row = (int)ld_stream_u16(perm + i) + (int)perm_off[chunk] * 65536;
    const int start = cs[chunk];
    const unsigned length = cl[chunk];
    const int nblk = (int)(length & 0xffffu) + 1;
    const int c_off = (int)(length >> 16) * seg_size;
    const real *v = value + start + lane;
    const unsigned short *c = col + start / BS + lane;
    const int nmax = N - 1;
    real acc = real(0);
    int h = 0;
    // U column blocks per trip: the U column loads and U*BS value loads are independent
    constexpr int U = BS <= 2 ? 4 : (BS <= 5 ? 2 : 1);
    for (; h + U <= nblk; h += U) {
        int cc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) cc[u] = (int)ld_stream_u16(c + (h + u) * 32) + c_off;
        real vv[U][BS];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int b = 0; b < BS; ++b) vv[u][b] = ld_stream(v + ((h + u) * BS + b) * 32);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int b = 0; b < BS; ++b) acc += vv[u][b] * ld_nc(x + min(cc[u] + b, nmax));
    }
    for (; h < nblk; ++h) {
        const int cc = (int)ld_stream_u16(c + h * 32) + c_off;
#pragma unroll
        for (int b = 0; b < BS; ++b) acc += ld_stream(v + (h * BS + b) * 32) * ld_nc(x + min(cc + b, nmax));
    }
    if (mode) {
        // write plan (nsp_amb_plan): 2 = this lane holds its whole row (plain store), 1 = the row has entries in
        // several column segments (red into the zeroed entry), 0 = padding lane
        const unsigned m = (unsigned)(ld_nc(reinterpret_cast<const long long *>(mode) + chunk) >> (2 * lane)) & 3u;
        if (m == 2u)
            y[row] = acc;
        else if (m == 1u)
            atomicAdd(y + row, acc);
    } else if (row < M) {
        atomicAdd(y + row, acc);
    }
}

template <typename real>
__global__ void __launch_bounds__(256)
amb_zero_rows_kernel(real *__restrict__ y, const int *__restrict__ rows, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[rows[i]] = real(0);
}

template <typename real, int BS>
struct AmbLaunch {
    static int run(nsp_context *ctx, const nsp_amb *mat, const real *x, real *y, int bs, int tb,
                   const unsigned long long *mode)
    {
        if (bs == BS) {
            const int lanes = mat->c_size * 32;
            const int grid = (lanes + tb - 1) / tb;
            amb_spmv_kernel<real, BS><<<grid, tb, 0, ctx->stream>>>(
                y, (const real *)mat->d_sellcs_val, mat->d_sellcs_col, mat->d_cl, mat->d_cs, x,
                mat->d_s_write_permutation, mat->d_s_write_permutation_offset, lanes, (int)mat->seg_size, mat->M,
                mat->N, mode);
            return 0;
        }
        return AmbLaunch<real, BS + 1>::run(ctx, mat, x, y, bs, tb, mode);
    }
};

template <typename real>
struct AmbLaunch<real, kAmbMaxBlock + 1> {
    static int run(nsp_context *ctx, const nsp_amb *, const real *, real *, int bs, int, const unsigned long long *)
    {
        return ctx->fail(-2, "nsp_spmv_amb: block_size " + std::to_string(bs) + " outside [1, 20]");
    }
};

template <typename real>
int amb_spmv(nsp_context *ctx, const nsp_amb *mat, const real *x, real *y)
{
    if (!mat || !y || (!x && mat->N > 0)) return ctx->fail(-2, "nsp_spmv_amb: bad argument");
    if (mat->M <= 0) return 0;
    // write plan: most rows live in ONE column segment and are written by a plain store; only the others need a
    // zeroed y (the reference zeroes all of y in a kernel and adds every virtual row atomically,
    // kernel_spmv_amb.cu:10-19, :70-76)
    const unsigned long long *mode = nullptr;
    auto it = ctx->amb_plans.find(mat->d_cs);
    if (it != ctx->amb_plans.end() && mat->c_size > 0 && it->second.M == mat->M && it->second.c_size == mat->c_size) {
        const nsp_amb_plan &wp = it->second;
        mode = wp.d_mode;
        if ((long long)wp.n_zero_rows * 3 > (long long)mat->M) {
            NSP_CUDA_TRY(ctx, cudaMemsetAsync(y, 0, sizeof(real) * (size_t)mat->M, ctx->stream));
        } else if (wp.n_zero_rows > 0) {
            amb_zero_rows_kernel<real><<<(wp.n_zero_rows + 255) / 256, 256, 0, ctx->stream>>>(y, wp.d_zero_rows, wp.n_zero_rows);
            ctx->launches += 1;
        }
    } else {
        NSP_CUDA_TRY(ctx, cudaMemsetAsync(y, 0, sizeof(real) * (size_t)mat->M, ctx->stream));
    }
    if (mat->c_size <= 0) return 0;
    int tb = (int)mat->thread_block;
    if (tb < 32 || tb > 256 || (tb & 31)) tb = 256;
    if (AmbLaunch<real, 1>::run(ctx, mat, x, y, mat->block_size, tb, mode) != 0) return -2;
    ctx->launches += 1;
    NSP_CUDA_TRY(ctx, cudaGetLastError());
    return 0;
}

template int amb_spmv<float>(nsp_context *, const nsp_amb *, const float *, float *);
template int amb_spmv<double>(nsp_context *, const nsp_amb *, const double *, double *);

}  // namespace nsp
