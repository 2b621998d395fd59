// nsparse-b200: extern "C" entry points of the AMB path (include/nsparse_b200.h).
#include <string.h>

#include "../../include/nsparse_b200.h"
#include "amb.h"
#include "context.h"

namespace {

void forget_plan(nsp_context *ctx, const void *d_cs)
{
    auto it = ctx->amb_plans.find(d_cs);
    if (it == ctx->amb_plans.end()) return;
    cudaFree(it->second.d_mode);
    cudaFree(it->second.d_zero_rows);
    ctx->amb_plans.erase(it);
}

template <typename real>
int free_arrays(nsp_context *ctx, nsp_amb *m)
{
    forget_plan(ctx, m->d_cs);
    cudaFree(m->d_cs);
    cudaFree(m->d_cl);
    cudaFree(m->d_sellcs_col);
    cudaFree(m->d_sellcs_val);
    cudaFree(m->d_s_write_permutation);
    cudaFree(m->d_s_write_permutation_offset);
    cudaFree(m->d_write_permutation);
    m->d_cs = nullptr;
    m->d_cl = nullptr;
    m->d_sellcs_col = nullptr;
    m->d_sellcs_val = nullptr;
    m->d_s_write_permutation = nullptr;
    m->d_s_write_permutation_offset = nullptr;
    m->d_write_permutation = nullptr;
    return 0;
}

// evaluate_spmv (convert_amb.cu:556-600): mean of the runs after the first, for every launch shape
template <typename real>
int time_spmv(nsp_context *ctx, nsp_amb *m, const real *x, real *y, float *best_ms, long long *best_tb)
{
    cudaEvent_t e0, e1;
    NSP_CUDA_TRY(ctx, cudaEventCreate(&e0));
    NSP_CUDA_TRY(ctx, cudaEventCreate(&e1));
    *best_ms = 1e30f;
    for (int tb = 64; tb <= 256; tb *= 2) {
        m->thread_block = tb;
        m->thread_grid = ((long long)m->c_size * 32 + tb - 1) / tb;
        float acc = 0.f;
        const int reps = 3;
        for (int i = 0; i < reps; ++i) {
            cudaEventRecord(e0, ctx->stream);
            if (nsp::amb_spmv<real>(ctx, m, x, y) != 0) return -1;
            cudaEventRecord(e1, ctx->stream);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (i > 0) acc += ms;
        }
        acc /= reps - 1;
        if (acc < *best_ms) {
            *best_ms = acc;
            *best_tb = tb;
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}

template <typename real>
int csr2amb(nsp_context *ctx, int M, int N, int nnz, const int *d_rpt, const int *d_col, const real *d_val,
            long long seg_size, int block_size, int autotune, const real *d_x, nsp_amb *out)
{
    if (!ctx) return NSP_ERR_ARG;
    cudaSetDevice(ctx->device);
    if (!autotune) return nsp::amb_convert<real>(ctx, M, N, nnz, d_rpt, d_col, d_val, seg_size, block_size, out);
    // timing search of the reference's `AT` build: every candidate (segment size, block size) is
    // built and timed; seg_size / block_size, when given, restrict the search
    if (!d_x) return ctx->fail(NSP_ERR_ARG, "nsp_csr2amb: autotune needs d_x");
    real *d_y = nullptr;
    NSP_CUDA_TRY(ctx, cudaMalloc((void **)&d_y, sizeof(real) * (size_t)(M > 0 ? M : 1)));
    long long segs[5] = {65536, 0, 0, 0, 0};
    int nseg = 1;
    if (seg_size > 0)
        segs[0] = seg_size;
    else if (N < 128 * 1024)
        for (int i = 1; i < 5; ++i) segs[nseg++] = N < 100 ? i : i * 1024;
    float best = 1e30f;
    long long best_seg = segs[0], best_tb = 256;
    int best_bs = block_size > 0 ? block_size : 1;
    for (int s = 0; s < nseg; ++s) {
        const int b_lo = block_size > 0 ? block_size : 1, b_hi = block_size > 0 ? block_size : nsp::kAmbMaxBlock;
        for (int bs = b_lo; bs <= b_hi; ++bs) {
            nsp_amb m;
            if (nsp::amb_convert<real>(ctx, M, N, nnz, d_rpt, d_col, d_val, segs[s], bs, &m) != 0) {
                cudaFree(d_y);
                return -1;
            }
            float ms = 0.f;
            long long tb = 256;
            const int rc = time_spmv<real>(ctx, &m, d_x, d_y, &ms, &tb);
            free_arrays<real>(ctx, &m);
            if (rc != 0) {
                cudaFree(d_y);
                return -1;
            }
            if (ms < best) {
                best = ms;
                best_seg = segs[s];
                best_bs = bs;
                best_tb = tb;
            }
        }
    }
    cudaFree(d_y);
    const int rc = nsp::amb_convert<real>(ctx, M, N, nnz, d_rpt, d_col, d_val, best_seg, best_bs, out);
    if (rc == 0) {
        out->thread_block = best_tb;
        out->thread_grid = ((long long)out->c_size * 32 + best_tb - 1) / best_tb;
    }
    return rc;
}

template <typename real>
int spmv_host(nsp_context *ctx, const nsp_amb *mat, const real *h_x, real *h_y)
{
    if (!ctx || !mat) return NSP_ERR_ARG;
    cudaSetDevice(ctx->device);
    // x / y staging on the device: grow-only buffers of the context (an iterative solver calls this in a loop)
    const size_t need = sizeof(real) * ((size_t)(mat->N > 0 ? mat->N : 1) + (size_t)(mat->M > 0 ? mat->M : 1)) + 256;
    if (need > ctx->spmv_stage_bytes) {
        NSP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_spmv_stage);
        ctx->d_spmv_stage = nullptr;
        ctx->spmv_stage_bytes = 0;
        NSP_CUDA_TRY(ctx, cudaMalloc((void **)&ctx->d_spmv_stage, need));
        ctx->spmv_stage_bytes = need;
    }
    real *d_x = reinterpret_cast<real *>(ctx->d_spmv_stage);
    real *d_y = reinterpret_cast<real *>(ctx->d_spmv_stage + ((sizeof(real) * (size_t)(mat->N > 0 ? mat->N : 1) + 255) & ~size_t(255)));
    cudaStream_t st = ctx->stream;
    int rc = 0;
    if (cudaMemcpyAsync(d_x, h_x, sizeof(real) * (size_t)mat->N, cudaMemcpyHostToDevice, st) != cudaSuccess) rc = -1;
    if (rc == 0) rc = nsp::amb_spmv<real>(ctx, mat, d_x, d_y);
    if (rc == 0 && cudaMemcpyAsync(h_y, d_y, sizeof(real) * (size_t)mat->M, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = -1;
    if (cudaStreamSynchronize(st) != cudaSuccess) rc = -1;
    if (rc == -1 && ctx->err.empty()) ctx->fail(-1, "nsp_spmv_amb_host: CUDA copy failed");
    return rc;
}

}  // namespace

extern "C" {

int nsp_csr2amb_s(nsp_context *ctx, int M, int N, int nnz, const int *d_rpt, const int *d_col, const float *d_val,
                  long long seg_size, int block_size, int autotune, const float *d_x, nsp_amb *out)
{
    return csr2amb<float>(ctx, M, N, nnz, d_rpt, d_col, d_val, seg_size, block_size, autotune, d_x, out);
}

int nsp_csr2amb_d(nsp_context *ctx, int M, int N, int nnz, const int *d_rpt, const int *d_col, const double *d_val,
                  long long seg_size, int block_size, int autotune, const double *d_x, nsp_amb *out)
{
    return csr2amb<double>(ctx, M, N, nnz, d_rpt, d_col, d_val, seg_size, block_size, autotune, d_x, out);
}

int nsp_amb_free(nsp_context *ctx, nsp_amb *mat)
{
    if (!ctx || !mat) return NSP_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    return free_arrays<float>(ctx, mat);
}

int nsp_spmv_amb_s(nsp_context *ctx, const nsp_amb *mat, const float *d_x, float *d_y)
{
    if (!ctx) return NSP_ERR_ARG;
    cudaSetDevice(ctx->device);
    return nsp::amb_spmv<float>(ctx, mat, d_x, d_y);
}

int nsp_spmv_amb_d(nsp_context *ctx, const nsp_amb *mat, const double *d_x, double *d_y)
{
    if (!ctx) return NSP_ERR_ARG;
    cudaSetDevice(ctx->device);
    return nsp::amb_spmv<double>(ctx, mat, d_x, d_y);
}

int nsp_memcpy_d2h(nsp_context *ctx, void *h_dst, const void *d_src, size_t bytes)
{
    if (!ctx || (bytes && (!h_dst || !d_src))) return NSP_ERR_ARG;
    cudaSetDevice(ctx->device);
    if (bytes == 0) return 0;
    NSP_CUDA_TRY(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    NSP_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int nsp_spmv_amb_host_s(nsp_context *ctx, const nsp_amb *mat, const float *h_x, float *h_y)
{
    return spmv_host<float>(ctx, mat, h_x, h_y);
}

int nsp_spmv_amb_host_d(nsp_context *ctx, const nsp_amb *mat, const double *h_x, double *h_y)
{
    return spmv_host<double>(ctx, mat, h_x, h_y);
}

}  // extern "C"
