// nsparse-b200: multi-GPU hash SpGEMM from ONE process (C ABI nsp_mgpu_*, include/nsparse_b200.h).
//
// New relative to the reference (single GPU; modelled on its driver flow, spgemm_hash.cu:14-94): A (host CSR) is
// cut into `ngpu` contiguous row blocks of ~equal intermediate products, B is replicated, one host thread per
// GPU runs the single-GPU pipeline on its block, and every GPU ends up with the FULL C: the numeric kernels
// count finished tiles and each GPU's pusher kernel stores them into the other GPUs' arrays over NVLink peer
// memory (peer_push.cu).  The torch.distributed path of nsparse_b200/multi_gpu.py does the same with one
// process per GPU and CUDA IPC; here peer access inside the process makes every pointer directly usable.
#include "../../include/nsparse_b200.h"

#include <omp.h>

#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "context.h"

struct nsp_mgpu {
    int n = 0;
    std::vector<int> dev;
    std::vector<nsp_context *> ctx;
    int M = 0, K = 0, N = 0, val_bytes = 0;
    std::vector<int> cuts;           // n + 1 row cuts of A
    std::vector<long long> disp;     // n + 1 element displacements of the blocks in C
    long long total_ip = 0;
    struct Dev {
        int *a_rpt = nullptr, *a_col = nullptr, *b_rpt = nullptr, *b_col = nullptr;
        void *a_val = nullptr, *b_val = nullptr;
        long long *rpt_local = nullptr;
        size_t a_rows_cap = 0, a_nnz_cap = 0, b_rows_cap = 0, b_nnz_cap = 0;
        int val_bytes = 0;
        long long nnz = 0, ip = 0, a_nnz = 0;
        int rc = 0;
        double ms_symbolic = 0, ms_numeric = 0;
    };
    std::vector<Dev> d;
    bool symbolic_done = false;
    std::string err;
    int fail(int code, const std::string &m)
    {
        err = m;
        return code;
    }
};

namespace {

__global__ void rebase_rpt_kernel(const long long *__restrict__ local, int rows, long long add, long long *__restrict__ out,
                                  long long *__restrict__ last, long long total)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) out[i] = local[i] + add;
    if (i == 0) *last = total;
}

template <typename F>
void for_each_gpu(nsp_mgpu *mg, F &&f)
{
    std::vector<std::thread> th;
    for (int g = 0; g < mg->n; ++g)
        th.emplace_back([&, g]() {
            cudaSetDevice(mg->dev[g]);
            mg->d[g].rc = f(g);
        });
    for (auto &t : th) t.join();
}

int first_error(nsp_mgpu *mg, const char *what)
{
    for (int g = 0; g < mg->n; ++g)
        if (mg->d[g].rc != 0)
            return mg->fail(mg->d[g].rc, std::string(what) + " on GPU " + std::to_string(mg->dev[g]) + ": " +
                                             nsp_last_error(mg->ctx[g]));
    return 0;
}

#define MG_TRY(ctx, expr)                                                                      \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) return (ctx)->fail(-1, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

template <typename T>
int grow(nsp_context *ctx, T *&p, size_t &cap, size_t want)
{
    if (want <= cap && p) return 0;
    cudaFree(p);
    p = nullptr;
    cap = 0;
    MG_TRY(ctx, cudaMalloc((void **)&p, sizeof(T) * (want ? want : 1)));
    cap = want;
    return 0;
}

template <typename real>
int mgpu_symbolic(nsp_mgpu *mg, int M, int K, int N, const int *a_rpt, const int *a_col, const real *a_val, const int *b_rpt,
                  const int *b_col, const real *b_val, long long *h_nnz, long long *h_ip)
{
    if (!mg || M < 0 || K < 0 || N < 0 || !a_rpt || !b_rpt) return mg ? mg->fail(NSP_ERR_ARG, "nsp_mgpu_spgemm_symbolic: bad argument") : NSP_ERR_ARG;
    mg->symbolic_done = false;
    mg->M = M;
    mg->K = K;
    mg->N = N;
    mg->val_bytes = (int)sizeof(real);
    // ---- row cuts of ~equal intermediate products (get_spgemm_flop's quantity, kernel_spgemm_cu_csr.cu:18-33) ----
    std::vector<long long> pre((size_t)M + 1, 0);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < M; ++i) {
        long long s = 0;
        for (int j = a_rpt[i]; j < a_rpt[i + 1]; ++j) s += b_rpt[a_col[j] + 1] - b_rpt[a_col[j]];
        pre[(size_t)i + 1] = s;
    }
    for (int i = 0; i < M; ++i) pre[(size_t)i + 1] += pre[i];
    mg->total_ip = pre[M];
    mg->cuts.assign(mg->n + 1, M);
    mg->cuts[0] = 0;
    for (int p = 1; p < mg->n; ++p) {
        const long long target = mg->total_ip / mg->n * p;
        int lo = 0, hi = M;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (pre[mid] < target) lo = mid + 1; else hi = mid;
        }
        mg->cuts[p] = lo < mg->cuts[p - 1] ? mg->cuts[p - 1] : lo;
    }
    const int b_nnz = b_rpt[K];
    for_each_gpu(mg, [&](int g) -> int {
        nsp_context *ctx = mg->ctx[g];
        nsp_mgpu::Dev &d = mg->d[g];
        const auto t0 = std::chrono::steady_clock::now();
        const int r0 = mg->cuts[g], r1 = mg->cuts[g + 1], rows = r1 - r0;
        const int lo = a_rpt[r0], nnz = a_rpt[r1] - lo;
        if (d.val_bytes != (int)sizeof(real)) {        // precision switch: the value buffers are re-made
            cudaFree(d.a_val);
            cudaFree(d.b_val);
            cudaFree(d.a_col);
            cudaFree(d.b_col);
            d.a_val = d.b_val = nullptr;
            d.a_col = d.b_col = nullptr;
            d.a_nnz_cap = d.b_nnz_cap = 0;
            d.val_bytes = (int)sizeof(real);
        }
        size_t cap2 = d.a_rows_cap;
        if (grow(ctx, d.a_rpt, d.a_rows_cap, (size_t)rows + 1) != 0) return -1;
        if (grow(ctx, d.rpt_local, cap2, (size_t)rows + 1) != 0) return -1;
        if ((size_t)nnz > d.a_nnz_cap || !d.a_col) {
            cudaFree(d.a_col);
            cudaFree(d.a_val);
            d.a_col = nullptr;
            d.a_val = nullptr;
            MG_TRY(ctx, cudaMalloc((void **)&d.a_col, sizeof(int) * ((size_t)nnz + 1)));
            MG_TRY(ctx, cudaMalloc((void **)&d.a_val, sizeof(real) * ((size_t)nnz + 1)));
            d.a_nnz_cap = (size_t)nnz;
        }
        if (grow(ctx, d.b_rpt, d.b_rows_cap, (size_t)K + 1) != 0) return -1;
        if ((size_t)b_nnz > d.b_nnz_cap || !d.b_col) {
            cudaFree(d.b_col);
            cudaFree(d.b_val);
            d.b_col = nullptr;
            d.b_val = nullptr;
            MG_TRY(ctx, cudaMalloc((void **)&d.b_col, sizeof(int) * ((size_t)b_nnz + 1)));
            MG_TRY(ctx, cudaMalloc((void **)&d.b_val, sizeof(real) * ((size_t)b_nnz + 1)));
            d.b_nnz_cap = (size_t)b_nnz;
        }
        // the block's row pointer, rebased to 0
        std::vector<int> rp((size_t)rows + 1);
        for (int i = 0; i <= rows; ++i) rp[i] = a_rpt[r0 + i] - lo;
        MG_TRY(ctx, cudaMemcpy(d.a_rpt, rp.data(), sizeof(int) * ((size_t)rows + 1), cudaMemcpyHostToDevice));
        MG_TRY(ctx, cudaMemcpy(d.a_col, a_col + lo, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice));
        MG_TRY(ctx, cudaMemcpy(d.a_val, a_val + lo, sizeof(real) * (size_t)nnz, cudaMemcpyHostToDevice));
        MG_TRY(ctx, cudaMemcpy(d.b_rpt, b_rpt, sizeof(int) * ((size_t)K + 1), cudaMemcpyHostToDevice));
        MG_TRY(ctx, cudaMemcpy(d.b_col, b_col, sizeof(int) * (size_t)b_nnz, cudaMemcpyHostToDevice));
        MG_TRY(ctx, cudaMemcpy(d.b_val, b_val, sizeof(real) * (size_t)b_nnz, cudaMemcpyHostToDevice));
        d.nnz = d.ip = 0;
        d.a_nnz = nnz;
        const int rc = nsp::spgemm_symbolic(ctx, rows, K, N, d.a_rpt, d.a_col, d.b_rpt, d.b_col, d.rpt_local, &d.nnz, &d.ip);
        d.ms_symbolic = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return rc;
    });
    if (int rc = first_error(mg, "symbolic phase")) return rc;
    mg->disp.assign(mg->n + 1, 0);
    for (int g = 0; g < mg->n; ++g) mg->disp[g + 1] = mg->disp[g] + mg->d[g].nnz;
    if (h_nnz) *h_nnz = mg->disp[mg->n];
    if (h_ip) *h_ip = mg->total_ip;
    mg->symbolic_done = true;
    return 0;
}

template <typename real>
int mgpu_numeric(nsp_mgpu *mg, long long *const *c_rpt64, int *const *c_col, real *const *c_val)
{
    if (!mg || !c_rpt64 || !c_col || !c_val) return mg ? mg->fail(NSP_ERR_ARG, "nsp_mgpu_spgemm_numeric: bad argument") : NSP_ERR_ARG;
    if (!mg->symbolic_done || mg->val_bytes != (int)sizeof(real))
        return mg->fail(NSP_ERR_ARG, "nsp_mgpu_spgemm_numeric: call nsp_mgpu_spgemm_symbolic of the same precision first");
    const long long tot = mg->disp[mg->n];
    // Phase A on every GPU, THEN phase B: whatever allocates device memory or loads code happens before the first
    // pusher kernel starts spinning anywhere -- cudaMalloc on one GPU of a process with peer access enabled
    // synchronises with its peers, i.e. it would wait for their pushers, which wait for kernels queued behind it.
    for_each_gpu(mg, [&](int g) -> int {
        return nsp::spgemm_numeric_reserve<real>(mg->ctx[g], mg->N, mg->d[g].a_nnz, mg->d[g].nnz, mg->cuts[g + 1] - mg->cuts[g],
                                                 mg->n - 1);
    });
    if (int rc = first_error(mg, "numeric phase (reserve)")) return rc;
    for_each_gpu(mg, [&](int g) -> int {
        nsp_context *ctx = mg->ctx[g];
        nsp_mgpu::Dev &d = mg->d[g];
        const auto t0 = std::chrono::steady_clock::now();
        const int r0 = mg->cuts[g], rows = mg->cuts[g + 1] - r0;
        // the other GPUs' full arrays (peer access is enabled: plain pointers)
        void *pc[nsp::kMaxPeerOut], *pv[nsp::kMaxPeerOut], *pr[nsp::kMaxPeerOut];
        int np = 0;
        for (int q = 0; q < mg->n; ++q)
            if (q != g) {
                pc[np] = c_col[q];
                pv[np] = c_val[q];
                pr[np] = c_rpt64[q];
                ++np;
            }
        int rc = nsp_spgemm_set_peers(ctx, np, pc, pv, mg->disp[g]);
        if (rc == 0)
            rc = nsp::spgemm_numeric<real>(ctx, rows, mg->K, mg->N, d.a_rpt, d.a_col, (const real *)d.a_val, d.b_rpt, d.b_col,
                                           (const real *)d.b_val, d.rpt_local, c_col[g] + mg->disp[g], c_val[g] + mg->disp[g]);
        nsp_spgemm_set_peers(ctx, 0, nullptr, nullptr, 0);
        if (rc != 0) return rc;
        // row pointer: own rows rebased by the block's displacement, then to the peers
        rebase_rpt_kernel<<<(rows + 255) / 256 + 1, 256, 0, ctx->stream>>>(d.rpt_local, rows, mg->disp[g], c_rpt64[g] + r0,
                                                                           c_rpt64[g] + mg->M, tot);
        ctx->launches += 1;
        MG_TRY(ctx, cudaGetLastError());
        if (np > 0 && rows > 0) {
            rc = nsp_push_to_peers(ctx, np, pr, sizeof(long long) * (size_t)r0, c_rpt64[g] + r0, sizeof(long long) * (size_t)rows);
            if (rc != 0) return rc;
        }
        int err = 0;
        rc = nsp_spgemm_peers_status(ctx, &err);     // synchronises the GPU's stream: its block has left for the peers
        d.ms_numeric = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return rc;
    });
    return first_error(mg, "numeric phase");
}

}  // namespace

extern "C" {

int nsp_mgpu_create(nsp_mgpu **out, int ngpu, const int *devices)
{
    if (!out || ngpu < 1 || ngpu > nsp::kMaxPeerOut + 1) return NSP_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count < ngpu) {
        fprintf(stderr, "nsparse_b200: %d GPUs asked for, %d visible\n", ngpu, count);
        return NSP_ERR_CUDA;
    }
    nsp_mgpu *mg = new nsp_mgpu();
    mg->n = ngpu;
    mg->d.resize(ngpu);
    for (int g = 0; g < ngpu; ++g) mg->dev.push_back(devices ? devices[g] : g);
    int cur = 0;
    cudaGetDevice(&cur);
    for (int g = 0; g < ngpu; ++g) {
        nsp_context *c = nullptr;
        if (nsp_create(&c, mg->dev[g]) != 0) {
            for (auto *x : mg->ctx) nsp_destroy(x);
            delete mg;
            cudaSetDevice(cur);
            return NSP_ERR_CUDA;
        }
        mg->ctx.push_back(c);
    }
    for (int g = 0; g < ngpu; ++g) {
        cudaSetDevice(mg->dev[g]);
        for (int q = 0; q < ngpu; ++q) {
            if (q == g) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, mg->dev[g], mg->dev[q]);
            if (!can) {
                fprintf(stderr, "nsparse_b200: GPU %d cannot access GPU %d (no NVLink / peer access)\n", mg->dev[g], mg->dev[q]);
                for (auto *x : mg->ctx) nsp_destroy(x);
                delete mg;
                cudaSetDevice(cur);
                return NSP_ERR_CUDA;
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(mg->dev[q], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                fprintf(stderr, "nsparse_b200: cudaDeviceEnablePeerAccess(%d -> %d): %s\n", mg->dev[g], mg->dev[q], cudaGetErrorString(e));
                for (auto *x : mg->ctx) nsp_destroy(x);
                delete mg;
                cudaSetDevice(cur);
                return NSP_ERR_CUDA;
            }
            cudaGetLastError();
        }
    }
    cudaSetDevice(cur);
    *out = mg;
    return 0;
}

int nsp_mgpu_destroy(nsp_mgpu *mg)
{
    if (!mg) return 0;
    int cur = 0;
    cudaGetDevice(&cur);
    for (int g = 0; g < mg->n; ++g) {
        cudaSetDevice(mg->dev[g]);
        cudaDeviceSynchronize();
        nsp_mgpu::Dev &d = mg->d[g];
        cudaFree(d.a_rpt); cudaFree(d.a_col); cudaFree(d.a_val);
        cudaFree(d.b_rpt); cudaFree(d.b_col); cudaFree(d.b_val);
        cudaFree(d.rpt_local);
        nsp_destroy(mg->ctx[g]);
    }
    cudaSetDevice(cur);
    delete mg;
    return 0;
}

const char *nsp_mgpu_last_error(nsp_mgpu *mg) { return mg ? mg->err.c_str() : "null handle"; }
int nsp_mgpu_ngpu(nsp_mgpu *mg) { return mg ? mg->n : 0; }
nsp_context *nsp_mgpu_context(nsp_mgpu *mg, int g) { return (mg && g >= 0 && g < mg->n) ? mg->ctx[g] : nullptr; }

int nsp_mgpu_block(nsp_mgpu *mg, int g, int *device, int *row0, int *row1, long long *elem0, long long *nnz,
                   double *ms_symbolic, double *ms_numeric)
{
    if (!mg || g < 0 || g >= mg->n || !mg->symbolic_done) return NSP_ERR_ARG;
    if (device) *device = mg->dev[g];
    if (row0) *row0 = mg->cuts[g];
    if (row1) *row1 = mg->cuts[g + 1];
    if (elem0) *elem0 = mg->disp[g];
    if (nnz) *nnz = mg->d[g].nnz;
    if (ms_symbolic) *ms_symbolic = mg->d[g].ms_symbolic;
    if (ms_numeric) *ms_numeric = mg->d[g].ms_numeric;
    return 0;
}

int nsp_mgpu_spgemm_symbolic_s(nsp_mgpu *mg, int M, int K, int N, const int *h_a_rpt, const int *h_a_col, const float *h_a_val,
                               const int *h_b_rpt, const int *h_b_col, const float *h_b_val, long long *h_nnz_c, long long *h_intprod)
{
    return mgpu_symbolic<float>(mg, M, K, N, h_a_rpt, h_a_col, h_a_val, h_b_rpt, h_b_col, h_b_val, h_nnz_c, h_intprod);
}

int nsp_mgpu_spgemm_symbolic_d(nsp_mgpu *mg, int M, int K, int N, const int *h_a_rpt, const int *h_a_col, const double *h_a_val,
                               const int *h_b_rpt, const int *h_b_col, const double *h_b_val, long long *h_nnz_c, long long *h_intprod)
{
    return mgpu_symbolic<double>(mg, M, K, N, h_a_rpt, h_a_col, h_a_val, h_b_rpt, h_b_col, h_b_val, h_nnz_c, h_intprod);
}

int nsp_mgpu_spgemm_numeric_s(nsp_mgpu *mg, long long *const *d_c_rpt64, int *const *d_c_col, float *const *d_c_val)
{
    return mgpu_numeric<float>(mg, d_c_rpt64, d_c_col, d_c_val);
}

int nsp_mgpu_spgemm_numeric_d(nsp_mgpu *mg, long long *const *d_c_rpt64, int *const *d_c_col, double *const *d_c_val)
{
    return mgpu_numeric<double>(mg, d_c_rpt64, d_c_col, d_c_val);
}

}  // extern "C"
