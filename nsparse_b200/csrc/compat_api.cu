// nsparse-b200: the nsparse.h entry points (C++ linkage, one build per precision) on top of the
// core library.  The UNCHANGED sample drivers of the reference (cuda-c/src/sample/spgemm/
// spgemm_hash.cu, cuda-c/src/sample/spmv/spmv_amb.cu) link against this: same names, argument
// meaning, ownership rules, printed lines and error behaviour (abort) as cuda-c/src/nsparse.cu,
// kernel_spgemm_hash_*.cu:1035-1075, kernel_spgemm_cu_csr.cu:35-57, convert_amb.cu:835-929 and
// kernel_spmv_amb.cu:98-104.  Built with -DFLOAT into libnsparse_s.a and -DDOUBLE into libnsparse_d.a.
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <string>
#include <vector>

#include <helper_cuda.h>
#include <nsparse.h>

#include "../../include/nsparse_b200.h"

static_assert(sizeof(sfPlan) == 48, "sfPlan layout (cuda-c/inc/nsparse.h:50-59)");
static_assert(sizeof(sfCSR) == 72, "sfCSR layout (cuda-c/inc/nsparse.h:62-75)");
static_assert(sizeof(sfAMB) == 176, "sfAMB layout (cuda-c/inc/nsparse.h:78-107)");

#ifdef FLOAT
#define NSP_PREC(name) name##_s
#else
#define NSP_PREC(name) name##_d
#endif

namespace {

// one lazily created context per device of the process, legacy default stream: the drivers bracket the
// calls with events on stream 0 (spgemm_hash.cu:40-44, spmv_amb.cu:47-52)
struct DevState {
    nsp_context *ctx = nullptr;
    long long *d_rpt64 = nullptr;     // int64 row pointer of the core, narrowed into sfCSR's int d_rpt
    int rpt64_cap = -1;
};
DevState g_dev[64];

DevState &dev_state()
{
    int dev = 0;
    checkCudaErrors(cudaGetDevice(&dev));
    DevState &s = g_dev[dev & 63];
    if (!s.ctx) {
        if (nsp_create(&s.ctx, dev) != 0) {
            fprintf(stderr, "nsparse-b200: cannot create a context on device %d (needs a B200, sm_100)\n", dev);
            exit(EXIT_FAILURE);
        }
    }
    return s;
}

nsp_context *the_context() { return dev_state().ctx; }

void check(int rc, const char *what)
{
    if (rc != 0) {
        fprintf(stderr, "nsparse-b200: %s failed (%d): %s\n", what, rc, nsp_last_error(the_context()));
        exit(EXIT_FAILURE);
    }
}

}  // namespace

// ---- MatrixMarket reader -----------------------------------------------------------------------------
// Same observable behaviour as convert_file_csr (nsparse.cu:14-136): a first line containing
// "general" means the entries are taken as they are, otherwise off-diagonal entries are mirrored;
// '%' lines are skipped; 1-based indices; a missing value reads as 1; entries are appended to their
// rows in file order (mirror right after the original); nothing is sorted or merged.
void init_csr_matrix_from_file(sfCSR *mat, char *file_name)
{
    FILE *fp = fopen(file_name, "r");
    if (!fp) {
        printf("Cannot find file\n");
        exit(1);
    }
    fclose(fp);
    printf("Read mtx file: %s\n", file_name);
    // parallel parse, entry-for-entry the reference reader's result (csrc/mtx_reader.cpp)
    int M = 0, N = 0, nnz_max = 0;
    long long nnz = 0;
    void *val = nullptr;
    const int rc = nsp_read_mtx(file_name, sizeof(real) == 8, 0, &M, &N, &nnz, &nnz_max, &mat->rpt, &mat->col, &val);
    if (rc != 0) {
        printf("Cannot read the matrix (nsp_read_mtx: %d)\n", rc);
        exit(1);
    }
    mat->val = (real *)val;
    mat->M = M;
    mat->N = N;
    mat->nnz = (int)nnz;
    mat->nnz_max = nnz_max;
    mat->matrix_name = file_name;
    mat->d_rpt = nullptr;
    mat->d_col = nullptr;
    mat->d_val = nullptr;
}

// ---- copies, plans, vectors, release (nsparse.cu:146-235) ---------------------------------------------
void csr_memcpy(sfCSR *mat)
{
    const size_t nnz = mat->nnz > 0 ? mat->nnz : 1;
    checkCudaErrors(cudaMalloc((void **)&(mat->d_rpt), sizeof(int) * ((size_t)mat->M + 1)));
    checkCudaErrors(cudaMalloc((void **)&(mat->d_col), sizeof(int) * nnz));
    checkCudaErrors(cudaMalloc((void **)&(mat->d_val), sizeof(real) * nnz));
    checkCudaErrors(cudaMemcpy(mat->d_rpt, mat->rpt, sizeof(int) * ((size_t)mat->M + 1), cudaMemcpyHostToDevice));
    checkCudaErrors(cudaMemcpy(mat->d_col, mat->col, sizeof(int) * (size_t)mat->nnz, cudaMemcpyHostToDevice));
    checkCudaErrors(cudaMemcpy(mat->d_val, mat->val, sizeof(real) * (size_t)mat->nnz, cudaMemcpyHostToDevice));
}

void csr_memcpyDtH(sfCSR *mat)
{
    const size_t nnz = mat->nnz > 0 ? mat->nnz : 1;
    mat->rpt = (int *)malloc(sizeof(int) * ((size_t)mat->M + 1));
    mat->col = (int *)malloc(sizeof(int) * nnz);
    mat->val = (real *)malloc(sizeof(real) * nnz);
    checkCudaErrors(cudaMemcpy(mat->rpt, mat->d_rpt, sizeof(int) * ((size_t)mat->M + 1), cudaMemcpyDeviceToHost));
    checkCudaErrors(cudaMemcpy(mat->col, mat->d_col, sizeof(int) * (size_t)mat->nnz, cudaMemcpyDeviceToHost));
    checkCudaErrors(cudaMemcpy(mat->val, mat->d_val, sizeof(real) * (size_t)mat->nnz, cudaMemcpyDeviceToHost));
}

void init_plan(sfPlan *plan) { plan->isPlan = FALSE; }

void set_plan(sfPlan *plan, size_t seg_size, int block_size)
{
    plan->isPlan = TRUE;
    plan->seg_size = seg_size > USHORT_MAX ? USHORT_MAX : seg_size;
    plan->block_size = (block_size < 1 || block_size > MAX_BLOCK_SIZE) ? 1 : block_size;
}

void init_vector(real *x, int row)
{
    // the reference seeds with time(NULL); NSPARSE_SEED makes runs reproducible
    const char *s = getenv("NSPARSE_SEED");
    srand48(s ? atol(s) : (long)time(NULL));
    for (int i = 0; i < row; ++i) x[i] = (real)drand48();
}

void release_cpu_csr(sfCSR mat)
{
    free(mat.rpt);
    free(mat.col);
    free(mat.val);
}

void release_csr(sfCSR mat)
{
    cudaFree(mat.d_rpt);
    cudaFree(mat.d_col);
    cudaFree(mat.d_val);
}

void release_cpu_amb(sfAMB mat)
{
    free(mat.cs);
    free(mat.cl);
    free(mat.sellcs_val);
    free(mat.sellcs_col);
    free(mat.s_write_permutation);
    free(mat.s_write_permutation_offset);
}

void release_amb(sfAMB mat)
{
    // the seven arrays (cudaFree, like nsparse.cu:226-235) and the SpMV write plan the library keeps for them
    nsp_amb core;
    memset(&core, 0, sizeof(core));
    core.d_cs = mat.d_cs;
    core.d_cl = mat.d_cl;
    core.d_sellcs_col = mat.d_sellcs_col;
    core.d_sellcs_val = mat.d_sellcs_val;
    core.d_s_write_permutation = mat.d_s_write_permutation;
    core.d_s_write_permutation_offset = mat.d_s_write_permutation_offset;
    core.d_write_permutation = mat.d_write_permutation;
    nsp_amb_free(the_context(), &core);
}

// ---- CPU SpMV and the two comparators (nsparse.cu:240-353) ----------------------------------------------
void csr_kernel(real *y, sfCSR *cpu_mat, real *x)
{
    const int *rpt = cpu_mat->rpt, *col = cpu_mat->col;
    const real *val = cpu_mat->val;
    for (int i = 0; i < cpu_mat->M; ++i) {
        real ans = 0;
        for (int j = rpt[i]; j < rpt[i + 1]; ++j) ans += val[j] * x[col[j]];
        y[i] = ans;
    }
}

static real tolerance_scale()
{
#ifdef FLOAT
    return (real)1000;
#else
    return (real)1000 * 1000;
#endif
}

void ans_check(real *csr_ans, real *ans_vec, int N)
{
    int fails = 0;
    const real scale = tolerance_scale();
    for (int i = 0; i < N && fails < 10; ++i) {
        real delta = ans_vec[i] - csr_ans[i], base = ans_vec[i];
        if (delta < 0) delta = -delta;
        if (base < 0) base = -base;
        if (delta * 100 * scale > base) {
            printf("i=%d, ans=%e, csr=%e, delta=%e\n", i, ans_vec[i], csr_ans[i], delta);
            ++fails;
        }
    }
    printf(fails ? "Calculation Result is Incorrect\n" : "Calculation Result is Correct\n");
}

void check_spgemm_answer(sfCSR c, sfCSR ans)
{
    if (c.nnz != ans.nnz) {
        printf("nnz is not correct: %d (correct), %d (incorrect)\n", ans.nnz, c.nnz);
        return;
    }
    for (int i = 0; i < c.M + 1; ++i)
        if (c.rpt[i] != ans.rpt[i]) {
            printf("rpt[%d] is not correct: %d (correct),%d (incorrect)\n", i, ans.rpt[i], c.rpt[i]);
            return;
        }
    for (int i = 0; i < c.nnz; ++i)
        if (c.col[i] != ans.col[i]) {
            printf("col[%d] is not correct: %d (correct), %d (incorrect)\n", i, ans.col[i], c.col[i]);
            return;
        }
    int fails = 0;
    const real scale = tolerance_scale();
    for (int i = 0; i < c.nnz && fails < 10; ++i) {
        real delta = ans.val[i] - c.val[i], base = ans.val[i];
        if (delta < 0) delta = -delta;
        if (base < 0) base = -base;
        if (delta * 1000 * scale > base) {
            printf("val[%d]: ans=%e, c=%e, delta=%e\n", i, ans.val[i], c.val[i], delta);
            ++fails;
        }
    }
    printf(fails ? "Calculation Result is Incorrect\n" : "Calculation Result is Correct\n");
}

// ---- hash SpGEMM ---------------------------------------------------------------------------------------
void get_spgemm_flop(sfCSR *a, sfCSR *b, int M, long long int *flop)
{
    long long f = 0;
    check(nsp_spgemm_flop(the_context(), M, a->d_rpt, a->d_col, b->d_rpt, &f), "get_spgemm_flop");
    *flop = f;
}

// C = A * B.  Allocates c->d_rpt / d_col / d_val with cudaMalloc (the caller frees them with
// release_csr), sets c->M, c->N, c->nnz, and returns with the device idle, like
// kernel_spgemm_hash_*.cu:1035-1075.  The 64-bit row pointer of the core is narrowed to sfCSR's
// int; a product with more than INT_MAX entries aborts with a message (the reference wraps).
void spgemm_kernel_hash(sfCSR *a, sfCSR *b, sfCSR *c)
{
    DevState &ds = dev_state();
    nsp_context *ctx = ds.ctx;
    const int M = a->M, K = a->N, N = b->N;
    c->M = M;
    c->N = N;
    if (M > ds.rpt64_cap) {
        cudaFree(ds.d_rpt64);
        checkCudaErrors(cudaMalloc((void **)&ds.d_rpt64, sizeof(long long) * ((size_t)M + 1)));
        ds.rpt64_cap = M;
    }
    long long *d_rpt64 = ds.d_rpt64;
    long long nnz = 0, ip = 0;
    check(nsp_spgemm_symbolic(ctx, M, K, N, a->d_rpt, a->d_col, b->d_rpt, b->d_col, d_rpt64, &nnz, &ip),
          "spgemm_kernel_hash (symbolic)");
    if (nnz > (long long)INT_MAX) {
        fprintf(stderr, "nsparse-b200: nnz(C) = %lld does not fit sfCSR's int nnz; use nsp_spgemm_symbolic/"
                        "nsp_spgemm_numeric_* (64-bit row pointer) for this product\n", nnz);
        exit(EXIT_FAILURE);
    }
    c->nnz = (int)nnz;
    c->nnz_max = 0;
    checkCudaErrors(cudaMalloc((void **)&(c->d_rpt), sizeof(int) * ((size_t)M + 1)));
    checkCudaErrors(cudaMalloc((void **)&(c->d_col), sizeof(int) * (size_t)(nnz > 0 ? nnz : 1)));
    checkCudaErrors(cudaMalloc((void **)&(c->d_val), sizeof(real) * (size_t)(nnz > 0 ? nnz : 1)));
    check(nsp_rpt64_to_rpt32(ctx, M, d_rpt64, nnz, c->d_rpt), "spgemm_kernel_hash (row pointer)");
    check(NSP_PREC(nsp_spgemm_numeric)(ctx, M, K, N, a->d_rpt, a->d_col, a->d_val, b->d_rpt, b->d_col, b->d_val,
                                       d_rpt64, c->d_col, c->d_val),
          "spgemm_kernel_hash (numeric)");
    check(nsp_sync(ctx), "spgemm_kernel_hash (sync)");
}

// C = A * B on `ngpu` GPUs of this process (extension; the reference is single GPU).  a, b: HOST CSR (rpt / col /
// val as init_csr_matrix_from_file leaves them).  c: array of ngpu sfCSR; on return c[g].d_rpt / d_col / d_val
// hold the FULL product on GPU g (cudaMalloc'ed there, released with release_csr_mgpu), c[g].M / N / nnz are set,
// and every GPU is idle.  Rows are cut by equal intermediate products, B is replicated, the row blocks of C are
// gathered over NVLink peer memory while they are computed (nsp_mgpu_*, include/nsparse_b200.h).
void spgemm_kernel_hash_mgpu(sfCSR *a, sfCSR *b, sfCSR *c, int ngpu)
{
    static nsp_mgpu *mg = nullptr;
    if (mg && nsp_mgpu_ngpu(mg) != ngpu) {
        nsp_mgpu_destroy(mg);
        mg = nullptr;
    }
    if (!mg && nsp_mgpu_create(&mg, ngpu, nullptr) != 0) {
        fprintf(stderr, "nsparse-b200: cannot set up %d GPUs with peer access\n", ngpu);
        exit(EXIT_FAILURE);
    }
    auto mcheck = [&](int rc, const char *what) {
        if (rc != 0) {
            fprintf(stderr, "nsparse-b200: %s failed (%d): %s\n", what, rc, nsp_mgpu_last_error(mg));
            exit(EXIT_FAILURE);
        }
    };
    int cur = 0;
    checkCudaErrors(cudaGetDevice(&cur));
    const int M = a->M, K = a->N, N = b->N;
    long long nnz = 0, ip = 0;
    mcheck(NSP_PREC(nsp_mgpu_spgemm_symbolic)(mg, M, K, N, a->rpt, a->col, a->val, b->rpt, b->col, b->val, &nnz, &ip),
           "spgemm_kernel_hash_mgpu (symbolic)");
    if (nnz > (long long)INT_MAX) {
        fprintf(stderr, "nsparse-b200: nnz(C) = %lld does not fit sfCSR's int nnz; use nsp_mgpu_spgemm_* (64-bit row pointer)\n", nnz);
        exit(EXIT_FAILURE);
    }
    long long *rpt64[8];
    int *col[8];
    real *val[8];
    for (int g = 0; g < ngpu; ++g) {
        int dev = g;
        nsp_mgpu_block(mg, g, &dev, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        checkCudaErrors(cudaSetDevice(dev));
        c[g].M = M;
        c[g].N = N;
        c[g].nnz = (int)nnz;
        c[g].nnz_max = 0;
        c[g].rpt = nullptr;
        c[g].col = nullptr;
        c[g].val = nullptr;
        checkCudaErrors(cudaMalloc((void **)&rpt64[g], sizeof(long long) * ((size_t)M + 1)));
        checkCudaErrors(cudaMalloc((void **)&(c[g].d_rpt), sizeof(int) * ((size_t)M + 1)));
        checkCudaErrors(cudaMalloc((void **)&(c[g].d_col), sizeof(int) * (size_t)(nnz > 0 ? nnz : 1)));
        checkCudaErrors(cudaMalloc((void **)&(c[g].d_val), sizeof(real) * (size_t)(nnz > 0 ? nnz : 1)));
        col[g] = c[g].d_col;
        val[g] = c[g].d_val;
    }
    mcheck(NSP_PREC(nsp_mgpu_spgemm_numeric)(mg, rpt64, col, val), "spgemm_kernel_hash_mgpu (numeric)");
    for (int g = 0; g < ngpu; ++g) {
        int dev = g;
        nsp_mgpu_block(mg, g, &dev, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
        checkCudaErrors(cudaSetDevice(dev));
        nsp_context *ctx = nsp_mgpu_context(mg, g);
        if (nsp_rpt64_to_rpt32(ctx, M, rpt64[g], nnz, c[g].d_rpt) != 0 || nsp_sync(ctx) != 0) {
            fprintf(stderr, "nsparse-b200: spgemm_kernel_hash_mgpu (row pointer) failed: %s\n", nsp_last_error(ctx));
            exit(EXIT_FAILURE);
        }
        checkCudaErrors(cudaFree(rpt64[g]));
    }
    checkCudaErrors(cudaSetDevice(cur));
}

// release_csr for the array spgemm_kernel_hash_mgpu filled (c[g] lives on GPU g)
void release_csr_mgpu(sfCSR *c, int ngpu)
{
    int cur = 0;
    checkCudaErrors(cudaGetDevice(&cur));
    for (int g = 0; g < ngpu; ++g) {
        checkCudaErrors(cudaSetDevice(g));
        release_csr(c[g]);
    }
    checkCudaErrors(cudaSetDevice(cur));
}

// ---- AMB SpMV ------------------------------------------------------------------------------------------
static void to_core(const sfAMB *m, const sfPlan *plan, nsp_amb *o)
{
    memset(o, 0, sizeof(*o));
    o->d_cs = m->d_cs;
    o->d_cl = m->d_cl;
    o->d_sellcs_col = m->d_sellcs_col;
    o->d_sellcs_val = m->d_sellcs_val;
    o->d_s_write_permutation = m->d_s_write_permutation;
    o->d_s_write_permutation_offset = m->d_s_write_permutation_offset;
    o->d_write_permutation = m->d_write_permutation;
    o->block_size = m->block_size;
    o->nnz = m->nnz;
    o->M = m->M;
    o->N = m->N;
    o->pad_M = m->pad_M;
    o->chunk = m->chunk;
    o->SIGMA = m->SIGMA;
    o->c_size = m->c_size;
    o->seg_size = (long long)m->seg_size;
    o->seg_num = (long long)m->seg_num;
    o->thread_block = plan ? (long long)plan->thread_block : 256;
    o->thread_grid = plan ? (long long)plan->thread_grid : 0;
}

// plan->isPlan == TRUE: build with plan->seg_size / plan->block_size; FALSE: choose them (the
// reference's footprint model by default, the reference's timing search with NSPARSE_AMB_AUTOTUNE=1)
// and write them back with isPlan = TRUE (convert_amb.cu:835-929).
void sf_csr2amb(sfAMB *mat, sfCSR *csr_mat, real *d_x, sfPlan *plan)
{
    nsp_context *ctx = the_context();
    long long seg = 0;
    int bs = 0;
    if (plan->isPlan == TRUE) {
        seg = (long long)plan->seg_size;
        bs = plan->block_size;
    }
    const char *at = getenv("NSPARSE_AMB_AUTOTUNE");
    const int autotune = (plan->isPlan != TRUE && at && atoi(at) != 0) ? 1 : 0;
    nsp_amb o;
    check(NSP_PREC(nsp_csr2amb)(ctx, csr_mat->M, csr_mat->N, csr_mat->nnz, csr_mat->d_rpt, csr_mat->d_col,
                                csr_mat->d_val, seg, bs, autotune, d_x, &o),
          "sf_csr2amb");
    memset(mat, 0, sizeof(*mat));
    mat->d_cs = o.d_cs;
    mat->d_cl = o.d_cl;
    mat->d_sellcs_col = o.d_sellcs_col;
    mat->d_sellcs_val = (real *)o.d_sellcs_val;
    mat->d_s_write_permutation = o.d_s_write_permutation;
    mat->d_s_write_permutation_offset = o.d_s_write_permutation_offset;
    mat->d_write_permutation = o.d_write_permutation;
    mat->block_size = o.block_size;
    mat->nnz = o.nnz;
    mat->M = o.M;
    mat->N = o.N;
    mat->pad_M = o.pad_M;
    mat->chunk = o.chunk;
    mat->SIGMA = o.SIGMA;
    mat->group_num_col = (int)o.seg_num;
    mat->nnz_max = csr_mat->nnz_max;
    mat->c_size = o.c_size;
    mat->seg_size = (size_t)o.seg_size;
    mat->seg_num = (size_t)o.seg_num;
    mat->matrix_name = csr_mat->matrix_name;
    plan->isPlan = TRUE;
    plan->SIGMA = o.SIGMA;
    plan->seg_size = (size_t)o.seg_size;
    plan->seg_num = (size_t)o.seg_num;
    plan->block_size = o.block_size;
    plan->thread_block = (size_t)o.thread_block;
    plan->thread_grid = (size_t)o.thread_grid;
}

void sf_spmv_amb(real *d_y, sfAMB *mat, real *d_x, sfPlan *plan)
{
    nsp_context *ctx = the_context();
    nsp_amb o;
    to_core(mat, plan, &o);
    check(NSP_PREC(nsp_spmv_amb)(ctx, &o, d_x, d_y), "sf_spmv_amb");
    check(nsp_sync(ctx), "sf_spmv_amb (sync)");
}
