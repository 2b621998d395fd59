// nsparse-b200: spgemm_cu_csr -- the comparison answer the sample driver checks spgemm_kernel_hash
// against (cuda-c/src/sample/spgemm/spgemm_hash.cu:60-68).  The reference computes it with the
// legacy cusparse{S,D}csrgemm (kernel_spgemm_cu_csr.cu:59-203), which CUDA 12 no longer has; this
// is the same role on the generic cusparseSpGEMM API, followed by a per-row column sort so that
// check_spgemm_answer (exact rpt / col comparison) applies.  Result is left on the HOST in
// c->rpt / col / val, as the reference does (:181-197).
#include <stdio.h>
#include <stdlib.h>

#include <cusparse.h>
#include <helper_cuda.h>
#include <nsparse.h>

#define CHECK_CUSPARSE(expr)                                                                       \
    do {                                                                                           \
        cusparseStatus_t _s = (expr);                                                              \
        if (_s != CUSPARSE_STATUS_SUCCESS) {                                                       \
            fprintf(stderr, "cuSPARSE error %d (%s) at %s:%d\n", (int)_s, cusparseGetErrorString(_s), \
                    __FILE__, __LINE__);                                                           \
            exit(EXIT_FAILURE);                                                                    \
        }                                                                                          \
    } while (0)

#ifdef FLOAT
static const cudaDataType dt = CUDA_R_32F;
#else
static const cudaDataType dt = CUDA_R_64F;
#endif

// spgemm_kernel_cu_csr (kernel_spgemm_cu_csr.cu:59-160): C = A * B with cuSPARSE, result left ON THE DEVICE in
// c->d_rpt / d_col / d_val (cudaMalloc'ed here, released by the caller with release_csr), c->M / N / nnz set.  Same
// signature as the reference so that its UNCHANGED driver cuda-c/src/sample/spgemm/spgemm_cu_csr.cu builds against
// this library (bin/spgemm_cu_csr_{s,d}); the legacy descriptors / operations it passes are accepted and ignored
// (general matrices, no transposition -- what the driver sets), the work is done by the generic cusparseSpGEMM API.
void spgemm_kernel_cu_csr(sfCSR *a, sfCSR *b, sfCSR *c, cusparseHandle_t *cusparseHandle, cusparseOperation_t *trans_a,
                          cusparseOperation_t *trans_b, cusparseMatDescr_t *descr_a, cusparseMatDescr_t *descr_b)
{
    (void)trans_a;
    (void)trans_b;
    (void)descr_a;
    (void)descr_b;
    cusparseHandle_t h = *cusparseHandle;
    c->M = a->M;
    c->N = b->N;
    cusparseSpMatDescr_t A, B, C;
    CHECK_CUSPARSE(cusparseCreateCsr(&A, a->M, a->N, a->nnz, a->d_rpt, a->d_col, a->d_val, CUSPARSE_INDEX_32I,
                                     CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, dt));
    CHECK_CUSPARSE(cusparseCreateCsr(&B, b->M, b->N, b->nnz, b->d_rpt, b->d_col, b->d_val, CUSPARSE_INDEX_32I,
                                     CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, dt));
    checkCudaErrors(cudaMalloc((void **)&c->d_rpt, sizeof(int) * ((size_t)c->M + 1)));
    CHECK_CUSPARSE(cusparseCreateCsr(&C, c->M, c->N, 0, c->d_rpt, nullptr, nullptr, CUSPARSE_INDEX_32I,
                                     CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, dt));
    const real alpha = (real)1, beta = (real)0;
    cusparseSpGEMMDescr_t desc;
    CHECK_CUSPARSE(cusparseSpGEMM_createDescr(&desc));
    const cusparseOperation_t op = CUSPARSE_OPERATION_NON_TRANSPOSE;
    size_t s1 = 0, s2 = 0;
    void *b1 = nullptr, *b2 = nullptr;
    CHECK_CUSPARSE(cusparseSpGEMM_workEstimation(h, op, op, &alpha, A, B, &beta, C, dt, CUSPARSE_SPGEMM_DEFAULT, desc,
                                                 &s1, nullptr));
    checkCudaErrors(cudaMalloc(&b1, s1 ? s1 : 1));
    CHECK_CUSPARSE(cusparseSpGEMM_workEstimation(h, op, op, &alpha, A, B, &beta, C, dt, CUSPARSE_SPGEMM_DEFAULT, desc,
                                                 &s1, b1));
    CHECK_CUSPARSE(cusparseSpGEMM_compute(h, op, op, &alpha, A, B, &beta, C, dt, CUSPARSE_SPGEMM_DEFAULT, desc, &s2,
                                          nullptr));
    checkCudaErrors(cudaMalloc(&b2, s2 ? s2 : 1));
    CHECK_CUSPARSE(cusparseSpGEMM_compute(h, op, op, &alpha, A, B, &beta, C, dt, CUSPARSE_SPGEMM_DEFAULT, desc, &s2,
                                          b2));
    int64_t rows = 0, cols = 0, nnz = 0;
    CHECK_CUSPARSE(cusparseSpMatGetSize(C, &rows, &cols, &nnz));
    if (nnz > 0x7fffffffll) {
        fprintf(stderr, "spgemm_kernel_cu_csr: nnz(C) = %lld does not fit sfCSR\n", (long long)nnz);
        exit(EXIT_FAILURE);
    }
    c->nnz = (int)nnz;
    checkCudaErrors(cudaMalloc((void **)&c->d_col, sizeof(int) * (size_t)(nnz ? nnz : 1)));
    checkCudaErrors(cudaMalloc((void **)&c->d_val, sizeof(real) * (size_t)(nnz ? nnz : 1)));
    CHECK_CUSPARSE(cusparseCsrSetPointers(C, c->d_rpt, c->d_col, c->d_val));
    CHECK_CUSPARSE(cusparseSpGEMM_copy(h, op, op, &alpha, A, B, &beta, C, dt, CUSPARSE_SPGEMM_DEFAULT, desc));
    cusparseSpGEMM_destroyDescr(desc);
    cusparseDestroySpMat(A);
    cusparseDestroySpMat(B);
    cusparseDestroySpMat(C);
    cudaFree(b1);
    cudaFree(b2);
    checkCudaErrors(cudaDeviceSynchronize());
}

// spgemm_cu_csr (kernel_spgemm_cu_csr.cu:162-203): the comparison answer of the hash driver's self-check, rows
// sorted by column, left on the HOST.
void spgemm_cu_csr(sfCSR *a, sfCSR *b, sfCSR *c)
{
    cusparseHandle_t h;
    CHECK_CUSPARSE(cusparseCreate(&h));
    cusparseOperation_t op = CUSPARSE_OPERATION_NON_TRANSPOSE;
    spgemm_kernel_cu_csr(a, b, c, &h, &op, &op, nullptr, nullptr);
    const long long nnz = c->nnz;
    // sort every row by column (check_spgemm_answer compares col[] element-wise)
    if (nnz > 0) {
        size_t sb = 0;
        void *buf = nullptr;
        int *perm = nullptr;
        real *sorted = nullptr;
        cusparseMatDescr_t md;
        CHECK_CUSPARSE(cusparseCreateMatDescr(&md));
        CHECK_CUSPARSE(cusparseXcsrsort_bufferSizeExt(h, c->M, c->N, (int)nnz, c->d_rpt, c->d_col, &sb));
        checkCudaErrors(cudaMalloc(&buf, sb ? sb : 1));
        checkCudaErrors(cudaMalloc((void **)&perm, sizeof(int) * (size_t)nnz));
        checkCudaErrors(cudaMalloc((void **)&sorted, sizeof(real) * (size_t)nnz));
        CHECK_CUSPARSE(cusparseCreateIdentityPermutation(h, (int)nnz, perm));
        CHECK_CUSPARSE(cusparseXcsrsort(h, c->M, c->N, (int)nnz, md, c->d_rpt, c->d_col, perm, buf));
        cusparseSpVecDescr_t vx;
        cusparseDnVecDescr_t vy;
        CHECK_CUSPARSE(cusparseCreateSpVec(&vx, nnz, nnz, perm, sorted, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, dt));
        CHECK_CUSPARSE(cusparseCreateDnVec(&vy, nnz, c->d_val, dt));
        CHECK_CUSPARSE(cusparseGather(h, vy, vx));
        checkCudaErrors(cudaMemcpy(c->d_val, sorted, sizeof(real) * (size_t)nnz, cudaMemcpyDeviceToDevice));
        cusparseDestroySpVec(vx);
        cusparseDestroyDnVec(vy);
        cusparseDestroyMatDescr(md);
        cudaFree(buf);
        cudaFree(perm);
        cudaFree(sorted);
    }
    csr_memcpyDtH(c);
    release_csr(*c);
    cusparseDestroy(h);
}

// sf_spmv_cu_csr (kernel_spmv_cu_csr.cu:9-31): y = A x with cuSPARSE; the reference calls the legacy
// cusparse{S,D}csrmv, which CUDA 12 no longer has -- same signature, generic cusparseSpMV underneath.  The SpMV
// buffer is kept between calls (the comparison driver times 100 of them).
void sf_spmv_cu_csr(real *d_y, sfCSR *mat, real *d_x, cusparseHandle_t *cusparseHandle, cusparseMatDescr_t *descr)
{
    (void)descr;
    static void *buf = nullptr;
    static size_t buf_bytes = 0;
    const real alpha = (real)1, beta = (real)0;
    cusparseSpMatDescr_t A;
    cusparseDnVecDescr_t x, y;
    CHECK_CUSPARSE(cusparseCreateCsr(&A, mat->M, mat->N, mat->nnz, mat->d_rpt, mat->d_col, mat->d_val, CUSPARSE_INDEX_32I,
                                     CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, dt));
    CHECK_CUSPARSE(cusparseCreateDnVec(&x, mat->N, d_x, dt));
    CHECK_CUSPARSE(cusparseCreateDnVec(&y, mat->M, d_y, dt));
    size_t need = 0;
    CHECK_CUSPARSE(cusparseSpMV_bufferSize(*cusparseHandle, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, A, x, &beta, y, dt,
                                           CUSPARSE_SPMV_ALG_DEFAULT, &need));
    if (need > buf_bytes) {
        cudaFree(buf);
        checkCudaErrors(cudaMalloc(&buf, need));
        buf_bytes = need;
    }
    CHECK_CUSPARSE(cusparseSpMV(*cusparseHandle, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, A, x, &beta, y, dt,
                                CUSPARSE_SPMV_ALG_DEFAULT, buf));
    cusparseDestroySpMat(A);
    cusparseDestroyDnVec(x);
    cusparseDestroyDnVec(y);
    checkCudaErrors(cudaDeviceSynchronize());
}
