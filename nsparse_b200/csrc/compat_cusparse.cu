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

void spgemm_cu_csr(sfCSR *a, sfCSR *b, sfCSR *c)
{
#ifdef FLOAT
    const cudaDataType dt = CUDA_R_32F;
#else
    const cudaDataType dt = CUDA_R_64F;
#endif
    cusparseHandle_t h;
    CHECK_CUSPARSE(cusparseCreate(&h));
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
    c->nnz = (int)nnz;
    checkCudaErrors(cudaMalloc((void **)&c->d_col, sizeof(int) * (size_t)(nnz ? nnz : 1)));
    checkCudaErrors(cudaMalloc((void **)&c->d_val, sizeof(real) * (size_t)(nnz ? nnz : 1)));
    CHECK_CUSPARSE(cusparseCsrSetPointers(C, c->d_rpt, c->d_col, c->d_val));
    CHECK_CUSPARSE(cusparseSpGEMM_copy(h, op, op, &alpha, A, B, &beta, C, dt, CUSPARSE_SPGEMM_DEFAULT, desc));

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
    cusparseSpGEMM_destroyDescr(desc);
    cusparseDestroySpMat(A);
    cusparseDestroySpMat(B);
    cusparseDestroySpMat(C);
    cusparseDestroy(h);
    cudaFree(b1);
    cudaFree(b2);
}
