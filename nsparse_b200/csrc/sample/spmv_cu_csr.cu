// cu_csr_{s,d} <A.mtx>: the cuSPARSE CSR SpMV comparison driver.  The reference's own
// cuda-c/src/sample/spmv/spmv_cu_csr.cu calls the legacy cusparse{S,D}csrmv directly (:50,55), which CUDA 12 no
// longer ships, so it cannot be compiled unchanged; this is the same flow and the same printed lines on
// sf_spmv_cu_csr (include/nsparse.h, generic cusparseSpMV underneath).
#include <stdio.h>
#include <stdlib.h>

#include <cuda.h>
#include <helper_cuda.h>
#include <cusparse.h>

#include <nsparse.h>

static void spmv_cu_csr(sfCSR *mat, real *x, real *y)
{
    real *d_x, *d_y;
    cudaEvent_t event[2];
    float exe_msec, ave_msec = 0, flops;
    cusparseHandle_t cusparseHandle = 0;
    cusparseMatDescr_t descr = 0;
    for (int i = 0; i < 2; i++) cudaEventCreate(&(event[i]));

    csr_memcpy(mat);
    checkCudaErrors(cudaMalloc((void **)&d_x, sizeof(real) * mat->N));
    checkCudaErrors(cudaMalloc((void **)&d_y, sizeof(real) * mat->M));
    checkCudaErrors(cudaMemcpy(d_x, x, sizeof(real) * mat->N, cudaMemcpyHostToDevice));

    cusparseCreate(&cusparseHandle);
    cusparseCreateMatDescr(&descr);
    cusparseSetMatType(descr, CUSPARSE_MATRIX_TYPE_GENERAL);
    cusparseSetMatIndexBase(descr, CUSPARSE_INDEX_BASE_ZERO);

    for (int i = 0; i < TRI_NUM; i++) {
        cudaEventRecord(event[0], 0);
        sf_spmv_cu_csr(d_y, mat, d_x, &cusparseHandle, &descr);
        cudaEventRecord(event[1], 0);
        cudaDeviceSynchronize();
        cudaEventElapsedTime(&exe_msec, event[0], event[1]);
        if (i > 0) ave_msec += exe_msec;
    }
    ave_msec /= TRI_NUM - 1;
    checkCudaErrors(cudaMemcpy(y, d_y, sizeof(real) * mat->M, cudaMemcpyDeviceToHost));
    flops = (float)(mat->nnz) * 2 / 1000 / 1000 / ave_msec;
    printf("SpMV using CSR format (cuSPARSE): %s, %f[GFLOPS], %f[ms]\n", mat->matrix_name, flops, ave_msec);

    cudaFree(d_x);
    cudaFree(d_y);
    release_csr(*mat);
    cusparseDestroyMatDescr(descr);
    cusparseDestroy(cusparseHandle);
    for (int i = 0; i < 2; i++) cudaEventDestroy(event[i]);
}

int main(int argc, char *argv[])
{
    if (argc < 2) {
        fprintf(stderr, "usage: %s A.mtx\n", argv[0]);
        return 1;
    }
    sfCSR mat;
    init_csr_matrix_from_file(&mat, argv[1]);
    real *x = (real *)malloc(sizeof(real) * mat.N);
    real *y = (real *)malloc(sizeof(real) * mat.M);
    init_vector(x, mat.N);
#ifdef sfDEBUG
    real *csr_y = (real *)malloc(sizeof(real) * mat.M);
    csr_kernel(csr_y, &mat, x);
#endif
    spmv_cu_csr(&mat, x, y);
#ifdef sfDEBUG
    ans_check(csr_y, y, mat.M);
    free(csr_y);
#endif
    free(x);
    free(y);
    release_cpu_csr(mat);
    return 0;
}
