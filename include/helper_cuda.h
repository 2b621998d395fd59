/*
 * helper_cuda.h -- minimal stand-in for the CUDA-samples header of the same name.
 * The nsparse sample drivers only use checkCudaErrors(); the CUDA samples are not
 * part of the toolkit any more, so nsparse-b200 ships this shim on its include path.
 */
#ifndef NSPARSE_B200_HELPER_CUDA_SHIM_H
#define NSPARSE_B200_HELPER_CUDA_SHIM_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime.h>

static inline void nsp_shim_check(cudaError_t err, const char *expr, const char *file, int line)
{
    if (err != cudaSuccess) {
        fprintf(stderr, "CUDA error at %s:%d code=%d (%s) \"%s\"\n", file, line, (int)err,
                cudaGetErrorName(err), expr);
        exit(EXIT_FAILURE);
    }
}
#define checkCudaErrors(expr) nsp_shim_check((expr), #expr, __FILE__, __LINE__)
#define getLastCudaError(msg) nsp_shim_check(cudaGetLastError(), msg, __FILE__, __LINE__)

#endif
