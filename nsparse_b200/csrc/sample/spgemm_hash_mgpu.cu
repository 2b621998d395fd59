// spgemm_hash_mgpu_{s,d} <A.mtx> [ngpu]: the multi-GPU sibling of the reference's sample driver
// (cuda-c/src/sample/spgemm/spgemm_hash.cu:14-94; same flow, same printed line format, same self-check).
// C = A * A on `ngpu` GPUs of this box through spgemm_kernel_hash_mgpu (include/nsparse.h), timed like the
// reference (mean of SPGEMM_TRI_NUM - 1 calls after one warm-up, C released between calls), then EVERY GPU's
// copy of C is compared with the single-GPU product of spgemm_kernel_hash via check_spgemm_answer.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>

#include <cuda.h>
#include <helper_cuda.h>

#include <nsparse.h>

int main(int argc, char **argv)
{
    if (argc < 2) {
        fprintf(stderr, "usage: %s A.mtx [ngpu]\n", argv[0]);
        return 1;
    }
    int ngpu = 0;
    checkCudaErrors(cudaGetDeviceCount(&ngpu));
    if (argc > 2) ngpu = atoi(argv[2]);
    if (ngpu < 1 || ngpu > 8) {
        fprintf(stderr, "ngpu must be 1..8\n");
        return 1;
    }
    sfCSR a, b, c[8];
    init_csr_matrix_from_file(&a, argv[1]);
    init_csr_matrix_from_file(&b, argv[1]);

    /* single-GPU answer on GPU 0 (also gives the flop count) */
    checkCudaErrors(cudaSetDevice(0));
    csr_memcpy(&a);
    csr_memcpy(&b);
    long long int flop_count = 0;
    get_spgemm_flop(&a, &b, a.M, &flop_count);
    sfCSR ans;
    spgemm_kernel_hash(&a, &b, &ans);
    csr_memcpyDtH(&ans);
    release_csr(ans);
    release_csr(a);
    release_csr(b);

    double ave_msec = 0;
    for (int i = 0; i < SPGEMM_TRI_NUM; i++) {
        if (i > 0) release_csr_mgpu(c, ngpu);
        const auto t0 = std::chrono::steady_clock::now();
        spgemm_kernel_hash_mgpu(&a, &b, c, ngpu);       /* returns with every GPU idle */
        const double msec = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (i > 0) ave_msec += msec;
    }
    ave_msec /= SPGEMM_TRI_NUM - 1;
    const double flops = (double)flop_count / 1000 / 1000 / ave_msec;
    printf("SpGEMM using CSR format (Hash-based, %d GPUs, host CSR in, C on every GPU): %s, %f[GFLOPS], %f[ms]\n", ngpu,
           a.matrix_name, flops, ave_msec);
    printf("(nnz of A): %d =>\n(Num of intermediate products): %lld =>\n(nnz of C): %d\n", a.nnz, flop_count / 2, c[0].nnz);
    for (int g = 0; g < ngpu; ++g) {
        checkCudaErrors(cudaSetDevice(g));
        csr_memcpyDtH(&c[g]);
        printf("GPU %d: ", g);
        check_spgemm_answer(c[g], ans);
        release_cpu_csr(c[g]);
    }
    release_csr_mgpu(c, ngpu);
    release_cpu_csr(ans);
    release_cpu_csr(a);
    release_cpu_csr(b);
    return 0;
}
