// Test infrastructure: runs the REFERENCE's own CSR->AMB conversion and AMB SpMV on a GPU and dumps
// every array of sfAMB, so oracle/amb.py (and through it the product) can be pinned to what the
// reference really produces.  Linked against the reference's cuda-c sources compiled where they lie
// (oracle/Makefile target ref_gpu; `__shfl_xor(` -> `__shfl_xor_sync(0xffffffffu, ` on a temporary
// copy because the originals do not compile for sm_70+).  Never part of the product.
//
//   dump_amb <in.mtx | in.csrbin> <seg_size> <block_size> <out.bin | -> [reps]
//
// in.csrbin (any name not ending in .mtx): int32 {M, N, nnz, sizeof(real)}, rpt int32[M+1], col int32[nnz],
// val real[nnz].  With reps > 0 the SpMV is also timed with the reference's protocol (spmv_amb.cu:45-62: mean
// of `reps` calls after one warm-up, cudaEvents on stream 0) and one JSON line is printed; "-" skips the dump.
//
// out.bin: int32 header {M, N, nnz_csr, pad_M, c_size, nnz_amb, block_size, seg_size, seg_num, sizeof(real)}
//          then rpt, col, val (CSR as read), cs, cl, sellcs_col, sellcs_val, s_write_permutation,
//          s_write_permutation_offset, write_permutation, x (N reals), y (M reals)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime.h>
#include <nsparse.h>

static void put(FILE *f, const void *p, size_t n)
{
    if (n && fwrite(p, 1, n, f) != n) {
        perror("fwrite");
        exit(2);
    }
}

template <typename T>
static void put_dev(FILE *f, const T *d, size_t count)
{
    T *h = (T *)malloc(sizeof(T) * (count ? count : 1));
    cudaMemcpy(h, d, sizeof(T) * count, cudaMemcpyDeviceToHost);
    put(f, h, sizeof(T) * count);
    free(h);
}

int main(int argc, char **argv)
{
    if (argc < 5) {
        fprintf(stderr, "usage: %s in.mtx seg_size block_size out.bin\n", argv[0]);
        return 1;
    }
    sfCSR mat;
    sfAMB amb;
    sfPlan plan;
    const size_t alen = strlen(argv[1]);
    if (alen >= 4 && !strcmp(argv[1] + alen - 4, ".mtx")) {
        init_csr_matrix_from_file(&mat, argv[1]);
    } else {
        FILE *fi = fopen(argv[1], "rb");
        int h[4];
        if (!fi || fread(h, sizeof(int), 4, fi) != 4 || h[3] != (int)sizeof(real)) {
            fprintf(stderr, "%s: cannot read raw CSR (value size must be %d)\n", argv[1], (int)sizeof(real));
            return 2;
        }
        mat.M = h[0];
        mat.N = h[1];
        mat.nnz = h[2];
        mat.nnz_max = 0;
        mat.matrix_name = argv[1];
        mat.rpt = (int *)malloc(sizeof(int) * (mat.M + 1));
        mat.col = (int *)malloc(sizeof(int) * (mat.nnz ? mat.nnz : 1));
        mat.val = (real *)malloc(sizeof(real) * (mat.nnz ? mat.nnz : 1));
        if (fread(mat.rpt, sizeof(int), mat.M + 1, fi) != (size_t)mat.M + 1 ||
            fread(mat.col, sizeof(int), mat.nnz, fi) != (size_t)mat.nnz ||
            fread(mat.val, sizeof(real), mat.nnz, fi) != (size_t)mat.nnz) {
            fprintf(stderr, "%s: short file\n", argv[1]);
            return 2;
        }
        fclose(fi);
        for (int i = 0; i < mat.M; ++i)
            if (mat.rpt[i + 1] - mat.rpt[i] > mat.nnz_max) mat.nnz_max = mat.rpt[i + 1] - mat.rpt[i];
    }
    csr_memcpy(&mat);
    init_plan(&plan);
    set_plan(&plan, (size_t)atol(argv[2]), atoi(argv[3]));
    real *x = (real *)malloc(sizeof(real) * mat.N);
    for (int i = 0; i < mat.N; ++i) x[i] = (real)((i * 37 + 11) % 101) / (real)101 + (real)0.25;
    real *d_x, *d_y;
    cudaMalloc((void **)&d_x, sizeof(real) * (mat.N + MAX_BLOCK_SIZE));
    cudaMalloc((void **)&d_y, sizeof(real) * (mat.M + WARP));
    cudaMemset(d_x, 0, sizeof(real) * (mat.N + MAX_BLOCK_SIZE));
    cudaMemcpy(d_x, x, sizeof(real) * mat.N, cudaMemcpyHostToDevice);
    sf_csr2amb(&amb, &mat, d_x, &plan);
    sf_spmv_amb(d_y, &amb, d_x, &plan);
    cudaDeviceSynchronize();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e));
        return 3;
    }
    const int reps = argc > 5 ? atoi(argv[5]) : 0;
    if (reps > 0) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        double sum = 0, best = 1e30;
        for (int i = 0; i <= reps; ++i) {
            cudaEventRecord(e0, 0);
            sf_spmv_amb(d_y, &amb, d_x, &plan);
            cudaEventRecord(e1, 0);
            cudaDeviceSynchronize();
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (i > 0) {
                sum += ms;
                if (ms < best) best = ms;
            }
        }
        printf("{\"impl\": \"cuda-c sf_csr2amb + sf_spmv_amb\", \"M\": %d, \"N\": %d, \"nnz\": %d, \"seg_size\": %d, "
               "\"block_size\": %d, \"reps\": %d, \"ms_mean\": %.6f, \"ms_min\": %.6f, \"gflops\": %.4f, \"value_bytes\": %d}\n",
               mat.M, mat.N, mat.nnz, (int)amb.seg_size, amb.block_size, reps, sum / reps, best,
               2.0 * mat.nnz / (sum / reps) / 1e6, (int)sizeof(real));
    }
    if (!strcmp(argv[4], "-")) return 0;
    FILE *f = fopen(argv[4], "wb");
    if (!f) {
        perror(argv[4]);
        return 2;
    }
    int hdr[10] = {mat.M, mat.N, mat.nnz, amb.pad_M, amb.c_size, amb.nnz, amb.block_size, (int)amb.seg_size,
                   (int)amb.seg_num, (int)sizeof(real)};
    put(f, hdr, sizeof(hdr));
    put(f, mat.rpt, sizeof(int) * (mat.M + 1));
    put(f, mat.col, sizeof(int) * mat.nnz);
    put(f, mat.val, sizeof(real) * mat.nnz);
    put_dev(f, amb.d_cs, amb.c_size);
    put_dev(f, amb.d_cl, amb.c_size);
    put_dev(f, amb.d_sellcs_col, amb.nnz / amb.block_size);
    put_dev(f, amb.d_sellcs_val, amb.nnz);
    put_dev(f, amb.d_s_write_permutation, (size_t)amb.c_size * amb.chunk);
    put_dev(f, amb.d_s_write_permutation_offset, amb.c_size);
    put_dev(f, amb.d_write_permutation, (size_t)amb.c_size * amb.chunk);
    put(f, x, sizeof(real) * mat.N);
    put_dev(f, d_y, mat.M);
    fclose(f);
    printf("dumped %s: M=%d N=%d nnz=%d c_size=%d nnz_amb=%d seg=%d bs=%d\n", argv[4], mat.M, mat.N, mat.nnz, amb.c_size,
           amb.nnz, (int)amb.seg_size, amb.block_size);
    return 0;
}
