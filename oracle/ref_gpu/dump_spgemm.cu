// Test infrastructure: runs the REFERENCE's own hash SpGEMM on a GPU, times it with the reference's
// protocol and dumps C, so that oracle/oracle.c and the CUDA product can be pinned to what the reference
// really computes, and so that bench.py has a same-box reference GPU number.  Never part of the product.
//
// Two builds of this file (oracle/Makefile target ref_spgemm):
//   default       the cuda-c tree: spgemm_kernel_hash (kernel_spgemm_hash_{s,d}.cu:1035-1075) compiled from
//                 where it lies with `__shfl_xor(` / `__shfl(` spelled `_sync(0xffffffffu, ` on a temporary copy
//                 (the originals do not compile for sm_70+), + nsparse.cu unmodified;
//   -DREF_CPP     the cuda-cpp tree: SpGEMM_Hash of HashSpGEMM_volta.hpp (:974-1010), header unmodified.
//
//   dump_spgemm <A> <B|-> <out.bin|-> [reps]
//
// <A>, <B>: *.mtx (read by the reference's own reader) or a raw CSR file
//           int32 {M, N, nnz, sizeof(real)}, rpt int32[M+1], col int32[nnz], val real[nnz];  "-" = B is A.
// out.bin:  int32 {M, N, nnz, sizeof(real)}, rpt int32[M+1], col int32[nnz], val real[nnz] of C.
// stdout:   one JSON line {"impl", "M", "N", "nnz_a", "ip", "nnz_c", "reps", "ms_mean", "ms_min", "gflops"}:
//           mean over `reps` calls after one warm-up, C released and re-allocated by every call exactly as
//           spgemm_hash.cu:35-52 does (the cudaMalloc of C is inside the timed region).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime.h>

#ifdef REF_CPP
#include <thrust/sort.h>
#include <thrust/device_vector.h>
#include <thrust/scan.h>
#include <helper_cuda.h>
#ifdef FLOAT
typedef float real;
#else
typedef double real;
#endif
#include <HashSpGEMM_volta.hpp>
typedef CSR<int, real> Mat;
#define M_ROWS(m) (m).nrow
#define M_COLS(m) (m).ncolumn
#define M_COL(m) (m).colids
#define M_VAL(m) (m).values
static const char *kImpl = "cuda-cpp HashSpGEMM_volta.hpp SpGEMM_Hash";
#else
#include <nsparse.h>
typedef sfCSR Mat;
#define M_ROWS(m) (m).M
#define M_COLS(m) (m).N
#define M_COL(m) (m).col
#define M_VAL(m) (m).val
#ifdef NSP_OURS   // the same driver linked against nsparse-b200's libnsparse_{s,d}.a (top-level Makefile: drivers)
static const char *kImpl = "nsparse-b200 spgemm_kernel_hash";
#else
static const char *kImpl = "cuda-c spgemm_kernel_hash (_sync spelling)";
#endif
#endif

static bool ends_with(const char *s, const char *suf)
{
    const size_t a = strlen(s), b = strlen(suf);
    return a >= b && !strcmp(s + a - b, suf);
}

static void load(Mat &m, char *path)
{
    if (ends_with(path, ".mtx")) {
#ifdef REF_CPP
        m.init_data_from_mtx(path);
#else
        init_csr_matrix_from_file(&m, path);
#endif
        return;
    }
    FILE *f = fopen(path, "rb");
    if (!f) {
        perror(path);
        exit(2);
    }
    int hdr[4];
    if (fread(hdr, sizeof(int), 4, f) != 4 || hdr[3] != (int)sizeof(real)) {
        fprintf(stderr, "%s: bad header (value size %d, built for %d)\n", path, hdr[3], (int)sizeof(real));
        exit(2);
    }
    M_ROWS(m) = hdr[0];
    M_COLS(m) = hdr[1];
    m.nnz = hdr[2];
    m.rpt = (int *)malloc(sizeof(int) * (hdr[0] + 1));
    M_COL(m) = (int *)malloc(sizeof(int) * (hdr[2] ? hdr[2] : 1));
    M_VAL(m) = (real *)malloc(sizeof(real) * (hdr[2] ? hdr[2] : 1));
    if (fread(m.rpt, sizeof(int), hdr[0] + 1, f) != (size_t)hdr[0] + 1 || fread(M_COL(m), sizeof(int), hdr[2], f) != (size_t)hdr[2] ||
        fread(M_VAL(m), sizeof(real), hdr[2], f) != (size_t)hdr[2]) {
        fprintf(stderr, "%s: short file\n", path);
        exit(2);
    }
    fclose(f);
#ifdef REF_CPP
    m.device_malloc = false;
#else
    m.matrix_name = path;
    m.nnz_max = 0;
#endif
}

int main(int argc, char **argv)
{
    if (argc < 4) {
        fprintf(stderr, "usage: %s A B|- out.bin|- [reps]\n", argv[0]);
        return 1;
    }
    const int reps = argc > 4 ? atoi(argv[4]) : 10;
    const bool same = !strcmp(argv[2], "-");
    Mat a, b, c;
    load(a, argv[1]);
    load(b, same ? argv[1] : argv[2]);
    long long ip = 0;
    for (int i = 0; i < a.nnz; ++i) ip += b.rpt[M_COL(a)[i] + 1] - b.rpt[M_COL(a)[i]];
#ifdef REF_CPP
    a.memcpyHtD();
    b.memcpyHtD();
#else
    csr_memcpy(&a);
    csr_memcpy(&b);
#endif
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double sum = 0, best = 1e30;
    for (int i = 0; i <= reps; ++i) {
        if (i > 0) {
#ifdef REF_CPP
            c.release_csr();
#else
            release_csr(c);
#endif
        }
        cudaEventRecord(e0, 0);
#ifdef REF_CPP
        SpGEMM_Hash(a, b, c);
#else
        spgemm_kernel_hash(&a, &b, &c);
#endif
        cudaEventRecord(e1, 0);
        cudaDeviceSynchronize();
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (i > 0 || reps == 0) {
            sum += ms;
            if (ms < best) best = ms;
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e));
        return 3;
    }
    const double mean = sum / (reps > 0 ? reps : 1);
    if (strcmp(argv[3], "-")) {
#ifdef REF_CPP
        c.memcpyDtH();
#else
        csr_memcpyDtH(&c);
#endif
        FILE *f = fopen(argv[3], "wb");
        if (!f) {
            perror(argv[3]);
            return 2;
        }
        int hdr[4] = {M_ROWS(c), M_COLS(c), c.nnz, (int)sizeof(real)};
        fwrite(hdr, sizeof(int), 4, f);
        fwrite(c.rpt, sizeof(int), (size_t)M_ROWS(c) + 1, f);
        fwrite(M_COL(c), sizeof(int), (size_t)c.nnz, f);
        fwrite(M_VAL(c), sizeof(real), (size_t)c.nnz, f);
        fclose(f);
    }
    printf("{\"impl\": \"%s\", \"M\": %d, \"N\": %d, \"nnz_a\": %d, \"ip\": %lld, \"nnz_c\": %d, \"reps\": %d, "
           "\"ms_mean\": %.6f, \"ms_min\": %.6f, \"gflops\": %.4f, \"value_bytes\": %d}\n",
           kImpl, M_ROWS(a), M_COLS(b), a.nnz, ip, c.nnz, reps, mean, best, 2.0 * (double)ip / mean / 1e6, (int)sizeof(real));
    return 0;
}
