/*
 * nsparse.h -- the drop-in contract of nsparse-b200.
 *
 * This header is what the UNCHANGED cuda-c sample drivers of EBD-CREST/nsparse
 * (cuda-c/src/sample/spgemm/spgemm_hash.cu, cuda-c/src/sample/spmv/spmv_amb.cu)
 * include.  It is written from scratch, but the four structs keep the field
 * names, order and types of the reference contract (cuda-c/inc/nsparse.h:50-121)
 * because the drivers touch fields directly and pass the structs by value:
 *
 *     sizeof(sfPlan) == 48, sizeof(sfCSR) == 72, sizeof(sfAMB) == 176   (-DDOUBLE, x86-64)
 *
 * (checked by static_assert in nsparse_b200/csrc/compat_api.cu).  Like the
 * reference header there is NO extern "C": the drivers are compiled by nvcc as
 * C++ and import the mangled names, so libnsparse_s.a / libnsparse_d.a are two
 * builds of the same sources with -DFLOAT / -DDOUBLE (reference: cuda-c/Makefile:99-113).
 *
 * The plain-C ABI used by everything that is not one of those drivers (ctypes,
 * cgo-style bindings, bench.py) is in nsparse_b200.h.
 */
#ifndef NSPARSE_B200_CONTRACT_H
#define NSPARSE_B200_CONTRACT_H

#include <stddef.h>
#include <cuda_runtime.h>

/* ---- precision switch (reference: nsparse.h:3-11) --------------------------------- */
#if defined(FLOAT)
typedef float real;
#else /* DOUBLE or nothing */
typedef double real;
#endif

/* ---- constants the drivers use ---------------------------------------------------- */
#define div_round_up(a, b) (((a) % (b) == 0) ? (a) / (b) : (a) / (b) + 1)

#define WARP_BIT 5
#define WARP 32                      /* d_y is padded by WARP in spmv_amb.cu:33           */
#define MAX_LOCAL_THREAD_NUM 1024
#define MAX_THREAD_BLOCK (MAX_LOCAL_THREAD_NUM / WARP)

#define TRI_NUM 101                  /* SpMV repetitions in the driver (first discarded)  */
#define TEST_NUM 2
#define SPGEMM_TRI_NUM 11            /* SpGEMM repetitions in the driver                   */

#define sfFLT_MAX 1000000000
#define SHORT_MAX 32768              /* sigma window of the AMB row sort                   */
#define SHORT_MAX_BIT 15
#define USHORT_MAX 65536             /* largest column segment (16-bit compressed column) */
#define USHORT_MAX_BIT 16
#define SCL_BORDER 16                /* d_cl = (blocks-1) | segment << SCL_BORDER          */
#define SCL_BIT ((1 << SCL_BORDER) - 1)
#define MAX_BLOCK_SIZE 20            /* d_x is padded by this in spmv_amb.cu:32            */

#define sfDEBUG                      /* drivers self-check when this is defined            */

typedef enum { FALSE, TRUE } BOOL;

/* ---- AMB plan: 48 bytes ----------------------------------------------------------- */
typedef struct {
    size_t thread_grid;              /* launch shape chosen by the planner                 */
    size_t thread_block;
    BOOL isPlan;                     /* TRUE: seg_size/block_size are given by the caller  */
    int SIGMA;
    size_t seg_size;                 /* columns per segment, <= 65536                      */
    size_t seg_num;
    int block_size;                  /* 1..MAX_BLOCK_SIZE consecutive x entries per block  */
} sfPlan;

/* ---- CSR container: 72 bytes ------------------------------------------------------ */
typedef struct {
    int *rpt;                        /* host mirror                                        */
    int *col;
    real *val;
    int *d_rpt;                      /* device arrays (cudaMalloc; freed by release_csr)   */
    int *d_col;
    real *d_val;
    int M;
    int N;
    int nnz;
    int nnz_max;                     /* longest row                                        */
    char *matrix_name;
} sfCSR;

/* ---- AMB container: 176 bytes ----------------------------------------------------- */
typedef struct {
    int *cs;                         /* host mirrors (unused by the drivers)               */
    unsigned int *cl;
    unsigned short *sellcs_col;
    real *sellcs_val;
    unsigned short *s_write_permutation;
    unsigned short *s_write_permutation_offset;
    int *write_permutation;
    int *d_cs;                       /* [c_size] value offset of each 32-row chunk         */
    unsigned int *d_cl;              /* [c_size] (blocks-1) | segment<<16                  */
    unsigned short *d_sellcs_col;    /* [nnz/block_size] block start column mod seg_size   */
    real *d_sellcs_val;              /* [nnz] chunk-column-major, block_size vals/block    */
    unsigned short *d_s_write_permutation;        /* [c_size*32] output row & 0xFFFF       */
    unsigned short *d_s_write_permutation_offset; /* [c_size]    output row >> 16          */
    int *d_write_permutation;        /* [c_size*32] full output row                        */
    int block_size;
    int nnz;                         /* stored values including padding                    */
    int M;
    int N;
    int pad_M;
    int chunk;
    int SIGMA;
    int group_num_col;
    int nnz_max;
    int c_size;
    size_t seg_size;
    size_t seg_num;
    char *matrix_name;
} sfAMB;

/* ---- SpGEMM binning state.  Kept for source compatibility only: nsparse-b200 plans on
 *      the device inside a cached context and never hands this to callers. ------------ */
typedef struct {
    cudaStream_t *stream;
    int *bin_size;
    int *bin_offset;
    int *d_bin_size;
    int *d_bin_offset;
    int *d_row_nz;
    int *d_row_perm;
    int max_intprod;
    int max_nz;
    int *d_max;
} sfBIN;

/* ---- host helpers (reference: cuda-c/src/nsparse.cu) ------------------------------ */
void init_csr_matrix_from_file(sfCSR *mat, char *file_name);   /* nsparse.cu:138 */
void csr_memcpy(sfCSR *mat);                                   /* nsparse.cu:146 */
void csr_memcpyDtH(sfCSR *mat);                                /* nsparse.cu:158 */
void init_vector(real *x, int row);                            /* nsparse.cu:190 */

void release_cpu_csr(sfCSR mat);
void release_cpu_amb(sfAMB mat);
void release_csr(sfCSR mat);
void release_amb(sfAMB mat);

/* ---- AMB SpMV path ---------------------------------------------------------------- */
void init_plan(sfPlan *plan);                                  /* nsparse.cu:171 */
void set_plan(sfPlan *plan, size_t seg_size, int block_size);  /* nsparse.cu:176 */
void sf_csr2amb(sfAMB *mat, sfCSR *csr_mat, real *d_x, sfPlan *plan);   /* convert_amb.cu:835 */
void sf_spmv_amb(real *d_y, sfAMB *mat, real *d_x, sfPlan *plan);       /* kernel_spmv_amb.cu:98 */
void csr_kernel(real *csr_ans, sfCSR *cpu_mat, real *rhs_vec); /* CPU SpMV, nsparse.cu:240 */
void ans_check(real *csr_ans, real *ans_vec, int N);           /* nsparse.cu:261 */

/* ---- hash SpGEMM path ------------------------------------------------------------- */
void get_spgemm_flop(sfCSR *a, sfCSR *b, int M, long long int *flop);   /* kernel_spgemm_cu_csr.cu:35 */
void spgemm_kernel_hash(sfCSR *a, sfCSR *b, sfCSR *c);         /* kernel_spgemm_hash_*.cu:1035 */
void spgemm_cu_csr(sfCSR *a, sfCSR *b, sfCSR *c);              /* comparison answer for the driver's self-check */
#ifdef CUSPARSE_H_   /* the cuSPARSE comparison entry points keep the reference's signatures (nsparse.h:160-165, kernel_spmv_cu_csr.cu:9) */
void spgemm_kernel_cu_csr(sfCSR *a, sfCSR *b, sfCSR *c, cusparseHandle_t *cusparseHandle, cusparseOperation_t *trans_a,
                          cusparseOperation_t *trans_b, cusparseMatDescr_t *descr_a, cusparseMatDescr_t *descr_b);
void sf_spmv_cu_csr(real *d_y, sfCSR *mat, real *d_x, cusparseHandle_t *cusparseHandle, cusparseMatDescr_t *descr);
#endif
void check_spgemm_answer(sfCSR c, sfCSR ans);                  /* nsparse.cu:300 */

/* ---- multi-GPU extension (not in the reference; does not disturb the symbols above) ------
 * a, b: HOST CSR; c: array of ngpu sfCSR, c[g].d_* = the full product on GPU g.  Driver:
 * nsparse_b200/csrc/sample/spgemm_hash_mgpu.cu (bin/spgemm_hash_mgpu_{s,d}). */
void spgemm_kernel_hash_mgpu(sfCSR *a, sfCSR *b, sfCSR *c, int ngpu);
void release_csr_mgpu(sfCSR *c, int ngpu);

#endif /* NSPARSE_B200_CONTRACT_H */
