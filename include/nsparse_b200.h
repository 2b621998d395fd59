/*
 * nsparse_b200.h -- plain C ABI of libnsparse_b200.so
 *
 * extern "C", raw pointers and sizes only; no torch / C++ types.  This is what a
 * binding (ctypes in nsparse_b200/_lib.py, cgo, JNI ...) loads.  Every entry point
 * names the reference interface it replaces (paths relative to the nsparse tree).
 *
 * Conventions
 *   - all functions return 0 on success, a negative nsp_status otherwise;
 *     nsp_last_error() gives the message of the last failure on that context.
 *   - "d_" arguments are DEVICE pointers on the context's GPU, "h_" are HOST pointers.
 *   - CSR is 0-based; rpt of the INPUTS is int32 (nnz < 2^31, like sfCSR); the row
 *     pointer of the PRODUCT is int64 because C = A*A of an R-MAT scale-20 graph has
 *     ~9e9 entries, which sfCSR's `int nnz` (nsparse.h:62-75) cannot hold.
 *     nsp_rpt64_to_rpt32() narrows it for callers that keep the sfCSR layout.
 *   - calls are asynchronous on the context's stream unless they return a host scalar
 *     (nnz, flop); nsp_sync() joins.
 */
#ifndef NSPARSE_B200_C_ABI_H
#define NSPARSE_B200_C_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nsp_context nsp_context;

typedef enum {
    NSP_OK = 0,
    NSP_ERR_CUDA = -1,        /* a CUDA runtime call failed                          */
    NSP_ERR_ARG = -2,         /* bad argument                                        */
    NSP_ERR_OVERFLOW = -3,    /* result does not fit the requested index width       */
    NSP_ERR_NOMEM = -4
} nsp_status;

/* ------------------------------------------------------------------------------------
 * context: device, stream, grow-only workspace arena.  Replaces the per-call
 * init_bin/release_bin (kernel_spgemm_hash_d.cu:33-68: 7 cudaStreamCreate + 5 cudaMalloc
 * per SpGEMM) with state that is created once.
 * ---------------------------------------------------------------------------------- */
int nsp_create(nsp_context **ctx, int device);
int nsp_destroy(nsp_context *ctx);
const char *nsp_last_error(nsp_context *ctx);
/* stream == NULL selects the legacy default stream (what the sample drivers time on). */
int nsp_set_stream(nsp_context *ctx, void *cuda_stream);
int nsp_sync(nsp_context *ctx);
/* tuning knobs, mostly for tests and measurements: "sym_bitmap_min" / "num_bitmap_min" (rows above this many
 * products / entries take the bitmap kernels), "sym_window_shift" / "num_window_shift" (log2 of the bitmap window),
 * "num_cap" (upper limit of the accumulator chunk), "sort" (0: the numeric
 * phase may leave the columns of a row unsorted -- SpGEMM_Hash_Numeric<sort = false> of cuda-cpp/inc/HashSpGEMM_volta.hpp:
 * 1018-1031; the hash classes then skip their per-row sort, the bitmap class is sorted by construction),
 * "hash_order" (how a hash-class row is ordered: 0 buckets, 1 bitonic sort always, 2 no shared-memory buckets),
 * "no_seg", "no_flat" (1: never the flat traversal of short B rows, -1: always), "no_ranges" (1: never the
 * column-range hash kernel for wide products with short B rows, -1: whenever it can run), "no_vec" (no 128-bit loads of
 * B.col), "no_fork" (long rows on the main stream), "gather_sm", "gather_tma", "push_sms", "dma_tile_log" (multi-GPU,
 * see nsp_spgemm_set_peers), "profile", "debug", "phase_timing".  NSP_OPTIONS="name=value,..." in the environment
 * sets them at nsp_create. */
int nsp_set_option(nsp_context *ctx, const char *name, long long value);
/* With option "profile" = 1 every row-class kernel launch is bracketed by CUDA events on the
 * context's stream.  nsp_profile_dump syncs, writes one line per launch
 * ("<kernel> <ms> <rows> <intermediate products> <A entries> <C entries>\n") into buf and clears
 * the log (<C entries> is 0 for the symbolic kernels). */
int nsp_profile_dump(nsp_context *ctx, char *buf, size_t buflen);
/* number of kernels this library launched on the context since creation (bench.py: gpu_launches) */
long long nsp_launch_count(nsp_context *ctx);

/* ------------------------------------------------------------------------------------
 * hash SpGEMM, C = A * B       A: M x K,  B: K x N,  C: M x N
 * ---------------------------------------------------------------------------------- */

/* get_spgemm_flop (kernel_spgemm_cu_csr.cu:35-57): flop = 2 * sum_i sum_{j in A_i} nnz(B_j). */
int nsp_spgemm_flop(nsp_context *ctx, int M,
                    const int *d_a_rpt, const int *d_a_col, const int *d_b_rpt,
                    long long *h_flop);

/* Symbolic phase = init_bin + set_max_bin + set_row_nnz (kernel_spgemm_hash_d.cu:33-198,
 * 1077-1185).  Writes the exclusive row pointer of C (int64, M+1 entries) and returns
 * nnz(C) and the number of intermediate products.  Synchronises (nnz is needed to
 * allocate C, as in the reference :1184). */
int nsp_spgemm_symbolic(nsp_context *ctx, int M, int K, int N,
                        const int *d_a_rpt, const int *d_a_col,
                        const int *d_b_rpt, const int *d_b_col,
                        long long *d_c_rpt64,
                        long long *h_nnz_c, long long *h_intprod);

/* Numeric phase = set_min_bin + calculate_value_col_bin (kernel_spgemm_hash_d.cu:201-246,
 * 1187-1288).  d_c_col / d_c_val must hold nnz(C) entries.  Rows of C come out with
 * ascending column indices; numerical zeros are kept (structural product).
 * Must follow nsp_spgemm_symbolic on the same context with the same A, B. */
int nsp_spgemm_numeric_s(nsp_context *ctx, int M, int K, int N,
                         const int *d_a_rpt, const int *d_a_col, const float *d_a_val,
                         const int *d_b_rpt, const int *d_b_col, const float *d_b_val,
                         const long long *d_c_rpt64, int *d_c_col, float *d_c_val);
int nsp_spgemm_numeric_d(nsp_context *ctx, int M, int K, int N,
                         const int *d_a_rpt, const int *d_a_col, const double *d_a_val,
                         const int *d_b_rpt, const int *d_b_col, const double *d_b_val,
                         const long long *d_c_rpt64, int *d_c_col, double *d_c_val);

/* The numeric phase for the rows [row0, row0 + nrows) only; every pointer is that of the FULL array.  The
 * multi-GPU pipeline calls it piece by piece so that finished pieces of C travel to the other GPUs (copy
 * engines over NVLink) while the next piece is computed. */
int nsp_spgemm_numeric_rows_s(nsp_context *ctx, int M, int K, int N, int row0, int nrows,
                              const int *d_a_rpt, const int *d_a_col, const float *d_a_val,
                              const int *d_b_rpt, const int *d_b_col, const float *d_b_val,
                              const long long *d_c_rpt64, int *d_c_col, float *d_c_val);
int nsp_spgemm_numeric_rows_d(nsp_context *ctx, int M, int K, int N, int row0, int nrows,
                              const int *d_a_rpt, const int *d_a_col, const double *d_a_val,
                              const int *d_b_rpt, const int *d_b_col, const double *d_b_val,
                              const long long *d_c_rpt64, int *d_c_col, double *d_c_val);

/* Narrow an int64 row pointer to the int32 one sfCSR carries; NSP_ERR_OVERFLOW if
 * nnz > INT_MAX (the reference silently wraps, kernel_spgemm_hash_d.cu:1183). */
int nsp_rpt64_to_rpt32(nsp_context *ctx, int M, const long long *d_rpt64, long long nnz, int *d_rpt32);

/* Fold of a device CSR with an int64 row pointer, for parity checks at sizes nobody wants on the host (C = A^2
 * of R-MAT scale 20 is 78 GB): h_hash2 = {position-dependent 64-bit hash of rpt[0..M], of col[0..nnz)} (exact,
 * independent of the order the kernel visits the entries in), h_sum2 = {sum of val, sum of val * (col % 1021 + 1)}
 * accumulated in double (floating-point atomics: equal up to rounding between runs).  Synchronises. */
int nsp_csr_fold_s(nsp_context *ctx, int M, long long nnz, const long long *d_rpt64, const int *d_col,
                   const float *d_val, unsigned long long *h_hash2, double *h_sum2);
int nsp_csr_fold_d(nsp_context *ctx, int M, long long nnz, const long long *d_rpt64, const int *d_col,
                   const double *d_val, unsigned long long *h_hash2, double *h_sum2);

/* spgemm_kernel_hash (kernel_spgemm_hash_d.cu:1035-1075) with HOST buffers: copies A and
 * B to the device, runs both phases and leaves C on the device inside the context.
 * nsp_spgemm_host_fetch_* then copies C into caller-provided host arrays (any of the
 * three may be NULL to skip it) and nsp_spgemm_host_release frees the device copy. */
int nsp_spgemm_host_s(nsp_context *ctx, int M, int K, int N,
                      const int *h_a_rpt, const int *h_a_col, const float *h_a_val,
                      const int *h_b_rpt, const int *h_b_col, const float *h_b_val,
                      long long *h_nnz_c);
int nsp_spgemm_host_d(nsp_context *ctx, int M, int K, int N,
                      const int *h_a_rpt, const int *h_a_col, const double *h_a_val,
                      const int *h_b_rpt, const int *h_b_col, const double *h_b_val,
                      long long *h_nnz_c);
int nsp_spgemm_host_fetch_s(nsp_context *ctx, long long *h_c_rpt64, int *h_c_col, float *h_c_val);
int nsp_spgemm_host_fetch_d(nsp_context *ctx, long long *h_c_rpt64, int *h_c_col, double *h_c_val);
/* Streams C through a caller-provided (pinned) staging buffer of `stage_bytes` and folds it
 * into a 64-bit checksum on the host; for results that do not fit host memory. */
int nsp_spgemm_host_drain(nsp_context *ctx, void *h_stage, size_t stage_bytes,
                          unsigned long long *h_checksum, long long *h_bytes);
/* The whole call end to end with overlap: A, B in (pinned) host memory, C streamed to the host through the
 * pinned staging buffer WHILE it is computed -- the numeric phase runs in `pieces` row ranges of ~equal output
 * and a second host thread drains every finished range on its own stream (order of the stream: row pointer,
 * then col and val of each range).  Returns nnz(C), the fold of the drained chunks and the bytes moved. */
int nsp_spgemm_host_stream_s(nsp_context *ctx, int M, int K, int N,
                             const int *h_a_rpt, const int *h_a_col, const float *h_a_val,
                             const int *h_b_rpt, const int *h_b_col, const float *h_b_val,
                             void *h_stage, size_t stage_bytes, int pieces, long long *h_nnz_c,
                             unsigned long long *h_checksum, long long *h_bytes);
int nsp_spgemm_host_stream_d(nsp_context *ctx, int M, int K, int N,
                             const int *h_a_rpt, const int *h_a_col, const double *h_a_val,
                             const int *h_b_rpt, const int *h_b_col, const double *h_b_val,
                             void *h_stage, size_t stage_bytes, int pieces, long long *h_nnz_c,
                             unsigned long long *h_checksum, long long *h_bytes);
int nsp_spgemm_host_release(nsp_context *ctx);

/* ------------------------------------------------------------------------------------
 * multi-GPU: allgatherv of C's row blocks over NVLink peer memory (new; the reference is single
 * GPU).  Every rank owns one full-size buffer per array of C; the buffers of the other ranks are
 * mapped into this process (CUDA IPC) and a rank PUSHES its block to all of them with plain
 * stores, at the displacement the block has in the full matrix -- one kernel per array, no
 * staging, no per-peer launch.  d_peer_bases: HOST array of `npeers` device pointers (the base of
 * the same array on every destination, any of which may be local); the block is
 * [byte_offset, byte_offset + nbytes) of every destination and starts at d_src; byte_offset, nbytes
 * and the pointers must be multiples of 4 bytes.
 * ---------------------------------------------------------------------------------- */
/* cudaMalloc + cudaIpcGetMemHandle: a buffer other processes of this node can map.  handle: 64 bytes. */
int nsp_peer_alloc(nsp_context *ctx, size_t bytes, void **d_ptr, unsigned char *handle64);
/* cudaIpcOpenMemHandle on the context's device (peer access to the exporting GPU is enabled lazily) */
int nsp_peer_open(nsp_context *ctx, const unsigned char *handle64, void **d_ptr);
int nsp_peer_close(nsp_context *ctx, void *d_ptr);
int nsp_peer_free(nsp_context *ctx, void *d_ptr);
/* cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault) on `cuda_stream` (NULL: the context's stream): with
 * a destination opened by nsp_peer_open this is a copy-engine transfer over NVLink that needs no SM. */
int nsp_copy_async(nsp_context *ctx, void *d_dst, const void *d_src, size_t bytes, void *cuda_stream);

/* Numeric phase overlapped with the allgatherv of C: after this call nsp_spgemm_numeric_* also sends every
 * entry of C it produces to d_peer_col[p] / d_peer_val[p] + elem_offset + (the same index) for p < npeers (at
 * most 7): the bases of the FULL C.col / C.val arrays of the other GPUs and the displacement of this rank's
 * row block in them (d_c_col / d_c_val passed to the numeric call must then be this rank's own full arrays
 * + elem_offset).  The block is cut into tiles of 2^20 .. 2^26 entries; the numeric kernels count finished
 * entries per tile and raise a flag in host-mapped memory when a tile is complete; the calling host thread
 * polls the flags while the kernels run and hands finished tiles to the copy engines (cudaMemcpyAsync, one
 * stream per peer, a short queue); what is finished but not queued when the kernels end is stored to all
 * peers by an SM copy kernel on the context's stream (option "gather_sm": 1 default, 0 copy engines only,
 * 2 SM stores only).  Option "gather_tma" = 1 selects the persistent TMA pusher kernel instead (tiles of 4096
 * entries, option "push_sms"; measured slower, kept as an option).  The transfer is complete when the
 * context's stream is.  npeers == 0 switches it off.
 * nsp_spgemm_peers_status synchronises and reports whether the pusher kernel ever gave up waiting (it never
 * should); nsp_spgemm_peers_stats returns how many tiles of the last product left through the copy engines
 * and through SM stores, and the time from the entry of the numeric call to the end of its kernels (what the
 * rank needed for its own block, the input of the feedback row partition). */
int nsp_spgemm_set_peers(nsp_context *ctx, int npeers, void *const *d_peer_col, void *const *d_peer_val,
                         long long elem_offset);
int nsp_spgemm_peers_status(nsp_context *ctx, int *h_error);
int nsp_spgemm_peers_stats(nsp_context *ctx, long long *h_copy_engine_tiles, long long *h_sm_tiles, double *h_kernel_ms);
int nsp_push_to_peers(nsp_context *ctx, int npeers, void *const *d_peer_bases, size_t byte_offset,
                      const void *d_src, size_t nbytes);

/* ------------------------------------------------------------------------------------
 * multi-GPU SpGEMM from ONE process (new; the reference is single GPU, its driver flow is
 * cuda-c/src/sample/spgemm/spgemm_hash.cu:14-94).  A and B are HOST CSR.  A is cut into ngpu contiguous row
 * blocks of ~equal intermediate products, B is replicated, one host thread per GPU runs the single-GPU
 * pipeline on its block and every GPU receives the FULL C (allgatherv over NVLink peer memory, overlapped with
 * the numeric phase as described at nsp_spgemm_set_peers).  Two phases like the single-GPU entry points, so
 * that the caller owns C:
 *   nsp_mgpu_spgemm_symbolic_*   uploads the blocks, symbolic phase on every GPU; returns nnz(C) and the products
 *   (caller allocates, ON GPU g, d_c_rpt64[g] int64[M+1], d_c_col[g] int32[nnz], d_c_val[g] real[nnz])
 *   nsp_mgpu_spgemm_numeric_*    numeric phase + gather; returns with every GPU idle and every copy complete
 * nsp_mgpu_block reports GPU g's rows [row0, row1), its displacement / entry count in C and the wall-clock
 * milliseconds its host thread spent in the two phases of the last product.  devices == NULL: GPUs 0..ngpu-1;
 * at most 8 GPUs; all pairs need peer access (NVLink).
 * ---------------------------------------------------------------------------------- */
typedef struct nsp_mgpu nsp_mgpu;
int nsp_mgpu_create(nsp_mgpu **mg, int ngpu, const int *devices);
int nsp_mgpu_destroy(nsp_mgpu *mg);
const char *nsp_mgpu_last_error(nsp_mgpu *mg);
int nsp_mgpu_ngpu(nsp_mgpu *mg);
nsp_context *nsp_mgpu_context(nsp_mgpu *mg, int g);      /* the per-GPU context (options, nsp_rpt64_to_rpt32, ...) */
int nsp_mgpu_block(nsp_mgpu *mg, int g, int *device, int *row0, int *row1, long long *elem0, long long *nnz,
                   double *ms_symbolic, double *ms_numeric);
int nsp_mgpu_spgemm_symbolic_s(nsp_mgpu *mg, int M, int K, int N,
                               const int *h_a_rpt, const int *h_a_col, const float *h_a_val,
                               const int *h_b_rpt, const int *h_b_col, const float *h_b_val,
                               long long *h_nnz_c, long long *h_intprod);
int nsp_mgpu_spgemm_symbolic_d(nsp_mgpu *mg, int M, int K, int N,
                               const int *h_a_rpt, const int *h_a_col, const double *h_a_val,
                               const int *h_b_rpt, const int *h_b_col, const double *h_b_val,
                               long long *h_nnz_c, long long *h_intprod);
int nsp_mgpu_spgemm_numeric_s(nsp_mgpu *mg, long long *const *d_c_rpt64, int *const *d_c_col, float *const *d_c_val);
int nsp_mgpu_spgemm_numeric_d(nsp_mgpu *mg, long long *const *d_c_rpt64, int *const *d_c_col, double *const *d_c_val);

/* nsp_push_to_peers through an NVSwitch multicast address (NVLS): d_multicast_base is the multicast
 * mapping of the same array on all GPUs; one multimem.st per 16 bytes reaches every copy, the local one
 * included.  Same alignment rules. */
int nsp_push_multicast(nsp_context *ctx, void *d_multicast_base, size_t byte_offset, const void *d_src,
                       size_t nbytes);

/* ------------------------------------------------------------------------------------
 * AMB SpMV, y = A * x
 * ---------------------------------------------------------------------------------- */

/* Device part of sfAMB (nsparse.h:78-107); arrays are cudaMalloc'ed by nsp_csr2amb_*. */
typedef struct {
    int *d_cs;
    unsigned int *d_cl;
    unsigned short *d_sellcs_col;
    void *d_sellcs_val;                 /* float* or double*                              */
    unsigned short *d_s_write_permutation;
    unsigned short *d_s_write_permutation_offset;
    int *d_write_permutation;
    int block_size;
    int nnz;                            /* stored values incl. padding                    */
    int M, N, pad_M;
    int chunk;                          /* 32                                             */
    int SIGMA;                          /* 32768                                          */
    int c_size;                         /* non-empty chunks                               */
    long long seg_size, seg_num;
    long long thread_grid, thread_block;/* launch shape picked by the planner             */
} nsp_amb;

/* sf_csr2amb (convert_amb.cu:835-929).  seg_size == 0 / block_size == 0 ask the planner
 * (footprint model of convert_amb.cu:785-797, optionally refined by timing when
 * autotune != 0, which is what the reference's `AT` build does).  d_x is only read when
 * autotune != 0. */
int nsp_csr2amb_s(nsp_context *ctx, int M, int N, int nnz,
                  const int *d_rpt, const int *d_col, const float *d_val,
                  long long seg_size, int block_size, int autotune, const float *d_x,
                  nsp_amb *out);
int nsp_csr2amb_d(nsp_context *ctx, int M, int N, int nnz,
                  const int *d_rpt, const int *d_col, const double *d_val,
                  long long seg_size, int block_size, int autotune, const double *d_x,
                  nsp_amb *out);
int nsp_amb_free(nsp_context *ctx, nsp_amb *mat);

/* sf_spmv_amb (kernel_spmv_amb.cu:98-104).  y[0..M) is overwritten.  x needs N entries
 * (no padding is read), y needs M entries. */
int nsp_spmv_amb_s(nsp_context *ctx, const nsp_amb *mat, const float *d_x, float *d_y);
int nsp_spmv_amb_d(nsp_context *ctx, const nsp_amb *mat, const double *d_x, double *d_y);

/* The same with HOST x / y (copies inside). */
int nsp_spmv_amb_host_s(nsp_context *ctx, const nsp_amb *mat, const float *h_x, float *h_y);
int nsp_spmv_amb_host_d(nsp_context *ctx, const nsp_amb *mat, const double *h_x, double *h_y);

/* plain device -> host copy on the context's stream (synchronous); lets a binding read the
 * cudaMalloc'ed arrays of nsp_amb back without a CUDA runtime binding of its own */
int nsp_memcpy_d2h(nsp_context *ctx, void *h_dst, const void *d_src, size_t bytes);

/* ------------------------------------------------------------------------------------
 * MatrixMarket coordinate file -> host CSR, parsed in parallel (OpenMP).  flags == 0 reproduces the
 * reference reader convert_file_csr (nsparse.cu:14-136) entry for entry: "general" in the banner or
 * off-diagonals mirrored right after the original, file order inside every row, nothing sorted or merged,
 * missing value = 1, at most `nz` entries.  NSP_MTX_SORT_MERGE orders every row by column and sums
 * duplicates.  The arrays are malloc'ed (free() or nsp_free_host); val is float* or double*.
 * ---------------------------------------------------------------------------------- */
#define NSP_MTX_SORT_MERGE 1
int nsp_read_mtx(const char *path, int is_double, int flags, int *M, int *N, long long *nnz, int *nnz_max,
                 int **h_rpt, int **h_col, void **h_val);
void nsp_free_host(void *p);

/* ------------------------------------------------------------------------------------
 * synthetic inputs (host side, OpenMP; counter-based so every rank and the numpy mirror
 * in nsparse_b200/gen.py produce identical edges)
 * ---------------------------------------------------------------------------------- */
/* Graph500 Kronecker edges, (a,b,c,d) = (0.57,0.19,0.19,0.05), no permutation, no noise. */
int nsp_gen_rmat_edges(int scale, long long n_edges, unsigned long long seed,
                       long long *h_src, long long *h_dst);

#ifdef __cplusplus
}
#endif
#endif /* NSPARSE_B200_C_ABI_H */
