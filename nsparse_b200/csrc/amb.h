// nsparse-b200: AMB (adaptive multi-level blocking) SpMV -- shared declarations.
#pragma once

#include "../../include/nsparse_b200.h"

struct nsp_context;

namespace nsp {

constexpr int kAmbMaxBlock = 20;      // MAX_BLOCK_SIZE (cuda-c/inc/nsparse.h:38)
constexpr int kAmbSigma = 32768;      // SHORT_MAX: rows per sort window (nsparse.h:31, convert_amb.cu:863)

// sf_csr2amb (convert_amb.cu:835-929).  seg_size == 0 / block_size == 0: planned with the
// reference's footprint model.  The seven device arrays of *out are cudaMalloc'ed.
template <typename real>
int amb_convert(nsp_context *ctx, int M, int N, int nnz, const int *rpt, const int *col, const real *val,
                long long seg_size, int block_size, nsp_amb *out);

// sf_spmv_amb (kernel_spmv_amb.cu:98-104): y[0..M) = A x
template <typename real>
int amb_spmv(nsp_context *ctx, const nsp_amb *mat, const real *x, real *y);

}  // namespace nsp
