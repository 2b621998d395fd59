// nsparse-b200: layout of the device-side SpGEMM plan and the kernel-class ladder.
#pragma once

#include "common.cuh"

struct nsp_context;

namespace nsp {

// d_bins layout (ints)
constexpr int kBinHist = 0;      // [kNumBins] rows per bin
constexpr int kBinStart = 32;    // [kNumBins] first slot of bin b in row_perm (heaviest bin first)
constexpr int kBinCursor = 64;   // [kNumBins] scatter cursors
constexpr int kBinQueue = 96;    // [kNumQueues] dynamic row-queue heads, one per kernel class
constexpr int kNumQueues = 16;
constexpr int kBinInts = 128;

// d_binsum layout (unsigned long long): per-bin sums used to pick lanes-per-B-row
constexpr int kSumIp = 0;        // [kNumBins] intermediate products of the rows in the bin
constexpr int kSumLen = 32;      // [kNumBins] A entries of the rows in the bin
constexpr int kSumCnt = 64;      // [kNumBins] nnz(C_i) of the rows in the bin (numeric plan only)
constexpr int kSumInts = 96;

// d_scalars layout (long long)
constexpr int kScalarIp = 0;     // total intermediate products (uncapped)
constexpr int kScalarNnz = 1;    // nnz(C)
constexpr int kScalarNnzB = 3;     // B.rpt[K]
constexpr int kScalarMaxLen = 4;   // longest row of A
constexpr int kScalarUnsorted = 2;   // entries of B below their predecessor in the same row

// Bin shifts: symbolic bins rows by min(intermediate products, N) with bin 0 = "<= 32"
// (IMB_PWMIN of the reference), numeric bins by nnz(C_i) with bin 0 = "<= 16" (B_PWMIN).
constexpr int kSymShift = 5;
constexpr int kNumShift = 4;

// A kernel class owns the bins [bin_lo, bin_hi]; its rows are row_perm[start[bin_hi] ..
// start[bin_lo] + hist[bin_lo]).
__device__ __forceinline__ void class_range(const int *bins, int bin_lo, int bin_hi, int &lo, int &hi)
{
    lo = bins[kBinStart + bin_hi];
    hi = bins[kBinStart + bin_lo] + bins[kBinHist + bin_lo];
}

int plan_reserve(nsp_context *ctx, int M);
int plan_by_intprod(nsp_context *ctx, int M, int K, int cap, const int *a_rpt, const int *a_col,
                    const int *b_rpt, const int *b_col);
int plan_by_count(nsp_context *ctx, int M, int shift, const int *a_rpt, int row0 = 0);
int scan_row_counts(nsp_context *ctx, int M, long long *rpt64);
int reserve_entry_segments(nsp_context *ctx, long long count, int nwin);
int build_entry_segments(nsp_context *ctx, const int *a_col, long long count, const int *b_rpt, const int *b_col,
                         int nwin, int wshift);

}  // namespace nsp
