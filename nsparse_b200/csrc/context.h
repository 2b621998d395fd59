// nsparse-b200: the per-GPU context behind the C ABI (include/nsparse_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <chrono>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"

// Everything a SpGEMM call needs between its symbolic and numeric phase.  Lives in the
// context's arena; replaces sfBIN (nsparse.h:110-121), which the reference rebuilds per call.
struct nsp_spgemm_state {
    int M = 0, K = 0, N = 0;
    int *d_row_cnt = nullptr;     // [M+1] nnz(C_i) after the symbolic phase
    int *d_row_ip = nullptr;      // [M+1] intermediate products of row i (saturated at INT_MAX)
    int *d_row_perm = nullptr;    // [M]   rows grouped by bin, heaviest bin first
    int *d_bins = nullptr;        // see spgemm_plan.h for the layout
    unsigned long long *d_binsum = nullptr;
    int *h_bins = nullptr;        // pinned mirrors
    unsigned long long *h_binsum = nullptr;
    long long *d_scalars = nullptr;   // [8]  total ip, nnz, ...
    long long *d_scan_tmp = nullptr;  // block sums of the row-pointer scan
    long long *h_scalars = nullptr;   // pinned mirror
    bool symbolic_done = false;
    long long b_nnz = 0;          // nnz(B) = B.rpt[K], read back by the symbolic plan
    long long a_nnz = 0;          // entries of A (sum of the row lengths seen by the symbolic plan)
    bool join_pending = false;    // the side-stream launch of this phase has not been joined yet
    bool has_multi_slab = true;   // some row of A has more than 1024 entries (second launch of the heavy numeric kernel)
    bool b_sorted = true;         // rows of B column-sorted (checked by the symbolic plan)
};

struct nsp_host_result {
    // C kept on the device by nsp_spgemm_host_*
    long long *d_rpt64 = nullptr;
    int *d_col = nullptr;
    void *d_val = nullptr;
    int M = 0;
    long long nnz = 0;
    int val_bytes = 0;
    // device copies of the inputs
    int *d_a_rpt = nullptr, *d_a_col = nullptr, *d_b_rpt = nullptr, *d_b_col = nullptr;
    void *d_a_val = nullptr, *d_b_val = nullptr;
    size_t a_nnz_cap = 0, b_nnz_cap = 0;
    int a_m_cap = 0, b_m_cap = 0;
    int in_val_bytes = 0;
    int *d_cut_rows = nullptr;        // row cuts of nsp_spgemm_host_stream_* (65 entries each)
    long long *d_cut_offs = nullptr;
};

// How the lanes of an AMB matrix write y (amb_convert.cu builds it, amb_spmv.cu uses it; keyed by the matrix's d_cs
// so that neither sfAMB nor nsp_amb -- fixed layouts -- has to carry it).  Per chunk two bits per lane: 0 = the lane
// pads the chunk (nothing to write), 1 = its row has entries in several column segments (red.global into a zeroed
// y), 2 = the lane holds the WHOLE row (plain store, no zeroing, no read-modify-write).  zero_rows: the rows that
// are not written by a plain store (several virtual rows, or none at all).
struct nsp_amb_plan {
    unsigned long long *d_mode = nullptr;
    int *d_zero_rows = nullptr;
    int n_zero_rows = 0;
    int M = 0, c_size = 0;        // of the matrix the plan was made for (checked before use)
};

// copy-engine gather of the multi-GPU SpGEMM (peer_dma.cu)
struct nsp_dma_push {
    int *d_tile_cnt = nullptr;
    int *h_done = nullptr, *d_done = nullptr;     // tile flags: pinned host memory and its device alias
    size_t cap = 0;
    cudaStream_t copy_st[nsp::kMaxPeerOut] = {};
    cudaEvent_t ev_copy[nsp::kMaxPeerOut] = {};
    // at most kSlots batches of copies are queued per peer (ev_slot: the end of each), so that what is left when the
    // kernels end can go out through SM stores instead (ev_sm: the ends of those launches)
    static constexpr int kSlots = 4, kSmSlots = 2;
    cudaEvent_t ev_slot[nsp::kMaxPeerOut][kSlots] = {};
    cudaEvent_t ev_sm[kSmSlots] = {};
    long long last_ce_tiles = 0, last_sm_tiles = 0;   // of the last gather: tiles sent by copy engines / by SM stores
    std::chrono::steady_clock::time_point t_numeric;  // entry of the numeric phase
    double last_kernel_ms = 0;                        // from there to the end of its kernels, as seen by the polling thread
    char *d_sort = nullptr;                       // keys / values / temporary storage of order_rows_by_tile
    size_t sort_bytes = 0;
    bool active = false;
};

struct nsp_prof_rec {
    std::string name;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    long long rows = 0, ip = 0, alen = 0, out = 0;
};

struct nsp_context {
    int device = 0;
    cudaStream_t stream = nullptr;   // nullptr = legacy default stream
    // side stream + events for the launch that runs next to the main row-class kernels (the long rows of the
    // heavy numeric class); forked from / joined into `stream`, so callers still see one stream
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaMemPool_t mem_pool = nullptr;   // stream-ordered pool of the AMB conversion's temporaries (kept between calls)
    int sm_count = 148;
    int max_smem_optin = nsp::kMaxSmemOptin;
    size_t l2_bytes = 0;

    // grow-only arena for plan / workspace arrays
    char *arena = nullptr;
    size_t arena_bytes = 0;
    size_t arena_used = 0;

    // window cuts per entry of A for the segment mode of the heavy numeric kernel (spgemm_plan.cu)
    int *d_seg = nullptr;
    size_t seg_cap = 0;
    long long opt_unsorted = 0;          // 1: sort = false numeric mode, the hash classes skip the per-row column sort
    long long opt_no_flat = 0;           // 1: never the flat traversal of the heavy kernels; -1: always (tests)
    long long opt_no_seg = 0;            // 1: never use the segment mode (tests, A/B measurements)

    // options (nsp_set_option)
    long long opt_sym_bitmap_min = -1;   // rows with min(ip,N) >  this go to the bitmap kernel (-1: default)
    long long opt_num_bitmap_min = -1;   // rows with nnz(C_i)  >  this go to the bitmap-rank kernel
    long long opt_sym_window_shift = 0;  // log2 of the symbolic bitmap window (0: default 20)
    long long opt_num_window_shift = 0;  // log2 of the numeric bitmap window (0: default 19)
    long long opt_num_cap = 0;           // > 0: upper limit of the accumulator chunk (tests)
    long long opt_no_fork = 0;           // 1: the long rows of the heavy numeric class run on the main stream
    long long opt_no_vec = 0;            // 1: never read B.col with 128-bit loads (tests)
    long long opt_phase_timing = 0;      // 1: the heavy numeric kernel accumulates cycles per phase (development)
    long long *d_phase = nullptr;
    long long *phase_cycles()
    {
        if (!d_phase) {
            cudaMalloc((void **)&d_phase, 12 * sizeof(long long));
            cudaMemset(d_phase, 0, 12 * sizeof(long long));
        }
        return d_phase;
    }
    long long opt_debug = 0;             // development only: bit 0 skip emit, 1 skip value pass, 2 skip zero-fill

    nsp_spgemm_state sp;
    nsp::PeerOut peer_out;   // nsp_spgemm_set_peers
    nsp_dma_push dma;
    long long opt_gather_tma = 0;        // 1: the TMA pusher kernel (peer_push.cu) instead of the copy engines
    long long opt_dma_tile_log = 0;      // > 0: log2 of the copy-engine tile (tests)
    long long opt_no_ranges = 0;         // 1: never num_hash_ranges_kernel (wide C, short B rows); -1: whenever it can run
    long long opt_hash_order = 0;        // 1: bitonic table sort always; 2: no shared-memory bucket ordering (measurements)
    long long opt_gather_sm = 1;         // 1: tiles still unsent when the kernels end leave through SM stores next to
                                         // the copy engines; 0: copy engines only; 2: SM stores only (tests)
    nsp::PeerOut last_push;  // the tile hand-off of the last product (diagnostics of nsp_spgemm_peers_status)
    // tile hand-off + pusher kernel of the multi-GPU allgatherv (peer_push.cu)
    cudaStream_t push_stream = nullptr;
    cudaEvent_t ev_push_fork = nullptr, ev_push_join = nullptr;
    int *d_push_ws = nullptr;            // [8 control ints][cap tile counters][cap queue slots]
    size_t push_cap = 0;
    bool push_active = false;
    int push_ctas = 0;
    bool peers_preloaded[2] = {false, false};   // numeric kernels loaded ahead of the pusher (fp32, fp64)
    long long opt_push_sms = 0;          // CTAs (= SMs) of the pusher kernel (0: default 16)
    nsp_host_result host;
    std::unordered_map<const void *, nsp_amb_plan> amb_plans;
    char *d_spmv_stage = nullptr;        // x / y of nsp_spmv_amb_host_* (grow-only)
    size_t spmv_stage_bytes = 0;

    long long launches = 0;
    std::string err;

    // per-launch CUDA-event timing of the row-class kernels (nsp_set_option("profile", 1))
    bool profile = false;
    std::vector<nsp_prof_rec> prof;
    cudaStream_t prof_stream = nullptr;
    bool prof_on_aux = false;
    void prof_begin(const char *name, long long rows, long long ip, long long alen, long long out = 0)
    {
        if (!profile) return;
        const cudaStream_t stream = prof_on_aux ? aux_stream : this->stream;
        nsp_prof_rec r;
        r.name = name;
        r.rows = rows;
        r.ip = ip;
        r.alen = alen;
        r.out = out;
        cudaEventCreate(&r.e0);
        cudaEventCreate(&r.e1);
        cudaEventRecord(r.e0, stream);
        prof.push_back(r);
    }
    void prof_end()
    {
        if (!profile || prof.empty()) return;
        cudaEventRecord(prof.back().e1, prof_on_aux ? aux_stream : stream);
    }

    int fail(int code, const std::string &msg)
    {
        err = msg;
        return code;
    }

    // Reserve `bytes` in the arena (256-byte aligned).  reset_arena() must have been called at
    // the start of the operation; the arena is re-allocated (after a device sync) when too small.
    int arena_reserve(size_t total_bytes);
    void arena_reset() { arena_used = 0; }
    template <typename T>
    T *arena_take(size_t count)
    {
        size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
        if (arena_used + bytes > arena_bytes) return nullptr;
        T *p = reinterpret_cast<T *>(arena + arena_used);
        arena_used += bytes;
        return p;
    }
};

namespace nsp {

int context_create(nsp_context **out, int device);
int context_destroy(nsp_context *ctx);

// ---- SpGEMM core (spgemm_plan.cu, spgemm_symbolic.cu, spgemm_numeric.cuh) -------------------
int spgemm_flop(nsp_context *ctx, int M, const int *a_rpt, const int *a_col, const int *b_rpt,
                long long *h_flop);
int spgemm_symbolic(nsp_context *ctx, int M, int K, int N, const int *a_rpt, const int *a_col,
                    const int *b_rpt, const int *b_col, long long *c_rpt64, long long *h_nnz,
                    long long *h_ip);
template <typename real>
int spgemm_numeric(nsp_context *ctx, int M, int K, int N, const int *a_rpt, const int *a_col,
                   const real *a_val, const int *b_rpt, const int *b_col, const real *b_val,
                   const long long *c_rpt64, int *c_col, real *c_val, int row0 = 0, int nrows = -1);
// everything the numeric phase with peers would allocate or load lazily, done ahead (single-process multi-GPU:
// see peer_push_reserve)
template <typename real>
int spgemm_numeric_reserve(nsp_context *ctx, int N, long long a_nnz, long long nnz_block, int rows, int npeers);
int rpt64_to_rpt32(nsp_context *ctx, int M, const long long *rpt64, long long nnz, int *rpt32);
// peer_push.cu
int peer_push_reserve(nsp_context *ctx, long long ntiles);
// peer_dma.cu
int dma_tile_log(long long nnz);
int peer_dma_reserve(nsp_context *ctx, long long ntiles, int npeers, int max_rows);
int peer_dma_begin(nsp_context *ctx, long long nnz_block);
int peer_dma_drive(nsp_context *ctx, const int *c_col_full, const void *c_val_full, int val_bytes);
int order_rows_by_tile(nsp_context *ctx, int *row_perm, int n, const long long *c_rpt, long long off, int tile_log);
void peer_dma_destroy(nsp_context *ctx);
int peer_push_begin(nsp_context *ctx, const int *c_col_full, const void *c_val_full, int val_bytes, long long nnz_block);
int peer_push_end(nsp_context *ctx);

}  // namespace nsp
