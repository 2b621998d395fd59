// nsparse-b200: context life cycle and the workspace arena.
#include "context.h"

#include <stdlib.h>

#include "../../include/nsparse_b200.h"

int nsp_context::arena_reserve(size_t total_bytes)
{
    if (total_bytes <= arena_bytes) return 0;
    // growing: nothing of ours may still be running on the old block
    cudaError_t e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess && aux_stream) e = cudaStreamSynchronize(aux_stream);
    sp.join_pending = false;
    if (e != cudaSuccess) return fail(-1, std::string("arena sync: ") + cudaGetErrorString(e));
    if (arena) cudaFree(arena);
    arena = nullptr;
    arena_bytes = 0;
    size_t want = total_bytes + total_bytes / 8 + (1u << 20);
    e = cudaMalloc((void **)&arena, want);
    if (e != cudaSuccess) return fail(-4, std::string("arena cudaMalloc: ") + cudaGetErrorString(e));
    arena_bytes = want;
    sp.symbolic_done = false;
    return 0;
}

namespace nsp {

int context_create(nsp_context **out, int device)
{
    if (!out) return -2;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        fprintf(stderr, "nsparse_b200: no CUDA device -- this library has no CPU fallback\n");
        return -1;
    }
    if (device < 0 || device >= count) return -2;
    if (cudaSetDevice(device) != cudaSuccess) return -1;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
    // the library holds sm_100a SASS only (arch-specific code is not forward compatible: sm_103 / sm_120 would
    // pass a "major >= 10" test and then fail at the first launch with "no kernel image")
    if (prop.major != 10 || prop.minor != 0) {
        fprintf(stderr, "nsparse_b200: device %d is sm_%d%d; this build targets sm_100a (B200) only\n", device,
                prop.major, prop.minor);
        return -1;
    }
    if ((int)prop.sharedMemPerBlockOptin < nsp::kMaxSmemOptin) {
        fprintf(stderr, "nsparse_b200: device %d offers %zu bytes of opt-in shared memory per block, the kernels "
                        "are laid out for %d\n", device, (size_t)prop.sharedMemPerBlockOptin, nsp::kMaxSmemOptin);
        return -1;
    }
    nsp_context *ctx = new nsp_context();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    ctx->l2_bytes = (size_t)prop.l2CacheSize;
    ctx->stream = nullptr;
    // The side stream of the long rows gets the HIGHEST priority: its launch and the main launch of the heavy class
    // become runnable at the same moment (both wait for the same predecessor), and which of the two the block
    // scheduler served first used to depend on such things as an event record queued in between -- with the main
    // launch first, the few long rows ran last and alone (measured: 80 ms instead of 25 ms for their launch, and no
    // tile of C finished before the end, profiles/r2_trace_dma_order_priority.txt).
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    if (cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, prio_greatest) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        delete ctx;
        return -1;
    }
    {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        if (cudaMemPoolCreate(&ctx->mem_pool, &props) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(ctx->mem_pool, cudaMemPoolAttrReleaseThreshold, &keep);
        } else {
            ctx->mem_pool = nullptr;       // fall back to the device's default pool
            cudaGetLastError();
        }
    }
    // NSP_OPTIONS="name=value,name=value": nsp_set_option for processes that cannot call it (the unchanged sample
    // drivers under compute-sanitizer, forced windows x chunks x slabs cases)
    if (const char *env = getenv("NSP_OPTIONS")) {
        std::string e(env);
        size_t pos = 0;
        while (pos < e.size()) {
            size_t end = e.find(',', pos);
            if (end == std::string::npos) end = e.size();
            const std::string kv = e.substr(pos, end - pos);
            const size_t eq = kv.find('=');
            if (eq != std::string::npos && nsp_set_option(ctx, kv.substr(0, eq).c_str(), atoll(kv.c_str() + eq + 1)) != 0)
                fprintf(stderr, "nsparse_b200: NSP_OPTIONS: unknown option '%s'\n", kv.substr(0, eq).c_str());
            pos = end + 1;
        }
    }
    *out = ctx;
    return 0;
}

int context_destroy(nsp_context *ctx)
{
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->aux_stream) {
        cudaStreamSynchronize(ctx->aux_stream);
        cudaStreamDestroy(ctx->aux_stream);
    }
    if (ctx->push_stream) {
        cudaStreamSynchronize(ctx->push_stream);
        cudaStreamDestroy(ctx->push_stream);
        cudaEventDestroy(ctx->ev_push_fork);
        cudaEventDestroy(ctx->ev_push_join);
    }
    cudaFree(ctx->d_push_ws);
    nsp::peer_dma_destroy(ctx);
    cudaFree(ctx->d_seg);
    cudaFree(ctx->d_spmv_stage);
    for (auto &kv : ctx->amb_plans) {
        cudaFree(kv.second.d_mode);
        cudaFree(kv.second.d_zero_rows);
    }
    ctx->amb_plans.clear();
    if (ctx->mem_pool) cudaMemPoolDestroy(ctx->mem_pool);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->arena) cudaFree(ctx->arena);
    if (ctx->d_phase) cudaFree(ctx->d_phase);
    for (auto &r : ctx->prof) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    ctx->prof.clear();
    if (ctx->sp.h_scalars) cudaFreeHost(ctx->sp.h_scalars);
    if (ctx->sp.h_bins) cudaFreeHost(ctx->sp.h_bins);
    if (ctx->sp.h_binsum) cudaFreeHost(ctx->sp.h_binsum);
    nsp_host_result &h = ctx->host;
    cudaFree(h.d_rpt64);
    cudaFree(h.d_col);
    cudaFree(h.d_val);
    cudaFree(h.d_a_rpt);
    cudaFree(h.d_a_col);
    cudaFree(h.d_a_val);
    cudaFree(h.d_b_rpt);
    cudaFree(h.d_b_col);
    cudaFree(h.d_b_val);
    cudaFree(h.d_cut_rows);
    cudaFree(h.d_cut_offs);
    delete ctx;
    return 0;
}

}  // namespace nsp
