// Microbenchmark: per-SM throughput of the candidate "accumulate one product" primitives on B200.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/atomics_bench scripts/micro/atomics_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int kSmemWords = 32768;   // 128 KiB of 4-byte words

enum Op { LDS4 = 0, STS1, ATOMS_OR, ATOMS_OR_TEST, ATOMS_ADD_F32, ATOMS_ADD_F64, RMW_F32, RMW_F64, REDG_F32, REDG_F64,
          REDG_OR, STG4, REDG_V4, LDS8_PLUS_LDS4, ATOMS_ADD_I32, REDG_F32_SORTED, ATOMS_F32_DENSE, ATOMS_F32_HOT, ATOMS_OR_SKEW, ATOMS_OR_SKEW_SWZ, NOPS };
const char *names[] = {"lds.32 random", "sts.u8 random", "atoms.or random", "test+atoms.or (50% set)", "atomicAdd f32 smem",
                       "atomicAdd f64 smem", "lds+fadd+sts f32 (racy)", "lds+dadd+sts f64 (racy)", "red.global f32 random(1MB/CTA)",
                       "red.global f64 random", "red.global.or random", "st.global.32 random", "red.global.v4.f32 random",
                       "lds.64 + lds.32 (rank lookup)", "atomicAdd i32 smem", "red.global f32 ascending (sorted ranks ~20 apart)",
                       "atomicAdd f32 smem, 32 consecutive slots per warp", "atomicAdd f32 smem, 12% of ops on 16 hot slots",
                       "atoms.or, R-MAT-skewed columns, transposed layout", "atoms.or, R-MAT-skewed columns, swizzled layout"};

template <int OP>
__global__ void __launch_bounds__(1024, 1) bench(int iters, float *g, unsigned *out)
{
    extern __shared__ __align__(16) unsigned smem[];
    const int t = threadIdx.x;
    for (int i = t; i < kSmemWords; i += 1024) smem[i] = (OP == ATOMS_OR_TEST) ? 0x55555555u : 0u;
    __syncthreads();
    unsigned h = (blockIdx.x * 1024 + t) * 2654435761u + 12345u;
    unsigned acc = 0;
    float *gb = g + (size_t)blockIdx.x * (1 << 18);   // 1 MiB of floats per CTA
    unsigned asc = t * 20;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        h = h * 1664525u + 1013904223u;
        const unsigned r = h >> 8;
        if (OP == LDS4) acc += smem[r & (kSmemWords - 1)];
        if (OP == STS1) reinterpret_cast<unsigned char *>(smem)[r & (kSmemWords * 4 - 1)] = 1;
        if (OP == ATOMS_OR) atomicOr(&smem[r & (kSmemWords - 1)], 1u << (r >> 20 & 31));
        if (OP == ATOMS_OR_TEST) {
            unsigned *w = &smem[r & (kSmemWords - 1)];
            const unsigned bit = 1u << (r >> 20 & 31);
            if (!(*(volatile unsigned *)w & bit)) acc += !(atomicOr(w, bit) & bit);
        }
        if (OP == ATOMS_ADD_F32) atomicAdd(reinterpret_cast<float *>(smem) + (r & (kSmemWords - 1)), 1.0f);
        if (OP == ATOMS_ADD_I32) atomicAdd(reinterpret_cast<int *>(smem) + (r & (kSmemWords - 1)), 1);
        if (OP == ATOMS_ADD_F64) atomicAdd(reinterpret_cast<double *>(smem) + (r & (kSmemWords / 2 - 1)), 1.0);
        if (OP == RMW_F32) {
            volatile float *p = reinterpret_cast<float *>(smem) + (r & (kSmemWords - 1));
            *p = *p + 1.0f;
        }
        if (OP == RMW_F64) {
            volatile double *p = reinterpret_cast<double *>(smem) + (r & (kSmemWords / 2 - 1));
            *p = *p + 1.0;
        }
        if (OP == REDG_F32) atomicAdd(gb + (r & ((1 << 18) - 1)), 1.0f);
        if (OP == REDG_F32_SORTED) {
            atomicAdd(gb + (asc & ((1 << 18) - 1)), 1.0f);
            asc += 1024 * 20;
        }
        if (OP == REDG_F64) atomicAdd(reinterpret_cast<double *>(gb) + (r & ((1 << 17) - 1)), 1.0);
        if (OP == REDG_OR) atomicOr(reinterpret_cast<unsigned *>(gb) + (r & ((1 << 18) - 1)), 1u << (r >> 20 & 31));
        if (OP == STG4) gb[r & ((1 << 18) - 1)] = 1.0f;
        if (OP == REDG_V4) {
            float *p = gb + (r & ((1 << 18) - 4));
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(1.0f), "f"(0.0f), "f"(1.0f), "f"(0.0f) : "memory");
        }
        if (OP == ATOMS_F32_DENSE) {
            unsigned hw = ((blockIdx.x * 32 + (t >> 5)) * 2654435761u + i) * 1664525u + 1013904223u;
            atomicAdd(reinterpret_cast<float *>(smem) + (((hw >> 8) + (t & 31)) & (kSmemWords - 1)), 1.0f);
        }
        if (OP == ATOMS_F32_HOT) {
            const unsigned a = (r & 7) == 0 ? ((r >> 3) & 15) : (r & (kSmemWords - 1));
            atomicAdd(reinterpret_cast<float *>(smem) + a, 1.0f);
        }
        if (OP == ATOMS_OR_SKEW || OP == ATOMS_OR_SKEW_SWZ) {
            // column with every bit 0 w.p. ~0.75 (AND of two random words), 20 bits
            unsigned h2 = h * 22695477u + 1u;
            const unsigned cc = (h >> 8) & (h2 >> 8) & ((1u << 20) - 1);
            const unsigned blk = cc >> 10;
            const unsigned sw = OP == ATOMS_OR_SKEW_SWZ ? ((blk ^ (blk >> 5)) & 31u) : 0u;
            atomicOr(&smem[(blk << 5) | ((cc ^ sw) & 31u)], 1u << ((cc >> 5) & 31u));
        }
        if (OP == LDS8_PLUS_LDS4) {
            const uint2 w = reinterpret_cast<uint2 *>(smem)[r & (kSmemWords / 4 - 1)];
            acc += __popc(w.x) + __popc(w.y) + smem[kSmemWords / 2 + (r & (kSmemWords / 4 - 1))];
        }
    }
    __syncthreads();
    if (acc == 0xdeadbeef || OP != LDS4) out[blockIdx.x * 1024 + t] = acc + smem[t];
}

template <int OP>
void run(float *g, unsigned *out, int sms, double mhz)
{
    const int iters = 4096;
    CK(cudaFuncSetAttribute(bench<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemWords * 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    bench<OP><<<sms, 1024, kSmemWords * 4>>>(iters, g, out);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    bench<OP><<<sms, 1024, kSmemWords * 4>>>(iters, g, out);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)sms * 1024 * iters;
    const double cyc_per_warp_instr = ms * 1e-3 * mhz * 1e6 / (32.0 * iters);   // per SM: 32 warps x iters warp-instr
    printf("%-52s %8.3f ms  %8.1f Gop/s  %6.2f cyc/warp-instr/SM  %6.3f cyc/lane\n", names[OP], ms, ops / ms / 1e6,
           cyc_per_warp_instr, cyc_per_warp_instr / 32.0);
}

int main()
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double mhz = clk / 1000.0;
    printf("%s  SMs=%d  clock=%.0f MHz (nominal; cycles assume it)\n", p.name, p.multiProcessorCount, mhz);
    float *g;
    unsigned *out;
    CK(cudaMalloc(&g, (size_t)p.multiProcessorCount * (1 << 20)));
    CK(cudaMemset(g, 0, (size_t)p.multiProcessorCount * (1 << 20)));
    CK(cudaMalloc(&out, (size_t)p.multiProcessorCount * 4096));
    const int s = p.multiProcessorCount;
    run<LDS4>(g, out, s, mhz);
    run<STS1>(g, out, s, mhz);
    run<LDS8_PLUS_LDS4>(g, out, s, mhz);
    run<ATOMS_OR>(g, out, s, mhz);
    run<ATOMS_OR_TEST>(g, out, s, mhz);
    run<ATOMS_ADD_I32>(g, out, s, mhz);
    run<ATOMS_ADD_F32>(g, out, s, mhz);
    run<ATOMS_ADD_F64>(g, out, s, mhz);
    run<RMW_F32>(g, out, s, mhz);
    run<RMW_F64>(g, out, s, mhz);
    run<REDG_F32>(g, out, s, mhz);
    run<REDG_F32_SORTED>(g, out, s, mhz);
    run<REDG_F64>(g, out, s, mhz);
    run<REDG_OR>(g, out, s, mhz);
    run<STG4>(g, out, s, mhz);
    run<REDG_V4>(g, out, s, mhz);
    run<ATOMS_F32_DENSE>(g, out, s, mhz);
    run<ATOMS_F32_HOT>(g, out, s, mhz);
    run<ATOMS_OR_SKEW>(g, out, s, mhz);
    run<ATOMS_OR_SKEW_SWZ>(g, out, s, mhz);
    return 0;
}
