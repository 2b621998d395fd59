// nsparse-b200: synthetic input generators (host side, OpenMP).
//
// Counter-based: the quadrant of edge e at Kronecker level l depends only on (seed, e, l), so the
// stream is identical for any thread count, on every rank, and in the numpy mirror
// (nsparse_b200/gen.py: rmat_edges_numpy).  Graph500 parameters (a,b,c,d) = (.57,.19,.19,.05),
// no vertex permutation, no noise.
#include <stdint.h>

#include "../../include/nsparse_b200.h"

static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

extern "C" int nsp_gen_rmat_edges(int scale, long long n_edges, unsigned long long seed, long long *h_src,
                                  long long *h_dst)
{
    if (scale < 1 || scale > 40 || n_edges < 0 || !h_src || !h_dst) return NSP_ERR_ARG;
    // thresholds on a 32-bit draw: a, a+b, a+b+c
    const uint32_t ta = 2448131358u, tab = 3264175144u, tabc = 4080218931u;
    const uint64_t key = splitmix64(seed);
#pragma omp parallel for schedule(static)
    for (long long e = 0; e < n_edges; ++e) {
        uint64_t src = 0, dst = 0;
        for (int l = 0; l < scale; ++l) {
            const uint32_t r = (uint32_t)(splitmix64(key ^ ((uint64_t)e * 64ull + (uint64_t)l)) >> 32);
            const uint64_t sbit = r >= tab;                    // quadrants c, d -> lower half
            const uint64_t dbit = (r >= ta && r < tab) || r >= tabc;   // quadrants b, d -> right half
            src = (src << 1) | sbit;
            dst = (dst << 1) | dbit;
        }
        h_src[e] = (long long)src;
        h_dst[e] = (long long)dst;
    }
    return NSP_OK;
}
