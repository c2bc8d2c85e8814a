// Issue-rate micro-benchmarks of the integer pipe: the denominators of the POPC / LOP3 roofline
// (SURVEY.md 8d asks for a measured POPC peak next to MEASURED_PEAKS.json).
#pragma once
#include "plm_common.cuh"

namespace plm {

constexpr int MICRO_CHAINS = 8;
constexpr int MICRO_UNROLL = 16;
constexpr double MICRO_OPS_PER_ITER = double(MICRO_CHAINS) * MICRO_UNROLL;

// MICRO_CHAINS independent dependency chains of POPC per thread: x <- popc(x) ^ c is two
// instructions, so the POPC is paired with a LOP3 that runs on the (4x wider) ALU lanes; what is
// timed is the POPC issue rate.
__global__ void __launch_bounds__(256) popc_rate_kernel(uint32_t *out, int iters) {
    uint32_t x[MICRO_CHAINS];
#pragma unroll
    for (int c = 0; c < MICRO_CHAINS; ++c) x[c] = threadIdx.x * 2654435761u + c * 40503u + blockIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < MICRO_UNROLL; ++u) {
#pragma unroll
            for (int c = 0; c < MICRO_CHAINS; ++c) {
                uint32_t p;
                asm volatile("popc.b32 %0, %1;" : "=r"(p) : "r"(x[c]));
                x[c] = p ^ x[(c + 1) % MICRO_CHAINS];
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < MICRO_CHAINS; ++c) s += x[c];
    if (s == 0xDEADBEEFu) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) lop3_rate_kernel(uint32_t *out, int iters) {
    uint32_t x[MICRO_CHAINS];
#pragma unroll
    for (int c = 0; c < MICRO_CHAINS; ++c) x[c] = threadIdx.x * 2654435761u + c * 40503u + blockIdx.x;
    const uint32_t k = blockIdx.x | 1u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < MICRO_UNROLL; ++u) {
#pragma unroll
            for (int c = 0; c < MICRO_CHAINS; ++c) {
                uint32_t r;
                asm volatile("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(x[c]), "r"(x[(c + 1) % MICRO_CHAINS]), "r"(k));
                x[c] = r;
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < MICRO_CHAINS; ++c) s += x[c];
    if (s == 0xDEADBEEFu) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

} // namespace plm
