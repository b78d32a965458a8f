// Shared device helpers for the morpheus_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/morpheus_b200.h"

namespace mb {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

__host__ __device__ inline uint32_t div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

// ---- hash-grid indexing (restates gridencoder.cu:46-79; uint32 wrap-around arithmetic) ---------
template <uint32_t D>
__device__ __forceinline__ uint32_t grid_index(uint32_t gridtype, uint32_t hashmap_size, uint32_t res,
                                               const uint32_t (&pg)[D]) {
    uint32_t stride = 1, index = 0;
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        if (stride <= hashmap_size) {
            index += pg[d] * stride;
            stride *= res;
        }
    }
    if (gridtype == 0 && stride > hashmap_size) {
        constexpr uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
        index = 0;
#pragma unroll
        for (uint32_t d = 0; d < D; d++) index ^= pg[d] * primes[d];
    }
    return index % hashmap_size;
}

// resolution rule of gridencoder.cu:133, all in float32
__device__ __forceinline__ uint32_t level_resolution(uint32_t level, float S, uint32_t H) {
    return (uint32_t)ceilf(exp2f((float)level * S) * (float)H);
}

// per-axis cell location, gridencoder.cu:140-160.  Returns fractional pos, writes pg and deriv.
__device__ __forceinline__ float locate(float x, uint32_t res, bool align_corners, uint32_t interp, uint32_t& pg,
                                        float& deriv) {
    float pos;
    if (align_corners) {
        pos = __fmul_rn(x, (float)(res - 1));
        pg = min((uint32_t)floorf(pos), res - 2);
    } else {
        pos = fminf(fmaxf(__fmaf_rn(x, (float)res, -0.5f), 0.0f), (float)(res - 1));
        pg = (uint32_t)floorf(pos);
    }
    pos = __fsub_rn(pos, (float)pg);
    if (interp == 1) {
        deriv = 6.0f * pos * (1.0f - pos);
        pos = pos * pos * (3.0f - 2.0f * pos);
    } else {
        deriv = 1.0f;
    }
    return pos;
}

__device__ __forceinline__ void red_add(float* addr, float v) { atomicAdd(addr, v); }  // result unused -> RED
__device__ __forceinline__ void red_add2(float* addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// cp.async helpers (LDGSTS)
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

}  // namespace mb
