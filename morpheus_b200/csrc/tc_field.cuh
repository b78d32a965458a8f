// Shared pieces of the tensor-core field kernels: operand-tile writers and input builders.
#pragma once
#include "field_common.cuh"
#include "tc_common.cuh"

namespace mb {
namespace tc {

constexpr int A_LO_OFF = 32768;         // lo-part offset (bytes) of a 128 x 128 fp16 operand tile

// `pitch` = bytes between 8-column core groups = 16 * rows of the tile (2048 for 128-row tiles, 1536 for 96-row tiles)
__device__ __forceinline__ void store_core(uint8_t* A, int m, int kc, const float (&v)[8], int lo_off = A_LO_OFF, int pitch = 2048) {
    uint4 hi, lo;
    split8(v, hi, lo);
    uint8_t* p = A + kc * pitch + (m >> 3) * 128 + (m & 7) * 16;
    *reinterpret_cast<uint4*>(p) = hi;
    *reinterpret_cast<uint4*>(p + lo_off) = lo;
}

// freq encoding in the tc feature order: cores 0..4 = [p(3), sin/cos bands (36), 0]
__device__ __forceinline__ void build_freq_tc(uint8_t* A, int m, const float p[3], int n_freq, int lo_off = A_LO_OFF) {
    float f[40];
    f[0] = p[0]; f[1] = p[1]; f[2] = p[2];
    float fr = 1.0f;
#pragma unroll
    for (int k = 0; k < 6; k++) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            float s = 0.f, c = 0.f;
            if (k < n_freq) sincosf(p[a] * fr, &s, &c);
            f[3 + 6 * k + a] = s;
            f[6 + 6 * k + a] = c;
        }
        fr *= 2.0f;
    }
    f[39] = 0.f;
#pragma unroll
    for (int c = 0; c < 5; c++) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = f[c * 8 + i];
        store_core(A, m, c, v, lo_off);
    }
}

// 4 grid levels (8 features) -> one core
__device__ __forceinline__ void build_grid_core_tc(uint8_t* A, int m, int kc, const GridCtx& g, int level0, const float p[3], int lo_off = A_LO_OFF) {
    float u[3];
#pragma unroll
    for (int d = 0; d < 3; d++) u[d] = __fdiv_rn(__fadd_rn(p[d], g.bound), g.two_bound);
    float v[8];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        float feat[2] = {0.f, 0.f};
        if ((uint32_t)(level0 + j) < g.n_levels) grid_eval(g, level0 + j, u, feat, nullptr);
        v[2 * j] = feat[0];
        v[2 * j + 1] = feat[1];
    }
    store_core(A, m, kc, v, lo_off);
}

// ---- rolled builders (small code footprint: the fused kernels are instruction-fetch bound when these are unrolled) ----------
// one feature (2-byte hi + 2-byte lo) / an even-aligned feature pair (4-byte stores) at tc column k of row m
__device__ __forceinline__ void store_one(uint8_t* A, int m, int k, float f, int lo_off = A_LO_OFF, int pitch = 2048) {
    const __half h = __float2half_rn(f);
    const __half l = __float2half_rn(f - __half2float(h));
    uint8_t* p = A + (k >> 3) * pitch + (m >> 3) * 128 + (m & 7) * 16 + (k & 7) * 2;
    *reinterpret_cast<__half*>(p) = h;
    *reinterpret_cast<__half*>(p + lo_off) = l;
}
__device__ __forceinline__ void store_pair(uint8_t* A, int m, int k, float f0, float f1, int lo_off = A_LO_OFF, int pitch = 2048) {
    const __half2 hh = __floats2half2_rn(f0, f1);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(f0 - hf.x, f1 - hf.y);
    uint8_t* p = A + (k >> 3) * pitch + (m >> 3) * 128 + (m & 7) * 16 + (k & 7) * 2;
    *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<const uint32_t*>(&hh);
    *reinterpret_cast<uint32_t*>(p + lo_off) = *reinterpret_cast<const uint32_t*>(&ll);
}
// frequency features of ONE axis a of point coordinate v: tc columns a, 3+6k+a (sin), 6+6k+a (cos)
__device__ __forceinline__ void freq_axis_tc(uint8_t* A, int m, int a, float v, int n_freq, int lo_off = A_LO_OFF, int pitch = 2048) {
    store_one(A, m, a, v, lo_off, pitch);
    float fr = 1.0f;
#pragma unroll 1
    for (int k = 0; k < 6; k++) {
        float s = 0.f, c = 0.f;
        if (k < n_freq) sincosf(v * fr, &s, &c);
        store_one(A, m, 3 + 6 * k + a, s, lo_off, pitch);
        store_one(A, m, 6 + 6 * k + a, c, lo_off, pitch);
        fr *= 2.0f;
    }
}
// grid levels [l0, l0+nl) at world point p -> tc columns kbase + 2l, kbase + 2l + 1 (same arithmetic as grid_eval)
__device__ __forceinline__ void gather_levels_tc(uint8_t* A, int m, int kbase, const GridCtx& g, int l0, int nl, const float p[3], int lo_off = A_LO_OFF,
                                                 int pitch = 2048) {
    float u[3];
#pragma unroll
    for (int d = 0; d < 3; d++) u[d] = __fdiv_rn(__fadd_rn(p[d], g.bound), g.two_bound);
#pragma unroll 1
    for (int l = l0; l < l0 + nl; l++) {
        float feat[2] = {0.f, 0.f};
        if ((uint32_t)l < g.n_levels) grid_eval(g, l, u, feat, nullptr);
        store_pair(A, m, kbase + 2 * l, feat[0], feat[1], lo_off, pitch);
    }
}

__device__ __forceinline__ float code_value(const mb_field_params& p, int v, int c, float t) {
    const int S = (int)p.code_len[v];
    t = fminf(fmaxf(t, 0.f), 1.f);
    const float g = __fsub_rn(__fmul_rn(t, 2.f), 1.f);
    const float pos = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), (float)(S - 1));
    int i0 = min(max((int)floorf(pos), 0), S - 1);
    const float w1 = pos - (float)i0, w0 = 1.f - w1;
    const float* line = p.code[v] + (size_t)c * S;
    float val = __ldg(line + i0) * w0;
    if (i0 + 1 <= S - 1) val += __ldg(line + i0 + 1) * w1;
    return val;
}


// original input-feature index of tc feature k (first layers are core-aligned; -1 = zero pad)
//   kind 1 (deform/topo L0): 0..38 -> k, 39 pad, 40..87 -> 39 + (k-40) (code), 88..95 pad
//   kind 2 (sdf L0)        : 0..38 -> k, 39 pad, 40..71 -> 39 + (k-40) (grid), 72,73 -> 71,72 (topo), 74..79 pad
__device__ __forceinline__ int tc_korig(int kind, int k) {
    if (kind == 1) { if (k < 39) return k; if (k == 39) return -1; if (k < 88) return 39 + (k - 40); return -1; }
    if (kind == 2) { if (k < 39) return k; if (k == 39) return -1; if (k < 72) return 39 + (k - 40); if (k < 74) return 71 + (k - 72); return -1; }
    return k;
}


}  // namespace tc
}  // namespace mb
