// Stand-alone multi-resolution hash-grid encoder (forward + dy_dx, grad scatter, input grad).
// Drop-in for the reference extension entry points grid_encode_forward / grid_encode_backward
// (external/encoders/gridencoder/src/gridencoder.h:12-13).  Arithmetic follows the reference
// kernel operation by operation (gridencoder.cu:83-249) with the FMA contractions nvcc applies to
// it written out explicitly, so results are bit-identical to the reference kernel on the same GPU
// (tests/test_gpu_parity.py::test_grid_encode_bit_exact_vs_reference_kernel checks this against oracle/_ref).
//
// B200 design: compulsory HBM traffic is 12 B in + 128 B out + 384 B dy_dx per sample (tables are L2-resident), but for
// incoherent points the binding resource is the L2 -> SM sector traffic of the 128 eight-byte gathers per sample (tools/time_grid.py).  One CTA = 128 consecutive samples x all levels.  outputs[l, b0:b0+128, :] is a
// contiguous 1 KB run per level -> coalesced float2 stores straight from registers; the dy_dx
// tile [128, L*D*C] is one contiguous global block, staged in shared memory and written back
// with 16-byte coalesced stores (the reference scatters 24-byte pieces at a 384-byte stride).
#include "common.cuh"

namespace mb {

constexpr int GE_TILE = 128;     // samples per CTA
constexpr int GE_THREADS = 256;  // 2 level-lanes x 128 samples

template <uint32_t D, uint32_t C>
__global__ void __launch_bounds__(GE_THREADS) grid_fwd_kernel(const float* __restrict__ inputs, const float* __restrict__ grid,
                                                              const int* __restrict__ offsets, float* __restrict__ outputs,
                                                              uint32_t B, uint32_t L, uint32_t max_level, float S, uint32_t H,
                                                              float* __restrict__ dy_dx, uint32_t gridtype, bool align_corners,
                                                              uint32_t interp) {
    extern __shared__ float s_dydx[];  // [GE_TILE][L*D*C + 1] when dy_dx != nullptr (+1: bank-conflict-free row pitch)
    const uint32_t b0 = blockIdx.x * GE_TILE;
    const uint32_t m = threadIdx.x % GE_TILE;
    const uint32_t lane_l = threadIdx.x / GE_TILE;  // 0..1
    const uint32_t b = b0 + m;
    const uint32_t row = L * D * C;
    const bool valid = b < B;

    float x[D];
    bool oob = false;
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        x[d] = valid ? inputs[b * D + d] : 0.0f;
        if (x[d] < 0 || x[d] > 1) oob = true;
    }

    for (uint32_t level = lane_l; level < L; level += GE_THREADS / GE_TILE) {
        float res_out[C];
        float res_grad[D][C];
#pragma unroll
        for (uint32_t c = 0; c < C; c++) {
            res_out[c] = 0.f;
#pragma unroll
            for (uint32_t d = 0; d < D; d++) res_grad[d][c] = 0.f;
        }
        const bool active = level < max_level;
        if (active && valid && !oob) {
            const float* tab = grid + (size_t)(uint32_t)offsets[level] * C;
            const uint32_t hashmap_size = offsets[level + 1] - offsets[level];
            const uint32_t res = level_resolution(level, S, H);
            float pos[D], deriv[D];
            uint32_t pg[D];
#pragma unroll
            for (uint32_t d = 0; d < D; d++) pos[d] = locate(x[d], res, align_corners, interp, pg[d], deriv[d]);
            // the 2^D corner values are fetched ONCE (one 8-byte load per corner for C = 2) and feed both the interpolation and the
            // corner differences of dy_dx (the reference re-reads them: 2^D + D 2^D loads per level).  Hashed levels (tables far larger
            // than L1) bypass L1 allocation so that the small dense tables of the coarse levels stay resident there.
            float cval[1u << D][C];
            const bool stream_level = hashmap_size > 16384u;
#pragma unroll
            for (uint32_t idx = 0; idx < (1u << D); idx++) {
                uint32_t pl[D];
#pragma unroll
                for (uint32_t d = 0; d < D; d++) pl[d] = (idx & (1u << d)) ? min(pg[d] + 1, res - 1) : pg[d];
                const uint32_t index = grid_index<D>(gridtype, hashmap_size, res, pl) * C;
                if (C == 2) {
                    float2 v;
                    if (stream_level) asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(tab + index));
                    else v = __ldg(reinterpret_cast<const float2*>(tab + index));
                    cval[idx][0] = v.x;
                    cval[idx][1 % C] = v.y;
                } else {
#pragma unroll
                    for (uint32_t c = 0; c < C; c++) cval[idx][c] = __ldg(tab + index + c);
                }
            }
#pragma unroll
            for (uint32_t idx = 0; idx < (1u << D); idx++) {
                float w = 1.0f;
#pragma unroll
                for (uint32_t d = 0; d < D; d++) w = __fmul_rn(w, (idx & (1u << d)) ? pos[d] : __fsub_rn(1.0f, pos[d]));
#pragma unroll
                for (uint32_t c = 0; c < C; c++) res_out[c] = __fmaf_rn(w, cval[idx][c], res_out[c]);
            }
            if (dy_dx) {
                const float scale = (float)(align_corners ? res - 1 : res);
#pragma unroll
                for (uint32_t gd = 0; gd < D; gd++) {
#pragma unroll
                    for (uint32_t idx = 0; idx < (1u << (D - 1)); idx++) {
                        float w = scale;
                        uint32_t cl = 0;
#pragma unroll
                        for (uint32_t nd = 0; nd < D - 1; nd++) {
                            const uint32_t d = (nd >= gd) ? (nd + 1) : nd;
                            if ((idx & (1u << nd)) == 0) {
                                w = __fmul_rn(w, __fsub_rn(1.0f, pos[d]));
                            } else {
                                w = __fmul_rn(w, pos[d]);
                                cl |= (1u << d);
                            }
                        }
#pragma unroll
                        for (uint32_t c = 0; c < C; c++) {
                            const float diff = __fsub_rn(cval[cl | (1u << gd)][c], cval[cl][c]);
                            res_grad[gd][c] = __fmaf_rn(__fmul_rn(w, diff), deriv[gd], res_grad[gd][c]);
                        }
                    }
                }
            }
        }
        // outputs[l, b, :]  (levels >= max_level are left to the caller's zero fill, grid.py:53)
        if (active && valid) {
            float* o = outputs + ((size_t)level * B + b) * C;
            if (C == 2) {
                *reinterpret_cast<float2*>(o) = make_float2(res_out[0], res_out[1 % C]);
            } else {
#pragma unroll
                for (uint32_t c = 0; c < C; c++) o[c] = res_out[c];
            }
        }
        if (dy_dx) {
            // inactive levels are written as zeros here (the reference relies on dy_dx.zero_(), grid.py:57)
            float* sd = s_dydx + (size_t)m * (row + 1) + level * D * C;
#pragma unroll
            for (uint32_t d = 0; d < D; d++)
#pragma unroll
                for (uint32_t c = 0; c < C; c++) sd[d * C + c] = res_grad[d][c];
        }
    }
    if (dy_dx) {
        __syncthreads();
        const uint32_t n_valid = min((uint32_t)GE_TILE, B - b0);
        const size_t total = (size_t)n_valid * row;  // floats, contiguous in global
        float* g = dy_dx + (size_t)b0 * row;
        if ((row % 4) == 0) {
            float4* g4 = reinterpret_cast<float4*>(g);
            for (uint32_t i = threadIdx.x; i < total / 4; i += GE_THREADS) {
                const uint32_t r = (i * 4) / row, j = (i * 4) % row;
                const float* sp = s_dydx + r * (row + 1) + j;
                g4[i] = make_float4(sp[0], sp[1], sp[2], sp[3]);
            }
        } else {
            for (uint32_t i = threadIdx.x; i < total; i += GE_THREADS) g[i] = s_dydx[(i / row) * (row + 1) + i % row];
        }
    }
}

// grad scatter: one thread = (sample, level); both channels of an entry go out as one red.v2
template <uint32_t D, uint32_t C>
__global__ void __launch_bounds__(GE_THREADS) grid_bwd_kernel(const float* __restrict__ grad, const float* __restrict__ inputs,
                                                              const int* __restrict__ offsets, float* __restrict__ grad_grid,
                                                              uint32_t B, uint32_t L, uint32_t max_level, float S, uint32_t H,
                                                              uint32_t gridtype, bool align_corners, uint32_t interp) {
    const uint32_t b = blockIdx.x * GE_TILE + threadIdx.x % GE_TILE;
    if (b >= B) return;
    float x[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        x[d] = inputs[b * D + d];
        if (x[d] < 0 || x[d] > 1) return;  // gridencoder.cu:279-284
    }
    for (uint32_t level = threadIdx.x / GE_TILE; level < max_level; level += GE_THREADS / GE_TILE) {
        float* gg = grad_grid + (size_t)(uint32_t)offsets[level] * C;
        const uint32_t hashmap_size = offsets[level + 1] - offsets[level];
        const uint32_t res = level_resolution(level, S, H);
        float pos[D], deriv;
        uint32_t pg[D];
#pragma unroll
        for (uint32_t d = 0; d < D; d++) pos[d] = locate(x[d], res, align_corners, interp, pg[d], deriv);
        float gcur[C];
#pragma unroll
        for (uint32_t c = 0; c < C; c++) gcur[c] = grad[((size_t)level * B + b) * C + c];
#pragma unroll
        for (uint32_t idx = 0; idx < (1u << D); idx++) {
            float w = 1.0f;
            uint32_t pl[D];
#pragma unroll
            for (uint32_t d = 0; d < D; d++) {
                if ((idx & (1u << d)) == 0) {
                    w = __fmul_rn(w, __fsub_rn(1.0f, pos[d]));
                    pl[d] = pg[d];
                } else {
                    w = __fmul_rn(w, pos[d]);
                    pl[d] = min(pg[d] + 1, res - 1);
                }
            }
            const uint32_t index = grid_index<D>(gridtype, hashmap_size, res, pl) * C;
            if (C == 2) {
                red_add2(gg + index, __fmul_rn(w, gcur[0]), __fmul_rn(w, gcur[1 % C]));
            } else {
#pragma unroll
                for (uint32_t c = 0; c < C; c++) red_add(gg + index + c, __fmul_rn(w, gcur[c]));
            }
        }
    }
}

// grad_inputs[b,d] = sum_{l,c} grad[l,b,c] * dy_dx[b,l,d,c]   (gridencoder.cu:353-378, same fma order)
template <uint32_t D, uint32_t C>
__global__ void grid_input_bwd_kernel(const float* __restrict__ grad, const float* __restrict__ dy_dx,
                                      float* __restrict__ grad_inputs, uint32_t B, uint32_t L) {
    const uint32_t t = threadIdx.x + blockIdx.x * blockDim.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const float* dd = dy_dx + (size_t)b * L * D * C;
    float result = 0;
    for (uint32_t l = 0; l < L; l++)
#pragma unroll
        for (uint32_t c = 0; c < C; c++)
            result = __fmaf_rn(grad[((size_t)l * B + b) * C + c], dd[l * D * C + d * C + c], result);
    grad_inputs[t] = result;
}

template <uint32_t D, uint32_t C>
static int launch_fwd(const float* inputs, const float* emb, const int* offsets, float* outputs, uint32_t B, uint32_t L,
                      uint32_t max_level, float S, uint32_t H, float* dy_dx, uint32_t gridtype, bool ac, uint32_t interp,
                      cudaStream_t st) {
    const size_t smem = dy_dx ? (size_t)GE_TILE * (L * D * C + 1) * sizeof(float) : 0;
    if (smem > 227 * 1024) { set_error("grid_encode_forward: L*D*C=%u too large for the dy_dx staging tile", L * D * C); return MB_EUNSUPPORTED; }
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(grid_fwd_kernel<D, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("grid_encode_forward: smem %zu: %s", smem, cudaGetErrorString(e)); return MB_ECUDA; }
    }
    grid_fwd_kernel<D, C><<<div_up(B, GE_TILE), GE_THREADS, smem, st>>>(inputs, emb, offsets, outputs, B, L, max_level, S, H,
                                                                        dy_dx, gridtype, ac, interp);
    return check_launch("grid_encode_forward");
}

template <uint32_t D, uint32_t C>
static int launch_bwd(const float* grad, const float* inputs, const int* offsets, float* gemb, uint32_t B, uint32_t L,
                      uint32_t max_level, float S, uint32_t H, const float* dy_dx, float* ginp, uint32_t gridtype, bool ac,
                      uint32_t interp, cudaStream_t st) {
    grid_bwd_kernel<D, C><<<div_up(B, GE_TILE), GE_THREADS, 0, st>>>(grad, inputs, offsets, gemb, B, L, max_level, S, H,
                                                                     gridtype, ac, interp);
    int rc = check_launch("grid_encode_backward");
    if (rc) return rc;
    if (dy_dx && ginp) {
        grid_input_bwd_kernel<D, C><<<div_up(B * D, 256), 256, 0, st>>>(grad, dy_dx, ginp, B, L);
        rc = check_launch("grid_encode_backward(inputs)");
    }
    return rc;
}

}  // namespace mb

#define MB_DISPATCH_DC(D, C, CALL)                                             \
    if (D == 3 && C == 2) { constexpr uint32_t D_ = 3, C_ = 2; CALL; }         \
    else if (D == 3 && C == 1) { constexpr uint32_t D_ = 3, C_ = 1; CALL; }    \
    else if (D == 3 && C == 4) { constexpr uint32_t D_ = 3, C_ = 4; CALL; }    \
    else if (D == 3 && C == 8) { constexpr uint32_t D_ = 3, C_ = 8; CALL; }    \
    else if (D == 2 && C == 2) { constexpr uint32_t D_ = 2, C_ = 2; CALL; }    \
    else if (D == 2 && C == 1) { constexpr uint32_t D_ = 2, C_ = 1; CALL; }    \
    else if (D == 2 && C == 4) { constexpr uint32_t D_ = 2, C_ = 4; CALL; }    \
    else if (D == 2 && C == 8) { constexpr uint32_t D_ = 2, C_ = 8; CALL; }    \
    else { mb::set_error("GridEncoding: unsupported D=%u C=%u (supported D in {2,3}, C in {1,2,4,8})", D, C); return MB_EUNSUPPORTED; }

extern "C" int mb_grid_encode_forward(const float* inputs, const void* embeddings, const int32_t* offsets, void* outputs,
                                      uint32_t B, uint32_t D, uint32_t C, uint32_t L, uint32_t max_level, float S, uint32_t H,
                                      void* dy_dx, uint32_t gridtype, int align_corners, uint32_t interp, int dtype,
                                      mb_stream_t stream) {
    if (!inputs || !embeddings || !offsets || !outputs) { mb::set_error("grid_encode_forward: null pointer"); return MB_EINVAL; }
    if (dtype != MB_DTYPE_F32) { mb::set_error("grid_encode_forward: only float32 embeddings are supported"); return MB_EUNSUPPORTED; }
    if (B == 0) return MB_OK;
    if (max_level > L) max_level = L;
    int rc = MB_OK;
    MB_DISPATCH_DC(D, C, rc = (mb::launch_fwd<D_, C_>(inputs, (const float*)embeddings, offsets, (float*)outputs, B, L, max_level, S,
                                                      H, (float*)dy_dx, gridtype, align_corners != 0, interp, (cudaStream_t)stream)))
    return rc;
}

extern "C" int mb_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings, const int32_t* offsets,
                                       void* grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, uint32_t max_level,
                                       float S, uint32_t H, const void* dy_dx, void* grad_inputs, uint32_t gridtype,
                                       int align_corners, uint32_t interp, int dtype, mb_stream_t stream) {
    (void)embeddings;
    if (!grad || !inputs || !offsets || !grad_embeddings) { mb::set_error("grid_encode_backward: null pointer"); return MB_EINVAL; }
    if (dtype != MB_DTYPE_F32) { mb::set_error("grid_encode_backward: only float32 is supported"); return MB_EUNSUPPORTED; }
    if (B == 0) return MB_OK;
    if (max_level > L) max_level = L;
    int rc = MB_OK;
    MB_DISPATCH_DC(D, C, rc = (mb::launch_bwd<D_, C_>((const float*)grad, inputs, offsets, (float*)grad_embeddings, B, L, max_level, S,
                                                      H, (const float*)dy_dx, (float*)grad_inputs, gridtype, align_corners != 0,
                                                      interp, (cudaStream_t)stream)))
    return rc;
}
