// Host-glue kernels of the render-and-loss step: each replaces a swarm of tiny eager launches of the reference
// (and of our own first torch-level host code) by one launch forward and one backward.
//
//   mb_ray_points_*   xyz = o[ray] + d[ray] * (t0+t1)/2                     morpheus.py:645-646   (+ segmented-sum backward
//                     instead of torch's index_put(accumulate) kernels, 0.5 ms each at M = 524 288)
//   mb_pack_arena_*   nn.utils.weight_norm (W = v * g/||v||_row, models/decoders.py:51-52) of the 18 dense layers on the
//                     hot path, transposed / padded into the parameter arena of include/morpheus_b200.h; backward maps
//                     the flat gradient arena back onto weight_v / weight_g / weight / bias
//   mb_sdf_loss_*     utils.get_sdf_loss (utils.py:91-113) on packed samples, forward sums + per-sample gradient
#include "common.cuh"

namespace mb {

// ---------------------------------------------------------------------------------------------------------------
__global__ void ray_points_fwd_kernel(const float* __restrict__ o, const float* __restrict__ d, const int64_t* __restrict__ ri,
                                      const float* __restrict__ t0, const float* __restrict__ t1, uint32_t M, float* __restrict__ xyz) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int64_t r = ri[i];
    const float tm = __fdiv_rn(__fadd_rn(t0[i], t1[i]), 2.0f);
#pragma unroll
    for (int a = 0; a < 3; a++) xyz[(size_t)i * 3 + a] = __fadd_rn(__ldg(o + r * 3 + a), __fmul_rn(__ldg(d + r * 3 + a), tm));
}

// one warp per ray: g_o[r] = sum g_xyz, g_d[r] = sum g_xyz * t_mid over the ray's packed samples
__global__ void ray_points_bwd_kernel(const int32_t* __restrict__ seg, uint32_t N, const float* __restrict__ t0, const float* __restrict__ t1,
                                      const float* __restrict__ g_xyz, float* __restrict__ g_o, float* __restrict__ g_d) {
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= N) return;
    const int s0 = seg[r], s1 = seg[r + 1];
    float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = s0 + lane; i < s1; i += 32) {
        const float tm = __fdiv_rn(__fadd_rn(t0[i], t1[i]), 2.0f);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float g = g_xyz[(size_t)i * 3 + c];
            a[c] += g;
            a[3 + c] += g * tm;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int c = 0; c < 6; c++) a[c] += __shfl_xor_sync(0xffffffffu, a[c], o);
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (g_o) g_o[(size_t)r * 3 + c] = a[c];
            if (g_d) g_d[(size_t)r * 3 + c] = a[3 + c];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// layer table, 8 x int64 per layer: {v_ptr (weight_v or weight), g_ptr (weight_g or 0), b_ptr, K, N, K_pad, N_pad, wt_off}
// grad table, 4 x int64 per layer: {gv_off, gg_off (-1 if none), gb_off, 0} (float offsets into the flat gradient buffer)
__global__ void pack_arena_fwd_kernel(const int64_t* __restrict__ tab, float* __restrict__ arena) {
    const int64_t* t = tab + (size_t)blockIdx.y * 8;
    const float* v = reinterpret_cast<const float*>(t[0]);
    const float* g = reinterpret_cast<const float*>(t[1]);
    const float* b = reinterpret_cast<const float*>(t[2]);
    const int K = (int)t[3], N = (int)t[4], Kp = (int)t[5], Np = (int)t[6];
    const size_t wt_off = (size_t)t[7], w_off = wt_off + (size_t)Kp * Np, b_off = w_off + (size_t)Kp * Np;
    const int n = blockIdx.x;
    if (n >= Np) return;
    __shared__ float s_scale;
    if (threadIdx.x < 32) {
        float s = 1.0f;
        if (g && n < N) {
            double acc = 0.0;
            for (int k = threadIdx.x; k < K; k += 32) { const double x = v[(size_t)n * K + k]; acc += x * x; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            s = __fdiv_rn(g[n], (float)sqrt(acc));
        }
        if (threadIdx.x == 0) s_scale = s;
    }
    __syncthreads();
    const float s = s_scale;
    for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
        float w = 0.f;
        if (n < N && k < K) {
            w = v[(size_t)n * K + k];
            if (g) w = __fmul_rn(w, s);
        }
        arena[w_off + (size_t)n * Kp + k] = w;
        arena[wt_off + (size_t)k * Np + n] = w;
    }
    if (threadIdx.x == 0) arena[b_off + n] = (n < N) ? b[n] : 0.f;
}

__global__ void pack_arena_bwd_kernel(const int64_t* __restrict__ tab, const int64_t* __restrict__ gtab, const float* __restrict__ g_arena,
                                      float* __restrict__ flat) {
    const int64_t* t = tab + (size_t)blockIdx.y * 8;
    const int64_t* gt = gtab + (size_t)blockIdx.y * 4;
    const float* v = reinterpret_cast<const float*>(t[0]);
    const float* g = reinterpret_cast<const float*>(t[1]);
    const int K = (int)t[3], N = (int)t[4], Np = (int)t[6];
    const size_t wt_off = (size_t)t[7], b_off = wt_off + 2 * (size_t)t[5] * Np;
    const int n = blockIdx.x;
    if (n >= N) return;
    float* gv = flat + gt[0];
    __shared__ float s_dot, s_norm;
    if (g) {
        if (threadIdx.x < 32) {
            double acc = 0.0, dot = 0.0;
            for (int k = threadIdx.x; k < K; k += 32) {
                const double x = v[(size_t)n * K + k];
                acc += x * x;
                dot += x * (double)g_arena[wt_off + (size_t)k * Np + n];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, o); dot += __shfl_xor_sync(0xffffffffu, dot, o); }
            if (threadIdx.x == 0) { s_dot = (float)dot; s_norm = (float)sqrt(acc); }
        }
        __syncthreads();
        const float norm = s_norm, dot = s_dot, gn = g[n];
        const float s = gn / norm, c = gn * dot / (norm * norm * norm);
        for (int k = threadIdx.x; k < K; k += blockDim.x)
            gv[(size_t)n * K + k] = g_arena[wt_off + (size_t)k * Np + n] * s - c * v[(size_t)n * K + k];
        if (threadIdx.x == 0) flat[gt[1] + n] = dot / norm;
    } else {
        for (int k = threadIdx.x; k < K; k += blockDim.x) gv[(size_t)n * K + k] = g_arena[wt_off + (size_t)k * Np + n];
    }
    if (threadIdx.x == 0) flat[gt[2] + n] = g_arena[b_off + n];
}

// ---------------------------------------------------------------------------------------------------------------
// utils.get_sdf_loss (utils.py:91-113) on packed samples: z = (t0+t1)/2, d_gt = depth[ray], mask = mask[ray].
// The reference normalises per SAMPLE (sum(dim=-1) over a size-1 axis, utils.py:106-111): n_i = front_i + band_i + 1e-8.
// out[0] += sum_i fs_i / n_i ; out[1] += sum_i |s_i - bound_i| band_i / n_i  (caller divides by the number of rays with depth)
__device__ __forceinline__ void sdf_loss_terms(float z, float dgt, float msk, bool has_mask, float s, float trunc, float& fs, float& sl,
                                               float& dfs, float& dsl) {
    const bool depth_ok = dgt > 0.f;
    const bool front = (z < (dgt - trunc)) || ((dgt < 0.f) && (z < 3.5f));
    const float bound = (dgt < 0.f) ? 10.f : (dgt - z);
    bool band = (fabsf(bound) <= trunc) && depth_ok;
    if (has_mask) band = band && (msk > 0.5f);
    const float n = (front ? 1.f : 0.f) + (band ? 1.f : 0.f) + 1e-8f;
    fs = sl = dfs = dsl = 0.f;
    if (front) {
        const float e = expf(-5.f * s) - 1.f, lin = s - bound;
        const float mx = fmaxf(e, lin);
        if (mx >= 0.f) {      // clamp(min=0) passes the gradient at 0 (torch: self >= min)
            fs = mx / n;
            dfs = ((e > lin) ? -5.f * expf(-5.f * s) : 1.f) / n;      // torch.max backward: ties go to the first argument's... (measure-zero)
            if (e == lin) dfs = 0.5f * (-5.f * expf(-5.f * s) + 1.f) / n;
        }
    }
    if (band) {
        const float df = s - bound;
        sl = fabsf(df) / n;
        dsl = ((df > 0.f) ? 1.f : (df < 0.f ? -1.f : 0.f)) / n;
    }
}

__global__ void sdf_loss_fwd_kernel(const float* __restrict__ t0, const float* __restrict__ t1, const int64_t* __restrict__ ri,
                                    const float* __restrict__ depth, const float* __restrict__ mask, const float* __restrict__ sdf,
                                    uint32_t M, float trunc, float* __restrict__ out) {
    float a = 0.f, b = 0.f;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        const int64_t r = ri[i];
        float fs, sl, dfs, dsl;
        sdf_loss_terms(__fdiv_rn(__fadd_rn(t0[i], t1[i]), 2.0f), depth[r], mask ? mask[r] : 1.f, mask != nullptr, sdf[i], trunc, fs, sl, dfs, dsl);
        a += fs;
        b += sl;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    __shared__ float sa[32], sb[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sa[w] = a; sb[w] = b; }
    __syncthreads();
    if (w == 0) {
        a = lane < (int)(blockDim.x >> 5) ? sa[lane] : 0.f;
        b = lane < (int)(blockDim.x >> 5) ? sb[lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
        if (lane == 0) { atomicAdd(out, a); atomicAdd(out + 1, b); }
    }
}

// g_sdf[i] = g_out[0] * d(fs_i)/ds + g_out[1] * d(sl_i)/ds   (g_out: device scalars, already divided by the ray count)
__global__ void sdf_loss_bwd_kernel(const float* __restrict__ t0, const float* __restrict__ t1, const int64_t* __restrict__ ri,
                                    const float* __restrict__ depth, const float* __restrict__ mask, const float* __restrict__ sdf,
                                    uint32_t M, float trunc, const float* __restrict__ g_out, float* __restrict__ g_sdf) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const int64_t r = ri[i];
    float fs, sl, dfs, dsl;
    sdf_loss_terms(__fdiv_rn(__fadd_rn(t0[i], t1[i]), 2.0f), depth[r], mask ? mask[r] : 1.f, mask != nullptr, sdf[i], trunc, fs, sl, dfs, dsl);
    g_sdf[i] = g_out[0] * dfs + g_out[1] * dsl;
}

}  // namespace mb

extern "C" int mb_ray_points_forward(const float* rays_o, const float* rays_d, const int64_t* ray_indices, const float* t_starts,
                                     const float* t_ends, uint32_t M, float* xyz, mb_stream_t stream) {
    using namespace mb;
    if (M == 0) return MB_OK;
    if (!rays_o || !rays_d || !ray_indices || !t_starts || !t_ends || !xyz) { set_error("ray_points_forward: null argument"); return MB_EINVAL; }
    ray_points_fwd_kernel<<<div_up(M, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, ray_indices, t_starts, t_ends, M, xyz);
    return check_launch("ray_points_forward");
}

extern "C" int mb_ray_points_backward(const int32_t* seg, uint32_t N, const float* t_starts, const float* t_ends, const float* g_xyz,
                                      float* g_o, float* g_d, mb_stream_t stream) {
    using namespace mb;
    if (N == 0) return MB_OK;
    if (!seg || !t_starts || !t_ends || !g_xyz) { set_error("ray_points_backward: null argument"); return MB_EINVAL; }
    ray_points_bwd_kernel<<<div_up(N * 32, 256), 256, 0, (cudaStream_t)stream>>>(seg, N, t_starts, t_ends, g_xyz, g_o, g_d);
    return check_launch("ray_points_backward");
}

extern "C" int mb_pack_arena_forward(const int64_t* layer_table, int n_layers, float* arena, mb_stream_t stream) {
    using namespace mb;
    if (!layer_table || !arena || n_layers <= 0) { set_error("pack_arena_forward: bad argument"); return MB_EINVAL; }
    pack_arena_fwd_kernel<<<dim3(128, n_layers), 128, 0, (cudaStream_t)stream>>>(layer_table, arena);
    return check_launch("pack_arena_forward");
}

extern "C" int mb_pack_arena_backward(const int64_t* layer_table, const int64_t* grad_table, int n_layers, const float* g_arena, float* flat_grads,
                                      mb_stream_t stream) {
    using namespace mb;
    if (!layer_table || !grad_table || !g_arena || !flat_grads || n_layers <= 0) { set_error("pack_arena_backward: bad argument"); return MB_EINVAL; }
    pack_arena_bwd_kernel<<<dim3(128, n_layers), 128, 0, (cudaStream_t)stream>>>(layer_table, grad_table, g_arena, flat_grads);
    return check_launch("pack_arena_backward");
}

extern "C" int mb_sdf_loss_forward(const float* t_starts, const float* t_ends, const int64_t* ray_indices, const float* depth, const float* mask,
                                   const float* sdf, uint32_t M, float truncation, float* out2, mb_stream_t stream) {
    using namespace mb;
    if (!t_starts || !t_ends || !ray_indices || !depth || !sdf || !out2) { set_error("sdf_loss_forward: null argument"); return MB_EINVAL; }
    if (M == 0) return MB_OK;
    const uint32_t grid = min(div_up(M, 256), (uint32_t)mb_sm_count() * 4u);
    sdf_loss_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(t_starts, t_ends, ray_indices, depth, mask, sdf, M, truncation, out2);
    return check_launch("sdf_loss_forward");
}

extern "C" int mb_sdf_loss_backward(const float* t_starts, const float* t_ends, const int64_t* ray_indices, const float* depth, const float* mask,
                                    const float* sdf, uint32_t M, float truncation, const float* g_out2, float* g_sdf, mb_stream_t stream) {
    using namespace mb;
    if (!t_starts || !t_ends || !ray_indices || !depth || !sdf || !g_out2 || !g_sdf) { set_error("sdf_loss_backward: null argument"); return MB_EINVAL; }
    if (M == 0) return MB_OK;
    sdf_loss_bwd_kernel<<<div_up(M, 256), 256, 0, (cudaStream_t)stream>>>(t_starts, t_ends, ray_indices, depth, mask, sdf, M, truncation, g_out2, g_sdf);
    return check_launch("sdf_loss_backward");
}
